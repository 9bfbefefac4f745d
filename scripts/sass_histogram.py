"""Opcode histogram per kernel from `cuobjdump -sass` on stdin: the evidence that the executor is tcgen05 / TMA / TMEM
code (UTCHMMA, UTMALDG, LDTM, UTCBAR ...) and that the fan kernel keeps its state in registers (profiles/README.md)."""
import collections
import re
import sys

kernels = collections.OrderedDict()
cur = None
for line in sys.stdin:
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)', line)
    if m and cur is not None:
        cur[m.group(1)] += 1
want = re.compile(sys.argv[1] if len(sys.argv) > 1 else r'gemm_tc_kernel|fan_kernel')
for name, hist in kernels.items():
    if not want.search(name):
        continue
    print('== %s: %d instructions' % (name, sum(hist.values())))
    base = collections.Counter()
    for op, n in hist.items():
        base[op.split('.')[0]] += n
    print('   by opcode: ' + ', '.join('%s %d' % kv for kv in base.most_common(24)))
    special = {op: n for op, n in hist.items() if re.match(r'UTC|UTMA|LDTM|STTM|UBLKCP|SYNCS|ELECT|HMMA|LDS\.128|LDS\.64|STS\.128|LDG\.E\.128|STG\.E\.128|BAR', op)}
    print('   of note: ' + ', '.join('%s %d' % kv for kv in sorted(special.items())))
