"""Summarise an executor timeline trace (EMPOSE_TC_TRACE=<file>, csrc/gemm_tc.cu): per job of one CTA, in SM cycles,
what each role waited for and how long it worked.

    producer: [job start, deps satisfied, last load issued]
    issuer:   [before tmem_empty wait, after it, first operands arrived, accumulator committed]
    epilogue: [before tmem_full wait, after it, chunks done, job barrier passed]      (warp 4 of the CTA)
"""
import sys
from collections import defaultdict

rows = defaultdict(dict)
for line in open(sys.argv[1]):
    c, r, j, *t = line.split()
    rows[int(c)][(int(r), int(j))] = [int(x) for x in t]
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = rows[cta]
t0 = min(v[0] for v in d.values() if v[0])
jobs = sorted({j for (_, j) in d})
print('cta %d, cycles relative to the first stamp' % cta)
print('%4s | %-28s | %-40s | %-40s' % ('job', 'producer start/deps/issued', 'issuer wait_empty/start/first_ops/commit', 'epilogue wait_full/start/chunks/done'))
prev_commit = prev_done = None
per = []
for j in jobs:
    p = d.get((0, j), [0] * 4); m = d.get((1, j), [0] * 4); e = d.get((2, j), [0] * 4)
    f = lambda v: '/'.join('%7d' % (x - t0) if x else '      -' for x in v)
    print('%4d | %s | %s | %s' % (j, f(p[:3]), f(m), f(e)))
    s4 = d.get((3, j))
    if s4:
        print('     |   epilogue job start: before job_full %d, +%d job slot ready, +%d fields / dependencies, +%d state requested' % (s4[0] - t0, s4[1] - s4[0], s4[2] - s4[1], s4[3] - s4[2]))
    if m[3] and e[3]:
        per.append((j, m[1] - m[0], m[3] - m[1], e[1] - e[0], e[2] - e[1], e[3] - e[2]))
if per:
    n = len(per)
    mid = per[n // 4: n - n // 4] or per
    avg = lambda k: sum(x[k] for x in mid) / len(mid)
    print('middle jobs: issuer waits for a free accumulator %.0f, issues + commits in %.0f; epilogue waits for the accumulator %.0f, '
          'works %.0f, barrier %.0f cycles' % (avg(1), avg(2), avg(3), avg(4), avg(5)))
    e_first = d.get((2, mid[0][0]))[3]; e_last = d.get((2, mid[-1][0]))[3]
    if len(mid) > 1: print('period per job: %.0f cycles' % ((e_last - e_first) / (len(mid) - 1)))
