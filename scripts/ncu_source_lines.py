"""Aggregate an .ncu-rep (captured with --import-source on, built with -lineinfo) per source line:
samples, executed warp instructions and the top stall reasons.  Usage: ncu_source_lines.py <rep> [top N]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
fname = None
hdr = None
agg = collections.OrderedDict()
tot_s = tot_i = 0
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == 'File Path':
        fname = row[1].split('/')[-1]
        continue
    if row[0] == 'Function Name':
        continue
    if row[0] == 'Line No':
        hdr = row
        ci = {h: i for i, h in enumerate(hdr)}
        stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or row[2] != '-':          # keep the per-source-line summary rows (Address == '-')
        continue
    try:
        s = int(row[ci['# Samples']]); n = int(row[ci['Instructions Executed']])
    except ValueError:
        continue
    key = (fname, int(row[0]))
    a = agg.setdefault(key, {'src': row[1].strip(), 's': 0, 'i': 0, 'st': collections.Counter()})
    a['s'] += s; a['i'] += n
    for h, i in stall_cols:
        try:
            a['st'][h[6:]] += int(row[i])
        except ValueError:
            pass
    tot_s += s; tot_i += n
print('total samples %d, warp instructions %d' % (tot_s, tot_i))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]['s'])[:top]:
    st = ', '.join('%s %d' % kv for kv in a['st'].most_common(3))
    print('%5.1f%% smp %5.1f%% ins  %s:%d  %s   [%s]' % (100.0 * a['s'] / max(tot_s, 1), 100.0 * a['i'] / max(tot_i, 1), f, ln, a['src'][:90], st))
