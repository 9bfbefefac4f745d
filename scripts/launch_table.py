"""Per-kernel table of ONE bench step from an ncu launch list (gpu__time_duration.sum CSV): python scripts/launch_table.py <csv> [-v]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
recs = [(r[ki], float(r[vi].replace(',', '')) * (1e-3 if r[ui] == 'ns' else 1.0)) for r in rows[start:] if len(r) > vi]
idx = [i for i, (n, _) in enumerate(recs) if 'prepare_kernel' in n or 'pack_offsets' in n]
# a step starts at its first pack_offsets / prepare launch
starts = [i for k, i in enumerate(idx) if k == 0 or idx[k - 1] != i - 1]
step = recs[starts[-2]:starts[-1]]
agg = collections.OrderedDict()
for n, v in step:
    short = n.split('(')[0].split('::')[-1][:50]
    d = agg.setdefault(short, [0, 0.0])
    d[0] += 1
    d[1] += v
tot = sum(v for _, v in step)
print('one step: %d launches, %.1f us (serialised under ncu)' % (len(step), tot))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('  %-50s %3d  %9.1f us  %5.1f%%' % (k, c, v, 100 * v / tot))
if '-v' in sys.argv:
    for n, v in step:
        print('    %-40s %8.1f' % (n.split('(')[0].split('::')[-1][:40], v))
