"""From an ncu launch list of a bench run (scripts/gpu_visit.sh launches), print `name regex skip` lines for one launch of
every kernel kind of the LGD step, counted in the third pass over the model (a warm-up step): what `ncu -k regex:<regex>
-s <skip> -c 1` needs to capture exactly that launch in a second run of the same command line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki = hdr.index('Kernel Name')
names = [r[ki] for r in rows[start:] if len(r) > ki]
steps = [i for i, n in enumerate(names) if 'pack_offsets' in n]
s0 = steps[2]
gemm_before = sum('gemm_tc_kernel' in n for n in names[:s0])
fan_before = sum('fan_kernel' in n for n in names[:s0])
for k, name in enumerate(('lstm', 'heads', 'blend_fwd', 'blend_t', 'chain')):
    print(name, 'gemm_tc', gemm_before + k)
print('fan_grad', 'fan_kernel', fan_before)
print('fan_fwd', 'fan_kernel', fan_before + 4)
for name in ('update_kernel', 'post_kernel', 'prepare_kernel'):
    print(name.replace('_kernel', ''), name, sum(name in n for n in names[:s0]) + (1 if name == 'update_kernel' else 0))
