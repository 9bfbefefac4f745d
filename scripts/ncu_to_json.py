"""Extract the numbers bench.py quotes from committed ncu captures (ncu --set full, one launch each).

    python scripts/ncu_to_json.py profiles/ncu_kernels.json chain=gpurun_out/x_prof_chain.ncu-rep lstm=... fan_grad=...

Writes {name: {kernel, duration_us, dram_read_bytes, dram_write_bytes, warp_instructions, tensor_active_pct,
issue_active_pct, l1_pct, registers}}; bench.py derives roofline.traffic (launch-weighted DRAM bytes of the executor)
and the fan kernel's instruction count from this file instead of hard-coded constants."""
import csv
import json
import subprocess
import sys

WANT = {'gpu__time_duration.sum': 'duration', 'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
        'smsp__inst_executed.sum': 'warp_instructions',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed': 'tensor_active_pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1_pct', 'launch__registers_per_thread': 'registers',
        'launch__grid_size': 'grid', 'launch__block_size': 'block'}
UNIT = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def read(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {'kernel': vals[hdr.index('Kernel Name')][:80], 'source': path.split('/')[-1]}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            x = float(v.replace(',', ''))
            key = WANT[h]
            if key == 'duration':
                rec['duration_us'] = x * UNIT.get(u, 1.0)
            elif key.startswith('dram_'):
                rec[key + '_bytes'] = x * UNIT.get(u, 1.0)
            else:
                rec[key] = x
    return rec


if __name__ == '__main__':
    dst = sys.argv[1]
    try:
        data = json.load(open(dst))
    except (OSError, ValueError):
        data = {}
    for spec in sys.argv[2:]:
        name, path = spec.split('=', 1)
        data[name] = read(path)
    with open(dst, 'w') as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))
