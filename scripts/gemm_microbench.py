"""Throughput experiments on the tcgen05 job executor (run on the GPU box).
EMPOSE_TC_DEBUG=0 normal, 1 TMA loads only, 2 MMAs only."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from empose_b200 import lib

dev = torch.device('cuda:0')
res = {}
PREC = {'tf32': lib.PRECISION_TF32, 'fp16': lib.PRECISION_FP16}[os.environ.get('EMPOSE_BENCH_PRECISION', 'fp16')]
shapes = [(128, 256, 64), (128 * 148, 256, 64), (128 * 148, 256, 1024), (4096, 2048, 1024), (4096, 2048, 704), (4096, 4096, 1024), (131072, 512, 512), (131072, 256, 576), (4096, 2048, 672), (131072, 512, 2048), (131072, 256, 512), (131072, 128, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split('x')) for a in sys.argv[1:]]
for (m, n, k) in shapes:
    a = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev)
    b = torch.zeros(n, device=dev)
    ms = lib.gemm_bench(a, w, b, PREC, reps=20)
    res['%dx%dx%d' % (m, n, k)] = {'ms': round(ms, 4), 'tflops': round(2.0 * m * n * k / ms / 1e9, 1),
                                  'load_TBps': round((m / 128) * (n / min(n, 256)) * (k / 32) * (128 + min(n, 256)) * 128 / ms / 1e9, 2)}
print(json.dumps({'mode': os.environ.get('EMPOSE_TC_DEBUG', '0'), 'precision': os.environ.get('EMPOSE_BENCH_PRECISION', 'fp16'), 'results': res}, indent=1))
