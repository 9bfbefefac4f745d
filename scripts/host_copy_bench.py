"""Host <-> device copy bandwidth of one box under torchrun: first rank 0 alone, then all ranks at once, with the
byte counts of one bench step (78 MB up, 74 MB down, concurrently on two streams).  Explains what bounds the end-to-end
figure at N GPUs when the device-resident figure scales (profiles/r02/README.md)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
up_bytes, down_bytes = 77873152, 74448896
h_in = torch.empty(up_bytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(down_bytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(up_bytes, dtype=torch.uint8, device=dev)
d_out = torch.empty(down_bytes, dtype=torch.uint8, device=dev)
s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def loop(n):
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(n):
        with torch.cuda.stream(s_up):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_down):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / n * 1e3


def barrier():
    if world > 1:
        dist.barrier()


loop(3)
barrier()
alone = loop(20) if rank == 0 else None
barrier()
together = loop(20)
if world > 1:
    t = torch.tensor([together], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    together = float(t.item())
if rank == 0:
    print(json.dumps({'n_gpus': world, 'cpus': len(os.sched_getaffinity(0)),
                      'rank0_alone_ms': alone, 'rank0_alone_up_GBps': up_bytes / alone / 1e6, 'rank0_alone_down_GBps': down_bytes / alone / 1e6,
                      'all_ranks_ms_max': together, 'per_gpu_up_GBps': up_bytes / together / 1e6, 'per_gpu_down_GBps': down_bytes / together / 1e6,
                      'aggregate_GBps_both_directions': world * (up_bytes + down_bytes) / together / 1e6}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
