"""The data-parallel training step on TWO B200s over NCCL (BASELINE config 5): what
``IterativeErrorFeedback.allreduce_gradients`` leaves in the flat gradient vector of the PRODUCT path -- CUDA tensors,
native forward / backward -- is the mean of the two ranks' local gradients (DDP semantics), and the two-bucket
form that reduces the dense bucket under the LSTM's backward-through-time sweep (``overlap_gradient_allreduce``) returns
the same bits and the same loss values.  Skipped on a box with fewer than two GPUs:

    gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu -q
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class _Batch(object):
    def __init__(self, inp, poses, shapes, joints):
        self.inp, self.seq_lengths = inp, inp['seq_lengths']
        self.poses_root, self.poses_body = poses[:, :, :3].contiguous(), poses[:, :, 3:].contiguous()
        self.shapes, self.joints_gt, self.marker_masks = shapes, joints, None
        self.batch_size, self.seq_length = poses.shape[0], poses.shape[1]

    def get_inputs(self, sf=None, ef=None, **kwargs):
        i = self.inp
        return {'marker_pos': i['marker_pos'], 'marker_oris': i['marker_oris'], 'offset_r': i['offset_r'], 'offset_t': i['offset_t'],
                'marker_masks': None}


def _worker(rank, world, port, npz, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    import util
    from empose_b200 import lib, synthetic
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    b, f = 40, 8
    net = util.build_module(npz, precision=lib.PRECISION_TF32, device=dev, m_fk_loss=0.1, m_pose_loss_weight=10.0)
    ctx = net.native_context(dev)
    p = synthetic.synth_window_params(b, f, seed=500 + rank, ragged=True, offsets=True)       # every rank its own shard
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    r = b * f
    poses, shapes = t(p['poses']), t(p['shapes'])
    rep = lambda x, *tail: x.unsqueeze(1).repeat(1, f, *([1] * len(tail))).reshape(r, *tail)
    pos, ori, joints = ctx.sensor_project(poses.reshape(r, 66), rep(shapes, 10), rep(t(p['offset_r']), 12, 3, 3), rep(t(p['offset_t']), 12, 3))
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    inp = dict(marker_pos=(pos + 0.01 * torch.randn(pos.shape, device=dev, generator=g)).reshape(b, f, 36), marker_oris=ori.reshape(b, f, 108),
               offset_r=t(p['offset_r']), offset_t=t(p['offset_t']), seq_lengths=t(p['seq_lengths']).to(torch.int32))
    batch = _Batch(inp, poses, shapes, joints.reshape(b, f, 66))
    net.train()
    results = {}
    net(batch)                                            # first call re-homes the parameters into the flat vectors
    for mode in ('plain', 'overlap'):                     # no optimiser step in between: both modes see the same parameters
        net.overlap_gradient_allreduce(mode == 'overlap', average=True)
        net.flat_gradients().zero_()
        out = net(batch)
        _, vals = net.backward(batch, out)
        if mode == 'plain':
            local = net.flat_gradients().clone()
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            want = torch.stack(gathered).double().mean(dim=0)
        net.allreduce_gradients(average=True)
        torch.cuda.synchronize(dev)
        results[mode] = (net.flat_gradients().clone(), vals)
    if rank == 0:
        got_plain, got_overlap = results['plain'][0], results['overlap'][0]
        scale = float(want.abs().max())
        np.save(os.path.join(out_dir, 'multirank.npy'), np.array([
            float((got_plain.double() - want).abs().max()) / scale,                 # all-reduced == mean of the local gradients
            float((got_overlap - got_plain).abs().max()),                           # two buckets == one all-reduce, bit for bit
            abs(results['plain'][1]['total_loss'] - results['overlap'][1]['total_loss']),
            float(net.lstm_bucket_end()), float(got_plain.numel()), scale]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_gradient_allreduce_on_two_gpus(smpl_npz, tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), smpl_npz, str(tmp_path)), nprocs=2, join=True)
    err_mean, err_overlap, d_loss, cut, n, scale = np.load(os.path.join(str(tmp_path), 'multirank.npy'))
    assert n == 5942472 and 0 < cut < n and scale > 0
    assert err_mean <= 1e-6, err_mean
    assert err_overlap == 0.0 and d_loss == 0.0


def _make_batch(ctx, dev, b, f, seed):
    from empose_b200 import synthetic
    p = synthetic.synth_window_params(b, f, seed=seed, ragged=True, offsets=True)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    r = b * f
    poses, shapes = t(p['poses']), t(p['shapes'])
    rep = lambda x, *tail: x.unsqueeze(1).repeat(1, f, *([1] * len(tail))).reshape(r, *tail)
    pos, ori, joints = ctx.sensor_project(poses.reshape(r, 66), rep(shapes, 10), rep(t(p['offset_r']), 12, 3, 3), rep(t(p['offset_t']), 12, 3))
    g = torch.Generator(device=dev).manual_seed(seed)
    inp = dict(marker_pos=(pos + 0.01 * torch.randn(pos.shape, device=dev, generator=g)).reshape(b, f, 36), marker_oris=ori.reshape(b, f, 108),
               offset_r=t(p['offset_r']), offset_t=t(p['offset_t']), seq_lengths=t(p['seq_lengths']).to(torch.int32))
    return inp, poses, shapes, joints.reshape(b, f, 66)


def _slice_batch(full, lo, hi):
    inp, poses, shapes, joints = full
    return _Batch({k: v[lo:hi].contiguous() for k, v in inp.items()}, poses[lo:hi].contiguous(), shapes[lo:hi].contiguous(),
                  joints[lo:hi].contiguous())


def _syncbn_worker(rank, world, port, npz, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    import util
    from empose_b200 import lib
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    b, f = 24, 8                                            # windows per rank
    kw = dict(precision=lib.PRECISION_FP32, device=dev, m_fk_loss=0.1, m_pose_loss_weight=10.0)      # exact arithmetic: a tight bar
    net = util.build_module(npz, **kw)
    full = _make_batch(net.native_context(dev), dev, world * b, f, seed=900)          # the SAME global batch on every rank
    shard = _slice_batch(full, rank * b, (rank + 1) * b)
    net.train()
    net(shard)                                              # re-homes the parameters; (one BatchNorm running-statistics update)
    net.sync_batchnorm(True)
    net.flat_gradients().zero_()
    out = net(shard)
    _, vals = net.backward(shard, out)
    net.allreduce_gradients(average=True)
    torch.cuda.synchronize(dev)
    calls = net.trainer(dev).sync_calls
    if rank == 0:
        ref = util.build_module(npz, **kw)                  # one device on the global batch, per-device statistics
        whole = _slice_batch(full, 0, world * b)
        ref.train()
        ref(shard)                                          # the same warm-up as above (running statistics do not enter train-mode outputs)
        ref.flat_gradients().zero_()
        ref_out = ref(whole)
        _, ref_vals = ref.backward(whole, ref_out)
        torch.cuda.synchronize(dev)
        g, g_ref = net.flat_gradients().double(), ref.flat_gradients().double()
        pose = torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1)
        pose_ref = torch.cat([ref_out['root_ori_hat'], ref_out['pose_hat']], dim=-1)[:b]
        # unsynchronised: the same shard with per-rank statistics must NOT reproduce the global-batch outputs
        net.sync_batchnorm(False)
        out_local = net(shard)
        pose_local = torch.cat([out_local['root_ori_hat'], out_local['pose_hat']], dim=-1)
        np.save(os.path.join(out_dir, 'syncbn.npy'), np.array([
            float((pose - pose_ref).abs().max()), float((pose_local - pose_ref).abs().max()),
            float((g - g_ref).norm() / g_ref.norm()), float(torch.dot(g, g_ref) / (g.norm() * g_ref.norm())), float(calls)]))
    else:
        net.sync_batchnorm(False)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_sync_batchnorm_equals_one_device_on_the_global_batch(smpl_npz, tmp_path):
    """SURVEY 8e caveat: with SyncBatchNorm (``empose_train_set_sync_batchnorm``) two ranks on half the windows each
    take the step ONE device takes on all of them -- train-mode outputs and the averaged gradient agree to rounding --
    while per-rank statistics (the DDP default) visibly do not."""
    mp.spawn(_syncbn_worker, args=(2, _free_port(), smpl_npz, str(tmp_path)), nprocs=2, join=True)
    d_pose, d_pose_local, g_rel, g_cos, calls = np.load(os.path.join(str(tmp_path), 'syncbn.npy'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import util
    util.report('sync_batchnorm_2gpu', pose_vs_global_batch=float(d_pose), pose_per_rank_statistics=float(d_pose_local),
                grad_rel=float(g_rel), grad_cos=float(g_cos), allreduce_calls=int(calls))
    assert calls == 40 + 10 + 0, calls             # N x 2 nets x 5 BatchNorm evaluations forward, 2 nets x 5 sites backward
    assert d_pose <= 2e-4, d_pose                   # train mode multiplies rounding ~5x per iteration
    assert d_pose_local >= 20 * max(d_pose, 1e-6), (d_pose_local, d_pose)
    assert g_rel <= 2e-2 and g_cos >= 0.9995, (g_rel, g_cos)
