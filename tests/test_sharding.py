"""Multi-rank host logic on CPU (gloo, world_size 2): window shards are independent, so the concatenation of
per-rank results equals the unsharded pass -- the property multi-GPU inference (bench.py --gpus N) relies on."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, npz, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    import util
    from empose_b200 import sharding, synthetic
    from oracle import ief as oracle_ief
    from oracle import sensors, smplh_lbs
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    smpl = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float64)
    topo = sensors.sensor_topology(smpl.faces.numpy())
    n_windows = 5                                              # odd on purpose: shards of 3 and 2
    params = synthetic.synth_window_params(n_windows, 6, seed=3, ragged=True, offsets=True)
    inp = util.oracle_inputs_from_params(smpl, topo, params, seed=3)
    kw = dict(n_markers=12, rnn_init=True, hidden_size=64, rnn_hidden_size=32)
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=2, rnn_init=True, hidden_size=64, rnn_hidden_size=32)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, **kw), torch.float64)
    inp64 = {k: (v.double() if v is not None and v.is_floating_point() else v) for k, v in inp.items()}
    local = oracle_ief.ief_forward(cfg, sd, smpl, topo, **sharding.shard_batch(inp64, rank, world))
    pose = sharding.gather_windows(torch.cat([local['root_ori_hat'], local['pose_hat']], dim=-1), n_windows, dist)
    joints = sharding.gather_windows(local['joints_hat'], n_windows, dist)
    if rank == 0:
        full = oracle_ief.ief_forward(cfg, sd, smpl, topo, **inp64)
        want = torch.cat([full['root_ori_hat'], full['pose_hat']], dim=-1)
        np.save(os.path.join(out_dir, 'err.npy'), np.array([(pose - want).abs().max().item(),
                                                            (joints - full['joints_hat']).abs().max().item()]))
    dist.barrier()
    dist.destroy_process_group()


def _train_worker(rank, world, port, npz, out_dir):
    """Data-parallel training step on CPU: per-shard oracle gradients -> ONE all-reduce of the flat vector."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    import util
    from empose_b200 import sharding, synthetic
    from oracle import ief as oracle_ief
    from oracle import sensors, smplh_lbs
    from oracle import train as oracle_train
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    smpl = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float64)
    topo = sensors.sensor_topology(smpl.faces.numpy())
    n_windows, n_frames = 4, 5
    params = synthetic.synth_window_params(n_windows, n_frames, seed=5, ragged=False, offsets=True)
    inp = util.oracle_inputs_from_params(smpl, topo, params, seed=5)
    kw = dict(n_markers=12, rnn_init=True, hidden_size=64, rnn_hidden_size=32, batch_norm=False)
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=2, rnn_init=True, hidden_size=64, rnn_hidden_size=32,
                               no_batch_norm=True)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, **kw), torch.float64)
    full = {k: (v.double() if v is not None and v.is_floating_point() else v) for k, v in inp.items()}
    full['poses_gt'] = torch.from_numpy(params['poses']).double()
    full['shapes_gt'] = torch.from_numpy(params['shapes']).double()
    full['joints_gt'] = torch.zeros(n_windows, n_frames, 66, dtype=torch.float64)
    weights = dict(pose_weight=10.0, shape_weight=1.0, r_weight=0.01, fk_weight=0.1)
    local = oracle_train.ief_train_step(cfg, sd, smpl, topo, **sharding.shard_batch(full, rank, world), **weights)
    names = sorted(local['grads'])
    flat = torch.cat([local['grads'][k].reshape(-1) for k in names])
    sharding.allreduce_mean_(flat, dist)                       # the single collective of the training step
    if rank == 0:
        whole = oracle_train.ief_train_step(cfg, sd, smpl, topo, **full, **weights)
        want = torch.cat([whole['grads'][k].reshape(-1) for k in names])
        np.save(os.path.join(out_dir, 'train_err.npy'), np.array([(flat - want).abs().max().item(), want.abs().max().item()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_whole_batch(smpl_npz, tmp_path):
    """Without BatchNorm and with equal shards, the all-reduced mean of the shard gradients is the whole-batch gradient."""
    port = _free_port()
    mp.spawn(_train_worker, args=(2, port, smpl_npz, str(tmp_path)), nprocs=2, join=True)
    err, scale = np.load(os.path.join(str(tmp_path), 'train_err.npy'))
    assert err < 1e-10 * max(scale, 1.0), (err, scale)


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    from empose_b200 import sharding
    for n in (1, 2, 5, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_equal_one(smpl_npz, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, smpl_npz, str(tmp_path)), nprocs=2, join=True)
    err = np.load(os.path.join(str(tmp_path), 'err.npy'))
    assert err[0] < 1e-12 and err[1] < 1e-12, err
