"""Benchmark of the LGD hot path: frames/s, LGD-RNN, 12 sensors, N=4, 32-frame windows (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--windows B] [--workload infer|train]

* default arm: the B200 path.  A "step" is one pass of ``IterativeErrorFeedback.forward`` over one batch
  of synthetic windows per GPU (BASELINE config 3: 4096 windows x 32 frames).  ``value`` is whole-job
  frames/s with inputs resident in HBM; ``e2e`` is the same through the host-buffer C-ABI call
  (pinned host inputs, H2D + D2H inside the timed region).  Adds ``roofline`` (tensor-core bound: the
  GEMM executor's achieved TFLOP/s from live CUDA-event timing) and ``cpu_baseline`` (rank 0, N=1).
* ``--impl reference``: the reference's own CPU implementation on the host cores.  In the build
  container that is the UNMODIFIED reference (``/root/reference`` through ``oracle.ref_shims``); on the
  GPU box, where the reference tree does not exist, it is the oracle restatement (``oracle/``).
* ``--workload train`` (BASELINE config 5, not the headline): one training step = zero_grad, train-mode forward,
  backward, ONE NCCL all-reduce of the flat gradient vector (N > 1), Adam step; 512 windows per GPU by default.
Under torchrun every rank drives its own GPU over its own shard of windows (no data-path collective:
windows are independent); NCCL is used for the barrier and the max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES = 32
FLOP_PER_FRAME = 26472448          # 2 x MACs of the learned layers, LGD-RNN-12-N4
BYTES_PER_FRAME = 1162             # algorithmic HBM bytes per frame
METRIC = 'frames/sec LGD-RNN-12 N=4 ws=32'
CEILING_FRAMES_PER_S = None        # tensor ceiling per GPU = measured sustained dense peak / FLOP_PER_FRAME (set in run_b200)
# launches of the tcgen05 executor in one step of the headline workload, by kind (N = 4 iterations, LSTM wavefront as
# ONE persistent launch): name in profiles/ncu_kernels.json -> launches per step
EXECUTOR_LAUNCHES = {'lstm': 1, 'heads': 1, 'blend_fwd': 5, 'blend_t': 4, 'chain': 4}


def ncu_kernels():
    """profiles/ncu_kernels.json: numbers extracted by scripts/ncu_to_json.py from the committed ncu --set full captures
    (one launch per kernel kind, 4096 windows x 32 frames, fp16 operand mode)."""
    path = os.path.join(ROOT, 'profiles', 'ncu_kernels.json')
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def executor_traffic():
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch of the executor, launch-weighted over
    the kinds of launch in a step; None when a kind has no committed capture."""
    k = ncu_kernels()
    if not all(name in k for name in EXECUTOR_LAUNCHES):
        return None, 'profiles/ncu_kernels.json lacks a capture for: ' + ', '.join(n for n in EXECUTOR_LAUNCHES if n not in k)
    total = sum(n * (k[name]['dram_read_bytes'] + k[name]['dram_write_bytes']) for name, n in EXECUTOR_LAUNCHES.items())
    note = '; '.join('%s %d x %.0f MB' % (name, n, (k[name]['dram_read_bytes'] + k[name]['dram_write_bytes']) / 1e6)
                     for name, n in EXECUTOR_LAUNCHES.items())
    return total / sum(EXECUTOR_LAUNCHES.values()), 'launch-weighted mean of the committed captures (profiles/ncu_kernels.json): ' + note


def asset_dir():
    d = os.path.join(tempfile.gettempdir(), 'empose_b200_assets')
    os.makedirs(d, exist_ok=True)
    return d


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'bf16_tflops': p['bf16_tflops'], 'bf16_tflops_sustained': p['bf16_tflops_sustained'],
                'hbm_gbs': p['hbm_gbs'], 'source': 'measured'}
    return {'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback'}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons for one GPU while the timed region runs."""

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        if self._run_nvml():
            return
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [p.strip() for p in out.split(',')]
                if len(parts) == 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def _run_nvml(self):
        """The same numbers through NVML (what nvidia-smi reads), every 5 ms: the timed region is ~150 ms, one nvidia-smi
        process per sample would see it once."""
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.index
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            if vis:
                idx = int(vis.split(',')[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons_fn = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [0x8, 0x40, 0x20, 0x4]          # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return False
        while not self.stop_flag.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = int(reasons_fn(h))
                self.samples.append([str(sm), str(max_sm)] + ['Active' if r & b else 'Not Active' for b in bits])
            except Exception:
                pass
            self.stop_flag.wait(0.005)
        return True

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons,
                'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
MESH = {'kind': None}          # --mesh: None = the regular lat-long mesh, 'mild' / 'wild' = sensor vertices of valence 5..7 / 4..11


def build_b200_model(device, precision='fp16'):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from empose_b200 import lib, synthetic
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.helpers.configuration import lgd_config
    from empose_b200.nn.models import IterativeErrorFeedback
    npz = synthetic.write_synthetic_smplh(asset_dir(), seed=0, irregular=MESH['kind'])
    cfg = lgd_config(n_markers=12, num_iterations=4, rnn_init=True, hidden_size=512, window_size=FRAMES)
    prec = {'fp16': lib.PRECISION_FP16, 'tf32': lib.PRECISION_TF32, 'fp32': lib.PRECISION_FP32}[precision]
    net = IterativeErrorFeedback(cfg, SMPLLayer(npz).to(dtype=torch.float32), precision=prec)
    sd = net.state_dict()
    for k, v in synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True).items():
        sd[k] = torch.from_numpy(np.asarray(v))
    net.load_state_dict(sd, strict=True)
    return net.to(device).eval()


def synth_device_inputs(ctx, n_windows, device, seed):
    """Random poses -> clean sensors through the CUDA projection -> + 1 cm noise."""
    from empose_b200 import synthetic
    p = synthetic.synth_window_params(n_windows, FRAMES, seed=seed)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(device)
    r = n_windows * FRAMES
    poses = t(p['poses']).reshape(r, 66)
    shapes = t(p['shapes']).unsqueeze(1).repeat(1, FRAMES, 1).reshape(r, 10)
    off_r = t(p['offset_r']).unsqueeze(1).repeat(1, FRAMES, 1, 1, 1).reshape(r, 12, 3, 3)
    off_t = t(p['offset_t']).unsqueeze(1).repeat(1, FRAMES, 1, 1).reshape(r, 12, 3)
    pos, ori, _ = ctx.sensor_project(poses, shapes, off_r, off_t)
    g = torch.Generator(device=device).manual_seed(seed)
    mpos = (pos + 0.01 * torch.randn(pos.shape, device=device, generator=g)).reshape(n_windows, FRAMES, 36).contiguous()
    mori = ori.reshape(n_windows, FRAMES, 108).contiguous()
    return dict(marker_pos=mpos, marker_oris=mori, offset_r=t(p['offset_r']).reshape(n_windows, 12, 9),
                offset_t=t(p['offset_t']), seq_lengths=t(p['seq_lengths']).to(torch.int32))


def timed(fn, steps, stream_sync):
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_sync()
    start.record()
    for _ in range(steps):
        fn()
    stop.record()
    stream_sync()
    return start.elapsed_time(stop)        # ms


def cpu_port_pass(n_windows, seed=123, threads=None):
    """One pass of the CPU oracle (or the unmodified reference when its tree is present) on `n_windows` windows.
    Returns (callable, kind)."""
    from empose_b200 import synthetic
    from oracle import ief as oracle_ief
    from oracle import ref_shims, sensors, smplh_lbs
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import util
    if threads:
        torch.set_num_threads(threads)
    npz = synthetic.write_synthetic_smplh(asset_dir(), seed=0)
    smpl = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float32)
    topo = sensors.sensor_topology(smpl.faces.numpy())
    params = synthetic.synth_window_params(n_windows, FRAMES, seed=seed)
    inp = util.oracle_inputs_from_params(smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float64), topo, params, seed=seed)
    weights = synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True)
    if ref_shims.reference_available():
        ref_shims.install(asset_dir(), seed=0)
        from empose.bodymodels.smpl import create_default_smpl_model
        from empose.data.data import AMASSBatch
        from empose.nn.models import create_model
        flags = ['--m_type', 'lgd', '--m_num_iterations', '4', '--m_hidden_size', '512', '--m_rnn_init', '--m_average_shape',
                 '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '12', '--window_size', '32']
        net = create_model(ref_shims.make_config(flags), create_default_smpl_model(device='cpu'))
        sd = net.state_dict()
        for k, v in weights.items():
            sd[k] = torch.from_numpy(np.asarray(v))
        net.load_state_dict(sd)
        net.eval()
        batch = AMASSBatch(list(range(n_windows)), inp['seq_lengths'], torch.from_numpy(params['poses']),
                           torch.from_numpy(params['shapes']), torch.zeros(n_windows, FRAMES, 3), torch.zeros(n_windows, FRAMES, 66))
        batch.marker_pos_synth, batch.marker_ori_synth = inp['marker_pos'], inp['marker_oris']
        batch.offset_t_augmented, batch.offset_r_augmented = inp['offset_t'], inp['offset_r']
        return (lambda: net(batch)), 'reference'
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=4, rnn_init=True)
    sd = util.torch_state_dict(weights)
    return (lambda: oracle_ief.ief_forward(cfg, sd, smpl, topo, **inp)), 'port'


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_windows = args.ref_windows
    fn, kind = cpu_port_pass(n_windows, threads=cores)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = n_windows * FRAMES * args.steps / dt
    sample = '%d windows x %d frames per step, %d steps, eval forward, %d torch threads' % (n_windows, FRAMES, args.steps, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1000.0 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'LGD-RNN (LSTM init), 12 sensors, N=4, ws=32 (BASELINE config 3), bounded CPU sample',
                   'windows_per_step': n_windows, 'frames_per_window': FRAMES},
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


class _HostBatch(object):
    """What the reference's evaluation loop hands to ``net(batch)``: an ``ABatch``-like object whose tensors live in pinned
    host memory and are moved by ``to_gpu()`` (``data.py:270-282``, ``eval/helpers.py:78-83``)."""

    def __init__(self, tensors, device=None):
        self.t = tensors
        self.device = device
        self.seq_lengths = tensors['seq_lengths']
        self.batch_size, self.seq_length = tensors['marker_pos'].shape[0], tensors['marker_pos'].shape[1]
        self.marker_masks = None

    def to_gpu(self, device):
        return _HostBatch({k: v.to(device, non_blocking=True) for k, v in self.t.items()}, device)

    def get_inputs(self, sf=None, ef=None, **kwargs):
        t = self.t
        return {'marker_pos': t['marker_pos'][:, sf:ef], 'marker_oris': t['marker_oris'][:, sf:ef], 'offset_r': t['offset_r'],
                'offset_t': t['offset_t'], 'marker_masks': None}


def train_subrecord(args, device, dist, rank, world, windows=512, steps=5):
    """BASELINE config 5 beside the headline: a short run of the data-parallel training step (train-mode forward, backward,
    ONE all-reduce of the flat gradient when N > 1, Adam).  Returns the dict that goes under ``train`` in the JSON line."""
    from empose_b200 import lib, synthetic
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.helpers.configuration import lgd_config
    from empose_b200.nn.models import IterativeErrorFeedback
    npz = synthetic.write_synthetic_smplh(asset_dir(), seed=0)
    cfg = lgd_config(n_markers=12, num_iterations=4, rnn_init=True, hidden_size=512, window_size=FRAMES, **TRAIN_FLAGS)
    net = IterativeErrorFeedback(cfg, SMPLLayer(npz).to(dtype=torch.float32), precision=lib.PRECISION_TF32)
    sd = net.state_dict()
    for k, v in synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True).items():
        sd[k] = torch.from_numpy(np.asarray(v))
    net.load_state_dict(sd, strict=True)
    net = net.to(device).eval()
    ctx = net.native_context(device)
    b = windows
    inp = synth_device_inputs(ctx, b, device, seed=2000 + rank)
    p = synthetic.synth_window_params(b, FRAMES, seed=2000 + rank)
    poses, shapes = torch.from_numpy(p['poses']).to(device), torch.from_numpy(p['shapes']).to(device)
    r = b * FRAMES
    _, _, joints = ctx.sensor_project(poses.reshape(r, 66), shapes.unsqueeze(1).repeat(1, FRAMES, 1).reshape(r, 10),
                                      torch.eye(3, device=device).repeat(r, 12, 1, 1), torch.zeros(r, 12, 3, device=device))
    batch = _TrainBatch(inp, poses, shapes, joints.reshape(b, FRAMES, 66))
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=cfg.lr)
    ar0, ar1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ar_ms, losses = [], []

    def step():
        opt.zero_grad()                      # scripts/train.py:136 (drops every .grad; the class re-attaches views of its flat vector)
        out = net(batch)
        _, vals = net.backward(batch, out)
        if dist is not None:
            ar0.record()
            net.allreduce_gradients(average=True)
            ar1.record()
        opt.step()
        losses.append(vals['total_loss'])
        if dist is not None:
            ar1.synchronize()
            ar_ms.append(ar0.elapsed_time(ar1))

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(device)

    for _ in range(3):
        step()                               # plain form: backward, then ONE all-reduce -- timed for reference (allreduce_ms)
    if dist is not None:
        net.overlap_gradient_allreduce(True, average=True)       # timed form: two buckets, the dense one under the LSTM sweep
        step()
    ms = timed(step, steps, barrier)
    if dist is not None:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    rec = {'metric': TRAIN_METRIC, 'value': world * b * FRAMES * steps / (ms / 1000.0), 'unit': 'frames/s', 'ms_per_step': ms / steps,
           'steps': steps, 'warmup': 3, 'windows_per_gpu': b, 'global_batch_windows': world * b, 'dtype': 'tf32',
           'allreduce_ms_unoverlapped': min(ar_ms[:3]) if ar_ms else 0.0,       # (the first one also sets the communicator up)
           'allreduce': 'two buckets inside backward: dense (iter-MLPs, heads) under the LSTM backward-through-time sweep on a side '
                        'stream, LSTM bucket after it' if world > 1 else 'none (one GPU)',
           'allreduce_bytes': int(net.flat_gradients().numel()) * 4 if world > 1 else 0,
           'launches_per_step': int(net._trainer.last_launch_count), 'loss_first_last': [losses[0], losses[-1]],
           'parallelism': 'data parallel: windows sharded, one NCCL all-reduce of the flat gradient (DDP semantics, per-shard BatchNorm statistics)'}
    net.invalidate()
    del net, opt, batch
    torch.cuda.empty_cache()
    return rec


def run_b200(args, rank, local_rank, world):
    from empose_b200 import lib
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    net = build_b200_model(device, args.precision)
    net.precision_guard = False          # the timed loops stay asynchronous; the guard's one read per forward is in the 'api' figure
    ctx = net.native_context(device)
    b = args.windows
    inp = synth_device_inputs(ctx, b, device, seed=1000 + rank)
    host = {k: v.cpu().pin_memory() for k, v in inp.items()}
    frames_per_step = b * FRAMES
    peaks = measured_peaks()
    ceiling = peaks['bf16_tflops_sustained'] * 1e12 / FLOP_PER_FRAME         # frames/s per GPU if the learned layers ran at the peak

    def step_device():
        ctx.forward(inp['marker_pos'], inp['marker_oris'], inp['offset_r'], inp['offset_t'], inp['seq_lengths'],
                    want_history=False, want_state=False)

    def step_host():
        # fresh windows: nobody consumes the final LSTM state, so it is not downloaded (models.py:489-492 reads it only
        # when the NEXT chunk of the same sequences follows)
        ctx.forward_host(host['marker_pos'], host['marker_oris'], host['offset_r'], host['offset_t'], host['seq_lengths'],
                         want_state=False)

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(device)

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = max_over_ranks(timed(step_device, args.steps, barrier))
    launches_per_step = ctx.last_launch_count
    sampler.stop_flag.set()
    sampler.join()
    value = world * frames_per_step * args.steps / (ms / 1000.0)

    # ---- end to end through the host-buffer C-ABI entry point ----
    for _ in range(2):
        step_host()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = max_over_ranks(timed(step_host, e2e_steps, barrier))
    e2e_sync_value = world * frames_per_step * e2e_steps / (ms_e2e / 1000.0)

    # ... and as a stream of requests, two in flight (empose_ief_submit_host / empose_ief_wait_host): what a host loop over
    # chunks does -- the upload of request k+1 and the download of request k-1 run under the pass of request k.  Every step
    # still uploads its inputs from pinned host memory and reads its results back; the host waits for each of them.
    def run_stream(steps):
        reqs, outs = [None, None], [None, None]
        for i in range(steps):
            slot = i % 2
            if reqs[slot] is not None:
                outs[slot] = ctx.wait_host(reqs[slot])
            reqs[slot] = ctx.submit_host(slot, host['marker_pos'], host['marker_oris'], host['offset_r'], host['offset_t'],
                                         host['seq_lengths'], out=outs[slot])
        for r in reqs:
            if r is not None:
                ctx.wait_host(r)

    run_stream(3)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    run_stream(e2e_steps)           # returns when the last result is on the host
    stop.record()
    barrier()
    ms_stream = max_over_ranks(start.elapsed_time(stop))
    e2e_value = world * frames_per_step * e2e_steps / (ms_stream / 1000.0)
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = frames_per_step * (66 + 10 + 66) * 4

    # ---- end to end through the drop-in class: net(batch) on a host batch, all five histories, outputs back on the host ----
    api_net = build_b200_model(device, args.precision)          # guard on: what a user of the reference surface gets
    host_batch = _HostBatch(host)

    def step_api():
        with torch.no_grad():
            out = api_net(host_batch.to_gpu(device))
        return {k: v.cpu() for k, v in out.items()}

    for _ in range(2):
        step_api()
    api_steps = max(3, min(args.steps, 5))
    ms_api = max_over_ranks(timed(step_api, api_steps, barrier))
    api = {'value': world * frames_per_step * api_steps / (ms_api / 1000.0), 'unit': 'frames/s', 'ms_per_step': ms_api / api_steps,
           'api': 'IterativeErrorFeedback.forward(batch) with the five N+1 histories kept on the module, host batch -> to_gpu() -> '
                  'outputs .cpu(); precision guard on (one scalar read per forward)',
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': frames_per_step * (63 + 3 + 10 + 66) * 4,
           'history_bytes_per_step': frames_per_step * 5 * 286 * 4}
    del api_net

    # ---- roofline of the dominant kernel (the tcgen05 GEMM executor): live CUDA-event timing of every launch ----
    ctx.set_profiling(True)
    prof_steps = max(2, min(args.steps, 5))
    for _ in range(prof_steps):
        step_device()
    torch.cuda.synchronize(device)
    gemm_ms, gemm_launches = ctx.profile_read()
    main_ms, main_launches = ctx.profile_read_main()
    ctx.set_profiling(False)
    achieved = FLOP_PER_FRAME * frames_per_step * prof_steps / (gemm_ms / 1000.0) / 1e12
    traffic, traffic_note = executor_traffic()
    clocks = sampler.summary()
    kern = ncu_kernels()
    fan = None
    if 'fan_grad' in kern and main_launches:
        # instruction roofline of the per-frame sub-model kernel: it moves its algorithmic bytes only (~3.7 KB per frame at
        # 0.8 TB/s) and issues no tensor instruction; what bounds it is instruction issue.  Warp instructions per launch from
        # the committed ncu capture (a fixed property of the build), duration live; peak = 4 schedulers x SMs x SM clock.
        n_grad = 4.0 * kern['fan_grad']['warp_instructions'] + (kern['fan_fwd']['warp_instructions'] if 'fan_fwd' in kern else 0.0)
        sm_hz = (clocks.get('sm_mhz') or peaks.get('sm_max_mhz', 1965.0)) * 1e6
        peak_ips = 4.0 * torch.cuda.get_device_properties(device).multi_processor_count * sm_hz
        fan = {'kernel': 'fan_kernel (SMPL-H sub-model forward + hand-derived reverse pass, fp32 SIMT)', 'bound': 'instruction issue',
               'share_of_step': (main_ms / prof_steps) / (ms / args.steps), 'launches_per_step': main_launches / prof_steps,
               'avg_launch_ms': main_ms / max(main_launches, 1),
               'warp_instructions_per_step': n_grad * (frames_per_step / 131072.0),
               'achieved': n_grad * (frames_per_step / 131072.0) / (main_ms / prof_steps / 1000.0) / 1e12,
               'peak': peak_ips / 1e12, 'unit': 'T warp-instructions/s',
               'frac': n_grad * (frames_per_step / 131072.0) / (main_ms / prof_steps / 1000.0) / peak_ips,
               'issue_active_pct_ncu': kern['fan_grad'].get('issue_active_pct'),
               'dram_bytes_per_launch_ncu': kern['fan_grad']['dram_read_bytes'] + kern['fan_grad']['dram_write_bytes']}
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / peaks['bf16_tflops_sustained'], 'traffic': traffic, 'traffic_note': traffic_note,
                'kernel': 'gemm_tc_kernel (tcgen05.mma kind::f16: learned layers, and the blend GEMMs as a 3-term fp16 split)'
                          if args.precision == 'fp16' else 'gemm_tc_kernel (tcgen05.mma kind::tf32; blend GEMMs kind::f16 split)',
                'peak_source': peaks['source'] + ' bf16 dense, sustained',
                'launches_per_step': gemm_launches / prof_steps, 'avg_launch_ms': gemm_ms / max(gemm_launches, 1),
                'kernel_share_of_step': (gemm_ms / prof_steps) / (ms / args.steps),
                # the whole step against the same ceiling: frames/s per GPU over (sustained peak / FLOP per frame)
                'step_frac': (value / world) / ceiling, 'step_ceiling_frames_per_s_per_gpu': ceiling,
                'e2e_step_frac': (e2e_value / world) / ceiling,
                'other_kernels': {'fan_kernel': fan},
                'algorithmic_flop_per_launch': FLOP_PER_FRAME * frames_per_step * prof_steps / max(gemm_launches, 1)}

    # ---- the same step with tf32 operands for the learned layers (what the reference's fp32 is closest to on tensor cores) ----
    tf32 = None
    if args.precision == 'fp16' and not args.no_companions:
        net32 = build_b200_model(device, 'tf32')
        c32 = net32.native_context(device)
        run32 = lambda: c32.forward(inp['marker_pos'], inp['marker_oris'], inp['offset_r'], inp['offset_t'], inp['seq_lengths'],
                                    want_history=False, want_state=False)
        for _ in range(3):
            run32()
        n32 = max(3, min(args.steps, 10))
        ms32 = max_over_ranks(timed(run32, n32, barrier))
        c32.set_profiling(True)
        for _ in range(2):
            run32()
        torch.cuda.synchronize(device)
        g32, _ = c32.profile_read()
        c32.set_profiling(False)
        a32 = FLOP_PER_FRAME * frames_per_step * 2 / (g32 / 1000.0) / 1e12
        v32 = world * frames_per_step * n32 / (ms32 / 1000.0)
        tf32 = {'value': v32, 'unit': 'frames/s', 'ms_per_step': ms32 / n32, 'executor_tflops': a32,
                'frac_of_tf32_rate': a32 / (peaks['bf16_tflops_sustained'] / 2.0), 'tf32_rate_tflops': peaks['bf16_tflops_sustained'] / 2.0,
                'step_frac_of_tf32_ceiling': (v32 / world) / (ceiling / 2.0),
                'note': 'tcgen05 kind::tf32 issues at half the f16 rate; same kernels, same launches'}
        del net32, c32

    train = None
    if not args.no_companions:
        train = train_subrecord(args, device, dist, rank, world)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fn, kind = cpu_port_pass(args.ref_windows)
        fn()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps
        cpu = {'value': args.ref_windows * FRAMES / dt, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': kind,
               'sample': '%d windows x %d frames, eval forward, mean of %d passes after 1 warm-up' % (args.ref_windows, FRAMES, reps)}
    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision,
            'data': 'synthetic',
            'config': {'workload': 'LGD-RNN (2x512 LSTM init), 12 sensors, N=4, ws=32, %d windows per GPU (BASELINE config 3)' % b,
                       'windows_per_gpu': b, 'frames_per_window': FRAMES, 'parallelism': 'windows sharded, no collective',
                       'mesh': args.mesh + ' (synthetic; sensor sub-mesh in fan form)',
                       'l2': 'per-step working set (~3 GB of activations and features) exceeds the 126 MB L2; no flush needed',
                       'arithmetic': {'fp16': 'learned layers: fp16 operands on tcgen05 (kind::f16), fp32 accumulate in TMEM; blend GEMMs as a '
                                              '3-term fp16 split (22 mantissa bits); per-frame SMPL math, LSTM cell state and all outputs fp32',
                                      'tf32': 'tf32 tensor cores (tcgen05), fp32 accumulate; blend GEMMs as a 3-term fp16 split; per-frame SMPL math fp32',
                                      'fp32': 'FFMA executor, fp32 everywhere (parity mode)'}[args.precision]},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_stream / e2e_steps,
                    'api': 'empose_ief_submit_host + empose_ief_wait_host: a stream of requests, two in flight; every request uploads its '
                           'inputs from pinned host buffers and downloads pose, shape and joints; the host waits for every result',
                    'frac_of_device_rate': e2e_value / value,
                    'sync': {'value': e2e_sync_value, 'unit': 'frames/s', 'ms_per_step': ms_e2e / e2e_steps,
                             'api': 'empose_ief_forward_host: ONE blocking call per step (copies overlap only inside the call)',
                             'frac_of_device_rate': e2e_sync_value / value}},
            'e2e_api': api,
            'gpu_launches': int(launches_per_step * args.steps), 'launches_per_step': int(launches_per_step),
            'roofline': roofline, 'tf32': tf32, 'train': train, 'cpu_baseline': cpu}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


class _TrainBatch(object):
    """The slice of the reference's AMASSBatch that forward / backward read (data.py:304-309, 433-459)."""

    def __init__(self, inp, poses, shapes, joints):
        self.inp = inp
        self.seq_lengths = inp['seq_lengths']
        self.poses_root, self.poses_body = poses[:, :, :3].contiguous(), poses[:, :, 3:].contiguous()
        self.shapes, self.joints_gt, self.marker_masks = shapes, joints, None
        self.batch_size, self.seq_length = poses.shape[0], poses.shape[1]

    def get_inputs(self, sf=None, ef=None, **kwargs):
        i = self.inp
        return {'marker_pos': i['marker_pos'], 'marker_oris': i['marker_oris'], 'offset_r': i['offset_r'],
                'offset_t': i['offset_t'], 'marker_masks': None}


TRAIN_FLAGS = dict(m_fk_loss=0.1, m_pose_loss_weight=10.0, m_reprojection_loss_weight=0.01, lr=5e-4)   # README.md:221
TRAIN_METRIC = 'frames/sec LGD-RNN-12 N=4 ws=32 training step (fwd + bwd + grad all-reduce + Adam)'


def cpu_train_pass(n_windows, seed=123):
    """One CPU training step (oracle restatement: forward, both backward passes; no optimiser) on n_windows windows."""
    from empose_b200 import synthetic
    from oracle import ief as oracle_ief
    from oracle import sensors, smplh_lbs
    from oracle import train as oracle_train
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import util
    npz = synthetic.write_synthetic_smplh(asset_dir(), seed=0)
    smpl = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float32)
    smpl64 = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float64)
    topo = sensors.sensor_topology(smpl.faces.numpy())
    params = synthetic.synth_window_params(n_windows, FRAMES, seed=seed)
    inp = util.oracle_inputs_from_params(smpl64, topo, params, seed=seed)
    poses, shapes = torch.from_numpy(params['poses']), torch.from_numpy(params['shapes'])
    with torch.no_grad():
        r = n_windows * FRAMES
        _, _, joints = oracle_ief.project_sensors(smpl, topo, poses.reshape(r, 66), shapes.unsqueeze(1).repeat(1, FRAMES, 1).reshape(r, 10),
                                                  torch.eye(3).repeat(r, 12, 1, 1), torch.zeros(r, 12, 3))
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=4, rnn_init=True)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True))
    return lambda: oracle_train.ief_train_step(cfg, sd, smpl, topo, poses_gt=poses, shapes_gt=shapes,
                                               joints_gt=joints.reshape(n_windows, FRAMES, 66), pose_weight=10.0, shape_weight=1.0,
                                               r_weight=0.01, fk_weight=0.1, **inp)


def run_b200_train(args, rank, local_rank, world):
    """BASELINE config 5: LGD-RNN-12 training step, windows sharded over the GPUs, one gradient all-reduce."""
    from empose_b200 import lib, synthetic
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.helpers.configuration import lgd_config
    from empose_b200.nn.models import IterativeErrorFeedback
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    npz = synthetic.write_synthetic_smplh(asset_dir(), seed=0)
    cfg = lgd_config(n_markers=12, num_iterations=4, rnn_init=True, hidden_size=512, window_size=FRAMES, **TRAIN_FLAGS)
    net = IterativeErrorFeedback(cfg, SMPLLayer(npz).to(dtype=torch.float32), precision=lib.PRECISION_TF32)
    sd = net.state_dict()
    for k, v in synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True).items():
        sd[k] = torch.from_numpy(np.asarray(v))
    net.load_state_dict(sd, strict=True)
    net = net.to(device)
    b = args.windows
    net.eval()
    ctx = net.native_context(device)
    inp = synth_device_inputs(ctx, b, device, seed=2000 + rank)
    p = synthetic.synth_window_params(b, FRAMES, seed=2000 + rank)
    poses = torch.from_numpy(p['poses']).to(device)
    shapes = torch.from_numpy(p['shapes']).to(device)
    r = b * FRAMES
    eye = torch.eye(3, device=device).repeat(r, 12, 1, 1)
    _, _, joints = ctx.sensor_project(poses.reshape(r, 66), shapes.unsqueeze(1).repeat(1, FRAMES, 1).reshape(r, 10), eye,
                                      torch.zeros(r, 12, 3, device=device))
    batch = _TrainBatch(inp, poses, shapes, joints.reshape(b, FRAMES, 66))
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=cfg.lr)
    losses = []

    def step():
        opt.zero_grad()
        out = net(batch)
        _, vals = net.backward(batch, out)
        if dist is not None:
            net.allreduce_gradients(average=True)
        opt.step()
        losses.append(vals['total_loss'])

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(device)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(step, args.steps, barrier)
    if dist is not None:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sampler.stop_flag.set()
    sampler.join()
    launches = net._trainer.last_launch_count
    value = world * b * FRAMES * args.steps / (ms / 1000.0)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fn = cpu_train_pass(args.ref_windows)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        cpu = {'value': args.ref_windows * FRAMES / dt, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '%d windows x %d frames, one training step (forward + gradients) after 1 warm-up' % (args.ref_windows, FRAMES)}
    if rank == 0:
        n_grad = int(net.flat_gradients().numel())
        print(json.dumps({
            'metric': TRAIN_METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'tf32', 'data': 'synthetic',
            'config': {'workload': 'LGD-RNN-12 N=4 ws=32 TRAINING step, %d windows per GPU (BASELINE config 5)' % b,
                       'windows_per_gpu': b, 'frames_per_window': FRAMES, 'global_batch_windows': world * b,
                       'parallelism': 'data parallel: windows sharded, one NCCL all-reduce of the %d-float flat gradient' % n_grad,
                       'optimizer': 'torch.optim.Adam on views of the flat parameter vector', 'loss_weights': TRAIN_FLAGS,
                       'l2': 'activations kept for the backward pass (~%.1f GB) exceed the 126 MB L2' % (r * 4 * 512 * 4 * 44 / 1e9)},
            'clocks': sampler.summary(), 'gpu_launches': int(launches * args.steps),
            'loss_first_last': [losses[0], losses[-1]], 'cpu_baseline': cpu}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


BIRNN_METRIC = 'frames/sec BiRNN-2x1024 6-sensor, one full stream (batch 1)'


def run_b200_birnn(args, rank, local_rank, world):
    """BASELINE config 4: the bidirectional 2 x 1024 LSTM baseline on ONE long 6-sensor stream (batch 1, ~15k frames).
    Along time the recurrence does not shard: under torchrun every rank runs its own replica stream."""
    from empose_b200 import lib, synthetic
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.helpers.configuration import Configuration
    from empose_b200.nn.models import SimpleRNN
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import util
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    npz = synthetic.write_synthetic_smplh(asset_dir(), seed=0)
    wk = dict(n_markers=6, hidden_size=1024, num_layers=2, bidirectional=True, estimate_shape=False)
    cfg = Configuration(dict(m_type='rnn', m_hidden_size=1024, m_num_layers=2, m_bidirectional=True, use_marker_pos=True,
                             use_marker_ori=True, n_markers=6, window_size=32))
    prec = {'fp16': lib.PRECISION_FP16, 'tf32': lib.PRECISION_TF32, 'fp32': lib.PRECISION_FP32}[args.precision]
    net = SimpleRNN(cfg, SMPLLayer(npz).to(dtype=torch.float32), precision=prec)
    sd = net.state_dict()
    for k, v in synthetic.synth_rnn_state_dict(seed=0, **wk).items():
        sd[k] = torch.from_numpy(np.asarray(v))
    net.load_state_dict(sd, strict=True)
    net = net.to(device).eval()
    f = args.frames
    g = torch.Generator().manual_seed(77 + rank)
    pos = (0.3 * torch.randn(1, f, 36, generator=g)).pin_memory()
    ori = (torch.eye(3).reshape(1, 1, 1, 9) + 0.05 * torch.randn(1, f, 12, 9, generator=g)).reshape(1, f, 108).pin_memory()
    lens = torch.tensor([f])
    z = torch.zeros(1, 12, 3)
    host = util.DuckBatch(pos, ori, z.unsqueeze(-1).repeat(1, 1, 1, 3), z, lens)
    dev_batch = host.to(device)

    def step_device():
        with torch.no_grad():
            net(dev_batch)

    def step_host():
        with torch.no_grad():
            out = net(host.to(device))
        return out['pose_hat'].cpu(), out['root_ori_hat'].cpu()

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(device)

    t0 = time.perf_counter()
    step_device()
    torch.cuda.synchronize(device)
    plan_s = time.perf_counter() - t0                      # first call: builds the execution plan (jobs + tensor maps)
    for _ in range(max(args.warmup, 3) - 1):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(step_device, args.steps, barrier)
    ms_e2e = timed(step_host, max(2, min(args.steps, 5)), barrier) / max(2, min(args.steps, 5))
    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    sampler.stop_flag.set()
    sampler.join()
    launches = net._ctx.last_launch_count
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import rnn as oracle_rnn
        n = min(f, 1500)
        sd32 = util.torch_state_dict(synthetic.synth_rnn_state_dict(seed=0, **wk))
        ocfg = oracle_rnn.RnnConfig(n_markers=6, hidden_size=1024, num_layers=2, bidirectional=True)
        fn = lambda: oracle_rnn.rnn_forward(ocfg, sd32, None, pos[:, :n], ori[:, :n], torch.tensor([n]))
        with torch.no_grad():
            fn()
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
        cpu = {'value': n / dt, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': 'the first %d frames of the stream, one pass after 1 warm-up' % n}
    if rank == 0:
        print(json.dumps({
            'metric': BIRNN_METRIC, 'value': world * f * args.steps / (ms / 1000.0), 'unit': 'frames/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
            'config': {'workload': 'BiRNN 2 x 1024, 6 sensors, one stream of %d frames, batch 1 (BASELINE config 4); replicas only across GPUs' % f,
                       'frames': f, 'plan_build_s': plan_s,
                       'note': 'fp16 mode: input projections of all frames as one tensor-core GEMM per layer, recurrence in the persistent '
                               'kernel (W_hh resident in shared memory across 128 cooperating CTAs, hidden vector handed over through L2); '
                               'latency-bound by design: ~4 us per time step'},
            'clocks': sampler.summary(), 'gpu_launches': int(launches * args.steps),
            'e2e': {'value': world * f / (ms_e2e / 1000.0), 'unit': 'frames/s', 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': f * 144 * 4, 'd2h_bytes_per_step': f * 66 * 4, 'api': 'SimpleRNN.forward on a host batch + .cpu()'},
            'cpu_baseline': cpu}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_b200_synth(args, rank, local_rank, world):
    """training-data synthesis (SMPLFK + SampleMarkersWithOffsets) on the device, windows sharded over GPUs."""
    from empose_b200 import synthetic
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.data.transforms import SMPLFK, SampleMarkersWithOffsets
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    npz = synthetic.write_synthetic_smplh(asset_dir(), seed=0)
    smpl = SMPLLayer(npz).to(device=device, dtype=torch.float32)
    files = synthetic.write_synthetic_offsets(asset_dir(), n_files=3, seed=0)
    b = args.windows
    p = synthetic.synth_window_params(b, FRAMES, seed=3000 + rank)

    class B(object):
        pass
    batch = B()
    poses = torch.from_numpy(p['poses']).to(device)
    batch.poses_root, batch.poses_body = poses[:, :, :3].contiguous(), poses[:, :, 3:].contiguous()
    batch.shapes = torch.from_numpy(p['shapes']).to(device)
    batch.trans = torch.zeros(b, FRAMES, 3, device=device)
    batch.batch_size, batch.seq_length = b, FRAMES
    fk, sampler = SMPLFK(smpl), SampleMarkersWithOffsets(smpl, files, noise_level=0)

    def step():
        sampler(fk(batch))

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(device)

    for _ in range(max(args.warmup, 3)):
        step()
    ms = timed(step, args.steps, barrier)
    if dist is not None:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ief as oracle_ief
        from oracle import sensors, smplh_lbs
        osmpl = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float32)
        topo = sensors.sensor_topology(osmpl.faces.numpy())
        n = 16 * FRAMES
        pp = torch.from_numpy(p['poses'][:16]).reshape(n, 66)
        ss = torch.from_numpy(p['shapes'][:16]).unsqueeze(1).repeat(1, FRAMES, 1).reshape(n, 10)
        eye, zero = torch.eye(3).repeat(n, 12, 1, 1), torch.zeros(n, 12, 3)
        with torch.no_grad():
            oracle_ief.project_sensors(osmpl, topo, pp, ss, eye, zero)
            t0 = time.perf_counter()
            oracle_ief.project_sensors(osmpl, topo, pp, ss, eye, zero)       # full mesh + sensor frames, once (the reference
            dt = time.perf_counter() - t0                                    # derives both marker sets from one mesh)
        cpu = {'value': n / dt, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '16 windows x 32 frames: full-mesh SMPL-H + sensor frames + offsets, one pass after 1 warm-up'}
    if rank == 0:
        print(json.dumps({
            'metric': 'frames/sec training-data synthesis (SMPLFK + SampleMarkersWithOffsets), 12 sensors, ws=32',
            'value': world * b * FRAMES * args.steps / (ms / 1000.0), 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp32 (pose blend 3xTF32)', 'data': 'synthetic',
            'config': {'workload': 'SMPLFK + SampleMarkersWithOffsets(noise_level=0) on %d windows x 32 frames per GPU: three passes of '
                                   'the sub-model kernels (joints; raw sensor frames; frames with offsets), the mesh is never materialised' % b,
                       'windows_per_gpu': b}, 'cpu_baseline': cpu}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_b200_metrics(args, rank, local_rank, world):
    """MetricsEngine.compute on the device (FK x2 + Procrustes + angular distances, one kernel)."""
    from empose_b200 import synthetic
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.eval.metrics import MetricsEngine
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    npz = synthetic.write_synthetic_smplh(asset_dir(), seed=0)
    me = MetricsEngine(SMPLLayer(npz).to(device=device, dtype=torch.float32))
    b = args.windows
    p = synthetic.synth_window_params(b, FRAMES, seed=4000 + rank)
    g = torch.Generator().manual_seed(4000 + rank)
    poses = torch.from_numpy(p['poses'])
    poses_hat = poses + 0.08 * torch.randn(poses.shape, generator=g)
    shapes = torch.from_numpy(p['shapes'])
    dv = lambda t: t.to(device)
    args_dev = (dv(poses[:, :, 3:]), dv(shapes), dv(poses_hat[:, :, 3:]), None, None, dv(poses[:, :, :3]), dv(poses_hat[:, :, :3]))

    def step():
        me.reset()
        me.compute(*args_dev)                   # includes the (frames x joints) tables coming back to the host
        return me.get_metrics()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        metrics = step()
    ms = 1000.0 * (time.perf_counter() - t0)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import metrics as oracle_metrics
        from oracle import smplh_lbs
        osmpl = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float32)
        n = 16 * FRAMES
        a = (poses[:16].reshape(n, 66), shapes[:16].unsqueeze(1).repeat(1, FRAMES, 1).reshape(n, 10), poses_hat[:16].reshape(n, 66),
             shapes[:16].unsqueeze(1).repeat(1, FRAMES, 1).reshape(n, 10))
        oracle_metrics.frame_metrics(osmpl, *a)
        t0 = time.perf_counter()
        oracle_metrics.frame_metrics(osmpl, *a)
        cpu = {'value': n / (time.perf_counter() - t0), 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '16 windows x 32 frames: two full-mesh FK passes + per-frame numpy Procrustes + angular distances'}
    if rank == 0:
        print(json.dumps({
            'metric': 'frames/sec MetricsEngine.compute (MPJPE, PA-MPJPE, MPJAE)', 'value': b * FRAMES * args.steps / (ms / 1000.0),
            'unit': 'frames/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (Procrustes in f64)', 'data': 'synthetic',
            'config': {'workload': 'MetricsEngine.compute + get_metrics on %d windows x 32 frames, wall clock incl. the result tables '
                                   'coming back to the host' % b, 'windows_per_gpu': b},
            'metrics': {k: float(v) for k, v in metrics.items()}, 'cpu_baseline': cpu}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='infer', choices=['infer', 'train', 'birnn', 'synth', 'metrics'])
    ap.add_argument('--frames', type=int, default=15000, help='stream length of the birnn workload')
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'tf32', 'fp32'], help='inference arithmetic of the learned layers')
    ap.add_argument('--windows', type=int, default=None, help='windows per GPU (inference: 4096 = BASELINE config 3; training: 512)')
    ap.add_argument('--ref-windows', type=int, default=16, help='windows per step of the CPU reference arm / baseline')
    ap.add_argument('--mesh', default='regular', choices=['regular', 'mild', 'wild'],
                    help='synthetic SMPL-H mesh of the inference workload: regular valence 6, or irregular sensor valences 5..7 / 4..11')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-companions', action='store_true', help='skip the tf32 and training sub-records (quick A/B runs)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    MESH['kind'] = None if args.mesh == 'regular' else args.mesh
    if args.windows is None:
        args.windows = 512 if args.workload == 'train' else 4096
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)')
    if args.workload == 'metrics':
        run_b200_metrics(args, rank, local_rank, world)
    elif args.workload == 'synth':
        run_b200_synth(args, rank, local_rank, world)
    elif args.workload == 'birnn':
        run_b200_birnn(args, rank, local_rank, world)
    elif args.workload == 'train':
        run_b200_train(args, rank, local_rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
