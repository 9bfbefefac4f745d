"""Diagnostic: tf32 vs fp32 executors, eval and train mode, at a few batch shapes (run on the GPU box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import util
from empose_b200 import lib as native, synthetic
from test_gpu_train import TrainBatch, build_train_module
import tempfile
dev = torch.device('cuda:0')
npz = synthetic.write_synthetic_smplh(os.path.join(tempfile.gettempdir(), 'empose_b200_assets'), seed=0)
flags = dict(n_markers=12, num_iterations=4, rnn_init=True, fk_weight=0.1, pose_weight=10.0)
for b, f, ragged in ((64, 32, True), (64, 32, False), (8, 32, True), (4, 8, True), (64, 8, False), (200, 4, False)):
    params = synthetic.synth_window_params(b, f, seed=91, ragged=ragged, offsets=True)
    ctx = util.build_module(npz, precision=native.PRECISION_FP32, device=dev).native_context(dev)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    r = b * f
    rep = lambda x, *tail: x.unsqueeze(1).repeat(1, f, *([1] * len(tail))).reshape(r, *tail)
    pos, ori, joints = ctx.sensor_project(t(params['poses']).reshape(r, 66), rep(t(params['shapes']), 10), rep(t(params['offset_r']), 12, 3, 3), rep(t(params['offset_t']), 12, 3))
    g = torch.Generator(device=dev).manual_seed(3)
    full = dict(marker_pos=(pos + 0.01 * torch.randn(pos.shape, device=dev, generator=g)).reshape(b, f, 36).cpu(), marker_oris=ori.reshape(b, f, 108).cpu(),
                offset_r=t(params['offset_r']).cpu(), offset_t=t(params['offset_t']).cpu(), seq_lengths=t(params['seq_lengths']).cpu(), marker_masks=None,
                poses_gt=t(params['poses']).cpu(), shapes_gt=t(params['shapes']).cpu(), joints_gt=joints.reshape(b, f, 66).cpu())
    live = util.valid_frame_mask(params['seq_lengths'], f)
    res = {}
    for mode in ('eval', 'train'):
        for prec in (native.PRECISION_FP32, native.PRECISION_TF32):
            net = build_train_module(npz, flags, prec, dev)
            if mode == 'eval':
                net.eval()
            with torch.no_grad():
                out = net(TrainBatch(full, dev))
            hist = np.stack([h.detach().cpu().numpy() for h in net.pose_hat_history])
            res[(mode, prec)] = (torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1).detach().cpu().numpy(), hist)
        a, c = res[(mode, native.PRECISION_FP32)], res[(mode, native.PRECISION_TF32)]
        per_iter = [float(np.abs(a[1][i][live] - c[1][i][live]).max()) for i in range(a[1].shape[0])]
        print('B=%d F=%d ragged=%s %s: tf32 vs fp32 rad %.3g, per iterate max|dpose| %s' % (b, f, ragged, mode, util.max_joint_angle_err(a[0][live], c[0][live]), ['%.2g' % v for v in per_iter]), flush=True)
