"""Parity of the CUDA training step (empose_train_*, through the model class) against the oracle's training
step and the gradients of the unmodified reference (tests/golden/train_*.npz).  Run on the B200 box: -m gpu.

How gradients are compared.  The training loss is only piecewise smooth: every PReLU (10 per MLP evaluation,
layers.py:28,61) and every L1 term (models.py:457) has a kink, and with ~10^5..10^6 pre-activations per step some
sit within 1e-7 of zero, so ANY change of arithmetic flips a few branch decisions and moves individual tensors
by ~1 % (the slope gradients of PReLU, sums with heavy cancellation, by more).  The unmodified reference does it
to itself: torch 2.11 on 4 instead of 1 CPU threads returns a 5 % different ``shape_net_iter.activation_fn``
gradient on the ``train_lgd_mlp12_n2`` fixture (see tests/golden/make_golden_train.py).  The bars are therefore
  FP32 executor : median per-tensor relative error <= 2e-4 (the algorithm is exact), every tensor <= 0.1 (PReLU
                  slopes <= 0.5), cosine of the whole gradient vector >= 0.9999;
  TF32 product  : median <= 3e-2, every tensor <= 0.35 (PReLU slopes unbounded but reported), cosine >= 0.99.
Loss values: 1e-4 / 2e-3 relative; forward outputs of the train-mode pass: 2e-5 / 5e-3 rad (batch statistics over
24..40 rows amplify TF32 rounding by 1/sigma).
"""
import sys

import numpy as np
import pytest
import torch

from empose_b200 import lib as native
from empose_b200 import synthetic
from oracle import ief as oracle_ief
from oracle import train as oracle_train

import util

sys.path.insert(0, util.GOLDEN_DIR)
from make_golden_train import sample_positions  # noqa: E402

pytestmark = pytest.mark.gpu

PNAME = {native.PRECISION_FP32: 'fp32', native.PRECISION_TF32: 'tf32'}
GRAD_MEDIAN = {native.PRECISION_FP32: 2e-4, native.PRECISION_TF32: 3e-2}
GRAD_MAX = {native.PRECISION_FP32: 0.1, native.PRECISION_TF32: 0.35}
GRAD_COS = {native.PRECISION_FP32: 0.9999, native.PRECISION_TF32: 0.99}
LOSS_TOL = {native.PRECISION_FP32: 1e-4, native.PRECISION_TF32: 2e-3}
FWD_RAD = {native.PRECISION_FP32: 2e-5, native.PRECISION_TF32: 5e-3}


def is_prelu_slope(net, key):
    mod = net
    for part in key.split('.')[:-1]:
        mod = getattr(mod, part) if not part.isdigit() else mod[int(part)]
    return isinstance(mod, torch.nn.PReLU)


class TrainBatch(util.DuckBatch):
    """DuckBatch plus the targets ``IterativeErrorFeedback.backward`` reads (models.py:649-660)."""

    def __init__(self, inp, dev):
        mv = lambda t: None if t is None else t.to(dev)
        super(TrainBatch, self).__init__(mv(inp['marker_pos']), mv(inp['marker_oris']), mv(inp['offset_r']), mv(inp['offset_t']),
                                         mv(inp['seq_lengths']), mv(inp['marker_masks']))
        self.poses_root = mv(inp['poses_gt'][:, :, :3])
        self.poses_body = mv(inp['poses_gt'][:, :, 3:])
        self.shapes = mv(inp['shapes_gt'])
        self.joints_gt = mv(inp['joints_gt'])


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'these tests need the B200'
    return torch.device('cuda:0')


def build_train_module(smpl_npz, flags, precision, dev):
    net = util.build_module(smpl_npz, n_markers=flags['n_markers'], num_iterations=flags['num_iterations'],
                            rnn_init=flags['rnn_init'], precision=precision, device=dev,
                            m_fk_loss=flags['fk_weight'], m_pose_loss_weight=flags['pose_weight'])
    return net.train()


@pytest.mark.parametrize('precision', [native.PRECISION_FP32, native.PRECISION_TF32], ids=PNAME.get)
@pytest.mark.parametrize('name', sorted(util.TRAIN_CASES))
def test_training_step_matches_oracle_and_reference(dev, smpl_npz, oracle_smpl, topology, name, precision):
    gold = util.load_golden(name)
    flags = util.TRAIN_CASES[name]
    cfg = oracle_ief.IefConfig(n_markers=flags['n_markers'], num_iterations=flags['num_iterations'], rnn_init=flags['rnn_init'])
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=flags['n_markers'], rnn_init=flags['rnn_init']),
                               torch.float64)
    inp64 = util.train_inputs(gold, torch.float64)
    want = oracle_train.ief_train_step(cfg, sd, oracle_smpl, topology, pose_weight=flags['pose_weight'], shape_weight=1.0,
                                       r_weight=0.01, fk_weight=flags['fk_weight'], **inp64)

    net = build_train_module(smpl_npz, flags, precision, dev)
    batch = TrainBatch(util.train_inputs(gold, torch.float32), dev)
    for p in net.parameters():
        p.grad = None
    out = net(batch)
    total, loss_vals = net.backward(batch, out)
    torch.cuda.synchronize()

    rows = []
    # forward outputs of the train-mode pass
    live = util.valid_frame_mask(gold['seq_lengths'], batch.seq_length)
    pose = torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1).cpu().numpy()
    want_pose = torch.cat([want['root_ori_hat'], want['pose_hat']], dim=-1).numpy()
    rad = util.max_joint_angle_err(pose[live], want_pose[live])
    mm = util.max_joint_pos_err_mm(out['joints_hat'].cpu().numpy()[live], want['joints_hat'].numpy()[live])
    # losses
    loss_err = {k: abs(loss_vals[k] - want['loss_vals'][k]) / max(1.0, abs(want['loss_vals'][k])) for k in loss_vals}
    # gradients: per tensor, relative to the tensor's own norm (floor for analytically-zero gradients)
    items = dict(net.named_parameters())
    worst, worst_slope, dot, n_got, n_want = 0.0, 0.0, 0.0, 0.0, 0.0
    for key, g_want in want['grads'].items():
        g = items[key].grad.detach().cpu().double().reshape(-1).numpy()
        w = g_want.reshape(-1).numpy()
        norm = float(np.sqrt((w * w).sum()))
        err = float(np.sqrt(((g - w) ** 2).sum()))
        # floor: tensors whose gradient is analytically zero (biases in front of a BatchNorm) hold ~1e-8 noise in the oracle
        rel = err / (norm + 1e-5 * np.sqrt(w.size))
        pos = sample_positions(key, w.size)
        ref_err = float(np.abs(g[pos] - gold['g/' + key + '/samples']).max() / (np.abs(gold['g/' + key + '/samples']).max() + 1e-6))
        rows.append((key, rel, ref_err, norm))
        dot += float((g * w).sum()); n_got += float((g * g).sum()); n_want += float((w * w).sum())
        if is_prelu_slope(net, key):
            worst_slope = max(worst_slope, rel)
        else:
            worst = max(worst, rel)
    # running statistics
    buf_err = 0.0
    bufs = dict(net.named_buffers())
    for key, b_want in want['buffers'].items():
        if 'num_batches' in key:
            assert int(bufs[key]) == int(b_want), key
            continue
        buf_err = max(buf_err, float((bufs[key].detach().cpu().double() - b_want).abs().max()))
    rows.sort(key=lambda r: -r[1])
    cosine = dot / np.sqrt(n_got * n_want)
    median = float(np.median([r[1] for r in rows]))
    util.report('train_step', case=name, precision=PNAME[precision], rad=rad, mm=mm, worst_grad_rel=worst, buf_err=buf_err,
                median_grad_rel=median, worst_slope_rel=worst_slope, cosine=cosine,
                loss_err=max(loss_err.values()), total_loss=loss_vals['total_loss'], want_total=want['loss_vals']['total_loss'],
                worst5=[(r[0], round(r[1], 6), round(r[2], 6)) for r in rows[:5]],
                launches=net._trainer.last_launch_count)
    assert np.isfinite(rad) and rad <= FWD_RAD[precision], rad
    assert max(loss_err.values()) <= LOSS_TOL[precision], loss_err
    assert median <= GRAD_MEDIAN[precision], (median, rows[:5])
    assert worst <= GRAD_MAX[precision], rows[:5]
    assert cosine >= GRAD_COS[precision], cosine
    if precision == native.PRECISION_FP32:
        assert worst_slope <= 0.5, rows[:5]
    assert buf_err <= (1e-5 if precision == native.PRECISION_FP32 else 2e-3), buf_err
    assert abs(float(total) - loss_vals['total_loss']) < 1e-6


def test_optimizer_step_and_eval_after_training(dev, smpl_npz):
    """scripts/train.py:125-152 shape: Adam on net.parameters(), zero_grad / forward / backward / step; then the eval
    path must see the updated weights and running statistics."""
    flags = util.TRAIN_CASES['train_lgd_rnn12_n4']
    gold = util.load_golden('train_lgd_rnn12_n4')
    net = build_train_module(smpl_npz, flags, native.PRECISION_TF32, dev)
    batch = TrainBatch(util.train_inputs(gold, torch.float32), dev)
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        out = net(batch)
        _, vals = net.backward(batch, out)
        opt.step()
        losses.append(vals['total_loss'])
    util.report('train_loop', losses=[round(v, 5) for v in losses])
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0], losses                     # the same batch eight times: the loss must go down
    flat = net.flat_parameters()
    assert flat is not None and net.pose_net_init.weight.data_ptr() >= flat.data_ptr()
    net.eval()
    with torch.no_grad():
        out_eval = net(batch)
    assert torch.isfinite(out_eval['pose_hat']).all()
    # gradient accumulation semantics: two backward passes without zero_grad double the gradient
    net.train()
    opt.zero_grad()
    out = net(batch)
    net.backward(batch, out)
    g1 = net.flat_gradients().clone()
    out = net(batch)
    net.backward(batch, out)
    g2 = net.flat_gradients()
    rel = float((g2 - 2 * g1).norm() / g1.norm())
    assert rel < 0.05, rel          # BatchNorm running stats do not enter train-mode outputs; only rounding differs
