"""Parity of the (Bi)RNN baseline on the B200 (empose_rnn_*, through the SimpleRNN class) against the golden outputs of
the unmodified reference and the oracle.  Run on the B200 box: -m gpu.  Bars as for the LGD path: <= 1e-4 rad per-joint
rotation, <= 0.1 mm joint position in the tensor-core modes; the FP32 executor is held to 2e-5 rad / 0.02 mm."""
import numpy as np
import pytest
import torch

from empose_b200 import lib as native
from empose_b200 import synthetic
from oracle import rnn as oracle_rnn

import util

pytestmark = pytest.mark.gpu

PRECISIONS = [native.PRECISION_FP32, native.PRECISION_TF32, native.PRECISION_FP16]
PNAME = {native.PRECISION_FP32: 'fp32', native.PRECISION_TF32: 'tf32', native.PRECISION_FP16: 'fp16'}


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'these tests need the B200'
    return torch.device('cuda:0')


def build_rnn(smpl_npz, flags, precision, dev):
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.helpers.configuration import Configuration
    from empose_b200.nn.models import SimpleRNN, create_model
    c, w = flags['cfg'], flags['weights']
    cfg = Configuration(dict(m_type='rnn', m_hidden_size=c['hidden_size'], m_num_layers=c['num_layers'],
                             m_bidirectional=c['bidirectional'], m_estimate_shape=c.get('estimate_shape', False),
                             m_average_shape=c.get('average_shape', False), m_fk_loss=0.1 if c.get('do_fk') else 0.0,
                             use_marker_pos=True, use_marker_ori=True, n_markers=c['n_markers'], window_size=32))
    net = create_model(cfg, SMPLLayer(smpl_npz).to(dtype=torch.float32))
    assert isinstance(net, SimpleRNN)
    net.precision = precision
    sd = net.state_dict()
    synth = synthetic.synth_rnn_state_dict(seed=0, **w)
    assert {k for k in sd if not k.startswith('smpl.')} == set(synth)
    for k, v in synth.items():
        sd[k] = torch.from_numpy(np.asarray(v))
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval()


@pytest.mark.parametrize('precision', PRECISIONS, ids=PNAME.get)
@pytest.mark.parametrize('name', sorted(util.RNN_CASES))
def test_rnn_matches_reference_golden(dev, smpl_npz, name, precision):
    gold = util.load_golden(name)
    flags = util.RNN_CASES[name]
    net = build_rnn(smpl_npz, flags, precision, dev)
    rad_tol, mm_tol = (2e-5, 0.02) if precision == native.PRECISION_FP32 else (1e-4, 0.1)
    c = 0
    while ('c%d_pose_hat' % c) in gold:
        tag = 'c%d_' % c
        pos, ori = torch.from_numpy(gold[tag + 'marker_pos']), torch.from_numpy(gold[tag + 'marker_oris'])
        lens = torch.from_numpy(gold[tag + 'seq_lengths'])
        z = torch.zeros(pos.shape[0], 12, 3)
        batch = util.DuckBatch(pos, ori, z.unsqueeze(-1).repeat(1, 1, 1, 3), z, lens).to(dev)
        with torch.no_grad():
            out = net(batch, is_new_sequence=(c == 0))
        live = util.valid_frame_mask(gold[tag + 'seq_lengths'], pos.shape[1])
        rad = max(util.max_joint_angle_err(out['pose_hat'].cpu().numpy()[live], gold[tag + 'pose_hat'][live]),
                  util.max_joint_angle_err(out['root_ori_hat'].cpu().numpy()[live], gold[tag + 'root_ori_hat'][live]))
        mm = 0.0
        if (tag + 'joints_hat') in gold:
            mm = util.max_joint_pos_err_mm(out['joints_hat'].cpu().numpy()[live], gold[tag + 'joints_hat'][live])
            np.testing.assert_allclose(out['shape_hat'].cpu().numpy()[live], gold[tag + 'shape_hat'][live], atol=5 * rad_tol, rtol=0)
        else:
            assert out['joints_hat'] is None and out['shape_hat'] is None
        state_err = float(np.abs(net.rnn.final_state[0].cpu().numpy() - gold[tag + 'final_h']).max())
        util.report('rnn_golden', case=name, chunk=c, precision=PNAME[precision], rad=rad, mm=mm, state_err=state_err,
                    launches=net._ctx.last_launch_count)
        assert rad <= rad_tol, rad
        assert mm <= mm_tol, mm
        assert state_err <= (5e-6 if precision == native.PRECISION_FP32 else 2e-3)
        np.testing.assert_allclose(net.rnn.final_state[1].cpu().numpy(), gold[tag + 'final_c'],
                                   atol=1e-5 if precision == native.PRECISION_FP32 else 5e-3, rtol=0)
        c += 1
    assert c >= 1


@pytest.mark.parametrize('precision', [native.PRECISION_FP32, native.PRECISION_FP16], ids=PNAME.get)
def test_birnn_stream_matches_oracle(dev, smpl_npz, oracle_smpl, precision):
    """BASELINE config 4 in the small: ONE long 6-sensor sequence through the bidirectional model (batch 1)."""
    flags = dict(cfg=dict(n_markers=6, hidden_size=256, num_layers=2, bidirectional=True),
                 weights=dict(n_markers=6, hidden_size=256, num_layers=2, bidirectional=True, estimate_shape=False))
    f = 300
    g = torch.Generator().manual_seed(3)
    pos = 0.3 * torch.randn(1, f, 36, generator=g)
    ori = (torch.eye(3).reshape(1, 1, 1, 9) + 0.05 * torch.randn(1, f, 12, 9, generator=g)).reshape(1, f, 108)
    lens = torch.tensor([f])
    sd = util.torch_state_dict(synthetic.synth_rnn_state_dict(seed=0, **flags['weights']), torch.float64)
    want = oracle_rnn.rnn_forward(oracle_rnn.RnnConfig(**flags['cfg']), sd, oracle_smpl, pos.double(), ori.double(), lens)
    net = build_rnn(smpl_npz, flags, precision, dev)
    z = torch.zeros(1, 12, 3)
    with torch.no_grad():
        out = net(util.DuckBatch(pos, ori, z.unsqueeze(-1).repeat(1, 1, 1, 3), z, lens).to(dev))
    rad = max(util.max_joint_angle_err(out['pose_hat'].cpu().numpy(), want['pose_hat'].numpy()),
              util.max_joint_angle_err(out['root_ori_hat'].cpu().numpy(), want['root_ori_hat'].numpy()))
    util.report('rnn_stream', precision=PNAME[precision], frames=f, rad=rad, launches=net._ctx.last_launch_count)
    assert rad <= (2e-5 if precision == native.PRECISION_FP32 else 1e-4), rad


def test_persistent_stream_with_state_carry_and_padding(dev, smpl_npz, oracle_smpl):
    """The persistent single-stream kernel (B = 1, fp16 mode): two chunks with the LSTM state carried across them
    (evaluate_real.py:61), the second one padded (len < F), uni- and bidirectional; checked against the oracle."""
    for bidir, hidden, layers in ((True, 256, 2), (False, 128, 3)):
        flags = dict(cfg=dict(n_markers=12, hidden_size=hidden, num_layers=layers, bidirectional=bidir),
                     weights=dict(n_markers=12, hidden_size=hidden, num_layers=layers, bidirectional=bidir, estimate_shape=False))
        sd = util.torch_state_dict(synthetic.synth_rnn_state_dict(seed=0, **flags['weights']), torch.float64)
        net = build_rnn(smpl_npz, flags, native.PRECISION_FP16, dev)
        g = torch.Generator().manual_seed(11)
        state = None
        z = torch.zeros(1, 12, 3)
        for c, (f, n_live) in enumerate(((96, 96), (80, 61))):
            pos = 0.3 * torch.randn(1, f, 36, generator=g)
            ori = (torch.eye(3).reshape(1, 1, 1, 9) + 0.05 * torch.randn(1, f, 12, 9, generator=g)).reshape(1, f, 108)
            lens = torch.tensor([n_live])
            want = oracle_rnn.rnn_forward(oracle_rnn.RnnConfig(**flags['cfg']), sd, oracle_smpl, pos.double(), ori.double(), lens,
                                          init_state=state)
            state = want['final_state']
            with torch.no_grad():
                out = net(util.DuckBatch(pos, ori, z.unsqueeze(-1).repeat(1, 1, 1, 3), z, lens).to(dev), is_new_sequence=(c == 0))
            live = util.valid_frame_mask(lens.numpy(), f)
            rad = max(util.max_joint_angle_err(out['pose_hat'].cpu().numpy()[live], want['pose_hat'].numpy()[live]),
                      util.max_joint_angle_err(out['root_ori_hat'].cpu().numpy()[live], want['root_ori_hat'].numpy()[live]))
            h_err = float((net.rnn.final_state[0].cpu().double() - state[0]).abs().max())
            c_err = float((net.rnn.final_state[1].cpu().double() - state[1]).abs().max())
            util.report('rnn_persistent', bidirectional=bidir, chunk=c, rad=rad, h_err=h_err, c_err=c_err,
                        launches=net._ctx.last_launch_count)
            assert net._ctx.last_launch_count < 40, 'the persistent path must not launch once per time step'
            assert rad <= 1e-4, rad
            assert h_err <= 2e-3 and c_err <= 5e-3, (h_err, c_err)


def test_birnn_large_ragged_batch_equals_subsets(dev, smpl_npz):
    """More 128-row tiles than SMs, ragged lengths, shape head with CTA-local scratch activations: the first layer of
    ``to_shape`` masks padded frames (``pad_packed_sequence`` zeros, models.py:300-309) by LOGICAL row -- a CTA's later tiles
    write their activations to the scratch rows of its first tile, which once also selected the sequence length.  Windows
    are independent, so every window of the big batch must equal the same window run in a small batch (bit for bit)."""
    flags = util.RNN_CASES['rnn_bi12_shape_fk']
    flags = dict(cfg=dict(flags['cfg'], hidden_size=128), weights=dict(flags['weights'], hidden_size=128))
    net = build_rnn(smpl_npz, flags, native.PRECISION_FP16, dev)
    b, f = 1300, 16                                   # 20800 rows = 163 row tiles > 148 SMs
    g = torch.Generator().manual_seed(5)
    pos = 0.3 * torch.randn(b, f, 36, generator=g)
    ori = (torch.eye(3).reshape(1, 1, 1, 9) + 0.05 * torch.randn(b, f, 12, 9, generator=g)).reshape(b, f, 108)
    lens = torch.randint(1, f + 1, (b,), generator=g).to(torch.int32)
    z = torch.zeros(b, 12, 3)
    mk = lambda sl: util.DuckBatch(pos[sl], ori[sl], z[sl].unsqueeze(-1).repeat(1, 1, 1, 3), z[sl], lens[sl]).to(dev)
    with torch.no_grad():
        full = net(mk(slice(0, b)), is_new_sequence=True)
        full = {k: full[k].clone() for k in ('pose_hat', 'shape_hat', 'joints_hat')}
        worst = 0.0
        for lo in (0, 640, 1240):                     # first tiles of a CTA, a later round, the tail
            part = net(mk(slice(lo, lo + 60)), is_new_sequence=True)
            live = torch.from_numpy(util.valid_frame_mask(lens[lo:lo + 60].numpy(), f)).to(dev)
            for k in ('pose_hat', 'shape_hat', 'joints_hat'):
                d = (full[k][lo:lo + 60] - part[k])[live].abs().max().item()
                worst = max(worst, d)
    util.report('birnn_large_ragged', worst_abs_diff=worst, rows=b * f)
    assert worst == 0.0, worst
