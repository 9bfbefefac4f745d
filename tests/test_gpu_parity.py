"""Parity of the CUDA path against the oracle and the golden vectors (run on the B200 box: -m gpu).

Everything here calls through the C ABI (``empose_b200.lib`` -> ``libempose_b200.so``).  Tolerances:
the task's bar is <= 1e-4 rad per-joint rotation and <= 0.1 mm joint position against the reference
path on identical inputs; the FP32 executor is held to a much tighter bound.
"""
import numpy as np
import pytest
import torch

from empose_b200 import lib as native
from empose_b200 import synthetic
from oracle import ief as oracle_ief

import util

pytestmark = pytest.mark.gpu

PARITY_RAD = 1e-4      # north_star tolerance, per-joint axis-angle
PARITY_MM = 0.1        # north_star tolerance, joint position
PRECISIONS = [native.PRECISION_FP32, native.PRECISION_TF32, native.PRECISION_FP16]
PNAME = {native.PRECISION_FP32: 'fp32', native.PRECISION_TF32: 'tf32', native.PRECISION_FP16: 'fp16'}


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'these tests need the B200'
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device('cuda:0')


# ---------------------------------------------------------------------------------------------------------------------
# the GEMM engine
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', PRECISIONS, ids=PNAME.get)
@pytest.mark.parametrize('m,n,k', [(128, 256, 32), (128, 16, 8), (300, 512, 296), (257, 80, 512), (1000, 192, 256),
                                   (4096, 2048, 656), (19, 66, 144)])
def test_gemm_engine(dev, precision, m, n, k):
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, generator=g).to(dev)
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).to(dev)
    bias = torch.randn(n, generator=g).to(dev)
    got = native.gemm_selftest(a, w, bias, precision)
    want = (a.double() @ w.double().T + bias.double()).float()
    err = (got - want).abs().max().item()
    # tf32 inputs carry 2^-11 relative rounding (truncation in the worst case 2^-10): |a||w| ~ 1 per term, sqrt(k) growth
    # (fp16 mode: operands AND, for N % 32 == 0, the outputs are rounded to fp16)
    tol = 2e-5 if precision == native.PRECISION_FP32 else 1e-2
    assert err < tol, 'max abs err %g' % err
    if precision == native.PRECISION_TF32:          # same operands pre-rounded: only accumulation order differs
        r = lambda t: (t.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)
        want_r = (r(a).double() @ r(w).double().T + bias.double()).float()
        assert (native.gemm_selftest(r(a), r(w), bias, precision) - want_r).abs().max().item() < 2e-5


# ---------------------------------------------------------------------------------------------------------------------
# SMPL sub-model -> sensors (empose_sensor_project)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', PRECISIONS, ids=PNAME.get)
def test_sensor_projection_matches_oracle(dev, smpl_npz, oracle_smpl, topology, precision):
    net = util.build_module(smpl_npz, precision=precision, device=dev)
    ctx = net.native_context(dev)
    p = synthetic.synth_window_params(5, 7, seed=4, offsets=True)
    r = 35
    t = lambda a: torch.from_numpy(np.asarray(a))
    poses = t(p['poses']).reshape(r, 66)
    shapes = t(p['shapes']).unsqueeze(1).repeat(1, 7, 1).reshape(r, 10)
    off_r = t(p['offset_r']).unsqueeze(1).repeat(1, 7, 1, 1, 1).reshape(r, 12, 3, 3)
    off_t = t(p['offset_t']).unsqueeze(1).repeat(1, 7, 1, 1).reshape(r, 12, 3)
    pos, ori, joints = ctx.sensor_project(poses.to(dev), shapes.to(dev), off_r.to(dev), off_t.to(dev))
    with torch.no_grad():
        o_pos, o_ori, o_j = oracle_ief.project_sensors(oracle_smpl, topology, poses.double(), shapes.double(),
                                                       off_r.double(), off_t.double())
    util.report('sensor_project', precision=PNAME[precision], joints_mm=util.max_joint_pos_err_mm(joints.cpu().numpy(), o_j.numpy()),
                pos_mm=util.max_joint_pos_err_mm(pos.cpu().numpy(), o_pos.numpy()),
                ori=float(np.abs(ori.cpu().numpy() - o_ori.numpy()).max()))
    assert util.max_joint_pos_err_mm(joints.cpu().numpy(), o_j.numpy()) < 0.005
    assert util.max_joint_pos_err_mm(pos.cpu().numpy(), o_pos.numpy()) < 0.01
    np.testing.assert_allclose(ori.cpu().numpy(), o_ori.numpy(), atol=2e-4, rtol=0)

    # golden vectors from the reference's own SMPLLayer + VirtualMarkerHelper
    gold = util.load_golden('smpl_sensors')
    gp = t(gold['poses']).reshape(-1, 66)
    f = gp.shape[0]
    pos, ori, _ = ctx.sensor_project(gp.to(dev), t(gold['shapes']).repeat(f, 1).to(dev),
                                     t(gold['offset_r']).repeat(f, 1, 1, 1).to(dev), t(gold['offset_t']).repeat(f, 1, 1).to(dev))
    assert util.max_joint_pos_err_mm(pos.cpu().numpy(), gold['sensor_pos'].reshape(f, 12, 3)) < 0.01
    np.testing.assert_allclose(ori.cpu().numpy(), gold['sensor_ori'].reshape(f, 12, 3, 3), atol=2e-4, rtol=0)


# ---------------------------------------------------------------------------------------------------------------------
# full-mesh SMPLLayer.forward (BASELINE config 1)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', PRECISIONS, ids=PNAME.get)
def test_smpl_layer_full_mesh(dev, smpl_npz, oracle_smpl, precision):
    from empose_b200.bodymodels.smpl import SMPLLayer
    from oracle import smplh_lbs
    layer = SMPLLayer(smpl_npz).to(device=dev, dtype=torch.float32)
    layer.precision = precision
    # known answer: zero pose and shape -> template mesh and regressed rest joints
    v0, j0 = layer(torch.zeros(1, 63, device=dev), torch.zeros(10, device=dev))
    np.testing.assert_allclose(v0[0].cpu().numpy(), oracle_smpl.v_template.numpy(), atol=2e-6, rtol=0)
    np.testing.assert_allclose(j0[0].cpu().numpy(), (oracle_smpl.j_regressor @ oracle_smpl.v_template).numpy(), atol=2e-6, rtol=0)
    # golden vectors from the reference's own SMPLLayer (tests/golden/smpl_sensors.npz)
    gold = util.load_golden('smpl_sensors')
    poses = torch.from_numpy(gold['poses']).reshape(-1, 66).to(dev)
    verts, joints = layer(poses_body=poses[:, 3:], betas=torch.from_numpy(gold['shapes']).to(dev), poses_root=poses[:, :3])
    assert verts.shape == (2, 6890, 3) and joints.shape == (2, 52, 3)
    assert util.max_joint_pos_err_mm(verts.cpu().numpy(), gold['verts']) < 0.01
    assert util.max_joint_pos_err_mm(joints.cpu().numpy(), gold['joints']) < 0.005
    # random poses with translation, more frames than one slab
    n = 2500
    g = torch.Generator().manual_seed(5)
    pb, pr = 0.2 * torch.randn(n, 63, generator=g), 0.2 * torch.randn(n, 3, generator=g)
    be, tr = torch.randn(n, 10, generator=g), torch.randn(n, 3, generator=g)
    verts, joints = layer(pb.to(dev), be.to(dev), poses_root=pr.to(dev), trans=tr.to(dev), window_size=1000)
    sel = [0, 1, 2047, 2048, 2499]
    hands = torch.zeros(len(sel), 90, dtype=torch.float64)
    ov, oj = smplh_lbs.lbs(oracle_smpl, torch.cat([pr[sel].double(), pb[sel].double(), hands], dim=1), be[sel].double(), tr[sel].double())
    util.report('smpl_full', precision=PNAME[precision], verts_mm=util.max_joint_pos_err_mm(verts[sel].cpu().numpy(), ov.numpy()),
                joints_mm=util.max_joint_pos_err_mm(joints[sel].cpu().numpy(), oj.numpy()))
    assert util.max_joint_pos_err_mm(verts[sel].cpu().numpy(), ov.numpy()) < 0.01
    assert util.max_joint_pos_err_mm(joints[sel].cpu().numpy(), oj.numpy()) < 0.005


# ---------------------------------------------------------------------------------------------------------------------
# the whole loop vs golden outputs of the unmodified reference
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', PRECISIONS, ids=PNAME.get)
@pytest.mark.parametrize('name', sorted(util.GOLDEN_CASES))
def test_ief_matches_reference_golden(dev, smpl_npz, name, precision):
    flags = util.GOLDEN_CASES[name]
    gold = util.load_golden(name)
    net = util.build_module(smpl_npz, precision=precision, device=dev, **flags)
    rad_tol, mm_tol = (2e-5, 0.02) if precision == native.PRECISION_FP32 else (PARITY_RAD, PARITY_MM)
    c = 0
    while ('c%d_pose_hat' % c) in gold:
        inp = util.chunk_inputs(gold, c)
        batch = util.DuckBatch(**inp).to(dev)
        with torch.no_grad():
            out = net(batch, is_new_sequence=(c == 0))
        tag = 'c%d_' % c
        live = util.valid_frame_mask(gold[tag + 'seq_lengths'], inp['marker_pos'].shape[1])
        got = {k: v.cpu().numpy() for k, v in out.items()}
        util.report('golden', case=name, chunk=c, precision=PNAME[precision],
                    rad=max(util.max_joint_angle_err(got['pose_hat'][live], gold[tag + 'pose_hat'][live]),
                            util.max_joint_angle_err(got['root_ori_hat'][live], gold[tag + 'root_ori_hat'][live])),
                    mm=util.max_joint_pos_err_mm(got['joints_hat'][live], gold[tag + 'joints_hat'][live]))
        assert util.max_joint_angle_err(got['pose_hat'][live], gold[tag + 'pose_hat'][live]) <= rad_tol
        assert util.max_joint_angle_err(got['root_ori_hat'][live], gold[tag + 'root_ori_hat'][live]) <= rad_tol
        assert util.max_joint_pos_err_mm(got['joints_hat'][live], gold[tag + 'joints_hat'][live]) <= mm_tol
        np.testing.assert_allclose(got['shape_hat'][live], gold[tag + 'shape_hat'][live], atol=rad_tol * 5, rtol=0)
        # all N+1 iterates (what IterativeErrorFeedback.backward consumes)
        hist = {'pose_hat_history': net.pose_hat_history, 'shape_hat_history': net.shape_hat_history,
                'joints_hat_history': net.joints_hat_history, 'markers_hat_history': net.markers_hat_history,
                'markers_ori_hat_history': net.markers_ori_hat_history}
        for k, lst in hist.items():
            g = np.stack([h.cpu().numpy() for h in lst])
            assert g.shape == gold[tag + k].shape, k
            tol = 5e-4 if 'ori' in k else rad_tol * 5
            np.testing.assert_allclose(g[:, live], gold[tag + k][:, live], atol=tol, rtol=0, err_msg=k)
        if flags['rnn_init']:
            np.testing.assert_allclose(net.rnn.final_state[0].cpu().numpy(), gold[tag + 'final_h'], atol=5e-6 if precision == native.PRECISION_FP32 else 2e-3, rtol=0)
        c += 1


# ---------------------------------------------------------------------------------------------------------------------
# vs the oracle on fresh seeded inputs: ragged lengths, dropped sensors, offsets, both sensor counts
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', PRECISIONS, ids=PNAME.get)
@pytest.mark.parametrize('n_markers,rnn_init,n_iter', [(12, True, 4), (6, True, 2), (12, False, 4)])
def test_ief_matches_oracle(dev, smpl_npz, oracle_smpl, topology, precision, n_markers, rnn_init, n_iter):
    b, f = 6, 32
    params = synthetic.synth_window_params(b, f, seed=31 + n_markers, ragged=True, offsets=True, drop_rate=0.05)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=5)
    cfg = oracle_ief.IefConfig(n_markers=n_markers, num_iterations=n_iter, rnn_init=rnn_init)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=n_markers, rnn_init=rnn_init))
    want = oracle_ief.ief_forward(cfg, sd, oracle_smpl, topology, **inp)
    net = util.build_module(smpl_npz, n_markers=n_markers, num_iterations=n_iter, rnn_init=rnn_init,
                            precision=precision, device=dev)
    with torch.no_grad():
        out = net(util.DuckBatch(**inp).to(dev))
    live = util.valid_frame_mask(params['seq_lengths'], f)
    rad_tol, mm_tol = (2e-5, 0.02) if precision == native.PRECISION_FP32 else (PARITY_RAD, PARITY_MM)
    pose = torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1).cpu().numpy()
    want_pose = torch.cat([want['root_ori_hat'], want['pose_hat']], dim=-1).numpy()
    rad = util.max_joint_angle_err(pose[live], want_pose[live])
    mm = util.max_joint_pos_err_mm(out['joints_hat'].cpu().numpy()[live], want['joints_hat'].numpy()[live])
    util.report('oracle', n_markers=n_markers, rnn_init=rnn_init, n_iter=n_iter, precision=PNAME[precision], rad=rad, mm=mm)
    assert rad <= rad_tol
    assert mm <= mm_tol


def test_window_shape_is_shared_and_batch_shards_are_independent(dev, smpl_npz):
    """Size-independent properties at a larger size: one shape per window (models.py:529-535) and
    window independence (what lets inference shard over GPUs without a collective)."""
    net = util.build_module(smpl_npz, precision=native.PRECISION_FP16, device=dev)
    ctx = net.native_context(dev)
    b, f = 512, 32
    p = synthetic.synth_window_params(b, f, seed=77, offsets=True)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    poses = t(p['poses']).reshape(b * f, 66)
    shapes = t(p['shapes']).unsqueeze(1).repeat(1, f, 1).reshape(b * f, 10)
    off_r = t(p['offset_r']).unsqueeze(1).repeat(1, f, 1, 1, 1).reshape(b * f, 12, 3, 3)
    off_t = t(p['offset_t']).unsqueeze(1).repeat(1, f, 1, 1).reshape(b * f, 12, 3)
    pos, ori, _ = ctx.sensor_project(poses, shapes, off_r, off_t)
    mpos = (pos + 0.01 * torch.randn(pos.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(3))).reshape(b, f, 36)
    mori = ori.reshape(b, f, 108)
    lengths = t(p['seq_lengths'])
    full = ctx.forward(mpos, mori, t(p['offset_r']), t(p['offset_t']), lengths, want_history=False)
    assert torch.isfinite(full['pose']).all() and torch.isfinite(full['joints']).all()
    assert (full['shape'] - full['shape'][:, :1]).abs().max().item() == 0.0
    half = ctx.forward(mpos[256:], mori[256:], t(p['offset_r'])[256:], t(p['offset_t'])[256:], lengths[256:], want_history=False)
    assert torch.equal(half['pose'], full['pose'][256:]) and torch.equal(half['joints'], full['joints'][256:])


def test_baseline_config3_full_size(dev, smpl_npz, oracle_smpl, topology):
    """BASELINE.json configs[2] at its full size -- LGD-RNN, 12 sensors, N=4, 4096 windows x 32 frames -- through
    size-independent properties: (i) eight windows picked across the batch equal, bit for bit, a run of those eight alone
    (window independence, i.e. what multi-GPU sharding relies on), (ii) those eight match the CPU oracle within the parity bar,
    (iii) every window has ONE shape (models.py:529-535), (iv) the host-buffer entry point returns the same bits."""
    b, f = 4096, 32
    net = util.build_module(smpl_npz, precision=native.PRECISION_FP16, device=dev)
    ctx = net.native_context(dev)
    params = synthetic.synth_window_params(b, f, seed=303, ragged=True, offsets=True)
    pick = np.array([0, 1, 777, 2047, 2048, 3001, 4094, 4095])
    sub_params = {k: (None if v is None else v[pick]) for k, v in params.items()}
    sub_inp = util.oracle_inputs_from_params(oracle_smpl, topology, sub_params, seed=303)
    # the full batch: measurements of every window through the CUDA projection, the eight picked ones replaced by the
    # oracle-made ones so that both runs and the oracle see identical inputs
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    r = b * f
    pos, ori, _ = ctx.sensor_project(t(params['poses']).reshape(r, 66), t(params['shapes']).unsqueeze(1).repeat(1, f, 1).reshape(r, 10),
                                     t(params['offset_r']).unsqueeze(1).repeat(1, f, 1, 1, 1).reshape(r, 12, 3, 3),
                                     t(params['offset_t']).unsqueeze(1).repeat(1, f, 1, 1).reshape(r, 12, 3))
    g = torch.Generator(device=dev).manual_seed(5)
    mpos = (pos + 0.01 * torch.randn(pos.shape, device=dev, generator=g)).reshape(b, f, 36)
    mori = ori.reshape(b, f, 108).clone()
    mpos[pick] = sub_inp['marker_pos'].to(dev)
    mori[pick] = sub_inp['marker_oris'].to(dev)
    lens, off_r, off_t = t(params['seq_lengths']), t(params['offset_r']), t(params['offset_t'])
    full = ctx.forward(mpos, mori, off_r, off_t, lens, want_history=False)
    alone = ctx.forward(mpos[pick], mori[pick], off_r[pick], off_t[pick], lens[pick], want_history=False)
    for k in ('pose', 'shape', 'joints'):
        assert torch.equal(full[k][pick], alone[k]), k                                              # (i)
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=4, rnn_init=True)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True))
    want = oracle_ief.ief_forward(cfg, sd, oracle_smpl, topology, **sub_inp)
    live = util.valid_frame_mask(sub_params['seq_lengths'], f)
    want_pose = torch.cat([want['root_ori_hat'], want['pose_hat']], dim=-1).numpy()
    rad = util.max_joint_angle_err(alone['pose'].cpu().numpy()[live], want_pose[live])
    mm = util.max_joint_pos_err_mm(alone['joints'].cpu().numpy()[live], want['joints_hat'].numpy()[live])
    util.report('config3_full', rad=rad, mm=mm, windows=b)
    assert rad <= PARITY_RAD and mm <= PARITY_MM, (rad, mm)                                         # (ii)
    assert (full['shape'] - full['shape'][:, :1]).abs().max().item() == 0.0                        # (iii)
    assert torch.isfinite(full['pose']).all() and torch.isfinite(full['joints']).all()
    pin = lambda x: x.cpu().contiguous().pin_memory()
    host = ctx.forward_host(pin(mpos), pin(mori), pin(off_r), pin(off_t), lens.cpu())
    assert torch.equal(host['pose'], full['pose'].cpu()) and torch.equal(host['joints'], full['joints'].cpu())   # (iv)


def test_host_buffer_entry_point_matches_device_entry_point(dev, smpl_npz, oracle_smpl, topology):
    net = util.build_module(smpl_npz, precision=native.PRECISION_FP16, device=dev)
    ctx = net.native_context(dev)
    params = synthetic.synth_window_params(4, 16, seed=9, ragged=True, offsets=True)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=2)
    d = {k: (v.to(dev) if v is not None else None) for k, v in inp.items()}
    a = ctx.forward(d['marker_pos'], d['marker_oris'], d['offset_r'], d['offset_t'], d['seq_lengths'], want_history=False)
    h = ctx.forward_host(inp['marker_pos'], inp['marker_oris'], inp['offset_r'], inp['offset_t'], inp['seq_lengths'])
    assert torch.equal(a['pose'].cpu(), h['pose']) and torch.equal(a['joints'].cpu(), h['joints'])
    assert torch.equal(a['lstm_state'].cpu(), h['lstm_state'])


def test_pipelined_host_entry_point_equals_one_pass(dev, smpl_npz):
    """empose_ief_forward_host cuts batches of >= 8192 windows into sub-batches whose PCIe copies overlap the compute of
    their neighbours; windows are independent, so the result must be bit-identical to the single device pass (inputs,
    offsets, ragged lengths, masks and the carried LSTM state all cross the sub-batch boundary)."""
    net = util.build_module(smpl_npz, num_iterations=2, precision=native.PRECISION_FP16, device=dev)
    ctx = net.native_context(dev)
    b, f = 8200, 2                                   # sub-batches of 4100 windows
    g = torch.Generator().manual_seed(9)
    p = synthetic.synth_window_params(b, f, seed=91, ragged=True, offsets=True)
    pos = 0.3 * torch.randn(b, f, 36, generator=g)
    eye = torch.eye(3).reshape(1, 1, 1, 9)
    ori = (eye + 0.05 * torch.randn(b, f, 12, 9, generator=g)).reshape(b, f, 108)
    masks = (torch.rand(b, f, 12, generator=g) > 0.02).float()
    state = 0.1 * torch.randn(2, 2, b, 512, generator=g)
    off_r, off_t, lens = torch.from_numpy(p['offset_r']), torch.from_numpy(p['offset_t']), torch.from_numpy(p['seq_lengths'])
    a = ctx.forward(pos.to(dev), ori.to(dev), off_r.to(dev), off_t.to(dev), lens.to(dev), marker_masks=masks.to(dev),
                    lstm_state=state.to(dev), is_new_sequence=False, want_history=False)
    pin = lambda t: t.contiguous().pin_memory()
    h = ctx.forward_host(pin(pos), pin(ori), pin(off_r), pin(off_t), lens, marker_masks=pin(masks), lstm_state=state,
                         is_new_sequence=False)
    for k in ('pose', 'shape', 'joints', 'lstm_state'):
        assert torch.equal(a[k].cpu(), h[k]), k


def test_streamed_requests_equal_blocking_calls(dev, smpl_npz):
    """empose_ief_submit_host / empose_ief_wait_host: five different requests streamed through two in-flight slots (the
    upload of one and the download of another run under the pass of a third) return exactly what one blocking
    empose_ief_forward_host call per request returns."""
    net = util.build_module(smpl_npz, num_iterations=2, precision=native.PRECISION_FP16, device=dev)
    ctx = net.native_context(dev)
    b, f = 300, 4
    pin = lambda t: t.contiguous().pin_memory()
    reqs_in = []
    for k in range(5):
        g = torch.Generator().manual_seed(100 + k)
        p = synthetic.synth_window_params(b, f, seed=200 + k, ragged=True, offsets=True)
        pos = 0.3 * torch.randn(b, f, 36, generator=g)
        ori = (torch.eye(3).reshape(1, 1, 1, 9) + 0.05 * torch.randn(b, f, 12, 9, generator=g)).reshape(b, f, 108)
        reqs_in.append((pin(pos), pin(ori), pin(torch.from_numpy(p['offset_r'])), pin(torch.from_numpy(p['offset_t'])),
                        pin(torch.from_numpy(p['seq_lengths']).to(torch.int32))))
    want = [ctx.forward_host(*r, want_state=False) for r in reqs_in]
    got, inflight, outs = [None] * 5, [None, None], [None, None]
    for k, r in enumerate(reqs_in):
        slot = k % 2
        if inflight[slot] is not None:
            j, req = inflight[slot]
            done = ctx.wait_host(req)
            got[j] = {n: done[n].clone() for n in ('pose', 'shape', 'joints')}
            outs[slot] = done
        inflight[slot] = (k, ctx.submit_host(slot, *r, out=outs[slot]))
    for slot in range(2):
        j, req = inflight[slot]
        done = ctx.wait_host(req)
        got[j] = {n: done[n].clone() for n in ('pose', 'shape', 'joints')}
    for k in range(5):
        for n in ('pose', 'shape', 'joints'):
            assert torch.equal(got[k][n], want[k][n]), (k, n)
    with pytest.raises(native.EmposeError):
        ctx.submit_host(0, reqs_in[0][0].clone(), *reqs_in[0][1:])          # not pinned


# ---------------------------------------------------------------------------------------------------------------------
# other configurations of the reference and edge cases
# ---------------------------------------------------------------------------------------------------------------------
def _compare_with_oracle(dev, smpl_npz, oracle_smpl, topology, b, f, precision, seed, cfg_kwargs, module_kwargs, ragged=True,
                         drop_rate=0.0, tag=''):
    params = synthetic.synth_window_params(b, f, seed=seed, ragged=ragged, offsets=True, drop_rate=drop_rate)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=seed)
    cfg = oracle_ief.IefConfig(**cfg_kwargs)
    sd = util.torch_state_dict(synthetic.synth_state_dict(
        seed=0, n_markers=cfg.n_markers, rnn_init=cfg.rnn_init, hidden_size=cfg.hidden_size, num_layers=cfg.num_layers,
        rnn_hidden_size=cfg.rnn_hidden_size, use_gradient=cfg.use_gradient, batch_norm=not cfg.no_batch_norm,
        use_marker_pos=cfg.use_marker_pos, use_marker_ori=cfg.use_marker_ori))
    want = oracle_ief.ief_forward(cfg, sd, oracle_smpl, topology, **inp)
    net = util.build_module(smpl_npz, n_markers=cfg.n_markers, num_iterations=cfg.num_iterations, rnn_init=cfg.rnn_init,
                            precision=precision, device=dev, hidden_size=cfg.hidden_size, **module_kwargs)
    with torch.no_grad():
        out = net(util.DuckBatch(**inp).to(dev))
    live = util.valid_frame_mask(params['seq_lengths'], f)
    pose = torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1).cpu().numpy()
    want_pose = torch.cat([want['root_ori_hat'], want['pose_hat']], dim=-1).numpy()
    rad = util.max_joint_angle_err(pose[live], want_pose[live])
    mm = util.max_joint_pos_err_mm(out['joints_hat'].cpu().numpy()[live], want['joints_hat'].numpy()[live])
    util.report('oracle_' + tag, precision=PNAME[precision], b=b, f=f, rad=rad, mm=mm)
    rad_tol, mm_tol = (2e-5, 0.02) if precision == native.PRECISION_FP32 else (PARITY_RAD, PARITY_MM)
    assert rad <= rad_tol and mm <= mm_tol, (rad, mm)
    return net, inp, want


def test_baseline_config2_lgd_no_rnn_256_windows(dev, smpl_npz, oracle_smpl, topology):
    """BASELINE.json configs[1]: LGD without RNN, 12 sensors, N=4, 256 windows x 32 frames, every frame checked."""
    _compare_with_oracle(dev, smpl_npz, oracle_smpl, topology, 256, 32, native.PRECISION_FP16, seed=61,
                         cfg_kwargs=dict(n_markers=12, num_iterations=4, rnn_init=False), module_kwargs={}, tag='config2')


def test_skip_connections_without_batch_norm(dev, smpl_npz, oracle_smpl, topology):
    """Not a released configuration.  Exact in FP32 mode; without BatchNorm's damping the TF32 rounding of six chained
    512-wide layers reaches ~5e-4 rad, so the TF32 bar is only claimed for the BatchNorm models."""
    precision = native.PRECISION_FP32
    _compare_with_oracle(dev, smpl_npz, oracle_smpl, topology, 5, 16, precision, seed=62,
                         cfg_kwargs=dict(n_markers=12, num_iterations=2, rnn_init=True, skip_connections=True, no_batch_norm=True),
                         module_kwargs=dict(m_skip_connections=True, m_no_batch_norm=True), tag='skip_nobn')


@pytest.mark.parametrize('precision', PRECISIONS, ids=PNAME.get)
def test_without_gradient_feature_and_position_only(dev, smpl_npz, oracle_smpl, topology, precision):
    _compare_with_oracle(dev, smpl_npz, oracle_smpl, topology, 4, 8, precision, seed=63,
                         cfg_kwargs=dict(n_markers=12, num_iterations=3, rnn_init=True, use_gradient=False),
                         module_kwargs=dict(m_use_gradient=False), tag='nograd')
    _compare_with_oracle(dev, smpl_npz, oracle_smpl, topology, 4, 8, precision, seed=64,
                         cfg_kwargs=dict(n_markers=6, num_iterations=2, rnn_init=False, use_marker_ori=False),
                         module_kwargs=dict(use_marker_ori=False), tag='posonly')


def test_smallest_and_odd_shapes(dev, smpl_npz, oracle_smpl, topology):
    for b, f in ((1, 1), (1, 3), (3, 1), (7, 5), (129, 3)):
        _compare_with_oracle(dev, smpl_npz, oracle_smpl, topology, b, f, native.PRECISION_FP32, seed=70 + b + f,
                             cfg_kwargs=dict(n_markers=12, num_iterations=2, rnn_init=True), module_kwargs={}, tag='odd')


def test_streaming_chunks_and_window_size(dev, smpl_npz, oracle_smpl, topology):
    """evaluate_real.py feeds one sequence in 256-frame chunks with is_new_sequence=(c == 0) (LSTM state carried);
    models.py:146-159 offers the same through window_size.  Both must equal the oracle run chunk by chunk."""
    f_total, chunk = 600, 256
    params = synthetic.synth_window_params(1, f_total, seed=81, offsets=True)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=81)
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=2, rnn_init=True)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True))
    net = util.build_module(smpl_npz, num_iterations=2, precision=native.PRECISION_FP32, device=dev)
    state, want_pose, got_pose = None, [], []
    for c, sf in enumerate(range(0, f_total, chunk)):
        ef = min(sf + chunk, f_total)
        sl = lambda t: t[:, sf:ef]
        part = dict(marker_pos=sl(inp['marker_pos']), marker_oris=sl(inp['marker_oris']), offset_r=inp['offset_r'],
                    offset_t=inp['offset_t'], seq_lengths=torch.tensor([ef - sf]), marker_masks=None)
        want = oracle_ief.ief_forward(cfg, sd, oracle_smpl, topology, init_state=state, **part)
        state = want['final_state']
        want_pose.append(torch.cat([want['root_ori_hat'], want['pose_hat']], dim=-1))
        with torch.no_grad():
            out = net(util.DuckBatch(**part).to(dev), is_new_sequence=(c == 0))
        got_pose.append(torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1).cpu())
    want_pose, got_pose = torch.cat(want_pose, dim=1).numpy(), torch.cat(got_pose, dim=1).numpy()
    assert util.max_joint_angle_err(got_pose, want_pose) <= 2e-5
    # the same split through window_size on one call
    whole = util.DuckBatch(**{**inp, 'seq_lengths': torch.tensor([f_total])}).to(dev)
    with torch.no_grad():
        out = net(whole, window_size=chunk)
    pose = torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1).cpu().numpy()
    assert pose.shape == (1, f_total, 66) and len(net.pose_hat_history) == 3
    assert util.max_joint_angle_err(pose, want_pose) <= 2e-5


def test_validation_loss_matches_reference_formula(dev, smpl_npz, oracle_smpl, topology):
    """IterativeErrorFeedback.backward in eval mode (eval/helpers.py:86): loss values from the histories."""
    b, f = 3, 8
    params = synthetic.synth_window_params(b, f, seed=91, ragged=True, offsets=True)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=91)
    net = util.build_module(smpl_npz, precision=native.PRECISION_FP32, device=dev, m_fk_loss=0.1, m_pose_loss_weight=10.0)
    batch = util.DuckBatch(**inp).to(dev)
    batch.poses_root = torch.from_numpy(params['poses'][:, :, :3]).to(dev)
    batch.poses_body = torch.from_numpy(params['poses'][:, :, 3:]).to(dev)
    batch.shapes = torch.from_numpy(params['shapes']).to(dev)
    batch.joints_gt = torch.zeros(b, f, 66, device=dev)
    with torch.no_grad():
        out = net(batch)
        total, vals = net.backward(batch, out)
    # recompute from the oracle's histories with the formula of models.py:648-674
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=4, rnn_init=True)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True))
    want = oracle_ief.ief_forward(cfg, sd, oracle_smpl, topology, **inp)
    lengths = inp['seq_lengths'].double()
    mask = torch.from_numpy(util.valid_frame_mask(params['seq_lengths'], f)).double()
    mm = lambda per_frame: ((per_frame * mask).sum(-1) / lengths).mean()
    gt_pose = torch.from_numpy(params['poses']).double()
    gt_shape = torch.from_numpy(params['shapes']).double().unsqueeze(1).repeat(1, f, 1)
    pose_l = sum(mm((gt_pose - h.double()).abs().mean(-1)) for h in want['history']['pose'])
    shape_l = sum(mm((gt_shape - h.double()).abs().mean(-1)) for h in want['history']['shape'])
    fk_l = 5 * mm(want['joints_hat'].double().reshape(b, f, 22, 3).norm(dim=-1).sum(-1))
    rec_l = sum(mm((h.double().reshape(b, f, 12, 3) - inp['marker_pos'].double().reshape(b, f, 12, 3)).norm(dim=-1).sum(-1))
                for h in want['history']['markers'])
    rec_l = rec_l + sum(mm((h.double().reshape(b, f, 12, 9) - inp['marker_oris'].double().reshape(b, f, 12, 9)).norm(dim=-1).sum(-1))
                        for h in want['history']['markers_ori'])
    want_total = (10.0 * pose_l + 0.1 * fk_l + 1.0 * shape_l + 0.01 * rec_l) / 5
    assert abs(total.item() - want_total.item()) <= 1e-4 * abs(want_total.item())
    assert abs(vals['pose'] - pose_l.item() / 5) <= 1e-5 and abs(vals['reconstruction'] - rec_l.item() / 5) <= 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# the two sub-model kernels (fan form / general) and sensor vertices of irregular valence
# ---------------------------------------------------------------------------------------------------------------------
def _lgd_run(net, dev, inp):
    with torch.no_grad():
        out = net(util.DuckBatch(**inp).to(dev))
    hist = {k: np.stack([h.cpu().numpy() for h in getattr(net, k)]) for k in
            ('pose_hat_history', 'shape_hat_history', 'joints_hat_history', 'markers_hat_history', 'markers_ori_hat_history')}
    return {k: v.cpu().numpy() for k, v in out.items()}, hist


@pytest.mark.parametrize('n_markers', [12, 6])
def test_fan_kernel_equals_general_kernel(dev, smpl_npz, oracle_smpl, topology, n_markers):
    """The production fan-form kernel (csrc/fan_kernel.cu) and the general index-table kernel (csrc/frame_kernels.cu) are two
    implementations of the same arithmetic: all N+1 iterates -- i.e. sensors, joints AND the gradient features that drive
    the updates -- must agree to float32 rounding (exact-arithmetic mode, so nothing else differs)."""
    params = synthetic.synth_window_params(5, 9, seed=21, ragged=True, offsets=True, drop_rate=0.05)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=8)
    net = util.build_module(smpl_npz, n_markers=n_markers, precision=native.PRECISION_FP32, device=dev)
    try:
        native.set_option('main_general', 1)
        out_g, hist_g = _lgd_run(net, dev, inp)
    finally:
        native.set_option('main_general', 0)
    out_f, hist_f = _lgd_run(net, dev, inp)
    live = util.valid_frame_mask(params['seq_lengths'], 9)
    for k in out_f:
        np.testing.assert_allclose(out_f[k][live], out_g[k][live], atol=5e-6, rtol=0, err_msg=k)
    for k in hist_f:
        np.testing.assert_allclose(hist_f[k][:, live], hist_g[k][:, live], atol=5e-5 if 'ori' in k else 5e-6, rtol=0, err_msg=k)
    util.report('fan_vs_general', n_markers=n_markers, pose=float(np.abs(out_f['pose_hat'][live] - out_g['pose_hat'][live]).max()),
                ori=float(np.abs(hist_f['markers_ori_hat_history'][:, live] - hist_g['markers_ori_hat_history'][:, live]).max()))


@pytest.mark.parametrize('precision', [native.PRECISION_FP32, native.PRECISION_FP16], ids=PNAME.get)
@pytest.mark.parametrize('kind', ['mild', 'wild'])
def test_irregular_valence_mesh(dev, asset_dir, kind, precision):
    """Sensor vertices of valence 4..11 (a real SMPL-H mesh is not regular): 'mild' runs the 8-slot / valence <= 7 fan kernel,
    'wild' the 12-slot one; both against the full-mesh oracle, and the general kernel on the same sub-model."""
    from oracle import sensors, smplh_lbs
    npz = synthetic.write_synthetic_smplh(asset_dir, seed=0, irregular=kind)
    osm = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float64)
    topo = sensors.sensor_topology(osm.faces.numpy())
    b, f = 4, 12
    params = synthetic.synth_window_params(b, f, seed=51, ragged=True, offsets=True)
    inp = util.oracle_inputs_from_params(osm, topo, params, seed=6)
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=4, rnn_init=True)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True))
    want = oracle_ief.ief_forward(cfg, sd, osm, topo, **inp)
    net = util.build_module(npz, precision=precision, device=dev)
    assert int(net.smpl.submodel_arrays()['sub.fan_dims'][0]) == 1
    live = util.valid_frame_mask(params['seq_lengths'], f)
    want_pose = torch.cat([want['root_ori_hat'], want['pose_hat']], dim=-1).numpy()
    rad_tol, mm_tol = (2e-5, 0.02) if precision == native.PRECISION_FP32 else (PARITY_RAD, PARITY_MM)
    for general in (0, 1):
        try:
            native.set_option('main_general', general)
            out, _ = _lgd_run(net, dev, inp)
        finally:
            native.set_option('main_general', 0)
        pose = np.concatenate([out['root_ori_hat'], out['pose_hat']], axis=-1)
        rad = util.max_joint_angle_err(pose[live], want_pose[live])
        mm = util.max_joint_pos_err_mm(out['joints_hat'][live], want['joints_hat'].numpy()[live])
        util.report('irregular', mesh=kind, general=general, precision=PNAME[precision], rad=rad, mm=mm)
        assert rad <= rad_tol and mm <= mm_tol, (kind, general, rad, mm)


# ---------------------------------------------------------------------------------------------------------------------
# evidence for the fp16 operand default: adversarial-but-plausible weights and inputs
# ---------------------------------------------------------------------------------------------------------------------
def _wide_bn_state_dict(sd, seed):
    """BatchNorm running variances log-uniform in [1e-4, 1e2] with the Linear in front scaled to match (what training
    produces: a unit with tiny variance has tiny pre-activations), so the network function is the same but the RAW
    weights span six orders of magnitude; the folded ones (what the kernels see) must not care."""
    g = torch.Generator().manual_seed(seed)
    sd = {k: v.clone() for k, v in sd.items()}
    for name in [k for k in sd if k.endswith('running_var')]:
        var = torch.exp(torch.empty_like(sd[name]).uniform_(float(np.log(1e-4)), float(np.log(1e2)), generator=g))
        ratio = torch.sqrt(var / sd[name])
        lin = name.replace('batch_norm.running_var', 'input_to_hidden') if 'batch_norm' in name else \
            '.'.join(name.split('.')[:-2] + [str(int(name.split('.')[-2]) - 1)])
        sd[lin + '.weight'] = sd[lin + '.weight'] * ratio.reshape(-1, 1)
        sd[lin + '.bias'] = sd[lin + '.bias'] * ratio
        mean = name.replace('running_var', 'running_mean')
        sd[mean] = sd[mean] * ratio
        sd[name] = var
    return sd


@pytest.mark.parametrize('precision', [native.PRECISION_FP16, native.PRECISION_TF32], ids=PNAME.get)
def test_wide_batchnorm_statistics(dev, smpl_npz, oracle_smpl, topology, precision):
    """BN running var in [1e-4, 1e2] (folded scale up to 100x / down to 0.1x of the raw weights): parity bar unchanged."""
    params = synthetic.synth_window_params(6, 32, seed=61, ragged=True, offsets=True)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=9)
    net = util.build_module(smpl_npz, precision=precision, device='cpu')
    sd = _wide_bn_state_dict(net.state_dict(), seed=4)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    assert net._effective_precision() == precision            # the folded weights are as tame as before
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=4, rnn_init=True)
    osd = {k: v.cpu() for k, v in sd.items() if not k.startswith('smpl.')}
    want = oracle_ief.ief_forward(cfg, osd, oracle_smpl, topology, **inp)
    with torch.no_grad():
        out = net(util.DuckBatch(**inp).to(dev))
    live = util.valid_frame_mask(params['seq_lengths'], 32)
    pose = torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1).cpu().numpy()
    want_pose = torch.cat([want['root_ori_hat'], want['pose_hat']], dim=-1).numpy()
    rad = util.max_joint_angle_err(pose[live], want_pose[live])
    mm = util.max_joint_pos_err_mm(out['joints_hat'].cpu().numpy()[live], want['joints_hat'].numpy()[live])
    util.report('wide_bn', precision=PNAME[precision], rad=rad, mm=mm)
    assert rad <= PARITY_RAD and mm <= PARITY_MM, (rad, mm)


@pytest.mark.parametrize('precision', PRECISIONS, ids=PNAME.get)
def test_short_sequences_in_long_windows(dev, smpl_npz, oracle_smpl, topology, precision):
    """seq_len << F: the gradient features carry the factor F / len (up to 32 here, loss.py:36-41 folded per frame), the
    largest inputs the iter-MLPs ever see.  Reported per precision; the bar is asserted for the fp32 executor and for the
    guard's choice (``IterativeErrorFeedback.precision_guard``)."""
    b, f = 6, 32
    params = synthetic.synth_window_params(b, f, seed=71, ragged=False, offsets=True)
    params['seq_lengths'] = np.asarray([1, 2, 3, 5, 8, 32], dtype=params['seq_lengths'].dtype)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=10)
    cfg = oracle_ief.IefConfig(n_markers=12, num_iterations=4, rnn_init=True)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=12, rnn_init=True))
    want = oracle_ief.ief_forward(cfg, sd, oracle_smpl, topology, **inp)
    net = util.build_module(smpl_npz, precision=precision, device=dev)
    with torch.no_grad():
        out = net(util.DuckBatch(**inp).to(dev))
    live = util.valid_frame_mask(params['seq_lengths'], f)
    pose = torch.cat([out['root_ori_hat'], out['pose_hat']], dim=-1).cpu().numpy()
    want_pose = torch.cat([want['root_ori_hat'], want['pose_hat']], dim=-1).numpy()
    rad = util.max_joint_angle_err(pose[live], want_pose[live])
    mm = util.max_joint_pos_err_mm(out['joints_hat'].cpu().numpy()[live], want['joints_hat'].numpy()[live])
    gmax = float(np.abs(np.stack([h.cpu().numpy() for h in net.pose_hat_history])[:, live]).max())
    util.report('short_sequences', precision=PNAME[precision], rad=rad, mm=mm, pose_max=gmax)
    rad_tol, mm_tol = (2e-5, 0.02) if precision == native.PRECISION_FP32 else (PARITY_RAD, PARITY_MM)
    assert rad <= rad_tol and mm <= mm_tol, (rad, mm)


def test_fp16_overflow_falls_back_to_tf32(dev, smpl_npz, oracle_smpl, topology):
    """Inputs beyond fp16's range (65504): the fp16 pass returns non-finite values, the guard repeats the pass on tf32 tensor
    cores and returns exactly what an explicit tf32 model returns; with the guard off the caller sees the non-finite result."""
    params = synthetic.synth_window_params(3, 8, seed=81, offsets=True)
    inp = util.oracle_inputs_from_params(oracle_smpl, topology, params, seed=11)
    inp['marker_pos'] = inp['marker_pos'] + 1.0e5
    batch = util.DuckBatch(**inp).to(dev)
    net = util.build_module(smpl_npz, precision=native.PRECISION_FP16, device=dev)
    ref = util.build_module(smpl_npz, precision=native.PRECISION_TF32, device=dev)
    with torch.no_grad():
        out, want = net(batch), ref(batch)
    assert net.precision_fallbacks == 1
    for k in out:
        assert torch.isfinite(out[k]).all() and torch.equal(out[k], want[k]), k
    net.precision_guard = False
    with torch.no_grad():
        raw = net(batch)
    assert not torch.isfinite(raw['pose_hat']).all()
    assert net.precision_fallbacks == 1
