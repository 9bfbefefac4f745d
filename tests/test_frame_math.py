"""The CUDA per-frame math (compiled for the host) vs the oracle: forward values and the hand-derived
reverse pass vs autograd.  No GPU needed; this is what de-risks the VJP before any kernel runs."""
import numpy as np
import pytest
import torch

from empose_b200 import submodel, synthetic
from oracle import ief as oracle_ief
from oracle import sensors

import host_math


@pytest.fixture(scope='module')
def sub(smpl_npz):
    s, _ = submodel.submodel_from_npz(smpl_npz)
    return s


def _case(n, seed, offsets=True):
    p = synthetic.synth_window_params(1, n, seed=seed, offsets=offsets)
    theta = p['poses'][0]
    beta = np.repeat(p['shapes'], n, axis=0) + 0.1 * np.random.RandomState(seed).standard_normal((n, 10)).astype(np.float32)
    off_r = np.repeat(p['offset_r'], n, axis=0)
    off_t = np.repeat(p['offset_t'], n, axis=0)
    return theta, beta, off_r, off_t


def _oracle(oracle_smpl, topology, theta, beta, off_r, off_t):
    t = lambda a: torch.from_numpy(np.asarray(a)).double()
    pose = t(theta).requires_grad_(True)
    shape = t(beta).requires_grad_(True)
    pos, ori, joints = oracle_ief.project_sensors(oracle_smpl, topology, pose, shape, t(off_r), t(off_t))
    return pose, shape, pos, ori, joints


EVAL = {'general': host_math.frame_eval, 'fan': host_math.fan_eval}


@pytest.mark.parametrize('form', ['general', 'fan'])
@pytest.mark.parametrize('use_double', [True, False])
def test_forward_matches_full_mesh_oracle(sub, oracle_smpl, topology, use_double, form):
    theta, beta, off_r, off_t = _case(6, seed=3)
    _, _, pos, ori, joints = _oracle(oracle_smpl, topology, theta, beta, off_r, off_t)
    z = np.zeros
    out = EVAL[form](sub, theta, beta, off_r, off_t, z((6, 12, 3)), z((6, 12, 9)), np.ones(12), np.ones(6),
                     want_grad=False, use_double=use_double)
    tol = 2e-6 if use_double else 5e-6        # sub-model constants are float32; metres
    np.testing.assert_allclose(out['sensor_pos'], pos.detach().numpy(), atol=tol, rtol=0)
    np.testing.assert_allclose(out['joints'], joints.detach().numpy(), atol=tol, rtol=0)
    np.testing.assert_allclose(out['sensor_ori'], ori.detach().numpy(), atol=5e-5 if not use_double else 2e-5, rtol=0)


@pytest.mark.parametrize('form', ['general', 'fan'])
@pytest.mark.parametrize('static_tree', [True, False])
@pytest.mark.parametrize('n_markers', [12, 6])
@pytest.mark.parametrize('use_double', [True, False])
def test_reverse_pass_matches_autograd(sub, oracle_smpl, topology, n_markers, use_double, static_tree, form):
    n = 8
    theta, beta, off_r, off_t = _case(n, seed=5)
    rng = np.random.RandomState(9)
    pose, shape, pos, ori, _ = _oracle(oracle_smpl, topology, theta, beta, off_r, off_t)
    meas_pos = pos.detach().numpy() + 0.01 * rng.standard_normal((n, 12, 3))
    meas_ori = ori.detach().numpy() + 0.05 * rng.standard_normal((n, 12, 3, 3))
    idx = list(range(12)) if n_markers == 12 else list(sensors.S_CONFIG_6)
    active = np.zeros(12, dtype=np.int32)
    active[idx] = 1
    coef = rng.uniform(0.5, 2.0, size=n)
    coef[2] = 0.0
    # per-frame energy, weighted per frame
    dp = pos[:, idx] - torch.from_numpy(meas_pos[:, idx])
    dr = (ori[:, idx] - torch.from_numpy(meas_ori[:, idx])).reshape(n, len(idx), 9)
    energy = (torch.sqrt((dp * dp).sum(-1)).sum(-1) + torch.sqrt((dr * dr).sum(-1)).sum(-1)) * torch.from_numpy(coef)
    g_pose, g_shape = torch.autograd.grad(energy.sum(), [pose, shape])
    out = EVAL[form](sub, theta, beta, off_r, off_t, meas_pos, meas_ori.reshape(n, 12, 9), active, coef,
                     use_double=use_double, static_tree=static_tree)
    scale = float(g_pose.abs().max())
    # float32: the orientation residual differentiates normalised cross products of ~1 cm edges taken from
    # ~0.5 m coordinates, so ~1e-3 relative noise is inherent to single precision (the reference has it too)
    tol = (2e-5 if use_double else 1.5e-3) * max(scale, 1.0)
    np.testing.assert_allclose(out['g_theta'], g_pose.numpy(), atol=tol, rtol=0)
    np.testing.assert_allclose(out['g_beta'], g_shape.numpy(), atol=tol, rtol=0)
    assert np.abs(out['g_theta'][2]).max() == 0.0


def test_rest_pose_known_answer(sub, oracle_smpl):
    """Zero pose and shape: skinned sub-mesh == template vertices, joints == regressed rest joints."""
    z = np.zeros
    eye = np.tile(np.eye(3, dtype=np.float32).reshape(1, 1, 9), (1, 12, 1))
    out = host_math.frame_eval(sub, z((1, 66)), z((1, 10)), eye, z((1, 12, 3)), z((1, 12, 3)), z((1, 12, 9)),
                               np.ones(12), np.ones(1), want_grad=False, use_double=True)
    ids = sub['sub.global_vertex_ids']                          # ring-major, -1 = padding slot of a sensor block
    np.testing.assert_allclose(out['verts'][0][ids >= 0], oracle_smpl.v_template.numpy()[ids[ids >= 0]], atol=1e-6, rtol=0)
    assert np.abs(out['verts'][0][ids < 0]).max() == 0.0
    rest = (oracle_smpl.j_regressor @ oracle_smpl.v_template).numpy()[:22]
    np.testing.assert_allclose(out['joints'][0], rest, atol=1e-6, rtol=0)


@pytest.mark.parametrize('form', ['general', 'fan'])
@pytest.mark.parametrize('static_tree', [True, False])
def test_weighted_reverse_pass_with_joint_term_matches_autograd(sub, oracle_smpl, topology, static_tree, form):
    """Training form of the reverse pass: w_s * sensor residual + w_j * sum_j ||J_j - Jgt_j|| (models.py:657-674)."""
    n = 5
    theta, beta, off_r, off_t = _case(n, seed=7)
    rng = np.random.RandomState(10)
    pose, shape, pos, ori, joints = _oracle(oracle_smpl, topology, theta, beta, off_r, off_t)
    meas_pos = pos.detach().numpy() + 0.01 * rng.standard_normal((n, 12, 3))
    meas_ori = ori.detach().numpy() + 0.05 * rng.standard_normal((n, 12, 3, 3))
    joints_gt = (joints.detach().numpy() + 0.02 * rng.standard_normal((n, 22, 3))).astype(np.float32)
    coef = rng.uniform(0.5, 2.0, size=n)
    w_s, w_j = 0.002, 0.1
    dp = pos - torch.from_numpy(meas_pos)
    dr = (ori - torch.from_numpy(meas_ori)).reshape(n, 12, 9)
    dj = joints - torch.from_numpy(joints_gt).double()
    energy = (w_s * (torch.sqrt((dp * dp).sum(-1)).sum(-1) + torch.sqrt((dr * dr).sum(-1)).sum(-1)) +
              w_j * torch.sqrt((dj * dj).sum(-1)).sum(-1)) * torch.from_numpy(coef)
    g_pose, g_shape = torch.autograd.grad(energy.sum(), [pose, shape])
    out = EVAL[form](sub, theta, beta, off_r, off_t, meas_pos, meas_ori.reshape(n, 12, 9), np.ones(12), coef,
                     use_double=True, static_tree=static_tree, sensor_weight=w_s,
                     joints_gt=joints_gt.reshape(n, 66), joint_weight=w_j)
    tol = 2e-5 * max(float(g_pose.abs().max()), 1.0)
    np.testing.assert_allclose(out['g_theta'], g_pose.numpy(), atol=tol, rtol=0)
    np.testing.assert_allclose(out['g_beta'], g_shape.numpy(), atol=tol, rtol=0)


@pytest.mark.parametrize('kind,force_maxd', [('mild', 0), ('wild', 0), (None, 7)])
@pytest.mark.parametrize('form', ['general', 'fan'])
def test_irregular_valence_meshes(asset_dir, kind, form, force_maxd):
    """Sensor vertices of valence 4..11 (a real SMPL-H mesh is not regular): forward vs the full-mesh oracle and the
    reverse pass vs autograd, for the general kernel arithmetic and for every fan instantiation (8 slots x valence <= 6 / 7,
    12 slots x valence <= 11; ``force_maxd`` runs the regular mesh through the valence-7 code)."""
    from oracle import smplh_lbs
    npz = synthetic.write_synthetic_smplh(asset_dir, seed=0, irregular=kind)
    sub_i, _ = submodel.submodel_from_npz(npz)
    assert sub_i['dims']['fan_ok']
    if kind:
        assert sorted(set(sub_i['sub.sensor_degree'].tolist())) == sorted(set(synthetic.IRREGULAR_VALENCES[kind]) | {6})
    osm = smplh_lbs.SmplhModel(npz, num_betas=10, dtype=torch.float64)
    topo = sensors.sensor_topology(osm.faces.numpy())
    n = 6
    theta, beta, off_r, off_t = _case(n, seed=11)
    rng = np.random.RandomState(12)
    pose, shape, pos, ori, joints = _oracle(osm, topo, theta, beta, off_r, off_t)
    meas_pos = pos.detach().numpy() + 0.01 * rng.standard_normal((n, 12, 3))
    meas_ori = ori.detach().numpy() + 0.05 * rng.standard_normal((n, 12, 3, 3))
    coef = rng.uniform(0.5, 2.0, size=n)
    dp = pos - torch.from_numpy(meas_pos)
    dr = (ori - torch.from_numpy(meas_ori)).reshape(n, 12, 9)
    energy = (torch.sqrt((dp * dp).sum(-1)).sum(-1) + torch.sqrt((dr * dr).sum(-1)).sum(-1)) * torch.from_numpy(coef)
    g_pose, g_shape = torch.autograd.grad(energy.sum(), [pose, shape])
    kw = dict(force_maxd=force_maxd) if form == 'fan' else {}
    out = EVAL[form](sub_i, theta, beta, off_r, off_t, meas_pos, meas_ori.reshape(n, 12, 9), np.ones(12), coef, use_double=True, **kw)
    np.testing.assert_allclose(out['sensor_pos'], pos.detach().numpy(), atol=2e-6, rtol=0)
    np.testing.assert_allclose(out['sensor_ori'], ori.detach().numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(out['joints'], joints.detach().numpy(), atol=2e-6, rtol=0)
    tol = 2e-5 * max(float(g_pose.abs().max()), 1.0)
    np.testing.assert_allclose(out['g_theta'], g_pose.numpy(), atol=tol, rtol=0)
    np.testing.assert_allclose(out['g_beta'], g_shape.numpy(), atol=tol, rtol=0)
