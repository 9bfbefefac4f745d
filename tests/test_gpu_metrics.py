"""The metrics engine on the B200 vs the unmodified reference MetricsEngine (tests/golden/metrics.npz) and the
oracle.  Run on the B200 box: -m gpu."""
import numpy as np
import pytest
import torch

from oracle import metrics as oracle_metrics

import util

pytestmark = pytest.mark.gpu


def test_metrics_engine_matches_reference(smpl_npz, oracle_smpl):
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.eval.metrics import MetricsEngine
    assert torch.cuda.is_available()
    dev = torch.device('cuda:0')
    gold = util.load_golden('metrics')
    me = MetricsEngine(SMPLLayer(smpl_npz).to(device=dev, dtype=torch.float32), keep_tables=True)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
    me.compute(t(gold['poses'][:, :, 3:]), t(gold['shapes']), t(gold['pose_hat'][:, :, 3:]), t(gold['shape_hat']), t(gold['seq_lengths']),
               pose_root=t(gold['poses'][:, :, :3]), pose_root_hat=t(gold['pose_hat'][:, :, :3]), frame_mask=t(gold['marker_masks']))
    me.compute(t(gold['poses'][:, :, 3:]), t(gold['shapes']), t(gold['pose_hat'][:, :, 3:]), None, None)
    eucl, eucl_pa, angle = (np.concatenate(x, axis=0) for x in (me.eucl_dists, me.eucl_dists_pa, me.angle_diffs))
    errs = dict(eucl=float(np.abs(eucl - gold['eucl']).max()), eucl_pa=float(np.abs(eucl_pa - gold['eucl_pa']).max()),
                angle=float(np.abs(angle - gold['angle']).max()))
    metrics = me.get_metrics()
    util.report('metrics_engine', **errs, **{k: float(v) for k, v in metrics.items()})
    assert errs['eucl'] <= 5e-6 and errs['eucl_pa'] <= 1e-5, errs          # metres
    assert errs['angle'] <= 5e-3, errs                                      # degrees
    for k, v in metrics.items():
        assert abs(float(v) - float(gold['m/' + k])) <= 2e-3 * max(1.0, abs(float(gold['m/' + k]))), k
    assert me.to_pretty_string(metrics, 'golden').splitlines()[0] == str(gold['pretty']).splitlines()[0]
    # joints given directly (compute_joint_dist) and a larger random batch against the oracle
    g = torch.Generator().manual_seed(5)
    n = 700
    pose, pose_hat = 0.3 * torch.randn(n, 66, generator=g), 0.3 * torch.randn(n, 66, generator=g)
    shape, shape_hat = torch.randn(n, 10, generator=g), torch.randn(n, 10, generator=g)
    want = oracle_metrics.frame_metrics(oracle_smpl, pose.double(), shape.double(), pose_hat.double(), shape_hat.double())
    got = me._context(dev).metrics(pose.to(dev), shape.to(dev), pose_hat.to(dev), shape_hat.to(dev))
    for w, gt_, tol in zip(want, got, (5e-6, 2e-5, 5e-3)):
        assert float(np.abs(gt_.cpu().numpy() - w).max()) <= tol
    me.reset()
    from oracle import smplh_lbs
    _, j = smplh_lbs.smpl_layer_forward(oracle_smpl, pose[:, 3:].double(), shape.double(), poses_root=pose[:, :3].double())
    _, jh = smplh_lbs.smpl_layer_forward(oracle_smpl, pose_hat[:, 3:].double(), shape_hat.double(), poses_root=pose_hat[:, :3].double())
    me.compute_joint_dist(j[:, :22].reshape(7, 100, 66).float().to(dev), jh[:, :22].reshape(7, 100, 66).float().to(dev))
    assert float(np.abs(me.eucl_dists[0] - want[0]).max()) <= 5e-6 and float(np.abs(me.eucl_dists_pa[0] - want[1]).max()) <= 2e-5


def test_metrics_aggregate_on_device_and_local_angles(smpl_npz, oracle_smpl):
    """get_metrics() from the device-side running moments equals the reference's statistics over the concatenated tables
    (mean over joints of per-joint means, population std over all selected entries), also across several compute() calls of
    different size and without keeping any table; angle_glob=False gives the angles between the LOCAL joint rotations."""
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.eval.metrics import MetricsEngine
    dev = torch.device('cuda:0')
    layer = SMPLLayer(smpl_npz).to(device=dev, dtype=torch.float32)
    me, keep = MetricsEngine(layer), MetricsEngine(layer, keep_tables=True)
    g = torch.Generator().manual_seed(17)
    for n, f in ((3, 40), (5, 7), (1, 1)):
        pose, pose_hat = 0.3 * torch.randn(n, f, 63, generator=g), 0.3 * torch.randn(n, f, 63, generator=g)
        shape = torch.randn(n, 10, generator=g)
        lens = torch.randint(1, f + 1, (n,), generator=g)
        for m in (me, keep):
            m.compute(pose.to(dev), shape.to(dev), pose_hat.to(dev), None, lens.to(dev))
    with pytest.raises(AttributeError):
        me.eucl_dists
    eucl, pa, ang = (np.concatenate(x, axis=0) for x in (keep.eucl_dists, keep.eucl_dists_pa, keep.angle_diffs))
    want = {'MPJPE [mm]': 1000 * np.mean(np.mean(eucl, axis=0)[keep.eucl_idxs]), 'MPJPE STD': 1000 * np.std(eucl[:, keep.eucl_idxs].astype(np.float64)),
            'PA-MPJPE [mm]': 1000 * np.mean(np.mean(pa, axis=0)[keep.eucl_idxs]), 'PA-MPJPE STD': 1000 * np.std(pa[:, keep.eucl_idxs].astype(np.float64)),
            'MPJAE [deg]': np.mean(np.mean(ang, axis=0)[keep.angle_idxs]), 'MPJAE STD': np.std(ang[:, keep.angle_idxs])}
    got = me.get_metrics()
    assert list(got) == list(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-5 * max(1.0, abs(want[k])), k
    # local angles: geodesic distance of exp(theta_j), exp(theta_hat_j)
    loc = MetricsEngine(layer, keep_tables=True)
    loc.angle_glob = False
    pose, pose_hat = 0.3 * torch.randn(2, 9, 63, generator=g), 0.3 * torch.randn(2, 9, 63, generator=g)
    loc.compute(pose.to(dev), torch.zeros(2, 10, device=dev), pose_hat.to(dev))
    from scipy.spatial.transform import Rotation
    a = Rotation.from_rotvec(pose.reshape(-1, 3).double().numpy())
    b = Rotation.from_rotvec(pose_hat.reshape(-1, 3).double().numpy())
    want_ang = np.rad2deg((a.inv() * b).magnitude()).reshape(18, 21)
    assert float(np.abs(loc.angle_diffs[0] - want_ang).max()) <= 5e-3
