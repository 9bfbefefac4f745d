"""The per-frame metrics restatement (oracle/metrics.py) vs the unmodified reference MetricsEngine (tests/golden/metrics.npz)."""
import numpy as np
import torch

from oracle import metrics as oracle_metrics

import util


def golden_frames(gold):
    """The frames the two golden ``compute`` calls keep, in order (masked call, then the unmasked one)."""
    b, f = gold['poses'].shape[:2]
    live = util.valid_frame_mask(gold['seq_lengths'], f) & (gold['marker_masks'] != 0).all(-1)
    rep = lambda a: np.repeat(a[:, None], f, axis=1)
    pose = np.concatenate([gold['poses'][live], gold['poses'].reshape(b * f, 66)])
    pose_hat = np.concatenate([gold['pose_hat'][live], gold['pose_hat'].reshape(b * f, 66)])
    shape = np.concatenate([rep(gold['shapes'])[live], rep(gold['shapes']).reshape(b * f, 10)])
    shape_hat = np.concatenate([gold['shape_hat'][live], rep(gold['shapes']).reshape(b * f, 10)])
    n_first = int(live.sum())
    # the second call passes no root poses: zero roots for both (metrics.py:211-213)
    pose[n_first:, :3] = 0.0
    pose_hat[n_first:, :3] = 0.0
    return pose, shape, pose_hat, shape_hat


def test_frame_metrics_match_reference(oracle_smpl):
    gold = util.load_golden('metrics')
    pose, shape, pose_hat, shape_hat = golden_frames(gold)
    t = lambda a: torch.from_numpy(a).double()
    eucl, eucl_pa, angle = oracle_metrics.frame_metrics(oracle_smpl, t(pose), t(shape), t(pose_hat), t(shape_hat))
    assert eucl.shape == gold['eucl'].shape and angle.shape == gold['angle'].shape
    np.testing.assert_allclose(eucl, gold['eucl'], atol=2e-6, rtol=0)
    np.testing.assert_allclose(eucl_pa, gold['eucl_pa'], atol=2e-6, rtol=0)
    np.testing.assert_allclose(angle, gold['angle'], atol=2e-3, rtol=0)        # degrees; the reference went through float32 axis-angles


def test_kernel_math_on_host_matches_oracle(smpl_npz, oracle_smpl):
    """csrc/metrics_math.h (what metrics_kernel runs per frame) compiled for the host: FK joints, Euclidean and
    Procrustes-aligned distances vs the restatement -- including near-degenerate point sets for the 3x3 SVD."""
    import host_math
    from empose_b200 import submodel
    sub, _ = submodel.submodel_from_npz(smpl_npz)
    g = torch.Generator().manual_seed(5)
    n = 64
    pose, pose_hat = 0.3 * torch.randn(n, 66, generator=g), 0.3 * torch.randn(n, 66, generator=g)
    shape, shape_hat = torch.randn(n, 10, generator=g), torch.randn(n, 10, generator=g)
    pose_hat[:4] = pose[:4]                                   # identical skeletons up to shape
    shape_hat[:2] = shape[:2]                                 # ... and fully identical ones (zero residual)
    want = oracle_metrics.frame_metrics(oracle_smpl, pose.double(), shape.double(), pose_hat.double(), shape_hat.double())
    eucl, pa, joints = host_math.metrics_eval(sub, pose.numpy(), shape.numpy(), pose_hat.numpy(), shape_hat.numpy())
    np.testing.assert_allclose(eucl, want[0], atol=3e-6, rtol=0)
    np.testing.assert_allclose(pa, want[1], atol=1e-5, rtol=0)
