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
