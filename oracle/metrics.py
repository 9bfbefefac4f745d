"""The metrics of the reference's ``MetricsEngine.compute`` restated per frame (test infrastructure).

Follows ``empose/eval/metrics.py``: joint positions through SMPL FK for ground truth and prediction (``:219-225``),
per-joint Euclidean distance (``:130-132``), the same after a per-frame Procrustes alignment with optimal scale
(``_procrustes``, ``:19-66``; ``:115-125``) and the angular distance of the GLOBAL joint orientations obtained with a zero
root (``:229-238``, ``helpers/utils.py:165-199``, ``quaternion.rotation_intrinsic_distance`` = geodesic angle).
Pinned by ``tests/golden/metrics.npz`` (the unmodified reference with a scipy stand-in for the quaternion package).
"""
import numpy as np
import torch

from oracle import smplh_lbs

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]      # configuration.py:118


def procrustes(x, y):
    """``metrics.py:19-66`` with ``compute_optimal_scale=True``: returns the aligned copy of ``y``."""
    mu_x, mu_y = x.mean(0), y.mean(0)
    x0, y0 = x - mu_x, y - mu_y
    norm_x, norm_y = np.sqrt((x0 ** 2).sum()), np.sqrt((y0 ** 2).sum())
    x0, y0 = x0 / norm_x, y0 / norm_y
    u, s, vt = np.linalg.svd(x0.T @ y0, full_matrices=False)
    v = vt.T
    t = v @ u.T
    det = np.linalg.det(t)
    v[:, -1] *= np.sign(det)
    s[-1] *= np.sign(det)
    t = v @ u.T
    return norm_x * s.sum() * (y0 @ t) + mu_x


def global_orientations(pose_body):
    """(n, 63) body pose -> (n, 22, 3, 3) global orientations with a zero root (metrics.py:229-236)."""
    n = pose_body.shape[0]
    pose = torch.cat([torch.zeros(n, 3, dtype=pose_body.dtype), pose_body], dim=-1).reshape(n * 22, 3)
    local = smplh_lbs.rodrigues(pose).reshape(n, 22, 3, 3)     # so3_exponential_map of the reference (helpers/so3.py), same map
    out = [None] * 22
    for j in range(22):
        out[j] = local[:, j] if SMPL_PARENTS[j] < 0 else out[SMPL_PARENTS[j]] @ local[:, j]
    return torch.stack(out, dim=1)


def frame_metrics(smpl, pose, shape, pose_hat, shape_hat):
    """
    :param pose, pose_hat: (n, 66) [root | body]; shape, shape_hat: (n, 10).
    :return: eucl (n, 22) metres, eucl_pa (n, 22), angle (n, 21) degrees -- one row per frame, all joints (the joint
             selections of metrics.py:82-95 are applied when aggregating).
    """
    with torch.no_grad():
        _, j = smplh_lbs.smpl_layer_forward(smpl, pose[:, 3:], shape, poses_root=pose[:, :3])
        _, j_hat = smplh_lbs.smpl_layer_forward(smpl, pose_hat[:, 3:], shape_hat, poses_root=pose_hat[:, :3])
        j, j_hat = j[:, :22].double().numpy(), j_hat[:, :22].double().numpy()
        eucl = np.sqrt(((j - j_hat) ** 2).sum(-1))
        aligned = np.stack([procrustes(j[i], j_hat[i]) for i in range(j.shape[0])])
        eucl_pa = np.sqrt(((j - aligned) ** 2).sum(-1))
        g, g_hat = global_orientations(pose[:, 3:].double()), global_orientations(pose_hat[:, 3:].double())
        rel = g.transpose(-1, -2) @ g_hat
        cos = ((rel.diagonal(dim1=-2, dim2=-1).sum(-1) - 1.0) / 2.0)
        skew = torch.stack([rel[..., 2, 1] - rel[..., 1, 2], rel[..., 0, 2] - rel[..., 2, 0], rel[..., 1, 0] - rel[..., 0, 1]], dim=-1) / 2.0
        angle = torch.atan2(skew.norm(dim=-1), cos)[:, 1:]
    return eucl, eucl_pa, np.rad2deg(angle.numpy())
