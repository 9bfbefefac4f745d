"""Shared helpers for the tests (golden loading, oracle driving, error metrics)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

#: flags of the golden cases, mirrored from tests/golden/make_golden.py
GOLDEN_CASES = {
    'lgd_rnn12_n4': dict(n_markers=12, num_iterations=4, rnn_init=True),
    'lgd_mlp12_n4': dict(n_markers=12, num_iterations=4, rnn_init=False),
    'lgd_rnn6_n2_real': dict(n_markers=6, num_iterations=2, rnn_init=True),
}


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + '.npz')) as z:
        return {k: z[k] for k in z.files}


def torch_state_dict(np_sd, dtype=torch.float32):
    out = {}
    for k, v in np_sd.items():
        t = torch.from_numpy(np.asarray(v))
        out[k] = t.to(dtype) if t.is_floating_point() else t
    return out


def chunk_inputs(gold, c, dtype=torch.float32):
    tag = 'c%d_' % c
    get = lambda k: torch.from_numpy(gold[tag + k]).to(dtype)
    masks = torch.from_numpy(gold[tag + 'marker_masks']) if (tag + 'marker_masks') in gold else None
    return dict(marker_pos=get('marker_pos'), marker_oris=get('marker_oris'), offset_r=get('offset_r'),
                offset_t=get('offset_t'), seq_lengths=torch.from_numpy(gold[tag + 'seq_lengths']),
                marker_masks=masks)


def valid_frame_mask(seq_lengths, n_frames):
    t = np.arange(n_frames)[None, :]
    return t < np.asarray(seq_lengths).reshape(-1, 1)


def max_joint_angle_err(pose_a, pose_b):
    """Largest per-joint axis-angle difference (rad) between two (..., 3k) pose arrays."""
    d = (np.asarray(pose_a, dtype=np.float64) - np.asarray(pose_b, dtype=np.float64))
    d = d.reshape(d.shape[:-1] + (-1, 3))
    return float(np.sqrt((d * d).sum(-1)).max())


def max_joint_pos_err_mm(j_a, j_b):
    d = (np.asarray(j_a, dtype=np.float64) - np.asarray(j_b, dtype=np.float64))
    d = d.reshape(d.shape[:-1] + (-1, 3))
    return float(np.sqrt((d * d).sum(-1)).max() * 1000.0)
