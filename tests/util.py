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


#: flags of the training golden cases, mirrored from tests/golden/make_golden_train.py
TRAIN_CASES = {
    'train_lgd_rnn12_n4': dict(n_markers=12, num_iterations=4, rnn_init=True, pose_weight=10.0, fk_weight=0.1),
    'train_lgd_mlp12_n2': dict(n_markers=12, num_iterations=2, rnn_init=False, pose_weight=1.0, fk_weight=0.0),
    'train_lgd_rnn6_n2_real': dict(n_markers=6, num_iterations=2, rnn_init=True, pose_weight=10.0, fk_weight=0.1),
}


#: (Bi)RNN golden cases, mirrored from tests/golden/make_golden_rnn.py
RNN_CASES = {
    'rnn_bi12_shape_fk': dict(cfg=dict(n_markers=12, hidden_size=1024, num_layers=2, bidirectional=True, estimate_shape=True,
                                       average_shape=True, do_fk=True),
                              weights=dict(n_markers=12, hidden_size=1024, num_layers=2, bidirectional=True, estimate_shape=True)),
    'rnn_bi6': dict(cfg=dict(n_markers=6, hidden_size=256, num_layers=2, bidirectional=True),
                    weights=dict(n_markers=6, hidden_size=256, num_layers=2, bidirectional=True, estimate_shape=False)),
    'rnn_uni12': dict(cfg=dict(n_markers=12, hidden_size=128, num_layers=3, bidirectional=False),
                      weights=dict(n_markers=12, hidden_size=128, num_layers=3, bidirectional=False, estimate_shape=False)),
}


def train_inputs(gold, dtype=torch.float32):
    get = lambda k: torch.from_numpy(gold[k]).to(dtype)
    masks = torch.from_numpy(gold['marker_masks']) if 'marker_masks' in gold else None
    return dict(marker_pos=get('marker_pos'), marker_oris=get('marker_oris'), offset_r=get('offset_r'),
                offset_t=get('offset_t'), seq_lengths=torch.from_numpy(gold['seq_lengths']), marker_masks=masks,
                poses_gt=get('poses'), shapes_gt=get('shapes'), joints_gt=get('joints_gt'))


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


# ----------------------------------------------------------------------------------------------------------------------
# product-side helpers
# ----------------------------------------------------------------------------------------------------------------------
class DuckBatch(object):
    """The slice of the reference's ``ABatch`` interface the model reads (data.py:304-309, 433-459)."""

    def __init__(self, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks=None):
        self.marker_pos, self.marker_oris = marker_pos, marker_oris
        self.offset_r, self.offset_t = offset_r, offset_t
        self.seq_lengths = seq_lengths
        self.marker_masks = marker_masks

    @property
    def batch_size(self):
        return self.marker_pos.shape[0]

    @property
    def seq_length(self):
        return self.marker_pos.shape[1]

    def get_inputs(self, sf=None, ef=None, **kwargs):
        return {'marker_pos': self.marker_pos[:, sf:ef], 'marker_oris': self.marker_oris[:, sf:ef],
                'offset_r': self.offset_r, 'offset_t': self.offset_t,
                'marker_masks': None if self.marker_masks is None else self.marker_masks[:, sf:ef]}

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return DuckBatch(mv(self.marker_pos), mv(self.marker_oris), mv(self.offset_r), mv(self.offset_t),
                         mv(self.seq_lengths), mv(self.marker_masks))


def build_module(smpl_npz, n_markers=12, num_iterations=4, rnn_init=True, precision=0, device='cpu', hidden_size=512,
                 **config_overrides):
    """An ``empose_b200`` IterativeErrorFeedback in eval mode carrying the deterministic synthetic weights."""
    from empose_b200 import synthetic
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.helpers.configuration import lgd_config
    from empose_b200.nn.models import IterativeErrorFeedback
    cfg = lgd_config(n_markers=n_markers, num_iterations=num_iterations, rnn_init=rnn_init, hidden_size=hidden_size,
                     **config_overrides)
    smpl = SMPLLayer(smpl_npz).to(dtype=torch.float32)
    net = IterativeErrorFeedback(cfg, smpl, precision=precision)
    sd = net.state_dict()
    synth = synthetic.synth_state_dict(seed=0, n_markers=n_markers, rnn_init=rnn_init, hidden_size=hidden_size,
                                       rnn_hidden_size=cfg.m_rnn_hidden_size, num_layers=cfg.m_num_layers,
                                       batch_norm=not cfg.m_no_batch_norm, use_gradient=cfg.m_use_gradient,
                                       use_marker_pos=cfg.use_marker_pos, use_marker_ori=cfg.use_marker_ori)
    for k, v in synth.items():
        sd[k] = torch.from_numpy(np.asarray(v))
    net.load_state_dict(sd, strict=True)
    return net.to(device).eval()


def oracle_inputs_from_params(oracle_smpl, topology, params, seed):
    """Synthetic measured sensors for ``synth_window_params`` output via the oracle projection (CPU, float64)."""
    from empose_b200 import synthetic
    from oracle import ief as oracle_ief
    b, f = params['poses'].shape[:2]
    t = lambda a: torch.from_numpy(np.asarray(a)).double()
    poses = t(params['poses']).reshape(b * f, 66)
    shapes = t(params['shapes']).unsqueeze(1).repeat(1, f, 1).reshape(b * f, 10)
    off_r = t(params['offset_r']).unsqueeze(1).repeat(1, f, 1, 1, 1).reshape(b * f, 12, 3, 3)
    off_t = t(params['offset_t']).unsqueeze(1).repeat(1, f, 1, 1).reshape(b * f, 12, 3)
    with torch.no_grad():
        pos, ori, _ = oracle_ief.project_sensors(oracle_smpl, topology, poses, shapes, off_r, off_t)
    mpos, mori = synthetic.synth_measurements(pos.reshape(b, f, 12, 3).numpy(), ori.reshape(b, f, 12, 3, 3).numpy(), seed=seed)
    masks = None if params['marker_masks'] is None else torch.from_numpy(params['marker_masks'])
    return dict(marker_pos=torch.from_numpy(mpos), marker_oris=torch.from_numpy(mori),
                offset_r=torch.from_numpy(params['offset_r']), offset_t=torch.from_numpy(params['offset_t']),
                seq_lengths=torch.from_numpy(params['seq_lengths']), marker_masks=masks)


def report(kind, **vals):
    """Append one JSON line with achieved errors to gpurun_out/parity_report.jsonl (kept under profiles/)."""
    import json
    out_dir = os.path.join(os.path.dirname(GOLDEN_DIR.rstrip('/')).rsplit('/tests', 1)[0], 'gpurun_out')
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, 'parity_report.jsonl'), 'a') as f:
            f.write(json.dumps(dict(kind=kind, **{k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in vals.items()})) + '\n')
    except OSError:
        pass
