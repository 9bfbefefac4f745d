"""The LGD / IEF reconstruction loop, restated (test infrastructure).

Follows ``IterativeErrorFeedback.forward`` (``empose/nn/models.py:485-632``) with
``prepare_inputs`` (``:106-125``), ``get_estimated_real_markers`` (``:471-483``) and
``reconstruction_loss`` (``empose/nn/loss.py:23-41``).  Pinned by ``tests/golden`` (outputs of the
unmodified reference modules).  Inference semantics only (BatchNorm uses running statistics).
"""
import torch

from oracle import nets
from oracle import sensors
from oracle import smplh_lbs


class IefConfig(object):
    """The handful of reference ``Configuration`` flags (configuration.py:150-209) the loop reads."""

    def __init__(self, n_markers=12, num_iterations=4, step_size=0.1, rnn_init=True, average_shape=True,
                 use_gradient=True, use_marker_pos=True, use_marker_ori=True, hidden_size=512, num_layers=2,
                 rnn_hidden_size=512, rnn_num_layers=2, skip_connections=False, no_batch_norm=False):
        self.n_markers = n_markers
        self.num_iterations = num_iterations
        self.step_size = step_size
        self.rnn_init = rnn_init
        self.average_shape = average_shape
        self.use_gradient = use_gradient
        self.use_marker_pos = use_marker_pos
        self.use_marker_ori = use_marker_ori
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.rnn_hidden_size = rnn_hidden_size
        self.rnn_num_layers = rnn_num_layers
        self.skip_connections = skip_connections
        self.no_batch_norm = no_batch_norm

    @property
    def marker_idxs(self):
        return list(range(12)) if self.n_markers == 12 else list(sensors.S_CONFIG_6)


def frame_mask_from_lengths(seq_lengths, n_frames):
    """``utils.py:105-123`` specialised to a padded window of ``n_frames``."""
    t = torch.arange(n_frames).unsqueeze(0)
    return t < torch.as_tensor(seq_lengths).reshape(-1, 1)


def reconstruction_energy(measured, predicted, seq_lengths, marker_masks):
    """``loss.py:23-41``: sum over sensors of the L2 norm, masked mean over frames, mean over the batch."""
    diff = predicted - measured
    per_frame = torch.sqrt((diff * diff).sum(dim=-1)).sum(dim=-1)               # (B,F)
    if marker_masks is not None:
        all_present = marker_masks.logical_not().any(dim=-1).logical_not()
        per_frame = per_frame * all_present
    mask = frame_mask_from_lengths(seq_lengths, per_frame.shape[1]).to(per_frame.dtype)
    per_window = (per_frame * mask).sum(-1) / torch.as_tensor(seq_lengths).to(per_frame.dtype)
    return per_window.mean()


def prepare_inputs(cfg, marker_pos, marker_oris):
    """``models.py:106-125``: (B,F,36)/(B,F,108) -> (B,F,12M) = [pos | ori] with the optional 6-sensor gather."""
    b, f = marker_pos.shape[0], marker_pos.shape[1]
    pos = marker_pos.reshape(b, f, -1, 3)
    ori = marker_oris.reshape(b, f, -1, 3, 3)
    if cfg.n_markers == 6:
        pos, ori = pos[:, :, sensors.S_CONFIG_6], ori[:, :, sensors.S_CONFIG_6]
    parts = []
    if cfg.use_marker_pos:
        parts.append(pos.reshape(b, f, -1))
    if cfg.use_marker_ori:
        parts.append(ori.reshape(b, f, -1))
    return torch.cat(parts, dim=-1)


def project_sensors(smpl, topology, pose, shape, offset_r, offset_t):
    """``models.py:471-483``: SMPL -> sensor frames -> offsets; always all 12 sensors, joints cut to 22."""
    verts, joints = smplh_lbs.smpl_layer_forward(smpl, pose[:, 3:], shape, poses_root=pose[:, :3])
    pos, ori, _ = sensors.sensor_frames(verts, topology)
    pos_c, ori_c = sensors.apply_offsets(pos, ori, offset_r, offset_t)
    return pos_c, ori_c, joints[:, :smplh_lbs.N_BODY_JOINTS]


def ief_forward(cfg, sd, smpl, topology, marker_pos, marker_oris, offset_r, offset_t, seq_lengths,
                marker_masks=None, init_state=None):
    """
    One pass of the hot path over a batch of windows.
    :param sd: state dict with the reference's keys (``rnn.lstm.*``, ``pose_net_init.*`` ...), torch tensors.
    :param marker_pos: (B,F,36), marker_oris: (B,F,108), offset_r: (B,12,3,3), offset_t: (B,12,3).
    :param seq_lengths: (B,) ints.  marker_masks: (B,F,12) or None.  init_state: LSTM (h,c) or None.
    :return: dict with pose_hat (B,F,63), root_ori_hat (B,F,3), shape_hat (B,F,10), joints_hat (B,F,66),
             ``history`` (lists of N+1 tensors: pose, shape, joints, markers, markers_ori, each (B,F,dof)),
             ``grad_history`` (N tensors pairs) and ``final_state`` of the LSTM.
    """
    dt = marker_pos.dtype
    sd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}
    smpl = smpl.to(dt)
    inputs = prepare_inputs(cfg, marker_pos, marker_oris)
    bsz, n_frames, dof = inputs.shape
    rows = bsz * n_frames
    off_r = offset_r.unsqueeze(1).repeat(1, n_frames, 1, 1, 1).reshape(rows, -1, 3, 3)
    off_t = offset_t.unsqueeze(1).repeat(1, n_frames, 1, 1).reshape(rows, -1, 3)
    flat_in = inputs.reshape(rows, dof)

    final_state = None
    if cfg.rnn_init:                                                             # models.py:511-518
        out, final_state = nets.lstm_packed(inputs, seq_lengths, sd, 'rnn.lstm', cfg.rnn_num_layers, init_state)
        pose = (out @ sd['pose_net_init.weight'].T + sd['pose_net_init.bias']).reshape(rows, -1)
        shape = (out @ sd['shape_net_init.weight'].T + sd['shape_net_init.bias']).reshape(rows, -1)
    else:                                                                        # models.py:520-526
        pose = nets.mlp_eval(flat_in, sd, 'pose_net_init', cfg.num_layers, cfg.skip_connections)
        shape = nets.mlp_eval(flat_in, sd, 'shape_net_init', cfg.num_layers, cfg.skip_connections)

    def window_mean(s):                                                          # models.py:529-532
        return s.reshape(bsz, n_frames, -1).mean(dim=1, keepdim=True).repeat(1, n_frames, 1).reshape(rows, -1)

    if cfg.average_shape:
        shape = window_mean(shape)

    n_pos = cfg.n_markers * 3 if cfg.use_marker_pos else 0
    hist = {'pose': [], 'shape': [], 'joints': [], 'markers': [], 'markers_ori': []}
    grads = []
    idx = cfg.marker_idxs
    for it in range(cfg.num_iterations + 1):
        pose = pose.detach().requires_grad_(True)
        shape = shape.detach().requires_grad_(True)
        m_pos, m_ori, joints = project_sensors(smpl, topology, pose, shape, off_r, off_t)
        hist['pose'].append(pose.detach().reshape(bsz, n_frames, -1))
        hist['shape'].append(shape.detach().reshape(bsz, n_frames, -1))
        hist['joints'].append(joints.detach().reshape(bsz, n_frames, -1))
        hist['markers'].append(m_pos.detach().reshape(bsz, n_frames, -1))
        hist['markers_ori'].append(m_ori.detach().reshape(bsz, n_frames, -1))
        if it == cfg.num_iterations:
            break
        feats = [flat_in, pose.detach(), shape.detach()]
        if cfg.use_gradient:                                                     # models.py:553-582
            energy = torch.zeros((), dtype=dt)
            if cfg.use_marker_pos:
                energy = energy + reconstruction_energy(
                    flat_in[:, :n_pos].reshape(bsz, n_frames, -1, 3),
                    m_pos.reshape(bsz, n_frames, -1, 3)[:, :, idx], seq_lengths, marker_masks)
            if cfg.use_marker_ori:
                energy = energy + reconstruction_energy(
                    flat_in[:, n_pos:].reshape(bsz, n_frames, -1, 9),
                    m_ori.reshape(bsz, n_frames, -1, 9)[:, :, idx], seq_lengths, marker_masks)
            g_pose, g_shape = torch.autograd.grad(energy, [pose, shape])
            g_pose, g_shape = g_pose * rows, g_shape * rows                      # models.py:578-579
            grads.append((g_pose.reshape(bsz, n_frames, -1), g_shape.reshape(bsz, n_frames, -1)))
            feats += [g_pose, g_shape]
        x = torch.cat(feats, dim=-1)
        d_pose = nets.mlp_eval(x, sd, 'pose_net_iter', cfg.num_layers, cfg.skip_connections)
        d_shape = nets.mlp_eval(x, sd, 'shape_net_iter', cfg.num_layers, cfg.skip_connections)
        if cfg.average_shape:
            d_shape = window_mean(d_shape)
        pose = pose.detach() + cfg.step_size * d_pose                            # models.py:591-592
        shape = shape.detach() + cfg.step_size * d_shape

    last_pose = hist['pose'][-1]
    return {'pose_hat': last_pose[:, :, 3:], 'root_ori_hat': last_pose[:, :, :3],
            'shape_hat': hist['shape'][-1], 'joints_hat': hist['joints'][-1],
            'history': hist, 'grad_history': grads, 'final_state': final_state}
