"""The (Bi)RNN baseline, restated (test infrastructure).

Follows ``SimpleRNN.forward`` (``empose/nn/models.py:291-317``) with ``prepare_inputs`` (``:106-125``), ``RNNLayer``
(``empose/nn/layers.py:80-157``), the BatchNorm-free ``to_shape`` MLP (``models.py:276-280``) and ``maybe_do_fk``
(``models.py:134-144``).  Pinned by ``tests/golden/rnn_*.npz`` (outputs of the unmodified reference).  Inference
semantics; ``m_learn_init_state`` is not restated (the reference's own implementation of it hands ``(c0, h0)`` to an
LSTM that expects ``(h0, c0)`` and cannot be used with a bidirectional LSTM, ``layers.py:121-131``).
"""
import torch

from oracle import ief
from oracle import nets
from oracle import smplh_lbs


class RnnConfig(object):
    def __init__(self, n_markers=12, hidden_size=1024, num_layers=2, bidirectional=True, estimate_shape=False,
                 average_shape=False, do_fk=False, use_marker_pos=True, use_marker_ori=True):
        self.n_markers = n_markers
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.bidirectional = bidirectional
        self.estimate_shape = estimate_shape
        self.average_shape = average_shape
        self.do_fk = do_fk
        self.use_marker_pos = use_marker_pos
        self.use_marker_ori = use_marker_ori


def rnn_forward(cfg, sd, smpl, marker_pos, marker_oris, seq_lengths, init_state=None):
    """
    :return: dict pose_hat (B,F,63), root_ori_hat (B,F,3), shape_hat (B,F,10) | None, joints_hat (B,F,66) | None and the
             LSTM ``final_state`` (h_n, c_n), each (layers * directions, B, H).
    """
    dt = marker_pos.dtype
    sd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}
    inputs = ief.prepare_inputs(cfg, marker_pos, marker_oris)
    out, final_state = nets.lstm_packed(inputs, seq_lengths, sd, 'rnn.lstm', cfg.num_layers, init_state, cfg.bidirectional)
    pose = out @ sd['to_pose.weight'].T + sd['to_pose.bias']                      # (B,F,66), models.py:298
    shape = None
    if cfg.estimate_shape:                                                        # models.py:302-306
        b, f, w = out.shape
        shape = nets.mlp_eval(out.reshape(b * f, w), sd, 'to_shape', 2, False).reshape(b, f, -1)
        if cfg.average_shape:
            shape = shape.mean(dim=1, keepdim=True).repeat(1, f, 1)
    joints = None
    if cfg.do_fk:                                                                 # models.py:134-144
        b, f = pose.shape[:2]
        _, j = smplh_lbs.smpl_layer_forward(smpl.to(dt), pose[:, :, 3:].reshape(b * f, -1), shape.reshape(b * f, -1),
                                            poses_root=pose[:, :, :3].reshape(b * f, -1))
        joints = j[:, :smplh_lbs.N_BODY_JOINTS].reshape(b, f, -1)
    return {'pose_hat': pose[:, :, 3:], 'root_ori_hat': pose[:, :, :3], 'shape_hat': shape, 'joints_hat': joints,
            'final_state': final_state}
