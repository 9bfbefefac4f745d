"""Constants and the configuration bag of the LGD path.

Mirrors the parts of ``empose/helpers/configuration.py`` the hot path reads: the ``CONSTANTS``
values at ``:32-34, 86-89, 104-107, 118`` and the model flags of ``Configuration.parse_cmd``
(``:150-209``).  When the reference package is installed its own ``Configuration`` objects work
unchanged with ``empose_b200.nn.models.create_model`` (only attributes are read).
"""
import json
import os

import torch


class _Constants(object):
    VERTEX_IDS = [3027, 3748, 5430, 5178, 5006, 4447, 4559, 1961, 1391, 1535, 959, 1072]
    S_CONFIG_6 = [0, 1, 2, 6, 7, 11]
    N_TRACKERS_WO_ROOT = 12
    N_JOINTS = 21
    MAX_INDEX_ROOT_AND_BODY = 66
    N_JOINTS_HAND = 15
    N_SHAPE_PARAMS = 10
    SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]
    DTYPE = torch.float32
    FPS = 60.0

    @property
    def DEVICE(self):
        return torch.device('cuda:0' if torch.cuda.is_available() else 'cpu')

    # The reference reads these four at import time and raises KeyError when unset
    # (configuration.py:25-28); here the KeyError is deferred to first use.
    @property
    def SMPL_MODELS_DIR(self):
        return os.environ['SMPL_MODELS']

    @property
    def DATA_DIR(self):
        return os.environ['EM_DATA_SYNTH']

    @property
    def EXPERIMENT_DIR(self):
        return os.environ['EM_EXPERIMENTS']

    @property
    def DATA_DIR_TEST(self):
        return os.environ['EM_DATA_REAL']


CONSTANTS = _Constants()

#: defaults of the flags the LGD path reads (reference configuration.py:150-209)
_DEFAULTS = dict(
    m_type='lgd', m_estimate_shape=False, m_shape_hidden_size=256, m_fk_loss=0.0, m_dropout=0.0, m_hidden_size=1024,
    m_num_layers=2, m_learn_init_state=False, m_bidirectional=False, m_num_iterations=4, m_dropout_hidden=0.0,
    m_step_size=0.1, m_reprojection_loss_weight=0.01, m_shape_loss_weight=1.0, m_pose_loss_weight=1.0,
    m_average_shape=False, m_use_gradient=False, m_skip_connections=False, m_no_batch_norm=False, m_rnn_init=False,
    m_rnn_denoiser=False, m_rnn_bidirectional=False, m_rnn_hidden_size=512, m_rnn_num_layers=2, use_marker_pos=False,
    use_marker_ori=False, use_marker_nor=False, n_markers=12, lr=0.001, window_size=120)


class Configuration(object):
    """Attribute bag with the reference's JSON round trip (configuration.py:137-144, 214-225)."""

    def __init__(self, adict=None, **overrides):
        self.__dict__.update(_DEFAULTS)
        self.__dict__.update(adict or {})
        self.__dict__.update(overrides)

    @staticmethod
    def from_json(json_path):
        with open(json_path, 'r') as f:
            return Configuration(json.load(f))

    def to_json(self, json_path):
        with open(json_path, 'w') as f:
            f.write(json.dumps(vars(self), indent=2, sort_keys=True))

    def __str__(self):
        return json.dumps(vars(self), indent=2, sort_keys=True, default=str)


def lgd_config(n_markers=12, num_iterations=4, rnn_init=True, hidden_size=512, window_size=32, **overrides):
    """The flag set of the released LGD models (reference README.md:51, 221)."""
    base = dict(m_type='lgd', m_num_iterations=num_iterations, m_hidden_size=hidden_size, m_rnn_init=rnn_init,
                m_average_shape=True, m_use_gradient=True, use_marker_pos=True, use_marker_ori=True,
                n_markers=n_markers, window_size=window_size)
    base.update(overrides)
    return Configuration(base)
