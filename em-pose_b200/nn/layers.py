"""Parameter containers of the learned layers.

The classes below exist so that ``state_dict()`` / ``load_state_dict()`` use exactly the key layout
of the reference (``empose/nn/layers.py:13-77`` MLP / LinearLayers, ``:80-157`` RNNLayer) and released
``model.pth`` files load with ``strict=True``.  They hold ``torch.nn`` parameter modules but have no
``forward``: the arithmetic runs in the CUDA library (``csrc/``), never in PyTorch.
"""
from torch import nn as nn


def _no_forward(self, *args, **kwargs):
    raise RuntimeError('%s is a parameter container; the LGD path runs in libempose_b200 '
                       '(call the IterativeErrorFeedback module instead)' % type(self).__name__)


class LinearLayers(nn.Module):
    """Two [Linear, BatchNorm1d?, PReLU, Dropout] groups under ``.layers`` (indices 0..7 / 0..5)."""
    forward = _no_forward

    def __init__(self, hidden_size, num_layers=2, dropout_p=0.0, use_skip=False, use_batch_norm=True):
        super(LinearLayers, self).__init__()
        self.hidden_size = hidden_size
        self.use_skip = use_skip
        mods = []
        for _ in range(num_layers):
            mods.append(nn.Linear(hidden_size, hidden_size))
            if use_batch_norm:
                norm = nn.BatchNorm1d(hidden_size)
                nn.init.uniform_(norm.weight)          # layers.py:27
                mods.append(norm)
            mods += [nn.PReLU(), nn.Dropout(dropout_p)]
        self.layers = nn.Sequential(*mods)


class MLP(nn.Module):
    """input_to_hidden, batch_norm, activation_fn, hidden_layers.{b}, hidden_to_output (layers.py:52-68)."""
    forward = _no_forward

    def __init__(self, input_size, output_size, hidden_size, num_layers=2, dropout_p=0.0, skip_connection=False,
                 use_batch_norm=True):
        super(MLP, self).__init__()
        self.input_to_hidden = nn.Linear(input_size, hidden_size)
        if use_batch_norm:
            self.batch_norm = nn.BatchNorm1d(hidden_size)
            nn.init.uniform_(self.batch_norm.weight)   # layers.py:58
        else:
            self.batch_norm = nn.Identity()
        self.activation_fn = nn.PReLU()
        self.dropout = nn.Dropout(dropout_p)
        self.hidden_to_output = nn.Linear(hidden_size, output_size)
        self.hidden_layers = nn.Sequential(*[
            LinearLayers(hidden_size, dropout_p=dropout_p, use_batch_norm=use_batch_norm, use_skip=skip_connection)
            for _ in range(num_layers)])


class RNNLayer(nn.Module):
    """``.lstm`` (torch layout, gate order i,f,g,o) plus the ``init_state`` / ``final_state`` carry (layers.py:108-114)."""
    forward = _no_forward

    def __init__(self, input_size, hidden_size, num_layers, output_size=None, bidirectional=False, dropout=0.0,
                 learn_init_state=False):
        super(RNNLayer, self).__init__()
        if learn_init_state or output_size is not None:
            raise NotImplementedError('learned initial states / an output layer inside RNNLayer are not supported')
        if dropout > 0.0:
            raise ValueError('input dropout > 0 is not supported (the released models use 0)')
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.is_bidirectional, self.num_directions, self.learn_init_state = bool(bidirectional), 2 if bidirectional else 1, False
        self.init_state = None
        self.final_state = None
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers, bidirectional=bool(bidirectional))
