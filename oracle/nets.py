"""The learned layers of the LGD model, restated functionally from a state dict (test infrastructure).

Follows ``empose/nn/layers.py`` of the reference (MLP ``:46-77``, LinearLayers ``:13-43``,
RNNLayer ``:80-157``); pinned by ``tests/golden``.
"""
import torch

BN_EPS = 1e-5   # torch.nn.BatchNorm1d default, used as-is by layers.py:26,57


def _linear(x, sd, prefix):
    return x @ sd[prefix + '.weight'].T + sd[prefix + '.bias']


def _bn_eval(x, sd, prefix):
    if prefix + '.running_mean' not in sd:      # m_no_batch_norm -> nn.Identity (layers.py:59-60)
        return x
    inv = torch.rsqrt(sd[prefix + '.running_var'] + BN_EPS)
    return (x - sd[prefix + '.running_mean']) * inv * sd[prefix + '.weight'] + sd[prefix + '.bias']


def _prelu(x, sd, prefix):
    alpha = sd[prefix + '.weight']
    return torch.clamp(x, min=0) + alpha * torch.clamp(x, max=0)


def mlp_eval(x, sd, prefix, num_blocks=2, skip=False):
    """
    ``MLP.forward`` in eval mode (layers.py:70-77): in->hidden, BN, PReLU, then ``num_blocks``
    ``LinearLayers`` groups of two Linear+BN+PReLU each (optional skip around a group, :35-43),
    then hidden->out.  Dropout is the identity in eval mode.
    """
    has_bn = (prefix + '.batch_norm.running_mean') in sd
    y = _linear(x, sd, prefix + '.input_to_hidden')
    y = _bn_eval(y, sd, prefix + '.batch_norm')
    y = _prelu(y, sd, prefix + '.activation_fn')
    for b in range(num_blocks):
        base = '%s.hidden_layers.%d.layers' % (prefix, b)
        stride = 4 if has_bn else 3              # [Linear, (BN), PReLU, Dropout] per layer
        z = y
        for l in range(2):
            z = _linear(z, sd, '%s.%d' % (base, l * stride))
            if has_bn:
                z = _bn_eval(z, sd, '%s.%d' % (base, l * stride + 1))
            z = _prelu(z, sd, '%s.%d' % (base, l * stride + (2 if has_bn else 1)))
        y = y + z if skip else z
    return _linear(y, sd, prefix + '.hidden_to_output')


def lstm_packed(x, seq_lengths, sd, prefix, num_layers, init_state=None, bidirectional=False):
    """
    Multi-layer LSTM with packed-sequence semantics (layers.py:133-157): for ``t >= seq_lengths[b]`` the output row is
    zero and (h, c) stop updating.  Gate order i, f, g, o.  With ``bidirectional`` the reverse direction of every layer
    walks each sequence from its last valid frame down to frame 0 (``pack_padded_sequence`` semantics -- equivalently:
    time runs F-1 .. 0 and the state of a sequence only starts moving at ``t = len - 1``) and a layer's output is
    ``[h_forward | h_reverse]``.
    :param x: (B, F, in).  :param init_state: (h0, c0) each (layers * directions, B, H) or None.
    :return: outputs (B, F, H * directions), (h_n, c_n) each (layers * directions, B, H).
    """
    bsz, n_frames, _ = x.shape
    hid = sd['%s.weight_hh_l0' % prefix].shape[1]
    lengths = torch.as_tensor(seq_lengths).to(torch.long).reshape(-1)
    dirs = 2 if bidirectional else 1
    layer_in = x
    h_n, c_n = [], []
    for layer in range(num_layers):
        outs_dir = []
        for d in range(dirs):
            sfx = '_reverse' if d == 1 else ''
            w_ih = sd['%s.weight_ih_l%d%s' % (prefix, layer, sfx)]
            w_hh = sd['%s.weight_hh_l%d%s' % (prefix, layer, sfx)]
            bias = sd['%s.bias_ih_l%d%s' % (prefix, layer, sfx)] + sd['%s.bias_hh_l%d%s' % (prefix, layer, sfx)]
            if init_state is None:
                h = torch.zeros(bsz, hid, dtype=x.dtype)
                c = torch.zeros(bsz, hid, dtype=x.dtype)
            else:
                h, c = init_state[0][layer * dirs + d].to(x.dtype), init_state[1][layer * dirs + d].to(x.dtype)
            outs = [None] * n_frames
            for t in (range(n_frames) if d == 0 else range(n_frames - 1, -1, -1)):
                gates = layer_in[:, t] @ w_ih.T + h @ w_hh.T + bias
                gi, gf, gg, go = gates.split(hid, dim=1)
                c_new = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
                h_new = torch.sigmoid(go) * torch.tanh(c_new)
                live = (t < lengths).to(x.dtype).unsqueeze(1)
                c = live * c_new + (1 - live) * c
                h = live * h_new + (1 - live) * h
                outs[t] = live * h_new
            outs_dir.append(torch.stack(outs, dim=1))
            h_n.append(h)
            c_n.append(c)
        layer_in = torch.cat(outs_dir, dim=-1)
    return layer_in, (torch.stack(h_n), torch.stack(c_n))
