"""One training step of the LGD / IEF model, restated (test infrastructure).

Follows what ``scripts/train.py:136-149`` runs: ``IterativeErrorFeedback.forward`` in train mode
(``empose/nn/models.py:485-632``: BatchNorm uses batch statistics and updates its running statistics,
``layers.py:26,57``) followed by ``IterativeErrorFeedback.backward`` (``models.py:634-688``).  Pinned by
``tests/golden/train_*.npz`` (losses, running statistics and parameter gradients of the unmodified reference).

Two properties of the reference that this restatement keeps on purpose:

* the forward pass itself leaves gradients in ``.grad``: ``reconstruction_error.backward(retain_graph=True)``
  (``models.py:576``) runs once per iteration and reaches every parameter upstream of that iterate, and
  ``scripts/train.py:138`` zeroes the gradients BEFORE the forward pass -- so the optimiser sees
  ``d(total_loss) + sum_i d(E_i)``;
* the iter-MLP inputs are detached (``models.py:549-551, 578-579``), so the only path from an iterate back to the
  parameters is the additive chain ``theta_i = theta_0 + step * sum_{k<i} dtheta_k``.
"""
import torch
import torch.nn.functional as F

from oracle import ief
from oracle import nets

BN_MOMENTUM = 0.1   # torch.nn.BatchNorm1d default (layers.py:26,57 construct it with defaults)


def is_trainable(key):
    return not (key.startswith('smpl.') or 'running_' in key or 'num_batches_tracked' in key)


class _Bn(object):
    """Train-mode BatchNorm1d over the rows of x; running statistics are updated in ``buffers``."""

    def __init__(self, buffers):
        self.buffers = buffers

    def __call__(self, x, params, prefix):
        if prefix + '.running_mean' not in self.buffers:
            return x
        y = F.batch_norm(x, self.buffers[prefix + '.running_mean'], self.buffers[prefix + '.running_var'],
                         params[prefix + '.weight'], params[prefix + '.bias'], training=True, momentum=BN_MOMENTUM,
                         eps=nets.BN_EPS)
        self.buffers[prefix + '.num_batches_tracked'] += 1
        return y


def mlp_train(x, params, bn, prefix, num_blocks=2, skip=False):
    """``MLP.forward`` in train mode with dropout p = 0 (layers.py:70-77)."""
    has_bn = (prefix + '.batch_norm.running_mean') in bn.buffers
    y = nets._linear(x, params, prefix + '.input_to_hidden')
    y = bn(y, params, prefix + '.batch_norm')
    y = nets._prelu(y, params, prefix + '.activation_fn')
    for b in range(num_blocks):
        base = '%s.hidden_layers.%d.layers' % (prefix, b)
        stride = 4 if has_bn else 3
        z = y
        for l in range(2):
            z = nets._linear(z, params, '%s.%d' % (base, l * stride))
            if has_bn:
                z = bn(z, params, '%s.%d' % (base, l * stride + 1))
            z = nets._prelu(z, params, '%s.%d' % (base, l * stride + (2 if has_bn else 1)))
        y = y + z if skip else z
    return nets._linear(y, params, prefix + '.hidden_to_output')


def padded_l1(gt, hat, seq_lengths):
    """``loss.py:13-20`` with ``nn.L1Loss(reduction='none')`` (models.py:457)."""
    unreduced = (gt - hat).abs().mean(-1)
    mask = ief.frame_mask_from_lengths(seq_lengths, unreduced.shape[1]).to(unreduced.dtype)
    return ((unreduced * mask).sum(-1) / torch.as_tensor(seq_lengths).to(unreduced.dtype)).mean()


def ief_train_step(cfg, sd, smpl, topology, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks,
                   poses_gt, shapes_gt, joints_gt, pose_weight=1.0, shape_weight=1.0, r_weight=0.01, fk_weight=0.0):
    """
    :param sd: state dict (reference keys).  Not modified.
    :param poses_gt: (B,F,66) [root | body], shapes_gt: (B,10), joints_gt: (B,F,66) or None.
    :return: dict with ``loss_vals`` (the five floats of models.py:676-680), ``grads`` {key: tensor} for every
             trainable key, ``buffers`` {key: tensor} (running statistics after the step) and the model outputs.
    """
    dt = marker_pos.dtype
    params = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in sd.items() if is_trainable(k)}
    buffers = {k: (v.detach().clone().to(dt) if v.is_floating_point() else v.detach().clone())
               for k, v in sd.items() if not is_trainable(k) and not k.startswith('smpl.')}
    bn = _Bn(buffers)
    names = sorted(params)
    plist = [params[k] for k in names]
    grads = {k: torch.zeros_like(params[k]) for k in names}

    def accumulate(scalar, retain):
        got = torch.autograd.grad(scalar, plist, retain_graph=retain, allow_unused=True)
        for k, g in zip(names, got):
            if g is not None:
                grads[k] += g

    smpl = smpl.to(dt)
    inputs = ief.prepare_inputs(cfg, marker_pos, marker_oris)
    bsz, n_frames, dof = inputs.shape
    rows = bsz * n_frames
    off_r = offset_r.unsqueeze(1).repeat(1, n_frames, 1, 1, 1).reshape(rows, -1, 3, 3)
    off_t = offset_t.unsqueeze(1).repeat(1, n_frames, 1, 1).reshape(rows, -1, 3)
    flat_in = inputs.reshape(rows, dof)

    if cfg.rnn_init:
        out, _ = nets.lstm_packed(inputs, seq_lengths, params, 'rnn.lstm', cfg.rnn_num_layers, None)
        pose = (out @ params['pose_net_init.weight'].T + params['pose_net_init.bias']).reshape(rows, -1)
        shape = (out @ params['shape_net_init.weight'].T + params['shape_net_init.bias']).reshape(rows, -1)
    else:
        pose = mlp_train(flat_in, params, bn, 'pose_net_init', cfg.num_layers, cfg.skip_connections)
        shape = mlp_train(flat_in, params, bn, 'shape_net_init', cfg.num_layers, cfg.skip_connections)

    def window_mean(s):
        return s.reshape(bsz, n_frames, -1).mean(dim=1, keepdim=True).repeat(1, n_frames, 1).reshape(rows, -1)

    if cfg.average_shape:
        shape = window_mean(shape)

    n_pos = cfg.n_markers * 3 if cfg.use_marker_pos else 0
    idx = cfg.marker_idxs
    hist = {'pose': [], 'shape': [], 'joints': [], 'markers': [], 'markers_ori': []}

    def recon(m_pos, m_ori):
        e = torch.zeros((), dtype=dt)
        if cfg.use_marker_pos:
            e = e + ief.reconstruction_energy(flat_in[:, :n_pos].reshape(bsz, n_frames, -1, 3),
                                              m_pos.reshape(bsz, n_frames, -1, 3)[:, :, idx], seq_lengths, marker_masks)
        if cfg.use_marker_ori:
            e = e + ief.reconstruction_energy(flat_in[:, n_pos:].reshape(bsz, n_frames, -1, 9),
                                              m_ori.reshape(bsz, n_frames, -1, 9)[:, :, idx], seq_lengths, marker_masks)
        return e

    for it in range(cfg.num_iterations + 1):
        m_pos, m_ori, joints = ief.project_sensors(smpl, topology, pose, shape, off_r, off_t)
        for k, v in (('pose', pose), ('shape', shape), ('joints', joints), ('markers', m_pos), ('markers_ori', m_ori)):
            hist[k].append(v)
        if it == cfg.num_iterations:
            break
        feats = [flat_in, pose.detach(), shape.detach()]
        if cfg.use_gradient:
            energy = recon(m_pos, m_ori)
            g_pose, g_shape = torch.autograd.grad(energy, [pose, shape], retain_graph=True)
            accumulate(energy, True)                                             # the side effect of models.py:576
            feats += [g_pose.detach() * rows, g_shape.detach() * rows]
        x = torch.cat(feats, dim=-1)
        d_pose = mlp_train(x, params, bn, 'pose_net_iter', cfg.num_layers, cfg.skip_connections)
        d_shape = mlp_train(x, params, bn, 'shape_net_iter', cfg.num_layers, cfg.skip_connections)
        if cfg.average_shape:
            d_shape = window_mean(d_shape)
        pose = pose + cfg.step_size * d_pose
        shape = shape + cfg.step_size * d_shape

    # ---- IterativeErrorFeedback.backward (models.py:634-688) ----
    n_hist = cfg.num_iterations + 1
    pose_t = torch.zeros((), dtype=dt)
    shape_t = torch.zeros((), dtype=dt)
    recon_t = torch.zeros((), dtype=dt)
    fk_t = torch.zeros((), dtype=dt)
    shapes_rep = shapes_gt.to(dt).unsqueeze(1).repeat(1, n_frames, 1)
    for i in range(n_hist):
        pose_t = pose_t + padded_l1(poses_gt.to(dt), hist['pose'][i].reshape(bsz, n_frames, -1), seq_lengths)
        shape_t = shape_t + padded_l1(shapes_rep, hist['shape'][i].reshape(bsz, n_frames, -1), seq_lengths)
        if fk_weight > 0.0:                                                      # always the FINAL joints (models.py:657-660)
            fk_t = fk_t + ief.reconstruction_energy(joints_gt.to(dt).reshape(bsz, n_frames, -1, 3),
                                                    hist['joints'][-1].reshape(bsz, n_frames, -1, 3), seq_lengths,
                                                    marker_masks)
        recon_t = recon_t + recon(hist['markers'][i], hist['markers_ori'][i])
    total = (pose_weight * pose_t + fk_weight * fk_t + shape_weight * shape_t + r_weight * recon_t) / n_hist
    accumulate(total, False)
    loss_vals = {'pose': float(pose_t.detach()) / n_hist, 'shape': float(shape_t.detach()) / n_hist,
                 'reconstruction': float(recon_t.detach()) / n_hist, 'fk': float(fk_t.detach()) / n_hist,
                 'total_loss': float(total.detach())}
    last_pose = hist['pose'][-1].detach().reshape(bsz, n_frames, -1)
    return {'loss_vals': loss_vals, 'grads': grads, 'buffers': buffers,
            'pose_hat': last_pose[:, :, 3:], 'root_ori_hat': last_pose[:, :, :3],
            'shape_hat': hist['shape'][-1].detach().reshape(bsz, n_frames, -1),
            'joints_hat': hist['joints'][-1].detach().reshape(bsz, n_frames, -1),
            'history': {k: [h.detach().reshape(bsz, n_frames, -1) for h in v] for k, v in hist.items()}}
