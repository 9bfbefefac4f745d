"""Generate the TRAINING golden fixtures by running the UNMODIFIED reference in train mode.

    python tests/golden/make_golden_train.py

For each case: ``net.train()``; ``optimizer.zero_grad()``; ``out = net(batch)``;
``net.backward(batch, out)`` exactly as ``scripts/train.py:136-149`` does, then record the loss
values, the BatchNorm running statistics after the step and the parameter gradients -- including the
gradients the forward pass itself leaves in ``.grad`` (``reconstruction_error.backward`` at
``empose/nn/models.py:576`` runs N times per forward and reaches every upstream parameter).

The full gradient of a case is 23 MB, so the fixture keeps, for every parameter tensor, its L2 norm,
its sum and 192 entries at seeded positions (small tensors are kept whole).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from empose_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402
import make_golden as mg  # noqa: E402

N_SAMPLES = 192

CASES = {
    'train_lgd_rnn12_n4': (
        ['--m_type', 'lgd', '--m_num_iterations', '4', '--m_hidden_size', '512', '--m_rnn_init', '--m_average_shape',
         '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '12', '--window_size', '32',
         '--m_reprojection_loss_weight', '0.01', '--m_fk_loss', '0.1', '--m_pose_loss_weight', '10.0'],
        dict(n_markers=12, rnn_init=True, hidden_size=512),
        dict(kind='amass', B=5, F=8, seed=31, ragged=True, offsets=True)),
    'train_lgd_mlp12_n2': (
        ['--m_type', 'lgd', '--m_num_iterations', '2', '--m_hidden_size', '512', '--m_average_shape',
         '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '12', '--window_size', '32'],
        dict(n_markers=12, rnn_init=False, hidden_size=512),
        dict(kind='amass', B=6, F=4, seed=32, ragged=False, offsets=False)),
    'train_lgd_rnn6_n2_real': (
        ['--m_type', 'lgd', '--m_num_iterations', '2', '--m_hidden_size', '512', '--m_rnn_init', '--m_average_shape',
         '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '6', '--window_size', '32',
         '--m_fk_loss', '0.1', '--m_pose_loss_weight', '10.0'],
        dict(n_markers=6, rnn_init=True, hidden_size=512),
        dict(kind='real', B=4, F=6, seed=33, ragged=True, offsets=True, drop_rate=0.1)),
}


def sample_positions(name, numel):
    """Seeded positions kept for a tensor (all of them when it is small)."""
    if numel <= N_SAMPLES:
        return np.arange(numel)
    seed = int.from_bytes(name.encode()[-4:].rjust(4, b'\0'), 'little') ^ (numel & 0xFFFF)
    return np.sort(np.random.RandomState(seed % (2 ** 31)).choice(numel, N_SAMPLES, replace=False))


def run_case(name, flags, weight_kwargs, spec, smpl_layer, out_dir):
    net = mg.build_reference_model(flags, weight_kwargs, smpl_layer)
    net.train()
    params = synthetic.synth_window_params(spec['B'], spec['F'], seed=spec['seed'], ragged=spec['ragged'],
                                           offsets=spec['offsets'], drop_rate=spec.get('drop_rate', 0.0))
    gt_pos, gt_ori, joints_gt, _, _ = mg.project_ground_truth(smpl_layer, params)
    marker_pos, marker_ori = synthetic.synth_measurements(gt_pos, gt_ori, seed=spec['seed'])
    batch = mg.make_batch(spec, params, marker_pos, marker_ori, joints_gt)
    if spec['kind'] == 'real':
        batch.joints_gt = torch.from_numpy(joints_gt)
    inputs = batch.get_inputs()
    for p in net.parameters():
        p.grad = None
    out = net(batch)
    total, loss_vals = net.backward(batch, out)
    record = {'marker_pos': inputs['marker_pos'].detach().numpy(), 'marker_oris': inputs['marker_oris'].detach().numpy(),
              'offset_t': inputs['offset_t'].numpy(), 'offset_r': inputs['offset_r'].numpy(),
              'seq_lengths': params['seq_lengths'], 'poses': params['poses'], 'shapes': params['shapes'],
              'joints_gt': joints_gt}
    if inputs['marker_masks'] is not None:
        record['marker_masks'] = inputs['marker_masks'].numpy()
    for k, v in loss_vals.items():
        record['loss_' + k] = np.asarray(v, dtype=np.float64)
    record['pose_hat'] = out['pose_hat'].detach().numpy()
    record['shape_hat'] = out['shape_hat'].detach().numpy()
    record['joints_hat'] = out['joints_hat'].detach().numpy()
    n_grad = 0
    for pname, p in net.named_parameters():
        if pname.startswith('smpl.'):
            continue
        g = p.grad.detach().numpy().astype(np.float64).reshape(-1)
        pos = sample_positions(pname, g.size)
        record['g/' + pname + '/norm'] = np.asarray(np.sqrt((g * g).sum()))
        record['g/' + pname + '/sum'] = np.asarray(g.sum())
        record['g/' + pname + '/samples'] = g[pos].astype(np.float32)
        n_grad += 1
    for bname, b in net.named_buffers():
        if 'running_' in bname or 'num_batches' in bname:
            record['b/' + bname] = b.detach().numpy()
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), **{k: np.asarray(v) for k, v in record.items()})
    print('%-26s grads=%d total_loss=%.6f' % (name, n_grad, loss_vals['total_loss']), loss_vals)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)   # with 4 threads torch 2.11 CPU autograd returned a 5 % different PReLU-slope gradient (1 and 8 agree with float64)
    asset_dir = os.path.join(tempfile.gettempdir(), 'empose_b200_assets')
    ref_shims.install(asset_dir, seed=mg.SMPL_SEED)
    from empose.bodymodels.smpl import create_default_smpl_model
    smpl_layer = create_default_smpl_model(device='cpu')
    for name, (flags, wk, spec) in CASES.items():
        run_case(name, flags, wk, spec, smpl_layer, HERE)


if __name__ == '__main__':
    main()
