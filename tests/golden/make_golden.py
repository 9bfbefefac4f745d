"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` exists):

    python tests/golden/make_golden.py

It imports the reference's own ``empose.nn.models`` / ``empose.bodymodels.smpl`` /
``empose.data.data`` through ``oracle.ref_shims`` (stand-ins only for the three absent third-party
imports), builds models with ``create_model``, loads the deterministic synthetic weights of
``empose_b200.synthetic.synth_state_dict`` and records inputs and outputs as small ``.npz`` files.
The SMPL-H model and the weights are NOT stored: both are regenerated bit-exactly from their seeds
by the tests (numpy ``RandomState`` streams).  ``/root/reference`` does not exist on the GPU box,
which is why these vectors are committed.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from empose_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402

SMPL_SEED = 0

CASES = {
    # name: (reference CLI flags, synth_state_dict kwargs, batch spec)
    'lgd_rnn12_n4': (
        ['--m_type', 'lgd', '--m_num_iterations', '4', '--m_hidden_size', '512', '--m_rnn_init', '--m_average_shape',
         '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '12', '--window_size', '32'],
        dict(n_markers=12, rnn_init=True, hidden_size=512),
        dict(kind='amass', B=3, F=8, seed=11, ragged=True, offsets=True)),
    'lgd_mlp12_n4': (
        ['--m_type', 'lgd', '--m_num_iterations', '4', '--m_hidden_size', '512', '--m_average_shape',
         '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '12', '--window_size', '32'],
        dict(n_markers=12, rnn_init=False, hidden_size=512),
        dict(kind='amass', B=4, F=4, seed=12, ragged=False, offsets=False)),
    'lgd_rnn6_n2_real': (
        ['--m_type', 'lgd', '--m_num_iterations', '2', '--m_hidden_size', '512', '--m_rnn_init', '--m_average_shape',
         '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '6', '--window_size', '32',
         '--m_fk_loss', '0.1', '--m_pose_loss_weight', '10.0'],
        dict(n_markers=6, rnn_init=True, hidden_size=512),
        dict(kind='real', B=2, F=6, seed=13, ragged=True, offsets=True, drop_rate=0.15, chunks=2)),
}


def build_reference_model(flags, weight_kwargs, smpl_layer):
    from empose.nn.models import create_model
    config = ref_shims.make_config(flags)
    net = create_model(config, smpl_layer)
    state = net.state_dict()
    synth = synthetic.synth_state_dict(seed=0, **weight_kwargs)
    missing = [k for k in state if not k.startswith('smpl.') and k not in synth]
    extra = [k for k in synth if k not in state]
    assert not missing and not extra, (missing, extra)
    for k, v in synth.items():
        assert tuple(state[k].shape) == tuple(v.shape), (k, state[k].shape, v.shape)
        state[k] = torch.from_numpy(np.asarray(v))
    net.load_state_dict(state, strict=True)
    net.eval()
    return net


def project_ground_truth(smpl_layer, params):
    """Reference SMPL + reference sensor frames + offsets on the ground-truth poses -> clean sensors."""
    from empose.data.virtual_sensors import VirtualMarkerHelper
    from empose.helpers.configuration import CONSTANTS as C
    b, f = params['poses'].shape[:2]
    poses = torch.from_numpy(params['poses']).reshape(b * f, 66)
    shapes = torch.from_numpy(params['shapes']).unsqueeze(1).repeat(1, f, 1).reshape(b * f, 10)
    with torch.no_grad():
        verts, joints = smpl_layer(poses_body=poses[:, 3:], betas=shapes, poses_root=poses[:, :3])
        pos, ori, _ = VirtualMarkerHelper(smpl_layer).get_virtual_pos_and_rot(verts, C.VERTEX_IDS)
        off_r = torch.from_numpy(params['offset_r']).unsqueeze(1).repeat(1, f, 1, 1, 1).reshape(b * f, 12, 3, 3)
        off_t = torch.from_numpy(params['offset_t']).unsqueeze(1).repeat(1, f, 1, 1).reshape(b * f, 12, 3)
        ori_c = torch.matmul(ori, off_r)
        pos_c = pos + torch.matmul(ori, off_t.unsqueeze(-1)).squeeze(-1)
    return (pos_c.reshape(b, f, 12, 3).numpy(), ori_c.reshape(b, f, 12, 3, 3).numpy(),
            joints[:, :22].reshape(b, f, 66).numpy(), verts, joints)


def make_batch(spec, params, marker_pos, marker_ori, joints_gt):
    from empose.data.data import AMASSBatch, RealBatch
    b, f = spec['B'], spec['F']
    lengths = torch.from_numpy(params['seq_lengths'])
    poses = torch.from_numpy(params['poses'])
    shapes = torch.from_numpy(params['shapes'])
    trans = torch.zeros(b, f, 3)
    if spec['kind'] == 'amass':
        batch = AMASSBatch(list(range(b)), lengths, poses, shapes, trans, torch.from_numpy(joints_gt))
        batch.marker_pos_synth = torch.from_numpy(marker_pos)
        batch.marker_ori_synth = torch.from_numpy(marker_ori)
        batch.offset_t_augmented = torch.from_numpy(params['offset_t'])
        batch.offset_r_augmented = torch.from_numpy(params['offset_r'])
    else:
        batch = RealBatch(list(range(b)), lengths.to(torch.int32), poses, shapes, trans,
                          torch.from_numpy(marker_pos), torch.from_numpy(marker_ori),
                          torch.from_numpy(params['marker_masks']),
                          offset_t=torch.from_numpy(params['offset_t']), offset_r=torch.from_numpy(params['offset_r']))
        batch.joints_hat = torch.from_numpy(joints_gt)
    return batch


def run_case(name, flags, weight_kwargs, spec, smpl_layer, out_dir):
    net = build_reference_model(flags, weight_kwargs, smpl_layer)
    chunks = spec.get('chunks', 1)
    record = {}
    for c in range(chunks):
        params = synthetic.synth_window_params(spec['B'], spec['F'], seed=spec['seed'] + 100 * c, ragged=spec['ragged'],
                                               offsets=spec['offsets'], drop_rate=spec.get('drop_rate', 0.0))
        if c > 0:   # one recording session: subject-specific offsets stay fixed across chunks
            params['offset_t'], params['offset_r'] = first['offset_t'], first['offset_r']
        else:
            first = params
        gt_pos, gt_ori, joints_gt, _, _ = project_ground_truth(smpl_layer, params)
        marker_pos, marker_ori = synthetic.synth_measurements(gt_pos, gt_ori, seed=spec['seed'] + 100 * c)
        batch = make_batch(spec, params, marker_pos, marker_ori, joints_gt)
        inputs = batch.get_inputs()           # for real batches this also zeroes the dropped sensors (data.py:283-307)
        out = net(batch, is_new_sequence=(c == 0))
        tag = 'c%d_' % c
        record[tag + 'marker_pos'] = inputs['marker_pos'].detach().numpy()
        record[tag + 'marker_oris'] = inputs['marker_oris'].detach().numpy()
        record[tag + 'offset_t'] = inputs['offset_t'].numpy()
        record[tag + 'offset_r'] = inputs['offset_r'].numpy()
        record[tag + 'seq_lengths'] = params['seq_lengths']
        if inputs['marker_masks'] is not None:
            record[tag + 'marker_masks'] = inputs['marker_masks'].numpy()
        for k in ('pose_hat', 'root_ori_hat', 'shape_hat', 'joints_hat'):
            record[tag + k] = out[k].detach().numpy()
        for hname in ('pose_hat_history', 'shape_hat_history', 'joints_hat_history', 'markers_hat_history',
                      'markers_ori_hat_history'):
            hist = getattr(net, hname)
            record[tag + hname] = np.stack([h.detach().reshape(spec['B'], spec['F'], -1).numpy() for h in hist])
        if net.rnn_init:
            record[tag + 'final_h'] = net.rnn.final_state[0].detach().numpy()
            record[tag + 'final_c'] = net.rnn.final_state[1].detach().numpy()
    n_params = sum(p.numel() for p in net.parameters() if p.requires_grad)
    record['n_trainable_params'] = np.asarray(n_params)
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), **{k: np.asarray(v) for k, v in record.items()})
    print('%-20s params=%d  keys=%d' % (name, n_params, len(record)))


def run_smpl_kat(smpl_layer, out_dir):
    """Two frames through the reference SMPLLayer wrapper + sensor helper (verts kept for the 12 sensors' 1-rings)."""
    params = synthetic.synth_window_params(1, 2, seed=21, offsets=True)
    gt_pos, gt_ori, joints22, verts, joints = project_ground_truth(smpl_layer, params)
    np.savez_compressed(os.path.join(out_dir, 'smpl_sensors.npz'),
                        poses=params['poses'], shapes=params['shapes'], offset_t=params['offset_t'],
                        offset_r=params['offset_r'], verts=verts.numpy().astype(np.float32),
                        joints=joints.numpy().astype(np.float32), sensor_pos=gt_pos, sensor_ori=gt_ori)
    print('smpl_sensors         verts', tuple(verts.shape))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    asset_dir = os.path.join(tempfile.gettempdir(), 'empose_b200_assets')
    ref_shims.install(asset_dir, seed=SMPL_SEED)
    from empose.bodymodels.smpl import create_default_smpl_model
    smpl_layer = create_default_smpl_model(device='cpu')
    for name, (flags, wk, spec) in CASES.items():
        run_case(name, flags, wk, spec, smpl_layer, HERE)
    run_smpl_kat(smpl_layer, HERE)


if __name__ == '__main__':
    main()
