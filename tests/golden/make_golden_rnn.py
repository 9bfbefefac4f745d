"""Generate the (Bi)RNN golden fixtures by running the UNMODIFIED reference ``SimpleRNN`` (models.py:265-317).

    python tests/golden/make_golden_rnn.py
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

CASES = {
    # released-model shape: BiRNN, 2 x 1024, 12 sensors, shape head with FK (one window, two chunks with state carry)
    'rnn_bi12_shape_fk': (
        ['--m_type', 'rnn', '--m_hidden_size', '1024', '--m_num_layers', '2', '--m_bidirectional', '--m_estimate_shape',
         '--m_average_shape', '--m_fk_loss', '0.1', '--use_marker_pos', '--use_marker_ori', '--n_markers', '12', '--window_size', '32'],
        dict(n_markers=12, hidden_size=1024, num_layers=2, bidirectional=True, estimate_shape=True),
        dict(kind='amass', B=3, F=7, seed=41, ragged=True, offsets=True, chunks=2)),
    'rnn_bi6': (
        ['--m_type', 'rnn', '--m_hidden_size', '256', '--m_num_layers', '2', '--m_bidirectional', '--use_marker_pos',
         '--use_marker_ori', '--n_markers', '6', '--window_size', '32'],
        dict(n_markers=6, hidden_size=256, num_layers=2, bidirectional=True, estimate_shape=False),
        dict(kind='real', B=4, F=9, seed=42, ragged=True, offsets=True, drop_rate=0.1, chunks=1)),
    'rnn_uni12': (
        ['--m_type', 'rnn', '--m_hidden_size', '128', '--m_num_layers', '3', '--use_marker_pos', '--use_marker_ori',
         '--n_markers', '12', '--window_size', '32'],
        dict(n_markers=12, hidden_size=128, num_layers=3, bidirectional=False, estimate_shape=False),
        dict(kind='amass', B=2, F=5, seed=43, ragged=False, offsets=False, chunks=2)),
}


def run_case(name, flags, weight_kwargs, spec, smpl_layer, out_dir):
    from empose.nn.models import create_model
    config = ref_shims.make_config(flags)
    net = create_model(config, smpl_layer)
    state = net.state_dict()
    synth = synthetic.synth_rnn_state_dict(seed=0, **weight_kwargs)
    missing = [k for k in state if not k.startswith('smpl.') and k not in synth]
    extra = [k for k in synth if k not in state]
    assert not missing and not extra, (missing, extra)
    for k, v in synth.items():
        assert tuple(state[k].shape) == tuple(v.shape), (k, state[k].shape, v.shape)
        state[k] = torch.from_numpy(np.asarray(v))
    net.load_state_dict(state, strict=True)
    net.eval()
    record = {}
    for c in range(spec['chunks']):
        params = synthetic.synth_window_params(spec['B'], spec['F'], seed=spec['seed'] + 100 * c, ragged=spec['ragged'],
                                               offsets=spec['offsets'], drop_rate=spec.get('drop_rate', 0.0))
        if c > 0:
            params['offset_t'], params['offset_r'] = first['offset_t'], first['offset_r']
        else:
            first = params
        gt_pos, gt_ori, joints_gt, _, _ = mg.project_ground_truth(smpl_layer, params)
        marker_pos, marker_ori = synthetic.synth_measurements(gt_pos, gt_ori, seed=spec['seed'] + 100 * c)
        batch = mg.make_batch(spec, params, marker_pos, marker_ori, joints_gt)
        inputs = batch.get_inputs()
        with torch.no_grad():
            out = net(batch, is_new_sequence=(c == 0))
        tag = 'c%d_' % c
        record[tag + 'marker_pos'] = inputs['marker_pos'].detach().numpy()
        record[tag + 'marker_oris'] = inputs['marker_oris'].detach().numpy()
        record[tag + 'seq_lengths'] = params['seq_lengths']
        for k in ('pose_hat', 'root_ori_hat', 'shape_hat', 'joints_hat'):
            if out[k] is not None:
                record[tag + k] = out[k].detach().numpy()
        record[tag + 'final_h'] = net.rnn.final_state[0].detach().numpy()
        record[tag + 'final_c'] = net.rnn.final_state[1].detach().numpy()
    record['n_trainable_params'] = np.asarray(sum(p.numel() for p in net.parameters() if p.requires_grad))
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), **{k: np.asarray(v) for k, v in record.items()})
    print('%-20s params=%d keys=%d name=%s' % (name, int(record['n_trainable_params']), len(record), net.model_name()))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    asset_dir = os.path.join(tempfile.gettempdir(), 'empose_b200_assets')
    ref_shims.install(asset_dir, seed=mg.SMPL_SEED)
    from empose.bodymodels.smpl import create_default_smpl_model
    smpl_layer = create_default_smpl_model(device='cpu')
    for name, (flags, wk, spec) in CASES.items():
        run_case(name, flags, wk, spec, smpl_layer, HERE)


if __name__ == '__main__':
    main()
