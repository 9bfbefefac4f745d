"""Golden vectors for the data-synthesis transforms: the UNMODIFIED reference ``SMPLFK`` + ``SampleMarkersWithOffsets``
(empose/data/transforms.py:163-226, 259-282) on a synthetic AMASS-shaped batch with stand-in offset files.

    python tests/golden/make_golden_synth.py
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

B, F, SEED = 4, 5, 51
NOISE_LEVELS = (-1, 0, 1, 3)


def make_batch():
    from empose.data.data import AMASSBatch
    params = synthetic.synth_window_params(B, F, seed=SEED)
    rng = np.random.RandomState(SEED)
    trans = (0.3 * rng.standard_normal((B, F, 3))).astype(np.float32)
    return AMASSBatch(list(range(B)), torch.from_numpy(params['seq_lengths']), torch.from_numpy(params['poses']),
                      torch.from_numpy(params['shapes']), torch.from_numpy(trans), torch.zeros(B, F, 66)), params, trans


def main():
    torch.set_num_threads(2)
    asset_dir = os.path.join(tempfile.gettempdir(), 'empose_b200_assets')
    ref_shims.install(asset_dir, seed=mg.SMPL_SEED)
    from empose.bodymodels.smpl import create_default_smpl_model
    from empose.data.transforms import SMPLFK, SampleMarkersWithOffsets
    smpl_layer = create_default_smpl_model(device='cpu')
    files = synthetic.write_synthetic_offsets(asset_dir, n_files=3, seed=0)
    record = {}
    for level in NOISE_LEVELS:
        torch.manual_seed(1234 + level)
        batch, params, trans = make_batch()
        with torch.no_grad():
            batch = SMPLFK(smpl_layer)(batch)
            sampler = SampleMarkersWithOffsets(smpl_layer, files, noise_level=level)
            batch = sampler(batch)
            batch = sampler(batch)                      # second call: the offset RandomState stream advances
        tag = 'n%d_' % level
        for k in ('joints_gt', 'marker_pos_vertex', 'marker_ori_vertex', 'marker_normal_vertex', 'marker_pos_synth',
                  'marker_ori_synth', 'marker_normal_synth', 'offset_t_augmented', 'offset_r_augmented'):
            record[tag + k] = getattr(batch, k).detach().numpy().astype(np.float32)
    record['poses'], record['shapes'], record['trans'] = params['poses'], params['shapes'], trans
    np.savez_compressed(os.path.join(HERE, 'data_synthesis.npz'), **record)
    print('data_synthesis', {k: v.shape for k, v in record.items() if k.startswith('n0_')})


if __name__ == '__main__':
    main()
