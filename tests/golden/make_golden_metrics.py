"""Golden vectors for the metrics engine: the UNMODIFIED reference ``MetricsEngine`` (empose/eval/metrics.py:69-330) on
synthetic predictions.  The absent ``quaternion`` package is stood in by scipy rotations (oracle/ref_shims.py); the
Euclidean and Procrustes parts are the reference's own numpy code.

    python tests/golden/make_golden_metrics.py
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

B, F, SEED = 5, 6, 61


def main():
    torch.set_num_threads(2)
    asset_dir = os.path.join(tempfile.gettempdir(), 'empose_b200_assets')
    ref_shims.install(asset_dir, seed=mg.SMPL_SEED)
    from empose.bodymodels.smpl import create_default_smpl_model
    from empose.eval.metrics import MetricsEngine
    smpl_layer = create_default_smpl_model(device='cpu')
    gt = synthetic.synth_window_params(B, F, seed=SEED, ragged=True, drop_rate=0.1)
    rng = np.random.RandomState(SEED)
    pose_hat = gt['poses'] + (0.08 * rng.standard_normal(gt['poses'].shape)).astype(np.float32)
    shape_hat = np.repeat(gt['shapes'][:, None], F, axis=1) + (0.3 * rng.standard_normal((B, F, 10))).astype(np.float32)
    t = torch.from_numpy
    record = {'poses': gt['poses'], 'shapes': gt['shapes'], 'pose_hat': pose_hat, 'shape_hat': shape_hat,
              'seq_lengths': gt['seq_lengths'], 'marker_masks': gt['marker_masks']}
    me = MetricsEngine(smpl_layer)
    with torch.no_grad():
        # evaluate_real.py / eval/helpers.py style call: body pose, shape per window, predicted shape per frame, root poses, masks
        me.compute(t(gt['poses'][:, :, 3:]), t(gt['shapes']), t(pose_hat[:, :, 3:]), t(shape_hat), t(gt['seq_lengths']),
                   pose_root=t(gt['poses'][:, :, :3]), pose_root_hat=t(pose_hat[:, :, :3]), frame_mask=t(gt['marker_masks']))
        me.compute(t(gt['poses'][:, :, 3:]), t(gt['shapes']), t(pose_hat[:, :, 3:]), None, None)      # train.py:158 style
    record['eucl'] = np.concatenate(me.eucl_dists, axis=0)
    record['eucl_pa'] = np.concatenate(me.eucl_dists_pa, axis=0)
    record['angle'] = np.concatenate(me.angle_diffs, axis=0)
    metrics = me.get_metrics()
    for k, v in metrics.items():
        record['m/' + k] = np.asarray(v, dtype=np.float64)
    record['pretty'] = np.asarray(me.to_pretty_string(metrics, 'golden'))
    np.savez_compressed(os.path.join(HERE, 'metrics.npz'), **record)
    print(metrics, record['eucl'].shape, record['angle'].shape)
    print(me.to_pretty_string(metrics, 'golden'))


if __name__ == '__main__':
    main()
