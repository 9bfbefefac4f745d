"""Data synthesis on the B200: empose_b200.data.transforms.SMPLFK / SampleMarkersWithOffsets vs the golden
outputs of the unmodified reference transforms (tests/golden/data_synthesis.npz).  Run on the B200 box: -m gpu."""
import os
import tempfile

import numpy as np
import pytest
import torch

from empose_b200 import lib as native
from empose_b200 import synthetic

import util

pytestmark = pytest.mark.gpu


class Batch(object):
    """The attributes of the reference's AMASSBatch the two transforms read (data.py:386-431)."""

    def __init__(self, poses, shapes, trans, dev):
        t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)
        poses = t(poses)
        self.poses_root, self.poses_body = poses[:, :, :3].contiguous(), poses[:, :, 3:].contiguous()
        self.shapes, self.trans = t(shapes), t(trans)
        self.batch_size, self.seq_length = poses.shape[0], poses.shape[1]


@pytest.mark.parametrize('precision', [native.PRECISION_FP32, native.PRECISION_TF32], ids=['fp32', 'tf32'])
def test_data_synthesis_matches_reference_transforms(smpl_npz, precision):
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.data.transforms import SMPLFK, SampleMarkersWithOffsets
    assert torch.cuda.is_available()
    dev = torch.device('cuda:0')
    gold = util.load_golden('data_synthesis')
    smpl = SMPLLayer(smpl_npz).to(device=dev, dtype=torch.float32)
    smpl.precision = precision
    files = synthetic.write_synthetic_offsets(os.path.join(tempfile.gettempdir(), 'empose_b200_assets'), n_files=3, seed=0)
    worst = {}
    for level in (-1, 0, 1, 3):
        torch.manual_seed(1234 + level)
        batch = Batch(gold['poses'], gold['shapes'], gold['trans'], dev)
        batch = SMPLFK(smpl)(batch)
        sampler = SampleMarkersWithOffsets(smpl, files, noise_level=level)
        batch = sampler(batch)
        batch = sampler(batch)
        assert batch.vertices is None                    # the mesh is never materialised
        tag = 'n%d_' % level
        for k, tol in (('joints_gt', 1e-5), ('marker_pos_vertex', 1e-5), ('marker_pos_synth', 2e-5), ('marker_ori_vertex', 2e-4),
                       ('marker_ori_synth', 2e-4), ('marker_normal_vertex', 2e-4), ('marker_normal_synth', 2e-4),
                       ('offset_t_augmented', 1e-7), ('offset_r_augmented', 1e-6)):
            got = getattr(batch, k).cpu().numpy()
            want = gold[tag + k].reshape(got.shape)
            if k == 'marker_normal_vertex':
                # the reference keeps the UN-normalised area-weighted vertex normal here (virtual_sensors.py:92-96 returns
                # get_vertex_normals as is; no model consumes it, models.py:122-123); we store the unit normal
                w3 = want.reshape(want.shape[:2] + (12, 3))
                want = (w3 / np.linalg.norm(w3, axis=-1, keepdims=True)).reshape(got.shape)
            err = float(np.abs(got - want).max())
            worst[k] = max(worst.get(k, 0.0), err)
            assert err <= tol, (level, k, err)
    util.report('data_synthesis', precision='fp32' if precision == native.PRECISION_FP32 else 'tf32', **worst)
    # with keep_vertices the full-mesh layer runs and joints agree with the sub-model route
    b1 = SMPLFK(smpl, keep_vertices=True)(Batch(gold['poses'], gold['shapes'], gold['trans'], dev))
    assert b1.vertices.shape == (4, 5, 6890 * 3)
    np.testing.assert_allclose(b1.joints_gt.cpu().numpy(), gold['n-1_joints_gt'], atol=1e-5, rtol=0)
