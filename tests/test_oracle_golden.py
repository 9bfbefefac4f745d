"""The oracle restatement vs outputs of the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from empose_b200 import synthetic
from oracle import ief as oracle_ief
from oracle import sensors
from oracle import smplh_lbs

import util


@pytest.mark.parametrize('name', sorted(util.GOLDEN_CASES))
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_ief_matches_reference(name, dtype, oracle_smpl, topology):
    gold = util.load_golden(name)
    flags = util.GOLDEN_CASES[name]
    cfg = oracle_ief.IefConfig(**flags)
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=flags['n_markers'],
                                                          rnn_init=flags['rnn_init']), dtype)
    state = None
    c = 0
    while ('c%d_pose_hat' % c) in gold:
        inp = util.chunk_inputs(gold, c, dtype)
        out = oracle_ief.ief_forward(cfg, sd, oracle_smpl, topology, init_state=state, **inp)
        state = out['final_state']
        tag = 'c%d_' % c
        live = util.valid_frame_mask(gold[tag + 'seq_lengths'], inp['marker_pos'].shape[1])
        # final outputs: parity bar of the task is 1e-4 rad / 0.1 mm; the restatement is far inside it
        for k, tol in (('pose_hat', 2e-5), ('root_ori_hat', 2e-5), ('shape_hat', 2e-5), ('joints_hat', 2e-6)):
            got = out[k].detach().numpy()
            np.testing.assert_allclose(got[live], gold[tag + k][live], atol=tol, rtol=0, err_msg=k)
        # every iterate
        for hk, gk in (('pose', 'pose_hat_history'), ('shape', 'shape_hat_history'), ('joints', 'joints_hat_history'),
                       ('markers', 'markers_hat_history'), ('markers_ori', 'markers_ori_hat_history')):
            got = np.stack([h.numpy() for h in out['history'][hk]])
            assert got.shape == gold[tag + gk].shape
            np.testing.assert_allclose(got[:, live], gold[tag + gk][:, live], atol=2e-5, rtol=0, err_msg=gk)
        if cfg.rnn_init:
            np.testing.assert_allclose(state[0].numpy(), gold[tag + 'final_h'], atol=2e-6, rtol=0)
            np.testing.assert_allclose(state[1].numpy(), gold[tag + 'final_c'], atol=2e-6, rtol=0)
        c += 1
    assert c >= 1


def test_param_count_known_answer():
    """README.md:228 of the reference: 'Model created with 5721419 trainable parameters' (LGD-RNN-6, N=2)."""
    gold = util.load_golden('lgd_rnn6_n2_real')
    assert int(gold['n_trainable_params']) == 5721419
    sd = synthetic.synth_state_dict(seed=0, n_markers=6, rnn_init=True)
    learned = sum(v.size for k, v in sd.items() if 'running_' not in k and 'num_batches' not in k)
    assert learned + 169 == 5721419          # + the 169 SMPL 'parameters' of the third-party BodyModel


def test_smpl_and_sensor_frames_match_reference(oracle_smpl, topology):
    gold = util.load_golden('smpl_sensors')
    poses = torch.from_numpy(gold['poses']).reshape(-1, 66).double()
    shapes = torch.from_numpy(gold['shapes']).double().repeat(poses.shape[0], 1)
    verts, joints = smplh_lbs.smpl_layer_forward(oracle_smpl, poses[:, 3:], shapes, poses_root=poses[:, :3])
    np.testing.assert_allclose(verts.numpy(), gold['verts'], atol=2e-6, rtol=0)
    np.testing.assert_allclose(joints.numpy(), gold['joints'], atol=2e-6, rtol=0)
    pos, ori, _ = sensors.sensor_frames(verts, topology)
    f = poses.shape[0]
    off_r = torch.from_numpy(gold['offset_r']).double().repeat(f, 1, 1, 1)
    off_t = torch.from_numpy(gold['offset_t']).double().repeat(f, 1, 1)
    pos_c, ori_c = sensors.apply_offsets(pos, ori, off_r, off_t)
    np.testing.assert_allclose(pos_c.numpy(), gold['sensor_pos'].reshape(f, 12, 3), atol=2e-6, rtol=0)
    np.testing.assert_allclose(ori_c.numpy(), gold['sensor_ori'].reshape(f, 12, 3, 3), atol=2e-5, rtol=0)
