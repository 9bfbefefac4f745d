"""The (Bi)RNN restatement (oracle/rnn.py) vs outputs of the unmodified reference SimpleRNN (tests/golden/rnn_*.npz)."""
import numpy as np
import pytest
import torch

from empose_b200 import synthetic
from oracle import rnn as oracle_rnn

import util


@pytest.mark.parametrize('name', sorted(util.RNN_CASES))
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_rnn_matches_reference(name, dtype, oracle_smpl):
    gold = util.load_golden(name)
    flags = util.RNN_CASES[name]
    cfg = oracle_rnn.RnnConfig(**flags['cfg'])
    sd = util.torch_state_dict(synthetic.synth_rnn_state_dict(seed=0, **flags['weights']), dtype)
    state, c = None, 0
    while ('c%d_pose_hat' % c) in gold:
        tag = 'c%d_' % c
        pos = torch.from_numpy(gold[tag + 'marker_pos']).to(dtype)
        ori = torch.from_numpy(gold[tag + 'marker_oris']).to(dtype)
        lens = torch.from_numpy(gold[tag + 'seq_lengths'])
        out = oracle_rnn.rnn_forward(cfg, sd, oracle_smpl, pos, ori, lens, init_state=state)
        state = out['final_state']
        live = util.valid_frame_mask(gold[tag + 'seq_lengths'], pos.shape[1])
        for k, tol in (('pose_hat', 2e-5), ('root_ori_hat', 2e-5), ('shape_hat', 2e-5), ('joints_hat', 5e-6)):
            if (tag + k) in gold:
                np.testing.assert_allclose(out[k].detach().numpy()[live], gold[tag + k][live], atol=tol, rtol=0, err_msg=k)
            else:
                assert out[k] is None, k
        np.testing.assert_allclose(state[0].numpy(), gold[tag + 'final_h'], atol=5e-6, rtol=0)
        np.testing.assert_allclose(state[1].numpy(), gold[tag + 'final_c'], atol=5e-6, rtol=0)
        c += 1
    assert c >= 1
    learned = sum(v.size for v in synthetic.synth_rnn_state_dict(seed=0, **flags['weights']).values())
    assert learned + 169 == int(gold['n_trainable_params'])      # + the 169 SMPL 'parameters' of the third-party BodyModel
