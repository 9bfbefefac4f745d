"""The training-step restatement (oracle/train.py) vs the unmodified reference in train mode
(tests/golden/train_*.npz, written by tests/golden/make_golden_train.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from empose_b200 import synthetic
from oracle import ief as oracle_ief
from oracle import train as oracle_train

import util

sys.path.insert(0, util.GOLDEN_DIR)
from make_golden_train import sample_positions  # noqa: E402


def run_oracle_case(name, dtype, oracle_smpl, topology):
    gold = util.load_golden(name)
    flags = util.TRAIN_CASES[name]
    cfg = oracle_ief.IefConfig(n_markers=flags['n_markers'], num_iterations=flags['num_iterations'],
                               rnn_init=flags['rnn_init'])
    sd = util.torch_state_dict(synthetic.synth_state_dict(seed=0, n_markers=flags['n_markers'],
                                                          rnn_init=flags['rnn_init']), dtype)
    inp = util.train_inputs(gold, dtype)
    out = oracle_train.ief_train_step(cfg, sd, oracle_smpl, topology, pose_weight=flags['pose_weight'],
                                      shape_weight=1.0, r_weight=0.01, fk_weight=flags['fk_weight'], **inp)
    return gold, out


@pytest.mark.parametrize('name', sorted(util.TRAIN_CASES))
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_train_step_matches_reference(name, dtype, oracle_smpl, topology):
    gold, out = run_oracle_case(name, dtype, oracle_smpl, topology)
    # the reference ran in float32; scalar reductions over all rows (PReLU slopes) carry ~3e-3 of float32
    # cancellation noise, which the float64 restatement does not share
    rel = 3e-4 if dtype == torch.float32 else 5e-3
    for k, v in out['loss_vals'].items():
        assert abs(v - float(gold['loss_' + k])) <= 2e-5 * max(1.0, abs(v)), (k, v, float(gold['loss_' + k]))
    n = 0
    for key, g in out['grads'].items():
        want_norm = float(gold['g/' + key + '/norm'])
        flat = g.reshape(-1).double().numpy()
        pos = sample_positions(key, flat.size)
        # the reference ran in float32: compare at float32 accuracy relative to the tensor's own scale
        # (floor: biases in front of a BatchNorm have an analytically zero gradient; the reference holds 1e-8 noise there)
        tol = 3e-4 * max(want_norm / np.sqrt(flat.size), 1e-7) + rel * np.abs(gold['g/' + key + '/samples']).max() + 3e-7
        np.testing.assert_allclose(flat[pos], gold['g/' + key + '/samples'], atol=tol, rtol=0, err_msg=key)
        assert abs(np.sqrt((flat * flat).sum()) - want_norm) <= max(2e-3, rel) * want_norm + 3e-7 * np.sqrt(flat.size), key
        n += 1
    assert n == sum(1 for k in gold if k.endswith('/norm'))
    for key, b in out['buffers'].items():
        np.testing.assert_allclose(b.numpy(), gold['b/' + key], atol=2e-5, rtol=1e-5, err_msg=key)
