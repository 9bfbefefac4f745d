import os
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def asset_dir():
    d = os.path.join(tempfile.gettempdir(), 'empose_b200_assets')
    os.makedirs(d, exist_ok=True)
    return d


@pytest.fixture(scope='session')
def smpl_npz(asset_dir):
    from empose_b200 import synthetic
    return synthetic.write_synthetic_smplh(asset_dir, seed=0)


@pytest.fixture(scope='session')
def oracle_smpl(smpl_npz):
    import torch
    from oracle import smplh_lbs
    return smplh_lbs.SmplhModel(smpl_npz, num_betas=10, dtype=torch.float64)


@pytest.fixture(scope='session')
def topology(oracle_smpl):
    from oracle import sensors
    return sensors.sensor_topology(oracle_smpl.faces.numpy())
