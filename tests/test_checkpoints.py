"""experiment folders written the reference's way load into the B200 classes with strict=True."""
import json
import os

import numpy as np
import pytest
import torch

from empose_b200 import checkpoints, synthetic
from oracle import ref_shims

import util


def _write_experiment(tmp_path, model_id, state_dict, config_dict):
    model_dir = os.path.join(str(tmp_path), '%d-IEF-2x512-N2-RNN-2x512-r0.01-ws32-lr0.0005-grad-n6-pos-ori' % model_id)
    os.makedirs(model_dir)
    with open(os.path.join(model_dir, 'config.json'), 'w') as f:
        json.dump(config_dict, f, indent=4, sort_keys=True)
    torch.save({'iteration': 7, 'epoch': 3, 'global_step': 1234, 'model_state_dict': state_dict, 'train_loss': torch.tensor([1.5]),
                'valid_loss': 1.25, 'test_eucl_mean': 31.8, 'test_angle_mean': 12.1}, os.path.join(model_dir, 'model.pth'))
    return model_dir


def test_reference_experiment_folder_loads_strict(smpl_npz, asset_dir, tmp_path):
    """config.json + model.pth as scripts/train.py:110-121,195-205 write them -- produced by the reference's own classes
    when its tree is present, by the mirror otherwise -- load into the B200 module with identical tensors."""
    from empose_b200.bodymodels.smpl import SMPLLayer
    from empose_b200.nn.models import IterativeErrorFeedback
    flags = ['--m_type', 'ief', '--m_num_iterations', '2', '--m_hidden_size', '512', '--m_rnn_init', '--m_average_shape',
             '--m_use_gradient', '--use_marker_pos', '--use_marker_ori', '--n_markers', '6', '--window_size', '32',
             '--m_fk_loss', '0.1', '--m_pose_loss_weight', '10.0', '--lr', '0.0005']
    if ref_shims.reference_available():
        ref_shims.install(asset_dir, seed=0)
        from empose.bodymodels.smpl import create_default_smpl_model
        from empose.nn.models import create_model
        cfg = ref_shims.make_config(flags)
        ref_net = create_model(cfg, create_default_smpl_model(device='cpu'))
        sd = ref_net.state_dict()
        config_dict = {k: v for k, v in vars(cfg).items()}
    else:
        net0 = util.build_module(smpl_npz, n_markers=6, num_iterations=2, m_fk_loss=0.1, m_pose_loss_weight=10.0, lr=0.0005)
        sd = net0.state_dict()
        config_dict = dict(vars(net0.config))
    for k, v in synthetic.synth_state_dict(seed=3, n_markers=6, rnn_init=True).items():
        sd[k] = torch.from_numpy(np.asarray(v))
    _write_experiment(tmp_path, 1615631737, sd, config_dict)

    net, config, model_dir, extra = checkpoints.load_model(1615631737, SMPLLayer(smpl_npz).to(dtype=torch.float32),
                                                           experiment_dir=str(tmp_path))
    assert isinstance(net, IterativeErrorFeedback) and not net.training
    assert config.m_num_iterations == 2 and config.n_markers == 6 and extra['global_step'] == 1234
    got = net.state_dict()
    assert set(got) == set(sd)
    for k in sd:
        assert torch.equal(got[k].cpu(), sd[k].to(got[k].dtype)), k
    info = checkpoints.describe_checkpoint(os.path.join(model_dir, 'model.pth'))
    assert info['trainable_parameters'] == 5721419                    # README.md:228 of the reference
    # round trip through the mirror's own writer
    out = os.path.join(str(tmp_path), 'again.pth')
    checkpoints.save_checkpoint(out, net, epoch=4)
    rest = checkpoints.load_model_weights(out, net)
    assert rest['epoch'] == 4
    with pytest.raises(ValueError):
        checkpoints.get_model_config(42, experiment_dir=str(tmp_path))


def test_checkpoint_loader_refuses_pickled_code(tmp_path, smpl_npz):
    """A model.pth is a downloaded file: the loader takes tensors and plain containers only, unless the caller opts in."""
    import os
    import pickle
    import util
    from empose_b200 import checkpoints
    net = util.build_module(smpl_npz)
    good = os.path.join(str(tmp_path), 'model.pth')
    checkpoints.save_checkpoint(good, net, epoch=3, train_loss=torch.tensor(0.5), valid_loss=0.25)
    extra = checkpoints.load_model_weights(good, net)
    assert extra['epoch'] == 3 and float(extra['train_loss']) == 0.5

    class Evil(object):
        def __reduce__(self):
            return (os.system, ('true',))
    bad = os.path.join(str(tmp_path), 'evil.pth')
    torch.save({'model_state_dict': net.state_dict(), 'payload': Evil()}, bad, pickle_module=pickle)
    with pytest.raises(ValueError, match='unrestricted pickle'):
        checkpoints.load_model_weights(bad, net)
