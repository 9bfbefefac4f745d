"""Checkpoint compatibility with the reference's experiment folders.

The reference stores a trained model as ``$EM_EXPERIMENTS/<id>-<summary>/{config.json, model.pth, cmd.txt, logs}``
(``scripts/train.py:110-121, 195-205``) and loads it back with ``get_model_config`` / ``load_model_weights`` / ``load_model``
(``empose/eval/helpers.py:131-164``).  The functions below are their mirrors for the B200 classes: because the module
mirrors keep the reference's state-dict keys -- including the ``smpl.bm.*`` buffers of the third-party body model -- a
released ``model.pth`` loads with ``strict=True``; BatchNorm folding, fp16 / tf32 rounding, K padding and the LSTM gate
permutation happen when the native context is built from the loaded module (``csrc/model.cu: empose_ief_create``), so no
converted file format exists or is needed.
"""
import glob
import json
import os

import torch

from empose_b200.helpers.configuration import Configuration


def get_model_dir(experiment_dir, model_id):
    """``helpers/utils.py:36-39``: the directory whose name starts with ``<model_id>-``."""
    found = glob.glob(os.path.join(experiment_dir, str(model_id) + "-*"), recursive=False)
    return None if len(found) == 0 else found[0]


def get_model_config(model_id, experiment_dir=None):
    """``eval/helpers.py:140-145``."""
    experiment_dir = experiment_dir or os.environ['EM_EXPERIMENTS']
    model_dir = get_model_dir(experiment_dir, model_id)
    if model_dir is None:
        raise ValueError("Cannot find model directory for experiment ID {}".format(model_id))
    return Configuration.from_json(os.path.join(model_dir, 'config.json')), model_dir


def _load_checkpoint(checkpoint_file, allow_pickle):
    """
    ``torch.load`` restricted to tensors and plain containers (``weights_only=True``): a ``model.pth`` is a downloaded
    file, and the unrestricted loader executes whatever pickle code it contains.  The reference's checkpoint dict holds
    tensors, numbers and the optimiser's state dict, all of which the restricted loader accepts; a checkpoint that needs
    more is refused unless the caller opts in with ``allow_pickle=True``.
    """
    try:
        return torch.load(checkpoint_file, map_location='cpu', weights_only=True)
    except Exception as err:                         # pickle.UnpicklingError and friends
        if not allow_pickle:
            raise ValueError('{} needs the unrestricted pickle loader ({}); pass allow_pickle=True only for files you trust'
                             .format(checkpoint_file, err))
        return torch.load(checkpoint_file, map_location='cpu', weights_only=False)


def load_model_weights(checkpoint_file, net, state_key='model_state_dict', strict=True, allow_pickle=False):
    """``eval/helpers.py:131-137``.  Returns the rest of the checkpoint dict (epoch, losses, optimiser state ...)."""
    if not os.path.exists(checkpoint_file):
        raise ValueError("Could not find model checkpoint {}.".format(checkpoint_file))
    checkpoint = _load_checkpoint(checkpoint_file, allow_pickle)
    net.load_state_dict(checkpoint[state_key], strict=strict)
    return {k: v for k, v in checkpoint.items() if k != state_key}


def load_model(model_id, smpl_model, experiment_dir=None, device=None, precision=None, allow_pickle=False):
    """
    ``eval/helpers.py:148-164`` without the data pipeline: config.json -> ``create_model`` -> ``model.pth``.
    :return: (net in eval mode, config, model_dir, the rest of the checkpoint dict)
    """
    from empose_b200.nn.models import create_model
    config, model_dir = get_model_config(model_id, experiment_dir)
    net = create_model(config, smpl_model)
    if precision is not None:
        net.precision = precision
    extra = load_model_weights(os.path.join(model_dir, 'model.pth'), net, allow_pickle=allow_pickle)
    if device is not None:
        net = net.to(device)
    return net.eval(), config, model_dir, extra


def save_checkpoint(checkpoint_file, net, optimizer=None, **fields):
    """The dict ``scripts/train.py:195-205`` writes (``model_state_dict``, ``optimizer_state_dict`` + bookkeeping fields)."""
    payload = dict(fields)
    payload['model_state_dict'] = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    if optimizer is not None:
        payload['optimizer_state_dict'] = optimizer.state_dict()
    torch.save(payload, checkpoint_file)


def describe_checkpoint(checkpoint_file, state_key='model_state_dict', allow_pickle=False):
    """Summary of a ``model.pth``: tensors per top-level module, trainable-parameter count as the reference prints it."""
    checkpoint = _load_checkpoint(checkpoint_file, allow_pickle)
    sd = checkpoint[state_key]
    groups = {}
    for k, v in sd.items():
        g = groups.setdefault(k.split('.')[0], [0, 0])
        g[0] += 1
        g[1] += v.numel()
    learned = sum(v.numel() for k, v in sd.items() if v.is_floating_point() and 'running_' not in k and
                  not any(k.startswith('smpl.bm.' + b) for b in ('v_template', 'shapedirs', 'J_regressor', 'posedirs', 'weights')))
    return {'groups': {k: {'tensors': n, 'elements': e} for k, (n, e) in groups.items()}, 'trainable_parameters': learned,
            'other_fields': sorted(k for k in checkpoint if k != state_key)}
