"""Swap the B200 classes into an installed reference ``empose`` package.

    import empose_b200.dropin; empose_b200.dropin.install()

After that ``scripts/evaluate_real.py`` and friends build ``empose_b200`` models through the reference's own factory
names (``empose.nn.models.create_model`` / ``IterativeErrorFeedback``, ``empose.bodymodels.smpl.SMPLLayer`` /
``create_default_smpl_model``).  Nothing else of the reference is touched.

The reference binds these names with ``from ... import`` (``empose/eval/helpers.py:20-25``, the scripts), so a module that
was imported BEFORE ``install()`` holds its own references to the old objects: ``install()`` therefore also rebinds them
in every already-loaded ``empose.*`` module and in ``__main__`` (a script that imported the factories at its top).
"""
import importlib
import sys


def install():
    from empose_b200.bodymodels import smpl as b200_smpl
    from empose_b200.nn import models as b200_models
    ref_models = importlib.import_module('empose.nn.models')
    ref_smpl = importlib.import_module('empose.bodymodels.smpl')
    ref_create = ref_models.create_model

    def create_model(config, *args):
        if config.m_type in ('ief', 'lgd'):
            return b200_models.IterativeErrorFeedback(config, *args)
        if config.m_type == 'rnn' and not getattr(config, 'm_learn_init_state', False):
            return b200_models.SimpleRNN(config, *args)
        return ref_create(config, *args)          # ResNet (and learned initial states) stay on the reference implementation

    swaps = {ref_models.IterativeErrorFeedback: b200_models.IterativeErrorFeedback,
             ref_models.SimpleRNN: b200_models.SimpleRNN,
             ref_models.create_model: create_model,
             ref_smpl.SMPLLayer: b200_smpl.SMPLLayer,
             ref_smpl.create_default_smpl_model: b200_smpl.create_default_smpl_model}
    for name, module in list(sys.modules.items()):
        if module is None or not (name == '__main__' or name == 'empose' or name.startswith('empose.')):
            continue
        for attr, value in list(vars(module).items()):
            try:
                new = swaps.get(value)
            except TypeError:                      # unhashable module attribute
                continue
            if new is not None:
                setattr(module, attr, new)
    return ref_models, ref_smpl
