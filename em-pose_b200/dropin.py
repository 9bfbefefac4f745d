"""Swap the B200 classes into an installed reference ``empose`` package.

    import empose_b200.dropin; empose_b200.dropin.install()

After that ``scripts/evaluate_real.py`` and friends build ``empose_b200`` models through the reference's own factory
names (``empose.nn.models.create_model`` / ``IterativeErrorFeedback``, ``empose.bodymodels.smpl.SMPLLayer`` /
``create_default_smpl_model``).  Nothing else of the reference is touched.
"""
import importlib


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

    ref_models.IterativeErrorFeedback = b200_models.IterativeErrorFeedback
    ref_models.SimpleRNN = b200_models.SimpleRNN
    ref_models.create_model = create_model
    ref_smpl.SMPLLayer = b200_smpl.SMPLLayer
    ref_smpl.create_default_smpl_model = b200_smpl.create_default_smpl_model
    return ref_models, ref_smpl
