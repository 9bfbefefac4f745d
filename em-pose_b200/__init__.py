"""empose-b200: a B200-native implementation of EM-POSE's learned-gradient-descent hot path.

Public surface (mirrors the reference's, see DESIGN.md / INTEGRATION.md):

* ``empose_b200.nn.models.create_model / IterativeErrorFeedback`` -- reference ``empose/nn/models.py:23-33, 369-688``
* ``empose_b200.bodymodels.smpl.SMPLLayer / create_default_smpl_model`` -- reference ``empose/bodymodels/smpl.py:24-165``
* ``empose_b200.dropin.install()`` -- swap the two classes into an installed reference ``empose`` package
* ``empose_b200.lib`` -- ctypes binding of the C-ABI in ``include/empose_b200.h``

Sub-modules are imported lazily so that ``import empose_b200`` stays cheap.
"""
__version__ = '0.1.0'
