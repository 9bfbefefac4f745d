"""Stand-ins that let the UNMODIFIED reference modules import and run here (test infrastructure).

``/root/reference`` imports three packages that are absent from this image (``quaternion`` gets a scipy-backed
stand-in for the four functions the metrics use): ``trimesh``
(``virtual_sensors.py:11``, ``smpl.py:12``), ``quaternion`` (``helpers/utils.py:12``) and
``human_body_prior`` (``smpl.py:20-21``).  ``install()`` registers minimal substitutes in
``sys.modules`` -- the third-party SMPL arithmetic comes from ``oracle.smplh_lbs`` -- sets the four
environment variables ``configuration.py:25-28`` reads at import time, writes a synthetic SMPL-H
model and puts the reference on ``sys.path``.  After that ``empose.nn.models`` etc. are the
reference's own code, which is what ``tests/golden/make_golden.py`` and ``bench.py --impl reference``
run.  Nothing is copied from or written to ``/root/reference``.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from oracle import sensors
from oracle import smplh_lbs
from empose_b200 import synthetic as smplh_synth

REFERENCE_ROOT = os.environ.get('EMPOSE_REFERENCE_ROOT', '/root/reference')


class _BodyModel(nn.Module):
    """Substitute for ``human_body_prior.body_model.body_model.BodyModel`` (buffer / parameter names and
    shapes as listed in SURVEY.md section 8b, which the released checkpoints' ``smpl.bm.*`` keys imply)."""

    def __init__(self, bm_path, num_betas=10, dtype=torch.float32, **_unused):
        super(_BodyModel, self).__init__()
        m = smplh_lbs.SmplhModel(bm_path, num_betas=num_betas, dtype=dtype)
        self._parents = m.parents
        self.register_buffer('v_template', m.v_template.unsqueeze(0))
        self.register_buffer('f', m.faces.to(torch.int32))
        self.register_buffer('shapedirs', m.shapedirs)
        self.register_buffer('J_regressor', m.j_regressor)
        self.register_buffer('posedirs', m.posedirs)
        self.register_buffer('kintree_table', m.kintree_table.to(torch.int32))
        self.register_buffer('weights', m.weights)
        for name, width in (('trans', 3), ('root_orient', 3), ('pose_body', 63), ('pose_hand', 90),
                            ('betas', num_betas)):
            self.register_parameter(name, nn.Parameter(torch.zeros(1, width, dtype=dtype)))

    def forward(self, root_orient, pose_body, betas, pose_hand, trans):
        view = object.__new__(smplh_lbs.SmplhModel)
        view.__dict__.update(dict(v_template=self.v_template[0], shapedirs=self.shapedirs, posedirs=self.posedirs,
                                  j_regressor=self.J_regressor, weights=self.weights, parents=self._parents,
                                  num_betas=self.shapedirs.shape[-1]))
        verts, joints = smplh_lbs.lbs(view, torch.cat([root_orient, pose_body, pose_hand], dim=1), betas, trans)
        return types.SimpleNamespace(v=verts, Jtr=joints)


class _Trimesh(object):
    def __init__(self, vertices, faces, process=False):
        self.vertex_faces = sensors.vertex_faces_table(np.asarray(faces), len(vertices))


def _quaternion_module():
    """Stand-in for the absent ``numpy-quaternion`` package: the four functions ``empose/eval/metrics.py:148-161`` and
    ``empose/data/transforms.py:106-116`` call, on top of ``scipy.spatial.transform.Rotation`` (unit quaternions)."""
    from scipy.spatial.transform import Rotation
    mod = types.ModuleType('quaternion')

    def from_rotation_vector(rv):
        rv = np.asarray(rv, dtype=np.float64)
        return Rotation.from_rotvec(rv.reshape(-1, 3)), rv.shape[:-1]

    def from_rotation_matrix(rm):
        rm = np.asarray(rm, dtype=np.float64)
        return Rotation.from_matrix(rm.reshape(-1, 3, 3)), rm.shape[:-2]

    def as_rotation_matrix(q):
        rot, shape = q
        return rot.as_matrix().reshape(shape + (3, 3))

    def rotation_intrinsic_distance(q1, q2):
        """2 |log(q1^-1 q2)|: the geodesic angle between the two rotations, in [0, pi]."""
        (r1, shape), (r2, _) = q1, q2
        return (r1.inv() * r2).magnitude().reshape(shape)

    mod.from_rotation_vector = from_rotation_vector
    mod.from_rotation_matrix = from_rotation_matrix
    mod.as_rotation_matrix = as_rotation_matrix
    mod.rotation_intrinsic_distance = rotation_intrinsic_distance
    return mod


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'empose'))


def install(asset_dir, seed=0):
    """Make ``import empose`` resolve to the unmodified reference.  Returns the synthetic model path."""
    if not reference_available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    os.makedirs(asset_dir, exist_ok=True)
    for var in ('EM_DATA_SYNTH', 'EM_EXPERIMENTS', 'SMPL_MODELS', 'EM_DATA_REAL'):
        os.environ.setdefault(var, asset_dir)
    model_path = smplh_synth.write_synthetic_smplh(os.environ['SMPL_MODELS'], seed=seed)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.modules.setdefault('quaternion', _quaternion_module())
    tm = types.ModuleType('trimesh')
    tm.Trimesh = _Trimesh
    sys.modules.setdefault('trimesh', tm)
    for name in ('human_body_prior', 'human_body_prior.body_model', 'human_body_prior.body_model.body_model',
                 'human_body_prior.tools', 'human_body_prior.tools.model_loader'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['human_body_prior.body_model.body_model'].BodyModel = _BodyModel
    sys.modules['human_body_prior.tools.model_loader'].load_vposer = lambda path: (None, None)
    return model_path


def make_config(flags):
    """Build a reference ``Configuration`` by driving its own argparse (``configuration.py:146-212``)."""
    from empose.helpers.configuration import Configuration
    saved = sys.argv
    sys.argv = ['oracle'] + list(flags)
    try:
        return Configuration.parse_cmd()
    finally:
        sys.argv = saved
