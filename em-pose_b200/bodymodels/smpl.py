"""SMPL-H layer with the reference's call surface (``empose/bodymodels/smpl.py:24-165``).

``SMPLLayer`` keeps the third-party ``BodyModel``'s buffers and parameters under ``.bm`` so that the
``smpl.bm.*`` keys of released checkpoints load unchanged, and exposes ``faces`` / ``vertex_faces`` as
the reference does.  ``forward`` / ``fk`` evaluate the full 6890-vertex mesh in the CUDA library
(``empose_smpl_forward``); the LGD loop itself never needs the full mesh and uses the sub-model of
``empose_b200.submodel``.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from empose_b200 import submodel as _submodel
from empose_b200.helpers.configuration import CONSTANTS as C


class _BodyModelBuffers(nn.Module):
    """Buffer / parameter holder with the names and shapes of ``human_body_prior`` ``BodyModel``."""

    def __init__(self, bm_path, num_betas=10, dtype=torch.float32):
        super(_BodyModelBuffers, self).__init__()
        with np.load(bm_path) as z:
            data = {k: z[k] for k in ('v_template', 'f', 'shapedirs', 'posedirs', 'J_regressor', 'kintree_table', 'weights')}
        t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dtype)
        n_v = data['v_template'].shape[0]
        self.register_buffer('v_template', t(data['v_template']).unsqueeze(0))
        self.register_buffer('f', torch.as_tensor(np.asarray(data['f']).astype(np.int32)))
        self.register_buffer('shapedirs', t(data['shapedirs'][:, :, :num_betas]))
        self.register_buffer('J_regressor', t(data['J_regressor']))
        pd = np.asarray(data['posedirs'], dtype=np.float64)
        self.register_buffer('posedirs', t(pd.reshape(n_v * 3, -1).T))
        self.register_buffer('kintree_table', torch.as_tensor(np.asarray(data['kintree_table']).astype(np.int32)))
        self.register_buffer('weights', t(data['weights']))
        for name, width in (('trans', 3), ('root_orient', 3), ('pose_body', 63), ('pose_hand', 90), ('betas', num_betas)):
            self.register_parameter(name, nn.Parameter(torch.zeros(1, width, dtype=dtype)))


def create_default_smpl_model(device=None, vposer_path=None):
    """``smpl.py:24-28``: loads ``$SMPL_MODELS/smplh_amass/neutral/model.npz`` as float32 on ``device``."""
    layer = SMPLLayer(os.path.join(C.SMPL_MODELS_DIR, 'smplh_amass/neutral/model.npz'))
    return layer.to(device=device if device is not None else C.DEVICE, dtype=torch.float32)


class SMPLLayer(nn.Module):
    def __init__(self, smpl_path, device=None, vposer_path=None):
        super(SMPLLayer, self).__init__()
        if vposer_path is not None:
            raise ValueError('VPoser is not part of the LGD path and is not supported')
        self.num_betas = C.N_SHAPE_PARAMS
        self.bm = _BodyModelBuffers(smpl_path, num_betas=self.num_betas, dtype=torch.float64)
        self.vposer = None
        self._vertex_faces = None
        self._faces = None
        self._topology = None
        self._ctx = None
        self._ctx_key = None
        self.precision = 0            # lib.PRECISION_TF32; set to lib.PRECISION_FP32 for the exact-arithmetic mode

    @property
    def faces(self):
        if self._faces is None:
            self._faces = self.bm.f.to(dtype=torch.int32)
        return self._faces

    def vertex_faces(self, n_vertices):
        """Per-vertex incident faces, -1 padded (``smpl.py:58-67``)."""
        if self._vertex_faces is None:
            table = _submodel.vertex_faces_table(self.faces.cpu().numpy(), n_vertices)
            self._vertex_faces = torch.from_numpy(table).to(dtype=torch.long, device=self.faces.device)
        return self._vertex_faces

    def sensor_topology(self, vertex_ids=None):
        """Sub-mesh faces / per-sensor faces / helper vertices as the reference derives them (``virtual_sensors.py:47-75``)."""
        if self._topology is None:
            ids = tuple(vertex_ids) if vertex_ids is not None else tuple(C.VERTEX_IDS)
            self._topology = _submodel.topology_from_faces(self.faces.cpu().numpy(), ids)
        return self._topology

    def submodel_arrays(self):
        """The ``sub.*`` arrays the C ABI consumes (float64 extraction, see submodel.py)."""
        bm = self.bm
        f64 = lambda t: t.detach().cpu().double().numpy()
        sub = _submodel.extract_submodel(f64(bm.v_template), f64(bm.shapedirs), f64(bm.posedirs), f64(bm.J_regressor),
                                         f64(bm.weights), bm.kintree_table.cpu().numpy(), self.sensor_topology())
        d = sub.pop('dims')
        sub['sub.dims'] = np.asarray([d['n_verts'], d['vp_dim'], d['n_faces'], d['max_degree'], d['n_skin'],
                                      d['n_sensors']], dtype=np.int32)
        sub.pop('sub.global_vertex_ids')
        return sub

    def fullmodel_arrays(self):
        """The ``smpl.*`` arrays of the full mesh for ``empose_smpl_create`` (float64 extraction, see submodel.py)."""
        bm = self.bm
        f64 = lambda t: t.detach().cpu().double().numpy()
        return _submodel.extract_fullmodel(f64(bm.v_template), f64(bm.shapedirs), f64(bm.posedirs), f64(bm.J_regressor),
                                           f64(bm.weights), bm.kintree_table.cpu().numpy())

    def invalidate(self):
        """Drop the packed full-mesh model: the next call re-reads ``self.bm`` (needed after writes through ``.data``)."""
        if self._ctx is not None:
            self._ctx.close()
        self._ctx, self._ctx_key = None, None

    def _native(self, device):
        from empose_b200 import lib as _lib
        if device.type != 'cuda':
            raise _lib.EmposeError('empose_b200 runs on CUDA devices only (no CPU fallback); got %s' % device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        key = (index, self.precision) + tuple((t.data_ptr(), t._version) for t in self.bm.buffers())
        if self._ctx is None or key != self._ctx_key:
            if self._ctx is not None:
                self._ctx.close()
            self._ctx = _lib.SmplContext(self.fullmodel_arrays(), self.precision, index)
            self._ctx_key = key
        return self._ctx

    def _fk(self, poses_body, betas, poses_root=None, trans=None, normalize_root=False):
        """``smpl.py:81-122``: zero hand pose, root / trans default to zero, betas broadcast and cut to 10."""
        assert poses_body.shape[1] >= C.N_JOINTS * 3                        # smpl.py:93
        if normalize_root:
            raise NotImplementedError('normalize_root is a data-normalisation option outside the LGD hot path')
        n = poses_body.shape[0]
        if len(betas.shape) == 1 or betas.shape[0] == 1:
            betas = betas.reshape(1, -1).repeat(n, 1)
        betas = betas[:, :self.num_betas]
        return self._native(poses_body.device).forward(poses_body[:, :C.N_JOINTS * 3], betas, poses_root, trans)

    def fk(self, poses_body, betas, poses_root=None, trans=None, normalize_root=False, window_size=None):
        """``smpl.py:124-147``.  ``window_size`` is accepted for compatibility: the library evaluates in slabs anyway."""
        if window_size is not None and normalize_root:
            raise ValueError("Are you sure you want to use root normalization with windowed evaluation?")
        return self._fk(poses_body, betas, poses_root, trans, normalize_root)

    def forward(self, *args, **kwargs):
        return self.fk(*args, **kwargs)
