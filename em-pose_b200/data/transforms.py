"""Training-data synthesis on the B200: the two transforms the reference applies right before the hot path.

Interface of ``SMPLFK`` (``empose/data/transforms.py:259-282``) and ``SampleMarkersWithOffsets`` (``:163-226``): same
constructor arguments, same attributes set on the batch, and the same random draws in the same order
(``np.random.RandomState(6273)`` for the subject whose offsets a window gets, torch's global generator for the offset
noise), so a run is reproducible against the reference.  What changes is the arithmetic: the reference
evaluates the full 6890-vertex SMPL-H mesh for every frame (``batch.vertices``, 83 KB per frame) and derives the sensor
frames from it with torch ops; here ONE pass of the sub-model kernels (``empose_sensor_project`` on a context made by
``empose_sensors_create``) goes from poses / shapes / offsets to sensor positions, orientations and joints and the mesh
is never materialised.  ``batch.vertices`` is therefore ``None`` unless ``SMPLFK(..., keep_vertices=True)``, and
``marker_normal_vertex`` holds the UNIT sensor normal (the reference leaves the un-normalised area-weighted vertex normal
there, ``virtual_sensors.py:92-96``; nothing consumes it: normals as inputs are rejected at ``models.py:122-123``).
"""
import numpy as np
import torch
from torch.distributions import MultivariateNormal

from empose_b200 import lib as _lib
from empose_b200.helpers.configuration import CONSTANTS as C


def _sensor_context(smpl_model, device):
    """One sub-model-only native context per (SMPL layer, device)."""
    cache = smpl_model.__dict__.setdefault('_sensor_contexts', {})
    index = device.index if device.index is not None else torch.cuda.current_device()
    if index not in cache:
        cache[index] = _lib.SensorContext(smpl_model.submodel_arrays(), getattr(smpl_model, 'precision', 0), index)
    return cache[index]


def _device_of(batch):
    dev = batch.poses_body.device
    if dev.type != 'cuda':
        raise _lib.EmposeError('empose_b200 data synthesis runs on CUDA tensors only (no CPU path); move the batch to the GPU first')
    return dev


class SMPLFK(object):
    """Sets ``joints_gt`` (22 joints, ``trans`` applied), ``joints_hat`` and ``vertices`` on the batch (transforms.py:259-282)."""

    def __init__(self, smpl_model, keep_vertices=False):
        self.smpl_model = smpl_model
        self.max_window_size = 1000
        self.keep_vertices = keep_vertices

    def __call__(self, batch):
        dev = _device_of(batch)
        n, f = batch.batch_size, batch.seq_length
        r = n * f
        pose = torch.cat([batch.poses_root, batch.poses_body], dim=-1).reshape(r, 66)
        shape = batch.shapes.unsqueeze(1).repeat(1, f, 1).reshape(r, -1)
        trans = batch.trans.reshape(r, 3)
        if self.keep_vertices:
            vertices, joints = self.smpl_model(poses_body=pose[:, 3:], betas=shape, poses_root=pose[:, :3], trans=trans,
                                               window_size=self.max_window_size)
            batch.vertices = vertices.reshape(n, f, -1)
            joints = joints[:, :(1 + C.N_JOINTS)]
        else:
            eye = torch.eye(3, device=dev).reshape(1, 1, 3, 3).expand(r, 12, 3, 3).contiguous()
            zero = torch.zeros(r, 12, 3, device=dev)
            _, _, joints = _sensor_context(self.smpl_model, dev).sensor_project(pose, shape, eye, zero)
            joints = joints + trans.unsqueeze(1)
            batch.vertices = None
        batch.joints_gt = joints.reshape(n, f, -1)
        batch.joints_hat = batch.joints_gt.clone().detach()
        return batch


class _OffsetBank(object):
    """The pre-estimated sensor-to-skin offsets of a set of subjects (one ``.npz`` per subject with ``means`` (M,3), ``covs``
    (M,3,3), ``r`` (M,3,3) and ``vertex_ids``; written by the reference's calibration, read at ``transforms.py:153-160``)."""

    def __init__(self, files):
        records = [np.load(path) for path in (files if isinstance(files, (list, tuple)) else [files])]
        stack = lambda key: torch.from_numpy(np.stack([np.asarray(rec[key]) for rec in records])).to(dtype=torch.float32)
        self.means, self.covs, self.rotations = stack('means'), stack('covs'), stack('r')       # (S,M,3), (S,M,3,3), (S,M,3,3)
        self.vertex_ids = [int(v) for v in records[-1]['vertex_ids']]
        self.n_sets, self.n_markers = int(self.means.shape[0]), int(self.means.shape[1])


class SampleMarkersWithOffsets(object):
    """Virtual sensors with pre-estimated sensor-to-skin offsets (the transform of ``transforms.py:139-226``), computed from
    poses instead of from a materialised mesh.

    Random streams are the reference's, so a run is reproducible against it draw for draw: which subject's offsets a
    window gets comes from ``numpy.random.RandomState(6273).randint`` (one call per batch), and the offset noise from ONE
    ``MultivariateNormal(means, covs).sample`` call on torch's global generator -- ``(N,)`` draws for ``noise_level`` 0,
    ``(N, F)`` for 1, none otherwise.  ``noise_level``: -1 mean offsets; 0 one noisy offset per window; 1 one per frame;
    2 zero translation; 3 zero translation and identity rotation."""

    def __init__(self, smpl_model, offset_files, noise_level=-1):
        if noise_level not in (-1, 0, 1, 2, 3):
            raise ValueError("Unknown noise level {}".format(noise_level))
        self.smpl_model = smpl_model
        self.noise_level = noise_level
        self.randomize = noise_level >= 0
        self.bank = _OffsetBank(offset_files)
        if list(self.bank.vertex_ids) != list(C.VERTEX_IDS):
            raise ValueError('the offset files must use the 12 sensor vertices of configuration.py:32-34 '
                             '(the SMPL sub-model is extracted for exactly those)')
        self.n_offsets, self.n_markers, self.vertex_ids = self.bank.n_sets, self.bank.n_markers, self.bank.vertex_ids
        self.normal_dists = MultivariateNormal(loc=self.bank.means, covariance_matrix=self.bank.covs)
        self.offset_rng = np.random.RandomState(6273)

    def _draw(self, n, f):
        """Per-window subject index, and translation (N, F, M, 3) / rotation (N, M, 3, 3) offsets on the host."""
        subject = torch.from_numpy(self.offset_rng.randint(0, self.n_offsets, n)).long()
        mean_t = self.bank.means[subject]                                           # (N, M, 3)
        window = torch.arange(n)
        if self.noise_level == 0:
            t = self.normal_dists.sample((n,))[window, subject].unsqueeze(1).expand(n, f, self.n_markers, 3)
        elif self.noise_level == 1:
            t = self.normal_dists.sample((n, f))[window, :, subject]               # advanced indices first: (N, F, M, 3)
        elif self.noise_level in (2, 3):
            t = torch.zeros(n, f, self.n_markers, 3)
        else:
            t = mean_t.unsqueeze(1).expand(n, f, self.n_markers, 3)
        if self.noise_level == 3:
            rot = torch.eye(3).expand(n, self.n_markers, 3, 3)
        else:
            rot = self.bank.rotations[subject]
        return mean_t, t, rot

    def __call__(self, batch):
        dev = _device_of(batch)
        n, f = batch.batch_size, batch.seq_length
        rows = n * f
        mean_t, t_off, r_off = self._draw(n, f)
        t_off = t_off.to(dev).reshape(rows, self.n_markers, 3)
        r_win = r_off.to(dev).contiguous()
        r_rows = r_win.unsqueeze(1).expand(n, f, self.n_markers, 3, 3).reshape(rows, self.n_markers, 3, 3)
        pose = torch.cat([batch.poses_root, batch.poses_body], dim=-1).reshape(rows, 66)
        shape = batch.shapes.unsqueeze(1).expand(n, f, batch.shapes.shape[-1]).reshape(rows, -1)
        trans = batch.trans.reshape(rows, 1, 3)
        # two passes of the sub-model kernels: the bare sensor frames, then the frames with the offsets applied
        # (R' = R R_off, p' = p + R t_off, models.py:478-479 -- the same arithmetic the hot path uses)
        ctx = _sensor_context(self.smpl_model, dev)
        identity = torch.eye(3, device=dev).expand(rows, self.n_markers, 3, 3).contiguous()
        bare_pos, bare_ori, _ = ctx.sensor_project(pose, shape, identity, torch.zeros(rows, self.n_markers, 3, device=dev))
        off_pos, off_ori, _ = ctx.sensor_project(pose, shape, r_rows, t_off)
        as_batch = lambda x: x.reshape(n, f, -1)
        batch.marker_pos_vertex, batch.marker_ori_vertex = as_batch(bare_pos + trans), as_batch(bare_ori)
        batch.marker_normal_vertex = as_batch(bare_ori[..., 2])
        batch.marker_pos_synth, batch.marker_ori_synth = as_batch(off_pos + trans), as_batch(off_ori)
        batch.marker_normal_synth = as_batch(off_ori[..., 2])
        # what a model may use to undo the offsets at test time: always the MEAN translation, and the rotation
        batch.offset_t_augmented = mean_t.to(dev)
        batch.offset_r_augmented = r_win.clone()
        return batch
