"""Training-data synthesis on the B200: the two transforms the reference applies right before the hot path.

Mirrors ``SMPLFK`` (``empose/data/transforms.py:259-282``) and ``SampleMarkersWithOffsets`` (``:163-226``): same
constructor arguments, same attributes set on the batch, same host-side random streams (``np.random.RandomState(6273)``
for the offset set, torch's global generator for the offset noise).  What changes is the arithmetic: the reference
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


class SampleMarkersWithOffsets(object):
    """Virtual sensors with pre-estimated sensor-to-skin offsets (transforms.py:139-226), from poses instead of vertices."""

    def __init__(self, smpl_model, offset_files, noise_level=-1):
        self.smpl_model = smpl_model
        self.randomize = noise_level >= 0
        self.noise_level = noise_level
        if not isinstance(offset_files, list):
            offset_files = [offset_files]
        self.n_markers = np.load(offset_files[0])['means'].shape[0]
        self.n_offsets = len(offset_files)
        self.offset_means = np.zeros([self.n_offsets, self.n_markers, 3])
        self.offset_covs = np.zeros([self.n_offsets, self.n_markers, 3, 3])
        self.r = np.zeros([self.n_offsets, self.n_markers, 3, 3])
        offset_data = None
        for i, offset_file in enumerate(offset_files):
            offset_data = np.load(offset_file)
            self.offset_means[i] = offset_data['means']
            self.offset_covs[i] = offset_data['covs']
            self.r[i] = offset_data['r']
        self.normal_dists = MultivariateNormal(loc=torch.from_numpy(self.offset_means).to(dtype=torch.float32),
                                               covariance_matrix=torch.from_numpy(self.offset_covs).to(dtype=torch.float32))
        self.vertex_ids = offset_data['vertex_ids'].tolist()
        if list(self.vertex_ids) != list(C.VERTEX_IDS):
            raise ValueError('the offset files must use the 12 sensor vertices of configuration.py:32-34 '
                             '(the SMPL sub-model is extracted for exactly those)')
        self.offset_rng = np.random.RandomState(6273)

    def __call__(self, batch):
        dev = _device_of(batch)
        n, f = batch.batch_size, batch.seq_length
        r_rows = n * f
        # ---- host side: which offset set, which noise (identical calls and streams as the reference) ----
        s_idxs = self.offset_rng.randint(0, self.n_offsets, n)
        offset_means = torch.from_numpy(self.offset_means[s_idxs]).to(dtype=torch.float32)
        local_offsets = offset_means.clone().unsqueeze(1).repeat(1, f, 1, 1)
        s_idx_t = torch.from_numpy(s_idxs).to(dtype=torch.long)
        if self.randomize:
            if self.noise_level == 0:
                noise = self.normal_dists.sample((n,))[torch.arange(n), s_idx_t]
                local_offsets = noise.unsqueeze(1).repeat(1, f, 1, 1)
            elif self.noise_level == 1:
                noise = self.normal_dists.sample((n, f))
                s = s_idx_t.unsqueeze(-1).repeat(1, f).reshape(-1)
                local_offsets = noise.reshape((n * f, self.n_offsets, -1, 3))[torch.arange(n * f), s].reshape((n, f, -1, 3))
            elif self.noise_level in (2, 3):
                local_offsets = torch.zeros_like(local_offsets)
            else:
                raise ValueError("Unknown noise level {}".format(self.noise_level))
        rot = torch.from_numpy(self.r).to(dtype=torch.float32)[s_idx_t].unsqueeze(1).repeat(1, f, 1, 1, 1)
        if self.randomize and self.noise_level == 3:
            rot = torch.eye(3).reshape(1, 1, 1, 3, 3).repeat(n, f, self.n_markers, 1, 1)
        local_offsets, rot = local_offsets.to(dev), rot.to(dev)
        # ---- device side: SMPL sub-model -> sensor frames -> offsets, twice (raw frames, then with offsets) ----
        pose = torch.cat([batch.poses_root, batch.poses_body], dim=-1).reshape(r_rows, 66)
        shape = batch.shapes.unsqueeze(1).repeat(1, f, 1).reshape(r_rows, -1)
        trans = batch.trans.reshape(r_rows, 1, 3)
        ctx = _sensor_context(self.smpl_model, dev)
        eye = torch.eye(3, device=dev).reshape(1, 1, 3, 3).expand(r_rows, 12, 3, 3).contiguous()
        zero = torch.zeros(r_rows, 12, 3, device=dev)
        pos0, ori0, _ = ctx.sensor_project(pose, shape, eye, zero)
        pos1, ori1, _ = ctx.sensor_project(pose, shape, rot.reshape(r_rows, 12, 3, 3), local_offsets.reshape(r_rows, 12, 3))
        batch.marker_pos_vertex = (pos0 + trans).reshape(n, f, -1)
        batch.marker_ori_vertex = ori0.reshape(n, f, -1)
        batch.marker_normal_vertex = ori0[..., 2].reshape(n, f, -1)
        batch.marker_pos_synth = (pos1 + trans).reshape(n, f, -1)
        batch.marker_ori_synth = ori1.reshape(n, f, -1)
        batch.marker_normal_synth = ori1[..., 2].reshape(n, f, -1)
        batch.offset_t_augmented = offset_means.clone().detach().to(dev)
        batch.offset_r_augmented = rot[:, 0].clone().detach()
        return batch
