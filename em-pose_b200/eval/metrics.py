"""The metrics engine with the reference's interface, computed on the B200.

Mirrors ``MetricsEngine`` (``empose/eval/metrics.py:69-345``): same constructor, ``reset`` / ``compute`` /
``compute_joint_dist`` / ``get_metrics`` / ``to_pretty_string`` / ``to_tensorboard_log``, same joint selections and the
same aggregation.  ``compute`` is where the reference spends its time -- two full-mesh ``smpl.fk`` calls plus a numpy SVD
per frame on the host (``metrics.py:115-125``); here the per-frame work (FK of the 22 body joints for ground truth and
prediction, Euclidean distances, Procrustes alignment, angular distances of the global orientations) is ONE kernel
(``empose_metrics_compute``) and only the (frames x joints) result tables come back to the host for aggregation.
"""
import numpy as np
import torch
from tabulate import tabulate

from empose_b200 import lib as _lib
from empose_b200.helpers.configuration import CONSTANTS as C

SMPL_JOINTS = ['root', 'l_hip', 'r_hip', 'spine1', 'l_knee', 'r_knee', 'spine2', 'l_ankle', 'r_ankle', 'spine3', 'l_foot',
               'r_foot', 'neck', 'l_collar', 'r_collar', 'head', 'l_shoulder', 'r_shoulder', 'l_elbow', 'r_elbow', 'l_wrist',
               'r_wrist']                                                     # configuration.py:115-117


class MetricsEngine(object):
    """Helper class to compute metrics over a dataset (``metrics.py:69``)."""

    def __init__(self, smpl_model):
        self.smpl_model = smpl_model
        self.eucl_dists = []
        self.eucl_dists_pa = []
        self.angle_diffs = []
        self.eucl_eval_joints = ['root', 'l_hip', 'r_hip', 'spine1', 'l_knee', 'r_knee', 'spine2', 'l_ankle', 'r_ankle',
                                 'spine3', 'neck', 'l_collar', 'r_collar', 'head', 'l_shoulder', 'r_shoulder',
                                 'l_elbow', 'r_elbow', 'l_wrist', 'r_wrist']
        self.angle_eval_joints = ['l_hip', 'r_hip', 'spine1', 'l_knee', 'r_knee', 'spine2', 'spine3',
                                  'neck', 'l_collar', 'r_collar', 'head', 'l_shoulder', 'r_shoulder',
                                  'l_elbow', 'r_elbow']
        self.eucl_idxs = [SMPL_JOINTS.index(j) for j in self.eucl_eval_joints]
        self.angle_idxs = [SMPL_JOINTS.index(j) - 1 for j in self.angle_eval_joints]
        self.angle_glob = True

    def reset(self):
        self.eucl_dists = []
        self.eucl_dists_pa = []
        self.angle_diffs = []

    @staticmethod
    def _masked_flatten(t, mask):
        return t.masked_select(mask.unsqueeze(-1)).reshape(-1, t.shape[-1])

    def _pad_shapes(self, s, n, mask):
        if len(s.shape) == 3:
            return self._masked_flatten(s, mask)
        return self._masked_flatten(s.unsqueeze(1).repeat(1, n, 1), mask)

    @staticmethod
    def _get_mask(seq_lengths, n, f, frame_mask, device):
        """``metrics.py:163-181``."""
        if seq_lengths is not None:
            mask = torch.arange(f, device=device).unsqueeze(0) < seq_lengths.to(device).reshape(-1, 1)
        else:
            mask = torch.ones(n, f, dtype=torch.bool, device=device)
        if frame_mask is not None:
            frame_mask = frame_mask.to(dtype=torch.bool, device=device)
            if len(frame_mask.shape) == 3:
                frame_mask = frame_mask.logical_not().any(dim=-1).logical_not()
            else:
                assert len(frame_mask.shape) == 2
            mask = torch.logical_and(mask, frame_mask)
        return mask

    def _context(self, device):
        if device.type != 'cuda':
            raise _lib.EmposeError('empose_b200 metrics run on CUDA tensors only (no CPU path)')
        cache = self.smpl_model.__dict__.setdefault('_sensor_contexts', {})
        index = device.index if device.index is not None else torch.cuda.current_device()
        if index not in cache:
            cache[index] = _lib.SensorContext(self.smpl_model.submodel_arrays(), getattr(self.smpl_model, 'precision', 0), index)
        return cache[index]

    def compute(self, pose, shape, pose_hat, shape_hat=None, seq_lengths=None, pose_root=None, pose_root_hat=None,
                frame_mask=None):
        """``metrics.py:183-241``: same arguments; the results are appended to ``eucl_dists`` / ``eucl_dists_pa`` /
        ``angle_diffs`` as (frames, joints) numpy arrays like the reference does."""
        n, f = pose.shape[0], pose.shape[1]
        if shape_hat is None:
            shape_hat = shape
        mask = self._get_mask(seq_lengths, n, f, frame_mask, pose.device)
        if mask.sum() == 0:
            return
        shape = self._pad_shapes(shape, f, mask)
        shape_hat = self._pad_shapes(shape_hat, f, mask)
        pose = self._masked_flatten(pose, mask)
        pose_hat = self._masked_flatten(pose_hat, mask)
        if pose_root is None:
            pose_root = torch.zeros([pose.shape[0], 3], dtype=pose.dtype, device=pose.device)
            pose_root_hat = torch.zeros([pose.shape[0], 3], dtype=pose.dtype, device=pose.device)
        else:
            pose_root = self._masked_flatten(pose_root, mask)
            pose_root_hat = self._masked_flatten(pose_root_hat, mask)
        eucl, eucl_pa, angle = self._context(pose.device).metrics(torch.cat([pose_root, pose], dim=-1), shape,
                                                                  torch.cat([pose_root_hat, pose_hat], dim=-1), shape_hat)
        self.eucl_dists.append(eucl.cpu().numpy())
        self.eucl_dists_pa.append(eucl_pa.cpu().numpy())
        if not self.angle_glob:
            raise NotImplementedError('local (non-global) angular distances are not built; the reference evaluates with angle_glob=True')
        self.angle_diffs.append(angle.cpu().numpy().astype(np.float64))

    def compute_joint_dist(self, joints, joints_hat, seq_lengths=None, frame_mask=None):
        """``metrics.py:243-265``: only the metrics on given 3D joints (N, F, 66)."""
        n, f = joints.shape[0], joints.shape[1]
        mask = self._get_mask(seq_lengths, n, f, frame_mask, joints.device)
        if mask.sum() == 0:
            return
        js = self._masked_flatten(joints, mask)[:, :(C.N_JOINTS + 1) * 3]
        js_hat = self._masked_flatten(joints_hat, mask)[:, :(C.N_JOINTS + 1) * 3]
        eucl, eucl_pa = _lib.metrics_from_joints(js, js_hat)
        self.eucl_dists.append(eucl.cpu().numpy())
        self.eucl_dists_pa.append(eucl_pa.cpu().numpy())

    def get_metrics(self, eucl_idxs_select=True, angle_idxs_select=True):
        """``metrics.py:287-330`` verbatim in behaviour."""
        if len(self.eucl_dists) > 0:
            eucl_dists = np.concatenate(self.eucl_dists, axis=0)
            eucl_dists_pa = np.concatenate(self.eucl_dists_pa, axis=0)
            eucl_idxs = self.eucl_idxs if eucl_idxs_select else list(range(eucl_dists.shape[1]))
            eucl_mean_all = np.mean(np.mean(eucl_dists, axis=0)[eucl_idxs])
            eucl_std_all = np.std(eucl_dists[:, eucl_idxs])
            eucl_mean_pa_all = np.mean(np.mean(eucl_dists_pa, axis=0)[eucl_idxs])
            eucl_std_pa_all = np.std(eucl_dists_pa[:, eucl_idxs])
        else:
            eucl_mean_all = eucl_std_all = eucl_mean_pa_all = eucl_std_pa_all = 0.0
        if len(self.angle_diffs) > 0:
            angle_diffs = np.concatenate(self.angle_diffs, axis=0)
            angle_idxs = self.angle_idxs if angle_idxs_select else list(range(angle_diffs.shape[1]))
            angle_mean_all = np.mean(np.mean(angle_diffs, axis=0)[angle_idxs])
            angle_std_all = np.std(angle_diffs[:, angle_idxs])
        else:
            angle_mean_all = angle_std_all = 0.0
        return {'MPJPE [mm]': eucl_mean_all * 1000.0, 'MPJPE STD': eucl_std_all * 1000.0,
                'PA-MPJPE [mm]': eucl_mean_pa_all * 1000.0, 'PA-MPJPE STD': eucl_std_pa_all * 1000.0,
                'MPJAE [deg]': angle_mean_all, 'MPJAE STD': angle_std_all}

    @staticmethod
    def to_pretty_string(metrics, model_name):
        headers, values = [], []
        for k in metrics:
            headers.append(k)
            values.append(metrics[k])
        return tabulate([[model_name] + values], headers=['Model'] + headers)

    @staticmethod
    def to_tensorboard_log(metrics, writer, global_step, prefix=''):
        writer.add_scalar('metrics/{}/mje mean'.format(prefix), metrics['MPJPE [mm]'], global_step)
        writer.add_scalar('metrics/{}/mje pa mean'.format(prefix), metrics['PA-MPJPE [mm]'], global_step)
        writer.add_scalar('metrics/{}/mae mean'.format(prefix), metrics['MPJAE [deg]'], global_step)
