"""The metrics engine behind the reference's interface, computed and AGGREGATED on the B200.

Interface of ``MetricsEngine`` (``empose/eval/metrics.py:69-345``): constructor, ``reset`` / ``compute`` /
``compute_joint_dist`` / ``get_metrics`` / ``to_pretty_string`` / ``to_tensorboard_log``, the joint selections and the six
numbers ``get_metrics`` returns.  The implementation shares nothing with it:

* per frame, FK of the 22 body joints for ground truth and prediction, Euclidean distances, the Procrustes alignment
  (the reference: a numpy SVD per frame on the host, ``metrics.py:115-125``) and the angular distances are ONE kernel
  (``empose_metrics_compute``);
* the tables never leave the device: every ``compute`` folds them into running per-joint sums and sums of squares
  (float64, on the device), and ``get_metrics`` turns those few numbers into mean / standard deviation -- the same
  statistics the reference takes over its concatenated host tables (mean over joints of per-joint means; population
  standard deviation over all selected entries), without the (frames x joints) device-to-host traffic and without
  keeping every frame of the dataset in host memory.

``keep_tables=True`` additionally keeps the per-frame tables (on the device) for inspection; ``eucl_dists`` /
``eucl_dists_pa`` / ``angle_diffs`` then return them as numpy arrays like the reference's attributes of the same name.
"""
import numpy as np
import torch
from tabulate import tabulate

from empose_b200 import lib as _lib
from empose_b200.helpers.configuration import CONSTANTS as C

#: body joints in SMPL order (configuration.py:115-117)
SMPL_JOINTS = ('root l_hip r_hip spine1 l_knee r_knee spine2 l_ankle r_ankle spine3 l_foot r_foot neck l_collar r_collar head '
               'l_shoulder r_shoulder l_elbow r_elbow l_wrist r_wrist').split()
#: joints NOT evaluated (metrics.py:78-86 lists the complements): feet for positions; root, ankles, feet, wrists for angles
_NO_POSITION = {'l_foot', 'r_foot'}
_NO_ANGLE = {'root', 'l_ankle', 'r_ankle', 'l_foot', 'r_foot', 'l_wrist', 'r_wrist'}
_KINDS = ('eucl', 'eucl_pa', 'angle')


class _RunningMoments(object):
    """Per-column count, sum and sum of squares of every table seen so far (float64, on the tables' device)."""

    def __init__(self):
        self.rows = 0
        self.s1 = None
        self.s2 = None

    def add(self, table):
        t = table.double()
        s1, s2 = t.sum(dim=0), (t * t).sum(dim=0)
        self.s1 = s1 if self.s1 is None else self.s1 + s1
        self.s2 = s2 if self.s2 is None else self.s2 + s2
        self.rows += int(table.shape[0])

    def mean_and_std(self, columns):
        """Mean over `columns` of the per-column means, and the population standard deviation over all their entries."""
        if self.rows == 0:
            return 0.0, 0.0
        idx = torch.as_tensor(columns, dtype=torch.long, device=self.s1.device)
        s1, s2 = self.s1.index_select(0, idx).cpu().numpy(), self.s2.index_select(0, idx).cpu().numpy()
        mean_of_means = float(np.mean(s1 / self.rows))
        n = self.rows * len(columns)
        grand = s1.sum() / n
        return mean_of_means, float(np.sqrt(max(s2.sum() / n - grand * grand, 0.0)))


class MetricsEngine(object):
    """Accumulates MPJPE, PA-MPJPE and MPJAE over a dataset (``metrics.py:69``)."""

    def __init__(self, smpl_model, keep_tables=False):
        self.smpl_model = smpl_model
        self.keep_tables = keep_tables
        self.eucl_eval_joints = [j for j in SMPL_JOINTS if j not in _NO_POSITION]
        self.angle_eval_joints = [j for j in SMPL_JOINTS if j not in _NO_ANGLE]
        self.eucl_idxs = [SMPL_JOINTS.index(j) for j in self.eucl_eval_joints]
        self.angle_idxs = [SMPL_JOINTS.index(j) - 1 for j in self.angle_eval_joints]       # angles exclude the root: joint j is column j-1
        self.angle_glob = True
        self.reset()

    def reset(self):
        self._moments = {k: _RunningMoments() for k in _KINDS}
        self._tables = {k: [] for k in _KINDS}

    # the reference's attribute names, as host tables (only with keep_tables=True)
    def _host_tables(self, kind):
        if not self.keep_tables:
            raise AttributeError('per-frame tables are not kept: construct MetricsEngine(smpl, keep_tables=True)')
        return [t.cpu().numpy().astype(np.float64 if kind == 'angle' else np.float32) for t in self._tables[kind]]

    eucl_dists = property(lambda self: self._host_tables('eucl'))
    eucl_dists_pa = property(lambda self: self._host_tables('eucl_pa'))
    angle_diffs = property(lambda self: self._host_tables('angle'))

    def _fold(self, **tables):
        for kind, t in tables.items():
            if t is None:
                continue
            self._moments[kind].add(t)
            if self.keep_tables:
                self._tables[kind].append(t)

    @staticmethod
    def _valid_rows(seq_lengths, n, f, frame_mask, device):
        """Flat indices (into N*F) of the frames that count (``metrics.py:163-181``): inside the sequence and, if a mask is
        given, flagged valid -- for an (N, F, M) mask, valid in all M entries."""
        keep = torch.ones(n, f, dtype=torch.bool, device=device)
        if seq_lengths is not None:
            keep &= torch.arange(f, device=device).unsqueeze(0) < seq_lengths.to(device).reshape(n, 1)
        if frame_mask is not None:
            fm = frame_mask.to(device=device, dtype=torch.bool)
            if fm.dim() == 3:
                fm = fm.all(dim=-1)
            elif fm.dim() != 2:
                raise ValueError('frame_mask must have shape (N, F) or (N, F, M)')
            keep &= fm
        return torch.nonzero(keep.reshape(-1), as_tuple=False).reshape(-1)

    @staticmethod
    def _rows(t, n, f, rows):
        """(N, F, D) or per-sequence (N, D) -> (len(rows), D)."""
        if t.dim() == 2:
            return t.index_select(0, torch.div(rows, f, rounding_mode='floor'))
        return t.reshape(n * f, t.shape[-1]).index_select(0, rows)

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
        """``metrics.py:183-241``, same arguments: pose / pose_hat (N, F, 63) without the root, shape (N, 10) or (N, F, 10),
        optional root poses (N, F, 3), sequence lengths and frame mask."""
        n, f = int(pose.shape[0]), int(pose.shape[1])
        rows = self._valid_rows(seq_lengths, n, f, frame_mask, pose.device)
        if rows.numel() == 0:
            return
        take = lambda t: self._rows(t, n, f, rows)
        zeros = torch.zeros(rows.numel(), 3, dtype=pose.dtype, device=pose.device)
        full = torch.cat([zeros if pose_root is None else take(pose_root), take(pose)], dim=-1)
        full_hat = torch.cat([zeros if pose_root is None else take(pose_root_hat), take(pose_hat)], dim=-1)
        eucl, eucl_pa, angle = self._context(pose.device).metrics(full, take(shape), full_hat, take(shape if shape_hat is None else shape_hat),
                                                                  angle_local=not self.angle_glob)
        self._fold(eucl=eucl, eucl_pa=eucl_pa, angle=angle)

    def compute_joint_dist(self, joints, joints_hat, seq_lengths=None, frame_mask=None):
        """``metrics.py:243-265``: the two position metrics from given joints (N, F, >= 66)."""
        n, f = int(joints.shape[0]), int(joints.shape[1])
        rows = self._valid_rows(seq_lengths, n, f, frame_mask, joints.device)
        if rows.numel() == 0:
            return
        width = (C.N_JOINTS + 1) * 3
        eucl, eucl_pa = _lib.metrics_from_joints(self._rows(joints, n, f, rows)[:, :width], self._rows(joints_hat, n, f, rows)[:, :width])
        self._fold(eucl=eucl, eucl_pa=eucl_pa)

    def get_metrics(self, eucl_idxs_select=True, angle_idxs_select=True):
        """The six numbers of ``metrics.py:287-330``: means in mm / degrees, standard deviations over all entries."""
        pos_cols = self.eucl_idxs if eucl_idxs_select else list(range(C.N_JOINTS + 1))
        ang_cols = self.angle_idxs if angle_idxs_select else list(range(C.N_JOINTS))
        out = {}
        for label, kind in (('MPJPE', 'eucl'), ('PA-MPJPE', 'eucl_pa')):
            mean, std = self._moments[kind].mean_and_std(pos_cols)
            out[label + ' [mm]'], out[label + ' STD'] = mean * 1000.0, std * 1000.0
        out['MPJAE [deg]'], out['MPJAE STD'] = self._moments['angle'].mean_and_std(ang_cols)
        return out

    @staticmethod
    def to_pretty_string(metrics, model_name):
        """One table row (``metrics.py:332-339``)."""
        names = list(metrics)
        return tabulate([[model_name] + [metrics[k] for k in names]], headers=['Model'] + names)

    @staticmethod
    def to_tensorboard_log(metrics, writer, global_step, prefix=''):
        """The three means under the reference's tags (``metrics.py:341-345``)."""
        for tag, key in (('mje mean', 'MPJPE [mm]'), ('mje pa mean', 'PA-MPJPE [mm]'), ('mae mean', 'MPJAE [deg]')):
            writer.add_scalar('metrics/{}/{}'.format(prefix, tag), metrics[key], global_step)
