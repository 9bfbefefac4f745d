"""ctypes binding of the C ABI declared in ``include/empose_b200.h``.

This is the only place Python touches the native library.  PyTorch is used for device memory and
streams only: tensors are handed over as raw device pointers (``tensor.data_ptr()``) together with
the current CUDA stream.  There is no fallback of any kind: if ``libempose_b200.so`` is missing or
no CUDA device is present, calls raise.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libempose_b200.so')

PRECISION_TF32 = 0      # tcgen05 kind::tf32 everywhere (training default)
PRECISION_FP32 = 1      # FFMA executor, exact arithmetic
PRECISION_FP16 = 2      # inference: fp16 operands for the learned layers (kind::f16), pose blend 3xTF32
_DTYPES = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.int64): 2}

#: every symbol include/empose_b200.h declares (checked by tests/test_cabi.py)
EXPORTED_SYMBOLS = ('empose_abi_version', 'empose_last_error', 'empose_set_option', 'empose_ief_create', 'empose_ief_destroy',
                    'empose_ief_forward', 'empose_ief_forward_host', 'empose_ief_submit_host', 'empose_ief_wait_host', 'empose_sensor_project',
                    'empose_ief_last_launch_count', 'empose_ief_set_profiling', 'empose_ief_profile_read', 'empose_ief_profile_read_main',
                    'empose_gemm_selftest', 'empose_gemm_bench', 'empose_smpl_create', 'empose_smpl_destroy',
                    'empose_smpl_forward', 'empose_train_layout', 'empose_train_sizes', 'empose_train_create',
                    'empose_train_destroy', 'empose_train_forward', 'empose_train_backward', 'empose_train_loss_values',
                    'empose_train_last_launch_count', 'empose_train_set_sync_batchnorm', 'empose_rnn_create', 'empose_rnn_destroy', 'empose_rnn_forward',
                    'empose_rnn_last_launch_count', 'empose_sensors_create', 'empose_metrics_compute', 'empose_metrics_joints')


#: empose_allreduce_fn of include/empose_b200.h: int fn(void* user, double* device_buf, int64_t count, void* stream)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)


class _DeviceDoubles(object):
    """A device buffer of the library seen through ``__cuda_array_interface__`` (zero-copy ``torch.as_tensor``)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {'shape': (int(count),), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}


class EmposeError(RuntimeError):
    pass


class Tensor(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char_p), ('data', ctypes.c_void_p), ('dtype', ctypes.c_int32),
                ('ndim', ctypes.c_int32), ('shape', ctypes.c_int64 * 4)]


class IefConfig(ctypes.Structure):
    _fields_ = [('n_markers', ctypes.c_int32), ('num_iterations', ctypes.c_int32), ('step_size', ctypes.c_float),
                ('rnn_init', ctypes.c_int32), ('average_shape', ctypes.c_int32), ('use_gradient', ctypes.c_int32),
                ('use_marker_pos', ctypes.c_int32), ('use_marker_ori', ctypes.c_int32), ('hidden_size', ctypes.c_int32),
                ('num_layers', ctypes.c_int32), ('rnn_hidden_size', ctypes.c_int32), ('rnn_num_layers', ctypes.c_int32),
                ('skip_connections', ctypes.c_int32), ('batch_norm', ctypes.c_int32), ('precision', ctypes.c_int32),
                ('device', ctypes.c_int32)]


class LossWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ('pose_weight', 'shape_weight', 'reprojection_weight', 'fk_weight')]


class RnnConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('n_markers', 'hidden_size', 'num_layers', 'bidirectional', 'learn_init_state',
                                              'estimate_shape', 'shape_hidden_size', 'average_shape', 'do_fk',
                                              'use_marker_pos', 'use_marker_ori', 'precision', 'device')]


class History(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('pose', 'shape', 'joints', 'markers', 'markers_ori')]


_lib = None


def load():
    """Load the native library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EmposeError('native library %s not found: build it with `python em-pose_b200/build.py` '
                          '(or __graft_entry__.build()); there is no Python/CPU fallback' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    i32 = ctypes.c_int32
    lib.empose_abi_version.restype = ctypes.c_int
    lib.empose_last_error.restype = ctypes.c_char_p
    lib.empose_set_option.restype = ctypes.c_int
    lib.empose_set_option.argtypes = [ctypes.c_char_p, i32]
    lib.empose_ief_create.restype = ctypes.c_int
    lib.empose_ief_create.argtypes = [ctypes.POINTER(IefConfig), ctypes.POINTER(Tensor), i32, ctypes.POINTER(vp)]
    lib.empose_ief_destroy.restype = None
    lib.empose_ief_destroy.argtypes = [vp]
    fwd_args = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, ctypes.POINTER(History), vp]
    lib.empose_ief_forward.restype = ctypes.c_int
    lib.empose_ief_forward.argtypes = fwd_args
    lib.empose_ief_forward_host.restype = ctypes.c_int
    lib.empose_ief_forward_host.argtypes = fwd_args
    lib.empose_ief_submit_host.restype = ctypes.c_int
    lib.empose_ief_submit_host.argtypes = fwd_args[:-1] + [i32, vp]
    lib.empose_ief_wait_host.restype = ctypes.c_int
    lib.empose_ief_wait_host.argtypes = [vp, i32]
    lib.empose_sensor_project.restype = ctypes.c_int
    lib.empose_sensor_project.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.empose_ief_last_launch_count.restype = ctypes.c_int64
    lib.empose_ief_last_launch_count.argtypes = [vp]
    lib.empose_ief_set_profiling.restype = ctypes.c_int
    lib.empose_ief_set_profiling.argtypes = [vp, i32]
    lib.empose_ief_profile_read.restype = ctypes.c_int
    lib.empose_ief_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]
    lib.empose_ief_profile_read_main.restype = ctypes.c_int
    lib.empose_ief_profile_read_main.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]
    lib.empose_gemm_selftest.restype = ctypes.c_int
    lib.empose_gemm_selftest.argtypes = [i32, vp, ctypes.c_int64, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, i32, i32,
                                         i32, vp]
    lib.empose_smpl_create.restype = ctypes.c_int
    lib.empose_smpl_create.argtypes = [ctypes.POINTER(Tensor), i32, i32, i32, ctypes.POINTER(vp)]
    lib.empose_smpl_destroy.restype = None
    lib.empose_smpl_destroy.argtypes = [vp]
    lib.empose_smpl_forward.restype = ctypes.c_int
    lib.empose_smpl_forward.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.empose_gemm_bench.restype = ctypes.c_int
    lib.empose_gemm_bench.argtypes = [i32, vp, ctypes.c_int64, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, i32, i32, i32,
                                      i32, ctypes.POINTER(ctypes.c_float), vp]
    i64p = ctypes.POINTER(ctypes.c_int64)
    lib.empose_train_layout.restype = ctypes.c_int
    lib.empose_train_layout.argtypes = [ctypes.POINTER(IefConfig), i32, ctypes.c_char_p, i32, ctypes.POINTER(i32), i64p, i64p]
    lib.empose_train_sizes.restype = ctypes.c_int
    lib.empose_train_sizes.argtypes = [ctypes.POINTER(IefConfig), i64p, i64p]
    lib.empose_train_set_sync_batchnorm.restype = ctypes.c_int
    lib.empose_train_set_sync_batchnorm.argtypes = [vp, ALLREDUCE_FN, vp, i32]
    lib.empose_train_create.restype = ctypes.c_int
    lib.empose_train_create.argtypes = [ctypes.POINTER(IefConfig), ctypes.POINTER(Tensor), i32, vp, vp, vp, ctypes.POINTER(vp)]
    lib.empose_train_destroy.restype = None
    lib.empose_train_destroy.argtypes = [vp]
    lib.empose_train_forward.restype = ctypes.c_int
    lib.empose_train_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, ctypes.POINTER(History), vp]
    lib.empose_train_backward.restype = ctypes.c_int
    lib.empose_train_backward.argtypes = [vp, vp, vp, vp, ctypes.POINTER(LossWeights), ctypes.POINTER(ctypes.c_float), vp, vp]
    lib.empose_train_loss_values.restype = ctypes.c_int
    lib.empose_train_loss_values.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.empose_train_last_launch_count.restype = ctypes.c_int64
    lib.empose_train_last_launch_count.argtypes = [vp]
    lib.empose_sensors_create.restype = ctypes.c_int
    lib.empose_sensors_create.argtypes = [ctypes.POINTER(Tensor), i32, i32, i32, ctypes.POINTER(vp)]
    lib.empose_metrics_compute.restype = ctypes.c_int
    lib.empose_metrics_compute.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.empose_metrics_joints.restype = ctypes.c_int
    lib.empose_metrics_joints.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.empose_rnn_create.restype = ctypes.c_int
    lib.empose_rnn_create.argtypes = [ctypes.POINTER(RnnConfig), ctypes.POINTER(Tensor), i32, ctypes.POINTER(vp)]
    lib.empose_rnn_destroy.restype = None
    lib.empose_rnn_destroy.argtypes = [vp]
    lib.empose_rnn_forward.restype = ctypes.c_int
    lib.empose_rnn_forward.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.empose_rnn_last_launch_count.restype = ctypes.c_int64
    lib.empose_rnn_last_launch_count.argtypes = [vp]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise EmposeError('empose_b200 error %d: %s' % (rc, load().empose_last_error().decode()))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_tensor_table(arrays):
    """{name: numpy array} -> (ctypes array of Tensor, keep-alive list)."""
    keep = []
    table = (Tensor * len(arrays))()
    for i, (name, arr) in enumerate(arrays.items()):
        arr = np.ascontiguousarray(arr)
        if arr.dtype not in _DTYPES:
            raise EmposeError('tensor %s has unsupported dtype %s' % (name, arr.dtype))
        if arr.ndim > 4:
            arr = arr.reshape(arr.shape[0], -1)
        bname = name.encode()
        keep += [arr, bname]
        table[i].name = bname
        table[i].data = arr.ctypes.data
        table[i].dtype = _DTYPES[arr.dtype]
        table[i].ndim = arr.ndim
        for d in range(arr.ndim):
            table[i].shape[d] = arr.shape[d]
    return table, keep


class IefContext(object):
    """Owns one ``empose_ief*``: packed weights, sub-model and cached plans on one device."""

    def __init__(self, config, arrays):
        """
        :param config: dict with the fields of ``empose_ief_config``.
        :param arrays: {name: numpy array}: the reference's state-dict keys plus the ``sub.*`` arrays.
        """
        lib = load()
        cfg = _config_struct(config)
        self.config = dict(config)
        table, keep = make_tensor_table(arrays)
        handle = ctypes.c_void_p()
        _check(lib.empose_ief_create(ctypes.byref(cfg), table, len(arrays), ctypes.byref(handle)))
        del keep
        self._handle = handle
        self.n_iter = int(config['num_iterations'])
        self.rnn_layers = int(config['rnn_num_layers']) if config['rnn_init'] else 0
        self.rnn_hidden = int(config['rnn_hidden_size'])
        self.device_index = int(config['device'])

    def close(self):
        if getattr(self, '_handle', None):
            load().empose_ief_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_launch_count(self):
        return int(load().empose_ief_last_launch_count(self._handle))

    def set_profiling(self, enable):
        _check(load().empose_ief_set_profiling(self._handle, int(bool(enable))))

    def profile_read(self):
        """(summed device ms of the GEMM executor launches, number of launches) since the last read."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _check(load().empose_ief_profile_read(self._handle, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def profile_read_main(self):
        """The same for the per-frame sub-model kernel (main_kernel)."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _check(load().empose_ief_profile_read_main(self._handle, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def forward(self, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks=None, lstm_state=None,
                is_new_sequence=True, want_history=True, want_state=True, history_out=None):
        """
        Device tensors in, device tensors out (all float32 / int32, contiguous, on this context's device).
        :param want_state: False skips the LSTM state output (a batch of fresh windows never needs it).
        :param history_out: a ``history`` dict returned by an earlier call with the same (B, F): its tensors are overwritten
            instead of allocating 5 x (N+1) new ones.
        :return: dict pose (B,F,66), shape (B,F,10), joints (B,F,66), history (dict of (N+1,B,F,dof)) or None,
                 lstm_state (2,L,B,H) or None.
        """
        import torch
        dev = marker_pos.device
        if dev.type != 'cuda':
            raise EmposeError('inputs must be CUDA tensors (no CPU path)')
        b, f = int(marker_pos.shape[0]), int(marker_pos.shape[1])
        f32 = lambda t: t.to(dtype=torch.float32).contiguous()
        marker_pos, marker_oris = f32(marker_pos).reshape(b, f, 36), f32(marker_oris).reshape(b, f, 108)
        offset_r, offset_t = f32(offset_r).reshape(b, 12, 9), f32(offset_t).reshape(b, 12, 3)
        seq_lengths = seq_lengths.to(device=dev, dtype=torch.int32).contiguous()
        if marker_masks is not None:
            marker_masks = f32(marker_masks).reshape(b, f, 12)
        opts = dict(dtype=torch.float32, device=dev)
        pose = torch.empty((b, f, 66), **opts)
        shape = torch.empty((b, f, 10), **opts)
        joints = torch.empty((b, f, 66), **opts)
        hist, hist_struct = None, None
        if want_history:
            n1 = self.n_iter + 1
            dof = (('pose', 66), ('shape', 10), ('joints', 66), ('markers', 36), ('markers_ori', 108))
            if history_out is not None and all(history_out[k].shape == (n1, b, f, d) and history_out[k].device == dev
                                               for k, d in dof):
                hist = history_out
            else:
                hist = {k: torch.empty((n1, b, f, d), **opts) for k, d in dof}
            hist_struct = History(*[hist[k].data_ptr() for k, _ in dof])
        state = None
        if self.rnn_layers and (want_state or (lstm_state is not None and not is_new_sequence)):
            if lstm_state is not None and not is_new_sequence:
                state = f32(lstm_state).reshape(2, self.rnn_layers, b, self.rnn_hidden).clone()
            else:
                state = torch.zeros((2, self.rnn_layers, b, self.rnn_hidden), **opts)
        _check(load().empose_ief_forward(
            self._handle, _ptr(marker_pos), _ptr(marker_oris), _ptr(offset_r), _ptr(offset_t), _ptr(seq_lengths),
            _ptr(marker_masks), _ptr(state), int(bool(is_new_sequence)), b, f, _ptr(pose), _ptr(shape), _ptr(joints),
            ctypes.byref(hist_struct) if hist_struct is not None else None, _stream()))
        return {'pose': pose, 'shape': shape, 'joints': joints, 'history': hist, 'lstm_state': state}

    def forward_host(self, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks=None,
                     lstm_state=None, is_new_sequence=True, want_state=True):
        """Host (CPU, ideally pinned) tensors in and out through ``empose_ief_forward_host``; synchronises.
        ``want_state=False`` skips the download of the final LSTM state (2 x L x B x H floats), which only a caller that
        streams the next chunk of the same sequences needs (``models.py:489-492``)."""
        import torch
        b, f = int(marker_pos.shape[0]), int(marker_pos.shape[1])
        f32 = lambda t: t.to(dtype=torch.float32).contiguous()
        marker_pos, marker_oris = f32(marker_pos), f32(marker_oris)
        offset_r, offset_t = f32(offset_r), f32(offset_t)
        seq_lengths = seq_lengths.to(dtype=torch.int32).contiguous()
        if marker_masks is not None:
            marker_masks = f32(marker_masks)
        # outputs are allocated pinned (torch's caching host allocator): no staging copy, truly asynchronous downloads
        pose = torch.empty((b, f, 66), dtype=torch.float32, pin_memory=True)
        shape = torch.empty((b, f, 10), dtype=torch.float32, pin_memory=True)
        joints = torch.empty((b, f, 66), dtype=torch.float32, pin_memory=True)
        state = None
        if self.rnn_layers and (want_state or (lstm_state is not None and not is_new_sequence)):
            state = torch.empty((2, self.rnn_layers, b, self.rnn_hidden), dtype=torch.float32, pin_memory=True)
            if lstm_state is not None and not is_new_sequence:
                state.copy_(lstm_state.reshape(state.shape))
        with torch.cuda.device(self.device_index):
            _check(load().empose_ief_forward_host(
                self._handle, _ptr(marker_pos), _ptr(marker_oris), _ptr(offset_r), _ptr(offset_t), _ptr(seq_lengths),
                _ptr(marker_masks), _ptr(state), int(bool(is_new_sequence)), b, f, _ptr(pose), _ptr(shape),
                _ptr(joints), None, _stream()))
        return {'pose': pose, 'shape': shape, 'joints': joints, 'lstm_state': state}

    def submit_host(self, slot, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks=None, out=None):
        """Streaming form of :meth:`forward_host` for fresh windows (``empose_ief_submit_host``): enqueue one request into the
        in-flight ``slot`` (0..3) and return at once; :meth:`wait_host` makes the results valid.  Inputs must be pinned
        float32 / int32 host tensors that stay untouched until then.  ``out`` (the dict a previous call on this slot
        returned) is reused for the results, otherwise pinned tensors are allocated."""
        import torch
        b, f = int(marker_pos.shape[0]), int(marker_pos.shape[1])
        for t in (marker_pos, marker_oris, offset_r, offset_t):
            if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_pinned():
                raise EmposeError('submit_host takes pinned, contiguous float32 host tensors')
        if seq_lengths.dtype != torch.int32 or not seq_lengths.is_pinned():
            raise EmposeError('submit_host takes seq_lengths as a pinned int32 host tensor')
        if marker_masks is not None and (marker_masks.dtype != torch.float32 or not marker_masks.is_pinned()):
            raise EmposeError('submit_host takes marker_masks as a pinned float32 host tensor')
        if out is None or tuple(out['pose'].shape) != (b, f, 66):
            out = {'pose': torch.empty((b, f, 66), dtype=torch.float32, pin_memory=True),
                   'shape': torch.empty((b, f, 10), dtype=torch.float32, pin_memory=True),
                   'joints': torch.empty((b, f, 66), dtype=torch.float32, pin_memory=True)}
        out['slot'] = int(slot)
        out['inputs'] = (marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks)      # kept alive until the wait
        with torch.cuda.device(self.device_index):
            _check(load().empose_ief_submit_host(
                self._handle, _ptr(marker_pos), _ptr(marker_oris), _ptr(offset_r), _ptr(offset_t), _ptr(seq_lengths),
                _ptr(marker_masks), None, 1, b, f, _ptr(out['pose']), _ptr(out['shape']), _ptr(out['joints']), None, int(slot), _stream()))
        return out

    def wait_host(self, request):
        """Blocks until the request :meth:`submit_host` returned has its results in ``request['pose' | 'shape' | 'joints']``."""
        _check(load().empose_ief_wait_host(self._handle, int(request['slot'])))
        request.pop('inputs', None)
        return request

    def sensor_project(self, poses, shapes, offset_r, offset_t):
        """(R,66), (R,10), (R,12,3,3), (R,12,3) device tensors -> sensor_pos (R,12,3), sensor_ori (R,12,3,3), joints (R,22,3)."""
        import torch
        r = int(poses.shape[0])
        f32 = lambda t: t.to(dtype=torch.float32).contiguous()
        poses, shapes, offset_r, offset_t = f32(poses), f32(shapes), f32(offset_r), f32(offset_t)
        opts = dict(dtype=torch.float32, device=poses.device)
        pos = torch.empty((r, 12, 3), **opts)
        ori = torch.empty((r, 12, 3, 3), **opts)
        joints = torch.empty((r, 22, 3), **opts)
        _check(load().empose_sensor_project(self._handle, _ptr(poses), _ptr(shapes), _ptr(offset_r), _ptr(offset_t), r,
                                            _ptr(pos), _ptr(ori), _ptr(joints), _stream()))
        return pos, ori, joints


def _config_struct(config):
    cfg = IefConfig()
    for name, _ in IefConfig._fields_:
        setattr(cfg, name, config[name])
    return cfg


def train_layout(config):
    """The flat parameter / running-statistics layout of a configuration (``empose_train_layout``).
    :return: (entries, n_params, n_buffers); entries = [(state-dict key, kind, offset, numel)], kind 0 = parameter
             (params / grads vectors), 1 = BatchNorm running statistic (bn_buffers vector); sizes in floats."""
    lib = load()
    cfg = _config_struct(config)
    entries = []
    name = ctypes.create_string_buffer(256)
    kind, off, numel = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
    i = 0
    while True:
        rc = lib.empose_train_layout(ctypes.byref(cfg), i, name, 256, ctypes.byref(kind), ctypes.byref(off), ctypes.byref(numel))
        if rc == -2:
            break
        _check(rc)
        entries.append((name.value.decode(), int(kind.value), int(off.value), int(numel.value)))
        i += 1
    n_p, n_b = ctypes.c_int64(), ctypes.c_int64()
    _check(lib.empose_train_sizes(ctypes.byref(cfg), ctypes.byref(n_p), ctypes.byref(n_b)))
    return entries, int(n_p.value), int(n_b.value)


class TrainContext(object):
    """Owns one ``empose_train*``: the training step over caller-owned flat parameter / gradient vectors."""

    def __init__(self, config, arrays, params, grads, bn_buffers):
        """
        :param arrays: {name: numpy array}: state dict + ``sub.*`` arrays (as for ``IefContext``).
        :param params, grads, bn_buffers: flat float32 CUDA tensors laid out as ``train_layout(config)`` says; they
                                          must outlive this object (bn_buffers may be None without BatchNorm).
        """
        lib = load()
        cfg = _config_struct(config)
        self.config = dict(config)
        table, keep = make_tensor_table(arrays)
        handle = ctypes.c_void_p()
        self._keep_alive = (params, grads, bn_buffers)
        _check(lib.empose_train_create(ctypes.byref(cfg), table, len(arrays), _ptr(params), _ptr(grads), _ptr(bn_buffers),
                                       ctypes.byref(handle)))
        del keep
        self._handle = handle
        self.n_iter = int(config['num_iterations'])
        self.device_index = int(config['device'])

    def close(self):
        if getattr(self, '_handle', None):
            load().empose_train_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_launch_count(self):
        return int(load().empose_train_last_launch_count(self._handle))

    def set_sync_batchnorm(self, enabled=True, group=None):
        """SyncBatchNorm over ``torch.distributed`` (``empose_train_set_sync_batchnorm``): the library hands every
        BatchNorm's batch statistics to ``dist.all_reduce`` between its own kernels, on the current stream, so means,
        variances and running statistics are those of the global batch (every rank must run the same number of rows)."""
        import torch
        import torch.distributed as dist
        if not enabled:
            _check(load().empose_train_set_sync_batchnorm(self._handle, ALLREDUCE_FN(0), None, 1))
            self._sync_cb = None
            return
        if not (dist.is_available() and dist.is_initialized()):
            raise EmposeError('SyncBatchNorm needs an initialised torch.distributed process group')
        device = torch.device('cuda', self.device_index)
        self.sync_calls = 0

        def allreduce(user, buf, count, stream):
            try:
                t = torch.as_tensor(_DeviceDoubles(buf, count), device=device)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)      # ordered on the current stream, like the kernels around it
                self.sync_calls += 1
                return 0
            except Exception:                                              # never let an exception cross the C boundary
                return 1

        self._sync_cb = ALLREDUCE_FN(allreduce)                             # keep the trampoline alive
        _check(load().empose_train_set_sync_batchnorm(self._handle, self._sync_cb, None, dist.get_world_size(group)))

    def forward(self, marker_pos, marker_oris, offset_r, offset_t, seq_lengths, marker_masks=None, want_history=True):
        """Train-mode forward pass; same tensors in / out as ``IefContext.forward`` (no LSTM state carry)."""
        import torch
        dev = marker_pos.device
        if dev.type != 'cuda':
            raise EmposeError('inputs must be CUDA tensors (no CPU path)')
        b, f = int(marker_pos.shape[0]), int(marker_pos.shape[1])
        f32 = lambda t: t.to(dtype=torch.float32).contiguous()
        marker_pos, marker_oris = f32(marker_pos).reshape(b, f, 36), f32(marker_oris).reshape(b, f, 108)
        offset_r, offset_t = f32(offset_r).reshape(b, 12, 9), f32(offset_t).reshape(b, 12, 3)
        seq_lengths = seq_lengths.to(device=dev, dtype=torch.int32).contiguous()
        if marker_masks is not None:
            marker_masks = f32(marker_masks).reshape(b, f, 12)
        opts = dict(dtype=torch.float32, device=dev)
        pose = torch.empty((b, f, 66), **opts)
        shape = torch.empty((b, f, 10), **opts)
        joints = torch.empty((b, f, 66), **opts)
        hist, hist_struct = None, None
        if want_history:
            n1 = self.n_iter + 1
            hist = {'pose': torch.empty((n1, b, f, 66), **opts), 'shape': torch.empty((n1, b, f, 10), **opts),
                    'joints': torch.empty((n1, b, f, 66), **opts), 'markers': torch.empty((n1, b, f, 36), **opts),
                    'markers_ori': torch.empty((n1, b, f, 108), **opts)}
            hist_struct = History(*[hist[k].data_ptr() for k in ('pose', 'shape', 'joints', 'markers', 'markers_ori')])
        _check(load().empose_train_forward(
            self._handle, _ptr(marker_pos), _ptr(marker_oris), _ptr(offset_r), _ptr(offset_t), _ptr(seq_lengths),
            _ptr(marker_masks), b, f, _ptr(pose), _ptr(shape), _ptr(joints),
            ctypes.byref(hist_struct) if hist_struct is not None else None, _stream()))
        self._shape = (b, f)
        return {'pose': pose, 'shape': shape, 'joints': joints, 'history': hist}

    _LOSS_KEYS = ('pose', 'shape', 'reconstruction', 'fk', 'total_loss')

    def loss_values(self):
        """The five loss values of the last ``backward(..., defer=True)`` (``empose_train_loss_values``).  Synchronises."""
        vals = (ctypes.c_float * 5)()
        _check(load().empose_train_loss_values(self._handle, vals))
        return dict(zip(self._LOSS_KEYS, [float(v) for v in vals]))

    def backward(self, poses_gt, shapes_gt, joints_gt, pose_weight, shape_weight, reprojection_weight, fk_weight,
                 defer=False, dense_ready_event=None):
        """Adds the step's gradients to the flat ``grads`` vector; returns the five loss values of
        ``models.py:676-680`` as a dict and synchronises -- unless ``defer``: then the pass is only enqueued (read the values
        with ``loss_values()``), and ``dense_ready_event`` (a ``torch.cuda.Event`` that has been recorded once, so that its
        handle exists) is recorded on the current stream when every gradient except the LSTM's is final."""
        import torch
        b, f = self._shape
        f32 = lambda t: None if t is None else t.to(dtype=torch.float32).contiguous()
        poses_gt = f32(poses_gt).reshape(b, f, 66)
        shapes_gt = f32(shapes_gt).reshape(b, 10)
        joints_gt = None if joints_gt is None else f32(joints_gt).reshape(b, f, 66)
        w = LossWeights(float(pose_weight), float(shape_weight), float(reprojection_weight), float(fk_weight))
        vals = None if defer else (ctypes.c_float * 5)()
        event = ctypes.c_void_p(int(dense_ready_event.cuda_event)) if dense_ready_event is not None else None
        _check(load().empose_train_backward(self._handle, _ptr(poses_gt), _ptr(shapes_gt), _ptr(joints_gt), ctypes.byref(w),
                                            vals, event, _stream()))
        return None if defer else dict(zip(self._LOSS_KEYS, [float(v) for v in vals]))


class SensorContext(object):
    """Owns a sub-model-only context (``empose_sensors_create``): poses / shapes / offsets -> 12 sensor frames + 22 joints."""

    def __init__(self, arrays, precision, device_index):
        table, keep = make_tensor_table(arrays)
        handle = ctypes.c_void_p()
        _check(load().empose_sensors_create(table, len(arrays), int(precision), int(device_index), ctypes.byref(handle)))
        del keep
        self._handle = handle
        self.device_index = int(device_index)

    def close(self):
        if getattr(self, '_handle', None):
            load().empose_ief_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    sensor_project = IefContext.sensor_project

    def metrics(self, pose, shape, pose_hat, shape_hat, want_angle=True, angle_local=False):
        """(R,66), (R,10) x2 CUDA tensors -> eucl (R,22), eucl_pa (R,22), angle_deg (R,21) | None (``empose_metrics_compute``);
        ``angle_local``: angles between the local joint rotations instead of the global orientations."""
        import torch
        r = int(pose.shape[0])
        f32 = lambda t: t.to(dtype=torch.float32).contiguous()
        pose, shape, pose_hat, shape_hat = f32(pose), f32(shape), f32(pose_hat), f32(shape_hat)
        opts = dict(dtype=torch.float32, device=pose.device)
        eucl, eucl_pa = torch.empty((r, 22), **opts), torch.empty((r, 22), **opts)
        angle = torch.empty((r, 21), **opts) if want_angle else None
        _check(load().empose_metrics_compute(self._handle, _ptr(pose), _ptr(shape), _ptr(pose_hat), _ptr(shape_hat), r,
                                             int(bool(angle_local)), _ptr(eucl), _ptr(eucl_pa), _ptr(angle), _stream()))
        return eucl, eucl_pa, angle


def metrics_from_joints(joints, joints_hat):
    """(R,66) CUDA tensors -> eucl (R,22), eucl_pa (R,22) (``empose_metrics_joints``)."""
    import torch
    if joints.device.type != 'cuda':
        raise EmposeError('inputs must be CUDA tensors (no CPU path)')
    r = int(joints.shape[0])
    joints, joints_hat = joints.to(dtype=torch.float32).contiguous(), joints_hat.to(dtype=torch.float32).contiguous()
    eucl = torch.empty((r, 22), dtype=torch.float32, device=joints.device)
    eucl_pa = torch.empty_like(eucl)
    _check(load().empose_metrics_joints(_ptr(joints), _ptr(joints_hat), r, _ptr(eucl), _ptr(eucl_pa), _stream()))
    return eucl, eucl_pa


class RnnContext(object):
    """Owns one ``empose_rnn*``: the (Bi)RNN baseline (``SimpleRNN``, models.py:265-317) on one device."""

    def __init__(self, config, arrays):
        lib = load()
        cfg = RnnConfig()
        for name, _ in RnnConfig._fields_:
            setattr(cfg, name, int(config[name]))
        self.config = dict(config)
        table, keep = make_tensor_table(arrays)
        handle = ctypes.c_void_p()
        _check(lib.empose_rnn_create(ctypes.byref(cfg), table, len(arrays), ctypes.byref(handle)))
        del keep
        self._handle = handle
        self.n_state = int(config['num_layers']) * (2 if config['bidirectional'] else 1)
        self.hidden = int(config['hidden_size'])

    def close(self):
        if getattr(self, '_handle', None):
            load().empose_rnn_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_launch_count(self):
        return int(load().empose_rnn_last_launch_count(self._handle))

    def forward(self, marker_pos, marker_oris, seq_lengths, lstm_state=None, is_new_sequence=True):
        """CUDA tensors in / out: pose (B,F,66), shape (B,F,10) | None, joints (B,F,66) | None, lstm_state (2,L*dirs,B,H)."""
        import torch
        dev = marker_pos.device
        if dev.type != 'cuda':
            raise EmposeError('inputs must be CUDA tensors (no CPU path)')
        b, f = int(marker_pos.shape[0]), int(marker_pos.shape[1])
        f32 = lambda t: t.to(dtype=torch.float32).contiguous()
        marker_pos, marker_oris = f32(marker_pos).reshape(b, f, 36), f32(marker_oris).reshape(b, f, 108)
        seq_lengths = seq_lengths.to(device=dev, dtype=torch.int32).contiguous()
        opts = dict(dtype=torch.float32, device=dev)
        pose = torch.empty((b, f, 66), **opts)
        shape = torch.empty((b, f, 10), **opts) if self.config['estimate_shape'] else None
        joints = torch.empty((b, f, 66), **opts) if self.config['do_fk'] else None
        if lstm_state is not None and not is_new_sequence:
            state = f32(lstm_state).reshape(2, self.n_state, b, self.hidden).clone()
        else:
            state = torch.zeros((2, self.n_state, b, self.hidden), **opts)
        _check(load().empose_rnn_forward(self._handle, _ptr(marker_pos), _ptr(marker_oris), _ptr(seq_lengths), _ptr(state),
                                         int(bool(is_new_sequence)), b, f, _ptr(pose), _ptr(shape), _ptr(joints), _stream()))
        return {'pose': pose, 'shape': shape, 'joints': joints, 'lstm_state': state}


class SmplContext(object):
    """Owns one ``empose_smpl*``: the full-mesh SMPL-H constants on one device."""

    def __init__(self, arrays, precision, device_index):
        table, keep = make_tensor_table(arrays)
        handle = ctypes.c_void_p()
        _check(load().empose_smpl_create(table, len(arrays), int(precision), int(device_index), ctypes.byref(handle)))
        del keep
        self._handle = handle
        self.n_verts = int(arrays['smpl.dims'][0])
        self.device_index = int(device_index)

    def close(self):
        if getattr(self, '_handle', None):
            load().empose_smpl_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, poses_body, betas, poses_root=None, trans=None):
        """(N,63), (N,10), optional (N,3), (N,3) CUDA tensors -> verts (N,V,3), joints (N,52,3)."""
        import torch
        if poses_body.device.type != 'cuda':
            raise EmposeError('inputs must be CUDA tensors (no CPU path)')
        n = int(poses_body.shape[0])
        f32 = lambda t: None if t is None else t.to(dtype=torch.float32).contiguous()
        poses_body, betas, poses_root, trans = f32(poses_body), f32(betas), f32(poses_root), f32(trans)
        opts = dict(dtype=torch.float32, device=poses_body.device)
        verts = torch.empty((n, self.n_verts, 3), **opts)
        joints = torch.empty((n, 52, 3), **opts)
        _check(load().empose_smpl_forward(self._handle, _ptr(poses_root), _ptr(poses_body), _ptr(betas), _ptr(trans), n,
                                          _ptr(verts), _ptr(joints), _stream()))
        return verts, joints


def set_option(key, value):
    """Development switch of the library (``empose_set_option``): e.g. ``set_option('main_general', 1)``."""
    rc = load().empose_set_option(key.encode(), int(value))
    if rc != 0:
        raise EmposeError(load().empose_last_error().decode())


def gemm_selftest(a, w, bias, precision=PRECISION_TF32):
    """C = A . W^T + bias on the job executor; a (M,K), w (N,K), bias (N,) CUDA float32 tensors."""
    import torch
    a, w = a.contiguous(), w.contiguous()
    m, k = a.shape
    n = w.shape[0]
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    _check(load().empose_gemm_selftest(precision, _ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(c),
                                       c.stride(0), m, n, k, _stream()))
    return c


def gemm_bench(a, w, bias, precision=PRECISION_TF32, reps=10):
    """Mean device milliseconds per launch of C = A . W^T + bias on the job executor."""
    import torch
    a, w = a.contiguous(), w.contiguous()
    m, k = a.shape
    n = w.shape[0]
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    ms = ctypes.c_float()
    _check(load().empose_gemm_bench(precision, _ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(c),
                                    c.stride(0), m, n, k, reps, ctypes.byref(ms), _stream()))
    return ms.value
