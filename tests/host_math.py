"""ctypes wrapper around tests/host_harness.cpp (the CUDA per-frame math compiled for the host)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_LIB = None

_F32_FIELDS = ['v_template', 'shapedirs', 'posedirs', 'j0', 'jdirs', 'skin_weight', 'jt_weight']
_I32_FIELDS = ['skin_joint', 'jt_ptr', 'jt_vert', 'parents', 'faces', 'sensor_vert', 'helper_vert', 'sensor_faces',
               'sensor_degree', 'vj_ptr', 'jvj_ptr', 'vinc_ptr', 'vinc_item', 'vinc_code']


class HostSub(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int) for n in ('n_verts', 'vp_dim', 'n_faces', 'max_degree', 'n_skin')] +
                [(n, ctypes.POINTER(ctypes.c_float)) for n in _F32_FIELDS] +
                [(n, ctypes.POINTER(ctypes.c_int)) for n in _I32_FIELDS] + [('n_vj', ctypes.c_int), ('use_static_tree', ctypes.c_int)])


def _lib():
    global _LIB
    if _LIB is None:
        out_dir = os.path.join(HERE, '_build')
        os.makedirs(out_dir, exist_ok=True)
        so = os.path.join(out_dir, 'host_harness.so')
        src = os.path.join(HERE, 'host_harness.cpp')
        hdr = os.path.join(ROOT, 'em-pose_b200', 'csrc', 'frame_math.h')
        hdr2 = os.path.join(ROOT, 'em-pose_b200', 'csrc', 'metrics_math.h')
        hdr3 = os.path.join(ROOT, 'em-pose_b200', 'csrc', 'fan_math.h')
        if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2), os.path.getmtime(hdr3)):
            subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-I', os.path.dirname(hdr), src,
                                   '-o', so])
        _LIB = ctypes.CDLL(so)
    return _LIB


def frame_eval(sub, theta, beta, off_r, off_t, meas_pos, meas_ori, active, coef, use_pos=True, use_ori=True,
               want_grad=True, use_double=False, static_tree=True, sensor_weight=1.0, joints_gt=None,
               joint_weight=0.0):
    """Run the per-frame math on the host.  All per-frame inputs are (n, ...) float32 arrays."""
    keep = []
    hs = HostSub()
    d = sub['dims']
    hs.n_verts, hs.vp_dim, hs.n_faces, hs.max_degree, hs.n_skin = (d['n_verts'], d['vp_dim'], d['n_faces'],
                                                                   d['max_degree'], d['n_skin'])
    for n in _F32_FIELDS:
        a = np.ascontiguousarray(sub['sub.' + n], dtype=np.float32)
        keep.append(a)
        setattr(hs, n, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    for n in _I32_FIELDS:
        a = np.ascontiguousarray(sub['sub.' + n], dtype=np.int32)
        keep.append(a)
        setattr(hs, n, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    hs.n_vj = int(sub['sub.vj_ptr'].shape[0]) - 1
    hs.use_static_tree = int(static_tree)
    n = theta.shape[0]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    theta, beta, off_r, off_t, meas_pos, meas_ori, coef = map(f32, (theta, beta, off_r, off_t, meas_pos, meas_ori, coef))
    active = np.ascontiguousarray(active, dtype=np.int32)
    out = {'sensor_pos': np.zeros((n, 12, 3)), 'sensor_ori': np.zeros((n, 12, 3, 3)), 'joints': np.zeros((n, 22, 3)),
           'g_theta': np.zeros((n, 66)), 'g_beta': np.zeros((n, 10)), 'verts': np.zeros((n, d['n_verts'], 3))}
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    jgt = None if joints_gt is None else f32(joints_gt)
    rc = _lib().host_frame_eval(ctypes.byref(hs), n, fp(theta), fp(beta), fp(off_r), fp(off_t), fp(meas_pos),
                                fp(meas_ori), active.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(use_pos),
                                int(use_ori), fp(coef), int(want_grad), int(use_double), ctypes.c_float(sensor_weight),
                                None if jgt is None else fp(jgt), ctypes.c_float(joint_weight), dp(out['sensor_pos']),
                                dp(out['sensor_ori']), dp(out['joints']), dp(out['g_theta']), dp(out['g_beta']),
                                dp(out['verts']))
    assert rc == 0
    return out


class HostFan(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int) for n in ('ok', 'slots', 'max_deg', 'n_part')] +
                [(n, ctypes.POINTER(ctypes.c_int)) for n in ('deg', 'helper', 'n_joints', 'part_ptr', 'joint')] +
                [('weight', ctypes.POINTER(ctypes.c_float))] +
                [(n, ctypes.POINTER(ctypes.c_int)) for n in ('jp_ptr', 'jp_idx')])


def fan_eval(sub, theta, beta, off_r, off_t, meas_pos, meas_ori, active, coef, use_pos=True, use_ori=True,
             want_grad=True, use_double=False, static_tree=True, sensor_weight=1.0, joints_gt=None, joint_weight=0.0,
             force_maxd=0, cta_forms=True):
    """The fan-form pass (csrc/fan_math.h, what the production kernel runs) on the host; same contract as frame_eval."""
    keep = []
    hs = HostSub()
    d = sub['dims']
    hs.n_verts, hs.vp_dim, hs.n_faces, hs.max_degree, hs.n_skin = (d['n_verts'], d['vp_dim'], d['n_faces'],
                                                                   d['max_degree'], d['n_skin'])
    for n in _F32_FIELDS:
        a = np.ascontiguousarray(sub['sub.' + n], dtype=np.float32)
        keep.append(a)
        setattr(hs, n, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    for n in _I32_FIELDS:
        a = np.ascontiguousarray(sub['sub.' + n], dtype=np.int32)
        keep.append(a)
        setattr(hs, n, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    hs.n_vj = int(sub['sub.vj_ptr'].shape[0]) - 1
    hs.use_static_tree = int(static_tree)
    hf = HostFan()
    hf.ok, hf.slots, hf.max_deg, hf.n_part = [int(v) for v in sub['sub.fan_dims']]
    if cta_forms:
        hf.ok |= 2          # reduce / local gradients in the "thread owns an item of every frame" form the kernel runs
    for field, key in (('deg', 'sensor_degree'), ('helper', 'fan_helper'), ('n_joints', 'fan_n_joints'),
                       ('part_ptr', 'fan_part_ptr'), ('joint', 'fan_joint'), ('jp_ptr', 'fan_jp_ptr'), ('jp_idx', 'fan_jp_idx')):
        a = np.ascontiguousarray(sub['sub.' + key], dtype=np.int32)
        keep.append(a)
        setattr(hf, field, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    w = np.ascontiguousarray(sub['sub.fan_weight'], dtype=np.float32)
    keep.append(w)
    hf.weight = w.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    n = theta.shape[0]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    theta, beta, off_r, off_t, meas_pos, meas_ori, coef = map(f32, (theta, beta, off_r, off_t, meas_pos, meas_ori, coef))
    active = np.ascontiguousarray(active, dtype=np.int32)
    out = {'sensor_pos': np.zeros((n, 12, 3)), 'sensor_ori': np.zeros((n, 12, 3, 3)), 'joints': np.zeros((n, 22, 3)),
           'g_theta': np.zeros((n, 66)), 'g_beta': np.zeros((n, 10))}
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    jgt = None if joints_gt is None else f32(joints_gt)
    rc = _lib().host_fan_eval(ctypes.byref(hs), ctypes.byref(hf), n, fp(theta), fp(beta), fp(off_r), fp(off_t), fp(meas_pos),
                              fp(meas_ori), active.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(use_pos),
                              int(use_ori), fp(coef), int(want_grad), int(use_double), int(force_maxd),
                              ctypes.c_float(sensor_weight), None if jgt is None else fp(jgt), ctypes.c_float(joint_weight),
                              dp(out['sensor_pos']), dp(out['sensor_ori']), dp(out['joints']), dp(out['g_theta']), dp(out['g_beta']))
    assert rc == 0, rc
    return out


def metrics_eval(sub, pose, shape, pose_hat, shape_hat):
    """The metrics kernel's per-frame math on the host: eucl (n,22), eucl_pa (n,22), ground-truth joints (n,22,3)."""
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    pose, shape, pose_hat, shape_hat = map(f32, (pose, shape, pose_hat, shape_hat))
    j0, jd = f32(sub['sub.j0']), f32(sub['sub.jdirs'])
    par = np.ascontiguousarray(sub['sub.parents'], dtype=np.int32)
    n = pose.shape[0]
    eucl, pa, joints = np.zeros((n, 22), np.float32), np.zeros((n, 22), np.float32), np.zeros((n, 22, 3), np.float32)
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    rc = _lib().host_metrics_eval(fp(j0), fp(jd), par.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n, fp(pose), fp(shape),
                                  fp(pose_hat), fp(shape_hat), fp(eucl), fp(pa), fp(joints))
    assert rc == 0
    return eucl, pa, joints
