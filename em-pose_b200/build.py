"""Build ``libempose_b200.so`` (the C-ABI library of include/empose_b200.h) in-tree with nvcc for sm_100a.

    python em-pose_b200/build.py            # or empose_b200.build.build_library()

The library lands next to this file so it travels with a snapshot of the repository; nothing is
installed into site-packages.  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libempose_b200.so')
SOURCES = ['model.cu', 'metrics.cu', 'rnn_model.cu', 'rnn_persistent.cu', 'train.cu', 'train_kernels.cu', 'smpl_full.cu', 'frame_kernels.cu', 'fan_kernel.cu', 'gemm_tc.cu', 'gemm_simt.cu']
HEADERS = ['common.cuh', 'frame_math.h', 'fan_math.h', 'frame_kernels.h', 'gemm_jobs.h', 'gemm_tc.h', 'model_internal.h', 'train_kernels.h', 'rnn_persistent.h', 'metrics_math.h',
           os.path.join('..', '..', 'include', 'empose_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    obj_dir = os.path.join(HERE, 'build')
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in HEADERS]
    objects = []
    procs = []
    for src in SOURCES:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        objects.append(obj)
        if force or _stale(obj, [src_path] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + ['-c', src_path, '-o', obj]
            if verbose:
                print(' '.join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out.decode()))
    if force or procs or _stale(LIB_PATH, objects):
        cmd = [_nvcc(), '-shared', '-o', LIB_PATH] + objects + ['-gencode', 'arch=compute_100a,code=sm_100a']
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose=True))
