"""Build libstellarscope_b200.so in-tree with nvcc for sm_100a.

The shared library is the product's only native artefact: hand-written CUDA
kernels plus the extern "C" boundary declared in include/stellarscope_b200.h.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.environ.get('STELLARSCOPE_B200_LIB') or os.path.join(HERE, 'libstellarscope_b200.so')
SOURCES = ['ss_api.cu', 'ss_layout.cu', 'ss_build.cu', 'ss_em.cu', 'ss_dense.cu', 'ss_reassign.cu', 'ss_count.cu', 'ss_umi.cu', 'ss_scores.cu', 'ss_mtx.cu']
HEADERS = ['ss_common.cuh', 'ss_em.cuh', 'ss_handle.h', 'ss_mtx_fmt.cuh', os.path.join('..', '..', 'include', 'stellarscope_b200.h')]
HOSTUTIL_SRC = os.path.join(CSRC, 'ss_hostutil.c')
HOSTUTIL = os.path.join(HERE, '_ss_hostutil.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def needs_build():
    if not os.path.exists(LIB):
        return True
    if os.environ.get('STELLARSCOPE_B200_LIB'):
        return False                  # an explicitly chosen library (a tuning variant) is used as it is
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(bdir, s.replace('.cu', '.o'))
        objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, s), '-o', o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(cmd)
    return LIB


def hostutil_needs_build():
    return (not os.path.exists(HOSTUTIL)) or os.path.getmtime(HOSTUTIL_SRC) > os.path.getmtime(HOSTUTIL)


def build_hostutil(force=False):
    """CPython helper of the Python boundary (plain C, gcc): row -> model map from the
    reference's dict of row sets."""
    if not force and not hostutil_needs_build():
        return HOSTUTIL
    import sysconfig
    inc = sysconfig.get_paths()['include']
    cmd = [os.environ.get('CC', 'gcc'), '-O2', '-fPIC', '-shared', '-I', inc, HOSTUTIL_SRC, '-o', HOSTUTIL]
    subprocess.check_call(cmd)
    return HOSTUTIL


if __name__ == '__main__':
    print(build_hostutil(force='--force' in sys.argv))
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
