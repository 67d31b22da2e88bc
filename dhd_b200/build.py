"""Builds dhd_b200/libdhd_b200.so (sm_100a only) in-tree with plain nvcc.

The shared library is the drop-in boundary: an ``extern "C"`` ABI declared in
``include/dhd_b200.h``.  It has no torch / Python dependency.
"""
import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB = os.path.join(_HERE, 'libdhd_b200.so')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(_HERE, '..', 'include', 'dhd_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_info=False):
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(_HERE, 'build'), exist_ok=True)
    for src in sources():
        obj = os.path.join(_HERE, 'build', os.path.basename(src) + '.o')
        cmd = ['nvcc', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
               '-lineinfo', '-Xcompiler', '-fPIC', '-c', src, '-o', obj]
        if ptxas_info:
            cmd += ['-Xptxas', '-v']
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = ['nvcc', '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + \
        ['-lcuda']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True, ptxas_info='-v' in sys.argv))
