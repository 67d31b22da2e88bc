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


def _compile(src, obj, verbose, ptxas_info):
    cmd = ['nvcc', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
           '-lineinfo', '-Xcompiler', '-fPIC', '-c', src, '-o', obj]
    if ptxas_info:
        cmd += ['-Xptxas', '-v']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj


def build(force=False, verbose=False, ptxas_info=False):
    """Per-file incremental: a source is recompiled when it, a shared header or the ABI header is newer than its
    object; the compiles run in parallel (one nvcc process per file)."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.join(_HERE, 'build'), exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(_HERE, '..', 'include', 'dhd_b200.h')]
    h_time = max(os.path.getmtime(h) for h in headers)
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(_HERE, 'build', os.path.basename(src) + '.o')
        objs.append(obj)
        if force or ptxas_info or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), h_time):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for f in [ex.submit(_compile, s, o, verbose, ptxas_info) for s, o in jobs]:
            f.result()
    # objects of deleted sources must not linger in the library
    for stale in set(glob.glob(os.path.join(_HERE, 'build', '*.cu.o'))) - set(objs):
        os.remove(stale)
    cmd = ['nvcc', '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + \
        ['-lcuda']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True, ptxas_info='-v' in sys.argv))
