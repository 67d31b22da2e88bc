"""TEST INFRASTRUCTURE -- builds the oracle's native pieces.

  oracle/_build/libdhd_oracle.so   gcc build of oracle/bev_pool_ref.c (our C restatement)
  oracle/_ref/libbev_pool_v2_ref.so  nvcc build of the REFERENCE's own, unmodified
        projects/mmdet3d_plugin/ops/bev_pool_v2/src/bev_pool_cuda.cu, compiled where it
        lies under /root/reference (that file has no torch dependency: it only defines the
        two kernels and the launchers `bev_pool_v2` / `bev_pool_v2_grad`).  It is the
        "reference bev_pool_v2 CUDA path" used on the GPU box as a second checker and as
        the CUDA baseline.  Never copied into the repo; `_ref/` is git-ignored.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_CU = '/root/reference/projects/mmdet3d_plugin/ops/bev_pool_v2/src/bev_pool_cuda.cu'


def build(verbose=False):
    out = os.path.join(_HERE, '_build')
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, 'libdhd_oracle.so')
    src = os.path.join(_HERE, 'bev_pool_ref.c')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ['gcc', '-O2', '-fPIC', '-shared', '-fopenmp', '-ffp-contract=off', '-o', so, src, '-lm']
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
    return so


def build_ref(verbose=False):
    """Compile the reference's own CUDA source (only where /root/reference exists)."""
    if not os.path.exists(REF_CU):
        return None
    out = os.path.join(_HERE, '_ref')
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, 'libbev_pool_v2_ref.so')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(REF_CU):
        cmd = ['nvcc', '-O3', '-shared', '-Xcompiler', '-fPIC', '-gencode',
               'arch=compute_100a,code=sm_100a', '-lineinfo', '-o', so, REF_CU]
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
    return so


if __name__ == '__main__':
    print(build(True))
    print(build_ref(True))
