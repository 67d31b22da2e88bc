"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/dhd_b200.h declares; host-only entry points validate their arguments."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from dhd_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'dhd_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dhd_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert 'dhd_mghs_pool_fwd' in names and 'dhd_bev_pool_v2_fwd' in names
    for n in names:
        assert hasattr(lib, n), 'libdhd_b200.so does not export %s' % n
    from dhd_b200 import _lib
    assert sorted(_lib.exported_symbols()) == names


def test_workspace_size_and_validation(lib):
    from dhd_b200._lib import MghsCfg
    cfg = MghsCfg()
    cfg.B, cfg.N, cfg.D, cfg.fH, cfg.fW, cfg.C = 4, 6, 44, 16, 44, 64
    cfg.Dx = cfg.Dy = 200
    cfg.n_pass = 4
    for p, dz in enumerate((1, 4, 4, 8)):
        cfg.dz[p] = dz
        cfg.mask_id[p] = p
    n = lib.dhd_mghs_workspace_bytes(ctypes.byref(cfg))
    F = 4 * 6 * 44 * 16 * 44
    assert n >= F * (12 + 32) and n < F * 64 + (8 << 20)
    cfg.C = 80                              # unsupported channel count -> 0 + message
    assert lib.dhd_mghs_workspace_bytes(ctypes.byref(cfg)) == 0
    assert b'C == 64' in lib.dhd_last_error()
    cfg.C = 64
    cfg.dz[3] = 40                          # 1+4+4+40 planes > DHD_MAX_PLANES
    assert lib.dhd_mghs_workspace_bytes(ctypes.byref(cfg)) == 0


def test_null_pointer_is_an_error_not_a_crash(lib):
    rc = lib.dhd_bev_pool_v2_fwd(64, 5, None, None, None, None, None, None, None, None, None)
    assert rc != 0 and b'null' in lib.dhd_last_error()
    assert lib.dhd_bev_pool_v2_fwd(64, 0, None, None, None, None, None, None, None, None, None) == 0
    assert lib.dhd_bev_pool_v2_fwd(300, 1, None, None, None, None, None, None, None, None, None) != 0


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    import torch
    from dhd_b200 import pool
    z = torch.zeros(1)
    i = torch.zeros(1, dtype=torch.int32)
    with pytest.raises(RuntimeError, match='CUDA'):
        pool.bev_pool_v2(z.view(1, 1, 1, 1, 1), z.view(1, 1, 1, 1, 1), i, i, i, (1, 1, 1, 1, 1), i, i)


def test_product_never_imports_oracle():
    for base in ('dhd_b200', 'projects'):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h')):
                    src = open(os.path.join(dp, f)).read()
                    assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), \
                        '%s imports the oracle' % os.path.join(dp, f)


def test_ctypes_struct_layouts_match_the_library(lib):
    """The ctypes mirrors of the ABI structs have the sizes the compiled library reports."""
    import ctypes
    from dhd_b200 import dense as D
    from dhd_b200._lib import MghsCfg
    from dhd_b200.stereo import StereoDesc
    from dhd_b200.train import PackDesc
    for which, cls in enumerate((MghsCfg, D.ConvSeg, D.ConvDesc, D.WgradDesc, StereoDesc, D.PredictorTailDesc, PackDesc)):
        assert lib.dhd_abi_sizeof(which) == ctypes.sizeof(cls), cls.__name__
    # the same fields at the same offsets for the two structs that grew this round
    assert D.ConvDesc.stride.offset == D.ConvDesc.seg.offset + D.MAX_SEGS * ctypes.sizeof(D.ConvSeg)
    assert D.WgradDesc.x_stride.offset == D.WgradDesc.accumulate.offset + 4


def test_ctypes_field_offsets_match_the_c_header(tmp_path):
    """Every field of every ABI struct sits at the offset the C compiler gives it in include/dhd_b200.h: a tiny C program
    prints offsetof() for each field (names parsed from the header) and the ctypes mirrors must agree one by one."""
    import ctypes
    import shutil
    import subprocess
    from dhd_b200 import dense as D
    from dhd_b200._lib import MghsCfg
    from dhd_b200.stereo import StereoDesc
    from dhd_b200.train import PackDesc
    if shutil.which('gcc') is None:
        pytest.skip('no C compiler')
    header = os.path.join(ROOT, 'include', 'dhd_b200.h')
    text = re.sub(r'/\*.*?\*/', '', open(header).read(), flags=re.S)
    structs = (('dhd_mghs_cfg', MghsCfg), ('dhd_conv_seg', D.ConvSeg), ('dhd_conv_desc', D.ConvDesc),
               ('dhd_wgrad_desc', D.WgradDesc), ('dhd_stereo_desc', StereoDesc),
               ('dhd_predictor_tail_desc', D.PredictorTailDesc), ('dhd_pack_desc', PackDesc))
    lines, expect = [], []
    for cname, cls in structs:
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), text, flags=re.S).group(1)
        names = []
        for decl in body.split(';'):
            parts = [p for p in decl.strip().split(',') if p.strip()]
            if not parts:
                continue
            names.append(re.sub(r'\[.*', '', parts[0].split()[-1]).lstrip('*'))
            names += [re.sub(r'\[.*', '', p.strip()).lstrip('*').strip() for p in parts[1:]]
        py_names = [f[0] for f in cls._fields_]
        assert [n.rstrip('_') for n in py_names] == names, cname          # `in_` mirrors the C field `in`
        for c_field, py_field in zip(names, py_names):
            lines.append('  printf("%%zu\\n", offsetof(%s, %s));' % (cname, c_field))
            expect.append(('%s.%s' % (cname, c_field), getattr(cls, py_field).offset))
        lines.append('  printf("%%zu\\n", sizeof(%s));' % cname)
        expect.append((cname + ' (sizeof)', ctypes.sizeof(cls)))
    src = tmp_path / 'offsets.c'
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "%s"\nint main(void) {\n%s\n  return 0;\n}\n'
                   % (header, '\n'.join(lines)))
    exe = tmp_path / 'offsets'
    subprocess.check_call(['gcc', '-std=c11', '-o', str(exe), str(src)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert len(got) == len(expect)
    for (what, want), have in zip(expect, got):
        assert have == want, '%s: C says %d, ctypes says %d' % (what, have, want)
