"""ctypes binding of libdhd_b200.so (the C-ABI in include/dhd_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  The library is built in-tree by ``dhd_b200.build``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdhd_b200.so')

MAX_PASSES = 4
MAX_PLANES = 32
LAYOUT_NHWC, LAYOUT_NCHW_COLLAPSE, LAYOUT_NCDHW, LAYOUT_NCDHW_CAT, LAYOUT_NHWC_BF16 = 0, 1, 2, 3, 4


class MghsCfg(ctypes.Structure):
    """struct dhd_mghs_cfg (include/dhd_b200.h)."""
    _fields_ = [
        ('B', ctypes.c_int32), ('N', ctypes.c_int32), ('D', ctypes.c_int32),
        ('fH', ctypes.c_int32), ('fW', ctypes.c_int32), ('C', ctypes.c_int32),
        ('Dx', ctypes.c_int32), ('Dy', ctypes.c_int32),
        ('x_lower', ctypes.c_float), ('x_interval', ctypes.c_float), ('x_size', ctypes.c_float),
        ('y_lower', ctypes.c_float), ('y_interval', ctypes.c_float), ('y_size', ctypes.c_float),
        ('n_pass', ctypes.c_int32),
        ('z_lower', ctypes.c_float * MAX_PASSES),
        ('z_interval', ctypes.c_float * MAX_PASSES),
        ('z_size', ctypes.c_float * MAX_PASSES),
        ('dz', ctypes.c_int32 * MAX_PASSES),
        ('mask_id', ctypes.c_int32 * MAX_PASSES),
    ]


_P = ctypes.c_void_p
_I = ctypes.c_int
_SIGNATURES = {
    'dhd_last_error': (ctypes.c_char_p, []),
    'dhd_abi_version': (ctypes.c_int, []),
    'dhd_abi_sizeof': (ctypes.c_size_t, [ctypes.c_int]),
    'dhd_bev_pool_v2_fwd': (ctypes.c_int, [ctypes.c_int, ctypes.c_int] + [_P] * 9),
    'dhd_bev_pool_v2_bwd': (ctypes.c_int, [ctypes.c_int, ctypes.c_int] + [_P] * 11),
    'dhd_height_to_mask': (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P,
                                          ctypes.c_int, _P, _P]),
    'dhd_mghs_workspace_bytes': (ctypes.c_size_t, [ctypes.POINTER(MghsCfg)]),
    'dhd_mghs_workspace_count_offset': (ctypes.c_size_t, [ctypes.POINTER(MghsCfg)]),
    'dhd_mghs_prepare': (ctypes.c_int, [ctypes.POINTER(MghsCfg)] + [_P] * 10 + [ctypes.c_int, _P]),
    'dhd_mghs_pool_fwd': (ctypes.c_int, [ctypes.POINTER(MghsCfg), _P, _P, _P, _P,
                                         ctypes.POINTER(_P), ctypes.c_int, _P]),
    'dhd_mghs_pool_bwd': (ctypes.c_int, [ctypes.POINTER(MghsCfg), _P, _P, _P, _P,
                                         ctypes.POINTER(_P), ctypes.c_int, _P, _P, _P]),
    'dhd_mghs_voxel_index': (ctypes.c_int, [ctypes.POINTER(MghsCfg), _P, _P, _P]),
    'dhd_conv2d_fwd': (ctypes.c_int, [_P, _P]),
    'dhd_conv2d_fwd_batch': (ctypes.c_int, [_P, _I, _P]),
    'dhd_conv_pair_mode': (ctypes.c_int, [_I]),
    'dhd_conv2d_stat_rows': (ctypes.c_int, [_P]),
    'dhd_stem_im2col': (ctypes.c_int, [_P] + [_I] * 7 + [_P, _I, _I, _I, _P]),
    'dhd_maxpool3s2': (ctypes.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P]),
    'dhd_upsample_nearest_add': (ctypes.c_int, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'dhd_pack_conv_weights_batch': (ctypes.c_int, [_P, _I, _P]),
    'dhd_bn_fwd_coeffs_partial': (ctypes.c_int, [_P, _I, _I, ctypes.c_float, _P, _P, ctypes.c_float, ctypes.c_float] + [_P] * 7),
    'dhd_bn_bwd_sums_coeffs': (ctypes.c_int, [_P, _I, _I, _P, _I, _I, ctypes.c_long, _I, _P, ctypes.c_float] + [_P] * 9),
    'dhd_adamw_flat': (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_long] + [ctypes.c_float] * 7 + [_P, _P]),
    'dhd_colsum_finish': (ctypes.c_int, [_P, _I, _I, _P, _P]),
    'dhd_conv2d_wgrad_workspace_bytes': (ctypes.c_size_t, [_P]),
    'dhd_conv2d_wgrad': (ctypes.c_int, [_P, _P]),
    'dhd_act_bwd_workspace_bytes': (ctypes.c_size_t, [_I]),
    'dhd_act_bwd': (ctypes.c_int, [_P, _I, _I, _P, _I, _I, ctypes.c_long, _I, _I, _P, _I, _I, _P, _P, _P, _I, _I, _P]),
    'dhd_pack_conv_weights': (ctypes.c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _I, _P]),
    'dhd_bn_apply': (ctypes.c_int, [_P, _I, _I, ctypes.c_long, _I, _P, _P, _I, _P, ctypes.c_long, _P, _I, _P, _I, _I, _P, ctypes.c_long, _P]),
    'dhd_bn_fwd_coeffs': (ctypes.c_int, [_P, _I, ctypes.c_float, _P, _P, ctypes.c_float, ctypes.c_float, _P, _P, _P, _P, _P, _P, _P]),
    'dhd_bn_bwd_coeffs': (ctypes.c_int, [_P, _I, _I, ctypes.c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'dhd_dropout': (ctypes.c_int, [_P, _I, _I, ctypes.c_long, _I, ctypes.c_float, _P, ctypes.c_uint, _P]),
    'dhd_affine_combine': (ctypes.c_int, [_P, _I, _I, _P, _I, _I, ctypes.c_long, _I, _P, _P, _P, _P, _I, _I, _P]),
    'dhd_se_gate_bwd': (ctypes.c_int, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P]),
    'dhd_gt_downsample': (ctypes.c_int, [_P, _I, _I, _I, _I, ctypes.c_float, ctypes.c_float, _I, _P, _P, _P]),
    'dhd_height_loss': (ctypes.c_int, [_P, _P, _P, _I, _I, _I, ctypes.c_float, _P, _P, _P, _I, _P]),
    'dhd_dcn_col2im_bwd': (ctypes.c_int, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    'dhd_occ_loss_workspace_bytes': (ctypes.c_size_t, []),
    'dhd_occ_ce_loss': (ctypes.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, ctypes.c_float, ctypes.c_float, ctypes.c_float, _I,
                                       _P, _P, _I, _P, _P]),
    'dhd_depth_head_bwd': (ctypes.c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    'dhd_sfa_gate_bwd_workspace_bytes': (ctypes.c_size_t, [_I, _I, _I]),
    'dhd_sfa_gate_bwd': (ctypes.c_int, [_I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P, _P, _I, _P, _P]),
    'dhd_sfa_gate_bwd_b16': (ctypes.c_int, [_I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P, _I, _I, _P, _I, _P, _P]),
    'dhd_sfa_dx_combine': (ctypes.c_int, [_P, _I, _I, _P, _I, _I, _P, _P, _I, _I, _I, _P, _I, _I, _P]),
    'dhd_add_rowvec': (ctypes.c_int, [_P, _P, _I, _I, _I, _P, _I, _I, _P]),
    'dhd_pack_nchw_to_nhwc': (ctypes.c_int, [_P] + [_I] * 4 + [_P] + [_I] * 4 + [_P]),
    'dhd_occ_argmax': (ctypes.c_int, [_P, ctypes.c_long, _I, _P, _P]),
    'dhd_predictor_tail': (ctypes.c_int, [_P, _P]),
    'dhd_launch_count': (ctypes.c_long, []),
    'dhd_maxpool3s2_bwd': (ctypes.c_int, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    'dhd_bn_apply_res16': (ctypes.c_int, [_P, _I, _I, ctypes.c_long, _I, _P, _P, _I, _P, _I, _I, _P, _I, _I, _P]),
    'dhd_upsample_bilinear_bwd_gather': (ctypes.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P, _P]),
    'dhd_maxpool2_bwd': (ctypes.c_int, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    'dhd_upsample_bilinear_bwd': (ctypes.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'dhd_maxpool2': (ctypes.c_int, [_P] + [_I] * 7 + [_P, _I, _I, _I, _I, _P]),
    'dhd_upsample_bilinear': (ctypes.c_int, [_P] + [_I] * 9 + [_P, _I, _I, _I, _I, _P]),
    'dhd_split_nhwc': (ctypes.c_int, [_P, ctypes.c_long, _I, _P] + [_I] * 4 + [_P]),
    'dhd_split_nhwc_mean': (ctypes.c_int, [_P, _I, _I, _I, _P] + [_I] * 4 + [_P, _P, _P]),
    'dhd_mean_workspace_bytes': (ctypes.c_size_t, [_I, _I, _I]),
    'dhd_unpack_nhwc_to_nchw': (ctypes.c_int, [_P] + [_I] * 8 + [_P, _P]),
    'dhd_mean_hw': (ctypes.c_int, [_P] + [_I] * 7 + [_P, _P, _P]),
    'dhd_linear_rows': (ctypes.c_int, [_P, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P, _P]),
    'dhd_gate_channels': (ctypes.c_int, [_P, _I, _I, _I, _P, _P] + [_I] * 4 + [_P, _P]),
    'dhd_sfa_mix': (ctypes.c_int, [_P] + [_I] * 7 + [_P, _P, _P] + [_I] * 4 + [_P]),
    'dhd_sfa_blend_b16': (ctypes.c_int, [_P] + [_I] * 5 + [_P, _P, _I, _I, _P, _I, _I, _P]),
    'dhd_sfa_fold_gate': (ctypes.c_int, [_P, _P, _I, _I, _I, _P, _P]),
    'dhd_stereo_cost_volume': (ctypes.c_int, [_P, _P]),
    'dhd_nchw_to_nhwc': (ctypes.c_int, [_P, _I, _I, _I, _P, _I, _P]),
    'dhd_dcn_im2col': (ctypes.c_int, [_P] + [_I] * 8 + [_P] + [_I] * 5 + [_P] + [_I] * 3 + [_P]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'dhd_b200: %s is missing -- run `python -m dhd_b200.build` (or '
            '__graft_entry__.build()). There is no CPU / PyTorch fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.dhd_abi_version() != 1:
        raise RuntimeError('dhd_b200: ABI version mismatch, rebuild the library')
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dhd_last_error().decode('utf-8', 'replace')
        raise RuntimeError('dhd_b200.%s failed (code %d): %s' % (what, rc, msg))
