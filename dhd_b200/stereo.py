"""Host side of the plane-sweep stereo cost volume (csrc/stereo.cu) -- the `calculate_cost_volumn` branch of
the camera-aware DepthNet (reference models/model_utils/depthnet.py:245-361) that DHD-M / DHD-L run per step.

ctypes over the C-ABI; torch owns the device memory and the stream.  No CPU / torch fallback: CPU tensors raise."""
import ctypes

import torch

from . import _lib

CAM_FLOATS = 40


class StereoDesc(ctypes.Structure):
    """struct dhd_stereo_desc (include/dhd_b200.h)."""
    _fields_ = [
        ('BN', ctypes.c_int32), ('C', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32),
        ('D', ctypes.c_int32), ('feat_bf16', ctypes.c_int32),
        ('prev', ctypes.c_void_p), ('curr', ctypes.c_void_p),
        ('frustum_u', ctypes.c_void_p), ('frustum_v', ctypes.c_void_p), ('frustum_d', ctypes.c_void_p),
        ('cam', ctypes.c_void_p), ('grid', ctypes.c_void_p),
        ('img_w', ctypes.c_float), ('img_h', ctypes.c_float), ('bias', ctypes.c_float),
        ('out_f32', ctypes.c_void_p),
        ('f32_sN', ctypes.c_int64), ('f32_sD', ctypes.c_int64), ('f32_sY', ctypes.c_int64), ('f32_sX', ctypes.c_int64),
        ('out_b16', ctypes.c_void_p),
        ('b16_ld', ctypes.c_int32), ('b16_coff', ctypes.c_int32), ('b16_parts', ctypes.c_int32),
        ('b16_part_stride', ctypes.c_int32), ('b16_cpad', ctypes.c_int32),
        ('grid_out', ctypes.c_void_p),
    ]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('dhd_b200.stereo: CUDA tensors required (there is no CPU path)')


def camera_table(k2s_sensor, intrins, post_rots, post_trans):
    """(B, N, ...) camera tensors of stereo_metas -> (B*N, 40) fp32 table of dhd_stereo_desc.cam.  The 3x3 inverses
    and the one 3x3 product are the reference's own torch calls (depthnet.py:273-284), so they round as there."""
    _need_cuda(k2s_sensor, intrins, post_rots, post_trans)
    B, N = post_trans.shape[:2]
    BN = B * N
    t = torch.zeros(BN, CAM_FLOATS, device=post_trans.device)
    t[:, 0:9] = torch.linalg.inv_ex(post_rots.float()).inverse.reshape(BN, 9)
    t[:, 9:12] = post_trans.float().reshape(BN, 3)
    rots = k2s_sensor[:, :, :3, :3].float().contiguous()
    t[:, 12:21] = rots.matmul(torch.linalg.inv_ex(intrins.float()).inverse).reshape(BN, 9)
    t[:, 21:24] = k2s_sensor[:, :, :3, 3].float().reshape(BN, 3)
    t[:, 24:33] = intrins.float().reshape(BN, 9)
    t[:, 33:37] = post_rots[..., :2, :2].float().reshape(BN, 4)
    t[:, 37:39] = post_trans[..., :2].float().reshape(BN, 2)
    return t


def frustum_axes(frustum):
    """(D, H, W, 3) template of create_frustum (lss_heightmap.py:105-134) -> its three axes (u (W), v (H), d (D)):
    the template is frustum[d][y][x] = (u[x], v[y], d[d]) by construction, so the kernel reads 3 small vectors instead
    of 12 bytes per point."""
    f = frustum.float()
    return f[0, 0, :, 0].contiguous(), f[0, :, 0, 1].contiguous(), f[:, 0, 0, 2].contiguous()


def to_nhwc(x, bf16=False):
    """(N, C, H, W) fp32 contiguous -> (N, H, W, C) fp32 or bf16: the layout the kernel gathers from (every
    bilinear tap is one contiguous row of C channels)."""
    _need_cuda(x)
    if x.dim() != 4 or x.dtype != torch.float32:
        raise ValueError('expected a (N, C, H, W) fp32 tensor')
    x = x.contiguous()
    N, C, H, W = x.shape
    out = torch.empty(N, H, W, C, device=x.device, dtype=torch.bfloat16 if bf16 else torch.float32)
    _lib.check(_lib.load().dhd_nchw_to_nhwc(ctypes.c_void_p(x.data_ptr()), N, C, H * W,
                                            ctypes.c_void_p(out.data_ptr()), int(bf16), _stream()), 'nchw_to_nhwc')
    return out


def cost_volume(prev, curr, depth_bins, img_hw, bias=0.0, frustum=None, cam=None, grid=None, out=None, out_act=None,
                want_grid=False):
    """softmax(-L1 matching cost) over the depth hypotheses.

    prev, curr : (BN, H, W, C) NHWC feature maps, both fp32 or both bf16 (see to_nhwc)
    frustum    : (D, H, W, 3) fp32 (u, v, d) template at the stereo resolution, or frustum_axes() of it (cacheable),
                 + cam = camera_table(...): the sampling
                 coordinates are computed in the kernel; or grid (BN, D*H, W, 2), coordinates computed elsewhere
    out        : optional (BN, D, H, W) fp32 tensor written in the reference's NCHW layout (allocated when neither
                 `out` nor `out_act` is given)
    out_act    : optional dense.Act (BN, H, W, >= D channels): the split-bf16 NHWC input of cost_volumn_net
    Returns (out, grid_used or None)."""
    _need_cuda(prev, curr, cam, grid, *(frustum if isinstance(frustum, (tuple, list)) else (frustum,)))
    if prev.shape != curr.shape or prev.dtype != curr.dtype or prev.dim() != 4:
        raise ValueError('prev / curr must be NHWC tensors of one shape and dtype')
    if prev.dtype not in (torch.float32, torch.bfloat16) or not prev.is_contiguous() or not curr.is_contiguous():
        raise ValueError('features must be contiguous fp32 or bf16')
    BN, H, W, C = curr.shape
    D = int(depth_bins)
    d = StereoDesc()
    d.BN, d.C, d.H, d.W, d.D = BN, C, H, W, D
    d.feat_bf16 = int(prev.dtype == torch.bfloat16)
    d.prev, d.curr = prev.data_ptr(), curr.data_ptr()
    keep = [prev, curr]
    if grid is not None:
        if tuple(grid.shape) != (BN, D * H, W, 2) or grid.dtype != torch.float32 or not grid.is_contiguous():
            raise ValueError('grid must be a contiguous fp32 (BN, D*H, W, 2) tensor')
        d.grid = grid.data_ptr()
        keep.append(grid)
    else:
        if frustum is None or cam is None:
            raise ValueError('either grid or frustum + cam must be given')
        fu, fv, fd = frustum if isinstance(frustum, (tuple, list)) else frustum_axes(frustum)
        if tuple(fu.shape) != (W,) or tuple(fv.shape) != (H,) or tuple(fd.shape) != (D,) or \
                tuple(cam.shape) != (BN, CAM_FLOATS):
            raise ValueError('frustum must be (D, H, W, 3) [or its (u, v, d) axes] and cam (BN, %d)' % CAM_FLOATS)
        cam = cam.float().contiguous()
        d.frustum_u, d.frustum_v, d.frustum_d, d.cam = fu.data_ptr(), fv.data_ptr(), fd.data_ptr(), cam.data_ptr()
        keep += [fu, fv, fd, cam]
    d.img_h, d.img_w = float(img_hw[0]), float(img_hw[1])
    d.bias = float(bias)
    if out is None and out_act is None:
        out = torch.empty(BN, D, H, W, device=curr.device)
    if out is not None:
        if tuple(out.shape) != (BN, D, H, W) or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError('out must be a (BN, D, H, W) fp32 CUDA tensor')
        d.out_f32 = out.data_ptr()
        d.f32_sN, d.f32_sD, d.f32_sY, d.f32_sX = out.stride()
    if out_act is not None:
        a = out_act
        if (a.N, a.H, a.W) != (BN, H, W) or a.C < D:
            raise ValueError('out_act does not match the cost volume')
        d.out_b16 = a.data.data_ptr()
        d.b16_ld, d.b16_coff, d.b16_parts, d.b16_part_stride, d.b16_cpad = a.ld, a.coff, a.parts, a.part_stride, a.C
        keep.append(a.data)
    g = None
    if want_grid:
        g = torch.empty(BN, D * H, W, 2, device=curr.device)
        d.grid_out = g.data_ptr()
    _lib.check(_lib.load().dhd_stereo_cost_volume(ctypes.byref(d), _stream()), 'stereo_cost_volume')
    return out, g
