"""BASELINE configs[4] (DHD-L, 6-cam 512x1408, 2-frame stereo) through the plugin's MGHS_Stereo built with the kwargs of
projects/configs/DHD/DHD-L.py: image features (B,6,512,32,88) + stereo features (B*6,128,128,352) of two frames ->
plane-sweep cost volume -> stereo DepthNet + HeightNet -> fused voxel pool -> (B,C,1,200,200) + (B,C,16,200,200).
CUDA events after warm-up; prints one JSON line.   Usage: python scripts/bench_dhdl.py [bf16|fp32] [B]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dhd_b200 import synth  # noqa: E402


def build(precision, B):
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.necks.lss_heightmap import MGHS_Stereo
    kw = dict(synth.DHD_L_VIEW_TRANSFORMER)
    torch.manual_seed(0)
    vt = MGHS_Stereo(precision=precision, **kw).eval().cuda()
    H, W = kw['input_size']
    rig = synth.synthetic_rig(B, 6, kw['input_size'], seed=1)
    s2e, e2g, K, pr, pt, bda = [t.cuda() for t in rig]
    x = torch.randn(B, 6, kw['in_channels'], H // 16, W // 16, device='cuda')
    mlp = vt.get_mlp_input(s2e, e2g, K, pr, pt, bda)
    k = torch.ones(1, 1, 5, 5, device='cuda') / 25.0
    C = synth.DHD_L_STEREO_CHANNELS
    feat = lambda: torch.nn.functional.conv2d(torch.randn(B * 6 * C, 1, H // 4, W // 4, device='cuda'), k,
                                              padding=2).view(B * 6, C, H // 4, W // 4)
    metas = dict(k2s_sensor=synth.synthetic_k2s_sensor(s2e), intrins=K, post_rots=pr, post_trans=pt,
                 frustum=vt.cv_frustum.cuda(), cv_downsample=4, downsample=vt.downsample, grid_config=vt.grid_config,
                 cv_feat_list=[feat(), feat()])
    return vt, [x, s2e, e2g, K, pr, pt, bda, mlp], metas


def timed(fn, it=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    vt, args, metas = build(precision, B)
    with torch.no_grad():
        bev, bev_z, depth, height = vt(args, metas)
        res = {'workload': 'BASELINE configs[4]: DHD-L view transformer (MGHS_Stereo), 6-cam 512x1408, D=88, C_in=512, '
                           'stereo C=128 @128x352, B=%d, %s' % (B, precision),
               'shapes': {'bev': list(bev.shape), 'bev_w_z': list(bev_z.shape), 'depth': list(depth.shape),
                          'height': list(height.shape)},
               'finite': bool(torch.isfinite(bev).all() and torch.isfinite(bev_z).all())}
        res['forward_ms'] = timed(lambda: vt(args, metas))
        res['samples_per_s'] = B / res['forward_ms'] * 1e3
        res['cost_volume_ms'] = timed(lambda: vt.depth_net.calculate_cost_volumn(metas))
        first = dict(metas, cv_feat_list=[None, metas['cv_feat_list'][1]])
        res['forward_first_frame_ms'] = timed(lambda: vt(args, first))
        vt.accelerate = True                      # bins cached across calls (MGHS.accelerate, LH:56)
        vt(args, metas)
        res['forward_cached_bins_ms'] = timed(lambda: vt(args, metas))
    print(json.dumps(res))


if __name__ == '__main__':
    main()
