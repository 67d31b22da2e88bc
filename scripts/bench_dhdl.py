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


build = synth.dhdl_view_transformer


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
