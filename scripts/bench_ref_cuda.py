"""MGHS.view_transform on one B200: the reference's CUDA path (oracle/ref_cuda_path.py: the
reference's own op sequence in torch CUDA ops + its unmodified bev_pool_v2 kernel from
oracle/_ref) against the fused path (height_to_mask + dhd_mghs_prepare + dhd_mghs_pool_fwd),
DHD-S, B=4, same inputs; checks the four outputs agree first.  Prints one JSON line.
Usage: python scripts/bench_ref_cuda.py [B] > profiles/r01_ref_cuda_vs_fused.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200.pool import MghsPool, height_to_mask  # noqa: E402
from oracle import mghs_oracle as O  # noqa: E402
from oracle import ref_cuda_path as R  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return sum(ts) / len(ts), ts[0]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    cfg = O.DHD_S
    inputs, depth, feat, height = O.synthetic_inputs(cfg, B, seed=3)
    inputs = tuple(t.cuda() for t in inputs)
    depth, feat, height = depth.cuda(), feat.cuda(), height.cuda()
    N, D = cfg['ncams'], depth.shape[1]
    fH, fW = depth.shape[-2:]
    C = cfg['C']
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample']).cuda()

    def ref():
        return R.view_transform_cuda(inputs, depth, feat, height, fr, cfg['height_range'], cfg['mask_range'],
                                     cfg['mask_grids'])

    grids = [cfg['bev_grid']] + list(cfg['mask_grids'])
    plan = MghsPool(B, N, D, fH, fW, C, grids[0]['x'], grids[0]['y'], [(g['z'], m) for m, g in enumerate(grids)])
    ws = torch.empty(plan.ws_bytes, dtype=torch.uint8, device='cuda')
    outs = plan.alloc_outputs('nhwc', 'cuda')
    f_nhwc = feat.view(B, N, C, fH, fW).permute(0, 1, 3, 4, 2).contiguous()
    _, s2e, _e2g, K, pr, pt, bda = inputs

    def ours():
        pm = height_to_mask(height, cfg['height_range'], cfg['mask_range'])
        plan.prepare(frustum=fr, sensor2ego=s2e, cam2imgs=K, post_rots=pr, post_trans=pt, bda=bda, workspace=ws)
        plan.raw_forward(depth, f_nhwc, pm, outs, 'nhwc', workspace=ws)

    def ours_pool_only():
        plan.raw_forward(depth, f_nhwc, height_to_mask(height, cfg['height_range'], cfg['mask_range']), outs,
                         'nhwc', workspace=ws)

    want = ref()
    # same geometry for the parity check: bins from the reference's own coordinates
    coor = O.ego_coor(fr, s2e, K, pr, pt, bda)
    pm = height_to_mask(height, cfg['height_range'], cfg['mask_range'])
    plan.prepare(coor=coor, workspace=ws)
    plan.raw_forward(depth, f_nhwc, pm, outs, 'nhwc', workspace=ws)
    torch.cuda.synchronize()
    max_err = 0.0
    for o, w in zip(outs, want):
        got = o.permute(0, 3, 1, 2)
        assert got.shape == w.shape
        assert torch.allclose(got, w, rtol=1e-5, atol=2e-6), 'fused path differs from the reference CUDA path'
        max_err = max(max_err, float((got - w).abs().max()))
    del want
    ref_avg, ref_min = timed(ref, n=10)
    our_avg, our_min = timed(ours)
    pool_avg, pool_min = timed(ours_pool_only)
    print(json.dumps({
        'what': 'MGHS.view_transform (4 passes) DHD-S B=%d on 1xB200: reference CUDA path vs fused path' % B,
        'reference_cuda_ms': ref_avg, 'reference_cuda_ms_min': ref_min,
        'fused_ms (height_to_mask + prepare from raw camera tensors + pool)': our_avg, 'fused_ms_min': our_min,
        'fused_cached_bins_ms (MGHS accelerate=True: pool only)': pool_avg,
        'speedup': ref_avg / our_avg, 'speedup_cached_bins': ref_avg / pool_avg,
        'samples_per_s_reference': B / ref_avg * 1e3, 'samples_per_s_fused': B / our_avg * 1e3,
        'parity': 'all four outputs allclose(rtol=1e-5, atol=2e-6) at full size; max abs diff %.3g' % max_err,
        'reference': 'oracle/ref_cuda_path.py: reference op sequence (LH:179-231, 303-371, 407-459; BP/bev_pool.py) in torch '
                     'CUDA ops + the unmodified reference kernel (oracle/_ref/libbev_pool_v2_ref.so), fp32',
    }))


if __name__ == '__main__':
    main()
