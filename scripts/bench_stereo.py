"""Times the fused plane-sweep cost-volume kernel (dhd_b200/csrc/stereo.cu) at DHD-L size -- B=2 samples x 6 cameras,
128 stereo channels (Swin-B stage 0), 1/4 map 128x352, D = 88 hypotheses -- next to the reference's own op sequence
(gen_grid + C/4 grid_sample groups + abs/sum + softmax, depthnet.py:245-361) run as torch CUDA ops through the oracle
restatement.  CUDA events after warm-up; prints one JSON line.
Usage: python scripts/bench_stereo.py [--no-ref] [--bn 12]"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dhd_b200 import stereo as S  # noqa: E402


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
    BN = int(sys.argv[sys.argv.index('--bn') + 1]) if '--bn' in sys.argv else 12
    C, H, W, D = 128, 128, 352, 88
    torch.manual_seed(0)
    k = torch.ones(1, 1, 5, 5, device='cuda') / 25.0
    feat = lambda: torch.nn.functional.conv2d(torch.randn(BN * C, 1, H, W, device='cuda'), k, padding=2).view(BN, C, H, W)
    prev, curr = feat(), feat()
    d = torch.arange(1.0, 45.0, 0.5).view(-1, 1, 1).expand(-1, H, W)
    u = torch.linspace(0, 4 * W - 1, W).view(1, 1, W).expand(D, H, W)
    v = torch.linspace(0, 4 * H - 1, H).view(1, H, 1).expand(D, H, W)
    frustum = torch.stack((u, v, d), -1).cuda()
    s = 4 * W / 1600.0
    B = BN // 6 if BN % 6 == 0 else 1
    N = BN // B
    intr = torch.tensor([[1266.0, 0.0, 816.0], [0.0, 1266.0, 491.0], [0.0, 0.0, 1.0]]).expand(B, N, 3, 3).contiguous().cuda()
    post_rots = torch.diag(torch.tensor([s, s, 1.0])).expand(B, N, 3, 3).contiguous().cuda()
    post_trans = torch.tensor([0.0, -140.0 * s, 0.0]).expand(B, N, 3).contiguous().cuda()
    a = math.radians(1.5)
    k2s = torch.eye(4)
    k2s[:3, :3] = torch.tensor([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]])
    k2s[:3, 3] = torch.tensor([0.05, 0.0, 0.8])
    k2s = k2s.expand(B, N, 4, 4).contiguous().cuda()
    cam = S.camera_table(k2s, intr, post_rots, post_trans)
    out = torch.empty(BN, D, H, W, device='cuda')
    res = {'workload': 'DHD-L stereo cost volume: BN=%d, C=%d, %dx%d map, D=%d, bias 5' % (BN, C, H, W, D)}
    res['to_nhwc_fp32_ms'] = timed(lambda: S.to_nhwc(curr))
    res['to_nhwc_bf16_ms'] = timed(lambda: S.to_nhwc(curr, bf16=True))
    for name, bf in (('fp32', False), ('bf16', True)):
        p, c = S.to_nhwc(prev, bf16=bf), S.to_nhwc(curr, bf16=bf)
        ms = timed(lambda: S.cost_volume(p, c, D, (4 * H, 4 * W), bias=5.0, frustum=frustum, cam=cam, out=out))
        res['kernel_%s_ms' % name] = ms
        esz = 2 if bf else 4
        taps = BN * D * H * W * 4.0 * C * esz                       # bytes the bilinear taps request (L1 / L2 side)
        algo = 2.0 * BN * H * W * C * esz + BN * D * H * W * 4.0    # each feature once + the result once (HBM side)
        res['kernel_%s_tap_TBps' % name] = taps / ms / 1e9
        res['kernel_%s_algorithmic_GBps' % name] = algo / ms / 1e6
    if '--no-ref' not in sys.argv:
        from oracle import dense_oracle as DO                       # the reference's op sequence, as torch CUDA ops
        def ref():
            g = DO.stereo_sampling_grid(frustum, k2s, intr, post_rots, post_trans, 4 * H, 4 * W)
            return DO.stereo_cost_volume(prev, curr, g, D, 5.0)
        want = ref()
        got, _ = S.cost_volume(S.to_nhwc(prev), S.to_nhwc(curr), D, (4 * H, 4 * W), bias=5.0, frustum=frustum, cam=cam)
        res['max_abs_diff_vs_reference_ops'] = float((got - want).abs().max())
        res['max_rel_diff_vs_reference_ops'] = float(((got - want).abs() / want.clamp_min(1e-6)).max())
        res['reference_ops_cuda_ms'] = timed(ref, it=3, warm=1)
        res['speedup_fp32_incl_layout'] = res['reference_ops_cuda_ms'] / (res['kernel_fp32_ms'] + 2 * res['to_nhwc_fp32_ms'])
        res['speedup_bf16_incl_layout'] = res['reference_ops_cuda_ms'] / (res['kernel_bf16_ms'] + 2 * res['to_nhwc_bf16_ms'])
    print(json.dumps(res))


if __name__ == '__main__':
    main()
