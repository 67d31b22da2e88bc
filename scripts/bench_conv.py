"""Times the tcgen05 convolution kernel on the DHD-S layer shapes (CUDA events, after warm-up).
Usage: python scripts/bench_conv.py [precision ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dhd_b200 import dense as D  # noqa: E402

SHAPES = [  # name, N, Cin, Cout, H, W, ksize, dilation
    ('heightnet 3x3 256->256 @24x16x44', 24, 256, 256, 16, 44, 3, 1),
    ('aspp 3x3 d6 256->256 @24x16x44', 24, 256, 256, 16, 44, 3, 6),
    ('aspp 1x1 1280->256 @24x16x44', 24, 1280, 256, 16, 44, 1, 1),
    ('depth_net 1x1 256->108 @24x16x44', 24, 256, 108, 16, 44, 1, 1),
    ('sfa 3x3 256->256 @4x200x200', 4, 256, 256, 200, 200, 3, 1),
    ('sfa 1x1 512->256 @4x200x200', 4, 512, 256, 200, 200, 1, 1),
    ('predicter 1x1 256->512 @4x200x200', 4, 256, 512, 200, 200, 1, 1),
    ('predicter 1x1 512->288 @4x200x200', 4, 512, 288, 200, 200, 1, 1),
]


def wgrad():
    """weight-gradient kernel (dhd_conv2d_wgrad) on the same shapes"""
    for name, N, Cin, Cout, H, W, k, dil in SHAPES:
        x = D.Act.empty(N, H, W, Cin, 1, 'cuda')
        x.data.normal_()
        dy = D.Act.empty(N, H, W, (Cout + 63) // 64 * 64, 1, 'cuda')
        dy.data.normal_()
        out = torch.empty(Cout, k * k, Cin, device='cuda')
        run = lambda: D.conv2d_wgrad(x, dy, Cout, ksize=k, dilation=dil, out=out)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 20
        e0.record()
        for _ in range(it):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        flops = 2.0 * N * H * W * Cout * Cin * k * k
        print('%-40s %-7s %8.1f us  %7.1f TFLOP/s algorithmic' % (name, 'wgrad', ms * 1e3, flops / ms / 1e9), flush=True)


def pair_ab():
    """CTA pairs (tcgen05.mma.cta_group::2) against one-CTA tiles, per layer shape: 20 launches captured in a CUDA
    graph (no host time in the figure), replayed once (burst) and back to back for ~1 s (sustained clocks)."""
    from dhd_b200 import _lib
    lib = _lib.load()
    for name, N, Cin, Cout, H, W, k, dil in SHAPES + [('predictor 3x3 256->256 @4x200x200 (2nd)', 4, 256, 256, 200, 200, 3, 1)]:
        x = D.Act.empty(N, H, W, Cin, 1, 'cuda')
        x.data.normal_()
        w = torch.randn(Cout, k * k, 1, Cin, device='cuda').to(torch.bfloat16)
        out = D.Act.empty(N, H, W, (Cout + 63) // 64 * 64, 1, 'cuda')
        res = {}
        for mode in (0, 2):
            lib.dhd_conv_pair_mode(mode)
            run = lambda: D.conv2d(x, w, Cout, ksize=k, dilation=dil, precision='bf16', segs=[dict(act='relu', out_act=out)])
            run()
            torch.cuda.synchronize()
            it = 20
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(it):
                    run()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            burst = e0.elapsed_time(e1) / it
            reps = max(1, int(1000.0 / (burst * it)))
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            res[mode] = (burst * 1e3, e0.elapsed_time(e1) / (it * reps) * 1e3)
        lib.dhd_conv_pair_mode(1)
        flops = 2.0 * N * H * W * Cout * Cin * k * k
        print('%-42s one-CTA %7.1f us burst %7.1f us sustained (%6.1f TF/s) | pairs %7.1f us burst %7.1f us sustained (%6.1f TF/s)' %
              (name, res[0][0], res[0][1], flops / res[0][1] / 1e6, res[2][0], res[2][1], flops / res[2][1] / 1e6), flush=True)


def main():
    if sys.argv[1:] == ['wgrad']:
        return wgrad()
    if sys.argv[1:] == ['pair']:
        return pair_ab()
    precs = sys.argv[1:] or ['bf16', 'bf16x3', 'fp32']
    for name, N, Cin, Cout, H, W, k, dil in SHAPES:
        for prec in precs:
            parts, terms = D.PRECISIONS[prec]
            x = D.Act.empty(N, H, W, Cin, parts, 'cuda')
            x.data.normal_()
            w = torch.randn(Cout, k * k, parts, Cin, device='cuda').to(torch.bfloat16)
            out = D.Act.empty(N, H, W, Cout, parts, 'cuda')
            run = lambda: D.conv2d(x, w, Cout, ksize=k, dilation=dil, precision=prec,
                                   segs=[dict(act='relu', out_act=out)])
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            it = 20
            e0.record()
            for _ in range(it):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / it
            flops = 2.0 * N * H * W * Cout * Cin * k * k
            print('%-40s %-7s %8.1f us  %7.1f TFLOP/s algorithmic  %7.1f TFLOP/s issued' %
                  (name, prec, ms * 1e3, flops / ms / 1e9, flops * len(terms) / ms / 1e9), flush=True)


if __name__ == '__main__':
    main()
