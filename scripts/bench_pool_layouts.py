"""dhd_mghs_pool_fwd per output layout and workload (CUDA events, 30 launches after warm-up): the NHWC stream kernel
(DHD-S / DHD-B encoders' layout), NCHW, and the (B, C, dz, Dy, Dx) layouts of collapse_z=False (DHD-M / DHD-L:
`ncdhw`, `ncdhw_cat`).  Bytes = every output byte once + depth + context + mask."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200 import synth as O  # noqa: E402
from dhd_b200.pool import MghsPool  # noqa: E402

PEAK = 6538.6


def run(name, B, input_size, depth_cfg, layouts):
    cfg = O.DHD_S
    N, ds, C = 6, 16, 64
    fH, fW = input_size[0] // ds, input_size[1] // ds
    D = int(round((depth_cfg[1] - depth_cfg[0]) / depth_cfg[2]))
    passes = [(cfg['bev_grid']['z'], 0)] + [(g['z'], i + 1) for i, g in enumerate(cfg['mask_grids'])]
    plan = MghsPool(B, N, D, fH, fW, C, cfg['bev_grid']['x'], cfg['bev_grid']['y'], passes)
    rig = [t.cuda() for t in O.synthetic_rig(B, N, input_size, seed=100)]
    s2e, e2g, K, pr, pt, bda = rig
    d = torch.arange(*depth_cfg, dtype=torch.float)                # MGHS.create_frustum (lss_heightmap.py:101-115)
    u = torch.linspace(0, input_size[1] - 1, fW, dtype=torch.float)
    v = torch.linspace(0, input_size[0] - 1, fH, dtype=torch.float)
    frustum = torch.empty(D, fH, fW, 3)
    frustum[..., 0], frustum[..., 1], frustum[..., 2] = u.view(1, 1, fW), v.view(1, fH, 1), d.view(-1, 1, 1)
    frustum = frustum.cuda()
    g = torch.Generator(device='cuda').manual_seed(1)
    depth = torch.rand(B * N, D, fH, fW, device='cuda', generator=g).softmax(1).contiguous()
    feat = torch.randn(B * N, fH, fW, C, device='cuda', generator=g)
    pixmask = torch.randint(1, 4, (B * N * fH * fW,), device='cuda', generator=g, dtype=torch.int8)
    plan.prepare(frustum=frustum, sensor2ego=s2e, cam2imgs=K, post_rots=pr, post_trans=pt, bda=bda)
    st = torch.cuda.current_stream()
    for layout in layouts:
        outs = plan.alloc_outputs(layout, 'cuda')
        nbytes = sum(o.numel() * o.element_size() for o in outs) + depth.numel() * 4 + feat.numel() * 4 + pixmask.numel()
        for _ in range(5):
            plan.raw_forward(depth, feat, pixmask, outs, layout)
        ts = []
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            plan.raw_forward(depth, feat, pixmask, outs, layout)
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        avg = sum(ts) / len(ts)
        print('%-34s %-10s %7.1f MB  avg %6.1f us  min %6.1f us  %5.0f GB/s = %.2f of the measured copy peak' %
              (name, layout, nbytes / 1e6, avg * 1e3, min(ts) * 1e3, nbytes / avg / 1e6, nbytes / avg / 1e6 / PEAK), flush=True)


def main():
    run('DHD-S B=4 (256x704, D=44)', 4, (256, 704), [1.0, 45.0, 1.0], ['nhwc', 'nhwc_bf16', 'nchw', 'ncdhw', 'ncdhw_cat'])
    run('DHD-L B=2 (512x1408, D=88)', 2, (512, 1408), [1.0, 45.0, 0.5], ['ncdhw_cat', 'ncdhw', 'nhwc'])


if __name__ == '__main__':
    os.environ['DHD_POOL_SWEEP'] = '1'          # the library re-reads its tunables on every call
    for pz in (sys.argv[1:] or ['9']):
        os.environ['DHD_POOL_PZ'] = pz
        print('DHD_POOL_PZ =', pz)
        main()
