"""dhd_maxpool3s2_bwd at the DHD-S training size (24 x 128x352x64 stem output), CUDA events; DHD_MAXPOOL_BWD=ref times the
first kernel.  Usage: python scripts/bench_maxpool_bwd.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200 import _lib  # noqa: E402
from dhd_b200 import dense as D  # noqa: E402
from dhd_b200.modules import _p, _stream  # noqa: E402

lib = _lib.load()
N, H, W, C = 24, 128, 352, 64
g = torch.Generator(device='cuda').manual_seed(0)
x = D.Act(torch.relu(torch.randn(N, H, W, C, device='cuda', generator=g)).bfloat16(), C, 1)
dy = D.Act(torch.randn(N, H // 2, W // 2, C, device='cuda', generator=g).bfloat16(), C, 1)
dx = D.Act.empty(N, H, W, C, 1, 'cuda')
run = lambda: _lib.check(lib.dhd_maxpool3s2_bwd(_p(x.data), x.ld, x.coff, _p(dy.data), dy.ld, dy.coff, N, H, W, C, _p(dx.data),
                                                dx.ld, dx.coff, 1, _stream()), 'maxpool3s2_bwd')
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
xr = x.data.float().permute(0, 3, 1, 2).requires_grad_()
torch.nn.functional.max_pool2d(xr, 3, 2, 1).backward(dy.data.float().permute(0, 3, 1, 2))
want = (xr.grad * (xr.detach() > 0)).permute(0, 2, 3, 1).bfloat16()
print(json.dumps({'kernel': os.environ.get('DHD_MAXPOOL_BWD', 'packed'), 'us': 1e3 * e0.elapsed_time(e1) / 20,
                  'equals_torch_at_full_size': bool(torch.equal(dx.data, want))}))
