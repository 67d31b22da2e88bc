set -x
python -m pytest tests/test_dense_gpu.py -x -q 2>&1 | tail -30
python -m pytest tests -m gpu -x -q --deselect tests/test_dense_gpu.py 2>&1 | tail -5
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r01_bench_poolv3.json; cat gpurun_out/r01_bench_poolv3.json
python - <<'P'
import torch
x = torch.empty(704*1024*1024//4, device='cuda')
for f in (lambda: x.zero_(), lambda: x.fill_(1.0)):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print('memset-like write of %.0f MB: %.1f us = %.0f GB/s' % (x.numel()*4/1e6, ms*1e3, x.numel()*4/ms/1e6))
P
