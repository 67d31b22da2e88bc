"""Pool-forward micro-benchmark on the DHD-S B=4 workload: sweeps the env tunables of
dhd_mghs_pool_fwd (csrc/mghs_pool.cu) and checks every variant bit-for-bit against the v2 kernel.
Usage: python scripts/bench_pool.py ["V=3,ZT=4096,THREADS=256,CELLCOST=2,HINT=0,PERSM=8" ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200.pipeline import HotPathStep, algorithmic_bytes  # noqa: E402
from oracle import mghs_oracle as O  # noqa: E402


def main():
    variants = sys.argv[1:] or ['V=2', 'V=3']
    cfg, B = O.DHD_S, 4
    step = HotPathStep(cfg, B, precision='bf16', use_graph=False)
    rig = O.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=100)
    host = step.make_host_inputs(rig, seed=100)
    step.alloc_static(host)
    step.upload(host)
    step._front()
    torch.cuda.synchronize()
    alg = algorithmic_bytes(cfg, B)['pool_fwd_bytes']
    os.environ['DHD_POOL_V'] = '2'
    for o in step.outs:
        o.fill_(float('nan'))
    step._pool()
    torch.cuda.synchronize()
    want = [o.clone() for o in step.outs]
    st = torch.cuda.current_stream()
    for v in variants:
        for k in list(os.environ):
            if k.startswith('DHD_POOL_'):
                del os.environ[k]
        for kv in v.split(','):
            k, val = kv.split('=')
            os.environ['DHD_POOL_' + k] = val
        step._front()          # prepare again: the chunk table depends on NCH / CELLCOST
        for o in step.outs:
            o.fill_(float('nan'))
        step._pool()
        torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(step.outs, want))
        for _ in range(5):
            step._pool()
        ts = []
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            step._pool()
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        avg = sum(ts) / len(ts)
        print('%-60s bit-equal=%s  avg %.1f us  min %.1f us  %.0f GB/s avg (%.1f%% of 6538.6)' %
              (v, same, avg * 1e3, ts[0] * 1e3, alg / avg / 1e6, 100 * alg / avg / 1e6 / 6538.6), flush=True)


def probes():
    import ctypes
    from dhd_b200 import _lib
    lib = _lib.load()
    nbytes = 696320000 // 256 * 256
    buf = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream()
    combos = [(0, 0, 8), (0, 0, 4), (1, 0, 8), (2, 0, 8), (3, 16384, 4), (3, 4096, 8), (3, 65536, 2), (4, 4096, 4),
              (4, 2048, 4), (4, 1024, 4), (4, 256, 4)]
    for mode, chunk, bps in combos:
        def run():
            _lib.check(lib.dhd_probe_write_bw(ctypes.c_void_p(buf.data_ptr()), nbytes, mode, max(chunk, 256), bps,
                                              ctypes.c_void_p(st.cuda_stream)), 'probe')
        for _ in range(3):
            run()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); run(); e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        avg = sum(ts) / len(ts)
        print('probe mode=%d chunk=%-6d blocks/SM=%d   avg %.1f us  min %.1f us  %.0f GB/s avg' %
              (mode, chunk, bps, avg * 1e3, ts[0] * 1e3, nbytes / avg / 1e6), flush=True)
    a = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    for name, fn in (('cudaMemset (torch zero_)', lambda: buf.zero_()), ('torch copy_ (r+w bytes)', lambda: buf.copy_(a))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        avg = sum(ts) / len(ts)
        mult = 2 if 'copy' in name else 1
        print('%-30s avg %.1f us  min %.1f us  %.0f GB/s avg' % (name, avg * 1e3, ts[0] * 1e3, mult * nbytes / avg / 1e6), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'probes':
        sys.argv.pop(1)
        probes()
    main()
