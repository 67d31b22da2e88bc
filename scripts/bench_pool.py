"""Pool-forward micro-benchmark on the DHD-S B=4 workload: sweeps the env tunables of
dhd_mghs_pool_fwd (csrc/mghs_pool.cu) and checks every variant bit-for-bit against the default configuration
(PROBE=1 = every cell treated as empty: the write-only ceiling of the kernel's own address pattern).
Usage: python scripts/bench_pool.py ["ZT=4096,THREADS=128,CELLCOST=8,HINT=1,PERSM=4,NCH=32768" ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200.pipeline import HotPathStep, algorithmic_bytes  # noqa: E402
from oracle import mghs_oracle as O  # noqa: E402


def main():
    variants = sys.argv[1:] or ['NCH=32768', 'PROBE=1']
    os.environ['DHD_POOL_SWEEP'] = '1'          # the library re-reads its tunables on every call
    cfg, B = O.DHD_S, 4
    step = HotPathStep(cfg, B, precision='bf16', use_graph=False)
    rig = O.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=100)
    host = step.make_host_inputs(rig, seed=100)
    step.alloc_static(host)
    step.upload(host)
    step._front()
    torch.cuda.synchronize()
    alg = algorithmic_bytes(cfg, B)['pool_fwd_bytes']
    for o in step.outs:
        o.fill_(float('nan'))
    step._pool()
    torch.cuda.synchronize()
    want = [o.clone() for o in step.outs]          # the default configuration is the reference every variant must equal
    st = torch.cuda.current_stream()
    for v in variants:
        for k in list(os.environ):
            if k.startswith('DHD_POOL_') and k != 'DHD_POOL_SWEEP':
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


if __name__ == '__main__':
    main()
