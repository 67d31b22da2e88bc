#!/usr/bin/env python
"""Benchmark of the DHD view-transform / voxel-occupancy hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch
of synthetic 6-camera DHD-S input (BASELINE.json configs[1]: B=4 samples per GPU,
256x704 images -> 16x44 features, D=44, C=64, grids 200x200x{1,4,4,8}).
`value`  : samples/s with the inputs resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the public plugin call with HOST (pinned) inputs, H2D of
           the step inputs and D2H of the step result inside the timed region.
`roofline`: the fused pool kernel's algorithmic bytes / its own CUDA-event time vs the
           measured HBM copy peak in MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference path (the
           reference has no CPU pool kernel; its Python + our C restatement) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'samples/s DHD-S 6-cam 256x704 hot path (view transform + voxel pool)'
UNIT = 'samples/s'
B_PER_GPU = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--layout', default='nhwc', choices=['nhwc', 'nchw'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = max([int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()] or [0])
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                    'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from dhd_b200.pipeline import HotPathStep, algorithmic_bytes
    from oracle import mghs_oracle as O   # synthetic input generator + cpu_baseline leg only

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cfg = O.DHD_S
    B = B_PER_GPU
    inputs, depth, feat, height = O.synthetic_inputs(cfg, B, seed=100 + rank)
    step = HotPathStep(cfg, B, layout=args.layout)
    host = step.pin_host_inputs(inputs, depth, feat, height)      # pinned host copies
    dev = step.to_device(host)                                    # resident copies
    st = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step.run(dev)
    barrier()

    # ---- device-resident timing (value) + per-launch timing of the dominant kernel
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    ev0.record(st)
    for i in range(args.steps):
        step.run(dev, pool_events=kev[i])
    ev1.record(st)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    pool_ms = sorted(a.elapsed_time(b) for a, b in kev)
    pool_ms_avg = sum(pool_ms) / len(pool_ms)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the step result
    for _ in range(3):
        step.run_e2e(host)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step.run_e2e(host)
    e1.record(st)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_total, ms_e2e, pool_ms_avg], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, pool_ms_avg = t.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        alg = algorithmic_bytes(cfg, B)
        achieved = alg['pool_fwd_bytes'] / (pool_ms_avg * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': world * B * args.steps / (ms_total * 1e-3), 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {
                'workload': 'BASELINE configs[1]: DHD-S, 6-cam 256x704 -> 16x44 feats, D=44, C=64, '
                            'grids 200x200x{1,4,4,8}, batch=%d per GPU' % B,
                'stages': step.stage_names(), 'layout': args.layout, 'samples_per_gpu': B,
                'l2': 'pool output working set %.0f MB per step > 126 MB L2 (no explicit flush)'
                      % (alg['pool_fwd_bytes'] / 1e6),
                'sharding': 'batch axis, one process per GPU, no data-path collective',
            },
            'clocks': clocks,
            'e2e': {'value': world * B * args.steps / (ms_e2e * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': step.h2d_bytes, 'd2h_bytes_per_step': step.d2h_bytes},
            'gpu_launches': step.launches_per_step * args.steps,
            'roofline': {
                'kernel': 'mghs_pool_%s_kernel (fused 4-pass voxel pool forward)' % args.layout,
                'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'frac_of_nominal_8TBs': achieved / 8000.0,
                'peak_source': peak_src, 'traffic': step.ncu_traffic_bytes(),
                'algorithmic_bytes_per_launch': alg['pool_fwd_bytes'],
                'kernel_ms_avg': pool_ms_avg, 'kernel_ms_min': pool_ms[0],
                'kernel_share_of_step': pool_ms_avg / (ms_total / args.steps),
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference_leg(cfg, seconds=15.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------- CPU oracle leg
def cpu_step(cfg, B, seed, threads):
    """One pass of the reference algorithm on the host: the oracle port of MGHS.view_transform
    (4x get_ego_coor-equivalent geometry, prepare, pool) + the pool backward."""
    from oracle import mghs_oracle as O
    O._PoolFn.threads = threads
    inputs, depth, feat, height = O.synthetic_inputs(cfg, B, seed=seed)
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    depth = depth.requires_grad_()
    feat = feat.requires_grad_()
    t0 = time.perf_counter()
    outs = O.view_transform(inputs, depth, feat, height, fr, cfg['height_range'], cfg['mask_range'],
                            cfg['mask_grids'])
    sum(o.sum() for o in outs).backward()
    return time.perf_counter() - t0


def cpu_reference_leg(cfg, seconds):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_step(cfg, 1, 0, cores)            # warm-up (page in, build the C oracle)
    n, t = 0, 0.0
    while t < seconds and n < 50:
        t += cpu_step(cfg, 1, n + 1, cores)
        n += 1
    return {'value': n / t, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d steps of B=1 DHD-S samples (view_transform fwd + pool bwd), oracle port '
                      'of the reference Python + C restatement of its CUDA kernels, %.1f s' % (n, t)}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import mghs_oracle as O
    cfg = O.DHD_S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for i in range(max(1, min(args.warmup, 2))):
        cpu_step(cfg, 1, i, cores)
    steps = min(args.steps, 20)
    t = 0.0
    for i in range(steps):
        t += cpu_step(cfg, 1, 10 + i, cores)
    v = steps / t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': max(1, min(args.warmup, 2)), 'ms_per_step': 1e3 * t / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1] DHD-S hot path on the host cores; each step a '
                               'bounded sample of B=1 (one 6-camera frame set)'},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d steps of B=1' % steps},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
