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

_OUT = sys.stdout

METRIC = 'samples/s DHD-S 6-cam 256x704 hot path (DepthNet/HeightNet, MGHS view transform + voxel pool, SFA, occupancy head)'
UNIT = 'samples/s'
B_PER_GPU = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'bf16x3', 'fp32'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step extra')
    ap.add_argument('--no-encoders', action='store_true', help='skip the widened-path (real encoders) extra')
    ap.add_argument('--no-dhdl', action='store_true', help='skip the DHD-L (configs[4]) view-transformer extra')
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
    from dhd_b200 import _lib
    from dhd_b200 import shard
    from dhd_b200.pipeline import HotPathStep, algorithmic_bytes, dense_flops
    from dhd_b200 import synth as O        # the product arm never touches oracle/

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cfg = O.DHD_S
    B = B_PER_GPU
    step = HotPathStep(cfg, B, precision=args.precision, use_graph=not args.no_graph)
    rig = O.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=100 + rank)
    host = step.make_host_inputs(rig, seed=100 + rank)          # pinned host copies
    step.alloc_static(host)
    step.upload(host)                                           # resident copies
    graphed = step.capture()
    st = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step.run()
    barrier()

    # ---- device-resident timing (value) + in-place timing of the dominant HBM kernel
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    profile_range = os.environ.get('DHD_PROFILE_RANGE', '0') != '0'     # ncu --profile-from-start off: the timed steps only
    if profile_range:
        torch.cuda.cudart().cudaProfilerStart()
    ev0.record(st)
    for i in range(args.steps):
        step.run(pool_events=kev[i])
    ev1.record(st)
    barrier()
    if profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    ms_total = ev0.elapsed_time(ev1)
    pool_ms = sorted(a.elapsed_time(b) for a, b in kev)
    pool_ms_avg = sum(pool_ms) / len(pool_ms)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the occupancy class map
    # (streamed: step i+1's H2D and step i-1's D2H overlap step i's kernels on side streams; every
    #  step's copies are issued and completed inside the timed region)
    def e2e_steps(n):
        step.e2e_open(host)
        for i in range(n):
            step.run_e2e_streamed(host, host if i + 1 < n else None)
        step.e2e_close()

    e2e_steps(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    e2e_steps(args.steps)
    e1.record(st)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    # the same without overlap (copies and kernels serialised on one stream), for the report
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(st)
    for _ in range(min(args.steps, 10)):
        step.run_e2e(host)
    s1.record(st)
    barrier()
    ms_e2e_serial = s0.elapsed_time(s1) / min(args.steps, 10)
    clocks = sampler.stop() if rank == 0 else None

    # ---- extras: pool backward (a10) and per-stage times, outside the timed regions
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        step.run_pool_bwd()
    b0.record(st)
    for _ in range(10):
        step.run_pool_bwd()
    b1.record(st)
    torch.cuda.synchronize()
    bwd_ms = b0.elapsed_time(b1) / 10
    stage_ms = {}
    for name, fn in (('front(pack+depth_net+HeightNet+mask+prepare)', step._front), ('pool_fwd', step._pool),
                     ('back(SFA+predictor+argmax)', step._back)):
        if step.graph is not None and name != 'pool_fwd':
            fn = step.graph[0].replay if name.startswith('front') else step.graph[1].replay
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        s0.record(st)
        for _ in range(5):
            fn()
        s1.record(st)
        torch.cuda.synchronize()
        stage_ms[name] = s0.elapsed_time(s1) / 5

    # ---- extras: the widened path -- the real BEV / voxel encoders between pool and SFA
    widened = None
    if not args.no_encoders:
        wide = HotPathStep(cfg, B, precision=args.precision, use_graph=not args.no_graph, encoders=True)
        wide.alloc_static(host)
        wide.upload(host)
        wide_graphed = wide.capture()
        for _ in range(3):
            wide.run()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nw = max(3, min(args.steps, 10))
        w0.record(st)
        for _ in range(nw):
            wide.run()
        w1.record(st)
        barrier()
        ms_wide = shard.max_over_ranks([w0.elapsed_time(w1) / nw], device='cuda')[0]
        from dhd_b200.pipeline import encoder_flops
        widened = {'ms_per_step': ms_wide, 'samples_per_s': world * B / (ms_wide * 1e-3), 'cuda_graph': bool(wide_graphed),
                   'launches_per_step': wide.launches_per_step, 'stages': wide.stage_names(),
                   'encoder_tflops_algorithmic': encoder_flops(B) / 1e12,
                   'what': 'same step with CustomResNet + FPN_LSS and the three UNets (DHD-S.py:106-131, random-init '
                           'weights) instead of resident encoder features: image features -> occupancy classes'}
        del wide
        torch.cuda.empty_cache()

    # ---- extras: the data-parallel TRAINING step of the path (forward, loss, backward, one gradient
    # all-reduce over NCCL, AdamW, weight re-pack), see dhd_b200.pipeline.TrainStep for what is trainable
    train = None
    if not args.no_train:
        from dhd_b200.pipeline import TrainStep
        del step.graph
        step.graph = None
        ts = TrainStep(cfg, B, bn='batch')
        ts.alloc_static(host)
        ts.upload(host)
        lib = _lib.load()
        n0 = lib.dhd_launch_count()
        ts.train_step()
        launches = lib.dhd_launch_count() - n0
        train_graphed = ts.capture_train() if not args.no_graph else False
        for _ in range(3):
            ts.train_step()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nt = max(3, min(args.steps, 10))
        t0.record(st)
        for _ in range(nt):
            ts.train_step()
        t1.record(st)
        barrier()
        ms_train = t0.elapsed_time(t1) / nt
        train = {'ms_per_step': ms_train, 'launches_per_step': launches, 'cuda_graph': bool(train_graphed),
                 'graph_error': getattr(ts, 'train_graph_error', None),
                 'loss': float(ts.loss[0]), 'loss_sem_scal': float(ts.loss[2]), 'loss_geo_scal': float(ts.loss[3]),
                 'trainable_params': ts.n_params,
                 'gradient_all_reduce_bytes': ts.n_params * 4,
                 'loss_height': float(ts.loss_height[0]),
                 'what': 'GT binning of the sparse gt_depth / gt_height maps + forward + losses (occupancy CE + sem_scal + geo_scal, height BCE) + backward of depth_net, HeightNet, SFA and predictor '
                         '(BatchNorm2d in training mode: batch statistics, trainable affine; the ASPP Dropout(0.5) on), fused pool fwd+bwd, one NCCL all-reduce of the fp32 gradient bucket, AdamW, '
                         'bf16 weight re-pack; encoders stand in as resident tensors (see TrainStep)'}
        del ts
        torch.cuda.empty_cache()
        if not args.no_encoders:
            # the same step with the real encoders in forward and backward: the occupancy loss reaches the pool and
            # depth_net through them (no stand-in tensor left between the image features and the losses)
            tse = TrainStep(cfg, B, encoders=True, bn='batch')
            tse.alloc_static(host)
            tse.upload(host)
            tse.train_step()
            ge = tse.capture_train() if not args.no_graph else False
            for _ in range(2):
                tse.train_step()
            barrier()
            t0.record(st)
            for _ in range(nt):
                tse.train_step()
            t1.record(st)
            barrier()
            ms_te = shard.max_over_ranks([t0.elapsed_time(t1) / nt], device='cuda')[0]
            train['with_encoders'] = {'ms_per_step': ms_te, 'samples_per_s': world * B / (ms_te * 1e-3), 'cuda_graph': bool(ge),
                                      'graph_error': getattr(tse, 'train_graph_error', None),
                                      'trainable_params': tse.n_params, 'loss': float(tse.loss[0])}
            del tse
            torch.cuda.empty_cache()

    dhdl = None
    if world == 1 and not args.no_dhdl:
        dhdl = dhdl_extra(args.precision if args.precision in ('bf16', 'fp32') else 'bf16')

    ms_total, ms_e2e, pool_ms_avg = shard.max_over_ranks([ms_total, ms_e2e, pool_ms_avg], device='cuda')
    if train is not None:
        train['ms_per_step'] = shard.max_over_ranks([train['ms_per_step']], device='cuda')[0]
        train['samples_per_s'] = world * B / (train['ms_per_step'] * 1e-3)

    if rank == 0:
        peak, peak_src = peaks()
        alg = algorithmic_bytes(cfg, B)
        fl = dense_flops(cfg, B)
        achieved = alg['pool_fwd_bytes'] / (pool_ms_avg * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': world * B * args.steps / (ms_total * 1e-3), 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': {'bf16': 'bf16', 'bf16x3': 'bf16x3 (split-bf16)',
                                           'fp32': 'f32 (6-term split-bf16)'}[args.precision],
            'data': 'synthetic',
            'config': {
                'workload': 'BASELINE configs[1]: DHD-S hot path inference, 6-cam 256x704 -> 16x44x256 image '
                            'features, D=44, C=64, grids 200x200x{1,4,4,8}, SFA 512->256 + predictor @200x200, '
                            'batch=%d per GPU; random-init weights' % B,
                'stages': step.stage_names(), 'samples_per_gpu': B, 'precision': args.precision,
                'pool_arithmetic': 'f32', 'cuda_graph': bool(graphed),
                'encoders': 'BEV/voxel encoders are outside the SURVEY 8 path: resident synthetic (B,512,200,200) bf16 NHWC features, the '
                            'form dhd_b200.encoders produces (extras.with_encoders runs the real encoders instead)',
                'l2': 'per-step working set > 1 GB (pool outputs 696 MB, BEV activations) > 126 MB L2, no explicit flush',
                'sharding': 'batch axis, one process per GPU, no data-path collective',
            },
            'clocks': clocks,
            'e2e': {'value': world * B * args.steps / (ms_e2e * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': step.h2d_bytes, 'd2h_bytes_per_step': step.d2h_bytes,
                    'how': 'HotPathStep.run_e2e_streamed: pinned host inputs -> H2D (side stream, overlaps the previous '
                           "step's kernels) -> step -> D2H of the uint8 class map to pinned host memory (side stream)"},
            'gpu_launches': step.launches_per_step * args.steps * 2,
            'roofline': {
                'kernel': 'mghs_pool_stream_kernel (fused 4-pass voxel pool forward, TMA bulk stores)',
                'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'frac_of_nominal_8TBs': achieved / 8000.0,
                'peak_source': peak_src, 'traffic': step.ncu_traffic_bytes(),
                'algorithmic_bytes_per_launch': alg['pool_fwd_bytes'],
                'kernel_ms_avg': pool_ms_avg, 'kernel_ms_min': pool_ms[0],
                'kernel_share_of_step': pool_ms_avg / (ms_total / args.steps),
            },
            'extras': {
                'stage_ms': stage_ms, 'pool_bwd_ms': bwd_ms, 'train_step': train, 'with_encoders': widened,
                'dhd_l_view_transformer': dhdl,
                'e2e_serialised_ms_per_step (H2D, kernels, D2H on one stream)': ms_e2e_serial,
                'dense_tflops_algorithmic': {k: v / 1e12 for k, v in fl.items()},
                'dense_tflop_per_s': sum(fl.values()) / 1e12 /
                (1e-3 * max(1e-9, stage_ms['front(pack+depth_net+HeightNet+mask+prepare)'] +
                            stage_ms['back(SFA+predictor+argmax)'])),
                'launches_per_step': step.launches_per_step,
                'graph_error': getattr(step, 'graph_error', None),
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference_leg(cfg, seconds=20.0)
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def dhdl_extra(precision, B=2):
    """BASELINE configs[4]: the plugin's MGHS_Stereo with the DHD-L.py kwargs at full size (6-cam 512x1408, C_in=512,
    D=88, two-frame stereo features 128 ch @128x352): cost volume -> stereo DepthNet + HeightNet -> fused pool."""
    from dhd_b200 import synth
    vt, vargs, metas = synth.dhdl_view_transformer(precision, B)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, it=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(it):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / it

    with torch.no_grad():
        ms = timed(lambda: vt(vargs, metas))
        ms_cv = timed(lambda: vt.depth_net.calculate_cost_volumn(metas))
    del vt, vargs, metas
    torch.cuda.empty_cache()
    return {'ms_per_step': ms, 'samples_per_s': B / ms * 1e3, 'samples': B, 'precision': precision,
            'stereo_cost_volume_ms': ms_cv,
            'what': 'MGHS_Stereo.forward(input, stereo_metas) of projects/configs/DHD/DHD-L.py: NCHW->NHWC of both stereo '
                    'features, fused plane-sweep cost volume (gen_grid + 32 grid_sample groups + softmax of the reference '
                    'in one kernel), cost_volumn_net, stereo DepthNet and HeightNet at C=512 on tcgen05, fused '
                    'collapse_z=False voxel pool -> (B,64,1,200,200) + (B,64,16,200,200)'}


# ----------------------------------------------------------------------- CPU oracle leg
_CPU_STATE = {}


def cpu_step(cfg, B, seed, threads):
    """One pass of the reference algorithm on the host for B samples: dense front (depth_net,
    HeightNet), the oracle port of MGHS.view_transform (4x geometry, prepare, pool), SFA and
    predictor on synthetic encoder features, class argmax -- the same stages as the GPU step."""
    from oracle import dense_oracle as DO
    from oracle import mghs_oracle as O
    O._PoolFn.threads = threads
    if 'sd' not in _CPU_STATE:
        import projects.mmdet3d_plugin  # noqa: F401  (parameter containers only; forward is the oracle's)
        from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
        from projects.mmdet3d_plugin.models.model_utils.depthnet import HeightNet
        from projects.mmdet3d_plugin.models.necks.mix import SFA
        torch.manual_seed(0)
        _CPU_STATE['sd'] = (HeightNet(256, 256, 65).eval().state_dict(), torch.nn.Conv2d(256, 108, 1).state_dict(),
                            SFA(512, 256).eval().state_dict(),
                            predictor(256, 256, 16, num_classes=18, loss_occ=None).eval().state_dict())
    hn_sd, dn_sd, sfa_sd, head_sd = _CPU_STATE['sd']
    N = cfg['ncams']
    fH, fW = cfg['input_size'][0] // cfg['downsample'], cfg['input_size'][1] // cfg['downsample']
    rig = O.synthetic_rig(B, N, cfg['input_size'], seed=seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B * N, 256, fH, fW, generator=g)
    enc = torch.randn(B, 512, 200, 200, generator=g)
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    t0 = time.perf_counter()
    with torch.no_grad():
        y = torch.nn.functional.conv2d(x, dn_sd['weight'], dn_sd['bias'])
        depth, feat = y[:, :44].softmax(1), y[:, 44:].contiguous()
        s2e, e2g, K, pr, pt, bda = rig
        mlp = torch.zeros(B, N, 27)
        height = DO.heightnet_forward(hn_sd, x, mlp).softmax(1)
        inputs = (torch.zeros(B, N, 1, fH, fW),) + tuple(rig)
        O.view_transform(inputs, depth, feat, height, fr, cfg['height_range'], cfg['mask_range'], cfg['mask_grids'])
        occ = DO.predictor_forward(head_sd, DO.sfa_forward(sfa_sd, enc))
        occ.argmax(-1).to(torch.uint8)
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
            'sample': '%d steps of B=1 DHD-S samples (same stages as the GPU step: dense front, 4-pass view '
                      'transform, SFA, predictor, argmax); oracle port = reference Python + torch CPU convs + '
                      'C restatement of the bev_pool_v2 kernels, %.1f s' % (n, t)}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import mghs_oracle as O
    cfg = O.DHD_S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    warm = max(1, min(args.warmup, 2))
    for i in range(warm):
        cpu_step(cfg, 1, i, cores)
    steps = min(args.steps, 20)
    t = 0.0
    for i in range(steps):
        t += cpu_step(cfg, 1, 10 + i, cores)
        if t > 120 and i >= 2:
            steps = i + 1
            break
    v = steps / t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * t / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1] DHD-S hot path inference on the host cores (reference '
                               'algorithm: oracle port, the reference ships no CPU kernel for bev_pool_v2); '
                               'each step a bounded sample of B=1 (one 6-camera frame set)'},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d steps of B=1' % steps},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), file=_OUT, flush=True)


if __name__ == '__main__':
    # stdout carries exactly ONE line (the JSON): libraries that write to fd 1 themselves (NCCL's version
    # banner) are sent to stderr for the duration of the run
    _OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
