#!/usr/bin/env python
"""Benchmark of the DHD view-transform / voxel-occupancy hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch
of synthetic 6-camera DHD-S input (BASELINE.json configs[1] / [2]: B=4 samples per GPU,
256x704 images -> 16x44 features, D=44, C=64, grids 200x200x{1,4,4,8}) -- the data-parallel
TRAINING step (forward + backward; north_star: one NCCL gradient all-reduce),
so the 1 -> 8 GPU curve measures a step that contains the collective.
`value`  : samples/s with the inputs resident in HBM (CUDA events, max over ranks), over a
           timed region of >= 2.5 s (steps x inner_repeats passes: sustained clocks).
`e2e`    : same metric with HOST (pinned) inputs, H2D of the step inputs (features, cameras,
           labels, LiDAR maps) and D2H of the step result (losses) inside the timed region.
`roofline`: the fused pool kernel's algorithmic bytes / its own CUDA-event time (in place,
           inside the training step) vs the measured HBM copy peak in MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the same training step (the
           reference has no CPU pool kernel; its Python + our C restatement) on the host cores.
`extras` : inference (image features -> class map; bf16 and fp32 precision modes, e2e),
           the widened steps with the real encoders, the plugin detector from camera images
           (forward_train incl. the image backbone / simple_test), DHD-L, the reference CUDA path.
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

METRIC = ('samples/s DHD-S 6-cam 256x704 hot path training step (fwd+bwd of DepthNet/HeightNet, MGHS view transform + voxel pool, SFA, '
          'occupancy head; NCCL gradient all-reduce); voxel-pool HBM GB/s vs peak in `roofline`')
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
    ap.add_argument('--quick', action='store_true', help='short timed regions (smoke runs under a profiler)')
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
MIN_REGION_S = 2.5          # every headline timed region lasts at least this long (sustained clocks, >= 20 clock samples)


def _region(run, K, min_seconds, barrier, st, shard, events=0):
    """Time K x R passes of `run` between two CUDA events, R chosen (identically on every rank: max over ranks of a
    calibration pass) so that the region lasts >= min_seconds.  `run(ev)` gets a (start, stop) event pair for the first
    `events` passes (in-place timing of one kernel) and None afterwards.  Returns (ms_total, passes, R, [kernel ms])."""
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    c0.record(st)
    for _ in range(K):
        run(None)
    c1.record(st)
    barrier()
    ms_pass = shard.max_over_ranks([c0.elapsed_time(c1) / K], device='cuda')[0]
    R = max(1, int(-(-min_seconds * 1e3 // max(1e-3, K * ms_pass))))
    n = K * R
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(events, n))]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for i in range(n):
        run(kev[i] if i < len(kev) else None)
    e1.record(st)
    barrier()
    return e0.elapsed_time(e1), n, R, sorted(a.elapsed_time(b) for a, b in kev)


def _e2e_region(step, host, K, min_seconds, barrier, st, shard, ms_hint):
    """Same, through the host-buffer entry point: every pass moves its inputs host -> device and its result back."""
    R = max(1, int(-(-min_seconds * 1e3 // max(1e-3, K * ms_hint))))
    n = K * R

    def go(m):
        step.e2e_open(host)
        for i in range(m):
            step.run_e2e_streamed(host, host if i + 1 < m else None)
        step.e2e_close()
    go(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    go(n)
    e1.record(st)
    barrier()
    return e0.elapsed_time(e1), n


def run_ours(args):
    import torch.distributed as dist
    from dhd_b200 import _lib
    from dhd_b200 import shard
    from dhd_b200.pipeline import HotPathStep, TrainStep, algorithmic_bytes, dense_flops, pool_traffic_bytes
    from dhd_b200 import synth as O        # the product arm never touches oracle/

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cfg = O.DHD_S
    B = B_PER_GPU
    K, W = args.steps, max(args.warmup, 3)
    st = torch.cuda.current_stream()
    lib = _lib.load()
    quick = args.quick
    min_s = 0.2 if quick else MIN_REGION_S

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ================================================================ headline: the data-parallel TRAINING step
    # samples/s of the hot path forward + backward; north_star: batch-axis shard + ONE NCCL gradient
    # all-reduce.  bf16 operands / fp32 accumulation, master weights and gradients; BatchNorm on batch statistics.
    ts = TrainStep(cfg, B, bn='batch')
    rig = O.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=100 + rank)
    host = ts.make_host_inputs(rig, seed=100 + rank)            # pinned host copies (features, cameras, labels, LiDAR maps)
    ts.alloc_static(host)
    ts.upload(host)                                             # resident copies
    n0 = lib.dhd_launch_count()
    ts.train_step()
    torch.cuda.synchronize()
    launches = lib.dhd_launch_count() - n0
    graphed = ts.capture_train() if not args.no_graph else False
    for _ in range(W):
        ts.train_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    profile_range = os.environ.get('DHD_PROFILE_RANGE', '0') != '0'     # ncu --profile-from-start off: the timed steps only
    if profile_range:
        torch.cuda.cudart().cudaProfilerStart()
    ms_total, n_timed, R, pool_ms = _region(ts.train_step, K, 0.0 if profile_range else min_s, barrier, st, shard, events=256)
    if profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    ms_e2e, n_e2e = _e2e_region(ts, host, K, 0.0 if profile_range else min_s, barrier, st, shard, ms_total / n_timed)
    clocks = sampler.stop() if rank == 0 else None
    pool_ms_avg = sum(pool_ms) / len(pool_ms)
    losses = [float(v) for v in ts.result.tolist()]
    hashes_equal = None
    if world > 1:                                                # replicas that stepped in lock-step hold identical weights
        h = ts.weight_hash().reshape(1)
        allh = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(allh, h)
        hashes_equal = bool(all(torch.equal(a, allh[0]) for a in allh))
        if not hashes_equal:
            raise RuntimeError('replica weights diverged after %d data-parallel steps: %s' % (n_timed, [float(a) for a in allh]))
    train_extra = {'launches_per_step': launches, 'cuda_graph': bool(graphed), 'graph_error': getattr(ts, 'train_graph_error', None),
                   'loss_occ': losses[0], 'loss_sem_scal': losses[2], 'loss_geo_scal': losses[3], 'loss_height': losses[4],
                   'trainable_params': ts.n_params, 'gradient_all_reduce_bytes': ts.n_params * 4,
                   'grad_clip_max_norm': ts.grad_clip, 'replica_weight_hashes_equal': hashes_equal}
    train_stages = ts.stage_names()
    h2d_train, d2h_train = ts.h2d_bytes, ts.d2h_bytes
    host_infer = {k: host[k] for k in ('x', 'sensor2ego', 'ego2global', 'cam2imgs', 'post_rots', 'post_trans', 'bda')}
    del ts
    torch.cuda.empty_cache()

    # ================================================================ extras
    extras = {'train_step': train_extra}
    # ---- the same step with the real BEV / voxel encoders in forward and backward
    if not args.no_encoders and not args.no_train:
        tse = TrainStep(cfg, B, encoders=True, bn='batch')
        tse.alloc_static(host)
        tse.upload(host)
        tse.train_step()
        ge = tse.capture_train() if not args.no_graph else False
        for _ in range(2):
            tse.train_step()
        ms, n, _, _ = _region(tse.train_step, max(3, min(K, 10)), 0.2 if quick else 1.0, barrier, st, shard)
        ms = shard.max_over_ranks([ms], device='cuda')[0]
        extras['train_step_with_encoders'] = {
            'ms_per_step': ms / n, 'samples_per_s': world * B * n / (ms * 1e-3), 'cuda_graph': bool(ge), 'timed_passes': n,
            'graph_error': getattr(tse, 'train_graph_error', None), 'trainable_params': tse.n_params,
            'gradient_all_reduce_bytes': tse.n_params * 4, 'loss_occ': float(tse.loss[0]),
            'what': 'the training step with CustomResNet + FPN_LSS and the three UNets (DHD-S.py:106-131) in forward and '
                    'backward: the occupancy loss reaches the pool and depth_net through them, no stand-in tensor'}
        del tse
        torch.cuda.empty_cache()

    # ---- inference (the round-1 headline): image features -> occupancy class map
    def inference(precision, encoders, seconds, images=False):
        step = HotPathStep(cfg, B, precision=precision, use_graph=not args.no_graph, encoders=encoders, images=images)
        h_in = step.make_host_inputs(rig, seed=100 + rank)
        if not images:
            h_in = host_infer
        step.alloc_static(h_in)
        step.upload(h_in)
        g = step.capture()
        for _ in range(3):
            step.run()
        ms, n, _, pms = _region(lambda ev: step.run(pool_events=ev), K, seconds, barrier, st, shard, events=64)
        out = {'precision': precision, 'ms_per_step': None, 'cuda_graph': bool(g), 'launches_per_step': step.launches_per_step,
               'stages': step.stage_names(), 'graph_error': getattr(step, 'graph_error', None)}
        if images:
            ms_e, n_e = _e2e_region(step, h_in, K, seconds, barrier, st, shard, ms / n)
            ms_e = shard.max_over_ranks([ms_e], device='cuda')[0]
            out['e2e'] = {'value': world * B * n_e / (ms_e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': step.h2d_bytes,
                          'd2h_bytes_per_step': step.d2h_bytes}
        if not encoders and not images and precision == args.precision:
            ms_e, n_e = _e2e_region(step, host_infer, K, seconds, barrier, st, shard, ms / n)
            ms_e = shard.max_over_ranks([ms_e], device='cuda')[0]
            out['e2e'] = {'value': world * B * n_e / (ms_e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': step.h2d_bytes,
                          'd2h_bytes_per_step': step.d2h_bytes}
            stage_ms = {}
            for name, fn in (('front(pack+depth_net+HeightNet+mask+prepare)', step._front), ('pool_fwd', step._pool),
                             ('back(SFA+predictor tail)', step._back)):
                if step.graph is not None and name != 'pool_fwd':
                    fn = step.graph[0].replay if name.startswith('front') else step.graph[1].replay
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                fn()
                s0.record(st)
                for _ in range(20):
                    fn()
                s1.record(st)
                torch.cuda.synchronize()
                stage_ms[name] = s0.elapsed_time(s1) / 20
            out['stage_ms'] = stage_ms
            fl = dense_flops(cfg, B)
            out['dense_tflops_algorithmic'] = {k: v / 1e12 for k, v in fl.items()}
            out['dense_tflop_per_s'] = sum(fl.values()) / 1e12 / (1e-3 * max(1e-9, stage_ms['front(pack+depth_net+HeightNet+mask+prepare)'] +
                                                                           stage_ms['back(SFA+predictor tail)']))
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                step.run_pool_bwd()
            b0.record(st)
            for _ in range(20):
                step.run_pool_bwd()
            b1.record(st)
            torch.cuda.synchronize()
            out['pool_bwd_ms'] = b0.elapsed_time(b1) / 20
        ms = shard.max_over_ranks([ms], device='cuda')[0]
        out['ms_per_step'] = ms / n
        out['value'] = world * B * n / (ms * 1e-3)
        out['timed_passes'] = n
        if pms:
            out['pool_fwd_ms_avg'] = sum(pms) / len(pms)
        del step
        torch.cuda.empty_cache()
        return out

    extras['inference'] = inference(args.precision, False, 0.2 if quick else 1.5)
    extras['inference']['what'] = ('HotPathStep: image features -> depth_net / HeightNet -> masks -> fused 4-pass pool -> [resident encoder '
                                   'features] -> SFA -> predictor with the fused Linear+Softplus+Linear+argmax tail -> uint8 class map')
    if args.precision != 'fp32' and not quick:
        extras['inference_fp32'] = inference('fp32', False, 0.5)
        extras['inference_fp32']['what'] = ('the same step in the precision mode the 1e-4 logits contract is tested in (6-term split-bf16 '
                                            'MMAs, fp32 accumulation; layer-by-layer head + dhd_occ_argmax)')
    if not args.no_encoders:
        extras['inference_with_encoders'] = inference(args.precision, True, 0.2 if quick else 1.0)
        from dhd_b200.pipeline import encoder_flops
        extras['inference_with_encoders']['encoder_tflops_algorithmic'] = encoder_flops(B) / 1e12
    if not args.no_encoders:
        # the whole detector: camera IMAGES -> ResNet-50 + CustomFPN -> view transformer -> encoders -> SFA -> head
        extras['inference_from_images'] = inference(args.precision, True, 0.2 if quick else 1.0, images=True)
        extras['inference_from_images']['what'] = (
            'DHD-S end to end from the 6 x 256x704 camera images (img_backbone ResNet-50 + img_neck CustomFPN of DHD-S.py:44-62 '
            'on the tcgen05 convolution kernel, then the widened hot path with the real BEV / voxel encoders) to the uint8 '
            'class map; e2e = pinned host images in, class map out')
    if world == 1 and not args.no_encoders and not args.no_train and not quick:
        # (one GPU only: an extra that fails on one rank must not be able to hang a multi-rank run)
        # the reference's own entry points on the whole DHD-S detector INCLUDING its image backbone / neck (DHD-S.py:44-62
        # trains them): DHD.forward_train on camera images -> backward -> all-reduce -> clip -> AdamW, and DHD.simple_test
        try:
            from dhd_b200 import synth as _synth
            from dhd_b200.detector_step import DetectorStep
            dstep = DetectorStep(_synth.dhd_s_model_cfg(args.precision if args.precision == 'bf16' else 'bf16', images=True), B,
                                 seed=rank)
            d_in, d_kw = dstep.make_inputs(300 + rank)
            for _ in range(2):
                d_losses = dstep.train_step(d_in, d_kw)
            ms, n, _, _ = _region(lambda ev: dstep.train_step(d_in, d_kw), 3, 0.5, barrier, st, shard)
            ms = shard.max_over_ranks([ms], device='cuda')[0]
            det = {'train_ms_per_step': ms / n, 'train_samples_per_s': world * B * n / (ms * 1e-3), 'train_timed_passes': n,
                   'trainable_params': dstep.n_params, 'gradient_all_reduce_bytes': dstep.n_params * 4,
                   'losses': {k: float(v) for k, v in d_losses.items()}, 'train_cuda_graph': False}
            graphed = dstep.capture_infer(d_in)
            run_inf = (lambda ev: dstep.infer_step_graphed()) if graphed else (lambda ev: dstep.infer_step(d_in))
            for _ in range(2):
                run_inf(None)
            ms, n, _, _ = _region(run_inf, 5, 0.3, barrier, st, shard)
            ms = shard.max_over_ranks([ms], device='cuda')[0]
            det.update(infer_ms_per_step=ms / n, infer_samples_per_s=world * B * n / (ms * 1e-3), infer_cuda_graph=bool(graphed),
                       what='the plugin DETECTOR driven the way the reference runner drives it, from 6 x 256x704 camera images: '
                            'DHD.forward_train (DHD_model.py:135-186; ResNet-50 + CustomFPN with batch-statistics BatchNorm and a '
                            'hand-written backward, view transformer, encoders, SFA, head, four losses) + backward + NCCL gradient '
                            'all-reduce + clip 5 + AdamW, eager launches through dhd_b200.autograd; and DHD.simple_test (bf16 NHWC '
                            'activation path, the whole call as one CUDA graph, class maps copied to the host)')
            extras['detector_api_from_images'] = det
            del dstep, d_in, d_kw
        except Exception as e:  # noqa: BLE001 -- an extra must never cost the bench line
            extras['detector_api_from_images'] = {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_dhdl:
        extras['dhd_l_view_transformer'] = dhdl_extra(args.precision if args.precision in ('bf16', 'fp32') else 'bf16')
    if world == 1 and not quick:
        extras['reference_cuda_path'] = reference_cuda_leg(B)

    ms_total, ms_e2e, pool_ms_avg = shard.max_over_ranks([ms_total, ms_e2e, pool_ms_avg], device='cuda')

    if rank == 0:
        peak, peak_src = peaks()
        alg = algorithmic_bytes(cfg, B)
        achieved = alg['pool_fwd_bytes'] / (pool_ms_avg * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': world * B * n_timed / (ms_total * 1e-3), 'unit': UNIT,
            'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms_total / n_timed, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {
                'workload': 'BASELINE configs[1] / configs[2]: DHD-S hot path TRAINING step, 6-cam 256x704 -> 16x44x256 image '
                            'features, D=44, C=64, grids 200x200x{1,4,4,8}, SFA 512->256 + predictor @200x200, batch=%d per GPU, '
                            'global batch %d; GT binning + forward + losses (occupancy CE + sem_scal + geo_scal, height BCE) + backward '
                            'of depth_net, HeightNet, pool, SFA, predictor + ONE NCCL all-reduce of the fp32 gradient bucket + grad '
                            'clip 5 + AdamW + bf16 weight re-pack; BatchNorm on batch statistics, ASPP Dropout on; random-init '
                            'weights' % (B, B * world),
                'stages': train_stages, 'samples_per_gpu': B, 'precision': 'bf16 operands, fp32 accumulation / master weights / gradients',
                'pool_arithmetic': 'f32', 'cuda_graph': bool(graphed),
                'inner_repeats': R, 'timed_passes': n_timed, 'timed_region_s': ms_total * 1e-3,
                'timing': 'steps x inner_repeats back-to-back passes between one pair of CUDA events (>= %.1f s, sustained clocks); '
                          'ms_per_step is per pass' % min_s,
                'encoders': 'BEV/voxel encoders are outside the SURVEY 8 path: resident synthetic (B,512,200,200) bf16 NHWC features feed the '
                            'SFA and resident synthetic gradients feed the pool backward (extras.train_step_with_encoders runs the real '
                            'encoders in forward and backward)',
                'l2': 'per-step working set > 2 GB (pool outputs 696 MB + their gradients, BEV activations) > 126 MB L2, no explicit flush',
                'sharding': 'batch axis, one process per GPU, weights replicated, one gradient all-reduce per step (NCCL)',
            },
            'clocks': clocks,
            'e2e': {'value': world * B * n_e2e / (ms_e2e * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': h2d_train, 'd2h_bytes_per_step': d2h_train, 'timed_passes': n_e2e,
                    'how': 'TrainStep.run_e2e_streamed: pinned host inputs (image features, camera tensors, voxel_semantics, '
                           'mask_camera, gt_depth, gt_height) -> H2D on a side stream (overlaps the previous step) -> training step -> '
                           'D2H of the five loss scalars to pinned host memory'},
            'gpu_launches': launches * (n_timed + n_e2e),
            'roofline': {
                'kernel': 'mghs_pool_stream_kernel (fused 4-pass voxel pool forward, TMA bulk stores), timed in place inside the training step',
                'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'frac_of_nominal_8TBs': achieved / 8000.0,
                'peak_source': peak_src, 'traffic': pool_traffic_bytes(),
                'algorithmic_bytes_per_launch': alg['pool_fwd_bytes'],
                'kernel_ms_avg': pool_ms_avg, 'kernel_ms_min': pool_ms[0], 'kernel_launches_timed': len(pool_ms),
                'kernel_share_of_step': pool_ms_avg / (ms_total / n_timed),
            },
            'extras': extras,
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference_leg(cfg, seconds=6.0 if quick else 20.0)
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_cuda_leg(B):
    """The reference's own CUDA path of MGHS.view_transform (its op sequence as torch CUDA ops + its UNMODIFIED
    bev_pool_v2 kernel, compiled from /root/reference into oracle/_ref) timed beside the fused path on the same
    inputs: the denominator of north_star's ">= 10x the reference bev_pool_v2 CUDA path" (baseline leg: the one place
    besides cpu_baseline where bench.py executes oracle/)."""
    try:
        from oracle import mghs_oracle as O
        from oracle import ref_cuda_path as R
    except Exception as e:  # noqa: BLE001
        return {'unavailable': repr(e)[:200]}
    if not R.available():
        return {'unavailable': 'oracle/_ref/libbev_pool_v2_ref.so not built (needs /root/reference at build time)'}
    from dhd_b200.pool import MghsPool, height_to_mask
    cfg = O.DHD_S
    inputs, depth, feat, height = O.synthetic_inputs(cfg, B, seed=3)
    inputs = tuple(t.cuda() for t in inputs)
    depth, feat, height = depth.cuda(), feat.cuda(), height.cuda()
    N, D = cfg['ncams'], depth.shape[1]
    fH, fW = depth.shape[-2:]
    C = cfg['C']
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample']).cuda()
    grids = [cfg['bev_grid']] + list(cfg['mask_grids'])
    plan = MghsPool(B, N, D, fH, fW, C, grids[0]['x'], grids[0]['y'], [(g['z'], m) for m, g in enumerate(grids)])
    ws = torch.empty(plan.ws_bytes, dtype=torch.uint8, device='cuda')
    outs = plan.alloc_outputs('nhwc', 'cuda')
    f_nhwc = feat.view(B, N, C, fH, fW).permute(0, 1, 3, 4, 2).contiguous()
    _, s2e, _e2g, Kc, pr, pt, bda = inputs

    def ref():
        return R.view_transform_cuda(inputs, depth, feat, height, fr, cfg['height_range'], cfg['mask_range'], cfg['mask_grids'])

    def ours():
        pm = height_to_mask(height, cfg['height_range'], cfg['mask_range'])
        plan.prepare(frustum=fr, sensor2ego=s2e, cam2imgs=Kc, post_rots=pr, post_trans=pt, bda=bda, workspace=ws)
        plan.raw_forward(depth, f_nhwc, pm, outs, 'nhwc', workspace=ws)

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    want = ref()
    coor = O.ego_coor(fr, s2e, Kc, pr, pt, bda)            # parity on identical coordinates first
    plan.prepare(coor=coor, workspace=ws)
    plan.raw_forward(depth, f_nhwc, height_to_mask(height, cfg['height_range'], cfg['mask_range']), outs, 'nhwc', workspace=ws)
    torch.cuda.synchronize()
    ok = all(torch.allclose(o.permute(0, 3, 1, 2), w, rtol=1e-5, atol=2e-6) for o, w in zip(outs, want))
    del want
    ms_ref, ms_ours = timed(ref, 10), timed(ours, 50)
    return {'reference_cuda_ms': ms_ref, 'fused_ms': ms_ours, 'speedup': ms_ref / ms_ours, 'outputs_match': bool(ok),
            'what': 'MGHS.view_transform forward, DHD-S B=%d, fp32: reference op sequence (4 x get_ego_coor + voxel_pooling_prepare_v2 '
                    'with argsort + zero-fill + the reference\'s unmodified bev_pool_v2 kernel + permute + collapse-Z cat, LH:179-231, '
                    '303-371, 407-459) vs height_to_mask + dhd_mghs_prepare + dhd_mghs_pool_fwd on the same inputs' % B}


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
    """One TRAINING step of the reference algorithm on the host for B samples -- the same stages as the GPU step:
    dense front (depth_net, HeightNet with BatchNorm on batch statistics), the oracle port of MGHS.view_transform
    (4x geometry, prepare with argsort, pool), height loss, SFA + predictor on resident encoder features, the three
    occupancy loss terms, torch autograd backward through all of it (the pool backward = the C restatement of
    bev_pool_grad_kernel; resident synthetic gradients at the pool outputs, as the GPU step), grad clip 5, AdamW."""
    from oracle import dense_oracle as DO
    from oracle import loss_oracle as LO
    from oracle import mghs_oracle as O
    O._PoolFn.threads = threads
    if 'sd' not in _CPU_STATE:
        import projects.mmdet3d_plugin  # noqa: F401  (parameter containers only; forward is the oracle's)
        from projects.mmdet3d_plugin.models.dense_heads.occ_head import nusc_class_frequencies, predictor
        from projects.mmdet3d_plugin.models.model_utils.depthnet import HeightNet
        from projects.mmdet3d_plugin.models.necks.mix import SFA
        torch.manual_seed(0)
        sds = [HeightNet(256, 256, 65).state_dict(), torch.nn.Conv2d(256, 108, 1).state_dict(), SFA(512, 256).state_dict(),
               predictor(256, 256, 16, num_classes=18, loss_occ=None).state_dict()]
        leaves = []
        for sd in sds:
            for k, v in sd.items():
                if v.dtype.is_floating_point and 'running' not in k:
                    sd[k] = v.clone().requires_grad_(True)
                    leaves.append(sd[k])
        import numpy as np
        cw = torch.from_numpy(1 / np.log(nusc_class_frequencies[:18] + 0.001)).float()
        _CPU_STATE['sd'] = tuple(sds) + (torch.optim.AdamW(leaves, lr=2e-4, weight_decay=1e-2), leaves, cw)
    hn_sd, dn_sd, sfa_sd, head_sd, opt, leaves, cw = _CPU_STATE['sd']
    N = cfg['ncams']
    h_in, w_in = cfg['input_size']
    fH, fW = h_in // cfg['downsample'], w_in // cfg['downsample']
    rig = O.synthetic_rig(B, N, cfg['input_size'], seed=seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B * N, 256, fH, fW, generator=g)
    enc = torch.randn(B, 512, 200, 200, generator=g)
    labels = torch.randint(0, 18, (B, 200, 200, 16), generator=g)
    mask = torch.rand(B, 200, 200, 16, generator=g) < 0.5
    hit = torch.rand(B, N, h_in, w_in, generator=g) < 0.02
    gt_depth = torch.where(hit, 1.0 + 59.0 * torch.rand(B, N, h_in, w_in, generator=g), torch.zeros(()))
    gt_height = torch.where(hit, -2.0 + 8.0 * torch.rand(B, N, h_in, w_in, generator=g), torch.zeros(()))
    fr = O.frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    shapes = [(B, dz * 64, 200, 200) for dz in (1, 4, 4, 8)]
    gouts = [1e-3 * torch.randn(sh, generator=g) for sh in shapes]
    t0 = time.perf_counter()
    DO.BN_TRAIN = True
    try:
        opt.zero_grad(set_to_none=True)
        y = torch.nn.functional.conv2d(x, dn_sd['weight'], dn_sd['bias'])
        depth, feat = y[:, :44].softmax(1), y[:, 44:].contiguous()
        mlp = torch.zeros(B, N, 27)
        height = DO.heightnet_forward(hn_sd, x, mlp).softmax(1)
        inputs = (torch.zeros(B, N, 1, fH, fW),) + tuple(rig)
        pooled = O.view_transform(inputs, depth, feat, height, fr, cfg['height_range'], cfg['mask_range'], cfg['mask_grids'])
        loss = sum((p * go).sum() for p, go in zip(pooled, gouts))           # resident gradients of the encoders' inputs
        loss = loss + LO.height_loss(gt_depth, gt_height, height, [1.0, 45.0, 0.5], 44, -1.0, 0.1, 0.1)
        occ = DO.predictor_forward(head_sd, DO.sfa_forward(sfa_sd, enc))
        terms = LO.predictor_loss(occ.reshape(-1, 18), labels.reshape(-1), mask.reshape(-1), cw)
        loss = loss + sum(terms.values())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(leaves, 5.0)
        opt.step()
    finally:
        DO.BN_TRAIN = False
    return time.perf_counter() - t0


CPU_SAMPLE = ('B=1 DHD-S training steps (same stages as the GPU step: dense front, 4-pass view transform, height loss, SFA, '
              'predictor, occupancy losses, backward, grad clip, AdamW); oracle port = reference Python + torch CPU '
              'convs / autograd + C restatement of the bev_pool_v2 forward / grad kernels')


def cpu_reference_leg(cfg, seconds):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_step(cfg, 1, 0, cores)            # warm-up (page in, build the C oracle)
    n, t = 0, 0.0
    while t < seconds and n < 50:
        t += cpu_step(cfg, 1, n + 1, cores)
        n += 1
    return {'value': n / t, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d %s, %.1f s' % (n, CPU_SAMPLE, t)}


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
        'config': {'workload': 'BASELINE configs[1] DHD-S hot path TRAINING step on the host cores (reference '
                               'algorithm: oracle port, the reference ships no CPU kernel for bev_pool_v2); '
                               'each step a bounded sample of B=1 (one 6-camera frame set)'},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d %s' % (steps, CPU_SAMPLE)},
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
