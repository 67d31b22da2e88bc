"""BASELINE configs[3] ("DHD-B": DHD-S topology at 6 x 384x1056) and configs[4] (DHD-L: 6 x 512x1408, two temporal frames)
on N GPUs -- one JSON line each, same timing rules as bench.py (barrier + synchronize, CUDA events, max over ranks, weak
scaling: every rank runs its own batch, training steps end in one NCCL gradient all-reduce).  Usage:
  python scripts/bench_configs.py                      (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_configs.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dhd_b200 import shard, synth  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    st = torch.cuda.current_stream()
    which = sys.argv[1:] or ['dhd_b', 'dhd_l']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, seconds, warm=3, kmin=5):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(st)
        for _ in range(kmin):
            fn()
        e1.record(st)
        barrier()
        ms = shard.max_over_ranks([e0.elapsed_time(e1) / kmin], device='cuda')[0]
        n = max(kmin, int(seconds * 1e3 / ms))
        barrier()
        e0.record(st)
        for _ in range(n):
            fn()
        e1.record(st)
        barrier()
        return shard.max_over_ranks([e0.elapsed_time(e1)], device='cuda')[0] / n, n

    def timed_infer(step, img_inputs):
        """Eager detector inference, then the same step as one CUDA graph (DetectorStep.capture_infer) when it captures."""
        ms_e, n_e = timed(lambda: step.infer_step(img_inputs), 1.0, warm=2, kmin=3)
        res = {'eager_ms_per_step': ms_e}
        if step.capture_infer(img_inputs):
            import numpy as np
            same = all(np.array_equal(a, b) for a, b in zip(step.infer_step(img_inputs), step.infer_step_graphed(img_inputs)))
            ms_g, n_g = timed(lambda: step.infer_step_graphed(), 1.0, warm=2, kmin=3)
            res.update(ms_per_step=ms_g, timed_passes=n_g, cuda_graph=True, graph_equals_eager=bool(same))
        else:
            res.update(ms_per_step=ms_e, timed_passes=n_e, cuda_graph=False, capture_error=step.capture_error[:200])
        return res

    out = []
    if 'dhd_b' in which:
        from dhd_b200.pipeline import HotPathStep, TrainStep
        cfg, B = synth.DHD_B, 4
        ts = TrainStep(cfg, B, bn='batch')
        host = ts.make_host_inputs(synth.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=100 + rank), seed=100 + rank)
        ts.alloc_static(host)
        ts.upload(host)
        ts.train_step()
        graphed = ts.capture_train()
        ms_t, n_t = timed(ts.train_step, 2.0)
        del ts
        torch.cuda.empty_cache()
        hp = HotPathStep(cfg, B, precision='bf16')
        hi = {k: host[k] for k in ('x', 'sensor2ego', 'ego2global', 'cam2imgs', 'post_rots', 'post_trans', 'bda')}
        hp.alloc_static(hi)
        hp.upload(hi)
        hp.capture()
        ms_i, n_i = timed(hp.run, 1.0)
        del hp
        torch.cuda.empty_cache()
        out.append({'config': 'BASELINE configs[3] "DHD-B": DHD-S topology at 6-cam 384x1056 (24x66 features), 200x200x16 grid, bf16, '
                              'batch 4 per GPU (the R101 image backbone is outside the hot path: synthetic image features)',
                    'n_gpus': world, 'unit': 'samples/s', 'scaling': 'weak',
                    'train_step': {'value': world * B / (ms_t * 1e-3), 'ms_per_step': ms_t, 'timed_passes': n_t, 'cuda_graph': bool(graphed),
                                   'what': 'hot-path training step (as bench.py value): fwd + losses + bwd + NCCL all-reduce + clip + AdamW'},
                    'inference': {'value': world * B / (ms_i * 1e-3), 'ms_per_step': ms_i, 'timed_passes': n_i,
                                  'what': 'HotPathStep: image features -> uint8 class map'}})
    for tag, depth, size, label in (('dhd_s_images', 50, (256, 704), 'DHD-S from camera images: ResNet-50 + CustomFPN (DHD-S.py:44-62) + the '
                                     'whole detector'),
                                    ('dhd_b_images', 101, (384, 1056), 'BASELINE configs[3] "DHD-B" from camera images: ResNet-101 at '
                                     '6-cam 384x1056 + CustomFPN + the whole detector')):
        if tag not in which:
            continue
        from dhd_b200.detector_step import DetectorStep
        B = 4
        step = DetectorStep(synth.dhd_s_model_cfg('bf16', images=True, input_size=size, depth=depth), B, seed=rank)
        img_inputs, kw = step.make_inputs(300 + rank)
        losses = step.train_step(img_inputs, kw)
        ms_t, n_t = timed(lambda: step.train_step(img_inputs, kw), 2.0, warm=2, kmin=3)
        inf = timed_infer(step, img_inputs)
        out.append({'config': label + ', 200x200x16 grid, bf16, batch 4 per GPU', 'n_gpus': world, 'unit': 'samples/s',
                    'scaling': 'weak',
                    'train_step': {'value': world * B / (ms_t * 1e-3), 'ms_per_step': ms_t, 'timed_passes': n_t, 'cuda_graph': False,
                                   'trainable_params': step.n_params, 'losses': {k: float(v) for k, v in losses.items()},
                                   'what': 'DHD.forward_train (DHD_model.py:135-186) on camera IMAGES through the plugin detector in '
                                           'train() mode: image backbone + neck with batch-statistics BatchNorm (trained, as DHD-S.py '
                                           'does), view transformer, BEV / voxel encoders, SFA, head, four losses, backward through '
                                           'every module, ONE NCCL gradient all-reduce, clip 5, AdamW; eager launches'},
                    'inference': dict(inf, value=world * B / (inf['ms_per_step'] * 1e-3),
                                      what='DHD.simple_test: camera images -> list of (200, 200, 16) uint8 class maps (incl. D2H)')})
        del step
        torch.cuda.empty_cache()
    if 'dhd_l' in which:
        from dhd_b200.detector_step import DetectorStep
        B = 2
        step = DetectorStep(synth.dhd_l_model_cfg('bf16'), B, seed=rank)
        img_inputs, kw = step.make_inputs(200 + rank)
        losses = step.train_step(img_inputs, kw)
        ms_t, n_t = timed(lambda: step.train_step(img_inputs, kw), 2.0, warm=2, kmin=3)
        inf = timed_infer(step, img_inputs)
        out.append({'config': 'BASELINE configs[4] DHD-L: 6-cam 512x1408 (32x88x512 image features, 128-ch stereo features at 128x352), '
                              'D=88, two temporal frames + stereo reference frame, 200x200x16 grid, bf16, batch 2 per GPU '
                              '(DHD-L.py samples_per_gpu; Swin-B + FPN are outside the hot path: synthetic per-frame features)',
                    'n_gpus': world, 'unit': 'samples/s', 'scaling': 'weak',
                    'train_step': {'value': world * B / (ms_t * 1e-3), 'ms_per_step': ms_t, 'timed_passes': n_t, 'cuda_graph': False,
                                   'trainable_params': step.n_params, 'losses': {k: float(v) for k, v in losses.items()},
                                   'what': 'DHD_stereo.forward_train (DHD_model.py:577-614) through the plugin detector: key frame with '
                                           'gradients (stereo cost volume, camera-aware DepthNet + HeightNet, fused pool fwd/bwd, '
                                           'pre-process nets, BEV encoder, three UNets, SFA, head, five losses), previous frame under '
                                           'no_grad, backward, ONE NCCL gradient all-reduce, clip 5, AdamW; eager launches'},
                    'inference': dict(inf, value=world * B / (inf['ms_per_step'] * 1e-3),
                                      what='DHD_stereo.simple_test (bf16 NHWC activation path): per-frame features -> list of '
                                           '(200, 200, 16) uint8 class maps (incl. D2H)')})
    if rank == 0:
        for o in out:
            print(json.dumps(o), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
