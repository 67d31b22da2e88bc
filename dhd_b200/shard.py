"""Multi-GPU plumbing of the hot path: the batch-sample axis is the only shard axis (a sample's
voxel reduction sums over its 6 cameras); one process per GPU, no data-path
collective.  The only collectives are bookkeeping: the max-over-ranks of the timed regions and
(optionally) gathering the per-rank class maps on rank 0, like mmdet's multi_gpu_test
(tools/test.py:267)."""
import torch
import torch.distributed as dist


def shard_samples(global_batch, world, rank):
    """Contiguous, balanced split of sample indices [0, global_batch) over `world` ranks."""
    if not 0 <= rank < world:
        raise ValueError('rank %d outside world %d' % (rank, world))
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def max_over_ranks(values, device='cpu'):
    """Element-wise MAX of a list of floats over all ranks (timing aggregation)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def samples_per_second(samples_all_ranks, ms):
    return samples_all_ranks / (ms * 1e-3)


def gather_on_rank0(t):
    """Concatenate per-rank results along dim 0 on rank 0 (None elsewhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != 0:
        return None
    return torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)], dim=0)


class GradBucket:
    """Flat fp32 gradient buffer of a parameter list: every `param.grad` is a view into it, so the
    backward kernels accumulate straight into the bucket and the data-parallel reduction is ONE
    all-reduce of one contiguous tensor (the reference: torch DDP's bucketed all-reduce over NCCL,
    tools/train.py:195 -> mmdet3d train_model -> MMDistributedDataParallel).  `all_reduce_async`
    launches it on the process group's stream (NCCL: overlaps whatever is enqueued afterwards) and
    `wait` joins it and divides by the world size."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev, n = self.params[0].device, sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view(p.shape)
            o += p.numel()
        self._works = []

    def zero(self):
        self.flat.zero_()

    def flatten_params(self):
        """Move every parameter's storage into ONE flat fp32 buffer (same order as the gradients): `p.data` becomes a
        view, modules keep their Parameter objects.  Lets the optimizer run as a single kernel over flat buffers."""
        if getattr(self, 'flat_params', None) is None:
            if any(p.dtype != torch.float32 for p in self.params):
                raise ValueError('flat parameters need fp32 master weights')
            self.flat_params = torch.empty_like(self.flat)
            o = 0
            for p in self.params:
                v = self.flat_params[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                o += p.numel()
        return self.flat_params


    def span(self, params):
        """[lo, hi) of the flat buffer covered by `params` (a contiguous run of this bucket's parameter list)."""
        ids = {id(p) for p in params if p.requires_grad}
        lo = hi = None
        o = 0
        for p in self.params:
            if id(p) in ids:
                lo = o if lo is None else lo
                hi = o + p.numel()
            o += p.numel()
        if lo is None:
            return 0, 0
        if sum(p.numel() for p in self.params if id(p) in ids) != hi - lo:
            raise ValueError('the parameters are not contiguous in the bucket')
        return lo, hi

    def all_reduce_async(self, lo=None, hi=None, bf16=False):
        """Launch the all-reduce of flat[lo:hi] (default: everything) on the process group's stream; it starts when the
        kernels enqueued so far have finished and overlaps whatever is enqueued afterwards (the rest of the backward).
        bf16=True: the slice travels as bf16 (half the bytes; torch DDP's bf16 compression hook) and is widened again
        in wait()."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        lo = 0 if lo is None else lo
        hi = self.flat.numel() if hi is None else hi
        if hi <= lo:
            return
        sl = self.flat[lo:hi]
        if bf16:
            buf = sl.to(torch.bfloat16)
            self._works.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True), sl, buf))
        else:
            self._works.append((dist.all_reduce(sl, op=dist.ReduceOp.SUM, async_op=True), sl, None))

    def wait(self):
        if not self._works:
            return
        inv = 1.0 / dist.get_world_size()
        for work, sl, buf in self._works:
            work.wait()
            if buf is not None:
                sl.copy_(buf)
            sl.mul_(inv)
        self._works = []


class FlatAdamW:
    """torch.optim.AdamW (decoupled weight decay; the reference's optimizer, DHD-S.py:262) as ONE kernel over the flat
    parameter / gradient buffers of a GradBucket (dhd_adamw_flat), instead of torch's multi-tensor launches over ~80
    tensors.  `step(grad_scale=t)` multiplies the gradient by the device scalar t on the fly (gradient clipping).
    `param_groups[0]` carries lr / betas / eps / weight_decay like a torch optimizer (schedulers write lr there)."""

    def __init__(self, bucket, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.bucket = bucket
        self.p = bucket.flatten_params()
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.t = 0
        self.param_groups = [dict(params=bucket.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]

    def step(self, grad_scale=None):
        import ctypes
        from . import _lib
        g = self.param_groups[0]
        self.t += 1
        b1, b2 = g['betas']
        P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        _lib.check(_lib.load().dhd_adamw_flat(P(self.p), P(self.bucket.flat), P(self.m), P(self.v), self.p.numel(),
                                              g['lr'], b1, b2, g['eps'], g['weight_decay'], 1.0 - b1 ** self.t,
                                              1.0 - b2 ** self.t, P(grad_scale),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'adamw_flat')
        # the kernel wrote the parameters through raw pointers: bump their version counters like an in-place torch op
        # would, so everything keyed on them (the compiled inference engines, dhd_b200.compat.EngineOwner) sees the step
        try:
            torch._C._increment_version(self.bucket.params)
        except (AttributeError, TypeError):             # no such hook in this torch: drop the engines explicitly instead
            from .compat import EngineOwner
            for m in getattr(self, 'modules', ()):
                if isinstance(m, EngineOwner):
                    m.invalidate()

    def state_dict(self):
        return dict(step=self.t, exp_avg=self.m, exp_avg_sq=self.v, param_groups=[{k: v for k, v in self.param_groups[0].items() if k != 'params'}])

    def load_state_dict(self, sd):
        self.t = int(sd['step'])
        self.m.copy_(sd['exp_avg'])
        self.v.copy_(sd['exp_avg_sq'])
        self.param_groups[0].update(sd['param_groups'][0])
