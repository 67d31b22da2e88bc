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
