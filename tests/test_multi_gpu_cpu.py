"""world_size-2 gloo tests (CPU) of the N>1 plumbing: batch-axis sharding, max-over-ranks timing
aggregation, result gathering -- the host logic bench.py and the plugin use under torchrun."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dhd_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = shard.shard_samples(7, world, rank)
        # each rank "processes" its samples: result row = sample id
        res = torch.tensor(mine, dtype=torch.int64).view(-1, 1)
        got = shard.gather_on_rank0(res)
        times = shard.max_over_ranks([10.0 + rank, 5.0 - rank])
        dist.barrier()
        q.put((rank, mine, None if got is None else got.flatten().tolist(), times))
    finally:
        dist.destroy_process_group()


def test_shard_is_a_partition():
    for world in (1, 2, 4, 8):
        for gb in (1, 7, 32):
            parts = [shard.shard_samples(gb, world, r) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(gb))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_aggregation():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, g0, t0), (r1, m1, g1, t1) = out
    assert m0 == [0, 1, 2, 3] and m1 == [4, 5, 6]
    assert g0 == list(range(7)) and g1 is None
    assert t0 == t1 == [11.0, 5.0]          # element-wise max over the two ranks
    assert shard.samples_per_second(8, 2.0) == 4000.0


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
        b = shard.GradBucket(list(lin.parameters()) + [frozen])
        b.zero()
        lin.weight.grad.add_(float(rank + 1))          # what a backward kernel does: accumulate in place
        lin.bias.grad.add_(10.0 * (rank + 1))
        b.all_reduce_async()
        b.wait()
        q.put((rank, lin.weight.grad.flatten().tolist(), lin.bias.grad.tolist(), b.flat.numel(),
               lin.weight.grad.data_ptr() == b.flat.data_ptr()))
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_single_all_reduce():
    """The data-parallel exchange of the training step: one flat bucket, one SUM all-reduce, mean."""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, w, bias, n, aliased in out:
        assert w == [1.5] * 15 and bias == [15.0] * 3      # mean of (1, 2) and of (10, 20)
        assert n == 18 and aliased


def _span_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        a, b_, c = torch.nn.Linear(4, 2), torch.nn.Linear(3, 3), torch.nn.Linear(2, 5)
        bucket = shard.GradBucket(list(a.parameters()) + list(b_.parameters()) + list(c.parameters()))
        spans = [bucket.span(list(m.parameters())) for m in (a, b_, c)]
        bucket.zero()
        for k, m in enumerate((a, b_, c)):
            for p in m.parameters():
                p.grad.add_(float((rank + 1) * (k + 1)))
        # the training step's order: the LAST modules' gradients are final first and travel while the rest is computed
        bucket.all_reduce_async(*spans[2])
        bucket.all_reduce_async(*spans[1], bf16=True)          # compressed exchange of one segment
        bucket.all_reduce_async(*spans[0])
        bucket.wait()
        try:
            bucket.span([a.weight, c.weight])
            contiguous_error = False
        except ValueError:
            contiguous_error = True
        q.put((rank, spans, [float(m.weight.grad.flatten()[0]) for m in (a, b_, c)], contiguous_error))
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_segmented_overlapped_exchange():
    """Per-segment all-reduces (head + SFA first, then the encoders, then the front) tile the flat buffer exactly and
    give the same mean as the single exchange; a bf16-compressed segment too (small integers are exact in bf16)."""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_span_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, spans, firsts, contiguous_error in out:
        assert spans == [(0, 10), (10, 22), (22, 37)]
        assert firsts == [1.5, 3.0, 4.5]                        # mean over ranks of (rank + 1) * (k + 1)
        assert contiguous_error
