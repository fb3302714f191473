"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: object sharding needs no data-path collective,
timing is reduced with MAX over ranks, and the training configuration's gradient all-reduce is bucketed."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mvtn_b200 import parallel
    r, lr, w = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    # (1) each rank takes a contiguous object range; the union is the batch, no exchange needed
    lo, hi = parallel.shard_range(33, r, w)
    owned = torch.zeros(33); owned[lo:hi] = 1
    dist.all_reduce(owned)
    assert torch.equal(owned, torch.ones(33))
    # (2) timing = max over ranks, throughput = sum of units / that time
    t = parallel.max_over_ranks(10.0 + rank)
    n = parallel.sum_over_ranks(hi - lo)
    assert t == 10.0 + world - 1 and n == 33
    # (3) bucketed gradient all-reduce (mean), several parameters per bucket and an odd one out
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((5, 3), (7,), (1000,), (2, 2))]
    params.append(torch.nn.Parameter(torch.zeros(3)))        # no grad: skipped
    for i, p in enumerate(params[:-1]):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    nb = parallel.allreduce_gradients(params, bucket_bytes=512)
    assert nb >= 2
    for i, p in enumerate(params[:-1]):
        assert torch.allclose(p.grad, torch.full_like(p, (i + 1) * (1 + world) / 2.0))
    # (4) hook-driven all-reduce launched during backward (configs[3]): same result as averaging the full gradients
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))
    unused = torch.nn.Parameter(torch.ones(3))                      # never reaches the loss: finish() must still reduce it
    sync = parallel.OverlappedGradientAllReduce(list(net.parameters()) + [unused], bucket_bytes=128)
    assert sync.enabled and len(sync.buckets) >= 3
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))
    ref.load_state_dict(net.state_dict())
    for step in range(3):
        x = torch.full((3, 6), float(rank + 1 + step))
        if step == 1:
            net.zero_grad(set_to_none=True)                    # a caller that detaches the gradients from the buckets anyway
        else:
            sync.zero_grad()
        net(x).square().sum().backward()
        ref.zero_grad(set_to_none=True)
        ref(x).square().sum().backward()
        assert sync.finish() == len(sync.buckets) and sync.stats["launched_in_backward"] >= len(sync.buckets) - 1
        for p, pr in zip(net.parameters(), ref.parameters()):
            want = pr.grad.clone(); dist.all_reduce(want); want /= world
            assert torch.allclose(p.grad, want, atol=1e-6)
            bi, i = sync._slot[id(p)]
            assert p.grad.data_ptr() == sync.buckets[bi]["views"][i].data_ptr()      # still (or again) a view of its bucket
        assert unused.grad is not None and float(unused.grad.abs().sum()) == 0.0
    sync.remove()
    parallel.barrier()
    dist.destroy_process_group()
    q.put((rank, lo, hi))


def test_object_sharding_and_reductions_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, 0, 17), (1, 17, 33)]


def test_single_process_helpers_are_noops():
    sys.path.insert(0, ROOT)
    from mvtn_b200 import parallel
    assert parallel.max_over_ranks(3.5) == 3.5 and parallel.sum_over_ranks(2) == 2.0
    assert parallel.allreduce_gradients([torch.nn.Parameter(torch.zeros(2))]) == 0
    parallel.barrier()
