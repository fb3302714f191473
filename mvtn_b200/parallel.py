"""Object-sharded data parallelism for the rendering path.

Every (object, view) is independent (renderer.py:105,141 replicate per view; no cross-object term), so
rendering needs NO collective: each rank renders a contiguous range of objects.  A collective (NCCL
all-reduce over NVLink) is only needed for the network gradients of the training configuration.
One process per GPU, launched with torchrun; rendezvous through RANK / WORLD_SIZE / MASTER_* env.
"""
import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process => (0, 0, 1))."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend: str = None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def shard_range(num_objects: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous object range [lo, hi) of `rank`: sizes differ by at most one, order preserved."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(num_objects, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_weight(weights: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous partition of ragged objects balancing sum(weights) (e.g. faces or points per object):
    cut i is placed where the prefix sum first reaches i/world of the total."""
    n = len(weights)
    total = float(sum(weights))
    cuts = [0]
    acc, j = 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while j < n and acc + weights[j] / 2.0 <= target:
            acc += weights[j]
            j += 1
        cuts.append(max(j, cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device-timed milliseconds) over all ranks."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_gradients(params, bucket_bytes: int = 32 << 20, average: bool = True):
    """Bucketed gradient all-reduce for the training configuration (BASELINE config 4): the MVTN
    regressor and CNN backbone gradients, ~60 MB fp32, in flat buckets sized for launch latency."""
    if not dist.is_initialized():
        return 0
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    n_buckets, i = 0, 0
    while i < len(grads):
        bucket, size = [], 0
        while i < len(grads) and (not bucket or size + grads[i].numel() * grads[i].element_size() <= bucket_bytes):
            bucket.append(grads[i])
            size += grads[i].numel() * grads[i].element_size()
            i += 1
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat)
        if average:
            flat /= world
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_buckets += 1
    return n_buckets


class OverlappedGradientAllReduce:
    """Hook-driven bucketed gradient all-reduce that runs DURING backward (BASELINE configs[3]: run_mvtn.py:168-224 sharded
    by object, NCCL all-reduce of the MVTN regressor + backbone gradients; SURVEY 8e "bucketed and overlapped with backward").

    Parameters are assigned to flat buckets in REVERSE registration order (the order autograd finishes them in, to first
    order) and their .grad tensors are VIEWS of those buckets, so autograd accumulates straight into the communication buffer:
    no copy in, no copy out.  A post-accumulate-grad hook counts finished gradients; the moment a bucket is complete its
    all-reduce (NCCL: ReduceOp.AVG) is launched asynchronously -- NCCL runs it on its own stream, under the rest of the backward
    pass: the renderer's backward and the view selector's come LAST in the graph, so every backbone bucket is in flight before
    the rasterizer's backward kernels start.  `finish()` -- call it after loss.backward() -- launches whatever is left and
    waits.

        sync = OverlappedGradientAllReduce(params)                           # once
        sync.zero_grad(); loss.backward(); sync.finish(); optimizer.step()   # every step

    Use sync.zero_grad() instead of optimizer.zero_grad(set_to_none=True): it zeroes the buckets (one memset each) and keeps
    the views; a gradient that was detached from its bucket anyway is copied back in and re-bound (slower, still correct).
    Single process / no process group: every call is a no-op (zero_grad falls back to setting grads to None)."""

    def __init__(self, params, bucket_bytes: int = 8 << 20, average: bool = True):
        self.enabled = dist.is_initialized() and dist.get_world_size() > 1
        self.average = average
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # dicts: params, offsets, flat, views, filled, handle
        self._slot = {}
        self._hooks = []
        self.launched_in_backward = 0      # buckets whose all-reduce was launched from a hook (statistics of the last step)
        self.stats = {}
        if not self.enabled:
            return
        self._avg_native = average and dist.get_backend() == "nccl"      # gloo has no ReduceOp.AVG: sum, then divide
        cur, size = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (size + nbytes > bucket_bytes or cur[0].dtype != p.dtype or cur[0].device != p.device):
                self._close(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self._close(cur)
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _close(self, plist):
        offs, o = [], 0
        for p in plist:
            offs.append(o)
            o += p.numel()
        flat = torch.zeros(o, dtype=plist[0].dtype, device=plist[0].device)
        views = [flat[offs[i]: offs[i] + p.numel()].view_as(p) for i, p in enumerate(plist)]
        b = {"params": list(plist), "offsets": offs, "numel": o, "flat": flat, "views": views, "handle": None, "filled": set()}
        for i, p in enumerate(plist):
            self._slot[id(p)] = (len(self.buckets), i)
            p.grad = views[i]
        self.buckets.append(b)

    def zero_grad(self):
        """Zero every gradient in place (one memset per bucket) and keep them bound to the buckets."""
        if not self.enabled:
            for p in self.params:
                p.grad = None
            return
        for b in self.buckets:
            b["flat"].zero_()
            for i, p in enumerate(b["params"]):
                if p.grad is not b["views"][i]:
                    p.grad = b["views"][i]

    def _launch(self, b):
        op = dist.ReduceOp.AVG if self._avg_native else dist.ReduceOp.SUM
        b["handle"] = dist.all_reduce(b["flat"], op=op, async_op=True)

    def _on_grad(self, p):
        bi, i = self._slot[id(p)]
        b = self.buckets[bi]
        if i in b["filled"]:
            return
        v = b["views"][i]
        if p.grad is not v and p.grad.data_ptr() != v.data_ptr():      # detached from the bucket (zero_grad(set_to_none=True)): copy in
            v.copy_(p.grad)
            p.grad = v
        b["filled"].add(i)
        if len(b["filled"]) == len(b["params"]):
            self._launch(b)
            self.launched_in_backward += 1

    def finish(self):
        """Launch the buckets that did not complete during backward (parameters without a gradient this step: their slots
        hold whatever zero_grad left there), wait for all of them and, without a native average, divide.  Returns the number
        of buckets."""
        if not self.enabled:
            return 0
        world = dist.get_world_size()
        for b in self.buckets:
            if b["handle"] is None:
                for i, p in enumerate(b["params"]):       # unfilled slots: the parameter had no gradient on this rank
                    if i not in b["filled"] and p.grad is not b["views"][i]:
                        b["views"][i].zero_()
                        p.grad = b["views"][i]
                self._launch(b)
        for b in self.buckets:
            b["handle"].wait()
            if self.average and not self._avg_native:
                b["flat"].div_(world)
            b["handle"] = None
            b["filled"] = set()
        n = len(self.buckets)
        self.stats = {"buckets": n, "launched_in_backward": self.launched_in_backward,
                      "reduce_op": "avg (nccl)" if self._avg_native else "sum + divide"}
        self.launched_in_backward = 0
        return n

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
