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
    order).  A post-accumulate-grad hook copies each finished gradient into its bucket slot; the moment a bucket is complete
    its all-reduce is launched asynchronously (NCCL runs it on its own stream, under the rest of the backward pass: the
    renderer's backward and the view selector's come LAST in the graph, so every backbone bucket is in flight before the
    rasterizer's backward kernels start).  `finish()` -- call it after loss.backward() -- launches whatever is left, waits,
    averages and scatters the buckets back into the .grad tensors.

        sync = OverlappedGradientAllReduce(params)         # once
        loss.backward(); sync.finish(); optimizer.step()   # every step

    Single process / no process group: every call is a no-op."""

    def __init__(self, params, bucket_bytes: int = 8 << 20, average: bool = True):
        self.enabled = dist.is_initialized() and dist.get_world_size() > 1
        self.average = average
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # dicts: params, offsets, flat (lazily), pending, handle
        self._slot = {}
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
        self._hooks = []
        if self.enabled:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.launched_in_backward = 0      # buckets whose all-reduce was launched from a hook (statistics of the last step)

    def _close(self, plist):
        offs, o = [], 0
        for p in plist:
            offs.append(o)
            o += p.numel()
        b = {"params": list(plist), "offsets": offs, "numel": o, "flat": None, "pending": len(plist), "handle": None,
             "filled": set()}
        for i, p in enumerate(plist):
            self._slot[id(p)] = (len(self.buckets), i)
        self.buckets.append(b)

    def _flat(self, b):
        if b["flat"] is None:
            p0 = b["params"][0]
            b["flat"] = torch.zeros(b["numel"], dtype=p0.dtype, device=p0.device)
        return b["flat"]

    def _on_grad(self, p):
        bi, i = self._slot[id(p)]
        b = self.buckets[bi]
        if i in b["filled"]:          # a parameter used twice in the graph: its hook fires once, after accumulation; be safe
            return
        flat = self._flat(b)
        flat[b["offsets"][i]: b["offsets"][i] + p.numel()].copy_(p.grad.reshape(-1))
        b["filled"].add(i)
        if len(b["filled"]) == len(b["params"]):
            b["handle"] = dist.all_reduce(flat, async_op=True)
            self.launched_in_backward += 1

    def finish(self):
        """Launch the buckets that did not complete during backward (parameters without a gradient this step), wait for
        all of them, average, and copy the reduced values back into the .grad tensors.  Returns the number of buckets."""
        if not self.enabled:
            return 0
        world = dist.get_world_size()
        for b in self.buckets:
            if b["handle"] is None:
                flat = self._flat(b)
                for i, p in enumerate(b["params"]):       # unfilled slots: the parameter had no gradient on this rank
                    if i not in b["filled"]:
                        flat[b["offsets"][i]: b["offsets"][i] + p.numel()].zero_()
                b["handle"] = dist.all_reduce(flat, async_op=True)
        for b in self.buckets:
            b["handle"].wait()
            flat = b["flat"]
            if self.average:
                flat.div_(world)
            for i, p in enumerate(b["params"]):
                src = flat[b["offsets"][i]: b["offsets"][i] + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = src.clone()
                else:
                    p.grad.copy_(src)
            b["handle"] = None
            b["filled"] = set()
        n = len(self.buckets)
        self.stats = {"buckets": n, "launched_in_backward": self.launched_in_backward}
        self.launched_in_backward = 0
        return n

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
