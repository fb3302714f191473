#!/usr/bin/env python
"""BASELINE config 4: a synthetic re-creation of MVTN's training step (run_mvtn.py:168-224) around the B200 renderer.

    view selector (PointNet-lite + MLP -> azim/elev offsets, models/mvtn.py:223-248)
      -> mvtn_b200.MVRenderer (mesh path, Phong)                        <- the hot path of this repo
      -> MVCNN: ResNet-18 on (B*M,3,H,W), max over views, Linear         (models/multi_view.py:54-70)
      -> CrossEntropy, backward (through the renderer into the selector), AdamW x 2

Objects shard by rank (32 per GPU by default); rendering needs no collective, the network gradients are
all-reduced in buckets over NCCL DURING backward (mvtn_b200.parallel.OverlappedGradientAllReduce: hook-driven, the buckets
of the backbone are in flight while the renderer's backward kernels run).

    python examples/train_step.py --steps 5
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_step.py --steps 5
"""
import argparse
import os
import sys
import time

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvtn_b200 import MVRenderer, Meshes, parallel, regualarize_rendered_views, synth  # noqa: E402


class ViewSelector(nn.Module):
    """Learned-spherical selector: spherical grid + tanh-bounded offsets predicted from the point cloud."""

    def __init__(self, nb_views, canonical_distance=2.2):
        super().__init__()
        az, el, di = synth.spherical_views(1, nb_views, canonical_distance)
        self.register_buffer("azim", az[0]); self.register_buffer("elev", el[0]); self.register_buffer("dist", di[0])
        self.nb_views = nb_views
        self.encoder = nn.Sequential(nn.Conv1d(3, 64, 1), nn.ReLU(), nn.Conv1d(64, 256, 1), nn.ReLU())
        self.head = nn.Sequential(nn.Linear(256 + 2 * nb_views, 256), nn.ReLU(), nn.Linear(256, 2 * nb_views), nn.Tanh())

    def forward(self, points):
        B = points.shape[0]
        feat = self.encoder(points.transpose(1, 2)).max(dim=2).values
        az, el, di = (t.expand(B, -1) for t in (self.azim, self.elev, self.dist))
        off = self.head(torch.cat([feat, az / 180.0, el / 90.0], dim=1))
        return az + off[:, :self.nb_views] * 180.0 / self.nb_views, (el + off[:, self.nb_views:] * 90.0).clamp(-89, 89), di


class MVCNN(nn.Module):
    def __init__(self, nb_classes=40):
        super().__init__()
        import torchvision
        net = torchvision.models.resnet18(weights=None)
        net.fc = nn.Identity()
        self.net = net
        self.fc = nn.Sequential(nn.LayerNorm(512), nn.Linear(512, nb_classes))

    def forward(self, images):                      # (B, M, 3, H, W)
        B, M = images.shape[:2]
        f = self.net(images.reshape(B * M, *images.shape[2:])).view(B, M, -1)
        return self.fc(f.max(dim=1).values)


class TrainStep:
    """One rank's share of the step: `batch` objects x `views` views.  step() returns the loss tensor (no host sync).
    sync: "overlap" (hook-driven buckets during backward), "after" (parallel.allreduce_gradients once backward is done:
    the un-overlapped baseline) or "none"."""

    def __init__(self, dev, rank, batch=32, views=12, image_size=224, faces=10000, amp=False, sync="overlap", collate=True,
                 view_reg=0.0, augment_training=False, crop_ratio=0.3):
        from mvtn_b200 import collate_meshes
        torch.manual_seed(1234)                         # same initial weights on every rank
        self.dev, self.amp, self.sync_mode = dev, amp, sync
        self.view_reg, self.augment_training, self.crop_ratio = view_reg, augment_training, crop_ratio      # config.yaml:42-44
        self.selector, self.cnn = ViewSelector(views).to(dev), MVCNN().to(dev)
        self.renderer = MVRenderer(views, image_size=image_size, pc_rendering=False, light_direction="random").to(dev)
        self.opt = torch.optim.AdamW(self.cnn.parameters(), lr=1e-3, weight_decay=0.01)
        self.opt_mvtn = torch.optim.AdamW(self.selector.parameters(), lr=1e-4, weight_decay=0.01)
        meshes = [Meshes([v], [f]) for v, f in synth.make_meshes(batch, faces, 4000 + 97 * rank)]   # this rank's objects
        self.points = torch.stack([m.verts_list()[0][torch.randperm(m.verts_list()[0].shape[0])[:2048]] for m in meshes]).to(dev)
        self.meshes = collate_meshes(meshes) if collate else meshes      # what the loader's collate_fn hands over
        self.targets = torch.randint(0, 40, (batch,), generator=torch.Generator().manual_seed(rank)).to(dev)
        self.params = list(self.cnn.parameters()) + list(self.selector.parameters())
        self.crit = nn.CrossEntropyLoss()
        self.overlap = parallel.OverlappedGradientAllReduce(self.params) if sync == "overlap" else None
        self.render_events = None

    def grad_bytes(self):
        return sum(p.numel() * p.element_size() for p in self.params)

    def step(self):
        azim, elev, dist = self.selector(self.points)
        if self.render_events is not None:
            self.render_events[0].record()
        images, _ = self.renderer(self.meshes, None, azim, elev, dist)
        if self.render_events is not None:
            self.render_events[1].record()
        # run_mvtn.py:186-187 (a no-op at the reference's defaults: view_reg 0, augment_training false)
        images = regualarize_rendered_views(images, self.view_reg, self.augment_training, self.crop_ratio)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            loss = self.crit(self.cnn(images), self.targets)
        if self.overlap is not None:
            self.overlap.zero_grad()              # gradients stay views of the communication buckets
        else:
            self.opt.zero_grad(set_to_none=True); self.opt_mvtn.zero_grad(set_to_none=True)
        loss.backward()
        if self.overlap is not None:
            self.overlap.finish()
        elif self.sync_mode == "after":
            parallel.allreduce_gradients(self.params)
        self.opt.step(); self.opt_mvtn.step()
        return loss

    def close(self):
        if self.overlap is not None:
            self.overlap.remove()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32, help="objects per GPU")
    ap.add_argument("--views", type=int, default=12)
    ap.add_argument("--image-size", type=int, default=224)
    ap.add_argument("--faces", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the CNN (the renderer stays fp32)")
    ap.add_argument("--sync", default="overlap", choices=["overlap", "after", "none"])
    a = ap.parse_args()
    rank, local_rank, world = parallel.init_distributed()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ts = TrainStep(dev, rank, a.batch, a.views, a.image_size, a.faces, a.amp, a.sync)
    t0 = None
    for step in range(a.steps + 2):
        if step == 2:
            torch.cuda.synchronize(); parallel.barrier(); t0 = time.time()
        loss = ts.step()
    torch.cuda.synchronize(); parallel.barrier()
    dt = (time.time() - t0) / a.steps
    g = sum(p.grad.abs().sum().item() for p in ts.selector.parameters() if p.grad is not None)
    if rank == 0:
        print(f"world {world}: {a.batch * world} objects x {a.views} views / step, {dt * 1e3:.1f} ms/step, "
              f"{a.batch * world * a.views / dt:.0f} views/s end to end (render + ResNet-18 fwd/bwd + all-reduce [{a.sync}]); "
              f"loss {loss.item():.3f}, |grad| into the view selector {g:.3e}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
