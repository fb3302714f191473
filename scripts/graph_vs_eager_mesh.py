"""End-to-end mesh step (collated pinned host batch -> H2D -> render -> backward -> sync) through MVRenderer with cuda_graph=False / True
as a function of the batch size: where replaying the captured step pays.  usage: python scripts/graph_vs_eager_mesh.py"""
import os, sys, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import MVRenderer, Meshes, synth, collate_meshes

dev = torch.device("cuda:0")
M, S, NF = 12, 224, 10000
for B in (1, 2, 4, 8, 16, 32):
    ml = [Meshes([v], [f]) for v, f in synth.make_meshes(B, NF, 1236)]
    host = collate_meshes(ml)
    az, el, di = (t.contiguous().pin_memory() for t in synth.circular_views(B, M))
    cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
    st = torch.cuda.current_stream()
    row = []
    for graph in (False, True):
        r = MVRenderer(M, image_size=S, pc_rendering=False, light_direction="fixed", cuda_graph=graph).to(dev).train()
        def step():
            a = az.to(dev, non_blocking=True).requires_grad_(); e = el.to(dev, non_blocking=True).requires_grad_(); d = di.to(dev, non_blocking=True).requires_grad_()
            img, _ = r(host, None, a, e, d)
            img.backward(cot)
            g = a.grad.cpu()
            return g
        for _ in range(8): step()
        ts = []
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row.append(statistics.median(ts))
    print("B = %2d (%3d views): eager %.3f ms   graph replay %.3f ms   (%.0f / %.0f k views/s)" % (B, B * M, row[0], row[1], B * M / row[0], B * M / row[1]))
