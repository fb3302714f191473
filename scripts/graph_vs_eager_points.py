"""Point step (forward + backward, resident inputs) through MVRenderer with cuda_graph=False / True at a given batch: where the
automatic switch (GRAPH_AUTO_MAX_VIEWS) should sit.  usage: python scripts/graph_vs_eager_points.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import MVRenderer, synth
dev = torch.device("cuda:0")
M, S, NP = 12, 224, 2048
for B in (1, 4, 8, 16, 32, 64):
    pts = synth.make_clouds(B, NP, 3).to(dev)
    az, el, di = (t.to(dev) for t in synth.learned_spherical_views(B, M, 5))
    cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
    row = []
    for graph in (False, True):
        r = MVRenderer(M, image_size=S, pc_rendering=True, points_per_pixel=4, compositor="alpha", background_color="black", cuda_graph=graph).to(dev)
        def step():
            a, e, d = (t.detach().requires_grad_() for t in (az, el, di))
            img, _ = r(None, pts, a, e, d)
            img.backward(cot)
        for _ in range(10): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): step()
        e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 50)
    print("B = %3d (%4d views): eager %.3f ms   graph replay %.3f ms" % (B, B * M, row[0], row[1]))
