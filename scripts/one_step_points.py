"""A few resident point steps (fwd + bwd) for ncu captures.  usage: python scripts/one_step_points.py [c3|c5] [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import ops, synth
dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, M, S, NP, K = (8, 20, 400, 16384, 4) if cfg == "c5" else (32, 12, 224, 2048, 4)
pts = synth.make_clouds(B, NP, 77).to(dev)
az, el, di = (t.to(dev) for t in (synth.spherical_views(B, M) if cfg == "c5" else synth.learned_spherical_views(B, M, 5)))
cot = torch.randn(B * M, 3, S, S, device=dev) / (3 * S * S)
col = torch.tensor([0.99999] * 3, device=dev); bg = torch.zeros(3, device=dev)
for _ in range(steps):
    a = az.detach().requires_grad_(); e = el.detach().requires_grad_(); d = di.detach().requires_grad_()
    img, _, _ = ops.render_points_from_angles(pts, col, M, a, e, d, 0.006, bg, S, points_per_pixel=K, compositor="alpha")
    img.backward(cot)
torch.cuda.synchronize()
