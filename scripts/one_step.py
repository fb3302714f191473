"""A few resident mesh steps (fwd + bwd) for ncu captures.  usage: python scripts/one_step.py [c2|c5] [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import ops, synth

dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, M, S, NF = (8, 20, 400, 100000) if cfg == "c5" else (32, 12, 224, 10000)
meshes = synth.make_meshes(B, NF, 1236)
nv = [v.shape[0] for v, _ in meshes]; nf = [f.shape[0] for _, f in meshes]
verts = torch.cat([v for v, _ in meshes]).to(dev); faces = torch.cat([f for _, f in meshes]).to(dev)
az, el, di = (t.to(dev) for t in (synth.circular_views(B, M) if S == 224 else synth.spherical_views(B, M)))
cot = torch.randn(B * M, 3, S, S, device=dev) / (3 * S * S)
col = torch.tensor([0.99999] * 3, device=dev); light = torch.tensor([[0.0, 1.0, 0.0]], device=dev)
for _ in range(steps):
    a = az.detach().requires_grad_(); e = el.detach().requires_grad_(); d = di.detach().requires_grad_()
    R, T, C, _ = ops._LookAt.apply(a, e, d)
    geom = ops.PackedMeshes.from_packed(verts, faces, nv, nf)
    img, _ = ops.render_meshes(geom, M, R, T, C, light, col, col, S)
    img.backward(cot)
torch.cuda.synchronize()
