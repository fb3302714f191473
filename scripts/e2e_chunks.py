"""End-to-end mesh step (collated pinned host batch -> H2D -> render -> backward -> D2H of the view gradients -> sync, every
step) as a function of MVRenderer(h2d_chunks=k).  usage: python scripts/e2e_chunks.py [c2|c5]"""
import os, sys, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import MVRenderer, Meshes, synth, collate_meshes

dev = torch.device("cuda:0")
B, M, S, NF = 32, 12, 224, 10000
if len(sys.argv) > 1 and sys.argv[1] == "c5":
    B, M, S, NF = 8, 20, 400, 100000
ml = [Meshes([v], [f]) for v, f in synth.make_meshes(B, NF, 1236)]
host = collate_meshes(ml)
az, el, di = (t.contiguous().pin_memory() for t in (synth.circular_views(B, M) if S == 224 else synth.spherical_views(B, M)))
cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
g_host = torch.empty(3, B, M, pin_memory=True)
st = torch.cuda.current_stream()
for k in (1, 2, 3, 4, 8, None):
    r = MVRenderer(M, image_size=S, pc_rendering=False, light_direction="fixed", h2d_chunks=k).to(dev).train()
    def step():
        a = az.to(dev, non_blocking=True).requires_grad_(); e = el.to(dev, non_blocking=True).requires_grad_(); d = di.to(dev, non_blocking=True).requires_grad_()
        img, _ = r(host, None, a, e, d)
        img.backward(cot)
        g_host[0].copy_(a.grad, non_blocking=True); g_host[1].copy_(e.grad, non_blocking=True); g_host[2].copy_(d.grad, non_blocking=True)
        st.synchronize()
    for _ in range(10): step()
    ts = []
    for _ in range(40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("h2d_chunks=%-4s  e2e %.4f ms / step (median; min %.4f)  grad checksum %.6e" % (k, statistics.median(ts), min(ts), float(g_host.double().abs().sum())))
