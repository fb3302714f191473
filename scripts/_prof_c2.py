import os, sys, cProfile, pstats, io
sys.path.insert(0, "/root/repo")
import torch
from mvtn_b200 import MVRenderer, Meshes, synth, collate_meshes
dev = torch.device("cuda:0")
B, M, S, NF = 32, 12, 224, 10000
ml = [Meshes([v], [f]) for v, f in synth.make_meshes(B, NF, 1236)]
host = collate_meshes(ml)
az, el, di = (t.contiguous().pin_memory() for t in synth.circular_views(B, M))
r = MVRenderer(M, image_size=S, pc_rendering=False, light_direction="fixed", h2d_chunks=1).to(dev).train()
cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
g_host = torch.empty(3, B, M, pin_memory=True)
st = torch.cuda.current_stream()
def fwd():
    a = az.to(dev, non_blocking=True).requires_grad_(); e = el.to(dev, non_blocking=True).requires_grad_(); d = di.to(dev, non_blocking=True).requires_grad_()
    img, _ = r(host, None, a, e, d)
    return img, a, e, d
def step():
    img, a, e, d = fwd()
    img.backward(cot)
    g_host[0].copy_(a.grad, non_blocking=True); g_host[1].copy_(e.grad, non_blocking=True); g_host[2].copy_(d.grad, non_blocking=True)
    st.synchronize()
for _ in range(20): step()
pr = cProfile.Profile()
for _ in range(300):
    pr.enable(); out = fwd(); pr.disable()
    out[0].backward(cot); st.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(60); print(s.getvalue()[:12000])
