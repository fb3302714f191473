import os, sys, cProfile, pstats, io
sys.path.insert(0, "/root/repo")
import torch
from mvtn_b200 import MVRenderer, synth
dev = torch.device("cuda:0")
B, M, S, NP = 1, 12, 224, 2048
pts_h = synth.make_clouds(B, NP, 7).pin_memory()
az, el, di = (t.contiguous().pin_memory() for t in synth.circular_views(B, M))
r = MVRenderer(M, image_size=S, pc_rendering=True, points_radius=0.006, points_per_pixel=1).to(dev)
cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
g_host = torch.empty(3, B, M, pin_memory=True)
st = torch.cuda.current_stream()
def step():
    a = az.to(dev, non_blocking=True).requires_grad_(); e = el.to(dev, non_blocking=True).requires_grad_(); d = di.to(dev, non_blocking=True).requires_grad_()
    img, _ = r(None, pts_h, a, e, d)
    img.backward(cot.view_as(img))
    g_host[0].copy_(a.grad, non_blocking=True); g_host[1].copy_(e.grad, non_blocking=True); g_host[2].copy_(d.grad, non_blocking=True)
    st.synchronize()
for _ in range(50): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(500): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45); print(s.getvalue()[:9000])
