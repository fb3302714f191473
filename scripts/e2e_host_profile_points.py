"""Host-side timeline of one end-to-end point-cloud step (BASELINE configs[0]: one 2048-point cloud x 12 views, 224^2) through
MVRenderer from pinned host tensors: where the ~0.3 ms of a launch-bound step go.  perf_counter only, no profiler.
usage: python scripts/e2e_host_profile_points.py [B]"""
import os, sys, time, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import MVRenderer, synth

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M, S, NP = 12, 224, 2048
pts_h = synth.make_clouds(B, NP, 7).pin_memory()
az, el, di = (t.contiguous().pin_memory() for t in synth.circular_views(B, M))
views_h = torch.stack([az, el, di]).pin_memory()
r = MVRenderer(M, image_size=S, pc_rendering=True, points_radius=0.006, points_per_pixel=1).to(dev)
cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
g_host = torch.empty(3, B, M, pin_memory=True)
st = torch.cuda.current_stream()

def step_three(rec):
    t0 = time.perf_counter()
    a = az.to(dev, non_blocking=True).requires_grad_(); e = el.to(dev, non_blocking=True).requires_grad_(); d = di.to(dev, non_blocking=True).requires_grad_()
    t1 = time.perf_counter()
    img, _ = r(None, pts_h, a, e, d)
    t2 = time.perf_counter()
    img.backward(cot.view_as(img))
    t3 = time.perf_counter()
    g_host[0].copy_(a.grad, non_blocking=True); g_host[1].copy_(e.grad, non_blocking=True); g_host[2].copy_(d.grad, non_blocking=True)
    t4 = time.perf_counter()
    st.synchronize()
    t5 = time.perf_counter()
    if rec is not None: rec.append([1e6 * (x - t0) for x in (t1, t2, t3, t4, t5)])

def step_one(rec):      # the three view tensors as rows of ONE pinned tensor: one H2D, one gradient tensor, one D2H
    t0 = time.perf_counter()
    v = views_h.to(dev, non_blocking=True).requires_grad_()
    a, e, d = v.unbind(0)
    t1 = time.perf_counter()
    img, _ = r(None, pts_h, a, e, d)
    t2 = time.perf_counter()
    img.backward(cot.view_as(img))
    t3 = time.perf_counter()
    g_host.copy_(v.grad, non_blocking=True)
    t4 = time.perf_counter()
    st.synchronize()
    t5 = time.perf_counter()
    if rec is not None: rec.append([1e6 * (x - t0) for x in (t1, t2, t3, t4, t5)])

names = ["views H2D enqueued", "forward returned", "backward returned", "D2H enqueued", "synchronized"]
for label, step in (("three view tensors", step_three), ("one (3,B,M) view tensor", step_one)):
    for _ in range(20): step(None)
    rec = []
    for _ in range(200): step(rec)
    print("==", label, "| graph:", getattr(r, "_graph_state", None) is not None)
    for i, n in enumerate(names):
        print("%-22s at %8.1f us" % (n, statistics.median(x[i] for x in rec)))
