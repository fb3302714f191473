"""Per-kernel device times of the resident point step.  usage: python scripts/points_times.py [c1|c3|c5]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import ops, synth
from mvtn_b200 import _lib as L
dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
B, M, S, NP, K = (8, 20, 400, 16384, 4) if cfg == "c5" else ((1, 12, 224, 2048, 1) if cfg == "c1" else (32, 12, 224, 2048, 4))
pts = synth.make_clouds(B, NP, 77).to(dev)
az, el, di = (t.to(dev) for t in (synth.spherical_views(B, M) if cfg == "c5" else synth.learned_spherical_views(B, M, 5)))
cot = torch.randn(B * M, 3, S, S, device=dev) / (3 * S * S)
col = torch.tensor([0.99999] * 3, device=dev); bg = torch.zeros(3, device=dev)
lib = L.load()
def step():
    a = az.detach().requires_grad_(); e = el.detach().requires_grad_(); d = di.detach().requires_grad_()
    img, _, _ = ops.render_points_from_angles(pts, col, M, a, e, d, 0.006, bg, S, points_per_pixel=K, compositor="alpha")
    img.backward(cot)
for _ in range(5): step()
torch.cuda.synchronize()
tot = 0.0
for nm in ("look_at_forward_one_cta_kernel", "pixel_table_kernel", "points_bin_kernel", "points_bin_scan_kernel", "points_tile_kernel", "points_backward_kernel", "points_backward_reduce", "look_at_backward_kernel"):
    lib.mvr_profile_enable(nm.encode())
    for _ in range(5): step()
    t, n = ctypes.c_double(0), ctypes.c_int(0)
    lib.mvr_profile_collect(ctypes.byref(t), ctypes.byref(n))
    tot += t.value / 5
    if n.value: print("%-28s %8.1f us per step (%d launches/step)" % (nm, 1e3 * t.value / 5, n.value // 5))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(30): step()
e1.record(); torch.cuda.synchronize()
print("sum of kernels %.1f us; step %.1f us" % (1e3 * tot, 1e3 * e0.elapsed_time(e1) / 30))
