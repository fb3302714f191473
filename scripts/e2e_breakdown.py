"""Host-side breakdown of one end-to-end MVRenderer step (CPU time per phase, no syncs inside, then total)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import MVRenderer, Meshes, ops, synth

dev = torch.device("cuda:0")
B, M, S = 32, 12, 224
meshes = synth.make_meshes(B, 10000, 1236)
ml = [Meshes([v], [f]) for v, f in meshes]
az, el, di = (t.contiguous().pin_memory() for t in synth.circular_views(B, M))
r = MVRenderer(M, image_size=S, pc_rendering=False, light_direction="fixed").to(dev)
cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
g_host = torch.empty(3, B, M, pin_memory=True)

def step(log=None):
    t = [time.perf_counter()]
    a = az.to(dev, non_blocking=True).requires_grad_(); e = el.to(dev, non_blocking=True).requires_grad_(); d = di.to(dev, non_blocking=True).requires_grad_()
    t.append(time.perf_counter())
    geom = ops.PackedMeshes([m.verts_list()[0] for m in ml], [m.faces_list()[0] for m in ml], dev)
    t.append(time.perf_counter())
    img, _ = r(geom, None, a, e, d)
    t.append(time.perf_counter())
    img.backward(cot)
    t.append(time.perf_counter())
    g_host[0].copy_(a.grad, non_blocking=True); g_host[1].copy_(e.grad, non_blocking=True); g_host[2].copy_(d.grad, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    t.append(time.perf_counter())
    if log is not None:
        log.append([t[i + 1] - t[i] for i in range(len(t) - 1)])

for _ in range(5):
    step()
log = []
for _ in range(20):
    step(log)
import statistics
names = ["h2d views", "pack+h2d+prepare (host)", "forward (host)", "backward (host)", "d2h + final sync"]
for i, nme in enumerate(names):
    print("%-28s %.3f ms" % (nme, 1e3 * statistics.median(x[i] for x in log)))
print("total %.3f ms" % (1e3 * statistics.median(sum(x) for x in log)))
# finer: inside PackedMeshes
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
