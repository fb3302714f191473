"""cProfile of the end-to-end MVRenderer step (host side), sorted by own time: where the Python microseconds go."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import MVRenderer, Meshes, synth

dev = torch.device("cuda:0")
B, M, S = 32, 12, 224
ml = [Meshes([v], [f]) for v, f in synth.make_meshes(B, 10000, 1236)]
az, el, di = (t.contiguous().pin_memory() for t in synth.circular_views(B, M))
r = MVRenderer(M, image_size=S, pc_rendering=False, light_direction="fixed").to(dev)
cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
g_host = torch.empty(3, B, M, pin_memory=True)

def step():
    a = az.to(dev, non_blocking=True).requires_grad_(); e = el.to(dev, non_blocking=True).requires_grad_(); d = di.to(dev, non_blocking=True).requires_grad_()
    img, _ = r(ml, None, a, e, d)
    t1 = time.perf_counter()
    img.backward(cot)
    g_host[0].copy_(a.grad, non_blocking=True); g_host[1].copy_(e.grad, non_blocking=True); g_host[2].copy_(d.grad, non_blocking=True)
    t2 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    return t1, t2

for _ in range(10):
    step()
ts = []
for _ in range(30):
    t0 = time.perf_counter(); t1, t2 = step(); t3 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
import statistics
print("host forward %.3f ms | host backward+d2h enqueue %.3f ms | final wait %.3f ms | total %.3f ms" % tuple(1e3 * statistics.median(x[i] for x in ts) for i in range(4)))
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(35)
