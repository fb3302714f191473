"""GPU-side timeline of one end-to-end MVRenderer step: CUDA events between the phases + host timestamps, so the
idle gaps (host-bound stretches) show up as the difference between event time and summed kernel time."""
import os, sys, time, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import MVRenderer, Meshes, ops, synth

dev = torch.device("cuda:0")
B, M, S = 32, 12, 224
ml = [Meshes([v], [f]) for v, f in synth.make_meshes(B, 10000, 1236)]
vs = [m.verts_list()[0] for m in ml]; fs = [m.faces_list()[0] for m in ml]
az, el, di = (t.contiguous().pin_memory() for t in synth.circular_views(B, M))
r = MVRenderer(M, image_size=S, pc_rendering=False, light_direction="fixed").to(dev)
cot = torch.randn(B, M, 3, S, S, device=dev) / (3 * S * S)
g_host = torch.empty(3, B, M, pin_memory=True)
st = torch.cuda.current_stream()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(st); return e

def step():
    marks = [("start", ev(), time.perf_counter())]
    a = az.to(dev, non_blocking=True).requires_grad_(); e = el.to(dev, non_blocking=True).requires_grad_(); d = di.to(dev, non_blocking=True).requires_grad_()
    geom = ops.PackedMeshes(vs, fs, dev)
    marks.append(("stage+h2d+prepare", ev(), time.perf_counter()))
    img, _ = r(geom, None, a, e, d)
    marks.append(("look_at+forward", ev(), time.perf_counter()))
    img.backward(cot)
    marks.append(("backward", ev(), time.perf_counter()))
    g_host[0].copy_(a.grad, non_blocking=True); g_host[1].copy_(e.grad, non_blocking=True); g_host[2].copy_(d.grad, non_blocking=True)
    marks.append(("d2h", ev(), time.perf_counter()))
    st.synchronize()
    t_end = time.perf_counter()
    return marks, t_end

for _ in range(10):
    step()
rows = []
for _ in range(30):
    marks, t_end = step()
    t0 = marks[0][2]
    rows.append([(name, marks[0][1].elapsed_time(e), 1e3 * (t - t0)) for name, e, t in marks[1:]] + [("sync", None, 1e3 * (t_end - t0))])
print("%-22s %12s %12s" % ("phase end", "gpu ms", "host ms"))
for i in range(len(rows[0])):
    g = [r_[i][1] for r_ in rows if r_[i][1] is not None]
    h = [r_[i][2] for r_ in rows]
    print("%-22s %12s %12.3f" % (rows[0][i][0], "%.3f" % statistics.median(g) if g else "-", statistics.median(h)))
# raw H2D rate of the staged bytes
buf = torch.empty(5763072 // 4, dtype=torch.float32).pin_memory(); dst = torch.empty_like(buf, device=dev)
for _ in range(3): dst.copy_(buf, non_blocking=True)
e0 = ev(); 
for _ in range(10): dst.copy_(buf, non_blocking=True)
e1 = ev(); st.synchronize()
print("H2D 5.76 MB pinned: %.3f ms each (%.1f GB/s)" % (e0.elapsed_time(e1) / 10, 5.763072e-3 / (e0.elapsed_time(e1) / 10) ))
