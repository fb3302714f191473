"""C5-scale sanity + timing: 100k-face meshes / 16k-point clouds, 20 views, 400x400 (BASELINE configs[4])."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
from mvtn_b200 import ops, synth

dev = torch.device("cuda:0")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B, M, S = 4, 20, 400
meshes = synth.make_meshes(B, 100000, 1240)
print("faces", [f.shape[0] for _, f in meshes])
az, el, di = synth.spherical_views(B, M)
R, T, C = orc.look_at(az.numpy().ravel(), el.numpy().ravel(), di.numpy().ravel())
Rd, Td, Cd = (torch.from_numpy(x).to(dev) for x in (R, T, C))
geom = ops.PackedMeshes([v for v, _ in meshes], [f for _, f in meshes], dev)
col = torch.full((3,), 0.99999, device=dev); light = torch.tensor([[0, 1.0, 0]], device=dev)
img, frag = ops.render_meshes(geom, M, Rd, Td, Cd, light, col, col, S)
print("mesh coverage", float((frag["pix_to_face"] >= 0).float().mean()), "counters", frag["counters"].tolist())
# oracle on 2 views of object 1
b = 1; s = slice(b * M, b * M + 2)
vp = meshes[b][0].numpy(); fp = meshes[b][1].numpy().astype(np.int32)
k00, k11 = ops.fov_projection_scale()
t0 = time.time()
o = orc.mesh_forward(vp, fp, [0, vp.shape[0]], [0, fp.shape[0]], orc.vertex_normals(vp, fp), np.full(3, 0.99999, np.float32), 2, R[s], T[s], C[s],
                     np.array([[0, 1.0, 0]], np.float32), np.full(3, 0.99999, np.float32), k00, k11, 0.5, S, S, 1, orc.PERSPECTIVE_CORRECT, fragments=False)
print("oracle 2 views %.1fs" % (time.time() - t0), "idx mismatches", int((frag["pix_to_face"][s].cpu().numpy() != o["pix_to_face"]).sum()),
      "img err", float(np.abs(img[s].cpu().numpy() - o["images"]).max()))
g = torch.randn(B * M, 3, S, S, device=dev)
def fwd():
    ops.render_meshes(geom, M, Rd, Td, Cd, light, col, col, S)
def fwdbwd():
    Rg = Rd.clone().requires_grad_(); Tg = Td.clone().requires_grad_(); Cg = Cd.clone().requires_grad_()
    im, _ = ops.render_meshes(geom, M, Rg, Tg, Cg, light, col, col, S); im.backward(g)
tf, tb = timeit(fwd), timeit(fwdbwd)
print("C5 mesh: %d views fwd %.3f ms (%.0f views/s)  fwd+bwd %.3f ms (%.0f views/s)" % (B * M, tf, B * M / tf * 1e3, tb, B * M / tb * 1e3))
# points
pts = synth.make_clouds(B, 16384, 1241).to(dev)
inv = (1.0 / di.reshape(-1)).to(dev)
imgp, fragp = ops.render_points(pts, col, M, Rd, Td, inv, 0.006, col * 0, S)
op = orc.points_forward(pts[b:b+1].cpu().numpy(), np.full(3, 0.99999, np.float32), 2, R[s], T[s], inv[s].cpu().numpy(), 0.006, np.zeros(3, np.float32), S, S, 1, 0, fragments=False)
print("points coverage", float((fragp["idx"] >= 0).float().mean()), "idx mismatches", int((fragp["idx"][s].cpu().numpy() != op["idx"]).sum()))
def pf():
    ops.render_points(pts, col, M, Rd, Td, inv, 0.006, col * 0, S)
tp = timeit(pf)
print("C5 points: %d views fwd %.3f ms (%.0f views/s)" % (B * M, tp, B * M / tp * 1e3))
