"""Development check on a GPU box: CUDA path vs CPU oracle on small seeded inputs; prints, never asserts."""
import os, sys, time, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
from mvtn_b200 import ops, synth, MVRenderer, Meshes
from mvtn_b200 import _lib as L

dev = torch.device("cuda:0")
def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

def section(name):
    print("\n=== " + name, flush=True)

def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()

def t_lookat():
    section("look_at")
    az = (torch.rand(64) * 360 - 180); el = torch.rand(64) * 170 - 85; di = torch.rand(64) * 2 + 1.2
    az[:4] = torch.tensor([0., 90., -90., 180.]); el[:4] = 0; di[:4] = 2.2
    R, T, C, bad = ops._LookAt.apply(az.to(dev), el.to(dev), di.to(dev))
    Ro, To, Co = orc.look_at(az.numpy(), el.numpy(), di.numpy())
    print("R", np.abs(R.cpu().numpy() - Ro).max(), "T", np.abs(T.cpu().numpy() - To).max(), "C", np.abs(C.cpu().numpy() - Co).max(), "bad", int(bad))
    el2 = el.clone(); el2[5] = 90.0
    _, _, _, bad = ops._LookAt.apply(az.to(dev), el2.to(dev), di.to(dev))
    Ro2, _, _ = orc.look_at(az.numpy(), el2.numpy(), di.numpy())
    print("degenerate: gpu bad", int(bad), "oracle bad", orc.count_invalid_rotations(Ro2))
    gR = torch.randn(64, 3, 3); gT = torch.randn(64, 3); gC = torch.randn(64, 3)
    a = az.to(dev).requires_grad_(); e = el.to(dev).requires_grad_(); d = di.to(dev).requires_grad_()
    R, T, C, _ = ops._LookAt.apply(a, e, d)
    ((R * gR.to(dev)).sum() + (T * gT.to(dev)).sum() + (C * gC.to(dev)).sum()).backward()
    ga, ge, gd = orc.look_at_backward(az.numpy(), el.numpy(), di.numpy(), gR.numpy(), gT.numpy(), gC.numpy())
    print("bwd rel: azim", rel(a.grad.cpu(), ga), "elev", rel(e.grad.cpu(), ge), "dist", rel(d.grad.cpu(), gd))

def mesh_case(name, meshes, M, H, K, views, persp=True, cull=False, vert_rgb=None, light=None, bwd=True):
    section("mesh " + name)
    verts = [m[0] for m in meshes]; faces = [m[1] for m in meshes]
    B = len(verts)
    az, el, di = views
    geom = ops.PackedMeshes(verts, faces, dev, vert_rgb=vert_rgb)
    R, T, C, bad = ops._LookAt.apply(az.reshape(-1).to(dev), el.reshape(-1).to(dev), di.reshape(-1).to(dev))
    N = B * M
    if light is None:
        light = torch.tensor([[0.3, 1.0, -0.5]])
    obj = torch.tensor([0.99999, 0.99999, 0.99999]); bg = torch.tensor([0.99999, 0.99999, 0.99999]) * 0.5
    Rg = R.detach().clone().requires_grad_(); Tg = T.detach().clone().requires_grad_(); Cg = C.detach().clone().requires_grad_()
    t0 = time.time()
    img, frag = ops.render_meshes(geom, M, Rg, Tg, Cg, light.to(dev), None if vert_rgb is not None else obj.to(dev), bg.to(dev), H,
                                  faces_per_pixel=K, cull_backfaces=cull, perspective_correct=persp, fragments=True)
    torch.cuda.synchronize()
    print("gpu fwd ok %.3fs" % (time.time() - t0), "counters", frag["counters"].tolist())
    # oracle
    vp = torch.cat(verts).numpy(); fp = torch.cat(faces).numpy().astype(np.int32)
    voff = np.array(geom.vert_off_host, np.int32); foff = np.array(geom.face_off_host, np.int32)
    nrm = orc.packed_vertex_normals(vp, fp, voff, foff)
    print("normals max err", np.abs(geom.vertex_normals().cpu().numpy() - nrm).max())
    k00, k11 = ops.fov_projection_scale()
    flags = (orc.PERSPECTIVE_CORRECT if persp else 0) | (orc.CULL_BACKFACES if cull else 0)
    rgb = obj.numpy() if vert_rgb is None else vert_rgb.numpy()
    t0 = time.time()
    o = orc.mesh_forward(vp, fp, voff, foff, nrm, rgb, M, R.cpu().numpy(), T.cpu().numpy(), C.cpu().numpy(), light.numpy(), bg.numpy(),
                         k00, k11, 0.5 if persp else -1.0, H, H, K, flags)
    print("oracle fwd %.2fs" % (time.time() - t0), "straddle", o["straddle"])
    p2f = frag["pix_to_face"].cpu().numpy()
    mism = (p2f != o["pix_to_face"])
    print("coverage", (o["pix_to_face"][..., 0] >= 0).mean(), "idx mismatches", int(mism.sum()), "of", mism.size)
    if mism.sum():
        w = np.argwhere(mism)[:5]
        for i in w:
            print("   at", i, "gpu", p2f[tuple(i)], "orc", o["pix_to_face"][tuple(i)], "zg", frag["zbuf"].cpu().numpy()[tuple(i)], "zo", o["zbuf"][tuple(i)])
    ok = ~mism
    zb = frag["zbuf"].cpu().numpy(); ba = frag["bary_coords"].cpu().numpy(); dd = frag["dists"].cpu().numpy()
    print("zbuf exact mism", int((zb[ok] != o["zbuf"][ok]).sum()), "bary exact mism", int((ba[ok] != o["bary"][ok]).sum()), "dists max err", np.abs(dd[ok] - o["dists"][ok]).max())
    okpix = ok[..., 0]
    im = img.detach().cpu().numpy(); 
    err = np.abs(im - o["images"]).transpose(0, 2, 3, 1)[okpix]
    print("image max abs err", err.max() if err.size else 0.0)
    if bwd:
        g = torch.randn(N, 3, H, H, generator=torch.Generator().manual_seed(5))
        (img * g.to(dev)).sum().backward()
        ob = orc.mesh_backward(vp, fp, voff, foff, nrm, rgb, M, R.cpu().numpy(), T.cpu().numpy(), C.cpu().numpy(), light.numpy(), k00, k11, H, H, K, flags,
                               p2f, g.numpy())
        print("bwd rel: gR", rel(Rg.grad.cpu(), ob["gR"]), "gT", rel(Tg.grad.cpu(), ob["gT"]), "gC", rel(Cg.grad.cpu(), ob["gC"]))

def points_case(name, B, Np, M, H, K, radius, mode, views, per_point=False):
    section("points " + name)
    pts = synth.make_clouds(B, Np, 11)
    az, el, di = views
    R, T, C, bad = ops._LookAt.apply(az.reshape(-1).to(dev), el.reshape(-1).to(dev), di.reshape(-1).to(dev))
    inv = (1.0 / di.reshape(-1))
    rgb = torch.rand(B, Np, 3, generator=torch.Generator().manual_seed(3)) if per_point else torch.tensor([0.99999] * 3)
    bg = torch.tensor([0.1, 0.2, 0.3])
    Rg = R.detach().clone().requires_grad_(); Tg = T.detach().clone().requires_grad_(); sg = inv.to(dev).requires_grad_()
    pg = pts.to(dev).requires_grad_(); fg = rgb.to(dev).requires_grad_()
    img, frag = ops.render_points(pg, fg, M, Rg, Tg, sg, radius, bg.to(dev), H, points_per_pixel=K, compositor=mode, fragments=True)
    torch.cuda.synchronize()
    flags = orc.COMPOSITE_ALPHA if mode == "alpha" else 0
    o = orc.points_forward(pts.numpy(), rgb.numpy(), M, R.cpu().numpy(), T.cpu().numpy(), inv.numpy(), radius, bg.numpy(), H, H, K, flags)
    idx = frag["idx"].cpu().numpy()
    mism = idx != o["idx"]
    print("coverage", (o["idx"][..., 0] >= 0).mean(), "idx mismatches", int(mism.sum()))
    ok = ~mism
    print("zbuf exact mism", int((frag["zbuf"].cpu().numpy()[ok] != o["zbuf"][ok]).sum()), "d2 exact mism", int((frag["dists"].cpu().numpy()[ok] != o["dists2"][ok]).sum()))
    print("image max abs err", np.abs(img.detach().cpu().numpy() - o["images"]).max())
    g = torch.randn(B * M, 3, H, H, generator=torch.Generator().manual_seed(6))
    (img * g.to(dev)).sum().backward()
    ob = orc.points_backward(pts.numpy(), rgb.numpy(), M, R.cpu().numpy(), T.cpu().numpy(), inv.numpy(), radius, H, H, K, flags, idx, g.numpy(), want_points=True, want_rgb=True)
    print("bwd rel: gR", rel(Rg.grad.cpu(), ob["gR"]), "gT", rel(Tg.grad.cpu(), ob["gT"]), "gs", rel(sg.grad.cpu(), ob["g_inv_dist"]),
          "gP", rel(pg.grad.cpu(), ob["grad_points"]), "gF", rel(fg.grad.cpu(), ob["grad_rgb"]))

def t_renderer():
    section("MVRenderer end-to-end")
    meshes = synth.make_meshes(2, 500, 21)
    ml = [Meshes([v], [f]) for v, f in meshes]
    az, el, di = synth.circular_views(2, 4)
    r = MVRenderer(4, image_size=64, pc_rendering=False, light_direction="fixed").cuda()
    a = az.to(dev).requires_grad_(); e = el.to(dev).requires_grad_(); d = di.to(dev).requires_grad_()
    img, cams = r(ml, None, a, e, d)
    print("mesh images", tuple(img.shape), float(img.min()), float(img.max()), float(img.mean()))
    img.square().mean().backward()
    print("grads", a.grad.abs().max().item(), e.grad.abs().max().item(), d.grad.abs().max().item())
    print("camera centers", cams.get_camera_center()[0].tolist())
    r = MVRenderer(4, image_size=64, pc_rendering=True, points_radius=0.02, points_per_pixel=3, background_color="black", compositor="alpha").cuda()
    pts = synth.make_clouds(2, 512, 5)
    a = az.to(dev).requires_grad_(); e = el.to(dev).requires_grad_(); d = di.to(dev).requires_grad_()
    img, cams = r(None, pts, a, e, d)
    print("point images", tuple(img.shape), float(img.min()), float(img.max()), float(img.mean()))
    img.square().mean().backward()
    print("grads", a.grad.abs().max().item(), e.grad.abs().max().item(), d.grad.abs().max().item())

run(t_lookat)
v4 = synth.circular_views(1, 4)
run(lambda: mesh_case("1x300f 4v 32px K=1", synth.make_meshes(1, 300, 1), 4, 32, 1, v4))
run(lambda: mesh_case("2x2000f 4v 64px K=1 spherical", synth.make_meshes(2, 2000, 2), 4, 64, 1, synth.learned_spherical_views(2, 4, 9)))
run(lambda: mesh_case("ragged 3 meshes K=3 50px", [synth.make_mesh(200, 4), synth.make_mesh(1200, 5), synth.make_mesh(60, 6)], 2, 50, 3, synth.learned_spherical_views(3, 2, 1)))
cube_v = torch.tensor([[-1,-1,-1],[1,-1,-1],[1,1,-1],[-1,1,-1],[-1,-1,1],[1,-1,1],[1,1,1],[-1,1,1]], dtype=torch.float32) * 0.55
cube_f = torch.tensor([[0,2,1],[0,3,2],[4,5,6],[4,6,7],[0,1,5],[0,5,4],[2,3,7],[2,7,6],[1,2,6],[1,6,5],[0,4,7],[0,7,3]])
run(lambda: mesh_case("cube big faces 96px K=2", [(cube_v, cube_f)], 4, 96, 2, synth.learned_spherical_views(1, 4, 3)))
run(lambda: mesh_case("cull + no persp", synth.make_meshes(1, 800, 7), 3, 40, 1, synth.learned_spherical_views(1, 3, 4), persp=False, cull=True))
m8 = synth.make_meshes(1, 600, 8)
run(lambda: mesh_case("per-vertex rgb + relative light", m8, 3, 48, 1, synth.learned_spherical_views(1, 3, 5), vert_rgb=torch.rand(m8[0][0].shape[0], 3)))
run(lambda: mesh_case("10k faces 224px 2 views", synth.make_meshes(1, 10000, 12), 2, 224, 1, synth.circular_views(1, 2), bwd=True))
run(lambda: points_case("K=1 norm r=.006 224", 1, 2048, 3, 224, 1, 0.006, "norm", synth.circular_views(1, 3)))
run(lambda: points_case("K=4 alpha r=.05 64 per-point", 2, 500, 3, 64, 4, 0.05, "alpha", synth.learned_spherical_views(2, 3, 2), per_point=True))
run(lambda: points_case("K=3 norm r=.04 50", 2, 300, 2, 50, 3, 0.04, "norm", synth.learned_spherical_views(2, 2, 3), per_point=True))
run(lambda: points_case("K=1 alpha r=.02 100", 1, 1000, 2, 100, 1, 0.02, "alpha", synth.circular_views(1, 2)))
run(t_renderer)
print("\nDONE")
