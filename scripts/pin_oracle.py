#!/usr/bin/env python
"""Pin the CPU oracle against PyTorch3D itself.  Run on a machine that HAS PyTorch3D (this build container does not:
no network, not vendored -- DESIGN.md section 2):

    python scripts/pin_oracle.py                # compare, print a report, exit 1 on any violation
    python scripts/pin_oracle.py --write        # also store PyTorch3D's outputs under tests/golden/pytorch3d_*.npz

`--write` turns "parity unpinned" into "pinned": tests/test_oracle_vs_pytorch3d.py checks the oracle against the stored
vectors on every machine afterwards (and against live PyTorch3D where it is importable).  The inputs are the seeded
synthetic configurations of mvtn_b200.synth; PyTorch3D is driven exactly as models/renderer.py drives it
(renderer.py:65-151): look_at_view_transform -> FoV cameras -> Meshes.extend(M) / Pointclouds.extend(M).scale_() ->
MeshRasterizer(bin_size=0 on CPU: RasterizeMeshesNaiveCpu) + HardPhongShader / PointsRasterizer + compositor, on CPU
tensors, fp32.  Gradients go to (azim, elev, dist) through PyTorch3D's own autograd.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (kind, kwargs)
    "mesh_12v": ("mesh", dict(nfaces=2000, M=12, S=96, K=1, seed=1236, view_kind="circular")),
    "mesh_k3": ("mesh", dict(nfaces=600, M=4, S=64, K=3, seed=77, view_kind="spherical")),
    "mesh_close": ("mesh", dict(nfaces=1500, M=3, S=96, K=2, seed=61, view_kind="close")),
    "points_12v": ("points", dict(N=2048, M=12, S=224, K=1, seed=1235, radius=0.006, compositor="norm")),
    "points_alpha_k4": ("points", dict(N=2048, M=6, S=128, K=4, seed=1237, radius=0.02, compositor="alpha")),
}


def case_inputs(name):
    from mvtn_b200 import synth
    kind, kw = CASES[name]
    M = kw["M"]
    if kind == "mesh":
        v, f = synth.make_mesh(kw["nfaces"], kw["seed"])
        if kw["view_kind"] == "circular":
            az, el, di = synth.circular_views(1, M)
        elif kw["view_kind"] == "spherical":
            az, el, di = synth.learned_spherical_views(1, M, kw["seed"])
        else:      # faces cross the near clip plane z = znear / 2 ([upstream] clip.py)
            az = torch.tensor([[15.0, 140.0, -80.0]]); el = torch.tensor([[10.0, -35.0, 50.0]]); di = torch.tensor([[1.12, 1.2, 1.3]])
        return dict(kind=kind, verts=v, faces=f, views=(az, el, di), **kw)
    pts = synth.make_clouds(1, kw["N"], kw["seed"])
    az, el, di = synth.learned_spherical_views(1, M, kw["seed"] + 1)
    return dict(kind=kind, points=pts, views=(az, el, di), **kw)


def cotangent(shape, seed=5):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ------------------------------------------------------------------------------------------------ PyTorch3D
def run_pytorch3d(inp):
    """The reference's own arithmetic.  Returns numpy arrays in the oracle's layouts (view-local indices, images (N,3,H,W))."""
    from pytorch3d.renderer import (AlphaCompositor, BlendParams, DirectionalLights, FoVOrthographicCameras,
                                    FoVPerspectiveCameras, HardPhongShader, MeshRasterizer, MeshRenderer,
                                    NormWeightedCompositor, PointsRasterizationSettings, PointsRasterizer, PointsRenderer,
                                    RasterizationSettings, TexturesVertex, look_at_view_transform)
    from pytorch3d.structures import Meshes, Pointclouds
    az, el, di = (t.clone().reshape(-1).requires_grad_() for t in inp["views"])
    M, S, K = inp["M"], inp["S"], inp["K"]
    R, T = look_at_view_transform(dist=di, elev=el, azim=az)
    out = {"R": R.detach().numpy(), "T": T.detach().numpy()}
    if inp["kind"] == "mesh":
        verts, faces = inp["verts"], inp["faces"]
        col = torch.full((1, verts.shape[0], 3), 0.99999)
        meshes = Meshes(verts=[verts], faces=[faces], textures=TexturesVertex(col)).extend(M)
        cameras = FoVPerspectiveCameras(R=R, T=T, znear=1.0, zfar=100.0, fov=60.0)
        rs = RasterizationSettings(image_size=S, blur_radius=0.0, faces_per_pixel=K, bin_size=0, cull_backfaces=False)
        lights = DirectionalLights(direction=((0.0, 1.0, 0.0),))
        rasterizer = MeshRasterizer(cameras=cameras, raster_settings=rs)
        frag = rasterizer(meshes)
        renderer = MeshRenderer(rasterizer=rasterizer, shader=HardPhongShader(
            cameras=cameras, lights=lights, blend_params=BlendParams(background_color=(0.5, 0.25, 0.75))))
        images = renderer(meshes, cameras=cameras, lights=lights)[..., :3].permute(0, 3, 1, 2)
        F = faces.shape[0]
        p2f = frag.pix_to_face.clone()
        local = torch.where(p2f >= 0, p2f - (torch.arange(M) * F).view(M, 1, 1, 1), p2f)
        out.update(index=local.numpy().astype(np.int32), zbuf=frag.zbuf.detach().numpy(),
                   bary=frag.bary_coords.detach().numpy(), dists=frag.dists.detach().numpy(),
                   normals=meshes[0].verts_normals_packed().numpy(), C=cameras.get_camera_center().detach().numpy())
    else:
        pts = inp["points"][0]
        feats = torch.full_like(pts, 0.99999)
        pc = Pointclouds(points=[pts], features=[feats]).extend(M)
        pc.scale_((1.0 / di)[:, None].expand(M, 3))
        cameras = FoVOrthographicCameras(R=R, T=T, znear=0.01)
        rs = PointsRasterizationSettings(image_size=S, radius=inp["radius"], points_per_pixel=K, bin_size=0)
        rasterizer = PointsRasterizer(cameras=cameras, raster_settings=rs)
        frag = rasterizer(pc)
        comp = (AlphaCompositor if inp["compositor"] == "alpha" else NormWeightedCompositor)(background_color=(0.0, 0.0, 0.0))
        images = PointsRenderer(rasterizer=rasterizer, compositor=comp)(pc).permute(0, 3, 1, 2)
        Np = pts.shape[0]
        idx = frag.idx.clone()
        local = torch.where(idx >= 0, idx - (torch.arange(M) * Np).view(M, 1, 1, 1), idx)
        out.update(index=local.numpy().astype(np.int32), zbuf=frag.zbuf.detach().numpy(), dists=frag.dists.detach().numpy())
    g = cotangent(images.shape)
    images.backward(g)
    out.update(images=images.detach().numpy(), g_azim=az.grad.numpy(), g_elev=el.grad.numpy(), g_dist=di.grad.numpy())
    return out


# ------------------------------------------------------------------------------------------------ oracle
def run_oracle(inp, cams=None):
    """The same configuration through oracle/mvr_oracle.c.  cams = (R, T, C) to rasterize from GIVEN cameras (PyTorch3D's:
    stage A of the parity protocol -- fragments are compared from identical cameras); None = the oracle's own look_at."""
    from mvtn_b200 import ops
    from oracle import oracle as orc
    az, el, di = (t.reshape(-1).numpy() for t in inp["views"])
    M, S, K = inp["M"], inp["S"], inp["K"]
    R0, T0, C0 = orc.look_at(az, el, di)
    R, T, C = (R0, T0, C0) if cams is None else cams
    out = {"R": R0, "T": T0, "C": C0}
    if inp["kind"] == "mesh":
        vp = inp["verts"].numpy(); fp = inp["faces"].numpy().astype(np.int32)
        voff = np.array([0, vp.shape[0]], np.int32); foff = np.array([0, fp.shape[0]], np.int32)
        nrm = orc.vertex_normals(vp, fp)
        k00, k11 = ops.fov_projection_scale()
        col = np.full(3, 0.99999, np.float32); light = np.array([[0, 1.0, 0]], np.float32)
        bg = np.array([0.5, 0.25, 0.75], np.float32)
        o = orc.mesh_forward(vp, fp, voff, foff, nrm, col, M, R, T, C, light, bg, k00, k11, 0.5, S, S, K, orc.PERSPECTIVE_CORRECT)
        g = cotangent(o["images"].shape).numpy()
        b = orc.mesh_backward(vp, fp, voff, foff, nrm, col, M, R, T, C, light, k00, k11, S, S, K, orc.PERSPECTIVE_CORRECT,
                              o["pix_to_face"], g)
        ga, ge, gd = orc.look_at_backward(az, el, di, b["gR"], b["gT"], b["gC"])
        out.update(index=o["pix_to_face"], zbuf=o["zbuf"], bary=o["bary"], dists=o["dists"], images=o["images"], normals=nrm,
                   g_azim=ga, g_elev=ge, g_dist=gd)
    else:
        pts = inp["points"].numpy()
        col = np.full(3, 0.99999, np.float32)
        inv = (1.0 / di).astype(np.float32)
        flags = orc.COMPOSITE_ALPHA if inp["compositor"] == "alpha" else 0
        o = orc.points_forward(pts, col, M, R, T, inv, inp["radius"], np.zeros(3, np.float32), S, S, K, flags)
        g = cotangent(o["images"].shape).numpy()
        b = orc.points_backward(pts, col, M, R, T, inv, inp["radius"], S, S, K, flags, o["idx"], g)
        ga, ge, gd = orc.look_at_backward(az, el, di, b["gR"], b["gT"], np.zeros_like(b["gT"]))
        gd = gd + b["g_inv_dist"] * (-1.0 / (di * di))       # scale = 1 / dist
        out.update(index=o["idx"], zbuf=o["zbuf"], dists=o["dists2"], images=o["images"], g_azim=ga, g_elev=ge, g_dist=gd)
    return out


# ------------------------------------------------------------------------------------------------ comparison
TOL = {"look_at_abs": 2e-6, "images_abs": 1e-5, "frag_rel": 1e-5, "grad_rel": 1e-4}


def compare(ref, inp):
    """ref: PyTorch3D outputs (live or stored).  Returns (report dict, list of violations).
    Stage A -- fragments, images from PyTorch3D's OWN cameras (bit-exact indices, 1e-5 elsewhere);
    stage B -- the oracle's look_at against PyTorch3D's; gradients end to end at 1e-4 (both chains are fp32 / fp64 mixes)."""
    C_ref = ref.get("C")
    if C_ref is None:      # orthographic path: the centre is not used
        C_ref = np.zeros((ref["R"].shape[0], 3), np.float32)
    mine = run_oracle(inp, cams=(ref["R"], ref["T"], C_ref))
    rep, bad = {}, []

    def rel(a, b):
        return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))

    rep["index_mismatches"] = int((mine["index"] != ref["index"]).sum())
    if rep["index_mismatches"]:
        bad.append(f"{rep['index_mismatches']} fragment indices differ")
    hit = ref["index"] >= 0
    for k in ("zbuf", "dists") + (("bary",) if "bary" in ref else ()):
        m = hit if ref[k].ndim == hit.ndim else hit[..., None] & np.ones_like(ref[k], bool)
        same = mine["index"] == ref["index"]
        m = m & (same if ref[k].ndim == hit.ndim else same[..., None])
        rep[k + "_max_abs"] = float(np.abs(mine[k][m] - ref[k][m]).max()) if m.any() else 0.0
        rep[k + "_bit_exact"] = bool((mine[k][m] == ref[k][m]).all())
        if rep[k + "_max_abs"] > TOL["frag_rel"] * max(1.0, float(np.abs(ref[k][m]).max()) if m.any() else 1.0):
            bad.append(f"{k} differs by {rep[k + '_max_abs']:.3g}")
    rep["images_max_abs"] = float(np.abs(mine["images"] - ref["images"]).max())
    if rep["images_max_abs"] > TOL["images_abs"]:
        bad.append(f"images differ by {rep['images_max_abs']:.3g}")
    if "normals" in ref:
        rep["normals_max_abs"] = float(np.abs(mine["normals"] - ref["normals"]).max())
        if rep["normals_max_abs"] > 1e-6:
            bad.append(f"vertex normals differ by {rep['normals_max_abs']:.3g}")
    rep["look_at_max_abs"] = max(float(np.abs(mine["R"] - ref["R"]).max()), float(np.abs(mine["T"] - ref["T"]).max()))
    if rep["look_at_max_abs"] > TOL["look_at_abs"]:
        bad.append(f"look_at differs by {rep['look_at_max_abs']:.3g}")
    g_m = np.concatenate([mine[k] for k in ("g_azim", "g_elev", "g_dist")])
    g_r = np.concatenate([ref[k] for k in ("g_azim", "g_elev", "g_dist")])
    rep["grad_views_rel"] = rel(g_m, g_r)
    if rep["grad_views_rel"] > TOL["grad_rel"]:
        bad.append(f"view gradients differ by {rep['grad_views_rel']:.3g} (relative)")
    return rep, bad


def fixture_path(name):
    return os.path.join(GOLD, f"pytorch3d_{name}.npz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true", help="store PyTorch3D's outputs as tests/golden/pytorch3d_<case>.npz")
    ap.add_argument("--cases", nargs="*", default=list(CASES))
    a = ap.parse_args()
    try:
        import pytorch3d
    except ImportError:
        sys.exit("pytorch3d is not importable here: run this script where it is installed (see the docstring)")
    failed = False
    for name in a.cases:
        inp = case_inputs(name)
        ref = run_pytorch3d(inp)
        rep, bad = compare(ref, inp)
        print(name, rep)
        for b in bad:
            print("  VIOLATION:", b)
            failed = True
        if a.write:
            np.savez_compressed(fixture_path(name), pytorch3d_version=pytorch3d.__version__, torch_version=torch.__version__, **ref)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
