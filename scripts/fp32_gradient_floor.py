#!/usr/bin/env python
"""What does "gradients within 1e-5 relative" mean for an fp32 rasterizer backward?  (CPU only, test infrastructure.)

PyTorch3D evaluates this chain (rasterize_meshes_backward + autograd of shading / projection) in fp32, on CPU and on GPU.
Here the SAME formulas (oracle/torch_ref.py, autograd) are evaluated twice on identical inputs and identical fragment
indices -- once in fp32, once in fp64 -- and the camera gradients compared.  The difference is pure fp32 rounding of a
legitimate evaluation order: any two correct fp32 implementations (PyTorch3D's CPU kernel, its CUDA kernel, ours) differ from
each other by about this much, so it is the floor under the mesh-gradient tolerance of tests/test_gpu_parity.py.

    python scripts/fp32_gradient_floor.py > profiles/r3_fp32_gradient_floor.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvtn_b200 import ops, synth          # noqa: E402
from oracle import oracle as orc          # noqa: E402
from oracle import torch_ref as tr        # noqa: E402

K00, K11 = ops.fov_projection_scale()


def grads(dtype, v, f, nrm, R, T, C, light, bg, rgb, p2f, gimg, H):
    Rd, Td, Cd = (torch.from_numpy(x).to(dtype).requires_grad_() for x in (R, T, C))
    loss = 0
    for n in range(R.shape[0]):
        img, _ = tr.render_mesh_view(v.to(dtype), f, torch.from_numpy(nrm).to(dtype), torch.from_numpy(rgb).to(dtype), Rd[n], Td[n],
                                     Cd[n], torch.from_numpy(light[0]).to(dtype), torch.from_numpy(bg).to(dtype), K00, K11, H, H,
                                     p2f=torch.from_numpy(p2f[n, ..., 0]).long())
        loss = loss + (img * torch.from_numpy(gimg[n]).to(dtype)).sum()
    loss.backward()
    return [t.grad.double().numpy() for t in (Rd, Td, Cd)]


def main():
    print("case                      faces  H    M   fp32-vs-fp64 autograd: max|d| / max|ref| per tensor (gR, gT, gC) | oracle(fp64 chain) vs fp64 autograd")
    cases = [("small", 300, 32, 4, 1), ("spherical", 2000, 64, 4, 2), ("c2-like", 10000, 224, 2, 12), ("dense_subpixel", 20000, 64, 2, 13)]
    for name, nf, H, M, seed in cases:
        v, f = synth.make_mesh(nf, seed)
        az, el, di = synth.learned_spherical_views(1, M, seed + 7)
        R, T, C = orc.look_at(az.numpy().ravel(), el.numpy().ravel(), di.numpy().ravel())
        nrm = orc.vertex_normals(v.numpy(), f.numpy())
        voff = np.array([0, v.shape[0]], np.int32); foff = np.array([0, f.shape[0]], np.int32)
        light = np.array([[0.3, 1.0, -0.5]], np.float32); bg = np.full(3, 0.99999, np.float32)
        rgb = np.full((v.shape[0], 3), 0.99999, np.float32)
        o = orc.mesh_forward(v.numpy(), f.numpy(), voff, foff, nrm, rgb, M, R, T, C, light, bg, K00, K11, 0.5, H, H, 1,
                             orc.PERSPECTIVE_CORRECT)
        gimg = np.random.RandomState(1).randn(M, 3, H, H).astype(np.float32)
        bw = orc.mesh_backward(v.numpy(), f.numpy(), voff, foff, nrm, rgb, M, R, T, C, light, K00, K11, H, H, 1,
                               orc.PERSPECTIVE_CORRECT, o["pix_to_face"], gimg)
        g64 = grads(torch.float64, v, f, nrm, R, T, C, light, bg, rgb, o["pix_to_face"], gimg, H)
        g32 = grads(torch.float32, v, f, nrm, R, T, C, light, bg, rgb, o["pix_to_face"], gimg, H)
        r32 = [float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(g32, g64)]
        ror = [float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max()) for a, b in zip((bw["gR"], bw["gT"], bw["gC"]), g64)]
        print(f"{name:24s} {f.shape[0]:6d} {H:4d} {M:3d}   " + "  ".join(f"{x:.2e}" for x in r32) + "   |   " + "  ".join(f"{x:.2e}" for x in ror))


if __name__ == "__main__":
    main()
