"""ctypes front-end of the CPU oracle (``oracle/mvr_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package ``mvtn_b200`` never does.

PARITY UNPINNED (see the header of mvr_oracle.c): PyTorch3D is absent, the reference has no tests.

All arrays are numpy, C-contiguous; float32 / int32 unless stated.  Functions mirror the C-ABI of
``include/mvr_b200.h`` one to one, with host pointers.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmvr_oracle.so")

PERSPECTIVE_CORRECT = 1
CULL_BACKFACES = 2
COMPOSITE_ALPHA = 4
RGB_PER_ELEMENT = 8


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc -O2 -ffp-contract=off -fopenmp)."""
    src = os.path.join(_HERE, "mvr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_count_invalid_rotations.restype = C.c_int
    return _lib


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))


# ------------------------------------------------------------------------------------------------
def look_at(azim, elev, dist):
    azim, elev, dist = _f(azim).ravel(), _f(elev).ravel(), _f(dist).ravel()
    n = azim.size
    R = np.empty((n, 3, 3), np.float32); T = np.empty((n, 3), np.float32); Cc = np.empty((n, 3), np.float32)
    lib().orc_look_at(_p(azim), _p(elev), _p(dist), C.c_int(n), _p(R), _p(T), _p(Cc))
    return R, T, Cc


def count_invalid_rotations(R):
    R = _f(R).reshape(-1, 3, 3)
    return lib().orc_count_invalid_rotations(_p(R), C.c_int(R.shape[0]))


def look_at_backward(azim, elev, dist, gR, gT, gC):
    azim, elev, dist = _f(azim).ravel(), _f(elev).ravel(), _f(dist).ravel()
    n = azim.size
    gR, gT, gC = _f(gR), _f(gT), _f(gC)
    ga = np.empty(n, np.float32); ge = np.empty(n, np.float32); gd = np.empty(n, np.float32)
    lib().orc_look_at_backward(_p(azim), _p(elev), _p(dist), C.c_int(n), _p(gR), _p(gT), _p(gC),
                               _p(ga), _p(ge), _p(gd))
    return ga, ge, gd


def project_perspective(verts, R, T, k00, k11):
    verts = _f(verts).reshape(-1, 3)
    out = np.empty_like(verts)
    lib().orc_project_perspective(_p(verts), C.c_int(verts.shape[0]), _p(_f(R)), _p(_f(T)),
                                  C.c_float(k00), C.c_float(k11), _p(out))
    return out


def project_orthographic(points, R, T, inv_dist):
    points = _f(points).reshape(-1, 3)
    out = np.empty_like(points)
    lib().orc_project_orthographic(_p(points), C.c_int(points.shape[0]), _p(_f(R)), _p(_f(T)),
                                   C.c_float(inv_dist), _p(out))
    return out


def rasterize_meshes(face_verts, first_idx, num_faces, H, W, K, flags, face_skip=None):
    """(F,3,3) NDC face corners -> pix_to_face (N,H,W,K) int32 packed, zbuf, bary (N,H,W,K,3), dists."""
    face_verts = _f(face_verts).reshape(-1, 3, 3)
    first_idx, num_faces = _i(first_idx), _i(num_faces)
    N = first_idx.size
    p2f = np.empty((N, H, W, K), np.int32); zbuf = np.empty((N, H, W, K), np.float32)
    bary = np.empty((N, H, W, K, 3), np.float32); dists = np.empty((N, H, W, K), np.float32)
    skip = None if face_skip is None else np.ascontiguousarray(face_skip, dtype=np.uint8)
    lib().orc_rasterize_meshes(_p(face_verts), _p(first_idx), _p(num_faces), _p(skip), C.c_int(N),
                               C.c_int(H), C.c_int(W), C.c_int(K), C.c_int(flags), _p(p2f), _p(zbuf),
                               _p(bary), _p(dists))
    return p2f, zbuf, bary, dists


def rasterize_meshes_backward(face_verts, pix_to_face, grad_zbuf, grad_bary, flags):
    face_verts = _f(face_verts).reshape(-1, 3, 3)
    p2f = _i(pix_to_face)
    N, H, W, K = p2f.shape
    out = np.empty_like(face_verts)
    lib().orc_rasterize_meshes_backward(_p(face_verts), _p(p2f), _p(_f(grad_zbuf)), _p(_f(grad_bary)),
                                        C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(K),
                                        C.c_int(face_verts.shape[0]), C.c_int(flags), _p(out))
    return out


def vertex_normals(verts, faces):
    verts = _f(verts).reshape(-1, 3); faces = _i(faces).reshape(-1, 3)
    out = np.empty_like(verts)
    lib().orc_vertex_normals(_p(verts), _p(faces), C.c_int(verts.shape[0]), C.c_int(faces.shape[0]), _p(out))
    return out


def packed_vertex_normals(verts, faces, vert_off, face_off):
    out = np.empty_like(_f(verts).reshape(-1, 3))
    for b in range(len(vert_off) - 1):
        out[vert_off[b]:vert_off[b + 1]] = vertex_normals(verts[vert_off[b]:vert_off[b + 1]],
                                                          faces[face_off[b]:face_off[b + 1]])
    return out


def mesh_forward(verts, faces, vert_off, face_off, normals, rgb, M, R, T, Cc, light, bg, k00, k11,
                 z_clip, H, W, K, flags, fragments=True):
    """Mirror of mvr_mesh_forward.  Returns dict(images, pix_to_face, zbuf, bary, dists, straddle)."""
    verts = _f(verts).reshape(-1, 3); faces = _i(faces).reshape(-1, 3)
    vert_off, face_off = _i(vert_off), _i(face_off)
    B = vert_off.size - 1
    N = B * M
    normals = _f(normals); rgb = _f(rgb); R = _f(R); T = _f(T); Cc = _f(Cc); bg = _f(bg)
    light = _f(light).reshape(-1, 3)
    light_stride = 0 if light.shape[0] == 1 else 3
    assert light.shape[0] in (1, N)
    if rgb.size != 3:
        flags |= RGB_PER_ELEMENT
        assert rgb.size == verts.size
    images = np.empty((N, 3, H, W), np.float32)
    p2f = np.empty((N, H, W, K), np.int32)
    zbuf = np.empty((N, H, W, K), np.float32) if fragments else None
    bary = np.empty((N, H, W, K, 3), np.float32) if fragments else None
    dists = np.empty((N, H, W, K), np.float32) if fragments else None
    counters = np.zeros(4, np.int64)
    lib().orc_mesh_forward(_p(verts), _p(faces), _p(vert_off), _p(face_off), _p(normals), _p(rgb),
                           C.c_int(B), C.c_int(M), _p(R), _p(T), _p(Cc), _p(light), C.c_int(light_stride),
                           _p(bg), C.c_float(k00), C.c_float(k11), C.c_float(z_clip), C.c_int(H),
                           C.c_int(W), C.c_int(K), C.c_int(flags), _p(images), _p(p2f), _p(zbuf),
                           _p(bary), _p(dists), _p(counters))
    return dict(images=images, pix_to_face=p2f, zbuf=zbuf, bary=bary, dists=dists, straddle=int(counters[0]))


def mesh_backward(verts, faces, vert_off, face_off, normals, rgb, M, R, T, Cc, light, k00, k11, H, W,
                  K, flags, pix_to_face, grad_images, want_verts=False, z_clip=None):
    """z_clip: the forward's near clip plane; None = [upstream] MeshRasterizer's default (znear / 2 = 0.5 with
    perspective_correct, no clipping otherwise)."""
    if z_clip is None:
        z_clip = 0.5 if (flags & PERSPECTIVE_CORRECT) else -1.0
    verts = _f(verts).reshape(-1, 3); faces = _i(faces).reshape(-1, 3)
    vert_off, face_off = _i(vert_off), _i(face_off)
    B = vert_off.size - 1
    N = B * M
    normals = _f(normals); rgb = _f(rgb); R = _f(R); T = _f(T); Cc = _f(Cc)
    light = _f(light).reshape(-1, 3)
    light_stride = 0 if light.shape[0] == 1 else 3
    if rgb.size != 3:
        flags |= RGB_PER_ELEMENT
    gR = np.empty((N, 3, 3), np.float32); gT = np.empty((N, 3), np.float32); gC = np.empty((N, 3), np.float32)
    gV = np.empty_like(verts) if want_verts else None
    gN = np.empty_like(verts) if want_verts else None
    lib().orc_mesh_backward(_p(verts), _p(faces), _p(vert_off), _p(face_off), _p(normals), _p(rgb),
                            C.c_int(B), C.c_int(M), _p(R), _p(T), _p(Cc), _p(light), C.c_int(light_stride),
                            C.c_float(k00), C.c_float(k11), C.c_float(z_clip), C.c_int(H), C.c_int(W), C.c_int(K),
                            C.c_int(flags), _p(_i(pix_to_face)), _p(_f(grad_images)), _p(gR), _p(gT),
                            _p(gC), _p(gV), _p(gN))
    return dict(gR=gR, gT=gT, gC=gC, grad_verts=gV, grad_normals=gN)


def rasterize_points(points, first_idx, num_points, radius, H, W, K):
    points = _f(points).reshape(-1, 3)
    first_idx, num_points = _i(first_idx), _i(num_points)
    rad = np.full(points.shape[0], radius, np.float32) if np.isscalar(radius) else _f(radius)
    N = first_idx.size
    idx = np.empty((N, H, W, K), np.int32); zbuf = np.empty((N, H, W, K), np.float32)
    d2 = np.empty((N, H, W, K), np.float32)
    lib().orc_rasterize_points(_p(points), _p(first_idx), _p(num_points), _p(rad), C.c_int(N), C.c_int(H),
                               C.c_int(W), C.c_int(K), _p(idx), _p(zbuf), _p(d2))
    return idx, zbuf, d2


def rasterize_points_backward(points, idx, grad_zbuf, grad_dists):
    points = _f(points).reshape(-1, 3); idx = _i(idx)
    N, H, W, K = idx.shape
    out = np.empty_like(points)
    lib().orc_rasterize_points_backward(_p(points), _p(idx), _p(_f(grad_zbuf)), _p(_f(grad_dists)),
                                        C.c_int(N), C.c_int(H), C.c_int(W), C.c_int(K),
                                        C.c_int(points.shape[0]), _p(out))
    return out


def composite_forward(features, alphas, idx, alpha_mode):
    """features (C,P), alphas/idx (N,K,H,W) -> (N,C,H,W)."""
    features, alphas, idx = _f(features), _f(alphas), _i(idx)
    N, K, H, W = idx.shape
    Cn, P = features.shape
    out = np.empty((N, Cn, H, W), np.float32)
    lib().orc_composite_forward(_p(features), _p(alphas), _p(idx), C.c_int(N), C.c_int(K), C.c_int(H),
                                C.c_int(W), C.c_int(Cn), C.c_int(P), C.c_int(int(alpha_mode)), _p(out))
    return out


def composite_backward(grad_out, features, alphas, idx, alpha_mode):
    features, alphas, idx = _f(features), _f(alphas), _i(idx)
    N, K, H, W = idx.shape
    Cn, P = features.shape
    gf = np.empty_like(features); ga = np.empty_like(alphas)
    lib().orc_composite_backward(_p(_f(grad_out)), _p(features), _p(alphas), _p(idx), C.c_int(N), C.c_int(K),
                                 C.c_int(H), C.c_int(W), C.c_int(Cn), C.c_int(P), C.c_int(int(alpha_mode)),
                                 _p(gf), _p(ga))
    return gf, ga


def points_forward(points, rgb, M, R, T, inv_dist, radius, bg, H, W, K, flags, fragments=True):
    """Mirror of mvr_points_forward.  points (B,Np,3)."""
    points = _f(points)
    B, Np, _ = points.shape
    N = B * M
    rgb = _f(rgb)
    if rgb.size != 3:
        flags |= RGB_PER_ELEMENT
        assert rgb.size == points.size
    images = np.empty((N, 3, H, W), np.float32)
    idx = np.empty((N, H, W, K), np.int32)
    zbuf = np.empty((N, H, W, K), np.float32) if fragments else None
    d2 = np.empty((N, H, W, K), np.float32) if fragments else None
    lib().orc_points_forward(_p(points), _p(rgb), C.c_int(B), C.c_int(Np), C.c_int(M), _p(_f(R)), _p(_f(T)),
                             _p(_f(inv_dist)), C.c_double(float(radius)), _p(_f(bg)), C.c_int(H), C.c_int(W),
                             C.c_int(K), C.c_int(flags), _p(images), _p(idx), _p(zbuf), _p(d2))
    return dict(images=images, idx=idx, zbuf=zbuf, dists2=d2)


def points_backward(points, rgb, M, R, T, inv_dist, radius, H, W, K, flags, idx, grad_images,
                    want_points=False, want_rgb=False):
    points = _f(points)
    B, Np, _ = points.shape
    N = B * M
    rgb = _f(rgb)
    if rgb.size != 3:
        flags |= RGB_PER_ELEMENT
    gR = np.empty((N, 3, 3), np.float32); gT = np.empty((N, 3), np.float32); gs = np.empty(N, np.float32)
    gP = np.empty_like(points) if want_points else None
    gRGB = np.empty_like(rgb) if want_rgb else None
    lib().orc_points_backward(_p(points), _p(rgb), C.c_int(B), C.c_int(Np), C.c_int(M), _p(_f(R)), _p(_f(T)),
                              _p(_f(inv_dist)), C.c_double(float(radius)), C.c_int(H), C.c_int(W), C.c_int(K),
                              C.c_int(flags), _p(_i(idx)), _p(_f(grad_images)), _p(gR), _p(gT), _p(gs),
                              _p(gP), _p(gRGB))
    return dict(gR=gR, gT=gT, g_inv_dist=gs, grad_points=gP, grad_rgb=gRGB)
