"""Generates the committed golden fixtures from the CPU oracle (run once in the build container):

    python tests/golden/make_golden.py

PyTorch3D (the reference's arithmetic) cannot be imported here, so these vectors pin the ORACLE (and
through it the CUDA path) against regressions; the analytic known-answer tests in tests/test_oracle.py pin
the oracle itself.  Geometry and cameras are stored, not regenerated, so the hashes do not depend on the
host's libm / BLAS.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from mvtn_b200 import synth, ops          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mesh_slice():
    """BASELINE configs[1] slice: one ~10k-face mesh x 12 circular views, 224x224, K=1."""
    v, f = synth.make_mesh(10000, 1236)
    az, el, di = synth.circular_views(1, 12)
    R, T, C = orc.look_at(az.numpy().ravel(), el.numpy().ravel(), di.numpy().ravel())
    vp = v.numpy(); fp = f.numpy().astype(np.int32)
    voff = np.array([0, vp.shape[0]], np.int32); foff = np.array([0, fp.shape[0]], np.int32)
    nrm = orc.vertex_normals(vp, fp)
    k00, k11 = ops.fov_projection_scale()
    col = np.full(3, 0.99999, np.float32); light = np.array([[0, 1.0, 0]], np.float32)
    o = orc.mesh_forward(vp, fp, voff, foff, nrm, col, 12, R, T, C, light, col, k00, k11, 0.5, 224, 224, 1,
                         orc.PERSPECTIVE_CORRECT)
    np.savez_compressed(os.path.join(HERE, "mesh_c2_slice.npz"), verts=vp, faces=fp, R=R, T=T, C=C,
                        k00=np.float32(k00), k11=np.float32(k11),
                        p2f_sha256=sha(o["pix_to_face"]), zbuf_sha256=sha(o["zbuf"]), bary_sha256=sha(o["bary"]),
                        covered=(o["pix_to_face"][..., 0] >= 0).sum(axis=(1, 2)).astype(np.int64),
                        image_sum=o["images"].astype(np.float64).sum(axis=(1, 2, 3)),
                        image_probe=o["images"][:, :, ::16, ::16].copy())
    print("mesh slice: covered", (o["pix_to_face"] >= 0).mean())


def points_c1():
    """BASELINE configs[0]: one 2048-point cloud, 12 circular views, 224x224, K=1, norm-weighted."""
    pts = synth.make_clouds(1, 2048, 1235)
    az, el, di = synth.circular_views(1, 12)
    R, T, C = orc.look_at(az.numpy().ravel(), el.numpy().ravel(), di.numpy().ravel())
    inv = (1.0 / di.numpy().ravel()).astype(np.float32)
    col = np.full(3, 0.99999, np.float32)
    o = orc.points_forward(pts.numpy(), col, 12, R, T, inv, 0.006, np.zeros(3, np.float32), 224, 224, 1, 0)
    o4 = orc.points_forward(pts.numpy(), col, 12, R, T, inv, 0.02, np.zeros(3, np.float32), 224, 224, 4,
                            orc.COMPOSITE_ALPHA)
    np.savez_compressed(os.path.join(HERE, "points_c1.npz"), points=pts.numpy(), R=R, T=T, inv_dist=inv,
                        idx_sha256=sha(o["idx"]), zbuf_sha256=sha(o["zbuf"]), d2_sha256=sha(o["dists2"]),
                        covered=(o["idx"][..., 0] >= 0).sum(axis=(1, 2)).astype(np.int64),
                        image_sum=o["images"].astype(np.float64).sum(axis=(1, 2, 3)),
                        idx4_sha256=sha(o4["idx"]), image4_sum=o4["images"].astype(np.float64).sum(axis=(1, 2, 3)))
    print("points: covered", (o["idx"] >= 0).mean(), (o4["idx"][..., 0] >= 0).mean())


def mesh_clip():
    """Close-up (dist 1.12 - 1.3, the range mvtn.py:33 transform_distance reaches): faces cross z_clip = 0.5 and are clipped
    ([upstream] clip.py).  One ~1.5k-face mesh x 3 views, 96x96, K = 2.  Only mesh_clip.npz is (re)written."""
    v, f = synth.make_mesh(1500, 61)
    az = np.array([15.0, 140.0, -80.0], np.float32); el = np.array([10.0, -35.0, 50.0], np.float32)
    di = np.array([1.12, 1.2, 1.3], np.float32)
    R, T, C = orc.look_at(az, el, di)
    vp = v.numpy(); fp = f.numpy().astype(np.int32)
    voff = np.array([0, vp.shape[0]], np.int32); foff = np.array([0, fp.shape[0]], np.int32)
    nrm = orc.vertex_normals(vp, fp)
    k00, k11 = ops.fov_projection_scale()
    col = np.full(3, 0.99999, np.float32); light = np.array([[0, 1.0, 0]], np.float32)
    o = orc.mesh_forward(vp, fp, voff, foff, nrm, col, 3, R, T, C, light, col, k00, k11, 0.5, 96, 96, 2, orc.PERSPECTIVE_CORRECT)
    assert o["straddle"] > 0
    np.savez_compressed(os.path.join(HERE, "mesh_clip.npz"), verts=vp, faces=fp, R=R, T=T, C=C,
                        k00=np.float32(k00), k11=np.float32(k11), straddle=np.int64(o["straddle"]),
                        p2f_sha256=sha(o["pix_to_face"]), zbuf_sha256=sha(o["zbuf"]), bary_sha256=sha(o["bary"]),
                        dists_sha256=sha(o["dists"]),
                        covered=(o["pix_to_face"][..., 0] >= 0).sum(axis=(1, 2)).astype(np.int64),
                        image_sum=o["images"].astype(np.float64).sum(axis=(1, 2, 3)),
                        image_probe=o["images"][:, :, ::8, ::8].copy())
    print("mesh clip: straddling faces", o["straddle"], "covered", (o["pix_to_face"][..., 0] >= 0).mean())


if __name__ == "__main__":
    if "--only-clip" not in sys.argv:
        mesh_slice()
        points_c1()
    mesh_clip()
