"""CUDA-graph capture of the device part of a render step.

The point path at MVTN's sizes is launch-bound: ~10 kernels of 20-200 us each plus the autograd plumbing cost more
host time than device time.  When shapes are fixed across iterations (a training loop over equal-sized batches), the
whole chain   look_at -> [mesh prepare] -> rasterize -> shade/composite   and its backward
rasterizer backward -> look_at backward   can be captured ONCE into two CUDA graphs and replayed with one launch each.
No tracing compiler is involved: torch.cuda.make_graphed_callables records the kernels our C ABI launches on the capture
stream (plus the memset / small fills), and keeps static input / gradient buffers.

    step = graphed_points_render(points, rgb, M=12, radius=0.006, bg=bg, image_size=224, sample_views=(az, el, di))
    images = step(azim, elev, dist)          # (B*M, 3, H, W); differentiable w.r.t. azim / elev / dist
    points.copy_(next_batch)                 # new data goes INTO the captured buffers

Restrictions (those of CUDA graphs): fixed B, M, image size, K and buffer addresses; the rotation-validity flag of
look_at is not read back (use MVRenderer for the guarded path); gradients flow to the views only.
"""
from typing import Sequence, Tuple

import torch

from . import _lib as L
from . import ops


def _views_like(sample_views: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
    out = []
    for t in sample_views:
        if not t.is_cuda:
            raise L.MVRError("sample_views must be CUDA tensors")
        out.append(t.detach().clone().to(torch.float32).requires_grad_())
    return tuple(out)


def graphed_points_render(points: torch.Tensor, rgb: torch.Tensor, M: int, radius: float, bg: torch.Tensor, image_size,
                          sample_views: Sequence[torch.Tensor], points_per_pixel: int = 1, compositor: str = "norm",
                          normalize=None, out_dtype=None, return_cameras: bool = False):
    """Capture look_at + point rasterization + compositing (and their backward) for a FIXED (B, N, 3) device tensor
    `points` (update it -- and `rgb` / `bg` -- in place between replays).
    Returns step(azim, elev, dist) -> images (B*M, 3, H, W); with return_cameras=True -> (images, cams, invalid, idx, mask)
    where cams is the flat R | T | C buffer (9n + 3n + 3n floats, n = B*M), invalid the rotation-validity flag of look_at,
    idx / mask the sparse fragment indices and the 1-bit hit mask (all live in captured buffers: copy what must outlive the
    next replay)."""
    ops._require_cuda(points, "points")
    rgb = rgb.to(points.device)
    bg = bg.to(points.device)

    def fn(az, el, di):
        img, (R, T, C, bad), frag = ops.render_points_from_angles(points, rgb, M, az, el, di, radius, bg, image_size,
                                                                  points_per_pixel=points_per_pixel, compositor=compositor,
                                                                  normalize=normalize, out_dtype=out_dtype)
        if return_cameras:
            idx, mask = frag._raw[0], frag._raw[1]      # sparse idx + 1-bit hit mask (ops._PointFragments completes them lazily)
            return img, torch.cat([R.reshape(-1), T.reshape(-1), C.reshape(-1)]), bad, idx, mask
        return img

    return torch.cuda.make_graphed_callables(fn, _views_like(sample_views))


def graphed_mesh_render(geom: "ops.PackedMeshes", M: int, light: torch.Tensor, obj_rgb, bg: torch.Tensor, image_size,
                        sample_views: Sequence[torch.Tensor], refresh_geometry: bool = True, return_cameras: bool = False,
                        **render_kwargs):
    """Capture [mvr_mesh_prepare] + look_at + mesh rasterization + Phong shading (and their backward) for a packed
    batch whose topology (counts, faces) is fixed.  With refresh_geometry=True the vertex positions may be updated in
    place (geom.verts.copy_(...)) between replays: packing and vertex normals are part of the graph.
    Returns step(azim, elev, dist) -> images (B*M, 3, H, W).  `light` (1,3) or None for the camera-relative light.
    return_cameras=True -> (images, cams, invalid, pix_to_face) as graphed_points_render does (cams: the flat R | T | C buffer;
    everything lives in captured buffers)."""
    geom.finish()

    def fn(az, el, di):
        if refresh_geometry:
            geom.refresh()
        R, T, C, bad = ops._LookAt.apply(az.reshape(-1), el.reshape(-1), di.reshape(-1))
        lt = C.detach() if light is None else light
        img, frag = ops.render_meshes(geom, M, R, T, C, lt, obj_rgb, bg, image_size, **render_kwargs)
        if return_cameras:
            return img, torch.cat([R.reshape(-1), T.reshape(-1), C.reshape(-1)]), bad, frag["pix_to_face"]
        return img

    return torch.cuda.make_graphed_callables(fn, _views_like(sample_views))
