"""Seeded synthetic inputs for BASELINE.json's configs: point clouds, meshes, views.

The real datasets (ModelNet40 / ScanObjectNN) are absent; every benchmark and parity test uses
these generators.  Objects follow the reference's normalisation contract -- centred and scaled
into the unit sphere (util.py:437-451 torch_center_and_normalize, p="2"; config.yaml:15).
"""
import math

import numpy as np
import torch


def center_and_normalize(points: torch.Tensor) -> torch.Tensor:
    """util.py:437-451 with p="2": subtract the mean, divide by the largest L2 norm."""
    c = points.mean(0)
    scale = torch.max(torch.norm(points - c, p=2, dim=1))
    return (points - c) * (1.0 / float(scale))


def make_cloud(n_points: int, seed: int) -> torch.Tensor:
    """Points on a randomly oriented, anisotropically scaled ellipsoid surface + N(0, 0.01) noise."""
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(n_points, 3, generator=g)
    p = p / p.norm(dim=1, keepdim=True)
    p = p * (0.5 + 0.5 * torch.rand(3, generator=g))
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    p = p @ q + 0.01 * torch.randn(n_points, 3, generator=g)
    return center_and_normalize(p).contiguous()


def make_clouds(batch: int, n_points: int, seed: int) -> torch.Tensor:
    return torch.stack([make_cloud(n_points, seed + 7919 * b) for b in range(batch)])


def uv_sphere(rings: int, segments: int):
    """Closed UV sphere: V = (rings-1)*segments + 2, F = 2*segments*(rings-1), CCW seen from outside."""
    verts = [(0.0, 1.0, 0.0)]
    for r in range(1, rings):
        th = math.pi * r / rings
        for s in range(segments):
            ph = 2 * math.pi * s / segments
            verts.append((math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)))
    verts.append((0.0, -1.0, 0.0))
    south = len(verts) - 1
    faces = []
    for s in range(segments):
        s1 = (s + 1) % segments
        faces.append((0, 1 + s1, 1 + s))
        base = 1 + (rings - 2) * segments
        faces.append((south, base + s, base + s1))
    for r in range(rings - 2):
        a, b = 1 + r * segments, 1 + (r + 1) * segments
        for s in range(segments):
            s1 = (s + 1) % segments
            faces.append((a + s, a + s1, b + s))
            faces.append((a + s1, b + s1, b + s))
    return torch.tensor(verts, dtype=torch.float32), torch.tensor(faces, dtype=torch.int64)


def sphere_dims_for_faces(target_faces: int):
    """rings, segments with 2*segments*(rings-1) ~= target_faces and segments ~= 2*rings."""
    rings = max(3, int(round(math.sqrt(target_faces / 4.0))))
    segments = max(3, int(round(target_faces / (2.0 * (rings - 1)))))
    return rings, segments


def make_mesh(target_faces: int, seed: int, amplitude: float = 0.2):
    """UV sphere radially perturbed by low-frequency noise (depth complexity > 1, varying normals),
    normalised into the unit sphere.  Returns verts (V,3) f32, faces (F,3) int64."""
    rings, segments = sphere_dims_for_faces(target_faces)
    v, f = uv_sphere(rings, segments)
    g = torch.Generator().manual_seed(seed)
    freqs = torch.randint(1, 5, (6, 3), generator=g).float()
    phase = 2 * math.pi * torch.rand(6, generator=g)
    amp = amplitude * (torch.rand(6, generator=g) - 0.5) * 2 / 3
    radial = 1.0 + (amp[None] * torch.sin(v @ freqs.T * 2.0 + phase[None])).sum(1)
    scale = 0.6 + 0.4 * torch.rand(3, generator=g)
    v = v * radial[:, None] * scale
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    v = center_and_normalize(v @ q)
    return v.contiguous(), f.contiguous()


def make_meshes(batch: int, target_faces: int, seed: int):
    return [make_mesh(target_faces, seed + 104729 * b) for b in range(batch)]


def unit_spherical_grid(nb_points: int):
    """util.py:359-391: azimuth/elevation (degrees) of a near-uniform grid on the sphere."""
    alpha = 4.0 * np.pi / nb_points
    d = np.sqrt(alpha)
    m_nu = int(np.round(np.pi / d))
    d_nu = np.pi / m_nu
    d_phi = alpha / d_nu
    azim, elev = [], []
    for m in range(m_nu):
        nu = np.pi * (m + 0.5) / m_nu
        m_phi = int(np.round(2 * np.pi * np.sin(nu) / d_phi))
        for n in range(m_phi):
            azim.append(2 * np.pi * n / m_phi)
            elev.append(nu - np.pi * 0.5)
    return np.rad2deg(azim)[:nb_points], np.rad2deg(elev)[:nb_points]


def circular_views(batch: int, nb_views: int, elevation: float = 30.0, distance: float = 2.2):
    """models/mvtn.py:13-36 CircularViewSelector (config.yaml:23-24 canonical 30 deg / 2.2)."""
    azim = torch.linspace(-180, 180, nb_views + 1)[:-1] - 90.0
    elev = torch.full_like(azim, elevation)
    dist = torch.full_like(azim, distance)
    return tuple(t.expand(batch, nb_views).contiguous() for t in (azim, elev, dist))


def spherical_views(batch: int, nb_views: int, distance: float = 2.2):
    """models/mvtn.py:39-75 SphericalViewSelector."""
    a, e = unit_spherical_grid(nb_views)
    azim = torch.from_numpy(a).float(); elev = torch.from_numpy(e).float()
    dist = torch.full_like(azim, distance)
    return tuple(t.expand(batch, nb_views).contiguous() for t in (azim, elev, dist))


def learned_spherical_views(batch: int, nb_views: int, seed: int, distance: float = 2.2):
    """models/mvtn.py:223-248 LearnedSphericalViewSelector with a seeded stand-in for the MLP:
    spherical grid + tanh-bounded offsets U(-1,1) * (180/M, 90)."""
    azim, elev, dist = spherical_views(batch, nb_views, distance)
    g = torch.Generator().manual_seed(seed)
    off = torch.rand(batch, 2 * nb_views, generator=g) * 2 - 1
    azim = azim + off[:, :nb_views] * 180.0 / nb_views
    elev = elev + off[:, nb_views:] * 90.0 * 0.9   # keep |elev| away from the +-90 degeneracy
    elev = elev.clamp(-89.0, 89.0)
    return azim.contiguous(), elev.contiguous(), dist.contiguous()
