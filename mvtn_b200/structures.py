"""Minimal stand-ins for the PyTorch3D containers MVRenderer's callers hand over (PyTorch3D itself is
not a dependency).  MVRenderer only duck-types: anything exposing verts_list() / faces_list() works."""
from typing import List, Sequence

import torch


class Meshes:
    """List-of-meshes container with the accessors renderer.py:67-68 uses."""

    def __init__(self, verts: Sequence[torch.Tensor], faces: Sequence[torch.Tensor]):
        if isinstance(verts, torch.Tensor):
            verts = list(verts) if verts.dim() == 3 else [verts]
        if isinstance(faces, torch.Tensor):
            faces = list(faces) if faces.dim() == 3 else [faces]
        if len(verts) != len(faces):
            raise ValueError("verts and faces must have the same length")
        self._verts: List[torch.Tensor] = list(verts)
        self._faces: List[torch.Tensor] = list(faces)

    def verts_list(self):
        return self._verts

    def faces_list(self):
        return self._faces

    def __len__(self):
        return len(self._verts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return Meshes(self._verts[i], self._faces[i])
        return Meshes([self._verts[i]], [self._faces[i]])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def extend(self, n: int):
        v, f = [], []
        for a, b in zip(self._verts, self._faces):
            v += [a.clone() for _ in range(n)]
            f += [b.clone() for _ in range(n)]
        return Meshes(v, f)


def unpack_mesh_list(meshes):
    """renderer.py:67-68 `[msh.verts_list()[0] for msh in meshes]` for a list of single-mesh objects or one
    batched object (run_mvtn.py:517-533 passes a batched Meshes)."""
    verts, faces = [], []
    if hasattr(meshes, "verts_list") and not isinstance(meshes, (list, tuple)):
        return list(meshes.verts_list()), list(meshes.faces_list())
    for m in meshes:
        if isinstance(m, (tuple, list)) and len(m) == 2:
            verts.append(m[0]); faces.append(m[1])
        else:
            verts.append(m.verts_list()[0]); faces.append(m.faces_list()[0])
    return verts, faces
