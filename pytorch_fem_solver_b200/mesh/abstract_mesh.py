"""Mesh container: `triangle`-style dict -> nested tensor dict with derived topology.

Mirrors the key layout and shapes of the reference's `AbstractMesh`
(torch_fem/mesh/abstract_mesh.py:10-316) with a different construction:

* single and batched meshes share one vectorised implementation (the reference loops over
  meshes in Python, meshes_tri.py:54-151);
* edge topology (interior/boundary split, edge->cells, normals, lengths) is built lazily on
  first access -- it is integer set-up outside the assembly hot path
  and `Patches` never needs it;
* edge->cell adjacency is sort-based and always ALIGNED with `interior_edges.vertices`
  (the reference's `neighbors` shortcut is not); cells of an edge are
  listed in ascending cell id, as the reference's brute-force path yields (:245-253).
"""

from __future__ import annotations

import abc
import os
from typing import Any

import numpy as np
import torch

from ..tensordict_lite import TensorDict

_KEY_MAP = {
    "vertices": ("vertices", "coordinates"),
    "vertex_markers": ("vertices", "markers"),
    "triangles": ("cells", "vertices"),
    "neighbors": ("cells", "neighbors"),
    "edges": ("edges", "vertices"),
    "edge_markers": ("edges", "markers"),
}
_LAZY_GROUPS = ("interior_edges", "boundary_edges")


def _to_tensor(value: Any) -> torch.Tensor:
    """ints -> torch.int32, floats -> default dtype (intent of reference :51-58), on the default device."""
    if isinstance(value, torch.Tensor):
        t = value
    else:
        t = torch.as_tensor(np.asarray(value))
    if t.dtype.is_floating_point:
        return t.to(dtype=torch.get_default_dtype(), device=torch.get_default_device())
    return t.to(dtype=torch.int32, device=torch.get_default_device())


class AbstractMesh(abc.ABC):
    """Triangle mesh (or a batch of equally sized meshes) held as nested tensors."""

    def __init__(self, triangulation: dict[str, Any]):
        self._triangulation = self._triangle_to_tensordict(triangulation)
        self._triangulation["cells", "coordinates"] = self.compute_coordinates_4_cells(
            self._triangulation["vertices", "coordinates"], self._triangulation["cells", "vertices"]
        )
        self._topology_ready = False

    # ---- mapping surface ---------------------------------------------------------------------
    def _needs_topology(self, key) -> bool:
        if self._topology_ready:
            return False
        if isinstance(key, tuple):
            return key[0] in _LAZY_GROUPS or key == ("cells", "length") or key[0] == "edges"
        return key in _LAZY_GROUPS or key in ("edges", "cells")

    def __getitem__(self, key):
        if self._needs_topology(key):
            self._build_optional_parameters()
        return self._triangulation[key]

    def __setitem__(self, key, value):
        self._triangulation[key] = value

    def __contains__(self, key):
        if self._needs_topology(key):
            self._build_optional_parameters()
        return key in self._triangulation

    def batch_size(self):
        """Leading batch dims: [] for one mesh, [F] for stacked meshes."""
        coords = self._triangulation["vertices", "coordinates"]
        return torch.Size(coords.shape[:-2])

    @property
    def is_batched(self) -> bool:
        return self._triangulation["vertices", "coordinates"].dim() == 3

    @property
    def device(self) -> torch.device:
        return self._triangulation["vertices", "coordinates"].device

    def to(self, device):
        """Move every tensor of the mesh (in place) and return self."""
        self._triangulation = self._triangulation.to(device)
        return self

    def copy_to(self, device) -> "AbstractMesh":
        """A copy of the mesh on `device`; the caller's object (and its tensors) stay where they are."""
        import copy

        other = copy.copy(self)
        other._triangulation = self._triangulation.to(device)
        for name, value in vars(self).items():  # tensor attributes of subclasses (Patches: centers, radius, ...)
            if isinstance(value, torch.Tensor):
                setattr(other, name, value.to(device))
        return other

    # ---- construction ------------------------------------------------------------------------
    def _triangle_to_tensordict(self, mesh_dict) -> TensorDict:
        groups: dict[str, dict] = {"vertices": {}, "cells": {}, "edges": {}}
        for key, value in mesh_dict.items():
            if key in _KEY_MAP:
                group, name = _KEY_MAP[key]
                groups[group][name] = _to_tensor(value)
        if "coordinates" not in groups["vertices"] or "vertices" not in groups["cells"]:
            raise ValueError("a mesh needs at least 'vertices' and 'triangles'")
        return TensorDict({g: TensorDict(content) for g, content in groups.items()}).auto_batch_size_()

    @staticmethod
    def compute_coordinates_4_cells(coordinates_4_vertices: torch.Tensor, vertices_4_cells: torch.Tensor):
        """`X[e,k,:] = coords[conn[e,k],:]` (reference :257-262; batched meshes_tri.py:33-41)."""
        return coordinates_4_vertices[vertices_4_cells.long()]

    # ---- lazily derived topology ---------------------------------------------------------------
    def _flat(self, key):
        """Tensor with a leading mesh axis (size 1 for a single mesh)."""
        t = self._triangulation[key]
        return t if self.is_batched else t.unsqueeze(0)

    def _store(self, key, value):
        self._triangulation[key] = value if self.is_batched else value.squeeze(0)

    def _build_optional_parameters(self):
        """Edges (if absent), interior/boundary split, edge->cells, normals, lengths."""
        self._topology_ready = True  # set first: the helpers below read through __getitem__
        td = self._triangulation
        conn = self._flat(("cells", "vertices")).long()  # (F,N,3)
        coords = self._flat(("vertices", "coordinates"))  # (F,V,2)
        n_mesh, n_cells, _ = conn.shape
        n_vert = coords.shape[1]
        device = conn.device
        perms = self._edges_permutations.to(device)

        # half-edge table, stably sorted by (mesh, min vertex, max vertex): hand-written kernels + CUB on a CUDA device
        # (`tfem_half_edges`, `tfem_edge_cells`, `tfem_interior_edge_geometry`), torch ops elsewhere
        native = (device.type == "cuda" and coords.dtype in (torch.float64, torch.float32) and coords.shape[-1] == 2
                  and 3 * n_mesh * n_cells < 2**31 and os.environ.get("TFEM_TOPOLOGY", "native") == "native")
        he = conn[:, :, perms]  # (F,N,3,2)
        if native:
            from .. import ops

            conn32 = conn.to(torch.int32).contiguous()
            he_sorted, cell_sorted, uniq, counts = ops.half_edges(conn32, n_vert, perms.tolist())
        else:
            lo, hi = he.min(-1).values, he.max(-1).values
            mesh_off = torch.arange(n_mesh, device=device).reshape(-1, 1, 1) * (n_vert * n_vert)
            he_key = (mesh_off + lo * n_vert + hi).reshape(-1)
            he_cell = torch.arange(n_cells, device=device).repeat_interleave(3).repeat(n_mesh)
            he_sorted, he_order = torch.sort(he_key, stable=True)
            cell_sorted = he_cell[he_order]
            uniq = counts = None

        if "vertices" not in td["edges"]:
            if uniq is None:
                uniq, counts = torch.unique_consecutive(he_sorted, return_counts=True)
            per_mesh = torch.bincount(torch.div(uniq, n_vert * n_vert, rounding_mode="floor"), minlength=n_mesh)
            if not bool((per_mesh == per_mesh[0]).all()):
                raise ValueError("stacked meshes must have the same number of edges")
            local = uniq % (n_vert * n_vert)
            edges = torch.stack([torch.div(local, n_vert, rounding_mode="floor"), local % n_vert], -1)
            self._store(("edges", "vertices"), edges.reshape(n_mesh, -1, 2).to(torch.int32))
            # the reference stores the incidence count as marker (1 = boundary), :264-281
            self._store(("edges", "markers"), counts.reshape(n_mesh, -1, 1).to(torch.int32))

        edge_vertices = self._flat(("edges", "vertices"))  # (F,E,2)
        markers = self._flat(("edges", "markers")).reshape(n_mesh, -1)
        boundary_mask = markers == 1
        n_boundary = boundary_mask.sum(1)
        if not bool((n_boundary == n_boundary[0]).all()):
            raise ValueError("stacked meshes must have the same number of boundary edges")
        v_boundary = edge_vertices[boundary_mask].reshape(n_mesh, -1, 2)
        v_interior = edge_vertices[~boundary_mask].reshape(n_mesh, -1, 2)

        def cells_of(edge_list, n_sides):
            if native:
                return ops.edge_cells(edge_list.to(torch.int32).contiguous(), n_vert, n_sides, he_sorted, cell_sorted).long()
            e = edge_list.long()
            key = (
                torch.arange(n_mesh, device=device).reshape(-1, 1) * (n_vert * n_vert)
                + e.min(-1).values * n_vert
                + e.max(-1).values
            ).reshape(-1)
            pos = torch.searchsorted(he_sorted, key)
            if key.numel() and not bool((he_sorted[pos.clamp_max(he_sorted.numel() - 1)] == key).all()):
                raise ValueError("an edge of the edge list belongs to no cell")
            sides = [cell_sorted[(pos + s).clamp_max(cell_sorted.numel() - 1)] for s in range(n_sides)]
            if n_sides == 2 and key.numel():
                second_ok = he_sorted[(pos + 1).clamp_max(he_sorted.numel() - 1)] == key
                if not bool(second_ok.all()):
                    raise ValueError("an edge marked interior has a single adjacent cell")
            return torch.stack(sides, -1).reshape(n_mesh, -1, n_sides)

        c_interior = cells_of(v_interior, 2)
        c_boundary = cells_of(v_boundary, 1)

        batch = torch.arange(n_mesh, device=device).reshape(-1, 1, 1)
        x_boundary = coords[batch, v_boundary.long()]
        if native:
            x_interior, length, normal = ops.interior_edge_geometry(coords.contiguous(), conn32, v_interior.to(torch.int32).contiguous(),
                                                                    c_interior.to(torch.int32).contiguous())
        else:
            x_interior = coords[batch, v_interior.long()]  # (F,E_i,2,2)
            vec = x_interior[..., 1:2, :] - x_interior[..., 0:1, :]  # (F,E_i,1,2)
            length = torch.linalg.vector_norm(vec, dim=-1, keepdim=True)
            normal = torch.stack([-vec[..., 1], vec[..., 0]], dim=-1) / length
            # orient from the first listed cell's centroid towards the second's (reference :143-162)
            cell_x = self._flat(("cells", "coordinates"))
            centroid = cell_x[batch, c_interior].mean(dim=-2)  # (F,E_i,2,2)
            towards = centroid[..., 1:2, :] - centroid[..., 0:1, :]
            flip = (normal * towards).sum(-1, keepdim=True) < 0
            normal = torch.where(flip, -normal, normal)

        self._store(("interior_edges", "cells"), c_interior)
        self._store(("interior_edges", "vertices"), v_interior)
        self._store(("interior_edges", "coordinates"), x_interior)
        self._store(("interior_edges", "length"), length)
        self._store(("interior_edges", "normals"), normal)
        self._store(("boundary_edges", "cells"), c_boundary)
        self._store(("boundary_edges", "vertices"), v_boundary)
        self._store(("boundary_edges", "coordinates"), x_boundary)

        # per-cell edge lengths, edges in `_edges_permutations` order (reference :283-309; its
        # `min` reduces a size-1 axis, so all three lengths are kept)
        x_he = coords[torch.arange(n_mesh, device=device).reshape(-1, 1, 1, 1), he]  # (F,N,3,2,2)
        cell_len = torch.linalg.vector_norm(x_he[..., 1, :] - x_he[..., 0, :], dim=-1)
        self._store(("cells", "length"), cell_len.reshape(n_mesh, n_cells, 3, 1, 1))
        self._triangulation.auto_batch_size_()

    @property
    @abc.abstractmethod
    def _edges_permutations(self) -> torch.Tensor:
        """Local vertex pairs of the three cell edges."""
        raise NotImplementedError
