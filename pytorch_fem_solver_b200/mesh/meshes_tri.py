"""Stack of equally sized triangular meshes (reference torch_fem/mesh/meshes_tri.py)."""

from __future__ import annotations

import torch

from .. import tensordict_lite
from .mesh_tri import MeshTri


class MeshesTri(MeshTri):
    """Batch of meshes sharing vertex/cell counts; tensors gain a leading mesh axis."""

    def __init__(self, triangulations):
        if isinstance(triangulations, (list, tuple)):
            triangulations = self._stack_triangulations(list(triangulations))
        super().__init__(triangulations)

    def _stack_triangulations(self, triangulations: list):
        """List of mesh dicts -> one dict of stacked tensors (reference :17-31)."""
        return tensordict_lite.stack([tensordict_lite.TensorDict(t) for t in triangulations], dim=0)

    @staticmethod
    def compute_coordinates_4_cells(coordinates_4_vertices: torch.Tensor, vertices_4_cells: torch.Tensor):
        """Per-mesh gather `out[f, ...] = values[f, index[f, ...]]` (reference :33-41)."""
        index = vertices_4_cells.long()
        batch = torch.arange(coordinates_4_vertices.size(0), device=index.device).reshape(-1, *([1] * (index.dim() - 1)))
        return coordinates_4_vertices[batch, index]

    @staticmethod
    def apply_mask(tensor: torch.Tensor, mask: torch.Tensor):
        """Row-wise boolean / index selection keeping the mesh axis (reference :43-52)."""
        return torch.stack([t[m] for t, m in zip(tensor, mask)], dim=0)
