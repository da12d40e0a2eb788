"""Square 5-vertex / 4-triangle patches (reference torch_fem/mesh/patches.py)."""

from __future__ import annotations

import math

import torch

from .meshes_tri import MeshesTri


class Patches(MeshesTri):
    """P patches: corner vertices `c + r*(+-1,+-1)` counter-clockwise, centre last."""

    def __init__(self, centers: torch.Tensor, radius: torch.Tensor):
        self.centers = torch.as_tensor(centers)
        self.radius = torch.as_tensor(radius)
        super().__init__(self._compute_patches(self.centers, self.radius))

    def _compute_patches(self, centers: torch.Tensor, radius: torch.Tensor):
        signs = self.signs_4_vertices.to(centers.device)
        coordinates = centers.unsqueeze(-2) + signs * radius.unsqueeze(-2)  # (P,5,2)
        n_patches = coordinates.shape[0]
        c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)
        self.rotation_matrix = torch.tensor([[c, -s], [s, c]], device=centers.device)
        self.rotated_signs = (self.rotation_matrix @ signs.to(self.rotation_matrix.dtype).mT).mT
        # one stacked dict directly: no per-patch Python loop (the reference builds P dicts, :39-45)
        return {
            "vertices": coordinates,
            "triangles": self.vertices_4_cells_4_patch.expand(n_patches, 4, 3),
            "vertex_markers": self.markers_4_vertices.expand(n_patches, 5, 1),
        }

    def refine_patches(self, refine_idx: torch.Tensor, maintain_old_patches: bool = False):
        """Split marked patches into four children plus one rotated patch (reference :49-135)."""
        signs = self.signs_4_vertices.to(self.centers.device)
        child_radius = 0.5 * self.radius[refine_idx]
        child_centers = self.centers[refine_idx, :].unsqueeze(-2) + signs[:-1, :] * child_radius.unsqueeze(-2)
        child_vertices = child_centers.unsqueeze(-2) + (signs * child_radius.unsqueeze(-2)).unsqueeze(-3)
        rotated_radius = 2 * child_radius / math.sqrt(2.0)
        rotated_centers = self.centers[refine_idx, :]
        rotated_vertices = rotated_centers.unsqueeze(-2) + self.rotated_signs.to(self.centers).unsqueeze(0) * rotated_radius.unsqueeze(-1)
        keep = slice(None) if maintain_old_patches else ~refine_idx
        radius = torch.cat([self.radius[keep], child_radius.repeat(4, 1), rotated_radius], dim=0)
        centers = torch.cat([self.centers[keep, :], child_centers.reshape(-1, 2), rotated_centers], dim=0)
        vertices = torch.cat(
            [self["vertices", "coordinates"][keep, ...], child_vertices.reshape(-1, 5, 2), rotated_vertices], dim=0
        )
        return centers, radius, vertices

    def uniform_refine(self, nb_refinements: int = 1):
        """Refine every patch (reference :137-149)."""
        centers, radius, vertices = self.centers, self.radius, None
        for _ in range(nb_refinements):
            centers, radius, vertices = self.refine_patches(torch.ones(self.batch_size()[0], dtype=torch.bool))
        return centers, radius, vertices

    @property
    def signs_4_vertices(self):
        return torch.tensor([[-1, -1], [1, -1], [1, 1], [-1, 1], [0, 0]], dtype=torch.int64)

    @property
    def vertices_4_cells_4_patch(self):
        return torch.tensor([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], dtype=torch.int64)

    @property
    def markers_4_vertices(self):
        return torch.tensor([[1], [1], [1], [1], [0]])
