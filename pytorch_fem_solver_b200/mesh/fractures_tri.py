"""Planar fractures embedded in 3-D (reference torch_fem/mesh/fractures_tri.py)."""

from __future__ import annotations

import torch

from .meshes_tri import MeshesTri


class FracturesTri(MeshesTri):
    """Stack of 2-D fracture meshes plus the affine map of each into 3-D."""

    def __init__(self, triangulations: list, fractures_3d_data: torch.Tensor):
        super().__init__(triangulations)
        self._compute_fracture_map(torch.as_tensor(fractures_3d_data).to(self["vertices", "coordinates"]))
        jac, shift = self["jacobian_fracture_map"], self["translation_vector"]
        self._triangulation["vertices", "coordinates_3d"] = (jac @ self["vertices", "coordinates"].mT + shift).mT
        self._triangulation["cells", "coordinates_3d"] = self.compute_coordinates_4_cells(
            self["vertices", "coordinates_3d"], self["cells", "vertices"]
        )

    def _build_optional_parameters(self):
        super()._build_optional_parameters()
        # reference :29-33 maps the unit normals with the full affine map (translation included)
        jac, shift = self._triangulation["jacobian_fracture_map"], self._triangulation["translation_vector"]
        normals = self._triangulation["interior_edges", "normals"]
        self._triangulation["interior_edges", "normals_3d"] = (jac.unsqueeze(-3) @ normals.mT + shift.unsqueeze(-3)).mT

    def _compute_fracture_map(self, fractures_3d_data: torch.Tensor):
        """Affine map x3 = J_f x2 + t fixed by the first three vertices (reference :35-67)."""
        v2 = self["vertices", "coordinates"][:, :3, :]
        v3 = fractures_3d_data[:, :3, :]
        homogeneous = torch.cat([v2, torch.ones_like(v3[..., :1])], dim=-1)
        affine = v3.mT @ torch.inverse(homogeneous).mT  # (F,3,3) = [J_f | t]
        jac = affine[..., :2]
        shift = affine[..., 2:3]
        normal = torch.cross(jac[..., 0], jac[..., 1], dim=-1)
        det = torch.linalg.vector_norm(normal, dim=-1).reshape(-1, 1, 1)
        pinv = torch.inverse(jac.mT @ jac) @ jac.mT
        self["jacobian_fracture_map"] = jac
        self["inv_jacobian_fracture_map"] = pinv
        self["det_jacobian_fracture_map"] = det
        self["translation_vector"] = shift
