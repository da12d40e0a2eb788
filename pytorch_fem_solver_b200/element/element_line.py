"""P1 line element for interior-edge quadrature, reference torch_fem/element/element_line.py."""

from __future__ import annotations

import torch

from .abstract_element import AbstractElement


class ElementLine(AbstractElement):
    """1-D line element on [-1, 1]."""

    @property
    def barycentric_grad(self):
        return torch.tensor([[-0.5], [0.5]])

    @property
    def reference_element_area(self):
        return 2.0

    def _compute_gauss_values(self):
        if self.integration_order == 2:
            node = 1.0 / torch.sqrt(torch.tensor(3.0))
            nodes = torch.stack([-node, node]).reshape(2, 1)
            weights = torch.tensor([0.5, 0.5])
        elif self.integration_order == 3:
            node = torch.sqrt(torch.tensor(3 / 5))
            nodes = torch.stack([torch.zeros_like(node), -node, node]).reshape(3, 1)
            weights = torch.tensor([8 / 18, 5 / 18, 5 / 18])
        else:
            raise NotImplementedError("Integration order not implemented")
        return nodes, weights.reshape(1, -1, 1, 1)

    def compute_barycentric_coordinates(self, x: torch.Tensor):
        return torch.cat([0.5 * (1.0 - x), 0.5 * (1.0 + x)], dim=-1)

    def compute_shape_functions(self, bar_coords: torch.Tensor, inv_map_jacobian: torch.Tensor):
        if self.polynomial_order != 1:
            raise NotImplementedError("Polynomial order not implemented")
        return bar_coords, self.barycentric_grad.to(inv_map_jacobian) @ inv_map_jacobian

    def compute_det_and_inv_map(self, map_jacobian: torch.Tensor):
        det = torch.linalg.vector_norm(map_jacobian, dim=-2, keepdim=True)
        return det.unsqueeze(-1), 1.0 / det
