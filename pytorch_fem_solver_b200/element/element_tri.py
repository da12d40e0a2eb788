"""Triangle element: P1 / P2 shape functions, quadrature orders 1-4 (1/3/4/6 points), reference torch_fem/element/element_tri.py."""

from __future__ import annotations

import torch

from .abstract_element import AbstractElement

# (nodes, weights) per integration order; same values as csrc/common.cuh::tri_table
# (reference element_tri.py:77-130)
_A, _B = 0.816847572980459, 0.091576213509771
_C, _D = 0.108103018168070, 0.445948490915965
_TABLES = {
    1: ([[1 / 3, 1 / 3]], [1.0]),
    2: ([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]], [1 / 3] * 3),
    3: ([[1 / 3, 1 / 3], [0.6, 0.2], [0.2, 0.6], [0.2, 0.2]], [-9 / 16] + [25 / 48] * 3),
    4: ([[_A, _B], [_B, _A], [_B, _B], [_C, _D], [_D, _C], [_D, _D]], [0.109951743655322] * 3 + [0.223381589678011] * 3),
}


class ElementTri(AbstractElement):
    """2-D triangular element."""

    @property
    def barycentric_grad(self):
        return torch.tensor([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])

    @property
    def reference_element_area(self):
        return 0.5

    @property
    def outward_normal(self):
        """Outward normals of the reference triangle's edges (not normalised)."""
        return torch.tensor([[1.0, 1.0], [-1.0, 0.0], [0.0, -1.0]])

    def _compute_gauss_values(self):
        if self.integration_order not in _TABLES:
            raise NotImplementedError("Integration order not implemented")
        nodes, weights = _TABLES[self.integration_order]
        return torch.tensor(nodes), torch.tensor(weights).reshape(-1, 1, 1)

    def compute_barycentric_coordinates(self, x: torch.Tensor):
        xi, eta = x[..., 0:1], x[..., 1:2]
        return torch.stack([1.0 - xi - eta, xi, eta], dim=-2)

    def compute_shape_functions(self, bar_coords: torch.Tensor, inv_map_jacobian: torch.Tensor):
        grads = self.barycentric_grad.to(inv_map_jacobian)  # (3,2): gradients of the barycentric coordinates
        if self.polynomial_order == 1:
            return bar_coords, grads @ inv_map_jacobian
        if self.polynomial_order == 2:
            # Lagrange P2 (reference element_tri.py:43-70): vertex functions l (2 l - 1), edge functions 4 l_a l_b on the
            # edges (1,2), (2,3), (3,1).  Host-side torch like the reference: no basis of either package assembles with
            # them (basis/basis.py:50-51 rejects order 2), only P1 is on the kernel path.
            lam = [bar_coords[..., k : k + 1, :] for k in range(3)]
            pairs = ((0, 1), (1, 2), (2, 0))
            values = [l * (2.0 * l - 1.0) for l in lam] + [4.0 * lam[a] * lam[b] for a, b in pairs]
            slopes = [(4.0 * lam[k] - 1.0) * grads[k : k + 1] for k in range(3)]
            slopes += [4.0 * (lam[b] * grads[a : a + 1] + lam[a] * grads[b : b + 1]) for a, b in pairs]
            return torch.cat(values, dim=-2), torch.cat(slopes, dim=-2) @ inv_map_jacobian
        raise NotImplementedError("Polynomial order not implemented")

    def compute_det_and_inv_map(self, map_jacobian: torch.Tensor):
        a, b = map_jacobian[..., 0:1, 0:1], map_jacobian[..., 0:1, 1:2]
        c, d = map_jacobian[..., 1:2, 0:1], map_jacobian[..., 1:2, 1:2]
        det = (a * d - b * c).unsqueeze(-3)  # signed, as in the reference (no abs)
        adj = torch.cat([torch.cat([d, -b], dim=-1), torch.cat([-c, a], dim=-1)], dim=-2).unsqueeze(-3)
        return det, adj / det
