"""1-D quadrature on interior edges for jump terms
(reference torch_fem/basis/interior_edges_basis.py and interior_edges_fracture_basis.py)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Optional, Tuple

import torch

from .. import forms, ops
from .abstract_basis import AbstractBasis


@dataclass
class EdgeLayout:
    edge_coords: torch.Tensor  # (E_total, 2, 2)
    n_edge_per_mesh: int
    lead: Tuple[int, ...]
    frac: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None  # jac, det, t

    @property
    def coords(self):
        return self.edge_coords

    @property
    def d(self) -> int:
        return 3 if self.frac is not None else 2

    @property
    def n_total(self) -> int:
        return self.edge_coords.shape[0]


class InteriorEdgesBasis(AbstractBasis):
    """Quadrature points / weights on the interior edges of a mesh."""

    def __init__(self, mesh, element):
        self._element = element
        self.mesh = ops.place_mesh(mesh)
        self._layout = self._compute_layout(self.mesh, element)
        nodes = element.gaussian_nodes.to(device=self.device, dtype=self.dtype)
        self.v = element.compute_barycentric_coordinates(nodes)  # (q, 2)
        lay = self._layout
        frac = lay.frac if lay.frac is not None else (None, None, None)
        inv_jac, v_grad, x_q, dx = ops.edge_geometry(lay.edge_coords, lay.n_edge_per_mesh, element.integration_order, *frac)
        q = x_q.shape[1]
        self._geometry = {
            "v_grad": v_grad.reshape(*lay.lead, 2, 1),
            "integration_points": x_q.reshape(*lay.lead, 1, q, lay.d),
            "_dx": dx.reshape(*lay.lead, q, 1, 1),
            "_inv_map_jacobian": inv_jac.reshape(*lay.lead, 1, 1),
        }
        # the reference fills DOF tables with the cell data of the mesh ("NOT CORRECT", its own
        # comment at interior_edges_basis.py:22); only the shapes matter to callers
        self._coords4global_dofs = self.mesh["vertices", "coordinates"]
        self._global_dofs4elements = self.mesh["cells", "vertices"]
        self._nodes4boundary_dofs = self.mesh["vertices", "markers"]
        self._coords4elements = self.mesh["cells", "coordinates"]
        n_dof = self._coords4global_dofs.size(-2)
        self._basis_parameters = {
            "bilinear_form_shape": (n_dof, n_dof),
            "linear_form_shape": (n_dof, 1),
            "inner_dofs": torch.nonzero(self._nodes4boundary_dofs != 1, as_tuple=True)[-2],
            "nb_dofs": n_dof,
        }

    def _compute_layout(self, mesh, element) -> EdgeLayout:
        if element.polynomial_order != 1:
            raise NotImplementedError("Polynomial order not implemented")
        x = mesh["interior_edges", "coordinates"]
        return EdgeLayout(x.reshape(-1, 2, 2).contiguous(), x.shape[-3], tuple(x.shape[:-2]))

    def _compute_dofs(self, mesh, element):  # pragma: no cover - tables are filled in __init__
        raise NotImplementedError

    def _compute_basis_parameters(self, *args):  # pragma: no cover
        raise NotImplementedError

    def integrate_functional(self, function: Callable[..., torch.Tensor], *args: Any, **kwargs: Any) -> torch.Tensor:
        """Per-edge integral (reference abstract_basis.py:65-72); `forms.Jump` runs fused."""
        if isinstance(function, forms.Jump) and len(args) >= 2 and not kwargs:
            normal, size = args[0], args[1]
            grad_edges = args[2] if len(args) > 2 else function.grad_edges
            lay = self._layout
            d = grad_edges.shape[-1]
            eta = ops.edge_jump(
                grad_edges.to(self.dtype).reshape(lay.n_total, 2, d).contiguous(),
                normal.to(self.dtype).expand(*lay.lead, 1, 1, d).reshape(lay.n_total, d).contiguous(),
                size.to(self.dtype).expand(*lay.lead, 1, 1, 1).reshape(lay.n_total).contiguous(),
                self._dx.reshape(lay.n_total, self.n_q).contiguous(),
            )
            return eta.reshape(*lay.lead, 1)
        return super().integrate_functional(function, *args, **kwargs)

    def integrate_bilinear_form(self, *args, **kwargs):
        raise NotImplementedError("edge bases carry no DOFs of their own; use integrate_functional")

    integrate_linear_form = integrate_bilinear_form


class InteriorEdgesFractureBasis(InteriorEdgesBasis):
    """Interior edges of stacked fracture meshes: 3-D points, dx scaled by det J_f."""

    def _compute_layout(self, mesh, element) -> EdgeLayout:
        lay = super()._compute_layout(mesh, element)
        lay.frac = (
            mesh["jacobian_fracture_map"].contiguous(),
            mesh["det_jacobian_fracture_map"].reshape(-1).contiguous(),
            mesh["translation_vector"].reshape(-1, 3).contiguous(),
        )
        return lay
