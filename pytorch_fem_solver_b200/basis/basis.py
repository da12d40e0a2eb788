"""P1 basis on a single planar mesh (reference torch_fem/basis/basis.py)."""

from __future__ import annotations

from typing import Callable, Optional

import torch

from .. import csr as csr_mod
from .. import ops
from .abstract_basis import AbstractBasis, CellLayout, LazyParameters
from .interior_edges_basis import InteriorEdgesBasis


class Basis(AbstractBasis):
    """DOFs = mesh vertices; global operators are (N_v, N_v)."""

    def _compute_layout(self, mesh, element) -> CellLayout:
        if element.polynomial_order != 1:
            raise NotImplementedError("Polynomial order not implemented")
        coords = mesh["vertices", "coordinates"].contiguous()
        conn = mesh["cells", "vertices"].to(torch.int32).contiguous()
        self._flat_dofs = conn
        self._n_dof_flat = coords.shape[-2]
        return CellLayout(coords, conn, conn.shape[0], coords.shape[0], (conn.shape[0],))

    def _compute_dofs(self, mesh, element):
        return (
            mesh["vertices", "coordinates"],
            mesh["cells", "vertices"],
            mesh["vertices", "markers"],
            mesh["cells", "coordinates"],
        )

    def _compute_basis_parameters(self, coords4global_dofs, global_dofs4elements, nodes4boundary_dofs):
        n_dof = coords4global_dofs.size(-2)
        eager = {
            "bilinear_form_shape": (n_dof, n_dof),
            "linear_form_shape": (n_dof, 1),
            "inner_dofs": torch.nonzero(nodes4boundary_dofs != 1, as_tuple=True)[-2],
            "nb_dofs": n_dof,
        }
        maps = lambda: csr_mod.coo_index_maps(global_dofs4elements)  # noqa: E731
        lazy = {
            "bilinear_form_idx": lambda: maps()[:2],
            "linear_form_idx": lambda: (maps()[2],),
        }
        return LazyParameters(eager, lazy)

    # ------------------------------------------------------------------ interpolation
    def _edge_interpolation_inputs(self, basis):
        mesh = basis.mesh
        lay = self._layout
        cells = mesh["interior_edges", "cells"].to(torch.int32).reshape(-1, 2).contiguous()
        first = self.mesh["cells", "coordinates"][..., 0, :].reshape(-1, 2).contiguous()
        inv = self._inv_map_jacobian.reshape(-1, 2, 2).contiguous()
        x_q = basis.integration_points.reshape(-1, basis.n_q, 2).contiguous()
        return cells, lay.conn, first, inv, x_q, cells.shape[0], lay.n_el_per_mesh

    def _edge_scatter_maps(self, basis, cells, conn, n_edge_per_mesh, n_el_per_mesh):
        """(seg, perm, inverse) of the (edge, side, local vertex) -> DOF relation: adjoint of edge interpolation."""
        key = ("edge_maps", id(basis))
        if key not in self._scatter_inverse:
            n_edge = cells.shape[0]
            mesh_of_edge = torch.arange(n_edge, device=cells.device) // n_edge_per_mesh
            global_cell = (mesh_of_edge[:, None] * n_el_per_mesh + cells.long()).reshape(-1)
            dof = conn.long()[global_cell].reshape(-1)  # (2E*3,)
            n_dof = self.n_dof_flat
            order = torch.argsort(dof, stable=True)
            seg = torch.zeros(n_dof + 1, dtype=torch.int64, device=cells.device)
            seg[1:] = torch.cumsum(torch.bincount(dof, minlength=n_dof), 0)
            self._scatter_inverse[key] = (seg.to(torch.int32), order.to(torch.int32), dof.to(torch.int32).contiguous())
        return self._scatter_inverse[key]

    def _interpolate_values(self, basis, nodal: torch.Tensor):
        u = nodal.to(self.dtype).reshape(-1).contiguous()
        needs_grad = u.requires_grad
        if basis is self:
            maps = (None, None, None)
            if needs_grad:
                pat = self.pattern
                maps = (pat.lin_seg, pat.lin_perm, self._inverse("linear"))
            val, grad = ops.interp_cells(u, self._dof_conn_flat(), self.v_grad.reshape(-1, 3, self._layout.d).contiguous(),
                                         self._element.integration_order, *maps)
            lead = self._layout.lead
            return val.reshape(*lead, self.n_q, 1, 1), grad.reshape(*lead, 1, 1, self._layout.d)
        cells, conn, first, inv, x_q, n_edge_per_mesh, n_el_per_mesh = self._edge_interpolation_inputs(basis)
        maps = self._edge_scatter_maps(basis, cells, conn, n_edge_per_mesh, n_el_per_mesh) if needs_grad else (None, None, None)
        val, grad = ops.interp_edges(u, cells, conn, first, inv, x_q, n_edge_per_mesh, n_el_per_mesh, *maps)
        lead = basis._layout.lead
        d = x_q.shape[-1]
        return val.reshape(*lead, 2, basis.n_q, 1, 1), grad.reshape(*lead, 2, 1, 1, d)

    _EDGE_BASIS = InteriorEdgesBasis

    def interpolate(self, basis: AbstractBasis, tensor: Optional[torch.Tensor] = None):
        """FE interpolant and its gradient at another basis' points (reference :98-177).

        With `tensor` (nodal values, (n_dof, 1)) returns `(values, gradients)`; without, returns
        two closures taking a function of the node coordinates."""
        if basis is not self and basis.__class__ is not self._EDGE_BASIS:
            raise NotImplementedError("Interpolation for this basis not implemented")
        if tensor is not None:
            return self._interpolate_values(basis, tensor)
        nodes = self._interpolation_nodes(basis)

        def interpolator(function: Callable[[torch.Tensor], torch.Tensor]) -> torch.Tensor:
            return self._interpolate_values(basis, function(nodes))[0]

        def interpolator_grad(function: Callable[[torch.Tensor], torch.Tensor]) -> torch.Tensor:
            return self._interpolate_values(basis, function(nodes))[1]

        return interpolator, interpolator_grad

    def _interpolation_nodes(self, basis=None):
        return self._coords4global_dofs
