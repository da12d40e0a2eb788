"""P1 basis on a discrete fracture network (reference torch_fem/basis/fracture_basis.py)."""

from __future__ import annotations

import torch

from .. import csr as csr_mod
from .. import ops
from ..tensordict_lite import TensorDict
from .abstract_basis import CellLayout, LazyParameters
from .basis import Basis
from .interior_edges_basis import InteriorEdgesFractureBasis


def _first_occurrence(inverse: torch.Tensor, n_groups: int) -> torch.Tensor:
    """Smallest flat index mapped to each group (the reference's scatter_reduce amin, :49-59)."""
    out = torch.full((n_groups,), inverse.numel() + 1, dtype=torch.int64, device=inverse.device)
    return out.scatter_reduce(0, inverse, torch.arange(inverse.numel(), device=inverse.device), reduce="amin")


class FractureBasis(Basis):
    """Fractures glued along traces: vertices with bit-identical 3-D coordinates share a DOF."""

    _EDGE_BASIS = InteriorEdgesFractureBasis

    def __init__(self, mesh, element):
        mesh = ops.place_mesh(mesh)
        self.global_triangulation = self._build_global_triangulation(mesh)
        super().__init__(mesh, element)

    def _build_global_triangulation(self, mesh) -> TensorDict:
        """Global numbering of vertices / triangles / edges and the trace sets (reference :28-129)."""
        coords2 = mesh["vertices", "coordinates"]
        n_f, n_v, _ = coords2.shape
        edges = mesh["edges", "vertices"]
        n_e = edges.shape[-2]
        device = coords2.device

        points3 = mesh["vertices", "coordinates_3d"].reshape(-1, 3)
        vertices_3d, to_global, multiplicity = torch.unique(points3, dim=0, return_inverse=True, return_counts=True)
        n_g = vertices_3d.shape[0]
        representative = _first_occurrence(to_global, n_g)
        offset_v = (torch.arange(n_f, device=device) * n_v).reshape(-1, 1, 1)
        triangles = to_global[mesh["cells", "vertices"].long() + offset_v].reshape(-1, 3)

        edges_global = to_global[edges.long() + offset_v].reshape(-1, 2)
        unique_edges, edge_to_global, edge_multiplicity = torch.unique(
            edges_global, dim=0, return_inverse=True, return_counts=True
        )
        trace_edges = torch.nonzero(edge_multiplicity > 1, as_tuple=True)[0]
        on_trace = torch.isin(edge_to_global, trace_edges).reshape(n_f, n_e)
        per_fracture = [torch.nonzero(row, as_tuple=True)[0] for row in on_trace]
        if len({int(t.numel()) for t in per_fracture}) == 1:
            traces_local = torch.stack(per_fracture, dim=0)  # (F, n_trace), as the reference reshapes it
        else:
            # fractures with different numbers of trace edges (e.g. a backbone crossed by several
            # planes): the reference's reshape(F, -1) cannot represent this; keep a padded table
            width = max(int(t.numel()) for t in per_fracture)
            traces_local = torch.full((n_f, width), -1, dtype=torch.int64, device=device)
            for f, t in enumerate(per_fracture):
                traces_local[f, : t.numel()] = t
        edge_representative = _first_occurrence(edge_to_global, unique_edges.shape[0])

        return TensorDict(
            vertices_3D=vertices_3d,
            vertices_2D=coords2.reshape(-1, 2)[representative],
            vertex_markers=mesh["vertices", "markers"].reshape(-1)[representative],
            triangles=triangles,
            edges=unique_edges,
            edge_markers=mesh["edges", "markers"].reshape(-1)[edge_representative],
            global2local_idx=to_global,
            local2global_idx=representative,
            traces__global_vertices_idx=torch.nonzero(multiplicity > 1, as_tuple=True)[0],
            traces_global_edges_idx=trace_edges,
            traces_local_edges_idx=traces_local,
        )

    def _compute_layout(self, mesh, element) -> CellLayout:
        if element.polynomial_order != 1:
            raise NotImplementedError("Polynomial order not implemented")
        coords = mesh["vertices", "coordinates"]
        conn = mesh["cells", "vertices"]
        n_f, n_v, _ = coords.shape
        n_c = conn.shape[1]
        self._flat_dofs = self.global_triangulation["triangles"].to(torch.int32).contiguous()
        self._n_dof_flat = self.global_triangulation["vertices_2D"].shape[-2]
        frac = (
            mesh["jacobian_fracture_map"].contiguous(),
            mesh["inv_jacobian_fracture_map"].contiguous(),
            mesh["det_jacobian_fracture_map"].reshape(-1).contiguous(),
            mesh["translation_vector"].reshape(-1, 3).contiguous(),
        )
        return CellLayout(
            coords.reshape(-1, 2).contiguous(), conn.to(torch.int32).reshape(-1, 3).contiguous(), n_c, n_v, (n_f, n_c), frac
        )

    def _compute_dofs(self, mesh, element):
        gt = self.global_triangulation
        boundary = torch.nonzero(gt["vertex_markers"] == 1)[:, 0]
        return gt["vertices_2D"], gt["triangles"], boundary, gt["vertices_2D"][gt["triangles"]]

    def _compute_basis_parameters(self, coords4global_dofs, global_dofs4elements, nodes4boundary_dofs):
        n_dof = coords4global_dofs.shape[-2]
        every = torch.arange(n_dof, device=coords4global_dofs.device)
        eager = {
            "bilinear_form_shape": (n_dof, n_dof),
            "linear_form_shape": (n_dof, 1),
            "inner_dofs": every[~torch.isin(every, nodes4boundary_dofs)],
            "nb_dofs": n_dof,
        }
        maps = lambda: csr_mod.coo_index_maps(global_dofs4elements)  # noqa: E731
        lazy = {"bilinear_form_idx": lambda: maps()[:2], "linear_form_idx": lambda: (maps()[2],)}
        return LazyParameters(eager, lazy)

    def _edge_interpolation_inputs(self, basis):
        mesh = basis.mesh
        lay = self._layout
        cells = mesh["interior_edges", "cells"]
        n_edge_per_mesh = cells.shape[-2]
        first = self.mesh["cells", "coordinates_3d"][..., 0, :].reshape(-1, 3).contiguous()
        inv = self._inv_map_jacobian.reshape(-1, 2, 3).contiguous()
        x_q = basis.integration_points.reshape(-1, basis.n_q, 3).contiguous()
        # NB the reference indexes the nodal vector with the fracture-LOCAL vertex ids here
        # (fracture_basis.py:229-231 gathers mesh["cells","vertices"]); kept for parity
        return cells.to(torch.int32).reshape(-1, 2).contiguous(), lay.conn, first, inv, x_q, n_edge_per_mesh, lay.n_el_per_mesh

    def _interpolation_nodes(self, basis=None):
        # own quadrature points: the interpolant gathers with the GLOBAL (glued) DOF ids, so the function is
        # evaluated at the global representative vertices.  Towards the edges the reference indexes the nodal
        # vector with fracture-LOCAL vertex ids (fracture_basis.py:229-231), kept as is.
        if basis is self:
            return self.global_triangulation["vertices_3D"]
        return self.mesh["vertices", "coordinates_3d"]
