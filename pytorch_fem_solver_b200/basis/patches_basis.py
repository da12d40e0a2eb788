"""Per-patch assembly: results are (P,5,5) / (P,5,1) (reference torch_fem/basis/patches_basis.py)."""

from __future__ import annotations

import torch

from .abstract_basis import AbstractBasis, CellLayout, LazyParameters


class PatchesBasis(AbstractBasis):
    """Every mesh of the batch is its own little FE space; nothing couples patches."""

    _tiled_residual_ok = False  # patches have their own one-launch kernel (tfem_batched_weak_residual)

    def __init__(self, mesh, element):
        self.nb_patches = mesh.batch_size()[0]
        super().__init__(mesh, element)
        self.patches_idx = torch.arange(self.nb_patches, device=self.device).unsqueeze(-1)

    def _compute_layout(self, mesh, element) -> CellLayout:
        if element.polynomial_order != 1:
            raise NotImplementedError("Polynomial order not implemented")
        coords = mesh["vertices", "coordinates"]
        conn = mesh["cells", "vertices"]
        n_p, n_v, _ = coords.shape
        n_c = conn.shape[1]
        flat_conn = conn.to(torch.int32).reshape(-1, 3).contiguous()
        offsets = (torch.arange(n_p, device=conn.device, dtype=torch.int32) * n_v).repeat_interleave(n_c)
        self._flat_dofs = (flat_conn + offsets[:, None]).contiguous()
        self._n_dof_flat = n_p * n_v
        self._n_local_dofs = n_v
        return CellLayout(coords.reshape(-1, 2).contiguous(), flat_conn, n_c, n_v, (n_p, n_c))

    def _compute_dofs(self, mesh, element):
        return (
            mesh["vertices", "coordinates"],
            mesh["cells", "vertices"],
            mesh["vertices", "markers"],
            mesh["cells", "coordinates"],
        )

    def _compute_basis_parameters(self, coords4global_dofs, global_dofs4elements, nodes4boundary_dofs):
        n_dof = coords4global_dofs.size(-2)
        n_loc = global_dofs4elements.size(-1)
        n_p = self.nb_patches
        patches_idx = torch.arange(n_p, device=global_dofs4elements.device).unsqueeze(-1)
        eager = {
            "bilinear_form_shape": (n_p, n_dof, n_dof),
            "linear_form_shape": (n_p, n_dof, 1),
            "inner_dofs": torch.nonzero(nodes4boundary_dofs != 1, as_tuple=True)[-2],
            "nb_dofs": n_dof,
        }
        lazy = {
            # reference :53-58: per patch, rows[9e+3i+j] = conn[e,j], cols[9e+3i+j] = conn[e,i]
            "bilinear_form_idx": lambda: (
                patches_idx,
                global_dofs4elements.repeat(1, 1, n_loc).reshape(n_p, -1),
                global_dofs4elements.repeat_interleave(n_loc, dim=-1).reshape(n_p, -1),
            ),
            "linear_form_idx": lambda: (patches_idx, global_dofs4elements.reshape(n_p, -1)),
        }
        return LazyParameters(eager, lazy)

    def reshape_for_assembly(self, local_matrices: torch.Tensor, form: str):
        if form == "bilinear":
            return local_matrices.reshape(self.nb_patches, -1)
        if form == "linear":
            return local_matrices.reshape(self.nb_patches, -1, 1)
        raise NotImplementedError(f"Unknown form type: {form}")

    def _matrix_result(self, values: torch.Tensor, layout=None) -> torch.Tensor:
        if layout == "values":
            return values
        pat = self.pattern
        n_loc = self._n_local_dofs
        out = torch.zeros(self.nb_patches * n_loc * n_loc, dtype=values.dtype, device=values.device)
        # flattened row r = p*n_loc + i and column c = p*n_loc + j land at p*n_loc^2 + i*n_loc + j
        out[pat.row_indices() * n_loc + pat.col.long() % n_loc] = values
        return out.reshape(self.nb_patches, n_loc, n_loc)

    def _vector_result(self, vec: torch.Tensor) -> torch.Tensor:
        return vec.reshape(self.nb_patches, self._n_local_dofs, 1)

    def reduce(self, tensor: torch.Tensor):
        """The single interior (centre) entry of every patch (reference :99-105)."""
        idx = self._basis_parameters["inner_dofs"]
        p = self.patches_idx.squeeze(-1)
        return tensor[p, idx, idx] if tensor.size(-1) != 1 else tensor[p, idx]
