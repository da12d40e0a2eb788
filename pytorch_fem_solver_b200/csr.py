"""Symbolic phase of the assembly: CSR pattern and the stable COO->CSR permutation.

Integer, one-time work done with torch ops on whatever device the connectivity lives on
("index-map construction may stay torch"); the numeric kernels
that consume these structures are the hand-written CUDA ones.

Contract: the CSR pattern is the sorted unique set of
`(rows_idx, cols_idx)` of the reference's `bilinear_form_idx`
(reference basis/basis.py:72-75):  `rows[9e+3i+j] = conn[e,j]`, `cols[9e+3i+j] = conn[e,i]`.
Structural zeros are kept: the pattern never depends on values.
"""

from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch

from . import _lib


@dataclass
class CsrPattern:
    """crow/col of the global operator plus the stable COO->CSR permutation."""

    n_dof: int
    n_el: int
    nnz: int
    crow: torch.Tensor  # (n_dof+1,) int32
    col: torch.Tensor  # (nnz,) int32
    seg: torch.Tensor  # (nnz+1,) int32: contributions of entry p are perm[seg[p]:seg[p+1]]
    perm: torch.Tensor  # (9*n_el,) int32 flat COO indices, stable-sorted by (row, col)
    lin_seg: torch.Tensor  # (n_dof+1,) int32
    lin_perm: torch.Tensor  # (3*n_el,) int32 flat indices of linear_form_idx sorted by DOF
    keys: torch.Tensor = field(repr=False, default=None)  # (nnz,) int64 row*n_dof+col, sorted

    def to(self, device) -> "CsrPattern":
        moved = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return CsrPattern(**moved)

    def row_indices(self) -> torch.Tensor:
        """Row of every stored entry (for conversions)."""
        counts = (self.crow[1:] - self.crow[:-1]).long()
        return torch.repeat_interleave(torch.arange(self.n_dof, device=self.crow.device), counts)

    def to_dense(self, values: torch.Tensor) -> torch.Tensor:
        out = torch.zeros((self.n_dof, self.n_dof), dtype=values.dtype, device=values.device)
        out[self.row_indices(), self.col.long()] = values
        return out

    def to_sparse_csr(self, values: torch.Tensor) -> torch.Tensor:
        return torch.sparse_csr_tensor(self.crow, self.col, values, size=(self.n_dof, self.n_dof), device=values.device)


def coo_index_maps(dof_conn: torch.Tensor):
    """The reference's index maps (basis/basis.py:72-77), for any leading batch shape."""
    conn = dof_conn.reshape(-1, dof_conn.shape[-1])
    n_loc = conn.shape[-1]
    rows = conn.repeat(1, n_loc).reshape(-1)
    cols = conn.repeat_interleave(n_loc, dim=-1).reshape(-1)
    return rows, cols, conn.reshape(-1)


def _coo_keys(conn: torch.Tensor, n_dof: int) -> torch.Tensor:
    n_el = conn.shape[0]
    if conn.is_cuda and conn.dtype == torch.int32 and n_el > 0:
        keys = torch.empty(9 * n_el, dtype=torch.int64, device=conn.device)
        _lib.call("tfem_coo_keys", None, conn.device, n_el, conn.data_ptr(), n_dof, keys.data_ptr())
        return keys
    rows, cols, _ = coo_index_maps(conn)
    return rows.long() * n_dof + cols.long()


def build_pattern(dof_conn: torch.Tensor, n_dof: int, extra_keys: torch.Tensor | None = None) -> CsrPattern:
    """Sorted-unique CSR pattern + stable permutation of the 9*n_el COO entries.

    `extra_keys` (row*n_dof+col) adds structural entries that receive no local contribution;
    the multi-GPU owner of an interface row uses them for columns only other ranks touch."""
    conn = dof_conn.reshape(-1, 3).contiguous()
    n_el = conn.shape[0]
    if 9 * n_el >= 2**31:
        raise ValueError("mesh too large for 32-bit COO indices")
    device = conn.device
    if conn.is_cuda and n_el > 0 and (extra_keys is None or not extra_keys.numel()) and os.environ.get("TFEM_SYMBOLIC", "native") == "native":
        return _build_pattern_native(conn.to(torch.int32).contiguous(), n_dof)
    keys = _coo_keys(conn.to(torch.int32) if conn.is_cuda else conn, n_dof)
    sorted_keys, perm = torch.sort(keys, stable=True)
    uniq, counts = torch.unique_consecutive(sorted_keys, return_counts=True)
    if extra_keys is not None and extra_keys.numel():
        # structural entries with no local contribution (rows completed by other ranks)
        merged = torch.unique(torch.cat([uniq, extra_keys.to(uniq)]))
        merged_counts = torch.zeros(merged.shape[0], dtype=counts.dtype, device=device)
        merged_counts[torch.searchsorted(merged, uniq)] = counts
        uniq, counts = merged, merged_counts
    nnz = int(uniq.shape[0])
    seg = torch.zeros(nnz + 1, dtype=torch.int64, device=device)
    seg[1:] = torch.cumsum(counts, 0)
    row_of = torch.div(uniq, n_dof, rounding_mode="floor")
    col = (uniq - row_of * n_dof).to(torch.int32)
    crow = torch.zeros(n_dof + 1, dtype=torch.int64, device=device)
    crow[1:] = torch.cumsum(torch.bincount(row_of, minlength=n_dof), 0)

    form = conn.reshape(-1).long()
    _, lin_perm = torch.sort(form, stable=True)
    lin_seg = torch.zeros(n_dof + 1, dtype=torch.int64, device=device)
    lin_seg[1:] = torch.cumsum(torch.bincount(form, minlength=n_dof), 0)
    return CsrPattern(
        n_dof=n_dof,
        n_el=n_el,
        nnz=nnz,
        crow=crow.to(torch.int32),
        col=col,
        seg=seg.to(torch.int32),
        perm=perm.to(torch.int32),
        lin_seg=lin_seg.to(torch.int32),
        lin_perm=lin_perm.to(torch.int32),
        keys=uniq,
    )


def _build_pattern_native(conn: torch.Tensor, n_dof: int) -> CsrPattern:
    """`build_pattern` in one C-ABI call (`tfem_csr_symbolic`: CUB radix sort / run-length encode / scans on the
    device); bit-identical to the torch program above (tests/test_kernels_gpu.py::test_native_symbolic_phase)."""
    import ctypes

    n_el = conn.shape[0]
    device = conn.device
    lib = _lib.load()
    need = ctypes.c_int64()
    if lib.tfem_csr_symbolic_workspace(n_el, n_dof, ctypes.byref(need)) != 0:
        raise _lib.TfemError("tfem_csr_symbolic_workspace failed")
    i32 = dict(dtype=torch.int32, device=device)
    workspace = torch.empty(need.value, dtype=torch.uint8, device=device)
    crow, lin_seg = torch.empty(n_dof + 1, **i32), torch.empty(n_dof + 1, **i32)
    col, seg, perm = torch.empty(9 * n_el, **i32), torch.empty(9 * n_el + 1, **i32), torch.empty(9 * n_el, **i32)
    lin_perm = torch.empty(3 * n_el, **i32)
    keys = torch.empty(9 * n_el, dtype=torch.int64, device=device)
    nnz_dev = torch.zeros(1, dtype=torch.int64, device=device)
    _lib.call("tfem_csr_symbolic", None, device, n_el, conn.data_ptr(), n_dof, workspace.data_ptr(), need.value, crow.data_ptr(), col.data_ptr(),
              seg.data_ptr(), perm.data_ptr(), lin_seg.data_ptr(), lin_perm.data_ptr(), keys.data_ptr(), nnz_dev.data_ptr())
    nnz = int(nnz_dev.item())  # the one synchronisation of the (one-time) symbolic phase
    return CsrPattern(n_dof=n_dof, n_el=n_el, nnz=nnz, crow=crow, col=col[:nnz].clone(), seg=seg[: nnz + 1].clone(), perm=perm,
                      lin_seg=lin_seg, lin_perm=lin_perm, keys=keys[:nnz].clone())


def build_tile_plan(*args, **kwargs):
    """Row-tile plan of the fused kernel; see `tileplan.build_tile_plan`."""
    from . import tileplan

    return tileplan.build_tile_plan(*args, **kwargs)
