"""Symbolic phase of the assembly: CSR pattern, COO->CSR permutation and the row-tile plan.

Integer, one-time work done with torch ops on whatever device the connectivity lives on
("index-map construction may stay torch"); the numeric kernels
that consume these structures are the hand-written CUDA ones.

Contract: the CSR pattern is the sorted unique set of
`(rows_idx, cols_idx)` of the reference's `bilinear_form_idx`
(reference basis/basis.py:72-75):  `rows[9e+3i+j] = conn[e,j]`, `cols[9e+3i+j] = conn[e,i]`.
Structural zeros are kept: the pattern never depends on values.
"""

from __future__ import annotations

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
        return torch.sparse_csr_tensor(self.crow, self.col, values, size=(self.n_dof, self.n_dof))


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


# ------------------------------------------------------------------------------------------------
# Row-tile plan of the fused kernel (tfem_tri_p1_assemble_csr)
# ------------------------------------------------------------------------------------------------


def _spread_bits16(v: torch.Tensor) -> torch.Tensor:
    v = v & 0xFFFF
    v = (v | (v << 8)) & 0x00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F
    v = (v | (v << 2)) & 0x33333333
    v = (v | (v << 1)) & 0x55555555
    return v


def morton_order(points: torch.Tensor) -> torch.Tensor:
    """Permutation sorting 2-D points along a Z-order curve (16 bits per axis)."""
    p = points[:, :2].to(torch.float64)
    lo = p.min(0).values
    span = (p.max(0).values - lo).clamp_min(1e-300)
    q = ((p - lo) / span * 65535.0).round().to(torch.int64)
    code = _spread_bits16(q[:, 0]) | (_spread_bits16(q[:, 1]) << 1)
    return torch.argsort(code, stable=True)


def block_tiles(points: torch.Tensor, rows_per_tile: int):
    """Tile id per row: square spatial blocks holding ~rows_per_tile points each.

    Returns (tile_of_row, n_tiles, largest tile)."""
    p = points[:, :2].to(torch.float64)
    n = p.shape[0]
    lo = p.min(0).values
    span = (p.max(0).values - lo).clamp_min(1e-300)
    area = float(span[0] * span[1])
    side = (area * rows_per_tile / max(n, 1)) ** 0.5
    nbx = max(int(round(float(span[0]) / side)), 1)
    nby = max(int(round(float(span[1]) / side)), 1)
    bx = ((p[:, 0] - lo[0]) / span[0] * nbx).floor().clamp(0, nbx - 1).long()
    by = ((p[:, 1] - lo[1]) / span[1] * nby).floor().clamp(0, nby - 1).long()
    used, tile_of_row, counts = torch.unique(by * nbx + bx, return_inverse=True, return_counts=True)
    return tile_of_row, int(used.shape[0]), int(counts.max().item())


@dataclass
class TilePlan:
    """Device arrays of struct tfem_tile_plan (one packed blob per tile) plus bookkeeping."""

    n_tiles: int
    tile_off: torch.Tensor  # (n_tiles+1,) int32 word offsets, multiples of 4
    blob: torch.Tensor  # int32 words, layout in include/tfem_b200.h
    max_vert: int
    max_elem: int
    max_out: int
    max_blob_words: int
    max_rows: int
    halo_factor: float  # tile elements / mesh elements (1.0 = every element computed once)
    index_bytes: int  # bytes of plan data the kernel reads per launch

    def c_struct(self) -> "_lib.TilePlan":
        s = _lib.TilePlan()
        s.n_tiles = self.n_tiles
        s.tile_off = self.tile_off.data_ptr()
        s.blob = self.blob.data_ptr()
        s.max_vert, s.max_elem, s.max_out, s.max_blob_words = self.max_vert, self.max_elem, self.max_out, self.max_blob_words
        return s

    def to(self, device) -> "TilePlan":
        moved = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return TilePlan(**moved)

    def sections(self, tile: int) -> dict:
        """Decode one tile's blob into named integer arrays (tests / debugging)."""
        off = self.tile_off.cpu().numpy().astype("int64")
        words = self.blob[int(off[tile]) : int(off[tile + 1])].cpu().numpy().astype("int64") & 0xFFFFFFFF
        n_vert, n_elem, n_rows, n_runs, n_corner, n_out = (int(w) for w in words[:6])
        pad4 = lambda n: (n + 3) & ~3  # noqa: E731
        out, pos = {"n_out": n_out}, 8
        for name, n in (("vert", n_vert), ("elem", n_elem), ("row_id", n_rows), ("row_meta", n_rows),
                        ("row_cptr", n_rows + 1), ("corner", n_corner), ("run_start", n_runs), ("run_meta", n_runs)):
            out[name] = words[pos : pos + n]
            pos += pad4(n)
        return out


def _ptr_from_sorted(group: torch.Tensor, n_groups: int) -> torch.Tensor:
    out = torch.zeros(n_groups + 1, dtype=torch.int64, device=group.device)
    out[1:] = torch.cumsum(torch.bincount(group, minlength=n_groups), 0)
    return out


def build_tile_plan(
    geom_conn: torch.Tensor,
    dof_conn: torch.Tensor,
    pattern: CsrPattern,
    row_points: torch.Tensor | None = None,
    rows_per_tile: int = 216,
    ordering: str = "block",
) -> TilePlan:
    """Partition CSR rows into tiles and precompute everything the fused kernel gathers.

    geom_conn (N,3): rows of `coords` of each element's vertices (batch offsets applied).
    dof_conn  (N,3): global DOF (CSR row/col) of each element vertex.
    row_points (n_dof,2+): a position per DOF, only used to cluster rows spatially; with
    None, tiles are runs of consecutive DOF ids.
    """
    device = dof_conn.device
    gconn = geom_conn.reshape(-1, 3).long()
    dconn = dof_conn.reshape(-1, 3).long()
    n_el = dconn.shape[0]
    n_dof = pattern.n_dof
    n_gv = int(gconn.max().item()) + 1 if n_el else 1
    crow = pattern.crow.long()

    # 1. rows -> tiles.  Only rows some element touches are clustered; rows without elements
    #    (isolated vertices, ghost columns of a multi-GPU owner) carry no work and are dealt out
    #    evenly afterwards so that their (zero) load entry is still written.
    active = torch.zeros(n_dof, dtype=torch.bool, device=device)
    active[dconn.reshape(-1)] = True
    active_rows = torch.nonzero(active, as_tuple=True)[0]
    n_active = int(active_rows.numel())
    tile_of_active = None
    if row_points is not None and ordering == "block" and n_active:
        tile_of_active, n_tiles, largest = block_tiles(row_points[active_rows], rows_per_tile)
        if largest > 2 * rows_per_tile:  # strongly graded mesh: fall back to balanced Z-order chunks
            tile_of_active = None
    if tile_of_active is None:
        if row_points is not None and ordering != "natural" and n_active:
            order = morton_order(row_points[active_rows])
        else:
            order = torch.arange(n_active, device=device)
        tile_of_active = torch.empty(n_active, dtype=torch.int64, device=device)
        tile_of_active[order] = torch.arange(n_active, device=device) // rows_per_tile
        n_tiles = max((n_active + rows_per_tile - 1) // rows_per_tile, 1)
    tile_of_row = torch.empty(n_dof, dtype=torch.int64, device=device)
    tile_of_row[active_rows] = tile_of_active
    idle_rows = torch.nonzero(~active, as_tuple=True)[0]
    tile_of_row[idle_rows] = torch.arange(idle_rows.numel(), device=device) % n_tiles

    # 2. (tile, element) incidences, tile-major / element ascending
    t_of_corner = tile_of_row[dconn]  # (N,3)
    pair_keys = torch.unique((t_of_corner * n_el + torch.arange(n_el, device=device)[:, None]).reshape(-1))
    pair_tile = torch.div(pair_keys, n_el, rounding_mode="floor")
    pair_elem = pair_keys - pair_tile * n_el
    elem_ptr = _ptr_from_sorted(pair_tile, n_tiles)

    # 3. (tile, geometry vertex) incidences and tile-local connectivity
    pair_verts = gconn[pair_elem]  # (n_pairs,3)
    vkeys_all = pair_tile[:, None] * n_gv + pair_verts
    vert_keys = torch.unique(vkeys_all.reshape(-1))
    vert_tile = torch.div(vert_keys, n_gv, rounding_mode="floor")
    tile_vert = vert_keys - vert_tile * n_gv
    vert_ptr = _ptr_from_sorted(vert_tile, n_tiles)
    local_v = torch.searchsorted(vert_keys, vkeys_all.reshape(-1)).reshape(-1, 3) - vert_ptr[pair_tile][:, None]
    max_vert = int((vert_ptr[1:] - vert_ptr[:-1]).max().item())
    max_elem = int((elem_ptr[1:] - elem_ptr[:-1]).max().item())
    if max_vert > 1024 or max_elem > 4096:
        raise ValueError(f"tile too large (vertices {max_vert} > 1024 or elements {max_elem} > 4096): lower rows_per_tile")
    tile_elem = local_v[:, 0] | (local_v[:, 1] << 10) | (local_v[:, 2] << 20)

    # 4. rows of each tile, ascending row id; slot of each row inside the tile's output image
    row_sorted = torch.argsort(tile_of_row * n_dof + torch.arange(n_dof, device=device))
    row_tile = tile_of_row[row_sorted]
    row_ptr = _ptr_from_sorted(row_tile, n_tiles)
    row_len = crow[row_sorted + 1] - crow[row_sorted]
    csum = torch.cumsum(row_len, 0) - row_len  # exclusive, over the tile-ordered rows
    tile_first = csum[row_ptr[:-1].clamp_max(max(n_dof - 1, 0))]
    out_base = csum - tile_first[row_tile]
    tile_out = torch.zeros(n_tiles, dtype=torch.int64, device=device).index_add_(0, row_tile, row_len)
    max_out = int(tile_out.max().item())
    if max_out > 65535 or int(row_len.max().item()) > 255:
        raise ValueError("tile output image too large for 16-bit slots: lower rows_per_tile")
    diag_pos = torch.searchsorted(pattern.keys, row_sorted * n_dof + row_sorted) - crow[row_sorted]
    diag_pos = torch.where(row_len > 0, diag_pos, torch.zeros_like(diag_pos))
    row_meta = out_base | (diag_pos << 16)
    row_rank = torch.empty(n_dof, dtype=torch.int64, device=device)
    row_rank[row_sorted] = torch.arange(n_dof, device=device)

    # 5. corners (row, incident element), grouped by tile-ordered row, element ascending
    flat = torch.arange(3 * n_el, device=device)
    corner_row = dconn.reshape(-1)
    corder = torch.argsort(row_rank[corner_row] * (3 * n_el) + flat)
    c_e = torch.div(corder, 3, rounding_mode="floor")
    c_k = corder - 3 * c_e
    c_row = corner_row[corder]
    c_tile = tile_of_row[c_row]
    el_local = torch.searchsorted(pair_keys, c_tile * n_el + c_e) - elem_ptr[c_tile]
    col_a = dconn[c_e, (c_k + 1) % 3]
    col_b = dconn[c_e, (c_k + 2) % 3]
    pos_a = torch.searchsorted(pattern.keys, c_row * n_dof + col_a) - crow[c_row]
    pos_b = torch.searchsorted(pattern.keys, c_row * n_dof + col_b) - crow[c_row]
    corner = el_local | (c_k << 12) | (pos_a << 16) | (pos_b << 24)
    row_corner_ptr = _ptr_from_sorted(row_rank[corner_row], n_dof)

    # 6. runs of consecutive rows inside a tile (contiguous CSR ranges)
    idx = torch.arange(n_dof, device=device)
    new_run = torch.ones(n_dof, dtype=torch.bool, device=device)
    if n_dof > 1:
        new_run[1:] = (row_tile[1:] != row_tile[:-1]) | (row_sorted[1:] != row_sorted[:-1] + 1)
    run_first = idx[new_run]
    run_last = torch.cat([run_first[1:], torch.tensor([n_dof], device=device)]) - 1
    run_start = crow[row_sorted[run_first]]
    run_len = crow[row_sorted[run_last] + 1] - run_start
    run_meta = out_base[run_first] | (run_len << 16)
    run_ptr = _ptr_from_sorted(row_tile[run_first], n_tiles)

    # 7. pack everything a tile needs into one 16 B aligned blob (a single TMA bulk copy per CTA)
    def pad4(t):
        return (t + 3) & ~3

    tiles = torch.arange(n_tiles, device=device)
    n_v, n_e, n_r, n_u = (p[1:] - p[:-1] for p in (vert_ptr, elem_ptr, row_ptr, run_ptr))
    corner_base = row_corner_ptr[row_ptr[:-1]]
    n_c = row_corner_ptr[row_ptr[1:]] - corner_base
    sizes = [torch.full_like(n_v, 8), pad4(n_v), pad4(n_e), pad4(n_r), pad4(n_r), pad4(n_r + 1), pad4(n_c), pad4(n_u), pad4(n_u)]
    tile_words = sum(sizes)
    tile_off = torch.zeros(n_tiles + 1, dtype=torch.int64, device=device)
    tile_off[1:] = torch.cumsum(tile_words, 0)
    total_words = int(tile_off[-1].item())
    if total_words >= 2**31:
        raise ValueError("tile plan too large for 32-bit word offsets")
    starts = [tile_off[:-1]]
    for size in sizes[:-1]:
        starts.append(starts[-1] + size)
    s_hdr, s_vert, s_elem, s_rid, s_rmeta, s_rcptr, s_corner, s_rstart, s_rmeta2 = starts
    blob = torch.zeros(total_words, dtype=torch.int64, device=device)
    # a vertex from the middle of each tile's (sorted) vertex list; tiles without elements get vertex 0
    middle = (vert_ptr[:-1] + torch.div(n_v, 2, rounding_mode="floor")).clamp_max(max(tile_vert.numel() - 1, 0))
    tile_base = torch.where(n_v > 0, tile_vert[middle] if tile_vert.numel() else torch.zeros_like(n_v), torch.zeros_like(n_v))
    for k, field in enumerate((n_v, n_e, n_r, n_u, n_c, tile_out, tile_base)):
        blob[s_hdr + k] = field
    blob[s_vert[vert_tile] + torch.arange(vert_tile.numel(), device=device) - vert_ptr[vert_tile]] = tile_vert
    blob[s_elem[pair_tile] + torch.arange(pair_tile.numel(), device=device) - elem_ptr[pair_tile]] = tile_elem
    row_local = torch.arange(n_dof, device=device) - row_ptr[row_tile]
    blob[s_rid[row_tile] + row_local] = row_sorted
    blob[s_rmeta[row_tile] + row_local] = row_meta
    cptr_tile = torch.repeat_interleave(tiles, n_r + 1)
    cptr_local = torch.arange(cptr_tile.numel(), device=device) - (row_ptr[:-1] + tiles)[cptr_tile]
    blob[s_rcptr[cptr_tile] + cptr_local] = row_corner_ptr[row_ptr[cptr_tile] + cptr_local] - corner_base[cptr_tile]
    blob[s_corner[c_tile] + torch.arange(c_tile.numel(), device=device) - corner_base[c_tile]] = corner
    run_tile = row_tile[run_first]
    run_local = torch.arange(run_tile.numel(), device=device) - run_ptr[run_tile]
    blob[s_rstart[run_tile] + run_local] = run_start
    blob[s_rmeta2[run_tile] + run_local] = run_meta

    blob32 = _wrap_u32(blob)
    tile_off32 = tile_off.to(torch.int32).contiguous()
    return TilePlan(
        n_tiles=n_tiles,
        tile_off=tile_off32,
        blob=blob32,
        max_vert=max_vert,
        max_elem=max_elem,
        max_out=max_out,
        max_blob_words=int(tile_words.max().item()),
        max_rows=int(n_r.max().item()),
        halo_factor=float(pair_keys.shape[0]) / max(n_el, 1),
        index_bytes=4 * (blob32.numel() + tile_off32.numel()),
    )


def _wrap_u32(t: torch.Tensor) -> torch.Tensor:
    """Store an unsigned 32-bit pattern in an int32 tensor (two's complement wrap)."""
    t = t & 0xFFFFFFFF
    return torch.where(t >= 2**31, t - 2**32, t).to(torch.int32).contiguous()
