"""Row-tile plan of the fused assembly kernel (`tfem_tri_p1_assemble_csr`).

Integer, one-time set-up in torch (any device).  CSR rows are clustered into tiles; for every tile
three packed, 16 B-aligned blobs are produced (layouts in include/tfem_b200.h):

  E ("early")  what the producer warp and the integration phase need: header, the tile's
               vertices (rows of `coords`) and its tile-local connectivity;
  LA ("late", entries)  segments of consecutive CSR slots and for every CSR entry of the tile ONE
               word holding its (at most two) contributions as (local-matrix slot, element)
               codes in increasing element order;
  LB ("late", rows)  owned row ids and for every row fixed-size chunks listing the elements of
               its load / diagonal entry.

Each blob is fetched by one TMA bulk copy, so per-tile sections are padded to whole 16 B units.
"""

from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import _lib
from .csr import CsrPattern

HEADER_WORDS = 12


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
    """Tile id per point: square spatial blocks holding ~rows_per_tile points each.

    Returns (tile_of_point, n_tiles, largest tile)."""
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
    used, tile_of_point, counts = torch.unique(by * nbx + bx, return_inverse=True, return_counts=True)
    return tile_of_point, int(used.shape[0]), int(counts.max().item())


def _pad4(t):
    return (t + 3) & ~3


def _ptr(group: torch.Tensor, n_groups: int) -> torch.Tensor:
    out = torch.zeros(n_groups + 1, dtype=torch.int64, device=group.device)
    out[1:] = torch.cumsum(torch.bincount(group, minlength=n_groups), 0)
    return out


def _excl_cumsum(t: torch.Tensor) -> torch.Tensor:
    return torch.cumsum(t, 0) - t


def _wrap_u32(t: torch.Tensor) -> torch.Tensor:
    """Store an unsigned 32-bit pattern in an int32 tensor (two's complement wrap)."""
    t = t & 0xFFFFFFFF
    return torch.where(t >= 2**31, t - 2**32, t).to(torch.int32).contiguous()


@dataclass
class _Section:
    name: str
    bits: int  # 32, 16 or 8
    count: torch.Tensor  # (n_tiles,) items per tile
    tile: torch.Tensor  # (n_items,) tile of each item
    local: torch.Tensor  # (n_items,) index of the item inside its tile
    value: torch.Tensor  # (n_items,) non-negative, < 2**bits


def _pack(sections: list, n_tiles: int, device):
    """Lay the sections of every tile out back to back; returns (tile_off int32, blob int32, words per tile)."""
    words = [_pad4((s.count * s.bits + 31) // 32) for s in sections]
    per_tile = sum(words)
    tile_off = torch.zeros(n_tiles + 1, dtype=torch.int64, device=device)
    tile_off[1:] = torch.cumsum(per_tile, 0)
    total = int(tile_off[-1].item())
    if total >= 2**31:
        raise ValueError("tile plan too large for 32-bit word offsets")
    blob = torch.zeros(total, dtype=torch.int64, device=device)
    start = tile_off[:-1].clone()
    for s, w in zip(sections, words):
        if s.value.numel():
            per_word = 32 // s.bits
            if int(s.value.max().item()) >= 2**s.bits or int(s.value.min().item()) < 0:
                raise ValueError(f"tile plan section {s.name}: value does not fit {s.bits} bits")
            pos = start[s.tile] + torch.div(s.local, per_word, rounding_mode="floor")
            shift = (s.local % per_word) * s.bits
            blob.index_add_(0, pos, s.value.long() << shift)  # disjoint bit ranges: add == or
        start = start + w
    return tile_off.to(torch.int32).contiguous(), _wrap_u32(blob), per_tile


@dataclass
class TilePlan:
    """Device arrays of struct tfem_tile_plan plus bookkeeping."""

    n_tiles: int
    e_off: torch.Tensor
    e_blob: torch.Tensor
    la_off: torch.Tensor
    la_blob: torch.Tensor
    lb_off: torch.Tensor
    lb_blob: torch.Tensor
    max_vert: int
    max_elem: int
    elem_stride: int  # row length of the shared local-matrix table; contribution codes index it directly
    max_e_words: int
    max_la_words: int
    max_lb_words: int
    max_rows: int
    max_out: int
    halo_factor: float  # tile elements / mesh elements (1.0 = every element integrated once)
    index_bytes: int  # bytes of plan data the kernel reads per launch
    consumer_threads: int = 0  # compute threads per CTA (128 / 256 / 384 / 512); 0 = library default
    tile_of_row: torch.Tensor = None  # (n_dof,) tile owning each CSR row
    tile_list: torch.Tensor = None  # optional (n,) int32 subset of tiles to run (see `subset`)
    reserve_ctas: int = 0  # CTA slots left free for kernels on other streams while this plan runs
    n_progress_tiles: int = 0  # the first tiles of `tile_list` report on `progress` when finished
    progress: torch.Tensor = None  # (1,) int32 device counter (see include/tfem_b200.h)

    def c_struct(self) -> "_lib.TilePlan":
        s = _lib.TilePlan()
        s.n_tiles = self.n_tiles if self.tile_list is None else int(self.tile_list.numel())
        s.tile_list = None if self.tile_list is None else self.tile_list.data_ptr()
        s.e_off, s.e_blob = self.e_off.data_ptr(), self.e_blob.data_ptr()
        s.la_off, s.la_blob = self.la_off.data_ptr(), self.la_blob.data_ptr()
        s.lb_off, s.lb_blob = self.lb_off.data_ptr(), self.lb_blob.data_ptr()
        s.max_vert, s.max_elem, s.max_e_words = self.max_vert, self.max_elem, self.max_e_words
        s.max_la_words, s.max_lb_words = self.max_la_words, self.max_lb_words
        s.elem_stride = self.elem_stride
        s.reserve_ctas = self.reserve_ctas
        s.n_progress_tiles = self.n_progress_tiles if self.progress is not None else 0
        s.progress = None if self.progress is None else self.progress.data_ptr()
        s.consumer_threads = int(os.environ.get("TFEM_TILED_CONSUMERS", self.consumer_threads))
        return s

    def subset(self, tile_ids: torch.Tensor, reserve_ctas: int = 0, n_progress_tiles: int = 0, progress: torch.Tensor = None) -> "TilePlan":
        """The same plan restricted to / reordered over some tiles (shares every array); used to
        assemble the tiles that hold multi-GPU interface rows first, with the first
        `n_progress_tiles` of them counted on `progress` so the exchange can start mid-launch."""
        import dataclasses

        return dataclasses.replace(self, tile_list=tile_ids.to(torch.int32).contiguous(), reserve_ctas=reserve_ctas,
                                   n_progress_tiles=n_progress_tiles, progress=progress)

    @property
    def consumer_warps(self) -> int:
        """Warps that report per finished tile on `progress` (library default: 256 consumer threads)."""
        return (int(os.environ.get("TFEM_TILED_CONSUMERS", self.consumer_threads)) or 256) // 32

    def to(self, device) -> "TilePlan":
        moved = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return TilePlan(**moved)

    def sections(self, tile: int) -> dict:
        """Decode one tile's blobs into named integer arrays (tests / debugging)."""
        import numpy as np

        def words_of(off, blob):
            o = off.cpu().numpy().astype("int64")
            return blob[int(o[tile]) : int(o[tile + 1])].cpu().numpy().astype("int64") & 0xFFFFFFFF

        def unpack(words, pos, n, bits):
            n_words = (n * bits + 31) // 32
            w = words[pos : pos + n_words]
            per = 32 // bits
            vals = np.stack([(w >> (k * bits)) & (2**bits - 1) for k in range(per)], axis=1).reshape(-1)[:n]
            return vals, pos + ((n_words + 3) & ~3)

        e = words_of(self.e_off, self.e_blob)
        names = ("n_vert", "n_elem", "n_rows", "n_runs", "n_out", "n_chunks", "base_vertex", "n_heavy_contrib", "n_heavy")
        out = {k: int(v) for k, v in zip(names, e[:HEADER_WORDS])}
        pos = HEADER_WORDS
        out["vert"], pos = unpack(e, pos, out["n_vert"], 32)
        out["elem"], pos = unpack(e, pos, out["n_elem"], 32)
        for off, blob, layout in (
            (self.la_off, self.la_blob, (
                ("run_start", out["n_runs"], 32),
                ("run_meta", out["n_runs"], 32),
                ("pair", out["n_out"], 32),
                ("heavy_seg", out["n_heavy"] + 1, 16),
                ("heavy_contrib", out["n_heavy_contrib"], 16),
                ("heavy_pos", out["n_heavy"], 32),
            )),
            (self.lb_off, self.lb_blob, (
                ("row_id", out["n_rows"], 32),
                ("row_chunk", 8 * out["n_chunks"], 16),
                ("row_diag", out["n_rows"], 32),
            )),
        ):
            lw = words_of(off, blob)
            pos = 0
            for name, n, bits in layout:
                out[name], pos = unpack(lw, pos, n, bits)
        return out


def build_tile_plan(
    geom_conn: torch.Tensor,
    dof_conn: torch.Tensor,
    pattern: CsrPattern,
    row_points: torch.Tensor | None = None,
    rows_per_tile: int = 192,
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
    arange = lambda n: torch.arange(n, device=device)  # noqa: E731

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
            order = arange(n_active)
        tile_of_active = torch.empty(n_active, dtype=torch.int64, device=device)
        tile_of_active[order] = arange(n_active) // rows_per_tile
        n_tiles = max((n_active + rows_per_tile - 1) // rows_per_tile, 1)
    tile_of_row = torch.empty(n_dof, dtype=torch.int64, device=device)
    tile_of_row[active_rows] = tile_of_active
    idle_rows = torch.nonzero(~active, as_tuple=True)[0]
    tile_of_row[idle_rows] = arange(idle_rows.numel()) % n_tiles
    tiles = arange(n_tiles)

    # 2. (tile, element) incidences, tile-major / element ascending
    pair_keys = torch.unique((tile_of_row[dconn] * n_el + arange(n_el)[:, None]).reshape(-1))
    pair_tile = torch.div(pair_keys, n_el, rounding_mode="floor")
    pair_elem = pair_keys - pair_tile * n_el
    elem_ptr = _ptr(pair_tile, n_tiles)
    n_e = elem_ptr[1:] - elem_ptr[:-1]

    # 3. (tile, geometry vertex) incidences and tile-local connectivity
    vkeys_all = pair_tile[:, None] * n_gv + gconn[pair_elem]
    vert_keys = torch.unique(vkeys_all.reshape(-1))
    vert_tile = torch.div(vert_keys, n_gv, rounding_mode="floor")
    tile_vert = vert_keys - vert_tile * n_gv
    vert_ptr = _ptr(vert_tile, n_tiles)
    n_v = vert_ptr[1:] - vert_ptr[:-1]
    local_v = torch.searchsorted(vert_keys, vkeys_all.reshape(-1)).reshape(-1, 3) - vert_ptr[pair_tile][:, None]
    max_vert, max_elem = int(n_v.max().item()), int(n_e.max().item())
    if max_vert > 1024 or max_elem > 4096:
        raise ValueError(f"tile too large (vertices {max_vert} > 1024 or elements {max_elem} > 4096): lower rows_per_tile")
    tile_elem = local_v[:, 0] | (local_v[:, 1] << 10) | (local_v[:, 2] << 20)
    middle = (vert_ptr[:-1] + torch.div(n_v, 2, rounding_mode="floor")).clamp_max(max(tile_vert.numel() - 1, 0))
    base_vertex = torch.where(n_v > 0, tile_vert[middle] if tile_vert.numel() else torch.zeros_like(n_v), torch.zeros_like(n_v))

    # 4. rows of each tile, ascending row id; runs of consecutive rows = contiguous CSR ranges
    row_sorted = torch.argsort(tile_of_row * n_dof + arange(n_dof))
    row_tile = tile_of_row[row_sorted]
    row_ptr = _ptr(row_tile, n_tiles)
    n_r = row_ptr[1:] - row_ptr[:-1]
    row_len = crow[row_sorted + 1] - crow[row_sorted]
    row_out = _excl_cumsum(row_len)  # image offset over all tile-ordered rows
    tile_out0 = row_out[row_ptr[:-1].clamp_max(max(n_dof - 1, 0))]  # image offset of each tile's first row
    n_out = torch.zeros(n_tiles, dtype=torch.int64, device=device).index_add_(0, row_tile, row_len)
    new_run = torch.ones(n_dof, dtype=torch.bool, device=device)
    if n_dof > 1:
        new_run[1:] = (row_tile[1:] != row_tile[:-1]) | (row_sorted[1:] != row_sorted[:-1] + 1)
    run_first = torch.nonzero(new_run, as_tuple=True)[0]
    run_last = torch.cat([run_first[1:], torch.tensor([n_dof], device=device)]) - 1
    whole_start = crow[row_sorted[run_first]]
    whole_len = crow[row_sorted[run_last] + 1] - whole_start
    whole_base = row_out[run_first] - tile_out0[row_tile[run_first]]
    # ... cut into segments of at most 32 entries: one warp, one lane per entry, no inner loop
    SEG = 32
    pieces = torch.div(whole_len + SEG - 1, SEG, rounding_mode="floor")
    piece_run = torch.repeat_interleave(arange(whole_len.numel()), pieces)
    piece_k = arange(piece_run.numel()) - _excl_cumsum(pieces)[piece_run]
    run_tile = row_tile[run_first][piece_run]
    run_ptr = _ptr(run_tile, n_tiles)
    n_u = run_ptr[1:] - run_ptr[:-1]
    run_start = whole_start[piece_run] + SEG * piece_k
    run_len = (whole_len[piece_run] - SEG * piece_k).clamp_max(SEG)
    run_base = whole_base[piece_run] + SEG * piece_k
    if int(n_out.max().item()) > 65535 or int(n_u.max().item()) > 65535:
        raise ValueError("tile image too large (entries or segments > 65535): lower rows_per_tile")

    # 5. every CSR entry of a tile with its contributions (element ascending = reference order)
    nnz = pattern.nnz
    ent_row = torch.repeat_interleave(arange(n_dof), row_len)  # tile-ordered row index of each image slot
    ent_global = crow[row_sorted[ent_row]] + (arange(nnz) - row_out[ent_row])
    ent_tile = row_tile[ent_row]
    ent_local = arange(nnz) - tile_out0[ent_tile]
    seg = pattern.seg.long()
    ent_cnt = seg[ent_global + 1] - seg[ent_global]
    ent_coff = _excl_cumsum(ent_cnt)
    total_c = int(ent_cnt.sum().item())
    c_ent = torch.repeat_interleave(arange(nnz), ent_cnt)
    c_within = arange(total_c) - ent_coff[c_ent]
    coo = pattern.perm.long()[seg[ent_global[c_ent]] + c_within]
    c_e = torch.div(coo, 9, rounding_mode="floor")
    c_i = torch.div(coo - 9 * c_e, 3, rounding_mode="floor")
    c_j = coo - 9 * c_e - 3 * c_i
    slot = torch.where(c_i == c_j, c_i, 3 + (c_i + c_j == 3).long() + 2 * (c_i + c_j == 2).long())  # K00 K11 K22 K01 K12 K20
    c_tile = ent_tile[c_ent]
    # sloc is [9][elem_stride]; codes `slot*elem_stride + element` index it directly.  The last
    # column (element elem_stride-1) is never written by the integration phase and holds zeros:
    # `zero_code` pads every fixed-length contribution list.
    elem_stride = (max_elem + 1 + 31) & ~31
    zero_code = elem_stride - 1
    if 9 * elem_stride > 65535:
        raise ValueError("tile has too many elements for 16-bit local-matrix indices: lower rows_per_tile")
    c_code = (torch.searchsorted(pair_keys, c_tile * n_el + c_e) - elem_ptr[c_tile]) + slot * elem_stride

    # 6. load-vector contributions per owned row
    lseg = pattern.lin_seg.long()
    lrow_cnt = lseg[row_sorted + 1] - lseg[row_sorted]
    lrow_off = _excl_cumsum(lrow_cnt)
    total_lc = int(lrow_cnt.sum().item())
    lc_row = torch.repeat_interleave(arange(n_dof), lrow_cnt)
    lc_within = arange(total_lc) - lrow_off[lc_row]
    lc_flat = pattern.lin_perm.long()[lseg[row_sorted[lc_row]] + lc_within]
    lc_e = torch.div(lc_flat, 3, rounding_mode="floor")
    lc_k = lc_flat - 3 * lc_e
    lc_tile = row_tile[lc_row]
    lc_code = (torch.searchsorted(pair_keys, lc_tile * n_el + lc_e) - elem_ptr[lc_tile]) + lc_k * elem_stride

    # 6b. who sums which entry.  One thread per entry handles entries with <= 2 contributions
    #     (every off-diagonal entry of a manifold mesh) from ONE packed word, without a loop; the
    #     diagonal of a row is summed by the row's thread together with its load entry (same element
    #     list); anything else with > 2 contributions (non-manifold edges, degenerate elements)
    #     goes to a short "heavy" list handled by a generic loop.
    ent_is_diag = pattern.col.long()[ent_global] == row_sorted[ent_row]
    off_slot = torch.zeros(nnz, dtype=torch.int64, device=device).index_add_(0, c_ent, (slot >= 3).long())
    diag_fast = ent_is_diag & (ent_cnt > 2) & (off_slot == 0) & (ent_cnt == lrow_cnt[ent_row])
    row_diag = torch.full((n_dof,), 0xFFFFFFFF, dtype=torch.int64, device=device)  # CSR position or none
    row_diag[ent_row[diag_fast]] = ent_global[diag_fast]
    heavy = (ent_cnt > 2) & ~diag_fast
    heavy_tile = ent_tile[heavy]
    n_h = torch.bincount(heavy_tile, minlength=n_tiles)
    heavy_local = arange(heavy_tile.numel()) - _excl_cumsum(n_h)[heavy_tile]
    # contributions of the heavy entries, with per-tile offsets (n_h + 1 values per tile)
    heavy_ids = torch.nonzero(heavy, as_tuple=True)[0]
    heavy_cnt = ent_cnt[heavy_ids]
    n_hc = torch.zeros(n_tiles, dtype=torch.int64, device=device).index_add_(0, heavy_tile, heavy_cnt)
    heavy_coff = _excl_cumsum(heavy_cnt)
    total_hc = int(heavy_cnt.sum().item())
    if int(n_hc.max().item()) > 65535:
        raise ValueError("tile has more than 65535 heavy contributions: lower rows_per_tile")
    hc_item = torch.repeat_interleave(arange(heavy_ids.numel()), heavy_cnt)
    hc_code = c_code[ent_coff[heavy_ids[hc_item]] + (arange(total_hc) - heavy_coff[hc_item])]
    hc_tile = heavy_tile[hc_item]
    tile_hc0 = _excl_cumsum(n_hc)
    hseg_tile = torch.repeat_interleave(tiles, n_h + 1)
    hseg_local = arange(hseg_tile.numel()) - (_excl_cumsum(n_h) + tiles)[hseg_tile]
    heavy_coff_ext = torch.cat([heavy_coff, torch.tensor([total_hc], device=device)])
    hseg_value = heavy_coff_ext[_excl_cumsum(n_h)[hseg_tile] + hseg_local] - tile_hc0[hseg_tile]

    # 6c. the packed pair of every entry: lo 16 bits = first contribution, hi 16 = second (zero_code
    #     when absent); 0xFFFFFFFF = "not mine" (diagonal done by the row thread, or heavy)
    first = torch.full((nnz,), zero_code, dtype=torch.int64, device=device)
    second = torch.full((nnz,), zero_code, dtype=torch.int64, device=device)
    first[c_ent[c_within == 0]] = c_code[c_within == 0]
    second[c_ent[c_within == 1]] = c_code[c_within == 1]
    pair = first | (second << 16)
    pair[ent_cnt > 2] = 0xFFFFFFFF

    # 6d. per-row chunks of 8 x u16: 7 contribution codes (k*elem_stride + element; zero_code pads)
    #     and the tile-local index of the row's next chunk (0 = none).  Chunk j < n_rows is the first
    #     chunk of row j; rows with more than 7 elements continue in chunks appended after n_rows.
    PER = 7
    row_chunks = torch.div(lrow_cnt + PER - 1, PER, rounding_mode="floor").clamp_min(1)
    row_extra = row_chunks - 1
    extra_off = _excl_cumsum(row_extra)
    extra_in_tile = extra_off - extra_off[row_ptr[:-1].clamp_max(max(n_dof - 1, 0))][row_tile]
    n_x = torch.zeros(n_tiles, dtype=torch.int64, device=device).index_add_(0, row_tile, row_extra)
    n_ch = n_r + n_x
    if int(n_ch.max().item()) > 65535:
        raise ValueError("tile has more than 65535 row chunks: lower rows_per_tile")
    chunk0 = _excl_cumsum(n_ch)
    total_ch = int(n_ch.sum().item())
    chunks = torch.full((total_ch, 8), zero_code, dtype=torch.int64, device=device)
    chunks[:, 7] = 0
    row_local = arange(n_dof) - row_ptr[row_tile]
    lc_q = torch.div(lc_within, PER, rounding_mode="floor")
    lc_chunk_local = torch.where(lc_q == 0, row_local[lc_row], n_r[lc_tile] + extra_in_tile[lc_row] + lc_q - 1)
    chunks[chunk0[lc_tile] + lc_chunk_local, lc_within - PER * lc_q] = lc_code
    # links: chunk q of a row -> chunk q+1
    link_row = torch.repeat_interleave(arange(n_dof), row_extra)
    link_q = arange(link_row.numel()) - extra_off[link_row]  # 0-based: link from chunk q to q+1
    link_tile = row_tile[link_row]
    link_to = n_r[link_tile] + extra_in_tile[link_row] + link_q
    link_from = torch.where(link_q == 0, row_local[link_row], link_to - 1)
    chunks[chunk0[link_tile] + link_from, 7] = link_to
    chunk_tile = torch.repeat_interleave(tiles, n_ch * 8)
    chunk_local = arange(total_ch * 8) - (chunk0 * 8)[chunk_tile]

    # 7. pack
    header = torch.stack([n_v, n_e, n_r, n_u, n_out, n_ch, base_vertex, n_hc, n_h, n_h * 0, n_h * 0, n_h * 0], dim=1).reshape(-1)
    hdr_tile = torch.repeat_interleave(tiles, HEADER_WORDS)
    hdr_local = arange(hdr_tile.numel()) % HEADER_WORDS
    e_sections = [
        _Section("header", 32, torch.full_like(n_v, HEADER_WORDS), hdr_tile, hdr_local, header),
        _Section("vert", 32, n_v, vert_tile, arange(vert_tile.numel()) - vert_ptr[vert_tile], tile_vert),
        _Section("elem", 32, n_e, pair_tile, arange(pair_tile.numel()) - elem_ptr[pair_tile], tile_elem),
    ]
    run_local = arange(run_tile.numel()) - run_ptr[run_tile]
    la_sections = [
        _Section("run_start", 32, n_u, run_tile, run_local, run_start),
        _Section("run_meta", 32, n_u, run_tile, run_local, run_base | (run_len << 16)),
        _Section("pair", 32, n_out, ent_tile, ent_local, pair),
        _Section("heavy_seg", 16, n_h + 1, hseg_tile, hseg_local, hseg_value),
        _Section("heavy_contrib", 16, n_hc, hc_tile, arange(total_hc) - tile_hc0[hc_tile], hc_code),
        _Section("heavy_pos", 32, n_h, heavy_tile, heavy_local, ent_global[heavy]),
    ]
    lb_sections = [
        _Section("row_id", 32, n_r, row_tile, row_local, row_sorted),
        _Section("row_chunk", 16, n_ch * 8, chunk_tile, chunk_local, chunks.reshape(-1)),
        _Section("row_diag", 32, n_r, row_tile, row_local, row_diag),
    ]
    e_off, e_blob, e_words = _pack(e_sections, n_tiles, device)
    la_off, la_blob, la_words = _pack(la_sections, n_tiles, device)
    lb_off, lb_blob, lb_words = _pack(lb_sections, n_tiles, device)
    return TilePlan(
        n_tiles=n_tiles,
        e_off=e_off,
        e_blob=e_blob,
        la_off=la_off,
        la_blob=la_blob,
        lb_off=lb_off,
        lb_blob=lb_blob,
        max_vert=max_vert,
        max_elem=max_elem,
        elem_stride=elem_stride,
        max_e_words=int(e_words.max().item()),
        max_la_words=int(la_words.max().item()),
        max_lb_words=int(lb_words.max().item()),
        max_rows=int(n_r.max().item()),
        max_out=int(n_out.max().item()),
        halo_factor=float(pair_keys.shape[0]) / max(n_el, 1),
        index_bytes=4 * (e_blob.numel() + la_blob.numel() + lb_blob.numel() + e_off.numel() + la_off.numel() + lb_off.numel()),
        tile_of_row=tile_of_row,
    )
