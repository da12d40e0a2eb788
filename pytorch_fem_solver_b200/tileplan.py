"""Row-tile plan of the fused assembly kernel (`tfem_tri_p1_assemble_csr`), format v2.

Integer, one-time set-up in torch (any device).  CSR rows are clustered into tiles; a tile's elements
are all elements touching its rows.  Everything the kernel needs per tile is split in two:

  template  the tile-LOCAL index structure, which depends on the mesh topology around the tile only:
            tile-local connectivity (TB blob) and, for every CSR entry / row of the tile, which local
            (element, slot) contributions it sums (TC blob).  Congruent tiles -- every interior tile of
            a lattice-numbered mesh -- share ONE template; the kernel keeps it resident in shared
            memory and refetches only when consecutive tiles differ.
  instance  what is specific to a tile: the rows of `coords` of its vertices, the global CSR offset
            of each of its entry segments and its global row ids (one small blob per tile).

Layouts are documented in include/tfem_b200.h.  Every blob is fetched by one TMA bulk copy, so sections
are padded to whole 16 B units.
"""

from __future__ import annotations

import dataclasses
import os
from dataclasses import dataclass

import torch

from . import _lib
from .csr import CsrPattern

INST_HEADER_WORDS = 4
TB_HEADER_WORDS = 8
SEG = 32  # entries per segment: one warp pass, one lane per entry
PER_CHUNK = 7  # contribution codes per 16 B row chunk (+ link)
# The CTA's local table (byte offsets for fp64; the fp32 kernels halve them).  With R = max_elem + 2 rows
# (row 0 = zeros, rows 1..n = tile elements, last row = scratch):
#   [0, 48 R)                 "dl" rows of 48 B: {diagonal term, load term} of local vertex 0, 1, 2 -- the row
#                             phase fetches a diagonal and its load term with one 16 B load
#   od_base[k] + 8 row        three arrays (k = 0, 1, 2) of the off-diagonal terms K01, K12, K20
# Consecutive rows of one array are consecutive in memory, so lanes reading the same kind of term of
# consecutive elements never collide on a bank; the arrays' base offsets (multiples of 8 B) and the order of
# the elements inside a tile are chosen per plan by `_choose_layout` to minimise bank conflicts.
DL_ROW_BYTES = 48
OD_PAD_BYTES = 128  # room behind each off-diagonal array for its base shift
MAX_VERT = 1024  # 10-bit tile-local vertex ids
MAX_ELEM = 880  # 16-bit byte codes: 48 R + 3 (8 R + 128) < 65536


def table_layout(max_elem: int, shift=(0, 0, 0)):
    """(rows R, od_base[3], table bytes) of the local table for tiles of at most `max_elem` elements."""
    rows = max_elem + 2
    start = DL_ROW_BYTES * rows
    od_base = [start + k * (8 * rows + OD_PAD_BYTES) + 8 * int(shift[k]) for k in range(3)]
    return rows, od_base, (start + 3 * (8 * rows + OD_PAD_BYTES) + 15) & ~15


def slot_code(row: torch.Tensor, slot: torch.Tensor, od_base) -> torch.Tensor:
    """Byte offset of local value `slot` (K00 K11 K22 K01 K12 K20 b0 b1 b2) of table row `row`."""
    base = torch.tensor(list(od_base), dtype=torch.int64, device=row.device)
    k = torch.where(slot >= 6, slot - 6, slot)
    dl = DL_ROW_BYTES * row + 16 * k + 8 * (slot >= 6).long()
    od = base[(slot - 3).clamp(0, 2)] + 8 * row
    return torch.where((slot >= 3) & (slot < 6), od, dl)


def _conflict_wavefronts(row, kind, active, lanes_per_group, group_bank):
    """Wavefronts of warp-wide shared-memory loads.  row/kind/active: (n_loads, 32); a lane's address is
    identified by (kind, row); `group_bank(kind, row)` gives its bank group (int array); lanes reading the
    same address share a wavefront.  Loads are served per group of `lanes_per_group` lanes."""
    import numpy as np

    n = row.shape[0]
    g = 32 // lanes_per_group
    key = (kind.astype(np.int64) * (1 << 20) + row).reshape(n * g, lanes_per_group)
    act = active.reshape(n * g, lanes_per_group)
    same = key[:, :, None] == key[:, None, :]
    earlier = np.tril(np.ones((lanes_per_group, lanes_per_group), dtype=bool), -1)
    first = act & ~(same & earlier[None] & act[:, None, :]).any(axis=2)
    bank = group_bank(kind, row).reshape(n * g, lanes_per_group)
    n_banks = int(bank.max()) + 1 if bank.size else 1
    counts = np.zeros((n * g, n_banks), dtype=np.int64)
    np.add.at(counts, (np.repeat(np.arange(n * g), lanes_per_group).reshape(n * g, lanes_per_group)[first], bank[first]), 1)
    return int(np.maximum(counts.max(axis=1), act.any(axis=1)).sum())


def _choose_layout(entry_rows, entry_slots, entry_active, row_rows, row_k, row_active, n_elem, max_elem, device=None):
    """Pick the element order inside a tile and the base shifts of the off-diagonal arrays that minimise the
    shared-memory wavefronts of the reduction phase on a representative tile.

    entry_*: (2, n_segs, 32) table row (natural element order, 1-based), slot (3..5) and validity of the first /
    second contribution of every entry lane; row_*: (n_rows, 7) of the row phase.  Returns (order, shift, stats):
    order 1 = ascending element id, 2 = even-position elements first, then odd-position ones."""
    import numpy as np

    best = None
    stats = {}
    pad = (-row_rows.shape[0]) % 32
    for order in (1, 2):
        def remap(r):
            loc = r - 1
            new = np.where(loc >= 0, (loc % order) * ((n_elem + order - 1) // order) + loc // order, -1)
            return np.where(r > 0, new + 1, 0)

        rr = np.concatenate([remap(row_rows), np.zeros((pad, 7), dtype=np.int64)]).reshape(-1, 32, 7)
        rk = np.concatenate([row_k, np.zeros((pad, 7), dtype=np.int64)]).reshape(-1, 32, 7)
        ra = np.concatenate([row_active, np.zeros((pad, 7), dtype=bool)]).reshape(-1, 32, 7)
        rows_wf = sum(_conflict_wavefronts(rr[:, :, j], rk[:, :, j], ra[:, :, j], 8, lambda k, r: (3 * r + k) % 8) for j in range(7))
        er = remap(entry_rows)
        wf_all = _entry_wavefronts_all_shifts(er, entry_slots - 3, entry_active, max_elem, device)  # (16, 16): shifts of K12, K20
        flat = int(np.argmin(wf_all))  # first minimum in (s1, s2) order, as a nested loop with a strict "<" would pick
        s1, s2 = divmod(flat, 16)
        wf = int(wf_all[s1, s2])
        if best is None or wf + rows_wf < best[0]:
            best = (wf + rows_wf, order, (0, s1, s2), wf, rows_wf)
        stats[order] = rows_wf
    return best[1], best[2], {"entries": best[3], "rows": best[4]}


def _entry_wavefronts_all_shifts(rows, kinds, active, max_elem, device=None):
    """Wavefronts of the entry phase's two table loads (`_conflict_wavefronts` with 16 lanes per group and bank
    `(od_base[kind] / 8 + row) % 16`) for all 16 x 16 base shifts of the K12 / K20 arrays at once: which lanes share an
    address does not depend on the shifts, only the banks do.  Evaluated with torch on `device` (the 10 M-element
    one-hot costs 60 ms in numpy on the host, 1 ms on the GPU the plan is being built for)."""
    import numpy as np

    _, od_base, _ = table_layout(max_elem, (0, 0, 0))
    dev = torch.device(device) if device is not None else torch.device("cpu")
    base8 = torch.tensor([b // 8 for b in od_base], dtype=torch.int64, device=dev)
    shifts = torch.arange(16, device=dev)
    total = torch.zeros((16, 16), dtype=torch.int64, device=dev)
    earlier = torch.tril(torch.ones((16, 16), dtype=torch.bool, device=dev), -1)
    for c in range(rows.shape[0]):
        row = torch.as_tensor(np.ascontiguousarray(rows[c])).to(dev).reshape(-1, 16)
        kind = torch.as_tensor(np.ascontiguousarray(kinds[c])).to(dev).reshape(-1, 16)
        act = torch.as_tensor(np.ascontiguousarray(active[c])).to(dev).reshape(-1, 16)
        key = kind * (1 << 20) + row
        same = key[:, :, None] == key[:, None, :]
        first = act & ~(same & earlier[None] & act[:, None, :]).any(dim=2)  # (G, 16): first lane of each distinct address
        bank0 = base8[kind] + row  # (G, 16)
        shift = torch.where(kind == 1, shifts[:, None, None, None], 0) + torch.where(kind == 2, shifts[None, :, None, None], 0)
        bank = (bank0[None, None] + shift) % 16  # (16, 16, G, 16)
        onehot = (bank[..., None] == shifts) & first[None, None, :, :, None]  # (16, 16, G, lanes, banks)
        worst = onehot.sum(dim=3).amax(dim=3)  # (16, 16, G)
        total += torch.maximum(worst, act.any(dim=1)[None, None].to(worst.dtype)).sum(dim=2)
    return total.cpu().numpy()


MAX_SEGS = 2047  # 16-bit entry codes seg * 32 + lane, 0xFFFF = none


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


def detect_lattice(pattern: CsrPattern):
    """Row stride W of a lattice-numbered mesh (DOF id = j*W + i), read off the CSR pattern alone.

    The commonest column-offset signature of the 7-entry rows of a structured triangulation is
    (-W-1, -W, -1, 0, 1, W, W+1) or (-W, -W+1, -1, 0, 1, W-1, W).  Returns W, or 0 when fewer than
    half of the rows follow one such stencil or the numbering does not wrap consistently.  Tiles of
    such a mesh are cut in INDEX space, so they do not depend on vertex positions (jittered vertices
    leave every interior tile congruent, which is what lets the templates be shared)."""
    n = pattern.n_dof
    if n < 16 or pattern.nnz < 7 * 8:
        return 0
    crow = pattern.crow.long()
    length = crow[1:] - crow[:-1]
    rows7 = torch.nonzero(length == 7, as_tuple=True)[0]
    if rows7.numel() * 2 < n:
        return 0
    sample = rows7[:: max(rows7.numel() // 4096, 1)]
    offs = pattern.col.long()[crow[sample][:, None] + torch.arange(7, device=crow.device)[None, :]] - sample[:, None]
    signatures, counts = torch.unique(offs, dim=0, return_counts=True)
    best = signatures[int(torch.argmax(counts))].tolist()
    if int(counts.max()) * 2 < sample.numel():
        return 0
    if best[2:5] != [-1, 0, 1] or best[0] != -best[6] or best[1] != -best[5] or best[6] - best[5] != 1:
        return 0
    col = pattern.col.long()
    row_of = pattern.row_indices()
    has_next = torch.zeros(n, dtype=torch.bool, device=crow.device)
    has_next[row_of[col == row_of + 1]] = True
    for width in (best[5], best[6]):
        if width < 2 or n % width != 0:
            continue
        last_in_row = (torch.arange(n, device=crow.device) % width) == width - 1
        populated = length > 0  # rows without entries (ghost columns of a multi-GPU owner) say nothing
        if not bool((has_next & last_in_row).any()) and bool(has_next[~last_in_row & populated].all()):
            return int(width)
    return 0


def lattice_shape(rows_per_tile: int):
    """(bx, by): tile extent in lattice columns / lattice rows for about `rows_per_tile` rows."""
    by = max(int(round((rows_per_tile / 1.3) ** 0.5)), 1)
    bx = max(rows_per_tile // by, 1)
    return bx, by


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
    bits: int  # 32 or 16
    count: torch.Tensor  # (n_tiles,) items per tile
    tile: torch.Tensor  # (n_items,) tile of each item
    local: torch.Tensor  # (n_items,) index of the item inside its tile
    value: torch.Tensor  # (n_items,) non-negative, < 2**bits
    fill: int = 0  # value of the items nobody sets (e.g. padding lanes of a segment)


def _pack(sections: list, n_tiles: int, device):
    """Lay the sections of every tile out back to back.

    Returns (tile_off int64 [n_tiles+1], words int64 in [0, 2**32), words per tile)."""
    words = [_pad4((s.count * s.bits + 31) // 32) for s in sections]
    per_tile = sum(words)
    tile_off = torch.zeros(n_tiles + 1, dtype=torch.int64, device=device)
    tile_off[1:] = torch.cumsum(per_tile, 0)
    total = int(tile_off[-1].item())
    if total >= 2**31:
        raise ValueError("tile plan too large for 32-bit word offsets")
    blob = torch.zeros(total, dtype=torch.int64, device=device)
    start = tile_off[:-1].clone()
    tiles = torch.arange(n_tiles, device=device)
    for s, w in zip(sections, words):
        per_word = 32 // s.bits
        if s.fill:
            # pre-fill the section's items (not the alignment padding behind them) with the pattern
            n_items = s.count
            item_tile = torch.repeat_interleave(tiles, n_items)
            item_local = torch.arange(item_tile.numel(), device=device) - _excl_cumsum(n_items)[item_tile]
            pos = start[item_tile] + torch.div(item_local, per_word, rounding_mode="floor")
            blob.index_add_(0, pos, torch.full_like(pos, s.fill) << ((item_local % per_word) * s.bits))
        if s.value.numel():
            if int(s.value.max().item()) >= 2**s.bits or int(s.value.min().item()) < 0:
                raise ValueError(f"tile plan section {s.name}: value does not fit {s.bits} bits")
            pos = start[s.tile] + torch.div(s.local, per_word, rounding_mode="floor")
            shift = (s.local % per_word) * s.bits
            blob.index_add_(0, pos, (s.value.long() - s.fill) << shift)  # disjoint bit ranges: add == or
        start = start + w
    return tile_off, blob, per_tile


def _hash_tiles(tile_off: torch.Tensor, blob: torch.Tensor, n_tiles: int, seed: int) -> torch.Tensor:
    """Two 61-bit positional hashes of every tile's words (no integer overflow: 16-bit halves times
    31-bit multipliers, at most 2**14 words per tile)."""
    device = blob.device
    per_tile = tile_off[1:] - tile_off[:-1]
    width = int(per_tile.max().item()) if n_tiles else 0
    if width >= 2**14:
        raise ValueError("tile blob too long to hash")
    gen = torch.Generator(device="cpu").manual_seed(seed)
    mult = torch.randint(1, 2**31 - 1, (4, max(width, 1)), generator=gen, dtype=torch.int64, device="cpu").to(device)
    word_tile = torch.repeat_interleave(torch.arange(n_tiles, device=device), per_tile)
    local = torch.arange(blob.numel(), device=device) - tile_off[word_tile]
    lo, hi = (blob & 0xFFFF) + 1, (blob >> 16) + 1
    out = torch.zeros((n_tiles, 2), dtype=torch.int64, device=device)
    out[:, 0].index_add_(0, word_tile, lo * mult[0, local] + hi * mult[1, local])
    out[:, 1].index_add_(0, word_tile, lo * mult[2, local] + hi * mult[3, local])
    out[:, 0] += per_tile * 1000003
    return out


def _gather_ranges(blob: torch.Tensor, start: torch.Tensor, length: torch.Tensor) -> torch.Tensor:
    """Concatenation of blob[start[k] : start[k] + length[k]] over k."""
    which = torch.repeat_interleave(torch.arange(start.numel(), device=blob.device), length)
    local = torch.arange(which.numel(), device=blob.device) - _excl_cumsum(length)[which]
    return blob[start[which] + local]


@dataclass
class TilePlan:
    """Device arrays of struct tfem_tile_plan plus bookkeeping."""

    n_tiles: int
    n_templates: int
    tile_desc: torch.Tensor  # (n_tiles, 4) int32: instance word offset, instance words, template, coords row of the base vertex
    inst_blob: torch.Tensor  # int32
    tpl_desc: torch.Tensor  # (n_templates, 4) int32: TB offset, TB words, TC offset, TC words
    tpl_blob: torch.Tensor  # int32
    default_order: torch.Tensor  # (n_tiles,) int32: all tiles, congruent ones adjacent
    max_vert: int
    max_elem: int
    max_rows: int
    max_segs: int
    max_inst_words: int
    max_tb_words: int
    max_tc_words: int
    halo_factor: float  # tile elements / mesh elements (1.0 = every element integrated once)
    index_bytes: int  # bytes of plan data one launch reads (instances + each template once + descriptors)
    lattice: tuple = None  # (W, bx, by) when the tiles are index-space blocks of a lattice-numbered mesh
    table_bytes: int = 0  # bytes (fp64) of the CTA's local table
    od_base: tuple = (0, 0, 0)  # byte offsets of the three off-diagonal arrays inside the table
    elem_order: int = 1  # 1 = tile elements in ascending id, 2 = even positions first, then odd
    layout_stats: dict = None  # modelled shared-memory wavefronts of the representative tile's reduction phase
    has_elem_ids: bool = False  # instances carry the global id of every tile element (sampled sources, fractures)
    consumer_threads: int = 0  # kernel selector: 0 = library default, 384 = classic kernel, 100 * NB + NC = role-specialised
    tile_of_row: torch.Tensor = None  # (n_dof,) tile owning each CSR row
    tile_list: torch.Tensor = None  # optional (n,) int32 subset / order of tiles to run (see `subset`)
    reserve_ctas: int = 0  # CTA slots left free for kernels on other streams while this plan runs
    n_progress_tiles: int = 0  # the first tiles of `tile_list` report on `progress` when finished
    progress: torch.Tensor = None  # (1,) int32 device counter (see include/tfem_b200.h)

    def c_struct(self) -> "_lib.TilePlan":
        s = _lib.TilePlan()
        order = self.default_order if self.tile_list is None else self.tile_list
        s.n_tiles = int(order.numel())
        s.tile_list = order.data_ptr()
        s.tile_desc, s.inst_blob = self.tile_desc.data_ptr(), self.inst_blob.data_ptr()
        s.tpl_desc, s.tpl_blob = self.tpl_desc.data_ptr(), self.tpl_blob.data_ptr()
        s.max_vert, s.max_elem = self.max_vert, self.max_elem
        s.max_inst_words, s.max_tb_words, s.max_tc_words = self.max_inst_words, self.max_tb_words, self.max_tc_words
        s.has_elem_ids = 1 if self.has_elem_ids else 0
        s.table_bytes = self.table_bytes
        for k in range(3):
            s.od_base[k] = self.od_base[k]
        s.reserve_ctas = self.reserve_ctas
        s.n_progress_tiles = self.n_progress_tiles if self.progress is not None else 0
        s.progress = None if self.progress is None else self.progress.data_ptr()
        s.consumer_threads = int(os.environ.get("TFEM_TILED_CONSUMERS", self.consumer_threads))
        self._keepalive = order  # the struct only holds raw pointers
        return s

    def op_args(self):
        """(tile_list, tile_desc, inst_blob, tpl_desc, tpl_blob, meta) for the registered op `assemble_csr_tiled`."""
        order = self.default_order if self.tile_list is None else self.tile_list
        meta = [self.max_vert, self.max_elem, self.max_inst_words, self.max_tb_words, self.max_tc_words, int(self.has_elem_ids),
                self.table_bytes, *self.od_base, int(os.environ.get("TFEM_TILED_CONSUMERS", self.consumer_threads)), self.reserve_ctas,
                self.n_progress_tiles if self.progress is not None else 0]
        return order, self.tile_desc, self.inst_blob, self.tpl_desc, self.tpl_blob, meta

    def subset(self, tile_ids: torch.Tensor, reserve_ctas: int = 0, n_progress_tiles: int = 0, progress: torch.Tensor = None) -> "TilePlan":
        """The same plan restricted to / reordered over some tiles (shares every array); used to
        assemble the tiles that hold multi-GPU interface rows first, with the first
        `n_progress_tiles` of them counted on `progress` so the exchange can start mid-launch."""
        return dataclasses.replace(self, tile_list=tile_ids.to(torch.int32).contiguous(), reserve_ctas=reserve_ctas,
                                   n_progress_tiles=n_progress_tiles, progress=progress)

    @property
    def consumer_warps(self) -> int:
        """Warps that report per finished tile on `progress`."""
        consumers = int(os.environ.get("TFEM_TILED_CONSUMERS", self.consumer_threads)) or default_consumers(self.max_elem)
        return consumers % 100 if consumers >= 1000 else consumers // 32  # role-specialised kernel: its reduction warps report

    def to(self, device) -> "TilePlan":
        moved = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items() if not k.startswith("_")}
        return TilePlan(**moved)

    def sections(self, tile: int) -> dict:
        """Decode one tile's instance and template into named integer arrays (tests / debugging)."""
        import numpy as np

        def unpack(words, pos, n, bits):
            n_words = (n * bits + 31) // 32
            w = words[pos : pos + n_words]
            per = 32 // bits
            vals = np.stack([(w >> (k * bits)) & (2**bits - 1) for k in range(per)], axis=1).reshape(-1)[:n]
            return vals, pos + ((n_words + 3) & ~3)

        inst_off, inst_words, tpl, _ = self.tile_desc[tile].tolist()
        inst = self.inst_blob[inst_off : inst_off + inst_words].cpu().numpy().astype("int64") & 0xFFFFFFFF
        tb_off, tb_words, tc_off, tc_words = self.tpl_desc[tpl].tolist()
        tb = self.tpl_blob[tb_off : tb_off + tb_words].cpu().numpy().astype("int64") & 0xFFFFFFFF
        tc = self.tpl_blob[tc_off : tc_off + tc_words].cpu().numpy().astype("int64") & 0xFFFFFFFF
        out = {"template": tpl, "inst_offset": inst_off, "tb_offset": tb_off, "tc_offset": tc_off}
        names = ("n_vert", "n_elem", "n_rows", "n_segs", "n_chunks", "n_heavy", "n_heavy_contrib")
        out.update({k: int(v) for k, v in zip(names, tb[:TB_HEADER_WORDS])})
        out["inst_header"] = inst[:INST_HEADER_WORDS].tolist()
        out["base_vertex"] = int(inst[3])
        pos = INST_HEADER_WORDS
        out["vert"], pos = unpack(inst, pos, out["n_vert"], 32)
        out["seg_start"], pos = unpack(inst, pos, out["n_segs"], 32)
        out["row_id"], pos = unpack(inst, pos, out["n_rows"], 32)
        if self.has_elem_ids:
            out["elem_id"], pos = unpack(inst, pos, out["n_elem"], 32)
        assert pos == inst_words
        out["elem"], pos = unpack(tb, TB_HEADER_WORDS, out["n_elem"], 32)
        assert pos == tb_words
        pos = 0
        out["pair"], pos = unpack(tc, pos, SEG * out["n_segs"], 32)
        out["row_chunk"], pos = unpack(tc, pos, 8 * out["n_chunks"], 16)
        out["row_diag"], pos = unpack(tc, pos, out["n_rows"], 16)
        out["heavy_seg"], pos = unpack(tc, pos, out["n_heavy"] + 1, 16)
        out["heavy_contrib"], pos = unpack(tc, pos, out["n_heavy_contrib"], 16)
        out["heavy_pos"], pos = unpack(tc, pos, out["n_heavy"], 16)
        assert pos == tc_words
        return out


def _layout_for_tile(t, ent_tile, ent_code, ent_cnt, c_ent, c_within, c_loc, slot, row_tile, row_ptr, lrow_cnt, lc_row, lc_within, lc_loc, lc_k,
                     n_segs, n_rows, n_elem, max_elem):
    """Gather tile t's reduction-phase accesses as numpy arrays and run `_choose_layout`."""
    import numpy as np

    entry_rows = np.zeros((2, n_segs * SEG), dtype=np.int64)
    entry_slots = np.full((2, n_segs * SEG), 3, dtype=np.int64)
    entry_active = np.zeros((2, n_segs * SEG), dtype=bool)
    light = (ent_tile[c_ent] == t) & (ent_cnt[c_ent] <= 2) & (slot >= 3) & (slot < 6)
    for which in (0, 1):
        sel = light & (c_within == which)
        code = ent_code[c_ent[sel]].cpu().numpy()
        entry_rows[which, code] = c_loc[sel].cpu().numpy() + 1
        entry_slots[which, code] = slot[sel].cpu().numpy()
        entry_active[which, code] = True
    # a lane whose entry has one contribution still reads the zero row for the other one
    entry_active[1] |= entry_active[0]
    row_rows = np.zeros((n_rows, PER_CHUNK), dtype=np.int64)
    row_k = np.zeros((n_rows, PER_CHUNK), dtype=np.int64)
    sel = (row_tile[lc_row] == t) & (lc_within < PER_CHUNK)
    local = (lc_row[sel] - row_ptr[t]).cpu().numpy()
    within = lc_within[sel].cpu().numpy()
    row_rows[local, within] = lc_loc[sel].cpu().numpy() + 1
    row_k[local, within] = lc_k[sel].cpu().numpy()
    row_active = np.ones((n_rows, PER_CHUNK), dtype=bool)
    return _choose_layout(entry_rows.reshape(2, n_segs, SEG), entry_slots.reshape(2, n_segs, SEG), entry_active.reshape(2, n_segs, SEG),
                          row_rows, row_k, row_active, n_elem, max_elem, device=ent_tile.device)


def default_consumers(max_elem: int) -> int:
    """Kernel the library picks for a plan (mirrors assemble_tiled.cu): 100 * NB + NC = the role-specialised kernel with
    NB integration and NC reduction warps; 384 = the classic kernel with that many consumer threads."""
    return 1212


class TileTooLarge(ValueError):
    """A tile exceeds what the kernel's 10-bit vertex ids / 16-bit table codes can address."""


def build_tile_plan(geom_conn, dof_conn, pattern, row_points=None, rows_per_tile: int = 336, ordering: str = "auto",
                    tile_shape: tuple | None = None, elem_ids: bool = False) -> "TilePlan":
    """`_build_tile_plan` with the tile size lowered until every tile fits the kernel's index widths
    (an unstructured mesh has more elements per row than a lattice)."""
    while True:
        try:
            return _build_tile_plan(geom_conn, dof_conn, pattern, row_points, rows_per_tile, ordering, tile_shape, elem_ids)
        except TileTooLarge:
            if rows_per_tile <= 8:
                raise
            rows_per_tile, tile_shape = max(rows_per_tile * 3 // 4, 8), None


def _build_tile_plan(
    geom_conn: torch.Tensor,
    dof_conn: torch.Tensor,
    pattern: CsrPattern,
    row_points: torch.Tensor | None = None,
    rows_per_tile: int = 336,
    ordering: str = "auto",
    tile_shape: tuple | None = None,
    elem_ids: bool = False,
) -> TilePlan:
    """Partition CSR rows into tiles and precompute everything the fused kernel gathers.

    geom_conn (N,3): rows of `coords` of each element's vertices (batch offsets applied).
    dof_conn  (N,3): global DOF (CSR row/col) of each element vertex.
    row_points (n_dof,2+): a position per DOF, only used to cluster rows spatially; with
    None, tiles are runs of consecutive DOF ids.
    ordering: "auto" = index-space blocks when the numbering is a lattice (`detect_lattice`), else
    "block" (square spatial blocks); "morton" / "natural" = balanced chunks along a Z-order curve /
    of consecutive ids.  `tile_shape` = (bx, by) overrides the lattice block extent.
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
    #    (isolated vertices, ghost columns of a multi-GPU owner) carry no work and are put into
    #    tiles of their own afterwards so that their (zero) load entry is still written.
    active = torch.zeros(n_dof, dtype=torch.bool, device=device)
    active[dconn.reshape(-1)] = True
    active_rows = torch.nonzero(active, as_tuple=True)[0]
    n_active = int(active_rows.numel())
    tile_of_active = None
    lattice = None
    if ordering in ("auto", "lattice") and n_active:
        width = detect_lattice(pattern)
        if width:
            bx, by = tile_shape if tile_shape is not None else lattice_shape(rows_per_tile)
            nbx = (width + bx - 1) // bx  # a leftover column / row of blocks is simply narrower
            li, lj = active_rows % width, torch.div(active_rows, width, rounding_mode="floor")
            ti = torch.div(li, bx, rounding_mode="floor")
            tj = torch.div(lj, by, rounding_mode="floor")
            used, tile_of_active = torch.unique(tj * nbx + ti, return_inverse=True)
            n_tiles = int(used.numel())
            lattice = (width, bx, by)
        elif ordering == "lattice":
            raise ValueError("the DOF numbering is not a lattice: use ordering='block'")
    if tile_of_active is None and row_points is not None and ordering in ("auto", "block") and n_active:
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
    # rows without elements get tiles of their own (no elements, only rows): dealing them out over the
    # other tiles would make otherwise congruent tiles differ and defeat the template sharing
    idle_rows = torch.nonzero(~active, as_tuple=True)[0]
    if idle_rows.numel():
        if n_active == 0:
            n_tiles = 0
        tile_of_row[idle_rows] = n_tiles + torch.div(arange(idle_rows.numel()), rows_per_tile, rounding_mode="floor")
        n_tiles += (int(idle_rows.numel()) + rows_per_tile - 1) // rows_per_tile
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
    if max_vert > MAX_VERT or max_elem > MAX_ELEM:
        raise TileTooLarge(f"tile too large (vertices {max_vert} > {MAX_VERT} or elements {max_elem} > {MAX_ELEM}): lower rows_per_tile")
    tile_elem = local_v[:, 0] | (local_v[:, 1] << 10) | (local_v[:, 2] << 20)
    base_vertex = torch.div(n_v, 2, rounding_mode="floor")  # tile-local index of a vertex near the middle of the id range

    # 4. rows of each tile, ascending row id; runs of consecutive rows = contiguous CSR ranges,
    #    cut into segments of at most 32 entries (one warp pass each)
    row_sorted = torch.argsort(tile_of_row * n_dof + arange(n_dof))
    row_tile = tile_of_row[row_sorted]
    row_ptr = _ptr(row_tile, n_tiles)
    n_r = row_ptr[1:] - row_ptr[:-1]
    row_len = crow[row_sorted + 1] - crow[row_sorted]
    new_run = torch.ones(n_dof, dtype=torch.bool, device=device)
    if n_dof > 1:
        new_run[1:] = (row_tile[1:] != row_tile[:-1]) | (row_sorted[1:] != row_sorted[:-1] + 1)
    run_first = torch.nonzero(new_run, as_tuple=True)[0]
    run_last = torch.cat([run_first[1:], torch.tensor([n_dof], device=device)]) - 1
    run_of_row = torch.cumsum(new_run.long(), 0) - 1
    whole_start = crow[row_sorted[run_first]]
    whole_len = crow[row_sorted[run_last] + 1] - whole_start
    pieces = torch.div(whole_len + SEG - 1, SEG, rounding_mode="floor")
    piece0 = _excl_cumsum(pieces)  # global index of each run's first segment
    piece_run = torch.repeat_interleave(arange(whole_len.numel()), pieces)
    piece_k = arange(piece_run.numel()) - piece0[piece_run]
    seg_tile = row_tile[run_first][piece_run]
    seg_ptr = _ptr(seg_tile, n_tiles)
    n_s = seg_ptr[1:] - seg_ptr[:-1]
    seg_start = whole_start[piece_run] + SEG * piece_k
    if int(n_s.max().item()) > MAX_SEGS:
        raise TileTooLarge("tile has too many entry segments: lower rows_per_tile")

    # 5. every CSR entry of a tile (tile-ordered rows, columns ascending) with its (segment, lane)
    #    and its contributions (element ascending = reference order)
    nnz = pattern.nnz
    row_out = _excl_cumsum(row_len)
    ent_row = torch.repeat_interleave(arange(n_dof), row_len)  # tile-ordered row index of each entry
    ent_k = arange(nnz) - row_out[ent_row]
    ent_global = crow[row_sorted[ent_row]] + ent_k
    ent_tile = row_tile[ent_row]
    ent_run = run_of_row[ent_row]
    ent_off = ent_global - whole_start[ent_run]
    ent_seg = piece0[ent_run] + torch.div(ent_off, SEG, rounding_mode="floor") - seg_ptr[ent_tile]  # tile-local segment
    ent_code = ent_seg * SEG + ent_off % SEG
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
    # the CTA's local table: row 0 stays zero (code 0 = "no contribution"), row 1 + l holds the tile's l-th element
    c_loc = torch.searchsorted(pair_keys, c_tile * n_el + c_e) - elem_ptr[c_tile]  # in ascending element id

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
    lc_loc = torch.searchsorted(pair_keys, lc_tile * n_el + lc_e) - elem_ptr[lc_tile]

    # 6a. table layout: element order inside a tile and base shifts of the off-diagonal arrays, chosen on a
    #     representative tile (the commonest tile shape) by counting shared-memory bank conflicts
    order, shift, layout_stats = 1, (0, 0, 0), None
    forced = os.environ.get("TFEM_TILE_LAYOUT")
    if forced:
        o, s1, s2 = (int(v) for v in forced.split(","))
        order, shift = o, (0, s1, s2)
    elif lattice is not None and n_el:
        signature = (n_e * (MAX_VERT + 1) + n_r) * (MAX_SEGS + 1) + n_s
        values, counts_sig = torch.unique(signature, return_counts=True)
        rep_tile = int(torch.nonzero(signature == values[int(torch.argmax(counts_sig))], as_tuple=True)[0][0])
        order, shift, layout_stats = _layout_for_tile(rep_tile, ent_tile, ent_code, ent_cnt, c_ent, c_within, c_loc, slot, row_tile, row_ptr, lrow_cnt,
                                                      lc_row, lc_within, lc_loc, lc_k, int(n_s[rep_tile]), int(n_r[rep_tile]), int(n_e[rep_tile]), max_elem)
    n_rows_table, od_base, table_bytes = table_layout(max_elem, shift)

    def table_row(loc, tile):  # table row of the tile's loc-th element (ascending id) under the chosen order
        per = torch.div(n_e[tile] + order - 1, order, rounding_mode="floor")
        return (loc % order) * per + torch.div(loc, order, rounding_mode="floor") + 1

    c_code = slot_code(table_row(c_loc, c_tile), slot, od_base)
    lc_code = slot_code(table_row(lc_loc, lc_tile), lc_k, od_base)
    elem_local = arange(pair_tile.numel()) - elem_ptr[pair_tile]
    elem_row = table_row(elem_local, pair_tile) - 1  # position of each tile element in the TB connectivity list

    # 6b. who sums which entry.  One lane per entry handles entries with <= 2 contributions (every
    #     off-diagonal entry of a manifold mesh) from ONE packed word, without a loop; the diagonal of
    #     a row is summed by the row's thread together with its load entry (same element list);
    #     anything else with > 2 contributions (non-manifold edges, degenerate elements) goes to a
    #     short "heavy" list handled by a generic loop.
    ent_is_diag = pattern.col.long()[ent_global] == row_sorted[ent_row]
    off_slot = torch.zeros(nnz, dtype=torch.int64, device=device).index_add_(0, c_ent, (slot >= 3).long())
    diag_fast = ent_is_diag & (ent_cnt > 2) & (off_slot == 0) & (ent_cnt == lrow_cnt[ent_row])
    row_diag = torch.full((n_dof,), 0xFFFF, dtype=torch.int64, device=device)
    row_diag[ent_row[diag_fast]] = ent_code[diag_fast]
    heavy = (ent_cnt > 2) & ~diag_fast
    heavy_tile = ent_tile[heavy]
    n_h = torch.bincount(heavy_tile, minlength=n_tiles)
    heavy_local = arange(heavy_tile.numel()) - _excl_cumsum(n_h)[heavy_tile]
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

    # 6c. the packed pair of every entry: lo 16 bits = first contribution, hi 16 = second (0 when
    #     absent); 0xFFFFFFFF = "not mine" (diagonal done by the row thread, heavy, or a padding lane)
    first = torch.zeros(nnz, dtype=torch.int64, device=device)
    second = torch.zeros(nnz, dtype=torch.int64, device=device)
    first[c_ent[c_within == 0]] = c_code[c_within == 0]
    second[c_ent[c_within == 1]] = c_code[c_within == 1]
    pair = first | (second << 16)
    pair[ent_cnt > 2] = 0xFFFFFFFF

    # 6d. per-row chunks of 8 x u16: 7 contribution codes ((element+1)*80 + 16 k, k = local vertex; 0 pads) and the
    #     tile-local index of the row's next chunk (0 = none).  Chunk j < n_rows is the first chunk of
    #     row j; rows with more than 7 elements continue in chunks appended after n_rows.
    row_chunks = torch.div(lrow_cnt + PER_CHUNK - 1, PER_CHUNK, rounding_mode="floor").clamp_min(1)
    row_extra = row_chunks - 1
    extra_off = _excl_cumsum(row_extra)
    extra_in_tile = extra_off - extra_off[row_ptr[:-1].clamp_max(max(n_dof - 1, 0))][row_tile]
    n_x = torch.zeros(n_tiles, dtype=torch.int64, device=device).index_add_(0, row_tile, row_extra)
    n_ch = n_r + n_x
    if int(n_ch.max().item()) > 65535:
        raise ValueError("tile has more than 65535 row chunks: lower rows_per_tile")
    chunk0 = _excl_cumsum(n_ch)
    total_ch = int(n_ch.sum().item())
    chunks = torch.zeros((total_ch, 8), dtype=torch.int64, device=device)
    row_local = arange(n_dof) - row_ptr[row_tile]
    lc_q = torch.div(lc_within, PER_CHUNK, rounding_mode="floor")
    lc_chunk_local = torch.where(lc_q == 0, row_local[lc_row], n_r[lc_tile] + extra_in_tile[lc_row] + lc_q - 1)
    chunks[chunk0[lc_tile] + lc_chunk_local, lc_within - PER_CHUNK * lc_q] = lc_code
    link_row = torch.repeat_interleave(arange(n_dof), row_extra)
    link_q = arange(link_row.numel()) - extra_off[link_row]  # 0-based: link from chunk q to q+1
    link_tile = row_tile[link_row]
    link_to = n_r[link_tile] + extra_in_tile[link_row] + link_q
    link_from = torch.where(link_q == 0, row_local[link_row], link_to - 1)
    chunks[chunk0[link_tile] + link_from, 7] = link_to
    chunk_tile = torch.repeat_interleave(tiles, n_ch * 8)
    chunk_local = arange(total_ch * 8) - (chunk0 * 8)[chunk_tile]

    # 7. pack: instances, and the two template blobs of every tile
    def header(columns, words):
        stacked = torch.stack(columns + [torch.zeros_like(n_v)] * (words - len(columns)), dim=1).reshape(-1)
        tile = torch.repeat_interleave(tiles, words)
        return _Section("header", 32, torch.full_like(n_v, words), tile, arange(tile.numel()) % words, stacked)

    seg_local = arange(seg_tile.numel()) - seg_ptr[seg_tile]
    inst_sections = [
        header([n_v, n_s, n_r, base_vertex], INST_HEADER_WORDS),
        _Section("vert", 32, n_v, vert_tile, arange(vert_tile.numel()) - vert_ptr[vert_tile], tile_vert),
        _Section("seg_start", 32, n_s, seg_tile, seg_local, seg_start),
        _Section("row_id", 32, n_r, row_tile, row_local, row_sorted),
    ]
    if elem_ids:  # global id of every tile element, in the order of the TB connectivity list: sampled sources
        # (f at the quadrature points of element e) and per-fracture metrics (fracture = e / elements per fracture)
        inst_sections.append(_Section("elem_id", 32, n_e, pair_tile, elem_row, pair_elem))
    tb_sections = [
        header([n_v, n_e, n_r, n_s, n_ch, n_h, n_hc], TB_HEADER_WORDS),
        _Section("elem", 32, n_e, pair_tile, elem_row, tile_elem),
    ]
    tc_sections = [
        _Section("pair", 32, n_s * SEG, ent_tile, ent_code, pair, fill=0xFFFFFFFF),
        _Section("row_chunk", 16, n_ch * 8, chunk_tile, chunk_local, chunks.reshape(-1)),
        _Section("row_diag", 16, n_r, row_tile, row_local, row_diag),
        _Section("heavy_seg", 16, n_h + 1, hseg_tile, hseg_local, hseg_value),
        _Section("heavy_contrib", 16, n_hc, hc_tile, arange(total_hc) - tile_hc0[hc_tile], hc_code),
        _Section("heavy_pos", 16, n_h, heavy_tile, heavy_local, ent_code[heavy]),
    ]
    inst_off, inst_blob, inst_words = _pack(inst_sections, n_tiles, device)
    tb_off, tb_blob, tb_words = _pack(tb_sections, n_tiles, device)
    tc_off, tc_blob, tc_words = _pack(tc_sections, n_tiles, device)

    # 8. congruent tiles share one template
    signature = torch.cat([_hash_tiles(tb_off, tb_blob, n_tiles, 1), _hash_tiles(tc_off, tc_blob, n_tiles, 2)], dim=1)
    _, template_of, counts = torch.unique(signature, dim=0, return_inverse=True, return_counts=True)
    n_templates = int(counts.numel())
    # representative = first tile of each template; templates numbered by decreasing population
    by_count = torch.argsort(counts, descending=True, stable=True)
    rank = torch.empty_like(by_count)
    rank[by_count] = arange(n_templates)
    template_of = rank[template_of]
    rep = torch.full((n_templates,), n_tiles, dtype=torch.int64, device=device).scatter_reduce_(0, template_of, tiles, reduce="amin")
    rep_tb_words, rep_tc_words = tb_words[rep], tc_words[rep]
    tpl_words = torch.stack([rep_tb_words, rep_tc_words], dim=1).reshape(-1)
    tpl_offsets = _excl_cumsum(tpl_words).reshape(-1, 2)
    tpl_blob = torch.empty(int(tpl_words.sum().item()), dtype=torch.int64, device=device)
    tb_part = _gather_ranges(tb_blob, tb_off[rep], rep_tb_words)
    tc_part = _gather_ranges(tc_blob, tc_off[rep], rep_tc_words)
    tb_dst = torch.repeat_interleave(tpl_offsets[:, 0], rep_tb_words) + (arange(tb_part.numel()) - torch.repeat_interleave(_excl_cumsum(rep_tb_words), rep_tb_words))
    tc_dst = torch.repeat_interleave(tpl_offsets[:, 1], rep_tc_words) + (arange(tc_part.numel()) - torch.repeat_interleave(_excl_cumsum(rep_tc_words), rep_tc_words))
    tpl_blob[tb_dst] = tb_part
    tpl_blob[tc_dst] = tc_part
    tpl_desc = torch.stack([tpl_offsets[:, 0], rep_tb_words, tpl_offsets[:, 1], rep_tc_words], dim=1).to(torch.int32).contiguous()
    base_row = tile_vert[(vert_ptr[:-1] + base_vertex).clamp_max(max(tile_vert.numel() - 1, 0))] if tile_vert.numel() else torch.zeros_like(n_v)
    base_row = torch.where(n_v > 0, base_row, torch.zeros_like(base_row))
    tile_desc = torch.stack([inst_off[:-1], inst_words, template_of, base_row], dim=1).to(torch.int32).contiguous()
    default_order = torch.argsort(template_of * n_tiles + tiles).to(torch.int32).contiguous()

    return TilePlan(
        n_tiles=n_tiles,
        n_templates=n_templates,
        tile_desc=tile_desc,
        inst_blob=_wrap_u32(inst_blob),
        tpl_desc=tpl_desc,
        tpl_blob=_wrap_u32(tpl_blob),
        default_order=default_order,
        max_vert=max_vert,
        max_elem=max_elem,
        max_rows=int(n_r.max().item()),
        max_segs=int(n_s.max().item()),
        max_inst_words=int(inst_words.max().item()),
        max_tb_words=int(tb_words.max().item()),
        max_tc_words=int(tc_words.max().item()),
        halo_factor=float(pair_keys.shape[0]) / max(n_el, 1),
        index_bytes=4 * (inst_blob.numel() + tpl_blob.numel() + tile_desc.numel() + tpl_desc.numel() + default_order.numel()),
        lattice=lattice,
        tile_of_row=tile_of_row,
        table_bytes=table_bytes,
        od_base=tuple(od_base),
        elem_order=order,
        layout_stats=layout_stats,
        has_elem_ids=bool(elem_ids),
    )
