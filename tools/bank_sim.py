#!/usr/bin/env python
"""Shared-memory wavefront model of the tiled kernel's reduction phase for one tile of a plan.

Counts, per warp-wide shared-memory load of phase C, how many 128 B wavefronts the access needs on a
32-bank x 4 B shared memory (64-bit loads are served per half-warp, 128-bit loads per quarter-warp; lanes
reading the same address share a wavefront).  Used to choose the table layout (row stride, element order)
that minimises bank conflicts; compare with `L1 Wavefronts Shared` of an ncu source-page export.
Measurement helper, not part of the product path."""
import sys

import numpy as np


def wavefronts(addresses, width, active=None):
    """addresses: (n_instr, 32) byte addresses; width 4, 8 or 16 bytes; active: bool mask or None."""
    addresses = np.asarray(addresses, dtype=np.int64)
    n = addresses.shape[0]
    if active is None:
        active = np.ones_like(addresses, dtype=bool)
    group = {4: 32, 8: 16, 16: 8}[width]
    total = 0
    for g in range(32 // group):
        a = addresses[:, g * group : (g + 1) * group]
        m = active[:, g * group : (g + 1) * group]
        worst = np.zeros(n, dtype=np.int64)
        words = width // 4
        for w in range(words):
            bank = ((a // 4) + w) % 32
            for b in range(32):
                hit = (bank == b) & m
                # distinct addresses per bank
                cnt = np.zeros(n, dtype=np.int64)
                for i in np.nonzero(hit.any(axis=1))[0]:
                    cnt[i] = len(np.unique(a[i][hit[i]]))
                worst = np.maximum(worst, cnt)
        total += int(np.maximum(worst, m.any(axis=1).astype(np.int64)).sum())
    return total


def phase_c(sec, consumers=384, row_bytes=80, remap=None):
    """Wavefronts of the table loads of phase C for one decoded tile (`TilePlan.sections`).
    remap: optional array mapping a table row (1 + element) to another row (element renumbering)."""
    def conv(code):
        code = np.asarray(code, dtype=np.int64)
        row, off = code // 80, code % 80
        if remap is not None:
            row = remap[row]
        return row * row_bytes + off

    pair = sec["pair"].reshape(-1, 32)
    mine = pair != 0xFFFFFFFF
    first = conv(np.where(mine, pair & 0xFFFF, 0))
    second = conv(np.where(mine, pair >> 16, 0))
    entries = wavefronts(first, 8) + wavefronts(second, 8)
    chunks = sec["row_chunk"].reshape(-1, 8)[: sec["n_rows"], :7]
    n_rows = sec["n_rows"]
    pad = (-n_rows) % 32
    codes = np.concatenate([chunks, np.zeros((pad, 7), dtype=chunks.dtype)]).reshape(-1, 32, 7)
    act = np.concatenate([np.ones(n_rows, bool), np.zeros(pad, bool)]).reshape(-1, 32)
    rows = sum(wavefronts(conv(codes[:, :, k]), 16, act) for k in range(7))
    return entries, rows


if __name__ == "__main__":
    import os

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch

    from pytorch_fem_solver_b200 import csr, meshgen

    nx, ny = 256, 128
    mesh = meshgen.structured_rectangle(nx, ny, jitter=0.25, seed=1234, topology=False)
    conn = torch.from_numpy(mesh["triangles"])
    coords = torch.from_numpy(mesh["vertices"])
    pat = csr.build_pattern(conn, coords.shape[0])
    plan = csr.build_tile_plan(conn, conn, pat, coords, 336, "auto", tile_shape=(21, 16))
    tile = int(torch.nonzero(plan.tile_desc[:, 2] == 0)[0, 0])
    sec = plan.sections(tile)
    print("tile", tile, "elements", sec["n_elem"], "rows", sec["n_rows"], "segments", sec["n_segs"])
    for rb in (72, 80, 88, 96, 104, 112, 120, 136):
        e, r = phase_c(sec, row_bytes=rb)
        print(f"row bytes {rb:4d}: entry loads {e:5d} wavefronts ({e / (2 * sec['n_segs']):.2f} per load), row loads {r:5d} ({r / (7 * ((sec['n_rows'] + 31) // 32)):.2f} per load)")


def deinterleave(n, s):
    """Permutation position -> new position ordering elements by (position mod s, position div s)."""
    pos = np.arange(n)
    order = np.lexsort((pos // s, pos % s))
    new = np.empty(n, dtype=np.int64)
    new[order] = np.arange(n)
    return new
