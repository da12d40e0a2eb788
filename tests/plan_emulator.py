"""CPU walk-through of the fused tile kernel's phases, driven by the oracle's local matrices.

Test infrastructure: lets the `-m "not gpu"` suite validate a tile plan (csr.build_tile_plan)
without a device.  It reads the plan arrays exactly as assemble_tiled.cu does.
"""

import numpy as np


def emulate_tiled(plan, coords, local_mat, local_vec, geom_conn, nnz, n_dof):
    """local_mat (N,3,3) symmetric, local_vec (N,3) from the oracle, indexed by GLOBAL element.

    The emulator recovers each tile element's global id by matching tile-local vertices, so
    it also checks tile_vert / tile_elem consistency."""
    tp = plan.tile_ptr.cpu().numpy().astype(np.int64)
    tile_vert = plan.tile_vert.cpu().numpy().astype(np.int64)
    tile_elem = plan.tile_elem.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    row_id = plan.row_id.cpu().numpy().astype(np.int64)
    row_meta = plan.row_meta.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    rcp = plan.row_corner_ptr.cpu().numpy().astype(np.int64)
    corner = plan.corner.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    run_start = plan.run_start.cpu().numpy().astype(np.int64)
    run_meta = plan.run_meta.cpu().numpy().astype(np.int64) & 0xFFFFFFFF

    elem_of = {tuple(v): e for e, v in enumerate(np.asarray(geom_conn).reshape(-1, 3).tolist())}
    csr_val = np.full(nnz, np.nan)
    load = np.full(n_dof, np.nan)
    for t in range(plan.n_tiles):
        v0, e0, r0, u0 = tp[t]
        v1, e1, r1, u1 = tp[t + 1]
        verts = tile_vert[v0:v1]
        sloc = np.zeros((9, e1 - e0))
        for el in range(e1 - e0):
            w = tile_elem[e0 + el]
            a, b, c = w & 1023, (w >> 10) & 1023, (w >> 20) & 1023
            ge = elem_of[(verts[a], verts[b], verts[c])]
            m = local_mat[ge]
            sloc[:, el] = [m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[1, 2], m[2, 0], *local_vec[ge]]
        n_out = 0
        if u1 > u0:
            n_out = (run_meta[u1 - 1] & 0xFFFF) + (run_meta[u1 - 1] >> 16)
        sout = np.zeros(n_out)
        for j in range(r1 - r0):
            meta = row_meta[r0 + j]
            base, pd = meta & 0xFFFF, (meta >> 16) & 0xFF
            diag = rhs = 0.0
            for c in range(rcp[r0 + j], rcp[r0 + j + 1]):
                cw = corner[c]
                el, k, pa, pb = cw & 0xFFF, (cw >> 12) & 3, (cw >> 16) & 0xFF, cw >> 24
                kb = 2 if k == 0 else k - 1
                diag += sloc[k, el]
                sout[base + pa] += sloc[3 + k, el]
                sout[base + pb] += sloc[3 + kb, el]
                rhs += sloc[6 + k, el]
            if rcp[r0 + j + 1] > rcp[r0 + j]:
                sout[base + pd] += diag
            load[row_id[r0 + j]] = rhs
        for r in range(u0, u1):
            base, length = run_meta[r] & 0xFFFF, run_meta[r] >> 16
            csr_val[run_start[r] : run_start[r] + length] = sout[base : base + length]
    return csr_val, load
