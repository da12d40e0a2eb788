"""CPU walk-through of the fused tile kernel's phases, driven by the oracle's local matrices.

Test infrastructure: lets the `-m "not gpu"` suite validate a tile plan (csr.build_tile_plan)
without a device.  It decodes the per-tile blobs exactly as assemble_tiled.cu does.
"""

import numpy as np


def emulate_tiled(plan, coords, local_mat, local_vec, geom_conn, nnz, n_dof):
    """local_mat (N,3,3) symmetric, local_vec (N,3) from the oracle, indexed by GLOBAL element.

    The emulator recovers each tile element's global id by matching tile-local vertices, so
    it also checks the vertex / connectivity sections for consistency."""
    off = plan.tile_off.cpu().numpy()
    assert np.all(off % 4 == 0), "blobs must start on 16 B boundaries (TMA bulk copy)"
    elem_of = {tuple(v): e for e, v in enumerate(np.asarray(geom_conn).reshape(-1, 3).tolist())}
    csr_val = np.full(nnz, np.nan)
    load = np.full(n_dof, np.nan)
    for t in range(plan.n_tiles):
        sec = plan.sections(t)
        assert off[t + 1] - off[t] <= plan.max_blob_words
        verts = sec["vert"]
        n_elem = len(sec["elem"])
        assert len(verts) <= plan.max_vert and n_elem <= plan.max_elem and sec["n_out"] <= plan.max_out
        sloc = np.zeros((9, n_elem))
        for el, w in enumerate(sec["elem"]):
            a, b, c = w & 1023, (w >> 10) & 1023, (w >> 20) & 1023
            ge = elem_of[(verts[a], verts[b], verts[c])]
            m = local_mat[ge]
            sloc[:, el] = [m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[1, 2], m[2, 0], *local_vec[ge]]
        sout = np.zeros(sec["n_out"])
        cptr = sec["row_cptr"]
        for j, (row, meta) in enumerate(zip(sec["row_id"], sec["row_meta"])):
            base, pd = meta & 0xFFFF, (meta >> 16) & 0xFF
            diag = rhs = 0.0
            for cw in sec["corner"][cptr[j] : cptr[j + 1]]:
                el, k, pa, pb = cw & 0xFFF, (cw >> 12) & 3, (cw >> 16) & 0xFF, cw >> 24
                kb = 2 if k == 0 else k - 1
                diag += sloc[k, el]
                sout[base + pa] += sloc[3 + k, el]
                sout[base + pb] += sloc[3 + kb, el]
                rhs += sloc[6 + k, el]
            if cptr[j + 1] > cptr[j]:
                sout[base + pd] += diag
            load[row] = rhs
        end = 0
        for start, meta in zip(sec["run_start"], sec["run_meta"]):
            base, length = meta & 0xFFFF, meta >> 16
            assert base == end, "runs must tile the image in order"
            end = base + length
            csr_val[start : start + length] = sout[base : base + length]
        assert end == sec["n_out"]
    return csr_val, load
