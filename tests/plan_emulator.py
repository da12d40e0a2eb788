"""CPU walk-through of the fused tile kernel's phases, driven by the oracle's local matrices.

Test infrastructure: lets the `-m "not gpu"` suite validate a tile plan (tileplan.build_tile_plan)
without a device.  It decodes the per-tile blobs exactly as assemble_tiled.cu does.
"""

import numpy as np


def emulate_tiled(plan, coords, local_mat, local_vec, geom_conn, nnz, n_dof):
    """local_mat (N,3,3) symmetric, local_vec (N,3) from the oracle, indexed by GLOBAL element.

    The emulator recovers each tile element's global id by matching tile-local vertices, so
    it also checks the vertex / connectivity sections for consistency."""
    for off in (plan.e_off, plan.l_off):
        assert np.all(off.cpu().numpy() % 4 == 0), "blobs must start on 16 B boundaries (TMA bulk copy)"
    elem_of = {tuple(v): e for e, v in enumerate(np.asarray(geom_conn).reshape(-1, 3).tolist())}
    csr_val = np.full(nnz, np.nan)
    load = np.full(n_dof, np.nan)
    for t in range(plan.n_tiles):
        sec = plan.sections(t)
        verts = sec["vert"]
        n_elem = sec["n_elem"]
        assert sec["n_vert"] <= plan.max_vert and n_elem <= plan.max_elem
        if sec["n_vert"]:
            assert sec["base_vertex"] in verts
        # phase B: every tile element integrated once -> sloc[9][n_elem]
        sloc = np.zeros((9, n_elem))
        for el, w in enumerate(sec["elem"]):
            a, b, c = w & 1023, (w >> 10) & 1023, (w >> 20) & 1023
            ge = elem_of[(verts[a], verts[b], verts[c])]
            m = local_mat[ge]
            sloc[:, el] = [m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[1, 2], m[2, 0], *local_vec[ge]]
        # phase C: one thread per CSR entry sums its contributions left to right
        seg, contrib = sec["ent_seg"], sec["contrib"]
        assert seg[0] == 0 and seg[-1] == sec["n_contrib"]
        stride = plan.elem_stride
        assert stride >= n_elem

        def store(position, value):
            assert np.isnan(csr_val[position]), "every CSR entry is written exactly once"
            csr_val[position] = value

        def entry_sum(o):
            acc, previous = 0.0, -1
            for code in contrib[seg[o] : seg[o + 1]]:
                slot, el = divmod(int(code), stride)
                assert slot < 6 and el >= previous, "contributions must come in increasing element order"
                previous = el
                acc += sloc[slot, el]
            return acc

        end = 0
        for start, meta in zip(sec["run_start"], sec["run_meta"]):  # light entries, run by run
            base, length = meta & 0xFFFF, meta >> 16
            assert base == end, "runs must tile the image in order"
            end = base + length
            for i in range(length):
                if seg[base + i + 1] - seg[base + i] <= 2:
                    store(start + i, entry_sum(base + i))
        assert end == sec["n_out"]
        for o, position in zip(sec["heavy"], sec["heavy_pos"]):  # generic loop
            assert seg[o + 1] - seg[o] > 2
            store(position, entry_sum(o))
        lseg = sec["lrow_seg"]
        for j, row in enumerate(sec["row_id"]):  # load entry + diagonal share the row's element list
            rhs = diag = 0.0
            for code in sec["lcontrib"][lseg[j] : lseg[j + 1]]:
                k, el = divmod(int(code), stride)
                assert k < 3
                rhs += sloc[6 + k, el]
                diag += sloc[k, el]
            load[row] = rhs
            if sec["row_diag"][j] != 0xFFFFFFFF:
                store(sec["row_diag"][j], diag)
    return csr_val, load
