"""CPU walk-through of the fused tile kernel's phases, driven by the oracle's local matrices.

Test infrastructure: lets the `-m "not gpu"` suite validate a tile plan (tileplan.build_tile_plan)
without a device.  It decodes the per-tile blobs exactly as assemble_tiled.cu does.
"""

import numpy as np


def emulate_tiled(plan, coords, local_mat, local_vec, geom_conn, nnz, n_dof):
    """local_mat (N,3,3) symmetric, local_vec (N,3) from the oracle, indexed by GLOBAL element.

    The emulator recovers each tile element's global id by matching tile-local vertices, so
    it also checks the vertex / connectivity sections for consistency."""
    for off in (plan.e_off, plan.la_off, plan.lb_off):
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
        # phase C: one thread per CSR entry adds the (at most two) contributions packed in its word
        stride = plan.elem_stride
        assert stride > n_elem, "the last column of the local-matrix table must stay free (zeros)"
        table = np.zeros((9, stride))
        table[:, :n_elem] = sloc
        zero_code = stride - 1

        def store(position, value):
            assert np.isnan(csr_val[position]), "every CSR entry is written exactly once"
            csr_val[position] = value

        def lookup(code, previous):
            slot, el = divmod(int(code), stride)
            if code != zero_code:
                assert slot < 6 and previous <= el < n_elem, "contributions must come in increasing element order"
            return table[slot, el], el

        end = 0
        for start, meta in zip(sec["run_start"], sec["run_meta"]):  # light entries, run by run
            base, length = meta & 0xFFFF, meta >> 16
            assert base == end, "runs must tile the image in order"
            end = base + length
            for i in range(length):
                word = int(sec["pair"][base + i])
                if word != 0xFFFFFFFF:
                    first, el = lookup(word & 0xFFFF, -1)
                    second, _ = lookup(word >> 16, el if (word & 0xFFFF) != zero_code else -1)
                    store(start + i, first + second)
        assert end == sec["n_out"]
        hseg = sec["heavy_seg"]
        assert hseg[0] == 0 and hseg[-1] == sec["n_heavy_contrib"]
        for h, position in enumerate(sec["heavy_pos"]):  # generic loop
            assert hseg[h + 1] - hseg[h] > 2
            acc, previous = 0.0, -1
            for code in sec["heavy_contrib"][hseg[h] : hseg[h + 1]]:
                value, previous = lookup(code, previous)
                acc += value
            store(position, acc)
        chunks = sec["row_chunk"].reshape(-1, 8)
        visited = np.zeros(len(chunks), dtype=bool)
        for j, row in enumerate(sec["row_id"]):  # load entry + diagonal share the row's element list
            rhs = diag = 0.0
            ch = j
            while True:
                assert not visited[ch]
                visited[ch] = True
                for code in chunks[ch, :7]:
                    k, el = divmod(int(code), stride)
                    assert k < 3 and (el < n_elem or code == zero_code)
                    rhs += table[6 + k, el]
                    diag += table[k, el]
                ch = int(chunks[ch, 7])
                if ch == 0:
                    break
                assert ch >= sec["n_rows"]
            load[row] = rhs
            if sec["row_diag"][j] != 0xFFFFFFFF:
                store(sec["row_diag"][j], diag)
        assert visited.all()
    return csr_val, load
