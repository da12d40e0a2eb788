"""CPU walk-through of the fused tile kernel's phases, driven by the oracle's local matrices.

Test infrastructure: lets the `-m "not gpu"` suite validate a tile plan (tileplan.build_tile_plan)
without a device.  It decodes the per-tile instance and its template exactly as assemble_tiled.cu does.
"""

import numpy as np

SEG = 32
N_SLOTS = 9


def emulate_tiled(plan, coords, local_mat, local_vec, geom_conn, nnz, n_dof, order=None):
    """local_mat (N,3,3) symmetric, local_vec (N,3) from the oracle, indexed by GLOBAL element.

    The emulator recovers each tile element's global id by matching tile-local vertices, so
    it also checks the vertex / connectivity sections for consistency."""
    desc = plan.tile_desc.cpu().numpy()
    tdesc = plan.tpl_desc.cpu().numpy()
    assert np.all(desc[:, 0] % 4 == 0) and np.all(desc[:, 1] % 4 == 0), "instances must be whole 16 B units (TMA bulk copy)"
    assert np.all(tdesc % 4 == 0), "template blobs must be whole 16 B units (TMA bulk copy)"
    assert desc[:, 1].max() <= plan.max_inst_words and tdesc[:, 1].max() <= plan.max_tb_words and tdesc[:, 3].max() <= plan.max_tc_words
    elem_of = {tuple(v): e for e, v in enumerate(np.asarray(geom_conn).reshape(-1, 3).tolist())}
    csr_val = np.full(nnz, np.nan)
    load = np.full(n_dof, np.nan)
    tiles = plan.default_order.cpu().numpy().tolist() if order is None else list(order)
    if order is None:
        assert sorted(tiles) == list(range(plan.n_tiles))
    for t in tiles:
        sec = plan.sections(t)
        verts = sec["vert"]
        n_elem = sec["n_elem"]
        assert sec["n_vert"] <= plan.max_vert and n_elem <= plan.max_elem
        assert sec["inst_header"][:3] == [sec["n_vert"], sec["n_segs"], sec["n_rows"]]
        if sec["n_vert"]:
            assert 0 <= sec["base_vertex"] < sec["n_vert"]
        # phase B: every tile element integrated once -> table[1 + el][9]; row 0 stays zero
        table = np.zeros((n_elem + 1, N_SLOTS))
        for el, w in enumerate(sec["elem"]):
            a, b, c = w & 1023, (w >> 10) & 1023, (w >> 20) & 1023
            ge = elem_of[(verts[a], verts[b], verts[c])]
            m = local_mat[ge]
            table[el + 1] = [m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[1, 2], m[2, 0], *local_vec[ge]]

        def store(code, value):
            position = sec["seg_start"][code // SEG] + code % SEG
            assert np.isnan(csr_val[position]), "every CSR entry is written exactly once"
            csr_val[position] = value

        def lookup(code, previous):
            el, slot = divmod(int(code), N_SLOTS)
            if code != 0:
                assert slot < 6 and max(previous, 1) <= el <= n_elem, "contributions must come in increasing element order"
            return table[el, slot], el

        # phase C: one lane per CSR entry adds the (at most two) contributions packed in its word
        pair = sec["pair"].reshape(-1, SEG)
        assert pair.shape[0] == sec["n_segs"]
        for s in range(sec["n_segs"]):
            for lane in range(SEG):
                word = int(pair[s, lane])
                if word != 0xFFFFFFFF:
                    first, el = lookup(word & 0xFFFF, 0)
                    second, _ = lookup(word >> 16, el)
                    store(s * SEG + lane, first + second)
        hseg = sec["heavy_seg"]
        assert hseg[0] == 0 and hseg[-1] == sec["n_heavy_contrib"]
        for h, code in enumerate(sec["heavy_pos"]):  # generic loop
            assert hseg[h + 1] - hseg[h] > 2
            acc, previous = 0.0, 0
            for contribution in sec["heavy_contrib"][hseg[h] : hseg[h + 1]]:
                value, previous = lookup(contribution, previous)
                acc += value
            store(int(code), acc)
        chunks = sec["row_chunk"].reshape(-1, 8)
        visited = np.zeros(len(chunks), dtype=bool)
        for j, row in enumerate(sec["row_id"]):  # load entry + diagonal share the row's element list
            rhs = diag = 0.0
            ch = j
            while True:
                assert not visited[ch]
                visited[ch] = True
                for code in chunks[ch, :7]:
                    el, k = divmod(int(code), N_SLOTS)
                    assert (k < 3 and 0 < el <= n_elem) or code == 0
                    rhs += table[el, 6 + k]
                    diag += table[el, k]
                ch = int(chunks[ch, 7])
                if ch == 0:
                    break
                assert ch >= sec["n_rows"]
            assert np.isnan(load[row])
            load[row] = rhs
            if sec["row_diag"][j] != 0xFFFF:
                store(int(sec["row_diag"][j]), diag)
        assert visited.all()
    return csr_val, load
