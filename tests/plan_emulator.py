"""CPU walk-through of the fused tile kernel's phases, driven by the oracle's local matrices.

Test infrastructure: lets the `-m "not gpu"` suite validate a tile plan (tileplan.build_tile_plan)
without a device.  It decodes the per-tile instance and its template exactly as assemble_tiled.cu does.
"""

import numpy as np

SEG = 32


def decode(code, plan):
    """(table row, slot) of a byte code: slot 0..5 = K00 K11 K22 K01 K12 K20, 6..8 = load (include/tfem_b200.h)."""
    rows = plan.max_elem + 2
    code = int(code)
    if code < 48 * rows:
        row, rest = divmod(code, 48)
        k, which = divmod(rest, 16)
        assert which in (0, 8)
        return row, k + (6 if which else 0)
    for m, base in enumerate(plan.od_base):
        if base <= code < base + 8 * rows:
            assert (code - base) % 8 == 0
            return (code - base) // 8, 3 + m
    raise AssertionError(f"code {code} outside the table")


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
        assert desc[t, 3] == (verts[sec["base_vertex"]] if sec["n_vert"] else 0)
        # phase B: every tile element integrated once -> table[1 + el][9]; row 0 stays zero
        table = np.zeros((n_elem + 1, 9))
        global_id = np.full(n_elem + 1, -1)
        for el, w in enumerate(sec["elem"]):
            a, b, c = w & 1023, (w >> 10) & 1023, (w >> 20) & 1023
            ge = elem_of[(verts[a], verts[b], verts[c])]
            m = local_mat[ge]
            table[el + 1] = [m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[1, 2], m[2, 0], *local_vec[ge]]
            global_id[el + 1] = ge
            if plan.has_elem_ids:
                assert sec["elem_id"][el] == ge, "elem_id must name the element of the same position in elem[]"
        assert len(set(global_id[1:].tolist())) == n_elem
        assert plan.table_bytes >= max(plan.od_base) + 8 * (plan.max_elem + 2) and all(b % 8 == 0 for b in plan.od_base)

        def store(code, value):
            position = sec["seg_start"][code // SEG] + code % SEG
            assert np.isnan(csr_val[position]), "every CSR entry is written exactly once"
            csr_val[position] = value

        def lookup(code, previous):
            if code == 0:
                return 0.0, previous
            el, slot = decode(code, plan)
            assert slot < 6 and 1 <= el <= n_elem and previous <= global_id[el], "contributions must come in increasing element order"
            return table[el, slot], global_id[el]

        # phase C: one lane per CSR entry adds the (at most two) contributions packed in its word
        pair = sec["pair"].reshape(-1, SEG)
        assert pair.shape[0] == sec["n_segs"]
        for s in range(sec["n_segs"]):
            for lane in range(SEG):
                word = int(pair[s, lane])
                if word != 0xFFFFFFFF:
                    first, el = lookup(word & 0xFFFF, -1)
                    second, _ = lookup(word >> 16, el)
                    store(s * SEG + lane, first + second)
        hseg = sec["heavy_seg"]
        assert hseg[0] == 0 and hseg[-1] == sec["n_heavy_contrib"]
        for h, code in enumerate(sec["heavy_pos"]):  # generic loop
            assert hseg[h + 1] - hseg[h] > 2
            acc, previous = 0.0, -1
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
                    if code == 0:
                        continue
                    el, k = decode(code, plan)
                    assert k < 3 and 0 < el <= n_elem
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
