"""Host-side symbolic phase (CSR pattern, permutation, tile plan) against the oracle; CPU only."""

import numpy as np
import pytest
import torch

from oracle import fem_oracle as fo
from pytorch_fem_solver_b200 import csr, meshgen
from tests.plan_emulator import emulate_tiled


def fan_mesh(n):
    """n triangles around one centre vertex: its row needs several 7-element chunks in the tile plan."""
    angle = 2 * np.pi * np.arange(n) / n
    vertices = np.concatenate([[[0.5, 0.5]], 0.5 + 0.4 * np.stack([np.cos(angle), np.sin(angle)], 1)])
    triangles = np.stack([np.zeros(n, dtype=np.int64), 1 + np.arange(n), 1 + (np.arange(n) + 1) % n], 1).astype(np.int32)
    return {"vertices": vertices, "triangles": triangles}


def nonmanifold_mesh():
    """Three triangles on one edge (entries with 3 contributions), a triangle with a repeated
    vertex (a diagonal that also receives off-diagonal slots) and an isolated vertex."""
    vertices = np.array([[0.0, 0.0], [1.0, 0.0], [0.5, 0.8], [0.5, -0.7], [0.4, 0.3], [0.9, 0.9], [0.1, 0.9], [2.0, 2.0]])
    triangles = np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4], [2, 5, 6], [2, 2, 5], [4, 1, 2]], dtype=np.int32)
    return {"vertices": vertices, "triangles": triangles}


def meshes():
    yield "structured4x4", meshgen.structured_rectangle(4, 4, topology=False)
    yield "structured37x21_jitter", meshgen.structured_rectangle(37, 21, jitter=0.25, seed=2, topology=False)
    yield "delaunay200", meshgen.delaunay_unit_square(200, seed=4)
    yield "permuted", meshgen.permute_mesh(meshgen.structured_rectangle(24, 24, jitter=0.2, topology=False))
    yield "fan23", fan_mesh(23)
    yield "nonmanifold", nonmanifold_mesh()


@pytest.mark.parametrize("name,mesh", list(meshes()))
def test_pattern_matches_oracle(name, mesh):
    conn = mesh["triangles"]
    n_dof = mesh["vertices"].shape[0]
    crow, col, perm, seg = fo.csr_pattern(conn, n_dof)
    pat = csr.build_pattern(torch.from_numpy(conn), n_dof)
    assert np.array_equal(pat.crow.numpy(), crow) and np.array_equal(pat.col.numpy(), col)
    assert np.array_equal(pat.perm.numpy(), perm) and np.array_equal(pat.seg.numpy(), seg)
    rows, cols, form = csr.coo_index_maps(torch.from_numpy(conn))
    o_rows, o_cols, o_form = fo.coo_index_maps(conn)
    assert np.array_equal(rows.numpy(), o_rows) and np.array_equal(cols.numpy(), o_cols)
    assert np.array_equal(form.numpy(), o_form)
    # linear map: lin_perm groups the flat linear_form_idx by DOF, stably
    lp, ls = pat.lin_perm.numpy(), pat.lin_seg.numpy()
    assert np.array_equal(o_form[lp], np.sort(o_form, kind="stable"))
    assert ls[-1] == 3 * conn.shape[0] and np.all(np.diff(lp.reshape(-1)[ls[5] : ls[6]]) > 0)


def test_kat_pattern_4x4():
    mesh = meshgen.structured_rectangle(4, 4, topology=False)
    pat = csr.build_pattern(torch.from_numpy(mesh["triangles"]), 25)
    assert pat.nnz == 137  # structural, not numeric (105): SURVEY.md 8(c)
    assert pat.crow[:7].tolist() == [0, 4, 9, 14, 19, 22, 27]


@pytest.mark.parametrize("name,mesh", list(meshes()))
@pytest.mark.parametrize("rows_per_tile,ordering", [(8, "block"), (16, "morton"), (7, "natural"), (256, "block"), (48, "auto")])
def test_tile_plan_reproduces_oracle(name, mesh, rows_per_tile, ordering):
    coords, conn = mesh["vertices"], mesh["triangles"]
    n_dof = coords.shape[0]
    tconn = torch.from_numpy(conn)
    pat = csr.build_pattern(tconn, n_dof)
    plan = csr.build_tile_plan(tconn, tconn, pat, torch.from_numpy(coords), rows_per_tile, ordering, elem_ids=(rows_per_tile == 16))
    assert plan.halo_factor >= 1.0
    geo = fo.tri_geometry(coords, conn, 3)
    local = fo.quad_reduce(fo.form_stiffness_mass(geo), geo["dx"])
    f_q = fo.source_sinsin(geo["integration_points"])
    lvec = fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]).reshape(-1, 3)
    if not np.isfinite(local).all():  # degenerate element: any symmetric local matrices exercise the plan
        rng = np.random.default_rng(0)
        local = rng.standard_normal(local.shape)
        local = local + np.swapaxes(local, -1, -2)
        lvec = rng.standard_normal(lvec.shape)
    vals, load = emulate_tiled(plan, coords, local, lvec, conn, pat.nnz, n_dof)
    crow, col, ref_vals = fo.scatter_bilinear_csr(local, conn, n_dof)
    ref_load = fo.scatter_linear(lvec, conn, n_dof).reshape(-1)
    assert not np.isnan(vals).any() and not np.isnan(load).any()
    np.testing.assert_allclose(vals, ref_vals, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(load, ref_load, rtol=1e-13, atol=1e-15)
    # every row is owned by exactly one tile
    owned = np.concatenate([plan.sections(t)["row_id"] for t in range(plan.n_tiles)])
    assert sorted(owned.tolist()) == list(range(n_dof))


def test_lattice_plan_shares_templates_and_picks_a_layout():
    """A lattice-numbered mesh is tiled in index space (independent of the jitter): congruent tiles share
    one template, and the table layout search beats the plain layout in the bank-conflict model."""
    from pytorch_fem_solver_b200 import tileplan

    mesh = meshgen.structured_rectangle(96, 64, jitter=0.25, seed=5, topology=False)
    conn = torch.from_numpy(mesh["triangles"])
    pat = csr.build_pattern(conn, mesh["vertices"].shape[0])
    assert tileplan.detect_lattice(pat) == 97
    plan = csr.build_tile_plan(conn, conn, pat, torch.from_numpy(mesh["vertices"]), 96, "auto", tile_shape=(12, 8))
    assert plan.lattice == (97, 12, 8)
    assert plan.n_templates <= 16 < plan.n_tiles  # interior + edge / corner shapes only
    populations = torch.bincount(plan.tile_desc[:, 2].long())
    assert int(populations.max()) >= (96 // 12 - 2) * (64 // 8 - 2)
    assert plan.layout_stats is not None and plan.elem_order in (1, 2)
    # an unstructured mesh is not mistaken for a lattice
    other = meshgen.delaunay_unit_square(300, seed=1)
    assert tileplan.detect_lattice(csr.build_pattern(torch.from_numpy(other["triangles"]), other["vertices"].shape[0])) == 0
