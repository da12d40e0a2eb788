"""Pin the numpy oracle against outputs of the unmodified reference (tests/golden)."""

import numpy as np
import pytest

from oracle import fem_oracle as fo

SINGLE = [
    ("structured4x4", (1, 2, 3, 4)),
    ("structured6x5_jitter", (3,)),
    ("delaunay60", (2, 3, 4)),
    ("structured3x3_neighbors", (2,)),
]
TOL = dict(rtol=1e-12, atol=1e-13)


def rel(a, b):
    return np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300)


def nonsym_form(geo):
    beta = np.array([[1.0, 0.5]])
    return geo["v"] @ fo.mT(geo["v_grad"] @ beta.T) + 0.25 * geo["v_grad"] @ fo.mT(geo["v_grad"])


@pytest.mark.parametrize("name,orders", SINGLE)
def test_single_mesh_geometry_and_forms(golden, name, orders):
    g = golden(name)
    coords, conn = g["in_vertices"], g["in_triangles"]
    n_dof = coords.shape[0]
    assert np.array_equal(g["mesh_cells_coordinates"], fo.gather_cells(coords, conn))
    for order in orders:
        p = f"o{order}_"
        geo = fo.tri_geometry(coords, conn, order)
        for key in ("v", "v_grad", "integration_points", "dx", "inv_map_jacobian"):
            assert geo[key].shape == g[p + key].shape, key
            np.testing.assert_allclose(geo[key], g[p + key], **TOL)
        rows, cols, form = fo.coo_index_maps(conn)
        assert np.array_equal(rows, g[p + "idx_rows"]) and np.array_equal(cols, g[p + "idx_cols"])
        assert np.array_equal(form, g[p + "linear_idx"])
        assert np.array_equal(fo.inner_dofs(g["in_vertex_markers"]), g[p + "inner_dofs"])

        for tag, form_fn in (
            ("A_km", fo.form_stiffness_mass),
            ("A_k", fo.form_stiffness),
            ("A_m", fo.form_mass),
            ("A_nonsym", nonsym_form),
        ):
            local = fo.quad_reduce(form_fn(geo), geo["dx"])
            dense = fo.scatter_bilinear_dense(local, conn, n_dof)
            assert rel(dense, g[p + tag]) < 1e-13, tag
            crow, col, vals = fo.scatter_bilinear_csr(local, conn, n_dof)
            assert rel(fo.csr_to_dense(crow, col, vals, n_dof), g[p + tag]) < 1e-13, tag

        f_q = fo.source_sinsin(geo["integration_points"])
        b = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), conn, n_dof)
        assert rel(b, g[p + "b_load"]) < 1e-13
        assert rel(fo.integrate_functional(f_q**2, geo["dx"]), g[p + "functional"]) < 1e-13
        inner = fo.inner_dofs(g["in_vertex_markers"])
        assert rel(fo.reduce_dense(g[p + "A_km"], inner), g[p + "reduced_A_km"]) == 0
        assert rel(fo.reduce_dense(b, inner), g[p + "reduced_b"]) < 1e-13

        grad_u = g[p + "residual_grad_in"]
        r = fo.scatter_linear(fo.quad_reduce(fo.form_weak_residual(geo, f_q, grad_u), geo["dx"]), conn, n_dof)
        assert rel(r, g[p + "residual"]) < 1e-13
        g_bar = fo.weak_residual_backward(geo, conn, g[p + "residual_cotangent"])
        assert g_bar.shape == g[p + "residual_grad_bar"].shape
        assert rel(g_bar, g[p + "residual_grad_bar"]) < 1e-13

        val, grad = fo.interpolate_self(geo, conn, g[p + "interp_u"])
        assert val.shape == g[p + "interp_self"].shape and grad.shape == g[p + "interp_self_grad"].shape
        assert rel(val, g[p + "interp_self"]) < 1e-13 and rel(grad, g[p + "interp_self_grad"]) < 1e-13


def test_kat_values_of_survey(golden):
    """Known answers quoted in SURVEY.md 8(c) for the 4x4 structured mesh."""
    g = golden("structured4x4")
    coords, conn = g["in_vertices"], g["in_triangles"]
    geo = fo.tri_geometry(coords, conn, 3)
    crow, col, k = fo.scatter_bilinear_csr(fo.quad_reduce(fo.form_stiffness(geo), geo["dx"]), conn, 25)
    dense = fo.csr_to_dense(crow, col, k, 25)
    assert dense[6, 6] == pytest.approx(4.0, abs=1e-14) and dense[6, 7] == pytest.approx(-1.0, abs=1e-14)
    assert col.shape[0] == 137 and list(crow[:7]) == [0, 4, 9, 14, 19, 22, 27]
    assert np.linalg.norm(dense) == pytest.approx(15.87450786638754, rel=1e-14)
    _, _, m = fo.scatter_bilinear_csr(fo.quad_reduce(fo.form_mass(geo), geo["dx"]), conn, 25)
    assert m.sum() == pytest.approx(1.0, rel=1e-14)


@pytest.mark.parametrize("name", [n for n, _ in SINGLE])
def test_interior_edges(golden, name):
    g = golden(name)
    coords, conn = g["in_vertices"], g["in_triangles"]
    edge_coords = g["mesh_interior_edges_coordinates"]
    geo_e = fo.edge_geometry(edge_coords, 2)
    for key in ("v", "v_grad", "integration_points", "dx", "inv_map_jacobian"):
        assert geo_e[key].shape == g["e2_" + key].shape, key
        np.testing.assert_allclose(geo_e[key], g["e2_" + key], **TOL)
    length, normals = fo.interior_edge_normals(edge_coords, g["mesh_cells_coordinates"], g["mesh_interior_edges_cells"])
    np.testing.assert_allclose(length, g["mesh_interior_edges_length"], **TOL)
    np.testing.assert_allclose(normals, g["mesh_interior_edges_normals"], **TOL)

    geo = fo.tri_geometry(coords, conn, 2)
    cells = g["mesh_cells_coordinates"]
    val, grad = fo.interpolate_edges(
        geo_e["integration_points"],
        g["mesh_interior_edges_cells"],
        conn,
        cells[..., [0], :],
        geo["inv_map_jacobian"],
        g["e2_interp_u"],
    )
    assert val.shape == g["e2_interp_edges"].shape and grad.shape == g["e2_interp_edges_grad"].shape
    np.testing.assert_allclose(val, g["e2_interp_edges"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(grad, g["e2_interp_edges_grad"], rtol=1e-11, atol=1e-12)
    h_e = g["mesh_interior_edges_length"][..., None, :, :]
    n_e = g["mesh_interior_edges_normals"][..., None, :, :]
    eta = fo.integrate_functional(fo.jump_integrand(grad, n_e, h_e), geo_e["dx"])
    assert eta.shape == g["e2_eta"].shape
    np.testing.assert_allclose(eta, g["e2_eta"], rtol=1e-10, atol=1e-12)
    val_c, grad_c = fo.interpolate_edges(
        geo_e["integration_points"],
        g["mesh_interior_edges_cells"],
        conn,
        cells[..., [0], :],
        geo["inv_map_jacobian"],
        g["e2_closure_nodal"],
    )
    np.testing.assert_allclose(val_c, g["e2_closure_edges"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(grad_c, g["e2_closure_edges_grad"], rtol=1e-11, atol=1e-12)


def test_patches(golden):
    g = golden("patches_l2")
    coords = fo.patch_vertices(g["in_centers"], g["in_radius"])
    assert np.array_equal(coords, g["mesh_vertices_coordinates"])
    n_p = coords.shape[0]
    conn = np.broadcast_to(fo.PATCH_CELLS, (n_p, 4, 3))
    assert np.array_equal(conn, g["mesh_cells_vertices"])
    flat_conn = (conn + 5 * np.arange(n_p)[:, None, None]).reshape(-1, 3)
    for order in (2, 4):
        p = f"o{order}_"
        geo = fo.tri_geometry(coords, conn, order)
        for key in ("v", "v_grad", "integration_points", "dx", "inv_map_jacobian"):
            assert geo[key].shape == g[p + key].shape, key
            np.testing.assert_allclose(geo[key], g[p + key], **TOL)
        local = fo.quad_reduce(fo.form_stiffness_mass(geo), geo["dx"])
        dense = fo.scatter_bilinear_dense(local.reshape(-1, 3, 3), flat_conn, 5 * n_p)
        blocks = np.stack([dense[5 * i : 5 * i + 5, 5 * i : 5 * i + 5] for i in range(n_p)])
        assert rel(blocks, g[p + "A_km"]) < 1e-13
        inner = g[p + "inner_dofs"]
        assert rel(fo.reduce_patches(blocks, inner), g[p + "reduced_A_km"]) < 1e-13
        f_q = fo.source_sinsin(geo["integration_points"])
        b = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), flat_conn, 5 * n_p).reshape(n_p, 5, 1)
        assert rel(b, g[p + "b_load"]) < 1e-13
        assert rel(fo.reduce_patches(b, inner), g[p + "reduced_b"]) < 1e-13
        assert rel(fo.integrate_functional(f_q**2, geo["dx"]), g[p + "functional"]) < 1e-13
        grad_u = g[p + "residual_grad_in"]
        r = fo.scatter_linear(
            fo.quad_reduce(fo.form_weak_residual(geo, f_q, grad_u), geo["dx"]), flat_conn, 5 * n_p
        ).reshape(n_p, 5, 1)
        assert rel(r, g[p + "residual"]) < 1e-13
        g_bar = fo.weak_residual_backward(geo, flat_conn.reshape(n_p, 4, 3), g[p + "residual_cotangent"])
        assert rel(g_bar, g[p + "residual_grad_bar"]) < 1e-13


def rhs3(points):
    x, y, z = points[..., [0]], points[..., [1]], points[..., [2]]
    return 6.0 * (y - y**2) * np.abs(x) - 2.0 * (np.abs(z) ** 3 - np.abs(x)) + 1.0


@pytest.mark.parametrize("name", ["fractures2_4x2", "fractures2_8x4"])
def test_fractures(golden, name):
    g = golden(name)
    v2, conn = g["in_vertices"], g["in_triangles"]
    fmap = fo.fracture_map(v2, g["in_fractures_3d_data"])
    for key, gk in (("jac", "jacobian_fracture_map"), ("inv", "inv_jacobian_fracture_map"),
                    ("det", "det_jacobian_fracture_map"), ("t", "translation_vector")):
        assert fmap[key].shape == g["mesh_" + gk].shape
        np.testing.assert_allclose(fmap[key], g["mesh_" + gk], rtol=1e-12, atol=1e-14)
    # use the reference's own maps downstream so unrelated rounding does not blur later checks
    fmap = {"jac": g["mesh_jacobian_fracture_map"], "inv": g["mesh_inv_jacobian_fracture_map"],
            "det": g["mesh_det_jacobian_fracture_map"], "t": g["mesh_translation_vector"]}
    v3 = fo.fracture_vertices_3d(v2, fmap)
    assert np.array_equal(v3, g["mesh_vertices_coordinates_3d"])
    n3 = fo.fracture_normals_3d(g["mesh_interior_edges_normals"], fmap)
    np.testing.assert_allclose(n3, g["mesh_interior_edges_normals_3d"], **TOL)

    gt = fo.global_triangulation(v3, v2, conn, g["in_edges"], g["in_vertex_markers"], g["in_edge_markers"])
    for key in ("vertices_3D", "vertices_2D", "vertex_markers", "triangles", "edges", "edge_markers",
                "global2local_idx", "local2global_idx", "traces__global_vertices_idx", "traces_global_edges_idx"):
        assert np.array_equal(gt[key], g["gt_" + key]), key
    assert np.array_equal(np.stack(gt["traces_local_edges_idx"]), g["gt_traces_local_edges_idx"])
    tris = gt["triangles"]
    n_g = gt["vertices_3D"].shape[0]
    gconn = tris.reshape(conn.shape)

    for order in (2, 4):
        p = f"o{order}_"
        geo = fo.tri_geometry(v2, conn, order, fracture=fmap)
        for key in ("v", "v_grad", "integration_points", "dx", "inv_map_jacobian"):
            assert geo[key].shape == g[p + key].shape, key
            np.testing.assert_allclose(geo[key], g[p + key], **TOL)
        rows, cols, form = fo.coo_index_maps(tris)
        assert np.array_equal(rows, g[p + "idx_rows"]) and np.array_equal(cols, g[p + "idx_cols"])
        assert np.array_equal(g[p + "inner_dofs"], np.nonzero(gt["vertex_markers"] != 1)[0])
        for tag, form_fn in (("A_k", fo.form_stiffness), ("A_km", fo.form_stiffness_mass)):
            local = fo.quad_reduce(form_fn(geo), geo["dx"]).reshape(-1, 3, 3)
            assert rel(fo.scatter_bilinear_dense(local, tris, n_g), g[p + tag]) < 1e-13
            crow, col, vals = fo.scatter_bilinear_csr(local, tris, n_g)
            assert rel(fo.csr_to_dense(crow, col, vals, n_g), g[p + tag]) < 1e-13
        f_q = rhs3(geo["integration_points"])
        b = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), tris, n_g)
        assert rel(b, g[p + "b_load"]) < 1e-13
        fun = fo.integrate_functional(f_q**2, geo["dx"])
        assert fun.shape == g[p + "functional"].shape and rel(fun, g[p + "functional"]) < 1e-13
        grad_u = g[p + "residual_grad_in"]
        r = fo.scatter_linear(fo.quad_reduce(fo.form_weak_residual(geo, f_q, grad_u), geo["dx"]), tris, n_g)
        assert rel(r, g[p + "residual"]) < 1e-13
        g_bar = fo.weak_residual_backward(geo, gconn, g[p + "residual_cotangent"])
        assert g_bar.shape == g[p + "residual_grad_bar"].shape and rel(g_bar, g[p + "residual_grad_bar"]) < 1e-13
        val, grad = fo.interpolate_self(geo, gconn, g[p + "interp_u"])
        assert val.shape == g[p + "interp_self"].shape and grad.shape == g[p + "interp_self_grad"].shape
        assert rel(val, g[p + "interp_self"]) < 1e-13 and rel(grad, g[p + "interp_self_grad"]) < 1e-13

    geo_e = fo.edge_geometry(g["mesh_interior_edges_coordinates"], 2, fracture=fmap)
    for key in ("v", "v_grad", "integration_points", "dx", "inv_map_jacobian"):
        assert geo_e[key].shape == g["e2_" + key].shape, key
        np.testing.assert_allclose(geo_e[key], g["e2_" + key], **TOL)
    geo = fo.tri_geometry(v2, conn, 2, fracture=fmap)
    val, grad = fo.interpolate_edges(
        geo_e["integration_points"], g["mesh_interior_edges_cells"], conn,
        g["mesh_cells_coordinates_3d"][..., [0], :], geo["inv_map_jacobian"], g["e2_interp_u"],
    )
    assert val.shape == g["e2_interp_edges"].shape and grad.shape == g["e2_interp_edges_grad"].shape
    np.testing.assert_allclose(val, g["e2_interp_edges"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(grad, g["e2_interp_edges_grad"], rtol=1e-11, atol=1e-12)
    h_e = g["mesh_interior_edges_length"][..., None, :, :]
    n_e = g["mesh_interior_edges_normals_3d"][..., None, :, :]
    eta = fo.integrate_functional(fo.jump_integrand(grad, n_e, h_e), geo_e["dx"])
    assert eta.shape == g["e2_eta"].shape
    np.testing.assert_allclose(eta, g["e2_eta"], rtol=1e-10, atol=1e-12)
