"""Kernel-level parity on the GPU against the oracle: precisions, edge cases, determinism and
size-independent properties at the full BASELINE config-2 size.  `pytest -m gpu`."""


import numpy as np
import pytest
import torch

import pytorch_fem_solver_b200 as tfem
from oracle import fem_oracle as fo
from pytorch_fem_solver_b200 import forms, meshgen
from tests import api_checks

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_basis(mesh_dict, order=3, dtype=torch.float64):
    previous = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(DEV):
            return tfem.Basis(tfem.MeshTri(mesh_dict), tfem.ElementTri(1, order))
    finally:
        torch.set_default_dtype(previous)


def oracle_system(mesh_dict, order=3, alpha=1.0, beta=1.0):
    coords, conn = mesh_dict["vertices"], mesh_dict["triangles"]
    geo = fo.tri_geometry(coords, conn, order)
    local = fo.quad_reduce(alpha * fo.form_stiffness(geo) + beta * fo.form_mass(geo), geo["dx"])
    crow, col, vals = fo.scatter_bilinear_csr(local, conn, coords.shape[0])
    f_q = fo.source_sinsin(geo["integration_points"])
    load = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), conn, coords.shape[0]).reshape(-1)
    return crow, col, vals, load


def relmax(a, b):
    return np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("path", ["tiled", "two_pass"])
def test_fp64_matches_oracle_all_orders(order, path):
    mesh = meshgen.structured_rectangle(40, 24, jitter=0.25, seed=order, topology=False)
    basis = make_basis(mesh, order)
    values, load = basis.assemble(forms.StiffnessMass(0.7, 1.3), forms.Load(), layout="values", path=path)
    crow, col, ref_vals, ref_load = oracle_system(mesh, order, 0.7, 1.3)
    assert np.array_equal(basis.pattern.crow.cpu().numpy(), crow)  # bit-exact pattern
    assert np.array_equal(basis.pattern.col.cpu().numpy(), col)
    assert relmax(values.cpu().numpy(), ref_vals) < 1e-12  # tolerance of BASELINE.json north_star (fp64)
    assert relmax(load.cpu().numpy().reshape(-1), ref_load) < 1e-12


@pytest.mark.parametrize("path", ["tiled", "two_pass"])
def test_fp32_within_1e5_of_fp64_oracle(path):
    # coarse mesh: in fp32 the edge vectors x1-x0 carry a relative error ~eps/h, so the stated
    # 1e-5 tolerance is only meaningful while h is not small
    mesh = meshgen.structured_rectangle(12, 10, jitter=0.2, seed=3, topology=False)
    basis = make_basis(mesh, 3, torch.float32)
    values, load = basis.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path=path)
    assert values.dtype == torch.float32
    _, _, ref_vals, ref_load = oracle_system(mesh, 3)
    assert relmax(values.double().cpu().numpy(), ref_vals) < 1e-5
    assert relmax(load.double().cpu().numpy().reshape(-1), ref_load) < 1e-5


def test_generic_path_fp32():
    mesh = meshgen.delaunay_unit_square(40, seed=2)
    basis = make_basis(mesh, 4, torch.float32)
    a = basis.integrate_bilinear_form(lambda b: b.v_grad @ b.v_grad.mT + b.v @ b.v.mT, layout="values")
    _, _, ref_vals, _ = oracle_system(mesh, 4)
    assert relmax(a.double().cpu().numpy(), ref_vals) < 1e-5


def test_clockwise_elements_keep_the_signed_determinant():
    """element_tri.py:139 uses det J without abs(): clockwise cells contribute negative weights."""
    mesh = meshgen.structured_rectangle(9, 7, jitter=0.2, seed=1, topology=False)
    mesh["triangles"][::3] = mesh["triangles"][::3][:, [0, 2, 1]]
    basis = make_basis(mesh, 3)
    crow, col, ref_vals, ref_load = oracle_system(mesh, 3)
    for path in ("tiled", "two_pass"):
        values, load = basis.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path=path)
        assert relmax(values.cpu().numpy(), ref_vals) < 1e-12
        assert relmax(load.cpu().numpy().reshape(-1), ref_load) < 1e-12
    assert (ref_vals[crow[:-1]] < 0).any() or (ref_vals < 0).any()


def test_isolated_vertices_and_single_triangle():
    mesh = {
        "vertices": np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [5.0, 5.0], [6.0, 6.0]]),
        "triangles": np.array([[0, 1, 2]], dtype=np.int32),
        "vertex_markers": np.ones((5, 1), dtype=np.int32),
    }
    basis = make_basis(mesh, 3)
    crow, col, ref_vals, ref_load = oracle_system(mesh, 3)
    assert basis.pattern.crow.tolist() == [0, 3, 6, 9, 9, 9]
    for path in ("tiled", "two_pass"):
        values, load = basis.assemble(forms.Stiffness(), forms.Load(forms.ConstSource(1.0)), layout="values", path=path)
        k_loc = 0.5 * np.array([2.0, -1, -1, -1, 1, 0, -1, 0, 1])  # SURVEY.md 8(c) unit right triangle
        np.testing.assert_allclose(values.cpu().numpy(), k_loc, atol=1e-15)
        np.testing.assert_allclose(load.cpu().numpy().reshape(-1), [1 / 6, 1 / 6, 1 / 6, 0, 0], atol=1e-15)


def test_empty_inputs_are_accepted():
    from pytorch_fem_solver_b200 import ops

    coords = torch.zeros((3, 2), dtype=torch.float64, device=DEV)
    conn = torch.zeros((0, 3), dtype=torch.int32, device=DEV)
    inv, vg, xq, dx = ops.tri_geometry(coords, conn, 1, 3, 2)
    assert inv.shape == (0, 2, 2) and xq.shape == (0, 3, 2) and dx.shape == (0, 3)
    out = ops.scatter(torch.zeros(0, dtype=torch.float64, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV),
                      torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV))
    assert out.shape == (0,)


def test_unsupported_orders_raise_like_the_reference():
    with pytest.raises(NotImplementedError):
        tfem.ElementTri(1, 5)
    with pytest.raises(NotImplementedError):
        tfem.ElementLine(1, 4)
    mesh = meshgen.structured_rectangle(2, 2)
    with pytest.raises(NotImplementedError):
        make_basis_with_p2(mesh)


def make_basis_with_p2(mesh):
    with torch.device(DEV):
        return tfem.Basis(tfem.MeshTri(mesh), tfem.ElementTri(2, 2))


def test_bitwise_determinism():
    mesh = meshgen.permute_mesh(meshgen.structured_rectangle(96, 64, jitter=0.25, topology=False))
    basis = make_basis(mesh, 3)
    runs = []
    for _ in range(3):
        for path in ("tiled", "two_pass"):
            v, b = basis.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path=path)
            runs.append((path, v.clone(), b.clone()))
    for path, v, b in runs[2:]:
        first = runs[0] if path == "tiled" else runs[1]
        assert torch.equal(v, first[1]) and torch.equal(b, first[2])
    # both kernels add contributions in increasing element order, so they agree closely
    assert relmax(runs[0][1].cpu().numpy(), runs[1][1].cpu().numpy()) < 1e-14


def test_weak_residual_large_against_oracle():
    mesh = meshgen.structured_rectangle(160, 120, jitter=0.25, seed=6, topology=False)
    basis = make_basis(mesh, 4)
    coords, conn = mesh["vertices"], mesh["triangles"]
    geo = fo.tri_geometry(coords, conn, 4)
    rng = np.random.default_rng(0)
    grad = rng.standard_normal(geo["integration_points"].shape)
    f_q = fo.source_sinsin(geo["integration_points"])
    ref = fo.scatter_linear(fo.quad_reduce(fo.form_weak_residual(geo, f_q, grad), geo["dx"]), conn, coords.shape[0])
    gu = torch.tensor(grad, device=DEV, requires_grad=True)
    r = basis.integrate_linear_form(forms.WeakResidual(), gu)
    assert relmax(r.detach().cpu().numpy(), ref) < 1e-12
    cot = rng.standard_normal(ref.shape)
    (r * torch.tensor(cot, device=DEV)).sum().backward()
    ref_bar = fo.weak_residual_backward(geo, conn, cot)
    assert relmax(gu.grad.cpu().numpy(), ref_bar) < 1e-12


@pytest.mark.parametrize("order", [2, 3, 4])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_weak_residual_tiled_one_launch(order, dtype):
    """`tfem_weak_residual_tiled` (rows sum their elements' terms in the tiled kernel) against the oracle and the
    element-kernel + scatter path, with a sampled source and without one, forward and adjoint."""
    mesh = meshgen.permute_mesh(meshgen.structured_rectangle(96, 80, jitter=0.25, seed=9, topology=False))
    basis = make_basis(mesh, order, dtype)
    assert basis.dtype == dtype
    coords, conn = mesh["vertices"], mesh["triangles"]
    geo = fo.tri_geometry(coords, conn, order)
    rng = np.random.default_rng(order)
    grad = rng.standard_normal(geo["integration_points"].shape)
    tol = 1e-12 if dtype == torch.float64 else 2e-5
    for source, f_q in ((lambda x: torch.sin(3 * x[..., :1]) * torch.cos(2 * x[..., 1:2]),
                         np.sin(3 * geo["integration_points"][..., :1]) * np.cos(2 * geo["integration_points"][..., 1:2])),
                        (None, None)):
        form = forms.WeakResidual(source) if source is not None else forms.WeakResidual(forms.Source())
        ref = fo.scatter_linear(fo.quad_reduce(fo.form_weak_residual(geo, f_q if f_q is not None else 0.0 * geo["integration_points"][..., :1], grad),
                                               geo["dx"]), conn, coords.shape[0])
        results = {}
        for path in ("tiled", "two_pass"):
            basis.residual_path = path
            gu = torch.tensor(grad, device=DEV, dtype=dtype, requires_grad=True)
            r = basis.integrate_linear_form(form, gu)
            assert relmax(r.detach().cpu().numpy(), ref) < tol, (path, order)
            cot = rng.standard_normal(ref.shape)
            (r * torch.tensor(cot, device=DEV, dtype=dtype)).sum().backward()
            assert relmax(gu.grad.cpu().numpy(), fo.weak_residual_backward(geo, conn, cot)) < tol
            results[path] = r.detach()
        if dtype == torch.float64:  # both sum a row's terms in increasing element order
            assert relmax(results["tiled"].cpu().numpy(), results["two_pass"].cpu().numpy()) < 1e-14


def test_weak_residual_tiled_two_fractures():
    """The same launch on a fracture network: grad u is 3-D, pulled back with J_f^+; rows glued along the trace."""
    meshes, data = meshgen.two_fracture_network(32, 12)
    v2 = np.stack([m["vertices"] for m in meshes])
    conn = np.stack([m["triangles"] for m in meshes])
    with api_checks.default_device(DEV):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 4))
    geo = fo.tri_geometry(v2, conn, 4, fracture=fo.fracture_map(v2, data))
    tris = basis.global_triangulation["triangles"].cpu().numpy()
    rng = np.random.default_rng(4)
    grad = rng.standard_normal(geo["integration_points"].shape)
    f_q = api_checks.rhs3_np(geo["integration_points"])
    ref = fo.scatter_linear(fo.quad_reduce(fo.form_weak_residual(geo, f_q, grad), geo["dx"]), tris, basis.pattern.n_dof)
    for path in ("tiled", "two_pass"):
        basis.residual_path = path
        r = basis.integrate_linear_form(forms.WeakResidual(api_checks.rhs3), torch.tensor(grad, device=DEV))
        assert relmax(r.cpu().numpy().reshape(-1), ref.reshape(-1)) < 1e-12, path


def test_patches_config3_size_fp32():
    """BASELINE config 3 shape: 4096 patches, 6-point quadrature, fp32 weak residual."""
    centers, radius = meshgen.generate_patches_info(6)
    previous = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        with torch.device(DEV):
            patches = tfem.Patches(torch.tensor(centers, dtype=torch.float32), torch.tensor(radius, dtype=torch.float32))
            basis = tfem.PatchesBasis(patches, tfem.ElementTri(1, 4))
        coords = fo.patch_vertices(centers, radius)
        conn = np.broadcast_to(fo.PATCH_CELLS, (4096, 4, 3))
        geo = fo.tri_geometry(coords, conn, 4)
        rng = np.random.default_rng(1)
        grad = rng.standard_normal(geo["integration_points"].shape)
        f_q = fo.source_sinsin(geo["integration_points"])
        flat = (conn + 5 * np.arange(4096)[:, None, None]).reshape(-1, 3)
        ref = fo.scatter_linear(fo.quad_reduce(fo.form_weak_residual(geo, f_q, grad), geo["dx"]), flat, 5 * 4096).reshape(4096, 5, 1)
        r = basis.integrate_linear_form(forms.WeakResidual(), torch.tensor(grad, dtype=torch.float32, device=DEV))
        assert r.shape == (4096, 5, 1)
        assert relmax(r.double().cpu().numpy(), ref) < 1e-5
        assert basis.reduce(r).shape == (4096, 1)
    finally:
        torch.set_default_dtype(previous)


@pytest.fixture(scope="module")
def config2():
    mesh = meshgen.structured_rectangle(2048, 1024, jitter=0.25, seed=1234, topology=False)
    return mesh, make_basis(mesh, 3)


def test_config2_sizes(config2):
    mesh, basis = config2
    assert mesh["triangles"].shape[0] == 4194304 and mesh["vertices"].shape[0] == 2100225
    assert basis.pattern.nnz == 14689281  # N_v + 2 (N_v + N_e - 1), SURVEY.md 8(d)


def test_config2_properties(config2):
    """Size-independent invariants at the full benchmark size (the oracle takes too long here)."""
    mesh, basis = config2
    pat = basis.pattern
    rows = pat.row_indices()
    m = basis.assemble(forms.Mass(), None, layout="values", path="tiled")[0]
    assert abs(float(m.sum()) - 1.0) < 1e-12  # sum of the mass matrix = area of the unit square
    k, load = basis.assemble(forms.Stiffness(), forms.Load(), layout="values", path="tiled")
    row_sum = torch.zeros(pat.n_dof, dtype=torch.float64, device=DEV).index_add_(0, rows, k)
    diag = k[torch.searchsorted(pat.keys, torch.arange(pat.n_dof, device=DEV) * (pat.n_dof + 1))]
    assert float((row_sum.abs() / diag).max()) < 1e-11  # constants are in the kernel of the stiffness matrix
    assert abs(float(load.sum()) - 8.0) < 1e-5  # int 2 pi^2 sin(pi x) sin(pi y) = 8, up to quadrature error
    # symmetry of the assembled operator
    transposed = torch.searchsorted(pat.keys, pat.col.long() * pat.n_dof + rows)
    km = basis.assemble(forms.StiffnessMass(), None, layout="values", path="tiled")[0]
    assert float((km - km[transposed]).abs().max() / km.abs().max()) < 1e-14
    # two independent kernels agree
    km2 = basis.assemble(forms.StiffnessMass(), None, layout="values", path="two_pass")[0]
    assert float((km - km2).abs().max() / km.abs().max()) < 1e-14


def test_config2_rows_against_oracle(config2):
    """Exact oracle comparison on a band of rows of the 4M-element mesh."""
    mesh, basis = config2
    coords, conn = mesh["vertices"], mesh["triangles"]
    band = 3 * 2049  # three vertex rows: all elements touching the first two are inside
    sel = (conn < band).all(axis=1)
    sub_conn = conn[sel]
    geo = fo.tri_geometry(coords[:band], sub_conn, 3)
    local = fo.quad_reduce(fo.form_stiffness_mass(geo), geo["dx"])
    crow, col, vals = fo.scatter_bilinear_csr(local, sub_conn, band)
    f_q = fo.source_sinsin(geo["integration_points"])
    load_ref = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), sub_conn, band).reshape(-1)
    values, load = basis.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path="tiled")
    full_rows = 2 * 2049
    pat = basis.pattern
    n = int(pat.crow[full_rows])
    assert n == crow[full_rows]
    assert np.array_equal(pat.col[:n].cpu().numpy(), col[:n])
    assert relmax(values[:n].cpu().numpy(), vals[:n]) < 1e-12
    assert relmax(load[:full_rows].cpu().numpy().reshape(-1), load_ref[:full_rows]) < 1e-12


def test_geometry_config2_roundtrip(config2):
    """x_q of the geometry kernel lie inside their triangles and dx sums to the area."""
    _, basis = config2
    assert abs(float(basis._dx.sum()) - 1.0) < 1e-12
    assert float(basis._dx.sum(-3).min()) > 0.0  # the 4-point rule has one negative weight; areas are positive
    pts = basis.integration_points
    assert pts.shape == (4194304, 4, 1, 2)
    assert float(pts.min()) >= 0.0 and float(pts.max()) <= 1.0


# ------------------------------------------------------------------ differentiable interpolation
def _torch_interp_cells(basis, u):
    """Plain torch restatement (differentiable) of Basis.interpolate(self, u)."""
    conn = basis._dof_conn_flat().long()
    local = u.reshape(-1)[conn]  # (N,3)
    val = torch.einsum("ni,nqi->nq", local, basis.v.reshape(-1, basis.n_q, 3).expand(conn.shape[0], -1, -1))
    grad = torch.einsum("ni,nic->nc", local, basis.v_grad.reshape(-1, 3, 2))
    return val, grad


def test_interpolate_cells_gradient_matches_torch_autograd():
    mesh = meshgen.structured_rectangle(30, 18, jitter=0.2, seed=5, topology=False)
    basis = make_basis(mesh, 3)
    n = basis.n_dof_flat
    gen = torch.Generator(device="cpu").manual_seed(1)
    u0 = torch.randn(n, 1, dtype=torch.float64, generator=gen).to(DEV)
    wv = torch.randn(basis._layout.n_total, basis.n_q, dtype=torch.float64, generator=gen).to(DEV)
    wg = torch.randn(basis._layout.n_total, 2, dtype=torch.float64, generator=gen).to(DEV)

    u = u0.clone().requires_grad_(True)
    val, grad = basis.interpolate(basis, u)
    loss = (val.reshape(-1, basis.n_q) * wv).sum() + (grad.reshape(-1, 2) ** 2 * wg).sum()
    (g_ours,) = torch.autograd.grad(loss, u)

    u_ref = u0.clone().requires_grad_(True)
    val_r, grad_r = _torch_interp_cells(basis, u_ref)
    assert relmax(val.detach().reshape(-1).cpu().numpy(), val_r.detach().reshape(-1).cpu().numpy()) < 1e-13
    loss_r = (val_r * wv).sum() + (grad_r**2 * wg).sum()
    (g_ref,) = torch.autograd.grad(loss_r, u_ref)
    assert relmax(g_ours.cpu().numpy(), g_ref.cpu().numpy()) < 1e-12

    # deterministic adjoint: same bits twice
    u2 = u0.clone().requires_grad_(True)
    val2, grad2 = basis.interpolate(basis, u2)
    (g_again,) = torch.autograd.grad((val2.reshape(-1, basis.n_q) * wv).sum() + (grad2.reshape(-1, 2) ** 2 * wg).sum(), u2)
    assert torch.equal(g_ours, g_again)


def test_jump_estimator_gradient_matches_torch_autograd():
    """d(sum eta_E)/du through interpolate(edges) + the fused jump kernel (example_jump.py:75-87)."""
    previous = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        with torch.device(DEV):
            mesh = tfem.MeshTri(meshgen.structured_rectangle(14, 11, jitter=0.2, seed=9))
            basis = tfem.Basis(mesh, tfem.ElementTri(1, 2))
            edges = tfem.InteriorEdgesBasis(mesh, tfem.ElementLine(1, 2))
    finally:
        torch.set_default_dtype(previous)
    n = basis.n_dof_flat
    gen = torch.Generator(device="cpu").manual_seed(2)
    u0 = torch.randn(n, 1, dtype=torch.float64, generator=gen).to(DEV)
    h_e = mesh["interior_edges", "length"].unsqueeze(-2)
    n_e = mesh["interior_edges", "normals"].unsqueeze(-2)

    u = u0.clone().requires_grad_(True)
    val, grad = basis.interpolate(edges, u)
    eta = edges.integrate_functional(forms.Jump(grad), n_e, h_e)
    weights = torch.linspace(0.5, 1.5, eta.numel(), dtype=torch.float64, device=DEV).reshape(eta.shape)
    (g_ours,) = torch.autograd.grad((eta * weights).sum() + 0.1 * (val**2).sum(), u)

    # torch restatement: gather the two cells of each edge and apply the same formulas
    cells = mesh["interior_edges", "cells"].reshape(-1, 2).long()
    conn = basis._dof_conn_flat().long()
    u_ref = u0.clone().requires_grad_(True)
    local = u_ref.reshape(-1)[conn[cells]]  # (E,2,3)
    inv = basis._inv_map_jacobian.reshape(-1, 2, 2)[cells]  # (E,2,2,2)
    first = mesh["cells", "coordinates"][..., 0, :].reshape(-1, 2)[cells]  # (E,2,2)
    x_q = edges.integration_points.reshape(-1, edges.n_q, 2)
    ref_pts = torch.einsum("esqc,esrc->esqr", x_q[:, None] - first[:, :, None], inv)  # (E,2,q,2)
    lam = torch.stack([1 - ref_pts[..., 0] - ref_pts[..., 1], ref_pts[..., 0], ref_pts[..., 1]], -1)  # (E,2,q,3)
    val_r = torch.einsum("esi,esqi->esq", local, lam)
    ghat = torch.tensor([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]], dtype=torch.float64, device=DEV)
    grad_r = torch.einsum("esi,ir,esrc->esc", local, ghat, inv)
    assert relmax(val.detach().reshape(-1).cpu().numpy(), val_r.detach().reshape(-1).cpu().numpy()) < 1e-12
    assert relmax(grad.detach().reshape(-1).cpu().numpy(), grad_r.detach().reshape(-1).cpu().numpy()) < 1e-12
    normal = n_e.reshape(-1, 2)
    jump = ((grad_r[:, 0] - grad_r[:, 1]) * normal).sum(-1)
    eta_r = h_e.reshape(-1) * jump**2 * edges._dx.reshape(-1, edges.n_q).sum(-1)
    assert relmax(eta.detach().reshape(-1).cpu().numpy(), eta_r.detach().cpu().numpy()) < 1e-12
    (g_ref,) = torch.autograd.grad((eta_r * weights.reshape(-1)).sum() + 0.1 * (val_r**2).sum(), u_ref)
    assert relmax(g_ours.cpu().numpy(), g_ref.cpu().numpy()) < 1e-12


def test_progress_counter_releases_waiting_pack():
    """Device-side dependency of the multi-GPU step: tiles listed first report on a counter, and
    tfem_iface_pack_after (started EARLIER, on another stream) gathers their rows once they are done."""
    from pytorch_fem_solver_b200 import ops

    mesh = meshgen.structured_rectangle(96, 64, jitter=0.2, seed=11, topology=False)
    basis = make_basis(mesh, 3)
    pat = basis.pattern
    plan = basis.tile_plan(64)
    first = 5
    progress = torch.zeros(1, dtype=torch.int32, device=DEV)
    order = torch.arange(plan.n_tiles, device=DEV).flip(0)  # any order: the LAST tiles are listed first here
    ordered = plan.subset(order, reserve_ctas=8, n_progress_tiles=first, progress=progress)
    rows = torch.nonzero(plan.tile_of_row >= plan.n_tiles - first, as_tuple=True)[0]
    crow = pat.crow.long()
    idx = torch.cat([torch.arange(int(crow[r]), int(crow[r + 1]), device=DEV) for r in rows.tolist()] + [pat.nnz + rows]).to(torch.int32)
    buffer = torch.zeros(pat.nnz + pat.n_dof, dtype=torch.float64, device=DEV)
    values, load = buffer[: pat.nnz], buffer[pat.nnz :]
    src = forms.SinSinSource()
    side = torch.cuda.Stream()
    outs = []
    for step in (1, 2):
        buffer.zero_()
        torch.cuda.synchronize()
        out = torch.full((idx.numel(),), -1.0, dtype=torch.float64, device=DEV)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # enqueued first: it has to wait on the device
            ops.pack_after_raw(out, buffer, idx, progress, step * first * ordered.consumer_warps)
        ops.assemble_csr_tiled(ordered.c_struct(), basis._layout.coords, 3, 0.7, 1.3, src.kind, src.params, values, load)
        torch.cuda.synchronize()
        assert int(progress.item()) == step * first * ordered.consumer_warps
        assert torch.equal(out, buffer[idx.long()])
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    crow_o, col_o, ref_vals, ref_load = oracle_system(mesh, 3, 0.7, 1.3)
    assert relmax(values.cpu().numpy(), ref_vals) < 1e-12 and relmax(load.cpu().numpy(), ref_load) < 1e-12


def _irregular_meshes():
    from tests.test_csr_symbolic import fan_mesh, nonmanifold_mesh

    yield "delaunay200", meshgen.delaunay_unit_square(200, seed=4)
    yield "permuted", meshgen.permute_mesh(meshgen.structured_rectangle(24, 24, jitter=0.2, topology=False))
    yield "fan23", fan_mesh(23)  # one row with 23 elements: chained 7-element chunks
    yield "nonmanifold", nonmanifold_mesh()  # entries with 3 contributions, a repeated vertex, an isolated vertex


@pytest.mark.parametrize("name,mesh", list(_irregular_meshes()))
@pytest.mark.parametrize("rows_per_tile", [8, 192])
def test_tiled_kernel_on_irregular_meshes(name, mesh, rows_per_tile):
    """The reduction phase's rarely taken paths (heavy entries, chunk links, short segments) against the
    generic two-pass kernels and, where the elements are non-degenerate, the oracle."""
    from pytorch_fem_solver_b200 import ops

    mesh = {"vertices": np.asarray(mesh["vertices"], dtype=np.float64), "triangles": np.asarray(mesh["triangles"], dtype=np.int32),
            "vertex_markers": np.ones((len(mesh["vertices"]), 1), dtype=np.int32)}
    basis = make_basis(mesh, 3)
    pat = basis.pattern
    plan = basis.tile_plan(rows_per_tile)
    src = forms.SinSinSource()
    values = torch.full((pat.nnz,), float("nan"), dtype=torch.float64, device=DEV)
    load = torch.full((pat.n_dof,), float("nan"), dtype=torch.float64, device=DEV)
    ops.assemble_csr_tiled(plan.c_struct(), basis._layout.coords, 3, 0.7, 1.3, src.kind, src.params, values, load)
    torch.cuda.synchronize()
    ref_values, ref_load = basis.assemble(forms.StiffnessMass(0.7, 1.3), forms.Load(), layout="values", path="two_pass")
    finite = torch.isfinite(ref_values)
    assert torch.equal(torch.isfinite(values), finite)  # a degenerate element poisons the same entries in both paths
    scale = float(ref_values[finite].abs().max())
    assert float((values[finite] - ref_values[finite]).abs().max()) <= 1e-12 * scale
    finite_l = torch.isfinite(ref_load.reshape(-1))
    assert torch.equal(torch.isfinite(load), finite_l)
    assert float((load[finite_l] - ref_load.reshape(-1)[finite_l]).abs().max()) <= 1e-12 * max(float(ref_load.reshape(-1)[finite_l].abs().max()), 1e-300)
    if name != "nonmanifold":
        crow, col, o_vals, o_load = oracle_system(mesh, 3, 0.7, 1.3)
        assert np.array_equal(pat.crow.cpu().numpy(), crow) and np.array_equal(pat.col.cpu().numpy(), col)
        assert relmax(values.cpu().numpy(), o_vals) < 1e-12 and relmax(load.cpu().numpy(), o_load) < 1e-12


def test_host_pipeline_matches_direct_assembly():
    """Three streams, two steps in flight: every step's host outputs equal a plain assembly of that step's inputs."""
    mesh = meshgen.structured_rectangle(64, 48, jitter=0.2, seed=21, topology=False)
    basis = make_basis(mesh, 3)
    pat = basis.pattern
    bilinear, load = forms.StiffnessMass(0.7, 1.3), forms.Load(forms.SinSinSource())
    rng = np.random.default_rng(5)
    base = mesh["vertices"]
    inputs = [torch.from_numpy(base + 1e-3 * rng.standard_normal(base.shape) * (i > 0)).pin_memory() for i in range(5)]
    outs = [(torch.empty(pat.nnz, dtype=torch.float64).pin_memory(), torch.empty((pat.n_dof, 1), dtype=torch.float64).pin_memory()) for _ in inputs]
    pipeline = basis.host_pipeline(bilinear, load, depth=2)
    for coords, (values, vec) in zip(inputs, outs):
        pipeline.step(coords, values, vec)
    pipeline.synchronize()
    for coords, (values, vec) in zip(inputs, outs):
        ref_values, ref_vec = torch.empty_like(values), torch.empty_like(vec)
        basis.assemble_from_host(coords, bilinear, load, ref_values, ref_vec, path="tiled")
        torch.cuda.synchronize()
        assert torch.equal(values, ref_values) and torch.equal(vec, ref_vec)
    assert not torch.equal(outs[0][0], outs[1][0])  # the steps really had different inputs


def test_h1_error_functional_fused_matches_generic():
    """examples/example_weak.py:113-124: the fused kernel against the generic integrand path and the oracle."""
    import math

    mesh = meshgen.structured_rectangle(40, 30, jitter=0.2, seed=3, topology=False)
    for order in (2, 3, 4):
        basis = make_basis(mesh, order)

        def exact(p):
            return torch.sin(math.pi * p[..., 0:1]) * torch.sin(math.pi * p[..., 1:2])

        def exact_grad(p):
            return math.pi * torch.cat([torch.cos(math.pi * p[..., 0:1]) * torch.sin(math.pi * p[..., 1:2]),
                                        torch.sin(math.pi * p[..., 0:1]) * torch.cos(math.pi * p[..., 1:2])], dim=-1)

        def u(p):
            return p[..., 0:1] * (1 - p[..., 0:1]) * p[..., 1:2] ** 2

        def gradient(p):
            x, y = p[..., 0:1], p[..., 1:2]
            return torch.cat([(1 - 2 * x) * y**2, 2 * x * (1 - x) * y], dim=-1)

        form = forms.H1Error(exact, exact_grad)
        fused = basis.integrate_functional(form, u, gradient)
        generic = basis.integrate_functional(lambda b, a, g: form(b, a, g), u, gradient)
        assert fused.shape == generic.shape
        assert relmax(fused.cpu().numpy(), generic.cpu().numpy()) < 1e-12
        geo = fo.tri_geometry(mesh["vertices"], mesh["triangles"], order)
        pts = torch.from_numpy(geo["integration_points"])
        integrand = (form(type("B", (), {"integration_points": pts})(), u, gradient)).numpy()
        ref = fo.integrate_functional(integrand, geo["dx"])
        assert relmax(fused.cpu().numpy(), ref) < 1e-12


@pytest.mark.parametrize("kind", ["structured", "permuted", "delaunay", "fan"])
def test_native_symbolic_phase(kind, monkeypatch):
    """`tfem_csr_symbolic` (CUB sort / run-length encode / scans on the device) against the torch program it replaces:
    every array of the pattern bit-identical."""
    from pytorch_fem_solver_b200 import csr

    if kind == "structured":
        mesh = meshgen.structured_rectangle(97, 53, topology=False)
    elif kind == "permuted":
        mesh = meshgen.permute_mesh(meshgen.structured_rectangle(64, 40, jitter=0.2, topology=False))
    elif kind == "delaunay":
        mesh = meshgen.delaunay_unit_square(400, seed=3)
    else:  # 23 triangles around one vertex: a row with 24 entries, a DOF with 23 linear-form terms
        angles = np.linspace(0.0, 2.0 * np.pi, 24)[:-1]
        ring = np.stack([np.cos(angles), np.sin(angles)], axis=1)
        mesh = {"vertices": np.concatenate([[[0.0, 0.0]], ring]), "triangles": np.array([[0, 1 + i, 1 + (i + 1) % 23] for i in range(23)], dtype=np.int32)}
    conn = torch.tensor(np.asarray(mesh["triangles"]), dtype=torch.int32, device=DEV)
    n_dof = int(np.asarray(mesh["vertices"]).shape[0])
    native = csr.build_pattern(conn, n_dof)
    monkeypatch.setenv("TFEM_SYMBOLIC", "torch")
    ref = csr.build_pattern(conn, n_dof)
    assert native.nnz == ref.nnz
    for name in ("crow", "col", "seg", "perm", "lin_seg", "lin_perm", "keys"):
        a, b = getattr(native, name), getattr(ref, name)
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), name


@pytest.mark.parametrize("kind", ["structured", "listed_edges", "delaunay", "patches", "fractures"])
def test_native_edge_topology(kind, monkeypatch):
    """`tfem_half_edges` / `tfem_edge_cells` / `tfem_interior_edge_geometry` against the torch program they replace:
    integer outputs bit-identical, lengths and normals to rounding."""

    def build():
        with api_checks.default_device(DEV):
            if kind == "structured":
                return tfem.MeshTri(meshgen.structured_rectangle(37, 23, jitter=0.2, seed=2, topology=False))
            if kind == "listed_edges":
                return tfem.MeshTri(meshgen.structured_rectangle(16, 9, jitter=0.2, seed=3))
            if kind == "delaunay":
                return tfem.MeshTri(meshgen.delaunay_unit_square(300, seed=5))
            if kind == "patches":
                centers, radius = meshgen.generate_patches_info(3)
                return tfem.Patches(torch.tensor(centers), torch.tensor(radius))
            meshes, data = meshgen.two_fracture_network(12, 5)
            return tfem.FracturesTri(meshes, torch.tensor(data))

    native = build()
    monkeypatch.setenv("TFEM_TOPOLOGY", "torch")
    ref = build()
    for group, names in (("edges", ("vertices", "markers")), ("interior_edges", ("cells", "vertices", "coordinates", "length", "normals")),
                         ("boundary_edges", ("cells", "vertices", "coordinates")), ("cells", ("length",))):
        for name in names:
            a, b = native[group, name], ref[group, name]
            assert a.shape == b.shape and a.dtype == b.dtype and a.device == b.device, (group, name)
            if a.dtype.is_floating_point:
                assert float((a - b).abs().max()) <= 4e-16 * max(float(b.abs().max()), 1.0), (group, name)
            else:
                assert torch.equal(a, b), (group, name)


def test_reduce_keeps_csr_sparse():
    """`reduce` of a CSR operator returns the interior block in compact numbering without densifying."""
    mesh = meshgen.structured_rectangle(12, 9, jitter=0.2, seed=1)
    basis = make_basis(mesh, 3)
    dense = basis.integrate_bilinear_form(forms.StiffnessMass(), layout="dense")
    csr_matrix = basis.integrate_bilinear_form(forms.StiffnessMass(), layout="csr")
    reduced = basis.reduce(csr_matrix)
    assert reduced.layout == torch.sparse_csr
    assert torch.equal(reduced.to_dense(), basis.reduce(dense))
