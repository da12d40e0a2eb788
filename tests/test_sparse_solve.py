"""Iterative solve on the assembled CSR system: host logic on CPU through the
oracle-backed stand-ins, the real kernels under `-m gpu`."""

import math

import numpy as np
import pytest
import scipy.sparse
import scipy.sparse.linalg
import torch

import pytorch_fem_solver_b200 as tfem
from pytorch_fem_solver_b200 import forms, meshgen, sparse
from pytorch_fem_solver_b200.basis import abstract_basis
from tests import cpu_shim


def poisson_problem(device, nx, ny, jitter=0.2):
    previous = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        with torch.device(device):
            mesh = tfem.MeshTri(meshgen.structured_rectangle(nx, ny, jitter=jitter, seed=5))
            basis = tfem.Basis(mesh, tfem.ElementTri(1, 3))
    finally:
        torch.set_default_dtype(previous)
    return basis


def check_solve(device, nx, ny, monkeypatch=None):
    """-Laplace u = 2 pi^2 sin(pi x) sin(pi y), u = 0 on the boundary: CG on the CSR system against
    scipy's direct solve of the same reduced system, and against the analytic solution."""
    basis = poisson_problem(device, nx, ny)
    if monkeypatch is not None:
        monkeypatch.setattr(abstract_basis, "DENSE_LIMIT", 16)  # force the CSR path on a small mesh
    stiffness = basis.integrate_bilinear_form(forms.Stiffness())
    load = basis.integrate_linear_form(forms.Load(forms.SinSinSource()))
    assert stiffness.layout == torch.sparse_csr
    if monkeypatch is None and stiffness.shape[0] <= abstract_basis.DENSE_LIMIT:
        pytest.skip("mesh too small for the CSR solve path")
    solution = basis.solve(stiffness, basis.solution_tensor(), load)
    info = basis.last_solve_info
    assert info.converged and info.relative_residual <= 1e-10
    # same reduced system, direct
    n = stiffness.shape[0]
    a = scipy.sparse.csr_matrix((stiffness.values().cpu().numpy(), stiffness.col_indices().cpu().numpy(),
                                 stiffness.crow_indices().cpu().numpy()), shape=(n, n))
    inner = basis._basis_parameters["inner_dofs"].cpu().numpy()
    direct = np.zeros(n)
    direct[inner] = scipy.sparse.linalg.spsolve(a[inner][:, inner].tocsc(), load.cpu().numpy().reshape(-1)[inner])
    ours = solution.cpu().numpy().reshape(-1)
    assert np.abs(ours - direct).max() <= 1e-8 * np.abs(direct).max()
    points = basis.mesh["vertices", "coordinates"].cpu().numpy().reshape(-1, 2)
    exact = np.sin(math.pi * points[:, 0]) * np.sin(math.pi * points[:, 1])
    h = 1.0 / min(nx, ny)
    assert np.abs(ours - exact).max() < 3.0 * h * h  # O(h^2) nodal error of P1
    return info


def test_csr_diagonal_and_cg_small():
    rng = np.random.default_rng(0)
    m = scipy.sparse.random(40, 40, density=0.2, random_state=1, format="csr")
    a = (m @ m.T + 40 * scipy.sparse.identity(40)).tocsr()
    a.sort_indices()
    crow, col = torch.from_numpy(a.indptr.astype(np.int32)), torch.from_numpy(a.indices.astype(np.int32))
    val = torch.from_numpy(a.data)
    assert np.allclose(sparse.csr_diagonal(crow, col, val).numpy(), a.diagonal())
    b = torch.from_numpy(rng.standard_normal(40))
    keep = torch.ones(40, dtype=torch.bool)
    keep[[0, 7, 39]] = False
    mp = pytest.MonkeyPatch()
    try:
        cpu_shim.install(mp)
        x, info = sparse.cg(crow, col, val, b, keep, rtol=1e-12, check_every=5)
    finally:
        mp.undo()
    assert info.converged
    idx = np.nonzero(keep.numpy())[0]
    ref = np.zeros(40)
    ref[idx] = np.linalg.solve(a.toarray()[np.ix_(idx, idx)], b.numpy()[idx])
    assert np.abs(x.numpy() - ref).max() < 1e-10


def test_cg_stops_on_breakdown_and_respects_limits():
    """A non-SPD system must end with `converged == False` after a bounded number of iterations (not
    10 n graph replays), a converged start vector needs no iteration, and tiny iteration limits are kept."""
    n = 30
    a = scipy.sparse.diags([np.r_[np.ones(n - 1), -1.0], 0.3 * np.ones(n - 1), 0.3 * np.ones(n - 1)], [0, 1, -1]).tocsr()
    a.sort_indices()
    crow, col = torch.from_numpy(a.indptr.astype(np.int32)), torch.from_numpy(a.indices.astype(np.int32))
    val = torch.from_numpy(a.data)
    b = torch.from_numpy(np.random.default_rng(1).standard_normal(n))
    mp = pytest.MonkeyPatch()
    try:
        cpu_shim.install(mp)
        singular = scipy.sparse.csr_matrix(np.zeros((n, n)) + np.diag(np.r_[np.ones(n - 1), 0.0]))
        zero_row = (torch.from_numpy(singular.indptr.astype(np.int32)), torch.from_numpy(singular.indices.astype(np.int32)),
                    torch.from_numpy(singular.data))
        _, info = sparse.cg(*zero_row, torch.ones(n, dtype=torch.float64), rtol=1e-12, check_every=5)
        assert not info.converged and info.iterations <= 2 * n + 100
        _, info = sparse.cg(crow, col, val, b, rtol=1e-12, check_every=5)
        assert info.iterations <= 2 * n + 100  # indefinite: either breaks down or happens to converge, never runs away
        good = (a.T @ a + scipy.sparse.identity(n)).tocsr()
        good.sort_indices()
        g = (torch.from_numpy(good.indptr.astype(np.int32)), torch.from_numpy(good.indices.astype(np.int32)), torch.from_numpy(good.data))
        x, info = sparse.cg(*g, b, rtol=1e-12, check_every=5)
        assert info.converged
        _, again = sparse.cg(*g, b, x0=x, rtol=1e-10)
        assert again.converged and again.iterations == 0
        _, short = sparse.cg(*g, b, rtol=1e-12, max_iterations=2)
        assert short.iterations <= 2 and not short.converged
    finally:
        mp.undo()


def _rvpinn_loss_check(device):
    """r^T G^-1 r and A^-1 b inside a loss: values and gradients against the dense formulas (example_weak.py:84-86,138)."""
    rng = np.random.default_rng(2)
    n = 60
    m = scipy.sparse.random(n, n, density=0.1, random_state=3, format="csr")
    a = (m @ m.T + n * scipy.sparse.identity(n)).tocsr()
    a.sort_indices()
    dense = torch.tensor(a.toarray(), device=device)
    matrix = torch.sparse_csr_tensor(torch.from_numpy(a.indptr.astype(np.int32)), torch.from_numpy(a.indices.astype(np.int32)),
                                     torch.from_numpy(a.data), size=(n, n)).to(device)
    keep = torch.ones(n, dtype=torch.uint8, device=device)
    keep[[0, 5, n - 1]] = 0
    idx = torch.nonzero(keep, as_tuple=True)[0]
    theta = torch.tensor(rng.standard_normal(n), device=device, requires_grad=True)
    r = torch.sin(theta) + 0.3 * theta  # any differentiable producer of the residual
    loss = sparse.inverse_quadratic_form(matrix, r.reshape(-1, 1), keep, rtol=1e-13)
    (g,) = torch.autograd.grad(loss, theta, retain_graph=True)
    theta_ref = theta.detach().clone().requires_grad_(True)
    r_ref = (torch.sin(theta_ref) + 0.3 * theta_ref)[idx]
    loss_ref = r_ref @ torch.linalg.solve(dense[idx][:, idx], r_ref)
    (g_ref,) = torch.autograd.grad(loss_ref, theta_ref, retain_graph=True)
    assert abs(float(loss) - float(loss_ref)) <= 1e-10 * abs(float(loss_ref))
    assert float((g - g_ref).abs().max()) <= 1e-9 * float(g_ref.abs().max())
    x = sparse.solve(matrix, r.reshape(-1, 1), keep, rtol=1e-13)
    w = torch.tensor(rng.standard_normal(n), device=device)
    (g2,) = torch.autograd.grad((x.reshape(-1) * w).sum(), theta)
    x_ref = torch.zeros(n, dtype=torch.float64, device=device)
    x_ref[idx] = torch.linalg.solve(dense[idx][:, idx], (torch.sin(theta_ref) + 0.3 * theta_ref)[idx])
    (g2_ref,) = torch.autograd.grad((x_ref * w).sum(), theta_ref)
    assert float((g2 - g2_ref).abs().max()) <= 1e-9 * float(g2_ref.abs().max())


def test_rvpinn_loss_is_differentiable_host_logic(monkeypatch):
    cpu_shim.install(monkeypatch)
    _rvpinn_loss_check("cpu")


@pytest.mark.gpu
def test_rvpinn_loss_is_differentiable_gpu():
    _rvpinn_loss_check("cuda")


def test_solve_csr_path_host_logic(monkeypatch):
    cpu_shim.install(monkeypatch)
    check_solve("cpu", 24, 20, monkeypatch)


@pytest.mark.gpu
def test_solve_small_forced_csr(monkeypatch):
    check_solve("cuda", 24, 20, monkeypatch)


@pytest.mark.gpu
def test_solve_quarter_million_unknowns():
    info = check_solve("cuda", 512, 512)
    assert info.iterations < 6000


@pytest.mark.gpu
def test_fused_cg_matches_unfused_and_is_reproducible():
    basis = poisson_problem("cuda", 96, 80)
    stiffness = basis._matrix_result(basis._assemble_fused(forms.Stiffness(), None)[0], "csr")
    load = basis.integrate_linear_form(forms.Load(forms.SinSinSource()))
    keep = torch.zeros(stiffness.shape[0], dtype=torch.uint8, device="cuda")
    keep[basis._basis_parameters["inner_dofs"]] = 1
    args = (stiffness.crow_indices().to(torch.int32), stiffness.col_indices().to(torch.int32), stiffness.values(), load, keep)
    x_fused, info = sparse.cg(*args, rtol=1e-11)
    x_again, _ = sparse.cg(*args, rtol=1e-11)
    x_plain, info_plain = sparse.cg(*args, rtol=1e-11, fused=False)
    x_eager, _ = sparse.cg(*args, rtol=1e-11, use_graph=False)
    assert info.converged and info_plain.converged
    assert torch.equal(x_fused, x_again)  # fixed-order reductions: bitwise reproducible
    assert torch.equal(x_fused, x_eager)  # graph replay == eager launches
    scale = float(x_plain.abs().max())
    assert float((x_fused - x_plain).abs().max()) <= 1e-8 * scale
    assert float(x_fused[keep == 0].abs().max()) == 0.0


@pytest.mark.gpu
def test_spmv_against_scipy_fp32_and_fp64():
    from pytorch_fem_solver_b200 import ops

    rng = np.random.default_rng(3)
    a = scipy.sparse.random(3000, 3000, density=0.003, random_state=2, format="csr")
    a.sort_indices()
    for dtype, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
        x = rng.standard_normal(3000).astype(dtype)
        keep = rng.random(3000) < 0.8
        args = [torch.from_numpy(a.indptr.astype(np.int32)).cuda(), torch.from_numpy(a.indices.astype(np.int32)).cuda(),
                torch.from_numpy(a.data.astype(dtype)).cuda(), torch.from_numpy(x).cuda()]
        y = ops.csr_spmv(*args).cpu().numpy()
        ref = a.astype(np.float64) @ x.astype(np.float64)
        assert np.abs(y - ref).max() <= tol * max(np.abs(ref).max(), 1.0)
        y = ops.csr_spmv(*args, torch.from_numpy(keep.astype(np.uint8)).cuda()).cpu().numpy()
        assert np.abs(y - ref * keep).max() <= tol * max(np.abs(ref).max(), 1.0)
