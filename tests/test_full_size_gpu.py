"""Parity at the FULL sizes of BASELINE.json's configs 2, 4 and 5, and on every branch of the tiled kernel's
source evaluation.  `pytest -m gpu`.

The checker is the CPU restatement of the reference (oracle/fem_oracle.py, oracle/torch_cpu_port.py --
both pinned to the reference's own outputs by tests/test_oracle_golden.py and tests/test_cpu_port.py);
tolerances are BASELINE.json's: 1e-12 relative to max |ref| in fp64."""

import math

import numpy as np
import pytest
import torch

import pytorch_fem_solver_b200 as tfem
from oracle import fem_oracle as fo
from pytorch_fem_solver_b200 import forms, meshgen
from tests import api_checks
from tests.test_kernels_gpu import DEV, make_basis, relmax

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config2():
    mesh = meshgen.structured_rectangle(2048, 1024, jitter=0.25, seed=1234, topology=False)
    return mesh, make_basis(mesh, 3)


def test_config2_every_row_against_the_cpu_restatement(config2):
    """Config 2 as bench.py runs it (4 194 304 elements, default tile plan): EVERY CSR value and EVERY load
    entry against the reference's tensor program on the CPU, and the tiled kernel against the two-pass
    kernels for the matrix and the load."""
    from oracle.torch_cpu_port import reference_assembly_cpu

    mesh, basis = config2
    matrix, load_ref = reference_assembly_cpu(torch.from_numpy(mesh["vertices"]), torch.from_numpy(mesh["triangles"]), 3)
    pat = basis.pattern
    assert torch.equal(pat.crow.cpu().long(), matrix.crow_indices()) and torch.equal(pat.col.cpu().long(), matrix.col_indices())
    values, load = basis.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path="tiled")
    plan = basis.tile_plan()
    assert plan.lattice is not None and plan.n_templates < 20  # the branch the benchmark runs
    assert relmax(values.cpu().numpy(), matrix.values().numpy()) < 1e-12
    assert relmax(load.cpu().numpy().reshape(-1), load_ref.numpy().reshape(-1)) < 1e-12
    values2, load2 = basis.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path="two_pass")
    assert relmax(values.cpu().numpy(), values2.cpu().numpy()) < 1e-13
    assert relmax(load.cpu().numpy(), load2.cpu().numpy()) < 1e-13


def test_config2_fp32_against_the_fp64_restatement(config2):
    """fp32 at the full config-2 size.  BASELINE.json's 1e-5 cannot hold on this mesh in ANY fp32 evaluation:
    the edge vectors x1 - x0 of elements of size h = 1/2048 carry a relative rounding error of eps32 / h
    ~ 1.2e-4, and the stiffness entries are quadratic in them.  Achieved and asserted here: 2e-3 relative
    to max |ref| on the matrix (the two-pass fp32 kernels land in the same place), 1e-4 on the load."""
    mesh, basis64 = config2
    ref_values, ref_load = basis64.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path="tiled")
    basis = make_basis(mesh, 3, torch.float32)
    values, load = basis.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path="tiled")
    values2, load2 = basis.assemble(forms.StiffnessMass(), forms.Load(), layout="values", path="two_pass")
    ref_values, ref_load = ref_values.cpu().numpy(), ref_load.cpu().numpy()
    err_tiled = relmax(values.cpu().numpy().astype(np.float64), ref_values)
    err_two_pass = relmax(values2.cpu().numpy().astype(np.float64), ref_values)
    print(f"fp32 config 2: matrix error tiled {err_tiled:.2e}, two-pass {err_two_pass:.2e}; "
          f"load error tiled {relmax(load.cpu().numpy().astype(np.float64), ref_load):.2e}")
    assert err_tiled < 2e-3 and err_two_pass < 2e-3
    assert relmax(load.cpu().numpy().astype(np.float64), ref_load) < 1e-4
    assert relmax(load2.cpu().numpy().astype(np.float64), ref_load) < 1e-4


# ---------------------------------------------------------------------------------------------------
# Source evaluation of the tiled kernel: base-point rotation small / medium / library, centroid expansion of
# degree 4 / 6 / 8 / sin() per point.  The branch depends on (frequency x tile extent) and (frequency x element
# size), so a frequency sweep on one mesh walks through all of them; the thresholds are restated below.
# ---------------------------------------------------------------------------------------------------
ROT_SMALL, ROT_MEDIUM = 0.04, 0.2
SPREAD3 = 0.4  # largest |barycentric offset| sum of the 4-point rule (assemble_tiled.cu: centroid_spread)
DEG4, DEG6, DEG8 = 4.0e-3 / SPREAD3, 3.0e-2 / SPREAD3, 1.0e-1 / SPREAD3


def _phase_extents(mesh, plan, wx, wy):
    """(largest base-vertex -> centroid phase, largest edge phase) over all (tile, element) pairs."""
    coords, conn = mesh["vertices"], mesh["triangles"]
    centroid = coords[conn].mean(axis=1)
    tile_of = plan.tile_of_row.cpu().numpy()[conn]  # tiles touching each element
    base = coords[plan.tile_desc[:, 3].cpu().numpy()]
    rot = 0.0
    for k in range(3):
        d = np.abs((centroid - base[tile_of[:, k]]) * np.array([wx, wy]))
        rot = max(rot, float(d.max()))
    edges = np.abs(np.stack([coords[conn[:, 1]] - coords[conn[:, 0]], coords[conn[:, 2]] - coords[conn[:, 0]]]) * np.array([wx, wy]))
    return rot, float(edges.max())


@pytest.mark.parametrize("rows_per_tile", [336, 12])
def test_tiled_source_branches_by_frequency(rows_per_tile):
    from pytorch_fem_solver_b200 import ops

    mesh = meshgen.structured_rectangle(64, 48, jitter=0.25, seed=9, topology=False)
    basis = make_basis(mesh, 3)
    plan = basis.tile_plan(rows_per_tile)
    pat = basis.pattern
    coords, conn = mesh["vertices"], mesh["triangles"]
    geo = fo.tri_geometry(coords, conn, 3)
    seen = set()
    for exponent in range(-9, 7):
        wx, wy = math.pi * 2.0**exponent, 0.6 * math.pi * 2.0**exponent
        rot, reach = _phase_extents(mesh, plan, wx, wy)
        seen.add(("rot", 0 if rot < ROT_SMALL else (1 if rot < ROT_MEDIUM else 2)))
        seen.add(("deg", 4 if reach < DEG4 else (6 if reach < DEG6 else (8 if reach < DEG8 else 0))))
        src = forms.SinSinSource(1.7, wx, wy)
        load = torch.full((pat.n_dof,), float("nan"), dtype=torch.float64, device=DEV)
        ops.assemble_csr_tiled(plan.c_struct(), basis._layout.coords, 3, 0.0, 0.0, src.kind, src.params, None, load)
        f_q = fo.source_sinsin(geo["integration_points"], 1.7, wx, wy)
        ref = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), conn, coords.shape[0]).reshape(-1)
        assert relmax(load.cpu().numpy(), ref) < 1e-12, (exponent, rot, reach)
    if rows_per_tile == 336:  # the sweep crosses every threshold
        assert seen == {("rot", 0), ("rot", 1), ("rot", 2), ("deg", 4), ("deg", 6), ("deg", 8), ("deg", 0)}, seen


@pytest.mark.parametrize("order", [1, 2, 4])
def test_tiled_source_other_rules_by_frequency(order):
    """The table-driven expansion (every rule but the 4-point one, which has its own code) on all branches."""
    from pytorch_fem_solver_b200 import ops

    mesh = meshgen.structured_rectangle(48, 40, jitter=0.25, seed=order, topology=False)
    basis = make_basis(mesh, order)
    plan = basis.tile_plan(96)
    pat = basis.pattern
    coords, conn = mesh["vertices"], mesh["triangles"]
    geo = fo.tri_geometry(coords, conn, order)
    for exponent in (-8, -5, -3, -1, 1, 3, 5):
        wx, wy = math.pi * 2.0**exponent, 0.6 * math.pi * 2.0**exponent
        src = forms.SinSinSource(1.7, wx, wy)
        load = torch.full((pat.n_dof,), float("nan"), dtype=torch.float64, device=DEV)
        ops.assemble_csr_tiled(plan.c_struct(), basis._layout.coords, order, 0.0, 0.0, src.kind, src.params, None, load)
        f_q = fo.source_sinsin(geo["integration_points"], 1.7, wx, wy)
        ref = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), conn, coords.shape[0]).reshape(-1)
        assert relmax(load.cpu().numpy(), ref) < 1e-12, (order, exponent)


# ---------------------------------------------------------------------------------------------------
# Config 4: two fractures of 524 288 elements each, 6-point rule, weak residual forward + adjoint
# ---------------------------------------------------------------------------------------------------
def test_config4_full_size_weak_residual_and_adjoint():
    nx, ny = 1024, 256
    meshes, data = meshgen.two_fracture_network(nx, ny)
    v2 = np.stack([m["vertices"] for m in meshes])
    conn = np.stack([m["triangles"] for m in meshes])
    with api_checks.default_device(DEV):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 4))
    assert conn.shape[0] * conn.shape[1] == 1048576
    fmap = fo.fracture_map(v2, data)
    geo = fo.tri_geometry(v2, conn, 4, fracture=fmap)
    tris = basis.global_triangulation["triangles"].cpu().numpy()
    n_g = basis.pattern.n_dof
    rng = np.random.default_rng(0)
    grad = rng.standard_normal(geo["integration_points"].shape)
    f_q = api_checks.rhs3_np(geo["integration_points"])
    ref = fo.scatter_linear(fo.quad_reduce(fo.form_weak_residual(geo, f_q, grad), geo["dx"]), tris, n_g)
    gu = torch.tensor(grad, device=DEV, requires_grad=True)
    basis.residual_path = "two_pass"  # element kernel + scatter
    r2 = basis.integrate_linear_form(forms.WeakResidual(api_checks.rhs3), gu.detach())
    assert relmax(r2.cpu().numpy().reshape(-1), ref.reshape(-1)) < 1e-12
    basis.residual_path = "auto"  # at this size: one launch of the tiled kernel
    assert basis._residual_tiled(forms.as_source(api_checks.rhs3))
    r = basis.integrate_linear_form(forms.WeakResidual(api_checks.rhs3), gu)
    assert relmax(r.detach().cpu().numpy().reshape(-1), ref.reshape(-1)) < 1e-12
    cot = rng.standard_normal(ref.shape)
    (r * torch.tensor(cot, device=DEV).reshape(r.shape)).sum().backward()
    ref_bar = fo.weak_residual_backward(geo, tris, cot)
    assert relmax(gu.grad.cpu().numpy(), ref_bar.reshape(gu.grad.shape)) < 1e-12


# ---------------------------------------------------------------------------------------------------
# Config 5: seven fractures of 1 200 128 elements each
# ---------------------------------------------------------------------------------------------------
def test_config5_full_size_seven_fractures():
    nx, ny = 1024, 586
    n_g = api_checks.check_seven_fractures(DEV, nx=nx, ny=ny)
    assert n_g == 7 * (nx + 1) * (ny + 1) - 6 * (ny + 1)
