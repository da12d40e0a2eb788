"""Parity checks of the torch_fem-compatible API against the reference's golden outputs.

Shared by `test_api_host_logic.py` (CPU, ops routed to the oracle: checks the host-side glue)
and `test_api_gpu.py` (`-m gpu`, real CUDA kernels through the C ABI).  Tolerances: fp64 values
within 1e-12 relative (BASELINE.json north_star); integer outputs bit-exact.
"""

from __future__ import annotations

import contextlib
import math

import numpy as np
import torch

import pytorch_fem_solver_b200 as tfem
from pytorch_fem_solver_b200 import forms

RTOL = 1e-12


@contextlib.contextmanager
def default_device(device):
    previous = torch.get_default_device()
    previous_dtype = torch.get_default_dtype()
    torch.set_default_device(device)
    torch.set_default_dtype(torch.float64)
    try:
        yield
    finally:
        torch.set_default_device(previous)
        torch.set_default_dtype(previous_dtype)


def npy(t):
    return t.detach().cpu().numpy()


def close(actual, expected, rtol=RTOL, what=""):
    actual = npy(actual) if isinstance(actual, torch.Tensor) else np.asarray(actual)
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} != {expected.shape}"
    scale = max(np.abs(expected).max(), 1e-300) if expected.size else 1.0
    err = np.abs(actual - expected).max() / scale if expected.size else 0.0
    assert err <= rtol, f"{what}: max error {err:.3e} relative to max |ref| > {rtol}"


def equal(actual, expected, what=""):
    actual = npy(actual) if isinstance(actual, torch.Tensor) else np.asarray(actual)
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} != {expected.shape}"
    assert np.array_equal(actual, expected), what


def rhs(x, y):
    return 2.0 * math.pi**2 * torch.sin(math.pi * x) * torch.sin(math.pi * y)


def bilinear_km(basis):
    return basis.v_grad @ basis.v_grad.mT + basis.v @ basis.v.mT


def bilinear_k(basis):
    return basis.v_grad @ basis.v_grad.mT


def bilinear_m(basis):
    return basis.v @ basis.v.mT


def bilinear_nonsym(basis):
    beta = torch.tensor([[1.0, 0.5]])
    return basis.v @ (basis.v_grad @ beta.mT).mT + 0.25 * basis.v_grad @ basis.v_grad.mT


def load(basis):
    x, y = torch.split(basis.integration_points, 1, dim=-1)
    return rhs(x, y) * basis.v


def functional(basis):
    x, y = torch.split(basis.integration_points, 1, dim=-1)
    return rhs(x, y) ** 2


def residual(basis, gradient_values):
    x, y = torch.split(basis.integration_points, 1, dim=-1)
    return rhs(x, y) * basis.v - basis.v_grad @ gradient_values.mT


def mesh_dict(g, with_neighbors=False):
    keys = ["vertices", "triangles", "vertex_markers", "edges", "edge_markers"] + (["neighbors"] if with_neighbors else [])
    return {k: g["in_" + k] for k in keys if "in_" + k in g}


def check_basis_attributes(basis, g, p):
    close(basis.v, g[p + "v"], what="v")
    close(basis.v_grad, g[p + "v_grad"], what="v_grad")
    close(basis.integration_points, g[p + "integration_points"], what="integration_points")
    close(basis._dx, g[p + "dx"], what="dx")
    close(basis._inv_map_jacobian, g[p + "inv_map_jacobian"], what="inv_map_jacobian")
    idx = basis._basis_parameters["bilinear_form_idx"]
    names = ["rows", "cols"] if len(idx) == 2 else ["patch", "rows", "cols"]
    for k, name in enumerate(names):
        equal(idx[k], g[p + "idx_" + name], what="bilinear_form_idx " + name)
    equal(basis._basis_parameters["linear_form_idx"][-1], g[p + "linear_idx"], what="linear_form_idx")
    equal(basis._basis_parameters["inner_dofs"], g[p + "inner_dofs"], what="inner_dofs")


def check_single_mesh(g, orders, device, aligned_edges=True):
    with default_device(device):
        mesh = tfem.MeshTri(mesh_dict(g))
        close(mesh["cells", "coordinates"], g["mesh_cells_coordinates"], rtol=0, what="cells.coordinates")
        check_mesh_topology(mesh, g, aligned_edges)
        for order in orders:
            p = f"o{order}_"
            basis = tfem.Basis(mesh, tfem.ElementTri(1, order))
            check_basis_attributes(basis, g, p)
            # generic path: arbitrary user callables, dense result like the reference
            for tag, fn in (("A_km", bilinear_km), ("A_k", bilinear_k), ("A_m", bilinear_m), ("A_nonsym", bilinear_nonsym)):
                close(basis.integrate_bilinear_form(fn), g[p + tag], what=tag)
            close(basis.integrate_linear_form(load), g[p + "b_load"], what="b_load")
            close(basis.integrate_functional(functional), g[p + "functional"], what="functional")
            a = basis.integrate_bilinear_form(bilinear_km)
            b = basis.integrate_linear_form(load)
            close(basis.reduce(a), g[p + "reduced_A_km"], what="reduce(A)")
            close(basis.reduce(b), g[p + "reduced_b"], what="reduce(b)")
            # CSR layout: pattern is the coalesced index map, values rebuild the dense matrix
            csr = basis.integrate_bilinear_form(bilinear_nonsym, layout="csr")
            close(csr.to_dense(), g[p + "A_nonsym"], what="csr nonsym")
            # fused named forms, both kernels
            for path in ("two_pass", "tiled"):
                for tag, form in (("A_km", forms.StiffnessMass()), ("A_k", forms.Stiffness()), ("A_m", forms.Mass())):
                    mat, vec = basis.assemble(form, forms.Load(forms.SinSinSource()), layout="dense", path=path)
                    close(mat, g[p + tag], what=f"fused {path} {tag}")
                    close(vec, g[p + "b_load"], what=f"fused {path} load")
            close(basis.integrate_bilinear_form(forms.StiffnessMass()), g[p + "A_km"], what="fused via integrate")
            close(basis.integrate_linear_form(forms.Load()), g[p + "b_load"], what="fused load via integrate")
            sampled = forms.Load(lambda pts: rhs(pts[..., 0:1], pts[..., 1:2]))
            close(basis.assemble(None, sampled, path="two_pass")[1], g[p + "b_load"], what="sampled load")

            # weak residual: generic callable, fused form, and the adjoint
            grad_in = torch.tensor(g[p + "residual_grad_in"])
            close(basis.integrate_linear_form(residual, grad_in), g[p + "residual"], what="residual generic")
            gu = grad_in.clone().requires_grad_(True)
            r = basis.integrate_linear_form(forms.WeakResidual(), gu)
            close(r, g[p + "residual"], what="residual fused")
            (r * torch.tensor(g[p + "residual_cotangent"])).sum().backward()
            close(gu.grad, g[p + "residual_grad_bar"], what="residual adjoint")
            gu2 = grad_in.clone().requires_grad_(True)
            r2 = basis.integrate_linear_form(residual, gu2)
            (r2 * torch.tensor(g[p + "residual_cotangent"])).sum().backward()
            close(gu2.grad, g[p + "residual_grad_bar"], what="residual adjoint (generic path)")

            u = torch.tensor(g[p + "interp_u"])
            val, grad = basis.interpolate(basis, u)
            close(val, g[p + "interp_self"], what="interpolate self")
            close(grad, g[p + "interp_self_grad"], what="interpolate self grad")
        check_edges(mesh, g, device)


def check_mesh_topology(mesh, g, aligned_edges=True):
    equal(mesh["interior_edges", "vertices"], g["mesh_interior_edges_vertices"], what="interior vertices")
    equal(mesh["boundary_edges", "vertices"], g["mesh_boundary_edges_vertices"], what="boundary vertices")
    close(mesh["interior_edges", "coordinates"], g["mesh_interior_edges_coordinates"], rtol=0, what="interior coordinates")
    close(mesh["interior_edges", "length"], g["mesh_interior_edges_length"], what="interior length")
    close(mesh["cells", "length"], g["mesh_cells_length"], what="cells length")
    if aligned_edges:
        equal(mesh["interior_edges", "cells"], g["mesh_interior_edges_cells"], what="interior cells")
        equal(mesh["boundary_edges", "cells"], g["mesh_boundary_edges_cells"], what="boundary cells")
        close(mesh["interior_edges", "normals"], g["mesh_interior_edges_normals"], what="normals")
    else:
        # with `neighbors` the reference lists cell pairs in sorted order, not aligned with the
        # edge list; the same SET of pairs must come out
        ours = {tuple(r) for r in npy(mesh["interior_edges", "cells"]).reshape(-1, 2).tolist()}
        theirs = {tuple(r) for r in g["mesh_interior_edges_cells"].reshape(-1, 2).tolist()}
        assert ours == theirs


def check_edges(mesh, g, device):
    p = "e2_"
    edges_basis = tfem.InteriorEdgesBasis(mesh, tfem.ElementLine(1, 2))
    close(edges_basis.v, g[p + "v"], what="edge v")
    close(edges_basis.v_grad, g[p + "v_grad"], what="edge v_grad")
    close(edges_basis.integration_points, g[p + "integration_points"], what="edge points")
    close(edges_basis._dx, g[p + "dx"], what="edge dx")
    close(edges_basis._inv_map_jacobian, g[p + "inv_map_jacobian"], what="edge inv")
    if "mesh_interior_edges_cells" in g and not np.array_equal(
        npy(mesh["interior_edges", "cells"]), g["mesh_interior_edges_cells"]
    ):
        return  # misaligned reference listing: downstream tensors are ordered differently
    basis = tfem.Basis(mesh, tfem.ElementTri(1, 2))
    u = torch.tensor(g[p + "interp_u"])
    val, grad = basis.interpolate(edges_basis, u)
    close(val, g[p + "interp_edges"], rtol=1e-12, what="interp edges")
    close(grad, g[p + "interp_edges_grad"], rtol=1e-12, what="interp edges grad")
    h_e = mesh["interior_edges", "length"].unsqueeze(-2)
    n_e = mesh["interior_edges", "normals"].unsqueeze(-2)

    def jump(_, normal, size):
        plus, minus = torch.unbind(grad, dim=-4)
        return size * ((plus * normal).sum(-1, keepdim=True) + (minus * -normal).sum(-1, keepdim=True)) ** 2

    close(edges_basis.integrate_functional(jump, n_e, h_e), g[p + "eta"], rtol=1e-12, what="eta generic")
    close(edges_basis.integrate_functional(forms.Jump(grad), n_e, h_e), g[p + "eta"], rtol=1e-12, what="eta fused")

    interp, interp_grad = basis.interpolate(edges_basis)

    def nodal(nodes):
        return torch.sin(2.0 * nodes[..., [0]]) * torch.cos(nodes[..., [1]])

    close(interp(nodal), g[p + "closure_edges"], rtol=1e-12, what="closure edges")
    close(interp_grad(nodal), g[p + "closure_edges_grad"], rtol=1e-12, what="closure edges grad")


def check_patches(g, device):
    with default_device(device):
        patches = tfem.Patches(torch.tensor(g["in_centers"]), torch.tensor(g["in_radius"]))
        close(patches["vertices", "coordinates"], g["mesh_vertices_coordinates"], rtol=0, what="patch vertices")
        equal(patches["cells", "vertices"], g["mesh_cells_vertices"], what="patch cells")
        close(patches["cells", "coordinates"], g["mesh_cells_coordinates"], rtol=0, what="patch cell coords")
        for order in (2, 4):
            p = f"o{order}_"
            basis = tfem.PatchesBasis(patches, tfem.ElementTri(1, order))
            check_basis_attributes(basis, g, p)
            a = basis.integrate_bilinear_form(bilinear_km)
            b = basis.integrate_linear_form(load)
            close(a, g[p + "A_km"], what="patch A_km")
            close(basis.integrate_bilinear_form(bilinear_k), g[p + "A_k"], what="patch A_k")
            close(b, g[p + "b_load"], what="patch load")
            close(basis.reduce(a), g[p + "reduced_A_km"], what="patch reduce A")
            close(basis.reduce(b), g[p + "reduced_b"], what="patch reduce b")
            close(basis.integrate_functional(functional), g[p + "functional"], what="patch functional")
            close(basis.integrate_bilinear_form(forms.StiffnessMass()), g[p + "A_km"], what="patch fused A_km")
            close(basis.integrate_linear_form(forms.Load()), g[p + "b_load"], what="patch fused load")
            grad_in = torch.tensor(g[p + "residual_grad_in"])
            close(basis.integrate_linear_form(residual, grad_in), g[p + "residual"], what="patch residual generic")
            gu = grad_in.clone().requires_grad_(True)
            r = basis.integrate_linear_form(forms.WeakResidual(), gu)
            close(r, g[p + "residual"], what="patch residual fused")
            close(basis.reduce(r.detach()), g[p + "reduced_residual"], what="patch reduced residual")
            (r * torch.tensor(g[p + "residual_cotangent"])).sum().backward()
            close(gu.grad, g[p + "residual_grad_bar"], what="patch residual adjoint")


def rhs3(points):
    x, y, z = torch.split(points, 1, dim=-1)
    return 6.0 * (y - y**2) * torch.abs(x) - 2.0 * (torch.abs(z) ** 3 - torch.abs(x)) + 1.0


def check_fractures(g, device):
    with default_device(device):
        n_f = g["in_vertices"].shape[0]
        keys = ["vertices", "triangles", "vertex_markers", "edges", "edge_markers", "neighbors"]
        meshes = [{k: g["in_" + k][f] for k in keys if "in_" + k in g} for f in range(n_f)]
        mesh = tfem.FracturesTri(meshes, torch.tensor(g["in_fractures_3d_data"]))
        for key in ("jacobian_fracture_map", "inv_jacobian_fracture_map", "det_jacobian_fracture_map", "translation_vector"):
            close(mesh[key], g["mesh_" + key], rtol=1e-12, what=key)
        close(mesh["vertices", "coordinates_3d"], g["mesh_vertices_coordinates_3d"], rtol=1e-15, what="coordinates_3d")
        close(mesh["cells", "coordinates_3d"], g["mesh_cells_coordinates_3d"], rtol=1e-15, what="cells coordinates_3d")
        equal(mesh["interior_edges", "vertices"], g["mesh_interior_edges_vertices"], what="interior vertices")
        close(mesh["interior_edges", "length"], g["mesh_interior_edges_length"], what="interior length")
        aligned = np.array_equal(npy(mesh["interior_edges", "cells"]), g["mesh_interior_edges_cells"])

        for order in (2, 4):
            p = f"o{order}_"
            basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, order))
            if order == 2:
                gt = basis.global_triangulation
                for key in ("vertices_3D", "vertices_2D"):
                    close(gt[key], g["gt_" + key], rtol=0, what=key)
                for key in ("vertex_markers", "triangles", "edges", "edge_markers", "global2local_idx", "local2global_idx",
                            "traces__global_vertices_idx", "traces_global_edges_idx", "traces_local_edges_idx"):
                    equal(gt[key], g["gt_" + key], what=key)
            check_basis_attributes(basis, g, p)
            close(basis.integrate_bilinear_form(bilinear_k), g[p + "A_k"], what="frac A_k")
            close(basis.integrate_bilinear_form(bilinear_km), g[p + "A_km"], what="frac A_km")
            close(basis.integrate_bilinear_form(forms.Stiffness()), g[p + "A_k"], what="frac fused A_k")
            close(basis.integrate_bilinear_form(forms.StiffnessMass()), g[p + "A_km"], what="frac fused A_km")

            def load3(b):
                return rhs3(b.integration_points) * b.v

            close(basis.integrate_linear_form(load3), g[p + "b_load"], what="frac load")
            close(basis.integrate_linear_form(forms.Load(rhs3)), g[p + "b_load"], what="frac sampled load")
            close(basis.integrate_functional(lambda b: rhs3(b.integration_points) ** 2), g[p + "functional"], what="frac functional")
            grad_in = torch.tensor(g[p + "residual_grad_in"])

            def residual3(b, gradient_values):
                return rhs3(b.integration_points) * b.v - b.v_grad @ gradient_values.mT

            close(basis.integrate_linear_form(residual3, grad_in), g[p + "residual"], what="frac residual generic")
            gu = grad_in.clone().requires_grad_(True)
            r = basis.integrate_linear_form(forms.WeakResidual(rhs3), gu)
            close(r, g[p + "residual"], what="frac residual fused")
            (r * torch.tensor(g[p + "residual_cotangent"])).sum().backward()
            close(gu.grad, g[p + "residual_grad_bar"], what="frac residual adjoint")
            u = torch.tensor(g[p + "interp_u"])
            val, grad = basis.interpolate(basis, u)
            close(val, g[p + "interp_self"], what="frac interpolate self")
            close(grad, g[p + "interp_self_grad"], what="frac interpolate self grad")

        p = "e2_"
        edges_basis = tfem.InteriorEdgesFractureBasis(mesh, tfem.ElementLine(1, 2))
        close(edges_basis.v_grad, g[p + "v_grad"], what="frac edge v_grad")
        close(edges_basis.integration_points, g[p + "integration_points"], what="frac edge points")
        close(edges_basis._dx, g[p + "dx"], what="frac edge dx")
        close(edges_basis._inv_map_jacobian, g[p + "inv_map_jacobian"], what="frac edge inv")
        close(mesh["interior_edges", "normals_3d"], g["mesh_interior_edges_normals_3d"], what="normals_3d") if aligned else None
        if aligned:
            basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 2))
            u = torch.tensor(g[p + "interp_u"])
            val, grad = basis.interpolate(edges_basis, u)
            close(val, g[p + "interp_edges"], rtol=1e-12, what="frac interp edges")
            close(grad, g[p + "interp_edges_grad"], rtol=1e-12, what="frac interp edges grad")
            h_e = mesh["interior_edges", "length"].unsqueeze(-2)
            n_e = mesh["interior_edges", "normals_3d"].unsqueeze(-2)
            close(edges_basis.integrate_functional(forms.Jump(grad), n_e, h_e), g[p + "eta"], rtol=1e-12, what="frac eta")
        return aligned


def check_seven_fractures(device, nx=8, ny=4, order=3):
    """BASELINE config 5 (scaled): backbone + six crossing planes.  The reference cannot represent this
    network (fracture_basis.py:84-92 reshapes the trace edges to (F,-1)), so parity is against the
    oracle's restatement of the same algorithm plus properties of the assembled operators."""
    from oracle import fem_oracle as fo
    from pytorch_fem_solver_b200 import meshgen

    meshes, data = meshgen.seven_fracture_network(nx, ny)
    v2 = np.stack([m["vertices"] for m in meshes])
    conn = np.stack([m["triangles"] for m in meshes])
    n_v = (nx + 1) * (ny + 1)
    n_g = 7 * n_v - 6 * (ny + 1)  # every crossing plane shares one column of ny+1 vertices with the backbone
    with default_device(device):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, order))
        gt = basis.global_triangulation
        assert gt["vertices_3D"].shape[-2] == n_g
        # oracle on the same inputs
        fmap = fo.fracture_map(v2, data)
        v3 = fo.fracture_vertices_3d(v2, fmap)
        close(mesh["vertices", "coordinates_3d"], v3, rtol=1e-15, what="coordinates_3d")
        o_gt = fo.global_triangulation(v3, v2, conn, np.stack([m["edges"] for m in meshes]),
                                       np.stack([m["vertex_markers"] for m in meshes]), np.stack([m["edge_markers"] for m in meshes]))
        for key in ("triangles", "global2local_idx", "local2global_idx", "vertex_markers"):
            equal(gt[key], o_gt[key], what="7 fractures " + key)
        tris = o_gt["triangles"]
        geo = fo.tri_geometry(v2, conn, order, fracture=fmap)
        for form, fn in ((forms.Stiffness(), fo.form_stiffness), (forms.StiffnessMass(), fo.form_stiffness_mass)):
            local = fo.quad_reduce(fn(geo), geo["dx"]).reshape(-1, 3, 3)
            crow, col, vals = fo.scatter_bilinear_csr(local, tris, n_g)
            ours = basis.integrate_bilinear_form(form)
            if ours.layout == torch.sparse_csr:
                equal(ours.crow_indices(), crow, what="7 fractures crow")
                equal(ours.col_indices(), col, what="7 fractures col")
                close(ours.values(), vals, what="7 fractures values")
                dense_rows = None
            else:
                close(ours, fo.csr_to_dense(crow, col, vals, n_g), what="7 fractures matrix")
        # stiffness annihilates constants; a trace vertex collects elements of two planes
        k_csr = fo.scatter_bilinear_csr(fo.quad_reduce(fo.form_stiffness(geo), geo["dx"]).reshape(-1, 3, 3), tris, n_g)
        row_sums = np.add.reduceat(k_csr[2], k_csr[0][:-1])
        assert np.abs(row_sums).max() < 1e-12 * np.abs(k_csr[2]).max()
        f_q = rhs3_np(geo["integration_points"])
        b = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]), tris, n_g)
        close(basis.integrate_linear_form(forms.Load(rhs3)), b, what="7 fractures load")
        # tangential gradient of a globally linear field is constant per plane: no jumps inside a plane
        edges_basis = tfem.InteriorEdgesFractureBasis(mesh, tfem.ElementLine(1, 2))
        basis2 = tfem.FractureBasis(mesh, tfem.ElementTri(1, 2))
        nodes = mesh["vertices", "coordinates_3d"]
        u_local = (0.3 * nodes[..., 0] - 1.1 * nodes[..., 1] + 0.7 * nodes[..., 2] + 0.2).reshape(-1, 1)
        val, grad = basis2.interpolate(edges_basis, u_local)
        h_e = mesh["interior_edges", "length"].unsqueeze(-2)
        n_e = mesh["interior_edges", "normals_3d"].unsqueeze(-2)
        eta = edges_basis.integrate_functional(forms.Jump(grad), n_e, h_e)
        assert float(eta.abs().max()) < 1e-20, float(eta.abs().max())
        # (values at the edge points are not compared: like the reference, FractureBasis.interpolate
        # to edges indexes the nodal vector with fracture-LOCAL vertex ids, fracture_basis.py:229-231)
        assert val.shape[:3] == grad.shape[:3] == (7, mesh["interior_edges", "cells"].shape[-2], 2)
    return n_g


def rhs3_np(points):
    x, y, z = np.split(points, 3, axis=-1)
    return 6.0 * (y - y**2) * np.abs(x) - 2.0 * (np.abs(z) ** 3 - np.abs(x)) + 1.0
