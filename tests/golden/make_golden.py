"""Generate golden input/output vectors by running the UNMODIFIED reference here.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

The reference (`torch_fem`, pure Python) imports `tensordict` and `matplotlib`,
neither of which is installed offline.  This script writes throw-away stand-ins
for both into a temporary directory (the `tensordict` stand-in re-exports this
repo's `tensordict_lite` container, `matplotlib.pyplot` is empty), prepends that
directory and /root/reference to `sys.path`, and then drives the reference's own
classes on small seeded meshes.  Everything it saves under `tests/golden/*.npz`
is an output of reference code, so the oracle (`oracle/fem_oracle.py`) and the
CUDA path can be pinned against the reference on a box where the reference
itself does not exist.

Two documented accommodations:
  * `AbstractMesh._triangle_to_tensordict` only accepts numpy dtypes
    (abstract_mesh.py:51-58) while `MeshesTri` hands it torch tensors
    (meshes_tri.py:22-29); for batched meshes the tensors are converted back to
    numpy int32/float64 before the original method runs.
  * meshes carry `edges`, `edge_markers` and, for batched meshes, `neighbors`
    (the reference's fall-back paths are broken).
"""

from __future__ import annotations

import math
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"


def install_standins() -> str:
    root = tempfile.mkdtemp(prefix="tfem_standins_")
    os.makedirs(os.path.join(root, "tensordict"))
    with open(os.path.join(root, "tensordict", "__init__.py"), "w") as fh:
        fh.write("from pytorch_fem_solver_b200.tensordict_lite import TensorDict, stack\n")
    os.makedirs(os.path.join(root, "matplotlib"))
    with open(os.path.join(root, "matplotlib", "__init__.py"), "w") as fh:
        fh.write("")
    with open(os.path.join(root, "matplotlib", "pyplot.py"), "w") as fh:
        fh.write("def subplots(*a, **k):\n    raise RuntimeError('plotting stand-in')\n")
    sys.path[:0] = [root, REFERENCE, REPO]
    return root


def import_reference():
    install_standins()
    import torch_fem  # noqa: E402  (the reference package)
    from torch_fem.mesh.abstract_mesh import AbstractMesh

    original = AbstractMesh._triangle_to_tensordict

    def tolerant(self, mesh_dict):
        converted = {}
        for key, value in mesh_dict.items():
            if isinstance(value, torch.Tensor):
                value = value.numpy()
                if np.issubdtype(value.dtype, np.integer):
                    value = value.astype(np.int32)
                elif np.issubdtype(value.dtype, np.floating):
                    value = value.astype(np.float64)
            converted[key] = value
        return original(self, converted)

    AbstractMesh._triangle_to_tensordict = tolerant
    return torch_fem


def npy(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def rhs(x, y):
    return 2.0 * math.pi**2 * torch.sin(math.pi * x) * torch.sin(math.pi * y)


def bilinear_km(basis):
    return basis.v_grad @ basis.v_grad.mT + basis.v @ basis.v.mT


def bilinear_k(basis):
    return basis.v_grad @ basis.v_grad.mT


def bilinear_m(basis):
    return basis.v @ basis.v.mT


def bilinear_nonsym(basis):
    """Convection-like non-symmetric form: exposes the transposed scatter."""
    beta = torch.tensor([[1.0, 0.5]])
    return basis.v @ (basis.v_grad @ beta.mT).mT + 0.25 * basis.v_grad @ basis.v_grad.mT


def load(basis):
    x, y = torch.split(basis.integration_points, 1, dim=-1)
    return rhs(x, y) * basis.v


def functional(basis):
    x, y = torch.split(basis.integration_points, 1, dim=-1)
    return rhs(x, y) ** 2


def grad_field(points):
    """Smooth stand-in for `neural_network.gradient` (any dimension)."""
    s = points.sum(-1, keepdim=True)
    comps = [torch.cos(2.0 * s + k) * (1.0 + 0.5 * points[..., [k]]) for k in range(points.shape[-1])]
    return torch.cat(comps, dim=-1)


def basis_dump(prefix, basis, out):
    out[prefix + "v"] = npy(basis.v)
    out[prefix + "v_grad"] = npy(basis.v_grad)
    out[prefix + "integration_points"] = npy(basis.integration_points)
    out[prefix + "dx"] = npy(basis._dx)
    out[prefix + "inv_map_jacobian"] = npy(basis._inv_map_jacobian)
    params = basis._basis_parameters
    idx = params["bilinear_form_idx"]
    for k, name in enumerate(["rows", "cols"] if len(idx) == 2 else ["patch", "rows", "cols"]):
        out[prefix + "idx_" + name] = npy(idx[k])
    out[prefix + "linear_idx"] = npy(params["linear_form_idx"][-1])
    out[prefix + "inner_dofs"] = npy(params["inner_dofs"])


def mesh_dump(prefix, mesh, out, keys):
    for key in keys:
        out[prefix + "mesh_" + "_".join(key) if isinstance(key, tuple) else prefix + "mesh_" + key] = npy(mesh[key])


def save(name, out):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(out)} arrays, {os.path.getsize(path)/1024:.1f} KiB")


EDGE_KEYS = [
    ("interior_edges", "cells"),
    ("interior_edges", "vertices"),
    ("interior_edges", "coordinates"),
    ("interior_edges", "length"),
    ("interior_edges", "normals"),
    ("boundary_edges", "cells"),
    ("boundary_edges", "vertices"),
    ("cells", "length"),
    ("cells", "coordinates"),
]


def single_mesh_case(tf, name, mesh_dict, orders, with_neighbors):
    """MeshTri + Basis (+ InteriorEdgesBasis) on one mesh, several quadrature orders."""
    from pytorch_fem_solver_b200 import meshgen  # noqa: F401

    out = {}
    md = dict(mesh_dict)
    if not with_neighbors:
        md.pop("neighbors", None)
    for key, value in md.items():
        out["in_" + key] = value
    mesh = tf.MeshTri(md)
    mesh_dump("", mesh, out, EDGE_KEYS)
    for order in orders:
        p = f"o{order}_"
        basis = tf.Basis(mesh, tf.ElementTri(1, order))
        basis_dump(p, basis, out)
        out[p + "A_km"] = npy(basis.integrate_bilinear_form(bilinear_km))
        out[p + "A_k"] = npy(basis.integrate_bilinear_form(bilinear_k))
        out[p + "A_m"] = npy(basis.integrate_bilinear_form(bilinear_m))
        out[p + "A_nonsym"] = npy(basis.integrate_bilinear_form(bilinear_nonsym))
        out[p + "b_load"] = npy(basis.integrate_linear_form(load))
        out[p + "functional"] = npy(basis.integrate_functional(functional))
        out[p + "reduced_A_km"] = npy(basis.reduce(basis.integrate_bilinear_form(bilinear_km)))
        out[p + "reduced_b"] = npy(basis.reduce(basis.integrate_linear_form(load)))

        # weak residual (examples/example_weak.py:64-75) and its gradient w.r.t. the field
        pts = basis.integration_points.detach().clone()
        g = grad_field(pts).detach().clone().requires_grad_(True)

        def residual(b, gradient_values):
            x, y = torch.split(b.integration_points, 1, dim=-1)
            return rhs(x, y) * b.v - b.v_grad @ gradient_values.mT

        r = basis.integrate_linear_form(residual, g)
        weights = torch.cos(torch.arange(r.numel(), dtype=r.dtype)).reshape(r.shape)
        (r * weights).sum().backward()
        out[p + "residual_grad_in"] = npy(g)
        out[p + "residual"] = npy(r)
        out[p + "residual_cotangent"] = npy(weights)
        out[p + "residual_grad_bar"] = npy(g.grad)

        # interpolate(self, u): values / gradients at own quadrature points
        n_dof = basis._basis_parameters["nb_dofs"]
        u = torch.sin(3.0 * torch.arange(n_dof, dtype=torch.get_default_dtype())).reshape(-1, 1)
        iu, igu = basis.interpolate(basis, u)
        out[p + "interp_u"] = npy(u)
        out[p + "interp_self"] = npy(iu)
        out[p + "interp_self_grad"] = npy(igu)

    # interior-edge basis and the jump estimator (examples/example_jump.py:47-87)
    # ElementLine(1, 3) cannot be used: interior_edges_basis.py:66-67 multiplies a (2,q) matrix
    # into (N_E,1,2,2) coordinates, which only type-checks for q = 2
    for line_order in (2,):
        p = f"e{line_order}_"
        edges_basis = tf.InteriorEdgesBasis(mesh, tf.ElementLine(1, line_order))
        out[p + "v"] = npy(edges_basis.v)
        out[p + "v_grad"] = npy(edges_basis.v_grad)
        out[p + "integration_points"] = npy(edges_basis.integration_points)
        out[p + "dx"] = npy(edges_basis._dx)
        out[p + "inv_map_jacobian"] = npy(edges_basis._inv_map_jacobian)
        basis = tf.Basis(mesh, tf.ElementTri(1, 2))
        n_dof = basis._basis_parameters["nb_dofs"]
        u = torch.sin(3.0 * torch.arange(n_dof, dtype=torch.get_default_dtype())).reshape(-1, 1)
        ie_u, ie_gu = basis.interpolate(edges_basis, u)
        out[p + "interp_u"] = npy(u)
        out[p + "interp_edges"] = npy(ie_u)
        out[p + "interp_edges_grad"] = npy(ie_gu)
        h_e = mesh["interior_edges", "length"].unsqueeze(-2)
        n_e = mesh["interior_edges", "normals"].unsqueeze(-2)

        def jump(_, normal, size):
            plus, minus = torch.unbind(ie_gu, dim=-4)
            return size * ((plus * normal).sum(-1, keepdim=True) + (minus * -normal).sum(-1, keepdim=True)) ** 2

        out[p + "eta"] = npy(edges_basis.integrate_functional(jump, n_e, h_e))

        # closure form: a callable evaluated at the mesh nodes (basis.py:161-177)
        interp, interp_grad = basis.interpolate(edges_basis)

        def nodal(nodes):
            return torch.sin(2.0 * nodes[..., [0]]) * torch.cos(nodes[..., [1]])

        out[p + "closure_nodal"] = npy(nodal(basis._coords4global_dofs))
        out[p + "closure_edges"] = npy(interp(nodal))
        out[p + "closure_edges_grad"] = npy(interp_grad(nodal))
    save(name, out)


def patches_case(tf, name, levels):
    from pytorch_fem_solver_b200 import meshgen

    centers, radius = meshgen.generate_patches_info(levels)
    centers = torch.tensor(centers)
    radius = torch.tensor(radius)
    out = {"in_centers": npy(centers), "in_radius": npy(radius)}
    patches = tf.Patches(centers, radius)
    out["mesh_vertices_coordinates"] = npy(patches["vertices", "coordinates"])
    out["mesh_cells_vertices"] = npy(patches["cells", "vertices"])
    out["mesh_cells_coordinates"] = npy(patches["cells", "coordinates"])
    out["mesh_vertices_markers"] = npy(patches["vertices", "markers"])
    # (Patches.refine_patches cannot be pinned here: the reference's own implementation raises -- it views the
    # children's 5-vertex coordinate block as (-1, 4, 2), mesh/patches.py:88-90,109-111 -- so the port is checked
    # through its geometric properties in tests/test_reference_mirrors.py.)
    for order in (2, 4):
        p = f"o{order}_"
        basis = tf.PatchesBasis(patches, tf.ElementTri(1, order))
        basis_dump(p, basis, out)
        a = basis.integrate_bilinear_form(bilinear_km)
        b = basis.integrate_linear_form(load)
        out[p + "A_km"] = npy(a)
        out[p + "A_k"] = npy(basis.integrate_bilinear_form(bilinear_k))
        out[p + "b_load"] = npy(b)
        out[p + "reduced_A_km"] = npy(basis.reduce(a))
        out[p + "reduced_b"] = npy(basis.reduce(b))
        out[p + "functional"] = npy(basis.integrate_functional(functional))

        pts = basis.integration_points.detach().clone()
        g = grad_field(pts).detach().clone().requires_grad_(True)

        def residual(bb, gradient_values):
            x, y = torch.split(bb.integration_points, 1, dim=-1)
            return rhs(x, y) * bb.v - bb.v_grad @ gradient_values.mT

        r = basis.integrate_linear_form(residual, g)
        weights = torch.cos(torch.arange(r.numel(), dtype=r.dtype)).reshape(r.shape)
        (r * weights).sum().backward()
        out[p + "residual_grad_in"] = npy(g)
        out[p + "residual"] = npy(r)
        out[p + "residual_cotangent"] = npy(weights)
        out[p + "residual_grad_bar"] = npy(g.grad)
        out[p + "reduced_residual"] = npy(basis.reduce(r.detach()))
    save(name, out)


def fracture_case(tf, name, meshes, data):
    out = {"in_fractures_3d_data": data}
    for key in meshes[0]:
        out["in_" + key] = np.stack([m[key] for m in meshes])
    mesh = tf.FracturesTri(triangulations=[dict(m) for m in meshes], fractures_3d_data=torch.tensor(data))
    for key in [
        "jacobian_fracture_map",
        "inv_jacobian_fracture_map",
        "det_jacobian_fracture_map",
        "translation_vector",
    ]:
        out["mesh_" + key] = npy(mesh[key])
    mesh_dump(
        "",
        mesh,
        out,
        EDGE_KEYS
        + [("vertices", "coordinates_3d"), ("cells", "coordinates_3d"), ("interior_edges", "normals_3d")],
    )

    def rhs3(points):
        x, y, z = torch.split(points, 1, dim=-1)
        return 6.0 * (y - y**2) * torch.abs(x) - 2.0 * (torch.abs(z) ** 3 - torch.abs(x)) + 1.0

    def load3(basis):
        return rhs3(basis.integration_points) * basis.v

    def functional3(basis):
        return rhs3(basis.integration_points) ** 2

    for order in (2, 4):
        p = f"o{order}_"
        basis = tf.FractureBasis(mesh, tf.ElementTri(1, order))
        if order == 2:
            for key, value in basis.global_triangulation.items():
                out["gt_" + key] = npy(value)
        basis_dump(p, basis, out)
        out[p + "A_k"] = npy(basis.integrate_bilinear_form(bilinear_k))
        out[p + "A_km"] = npy(basis.integrate_bilinear_form(bilinear_km))
        out[p + "b_load"] = npy(basis.integrate_linear_form(load3))
        out[p + "functional"] = npy(basis.integrate_functional(functional3))

        pts = basis.integration_points.detach().clone()
        g = grad_field(pts).detach().clone().requires_grad_(True)

        def residual(b, gradient_values):
            return rhs3(b.integration_points) * b.v - b.v_grad @ gradient_values.mT

        r = basis.integrate_linear_form(residual, g)
        weights = torch.cos(torch.arange(r.numel(), dtype=r.dtype)).reshape(r.shape)
        (r * weights).sum().backward()
        out[p + "residual_grad_in"] = npy(g)
        out[p + "residual"] = npy(r)
        out[p + "residual_cotangent"] = npy(weights)
        out[p + "residual_grad_bar"] = npy(g.grad)

        n_dof = basis._basis_parameters["nb_dofs"]
        u = torch.sin(3.0 * torch.arange(n_dof, dtype=torch.get_default_dtype())).reshape(-1, 1)
        iu, igu = basis.interpolate(basis, u)
        out[p + "interp_u"] = npy(u)
        out[p + "interp_self"] = npy(iu)
        out[p + "interp_self_grad"] = npy(igu)

    edges_basis = tf.InteriorEdgesFractureBasis(mesh, tf.ElementLine(1, 2))
    p = "e2_"
    out[p + "v"] = npy(edges_basis.v)
    out[p + "v_grad"] = npy(edges_basis.v_grad)
    out[p + "integration_points"] = npy(edges_basis.integration_points)
    out[p + "dx"] = npy(edges_basis._dx)
    out[p + "inv_map_jacobian"] = npy(edges_basis._inv_map_jacobian)
    basis = tf.FractureBasis(mesh, tf.ElementTri(1, 2))
    n_dof = basis._basis_parameters["nb_dofs"]
    u = torch.sin(3.0 * torch.arange(n_dof, dtype=torch.get_default_dtype())).reshape(-1, 1)
    ie_u, ie_gu = basis.interpolate(edges_basis, u)
    out[p + "interp_u"] = npy(u)
    out[p + "interp_edges"] = npy(ie_u)
    out[p + "interp_edges_grad"] = npy(ie_gu)
    h_e = mesh["interior_edges", "length"].unsqueeze(-2)
    n_e = mesh["interior_edges", "normals_3d"].unsqueeze(-2)

    def jump(_, normal, size):
        plus, minus = torch.unbind(ie_gu, dim=-4)
        return size * ((plus * normal).sum(-1, keepdim=True) + (minus * -normal).sum(-1, keepdim=True)) ** 2

    out[p + "eta"] = npy(edges_basis.integrate_functional(jump, n_e, h_e))
    save(name, out)


def main():
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(0)
    tf = import_reference()
    from pytorch_fem_solver_b200 import meshgen

    single_mesh_case(tf, "structured4x4", meshgen.structured_rectangle(4, 4), (1, 2, 3, 4), with_neighbors=False)
    single_mesh_case(
        tf, "structured6x5_jitter", meshgen.structured_rectangle(6, 5, jitter=0.25, seed=11), (3,), with_neighbors=False
    )
    single_mesh_case(tf, "delaunay60", meshgen.delaunay_unit_square(60, seed=0), (2, 3, 4), with_neighbors=False)
    # with `neighbors` the reference lists interior-edge cells in sorted-pair order
    # (abstract_mesh.py:198-228); kept to pin that the kernels are pure functions of the arrays
    single_mesh_case(
        tf, "structured3x3_neighbors", meshgen.structured_rectangle(3, 3, jitter=0.2, seed=5), (2,), with_neighbors=True
    )
    patches_case(tf, "patches_l2", 2)
    meshes, data = meshgen.two_fracture_network(4, 2, jitter=0.25)
    fracture_case(tf, "fractures2_4x2", meshes, data)
    fracture_case(tf, "fractures2_8x4", *meshgen.two_fracture_network(8, 4, jitter=0.25, seed=9))
    # The seven-plane network of BASELINE config 5 cannot be run through the reference:
    # fracture_basis.py:84-92 reshapes the trace-edge list to (F, -1), which needs every
    # fracture to carry the same number of trace edges (the backbone has six traces, the
    # crossing planes one each).  It is pinned through the oracle only.


if __name__ == "__main__":
    main()
