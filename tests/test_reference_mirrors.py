"""Counterparts of the reference's own tests that are not golden-vector comparisons
(SURVEY.md section 4): NN input derivatives vs finite differences, one embedded fracture == the planar
problem, PatchesBasis on one patch == Basis on the same 5-vertex mesh.  CPU runs go through the
oracle-backed stand-ins (host logic); `-m gpu` runs use the CUDA kernels."""

import numpy as np
import pytest
import torch

import pytorch_fem_solver_b200 as tfem
from pytorch_fem_solver_b200 import forms, meshgen
from tests import cpu_shim


class BoundaryConstrain(torch.nn.Module):
    def forward(self, inputs):
        x, y = torch.split(inputs, 1, dim=-1)
        return x * (x - 1) * y * (y - 1)


def test_network_derivatives_match_finite_differences():
    """reference tests/test_derivate_wrt_inputs.py:17-106 (fp64: atol 1e-8 / 1e-6 for the Laplacian)."""
    previous = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(0)
        net = tfem.FeedForwardNeuralNetwork(2, 1, 4, 25, boundary_condition_modifier=BoundaryConstrain())
        points = torch.rand(200, 4, 1, 2)  # any (.., q, 1, 2) set of points, as basis.integration_points
        x, y = torch.split(points, 1, dim=-1)
        h = 2.0**-9
        grad = net.gradient(points.clone())
        dx = (net(torch.cat([x + h, y], -1)) - net(torch.cat([x - h, y], -1))) / (2 * h)
        dy = (net(torch.cat([x, y + h], -1)) - net(torch.cat([x, y - h], -1))) / (2 * h)
        assert torch.allclose(dx, grad[..., :1], atol=1e-5) and torch.allclose(dy, grad[..., 1:], atol=1e-5)
        lap = net.laplacian(points.clone())
        d2x = (net(torch.cat([x + h, y], -1)) - 2 * net(points) + net(torch.cat([x - h, y], -1))) / h**2
        d2y = (net(torch.cat([x, y + h], -1)) - 2 * net(points) + net(torch.cat([x, y - h], -1))) / h**2
        assert torch.allclose(d2x + d2y, lap, atol=1e-4)
    finally:
        torch.set_default_dtype(previous)


def one_fracture_equals_planar(device):
    """reference tests/test_fracture_jump.py: a single fracture in the plane z = 0 reproduces the 2-D
    basis: matrices, load, interpolants at the edge points, jump estimator."""
    previous = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        mesh_dict = meshgen.structured_rectangle(6, 5, jitter=0.2, seed=8, corners_first=True)
        data = torch.tensor([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0]]])
        with torch.device(device):
            planar = tfem.MeshTri(mesh_dict)
            basis2 = tfem.Basis(planar, tfem.ElementTri(1, 2))
            edges2 = tfem.InteriorEdgesBasis(planar, tfem.ElementLine(1, 2))
            embedded = tfem.FracturesTri([{k: v.copy() for k, v in mesh_dict.items()}], data.to(device))
            basis3 = tfem.FractureBasis(embedded, tfem.ElementTri(1, 2))
            edges3 = tfem.InteriorEdgesFractureBasis(embedded, tfem.ElementLine(1, 2))
        # the fracture basis numbers its DOFs by the sorted 3-D vertices: local vertex i is global DOF to_global[i]
        to_global = basis3.global_triangulation["global2local_idx"].reshape(-1).long()
        a2 = basis2.integrate_bilinear_form(forms.StiffnessMass())
        a3 = basis3.integrate_bilinear_form(forms.StiffnessMass())[to_global][:, to_global]
        assert torch.allclose(a2, a3, rtol=1e-12, atol=1e-14)
        source = lambda points: 1.0 + points[..., :1] * points[..., 1:2]  # noqa: E731
        b2 = basis2.integrate_linear_form(forms.Load(source))
        b3 = basis3.integrate_linear_form(forms.Load(source))[to_global]
        assert torch.allclose(b2, b3, rtol=1e-12, atol=1e-15)
        gen = torch.Generator().manual_seed(3)
        u = torch.randn(b2.shape[0], 1, generator=gen, dtype=torch.float64).to(device)
        v2, g2 = basis2.interpolate(edges2, u)
        v3, g3 = basis3.interpolate(edges3, u)
        assert torch.allclose(v2.reshape(-1), v3.reshape(-1), rtol=1e-12, atol=1e-14)
        assert torch.allclose(g2.reshape(-1, 2), g3.reshape(-1, 3)[:, :2], rtol=1e-12, atol=1e-13)
        assert float(g3.reshape(-1, 3)[:, 2].abs().max()) < 1e-13
        eta2 = edges2.integrate_functional(forms.Jump(g2), planar["interior_edges", "normals"].unsqueeze(-2),
                                           planar["interior_edges", "length"].unsqueeze(-2))
        eta3 = edges3.integrate_functional(forms.Jump(g3), embedded["interior_edges", "normals_3d"].unsqueeze(-2),
                                           embedded["interior_edges", "length"].unsqueeze(-2))
        assert torch.allclose(eta2.reshape(-1), eta3.reshape(-1), rtol=1e-11, atol=1e-14)
    finally:
        torch.set_default_dtype(previous)


def one_patch_equals_basis(device):
    """reference tests/test_assembly_patches.py:1-74: PatchesBasis on the patch (centre .5,.5, r=.5) against Basis
    on the same 5-vertex / 4-triangle mesh; also the known answers of SURVEY.md 8(c)."""
    previous = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        with torch.device(device):
            patches = tfem.Patches(torch.tensor([[0.5, 0.5]]), torch.tensor([[0.5]]))
            pbasis = tfem.PatchesBasis(patches, tfem.ElementTri(1, 2))
            verts = patches["vertices", "coordinates"].reshape(5, 2).cpu().numpy()
            cells = patches["cells", "vertices"].reshape(4, 3).cpu().numpy().astype(np.int32)
            mesh = tfem.MeshTri({"vertices": verts, "triangles": cells, "vertex_markers": np.array([[1], [1], [1], [1], [0]], dtype=np.int32)})
            basis = tfem.Basis(mesh, tfem.ElementTri(1, 2))
        a_p = pbasis.integrate_bilinear_form(forms.Stiffness()).reshape(5, 5)
        a_b = basis.integrate_bilinear_form(forms.Stiffness())
        assert torch.allclose(a_p, a_b, rtol=1e-12, atol=1e-14)
        expected = torch.tensor([[1.0, 0, 0, 0, -1], [0, 1, 0, 0, -1], [0, 0, 1, 0, -1], [0, 0, 0, 1, -1], [-1, -1, -1, -1, 4]])
        assert torch.allclose(a_p.cpu(), expected, atol=1e-13)
        b_p = pbasis.integrate_linear_form(forms.Load(forms.ConstSource(1.0))).reshape(5)
        b_b = basis.integrate_linear_form(forms.Load(forms.ConstSource(1.0))).reshape(5)
        assert torch.allclose(b_p, b_b, rtol=1e-12, atol=1e-15)
        assert torch.allclose(b_p.cpu(), torch.tensor([1 / 6, 1 / 6, 1 / 6, 1 / 6, 1 / 3]), atol=1e-14)
    finally:
        torch.set_default_dtype(previous)


def test_one_fracture_equals_planar_host_logic(monkeypatch):
    cpu_shim.install(monkeypatch)
    one_fracture_equals_planar("cpu")


def test_one_patch_equals_basis_host_logic(monkeypatch):
    cpu_shim.install(monkeypatch)
    one_patch_equals_basis("cpu")


@pytest.mark.gpu
def test_one_fracture_equals_planar_gpu():
    one_fracture_equals_planar("cuda")


@pytest.mark.gpu
def test_one_patch_equals_basis_gpu():
    one_patch_equals_basis("cuda")


def test_refine_patches_children_tile_their_parent():
    """mesh/patches.py:49-149.  The reference's refine_patches raises on its own data (it reshapes the children's
    5-vertex blocks to (-1, 4, 2)), so the port is pinned by what the routine is meant to produce: four children of
    half the radius centred on the parent's quadrant centres, one patch of radius sqrt(2)/2 r rotated by 45 degrees
    about the parent's centre, old patches dropped or kept, and vertex blocks that are exactly the patches
    `Patches(centers, radius)` would build."""
    import math

    from pytorch_fem_solver_b200 import meshgen
    from pytorch_fem_solver_b200.mesh.patches import Patches

    centers, radius = meshgen.generate_patches_info(2)
    previous = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        _check_refinement(Patches, centers, radius, math)
    finally:
        torch.set_default_dtype(previous)


def _check_refinement(Patches, centers, radius, math):
    patches = Patches(torch.tensor(centers), torch.tensor(radius))
    marks = torch.zeros(16, dtype=torch.bool)
    marks[[1, 4, 5, 11]] = True
    for keep_old in (False, True):
        c, r, v = patches.refine_patches(marks, maintain_old_patches=keep_old)
        n_old = 16 if keep_old else 12
        assert c.shape == (n_old + 20, 2) and r.shape == (n_old + 20, 1) and v.shape == (n_old + 20, 5, 2)
        rebuilt = Patches(c[: n_old + 16], r[: n_old + 16])["vertices", "coordinates"]
        assert torch.allclose(v[: n_old + 16], rebuilt, rtol=0, atol=1e-15)  # kept patches and axis-aligned children
        parents_c, parents_r = torch.tensor(centers)[marks], torch.tensor(radius)[marks]
        children_c = c[n_old : n_old + 16].reshape(4, 4, 2)
        assert torch.allclose(children_c.mean(1), parents_c)  # the four children surround the parent's centre
        assert torch.allclose(r[n_old : n_old + 16], (0.5 * parents_r).repeat(4, 1))
        area = lambda rr: (2 * rr) ** 2  # noqa: E731
        assert torch.allclose(4 * area(r[n_old : n_old + 4]), area(parents_r))  # children tile the parent
        rot_c, rot_r, rot_v = c[n_old + 16 :], r[n_old + 16 :], v[n_old + 16 :]
        assert torch.equal(rot_c, parents_c) and torch.allclose(rot_r, parents_r / math.sqrt(2.0))
        assert torch.allclose(rot_v[:, 4], parents_c)  # centre vertex last
        corner = rot_v[:, :4] - parents_c.unsqueeze(1)
        assert torch.allclose(corner.norm(dim=-1), (math.sqrt(2.0) * rot_r).expand(-1, 4))  # a square of half-diagonal sqrt(2) r'
        assert torch.allclose(corner.abs().min(dim=-1).values, torch.zeros(4, 4, dtype=corner.dtype), atol=1e-15)  # corners on the axes
    c, r, v = patches.uniform_refine(1)
    assert c.shape[0] == 16 * 5 and torch.allclose(r[:64], torch.full((64, 1), 0.0625, dtype=r.dtype))


def test_p2_shape_functions_match_the_reference():
    """`ElementTri(2, k).compute_shape_functions` against outputs of the unmodified reference (element_tri.py:43-70;
    tests/golden/p2_shape.npz made by tests/golden/make_golden_p2.py); order 3 elements still raise."""
    import os

    import numpy as np
    import pytest
    import torch

    import pytorch_fem_solver_b200 as tfem

    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "p2_shape.npz"))
    previous = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        for order in (2, 3, 4):
            element = tfem.ElementTri(2, order)
            bar = element.compute_barycentric_coordinates(element.gaussian_nodes)
            assert np.array_equal(bar.numpy(), golden[f"o{order}_bar"])
            v, v_grad = element.compute_shape_functions(bar, torch.from_numpy(golden[f"o{order}_inv"]))
            assert v.shape == golden[f"o{order}_v"].shape and v_grad.shape == golden[f"o{order}_v_grad"].shape
            assert np.abs(v.numpy() - golden[f"o{order}_v"]).max() <= 1e-15
            assert np.abs(v_grad.numpy() - golden[f"o{order}_v_grad"]).max() <= 1e-14
            assert abs(float(v.sum(-2).max()) - 1.0) < 1e-14  # partition of unity
        cubic = tfem.ElementTri(3, 2)
        with pytest.raises(NotImplementedError):
            cubic.compute_shape_functions(bar, torch.eye(2).reshape(1, 1, 2, 2))
    finally:
        torch.set_default_dtype(previous)
