"""BASELINE config 4 end to end at test size: one VPINN step (MLP -> grad u_NN at the quadrature points -> weak residual
on a two-fracture network -> loss over the interior DOFs -> backward) through the fused kernels against the same step
through the reference's autograd route for the network (example_fracture_vpinns.py:104-113,259; example_weak.py:140)."""

import pytest
import torch

import pytorch_fem_solver_b200 as tfem
from pytorch_fem_solver_b200 import forms, meshgen
from tests import api_checks

pytestmark = pytest.mark.gpu
DEV = "cuda"


class BoundaryModifier(torch.nn.Module):
    def forward(self, x):
        return (x[..., :1] ** 2 - 1.0) * x[..., 1:2] * (x[..., 1:2] - 1.0) * (x[..., 2:3] ** 2 - 1.0)


@pytest.mark.parametrize("residual_path", ["tiled", "two_pass"])
def test_vpinn_step_matches_the_autograd_route(residual_path):
    meshes, data = meshgen.two_fracture_network(24, 10)
    with api_checks.default_device(DEV):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 4))
        torch.manual_seed(0)
        net = tfem.FeedForwardNeuralNetwork(3, 1, 6, 25, activation_function=torch.nn.ReLU(), boundary_condition_modifier=BoundaryModifier())
    basis.residual_path = residual_path
    points = basis.integration_points
    inner = basis._basis_parameters["inner_dofs"]
    form = forms.WeakResidual(api_checks.rhs3)
    results = {}
    for route in ("auto", "torch"):
        net.gradient_path = route
        assert (net._fused_spec(points) is not None) == (route == "auto")
        residual = basis.integrate_linear_form(form, net.gradient(points))
        loss = (residual.reshape(-1)[inner] ** 2).sum()
        grads = torch.autograd.grad(loss, list(net.parameters()))
        results[route] = (loss.detach(), residual.detach(), grads)
    loss_f, res_f, grads_f = results["auto"]
    loss_t, res_t, grads_t = results["torch"]
    assert float((res_f - res_t).abs().max()) <= 1e-12 * float(res_t.abs().max())
    assert abs(float(loss_f - loss_t)) <= 1e-12 * float(loss_t)
    scale = max(float(g.abs().max()) for g in grads_t)
    for a, b in zip(grads_f, grads_t):
        assert float((a - b).abs().max()) <= 1e-11 * scale
