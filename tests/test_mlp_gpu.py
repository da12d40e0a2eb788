"""Fused MLP producer on the GPU against the reference's autograd route
(model/neural_network.py:77-100): u_NN, grad u_NN and the parameter gradients of a loss built from both."""

import pytest
import torch

import pytorch_fem_solver_b200 as tfem

pytestmark = pytest.mark.gpu
DEV = "cuda"


class PolynomialBC(torch.nn.Module):
    """x (x - 1) y (y - 1) [z (z - 1)]: the modifiers of examples/example_patches.py:28-44 / example_fracture_vpinns.py:30-46."""

    def forward(self, x):
        out = torch.ones_like(x[..., :1])
        for c in range(x.shape[-1]):
            out = out * x[..., c : c + 1] * (x[..., c : c + 1] - 1.0)
        return out


def make_network(d, hidden, width, activation, dtype, modifier):
    torch.manual_seed(0)
    net = tfem.FeedForwardNeuralNetwork(d, 1, hidden, width, activation_function=activation, boundary_condition_modifier=modifier)
    return net.to(device=DEV, dtype=dtype)


@pytest.mark.parametrize("d,hidden,width,activation", [(2, 4, 15, torch.nn.Tanh()), (3, 6, 25, torch.nn.ReLU()), (3, 3, 32, torch.nn.Tanh()),
                                                        (1, 0, 7, torch.nn.Tanh()), (2, 7, 25, torch.nn.Tanh())])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("with_modifier", [True, False])
def test_fused_value_gradient_and_parameter_gradients(d, hidden, width, activation, dtype, with_modifier):
    net = make_network(d, hidden, width, activation, dtype, PolynomialBC() if with_modifier else None)
    generator = torch.Generator(device=DEV).manual_seed(1)
    points = torch.rand(5, 301, 3, 1, d, device=DEV, dtype=dtype, generator=generator)  # the shape of basis.integration_points
    assert net._fused_spec(points) is not None
    value, gradient = net.value_and_gradient(points)
    assert torch.equal(net.gradient(points.clone()), gradient)
    cot_v = torch.randn(value.shape, device=DEV, dtype=dtype, generator=generator)
    cot_g = torch.randn(gradient.shape, device=DEV, dtype=dtype, generator=generator)
    loss = (value * cot_v).sum() + (gradient * cot_g).sum() + (gradient**2).sum()
    fused_grads = torch.autograd.grad(loss, list(net.parameters()))

    net.gradient_path = "torch"  # the reference's route, in fp64
    ref = make_network(d, hidden, width, activation, torch.float64, PolynomialBC() if with_modifier else None)
    ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    ref.gradient_path = "torch"
    pts = points.double()
    ref_value = ref(pts)
    ref_gradient = ref.gradient(pts)
    ref_loss = (ref_value * cot_v.double()).sum() + (ref_gradient * cot_g.double()).sum() + (ref_gradient**2).sum()
    ref_grads = torch.autograd.grad(ref_loss, list(ref.parameters()))

    tol = 1e-11 if dtype == torch.float64 else 2e-4
    scale = lambda t: float(t.abs().max()) + 1e-300  # noqa: E731
    assert float((value.double() - ref_value).abs().max()) <= tol * scale(ref_value)
    assert float((gradient.double() - ref_gradient).abs().max()) <= tol * scale(ref_gradient)
    for got, want in zip(fused_grads, ref_grads):
        assert got.shape == want.shape
        assert float((got.double() - want).abs().max()) <= tol * max(scale(want), scale(ref_grads[0])), (got.shape,)


@pytest.mark.parametrize("name,activation", [("tanh2", torch.nn.Tanh()), ("relu3", torch.nn.ReLU())])
def test_fused_kernels_against_the_reference_golden_vectors(name, activation):
    """The kernels against outputs of the UNMODIFIED reference network (tests/golden/mlp_reference.npz)."""
    import os

    import numpy as np

    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mlp_reference.npz"))
    n = int(golden[f"{name}_n_layers"])
    d, width = golden[f"{name}_w0"].shape[1], golden[f"{name}_w0"].shape[0]
    net = tfem.FeedForwardNeuralNetwork(d, 1, n - 2, width, activation_function=activation).to(device=DEV, dtype=torch.float64)
    linears = [m for m in net._neural_network if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for k, m in enumerate(linears):
            m.weight.copy_(torch.from_numpy(golden[f"{name}_w{k}"]))
            m.bias.copy_(torch.from_numpy(golden[f"{name}_b{k}"]))
    points = torch.from_numpy(golden[f"{name}_points"]).to(DEV)
    assert net._fused_spec(points) is not None
    value, gradient = net.value_and_gradient(points)
    assert float((value.cpu() - torch.from_numpy(golden[f"{name}_value"])).abs().max()) <= 1e-12 * float(np.abs(golden[f"{name}_value"]).max())
    assert float((gradient.cpu() - torch.from_numpy(golden[f"{name}_gradient"])).abs().max()) <= 1e-12 * float(np.abs(golden[f"{name}_gradient"]).max())
    loss = (value * torch.from_numpy(golden[f"{name}_cot_v"]).to(DEV)).sum() + (gradient * torch.from_numpy(golden[f"{name}_cot_g"]).to(DEV)).sum()
    grads = torch.autograd.grad(loss, list(net.parameters()))
    scale = max(float(np.abs(golden[f"{name}_grad{k}"]).max()) for k in range(2 * n))
    for k, g in enumerate(grads):
        assert float((g.cpu() - torch.from_numpy(golden[f"{name}_grad{k}"])).abs().max()) <= 1e-12 * scale, k


def test_fused_backward_is_reproducible_and_ragged_sizes():
    net = make_network(3, 6, 25, torch.nn.ReLU(), torch.float64, PolynomialBC())
    for n in (1, 7, 8, 9, 1000, 4099):
        points = torch.rand(n, 3, device=DEV, dtype=torch.float64)
        runs = []
        for _ in range(2):
            value, gradient = net.value_and_gradient(points)
            runs.append(torch.cat([g.reshape(-1) for g in torch.autograd.grad((value**2).sum() + (gradient**2).sum(), list(net.parameters()))]))
        assert torch.equal(runs[0], runs[1])
        net.gradient_path = "torch"
        ref = torch.cat([g.reshape(-1) for g in torch.autograd.grad((net(points) ** 2).sum() + (net.gradient(points) ** 2).sum(), list(net.parameters()))])
        net.gradient_path = "auto"
        assert float((runs[0] - ref).abs().max()) <= 1e-11 * float(ref.abs().max())


def test_unsupported_shapes_take_the_autograd_route():
    wide = make_network(2, 2, 40, torch.nn.Tanh(), torch.float64, None)
    sigmoid = make_network(2, 2, 10, torch.nn.Sigmoid(), torch.float64, None)
    points = torch.rand(64, 2, device=DEV, dtype=torch.float64)
    for net in (wide, sigmoid):
        assert net._fused_spec(points) is None
        assert net.gradient(points).shape == points.shape
    # second derivatives w.r.t. the points stay on the autograd route
    net = make_network(2, 2, 10, torch.nn.Tanh(), torch.float64, PolynomialBC())
    assert net.laplacian(points).shape == (64, 1)
