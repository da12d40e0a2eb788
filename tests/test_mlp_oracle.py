"""The MLP oracle (oracle/mlp_oracle.py) against outputs of the UNMODIFIED reference network
(tests/golden/mlp_reference.npz, made by tests/golden/make_golden_mlp.py), and the package's own network class on the CPU."""

import os

import numpy as np
import pytest
import torch

from oracle import mlp_oracle as mo

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mlp_reference.npz"))
CASES = {"tanh2": "tanh", "relu3": "relu"}


def load_case(name):
    n = int(GOLDEN[f"{name}_n_layers"])
    weights = [GOLDEN[f"{name}_w{k}"] for k in range(n)]
    biases = [GOLDEN[f"{name}_b{k}"] for k in range(n)]
    return weights, biases


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_value_and_gradient_match_the_reference(name):
    weights, biases = load_case(name)
    value, grad = mo.value_and_gradient(weights, biases, GOLDEN[f"{name}_points"], CASES[name])
    assert np.abs(value[:, None] - GOLDEN[f"{name}_value"]).max() <= 1e-13 * np.abs(GOLDEN[f"{name}_value"]).max()
    assert np.abs(grad - GOLDEN[f"{name}_gradient"]).max() <= 1e-13 * np.abs(GOLDEN[f"{name}_gradient"]).max()


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_parameter_gradients_match_the_reference(name):
    weights, biases = load_case(name)
    w_bars, b_bars = mo.parameter_gradients(weights, biases, GOLDEN[f"{name}_points"], GOLDEN[f"{name}_cot_v"][:, 0], GOLDEN[f"{name}_cot_g"],
                                            CASES[name])
    scale = max(np.abs(GOLDEN[f"{name}_grad{k}"]).max() for k in range(2 * len(weights)))
    for k in range(len(weights)):
        assert np.abs(w_bars[k] - GOLDEN[f"{name}_grad{2 * k}"]).max() <= 1e-12 * scale, k
        assert np.abs(b_bars[k] - GOLDEN[f"{name}_grad{2 * k + 1}"]).max() <= 1e-12 * scale, k


def test_package_network_on_cpu_takes_the_autograd_route_and_matches():
    """Without CUDA tensors the fused kernels do not apply: `gradient` is the reference's reverse-mode route, and the
    packed parameter vector the kernels would read has the layout the C ABI documents."""
    import pytorch_fem_solver_b200 as tfem
    from pytorch_fem_solver_b200 import ops

    name = "tanh2"
    weights, biases = load_case(name)
    net = tfem.FeedForwardNeuralNetwork(2, 1, len(weights) - 2, weights[0].shape[0], activation_function=torch.nn.Tanh()).double()
    linears = [m for m in net._neural_network if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for m, w, b in zip(linears, weights, biases):
            m.weight.copy_(torch.from_numpy(w))
            m.bias.copy_(torch.from_numpy(b))
    points = torch.from_numpy(GOLDEN[f"{name}_points"])
    assert net._fused_spec(points) is None
    assert np.abs(net.gradient(points).detach().numpy() - GOLDEN[f"{name}_gradient"]).max() <= 1e-14
    packed = net._packed_parameters().detach().numpy()
    assert packed.size == ops.mlp_param_count(2, weights[0].shape[0], len(weights) - 2)
    expected = np.concatenate([np.concatenate([w.reshape(-1), b.reshape(-1)]) for w, b in zip(weights, biases)])
    assert np.array_equal(packed, expected)
