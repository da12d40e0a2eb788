"""MLP `u_NN` with input-gradient / Laplacian helpers (reference torch_fem/model/neural_network.py).

The network is the PRODUCER of `grad u_NN(x_q)` consumed by the weak-residual kernels; its dense
GEMM work stays with torch/cuBLAS."""

from __future__ import annotations

from typing import Optional

import torch


class IdentityBC(torch.nn.Module):
    """No boundary modifier: multiplies the network output by one."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.ones_like(x[..., :1])


class FeedForwardNeuralNetwork(torch.nn.Module):
    """Linear -> act -> (Linear -> act) x L -> Linear, times a boundary-condition modifier."""

    def __init__(
        self,
        input_dimension: int,
        output_dimension: int,
        nb_hidden_layers: int,
        neurons_per_layers: int,
        activation_function: torch.nn.Module = torch.nn.Tanh(),
        use_xavier_initialization: bool = False,
        boundary_condition_modifier: Optional[torch.nn.Module] = None,
    ):
        super().__init__()
        self._boundary_condition_modifier = boundary_condition_modifier or IdentityBC()
        widths = [input_dimension] + [neurons_per_layers] * (nb_hidden_layers + 1)
        layers: list[torch.nn.Module] = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            layers += [torch.nn.Linear(fan_in, fan_out), activation_function]
        layers.append(torch.nn.Linear(neurons_per_layers, output_dimension))
        self._neural_network = torch.nn.Sequential(*layers)
        if use_xavier_initialization:
            for layer in self._neural_network:
                if isinstance(layer, torch.nn.Linear):
                    torch.nn.init.xavier_uniform_(layer.weight)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._neural_network(x) * self._boundary_condition_modifier(x)

    def gradient(self, inputs: torch.Tensor) -> torch.Tensor:
        """d u / d x at `inputs`, differentiable w.r.t. the parameters (reference :85-100)."""
        inputs.requires_grad_(True)
        output = self.forward(inputs)
        return torch.autograd.grad(output, inputs, torch.ones_like(output), retain_graph=True, create_graph=True)[0]

    def laplacian(self, inputs: torch.Tensor) -> torch.Tensor:
        """Sum of unmixed second derivatives (reference :103-138)."""
        grads = self.gradient(inputs)
        total = torch.zeros_like(grads[..., :1])
        for i in range(inputs.shape[-1]):
            component = grads[..., i]
            second = torch.autograd.grad(component, inputs, torch.ones_like(component), retain_graph=True, create_graph=True)[0]
            total = total + second[..., i : i + 1]
        return total
