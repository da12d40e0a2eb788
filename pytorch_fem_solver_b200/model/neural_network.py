"""MLP `u_NN` with input-gradient / Laplacian helpers (reference torch_fem/model/neural_network.py).

The network is the PRODUCER of `grad u_NN(x_q)` consumed by the weak-residual kernels.  On a CUDA device
`gradient` / `value_and_gradient` run the fused forward-mode kernel `tfem_mlp_value_grad` --
one pass over the points, no autograd graph, no per-layer activation tensors -- and `loss.backward()` its
hand-written adjoint `tfem_mlp_value_grad_bwd`; shapes the kernel does not cover (width > 32, other
activations, several outputs, more than 7 square layers) take the reference's autograd route."""

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
        self._neural_network = self.build_network(input_dimension, output_dimension, nb_hidden_layers, neurons_per_layers,
                                                  activation_function, use_xavier_initialization)

    def build_network(self, input_dimension: int, output_dimension: int, nb_layers: int, neurons_per_layers: int,
                      activation_function: torch.nn.Module, use_xavier_initialization: bool) -> torch.nn.Sequential:
        """Linear(in, w) act, `nb_layers` x (Linear(w, w) act), Linear(w, out) (reference :50-75)."""
        widths = [input_dimension] + [neurons_per_layers] * (nb_layers + 1)
        layers: list[torch.nn.Module] = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            layers += [torch.nn.Linear(fan_in, fan_out), activation_function]
        layers.append(torch.nn.Linear(neurons_per_layers, output_dimension))
        network = torch.nn.Sequential(*layers)
        if use_xavier_initialization:
            for layer in network:
                if isinstance(layer, torch.nn.Linear):
                    torch.nn.init.xavier_uniform_(layer.weight)
        return network

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._neural_network(x) * self._boundary_condition_modifier(x)

    #: "auto": the fused kernel where it applies; "torch": always the reference's autograd route
    gradient_path = "auto"

    def _fused_spec(self, inputs: torch.Tensor):
        """(width, n_square, act) when the fused kernel covers this network and these inputs, else None."""
        if self.gradient_path != "auto" or not inputs.is_cuda or inputs.dtype not in (torch.float64, torch.float32):
            return None
        layers = list(self._neural_network)
        linears = [m for m in layers if isinstance(m, torch.nn.Linear)]
        acts = [m for m in layers if not isinstance(m, torch.nn.Linear)]
        if len(linears) < 2 or len(acts) != len(linears) - 1 or any(m.bias is None for m in linears):
            return None
        if isinstance(acts[0], torch.nn.Tanh) and all(isinstance(a, torch.nn.Tanh) for a in acts):
            act = 0
        elif isinstance(acts[0], torch.nn.ReLU) and all(isinstance(a, torch.nn.ReLU) for a in acts):
            act = 1
        else:
            return None
        width, d = linears[0].out_features, linears[0].in_features
        if inputs.shape[-1] != d or not 1 <= d <= 3 or width > 32 or linears[-1].out_features != 1 or linears[-1].in_features != width:
            return None
        if any(m.in_features != width or m.out_features != width for m in linears[1:-1]) or len(linears) - 2 > 7:
            return None
        if any(p.dtype != inputs.dtype or p.device != inputs.device for p in self._neural_network.parameters()):
            return None
        if any(True for _ in self._boundary_condition_modifier.parameters()):
            return None  # a trainable modifier needs the double backward through it
        return width, len(linears) - 2, act

    def _packed_parameters(self) -> torch.Tensor:
        """W0 | b0 | W1 | b1 | ... | w_out | b_out as one vector (differentiable: gradients flow back to the layers)."""
        return torch.cat([p.reshape(-1) for m in self._neural_network if isinstance(m, torch.nn.Linear) for p in (m.weight, m.bias)])

    def value_and_gradient(self, inputs: torch.Tensor):
        """`forward(inputs)` and `gradient(inputs)` from ONE pass (what a VPINN step needs: u_NN for the H1 error,
        grad u_NN for the residual).  u = N b, grad u = b grad N + N grad b with N the MLP body and b the
        boundary-condition modifier, whose own gradient comes from autograd on the (cheap, pointwise) modifier."""
        spec = self._fused_spec(inputs)
        if spec is None:
            return self.forward(inputs), self.gradient(inputs)
        from .. import ops

        width, n_square, act = spec
        flat = inputs.detach().reshape(-1, inputs.shape[-1]).contiguous()
        body, body_grad = ops.mlp_value_grad(self._packed_parameters(), flat, width, n_square, act)
        body = body.reshape(*inputs.shape[:-1], 1)
        body_grad = body_grad.reshape(inputs.shape)
        with torch.enable_grad():
            points = inputs.detach().requires_grad_(True)
            modifier = self._boundary_condition_modifier(points)
            modifier_grad = None
            if modifier.requires_grad:  # (IdentityBC and other constants do not depend on the points)
                modifier_grad = torch.autograd.grad(modifier, points, torch.ones_like(modifier), allow_unused=True)[0]
        modifier = modifier.detach()
        if modifier_grad is None:  # a modifier that does not depend on x (IdentityBC)
            return body * modifier, body_grad * modifier
        return body * modifier, body_grad * modifier + body * modifier_grad.detach()

    def gradient(self, inputs: torch.Tensor) -> torch.Tensor:
        """d u / d x at `inputs`, differentiable w.r.t. the parameters (reference :85-100).

        On the fused path the result is NOT a function of `inputs` in the autograd graph (the points are quadrature
        points): differentiate it once more w.r.t. the points through `laplacian`, or set `gradient_path = "torch"`."""
        inputs.requires_grad_(True)
        if self._fused_spec(inputs) is not None:
            return self.value_and_gradient(inputs)[1]
        return self._gradient_autograd(inputs)

    def _gradient_autograd(self, inputs: torch.Tensor) -> torch.Tensor:
        """The reference's route: reverse mode with `create_graph=True` (also differentiable w.r.t. `inputs`)."""
        inputs.requires_grad_(True)
        output = self.forward(inputs)
        return torch.autograd.grad(output, inputs, torch.ones_like(output), retain_graph=True, create_graph=True)[0]

    def laplacian(self, inputs: torch.Tensor) -> torch.Tensor:
        """Sum of unmixed second derivatives (reference :103-138)."""
        grads = self._gradient_autograd(inputs)  # second derivatives w.r.t. the points: the autograd route
        total = torch.zeros_like(grads[..., :1])
        for i in range(inputs.shape[-1]):
            component = grads[..., i]
            second = torch.autograd.grad(component, inputs, torch.ones_like(component), retain_graph=True, create_graph=True)[0]
            total = total + second[..., i : i + 1]
        return total
