"""One VPINN training loop at BASELINE config-4 size on one B200 (the shape of the reference's
examples/example_fracture_vpinns.py).

    python examples/vpinn_config4.py [nx ny steps]

Two fractures of nx x ny squares each (default 1024 x 256: 1 048 576 triangles, 6 291 456 quadrature points of the
6-point rule), u_NN = MLP(3 -> 25 x 7 -> 1, ReLU) x (x^2 - 1) y (y - 1) (z^2 - 1) (example_fracture_vpinns.py:30-46).  Per step:
grad u_NN at every quadrature point (fused forward-mode kernel `tfem_mlp_value_grad`), the weak residual
r_i = sum_q dx (f phi_i - grad phi_i . grad u_NN) summed to the glued DOFs in one launch (`tfem_weak_residual_tiled`),
loss = sum of r_i^2 over the interior DOFs (example_weak.py:140), `loss.backward()` through the residual's adjoint
(`tfem_weak_residual_bwd`) and the MLP's (`tfem_mlp_value_grad_bwd`), Adam.  The reference builds the same step from
autograd double backward and `index_put_` on dense tensors (4.7 s per step on the CPU at this size).
"""

import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pytorch_fem_solver_b200 as tfem  # noqa: E402
from pytorch_fem_solver_b200 import forms, meshgen  # noqa: E402


class BoundaryModifier(torch.nn.Module):
    def forward(self, x):
        return (x[..., :1] ** 2 - 1.0) * x[..., 1:2] * (x[..., 1:2] - 1.0) * (x[..., 2:3] ** 2 - 1.0)


def rhs(points):
    x, y, z = torch.split(points, 1, dim=-1)
    return 6.0 * (y - y**2) * torch.abs(x) - 2.0 * (torch.abs(z) ** 3 - torch.abs(x)) + 1.0


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(0)
    meshes, data = meshgen.two_fracture_network(nx, ny)
    with torch.device("cuda"):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 4))
        net = tfem.FeedForwardNeuralNetwork(3, 1, 6, 25, activation_function=torch.nn.ReLU(), boundary_condition_modifier=BoundaryModifier())
    optimizer = torch.optim.Adam(net.parameters(), lr=1e-3)
    residual_form = forms.WeakResidual(rhs)
    points = basis.integration_points
    inner = basis._basis_parameters["inner_dofs"]

    def step():
        optimizer.zero_grad(set_to_none=True)
        residual = basis.integrate_linear_form(residual_form, net.gradient(points))
        loss = (residual.reshape(-1)[inner] ** 2).sum()
        loss.backward()
        optimizer.step()
        return loss.detach()

    first = float(step())  # builds the tile plan, samples f once, loads the kernels
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    torch.cuda.synchronize()
    per_step = (time.perf_counter() - t0) / steps
    last = float(loss)
    print(f"two fractures {nx} x {ny}: {4 * nx * ny} triangles, {points.numel() // 3} quadrature points, {basis.pattern.n_dof} DOFs")
    print(f"{steps} training steps: {per_step * 1e3:.1f} ms per step; loss {first:.6e} -> {last:.6e}")
    assert last < first


if __name__ == "__main__":
    main()
