#!/usr/bin/env python
"""One forward + adjoint of the C3 (4^10 patches, fp32) and C4 (two fractures, 1.05 M elements, fp64) weak residuals and
of the C4 MLP producer: the launch sequence `ncu` captures for profiles/ (measurement helper, run under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pytorch_fem_solver_b200 as tfem  # noqa: E402
from pytorch_fem_solver_b200 import forms, meshgen  # noqa: E402

DEV = "cuda"


def rhs3(points):
    x, y, z = torch.split(points, 1, dim=-1)
    return 6.0 * (y - y**2) * torch.abs(x) - 2.0 * (torch.abs(z) ** 3 - torch.abs(x)) + 1.0


def run(basis, form, d, repeats=2):
    grad_u = torch.randn(*basis.integration_points.shape[:-1], d, device=DEV, dtype=basis.dtype, requires_grad=True)
    for _ in range(repeats):
        r = basis.integrate_linear_form(form, grad_u)
        grad_u.grad = None
        r.backward(torch.ones_like(r))
    torch.cuda.synchronize()


torch.set_default_dtype(torch.float32)
centers, radius = meshgen.generate_patches_info(10)
with torch.device(DEV):
    patches = tfem.Patches(torch.tensor(centers, dtype=torch.float32), torch.tensor(radius, dtype=torch.float32))
    run(tfem.PatchesBasis(patches, tfem.ElementTri(1, 4)), forms.WeakResidual(), 2)
torch.set_default_dtype(torch.float64)
meshes, data = meshgen.two_fracture_network(1024, 256)
with torch.device(DEV):
    basis = tfem.FractureBasis(tfem.FracturesTri(meshes, torch.tensor(data)), tfem.ElementTri(1, 4))
for path in ("tiled", "two_pass"):
    basis.residual_path = path
    run(basis, forms.WeakResidual(rhs3), 3)
torch.manual_seed(0)
net = tfem.FeedForwardNeuralNetwork(3, 1, 6, 25, activation_function=torch.nn.ReLU()).to(device=DEV, dtype=torch.float64)
points = basis.integration_points
for _ in range(2):
    value, gradient = net.value_and_gradient(points)
    torch.autograd.grad((value**2).sum() + (gradient**2).sum(), list(net.parameters()))
torch.cuda.synchronize()
print("ok")
