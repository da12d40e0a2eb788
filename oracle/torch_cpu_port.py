"""CPU baseline: the reference's own tensor program for the assembly path, op for op, in torch.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/fem_oracle.py header).  `/root/reference` does
not exist on the GPU box and is pure Python, so the "reference arm" of bench.py times this
port: the same ATen calls the reference issues, run with all host threads
(`cpu_baseline.kind = "port"`).  Checked against the numpy oracle in tests/test_cpu_port.py.

Reference lines restated (paths relative to /root/reference/torch_fem/):
  mesh/abstract_mesh.py:257-262   gather of cell coordinates
  basis/basis.py:87-96            Jacobian, integration points, weights
  element/element_tri.py:23-41,132-145   barycentric coords, shape gradients, det / inverse
  basis/abstract_basis.py:81-110  (integrand * dx).sum(-3) and index_put_(accumulate=True)
  basis/basis.py:72-77            bilinear_form_idx / linear_form_idx
and, because the dense (N_d, N_d) target of abstract_basis.py:81 cannot exist at 4M elements
(35 TB), the only route the reference's data has to CSR:
  torch.sparse_coo_tensor(bilinear_form_idx, local).coalesce().to_sparse_csr()  (SURVEY.md 8(d)).
"""

from __future__ import annotations

import math
import time

import torch

from .fem_oracle import TRI_QUADRATURE


def reference_assembly_cpu(coords: torch.Tensor, conn: torch.Tensor, order: int = 3, timings: dict | None = None):
    """(K+M) to CSR and the load vector of f = 2 pi^2 sin(pi x) sin(pi y), on the CPU."""
    t0 = time.perf_counter()
    dtype = coords.dtype
    nodes, weights = TRI_QUADRATURE[order]
    nodes = torch.tensor(nodes, dtype=dtype)
    weights = torch.tensor(weights, dtype=dtype).reshape(-1, 1, 1)
    bar_grad = torch.tensor([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]], dtype=dtype)
    n_dof = coords.shape[0]

    cells = coords[conn.long()]  # (N,3,2)
    jac = cells.mT @ bar_grad  # (N,2,2)
    ab, cd = torch.split(jac, 1, dim=-2)
    a, b = torch.split(ab, 1, dim=-1)
    c, d = torch.split(cd, 1, dim=-1)
    det = (a * d - b * c).unsqueeze(-3)
    inv = (1 / det) * torch.stack([torch.concat([d, -b], dim=-1), torch.concat([-c, a], dim=-1)], dim=-2)
    bar = torch.stack([1.0 - nodes[..., [0]] - nodes[..., [1]], nodes[..., [0]], nodes[..., [1]]], dim=-2)
    v = bar
    v_grad = bar_grad @ inv
    points = bar.mT @ cells.unsqueeze(-3)
    dx = 0.5 * weights * det
    t1 = time.perf_counter()

    local_a = ((v_grad @ v_grad.mT + v @ v.mT) * dx).sum(-3)
    t2 = time.perf_counter()
    x, y = torch.split(points, 1, dim=-1)
    rhs = 2.0 * math.pi**2 * torch.sin(math.pi * x) * torch.sin(math.pi * y)
    local_b = (rhs * v * dx).sum(-3)
    t3 = time.perf_counter()

    rows = conn.repeat(1, 3).reshape(-1)
    cols = conn.repeat_interleave(3).reshape(-1)
    matrix = torch.sparse_coo_tensor(torch.stack([rows, cols]).long(), local_a.reshape(-1), (n_dof, n_dof)).coalesce().to_sparse_csr()
    load = torch.zeros((n_dof, 1), dtype=dtype)
    load.index_put_((conn.reshape(-1).long(),), local_b.reshape(-1, 1), accumulate=True)
    t4 = time.perf_counter()
    if timings is not None:
        timings.update(geometry=t1 - t0, local_matrix=t2 - t1, local_load=t3 - t2, coo_to_csr=t4 - t3, total=t4 - t0)
    return matrix, load
