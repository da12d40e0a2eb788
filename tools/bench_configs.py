"""Auxiliary measurements for the parity-test configurations of BASELINE.json: C3 patches
weak residual (fp32), C4 two-fracture weak residual forward + adjoint (fp64), C5 seven-fracture stiffness + load
+ jump estimator (fp64).  Not bench.py lines: one table of device times (CUDA events, L2 flushed between
repetitions for the large cases) and achieved bandwidth over the bytes each op must move.

    python tools/bench_configs.py [--scale S]     (S < 1 shrinks the C4 / C5 meshes)
"""

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pytorch_fem_solver_b200 as tfem  # noqa: E402
from pytorch_fem_solver_b200 import forms, meshgen  # noqa: E402

DEV = "cuda"


def timed(fn, repeats, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for i in range(repeats):
        if flush is not None:
            flush.fill_(float(i))
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        fn()
        stop.record()
        torch.cuda.synchronize()
        total += start.elapsed_time(stop)
    return total / repeats * 1e3  # microseconds


def rhs3(points):
    x, y, z = torch.split(points, 1, dim=-1)
    return 6.0 * (y - y**2) * torch.abs(x) - 2.0 * (torch.abs(z) ** 3 - torch.abs(x)) + 1.0


def rhs2(points):
    x, y = torch.split(points, 1, dim=-1)
    return 2.0 * torch.pi**2 * torch.sin(torch.pi * x) * torch.sin(torch.pi * y)


def residual_case(name, basis, n_el, n_q, d, size, flush, rows, analytic=False):
    """Weak residual r = sum_q dx (f v - grad v . grad u) to the DOF vector, and its adjoint."""
    grad_u = torch.randn(*basis.integration_points.shape[:-1], d, device=DEV, dtype=basis.dtype, requires_grad=True)
    # f as a callable, evaluated once at the basis' points and cached on the form (the reference evaluates rhs(x, y) in torch
    # at every step, example_patches.py:102-113); `analytic=True` instead evaluates 2 pi^2 sin(pi x) sin(pi y) inside the kernel
    form = (forms.WeakResidual() if analytic else forms.WeakResidual(rhs2)) if d == 2 else forms.WeakResidual(rhs3)
    r = basis.integrate_linear_form(form, grad_u)
    cot = torch.randn_like(r)

    def forward():
        return basis.integrate_linear_form(form, grad_u)

    def forward_backward():
        out = basis.integrate_linear_form(form, grad_u)
        grad_u.grad = None
        out.backward(cot)

    t_f = timed(lambda: forward(), 20, flush)
    t_fb = timed(forward_backward, 20, flush)
    # the same calls replayed from a CUDA graph: the device time without Python / dispatcher overhead, which is
    # what a captured training step pays (the small patch case is otherwise bound by ~150 us of host work)
    graphed = {}
    graphed["forward+adjoint"] = None  # (autograd's backward runs on its own thread: not captured here)
    for label, fn in (("forward", forward),):
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fn()
            graphed[label] = timed(graph.replay, 20, flush)
        except Exception as error:  # noqa: BLE001
            graphed[label] = None
            print(f"# graph capture of {name} {label} failed: {error!r}", file=sys.stderr)
    n_dof = r.numel()
    bytes_f = 12 * n_el + size * (d * n_q * n_el + n_q * n_el + n_dof) + size * 2 * n_dof  # conn + grad_u + f_q + r + coords
    bytes_b = 12 * n_el + size * (d * n_q * n_el + n_dof)
    rows.append({"case": name + " forward", "elements": n_el, "us": round(t_f, 1), "GB/s": round(bytes_f / t_f / 1e3, 1),
                 "us_graph_replay": None if graphed["forward"] is None else round(graphed["forward"], 1),
                 "GB/s_graph_replay": None if graphed["forward"] is None else round(bytes_f / graphed["forward"] / 1e3, 1)})
    rows.append({"case": name + " forward+adjoint", "elements": n_el, "us": round(t_fb, 1), "GB/s": round((bytes_f + bytes_b) / t_fb / 1e3, 1),
                 "us_graph_replay": None if graphed["forward+adjoint"] is None else round(graphed["forward+adjoint"], 1),
                 "GB/s_graph_replay": None if graphed["forward+adjoint"] is None else round((bytes_f + bytes_b) / graphed["forward+adjoint"] / 1e3, 1)})


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--scale", type=float, default=1.0)
    args = parser.parse_args()
    rows = []
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=DEV)

    # ---- C3: 4096 patches (and 4^10 for a bandwidth-bound variant), fp32, 6-point quadrature ------------
    torch.set_default_dtype(torch.float32)
    for levels, label in ((6, "C3 patches P=4096 fp32"), (10, "C3 scaled P=4^10 fp32")):
        centers, radius = meshgen.generate_patches_info(levels)
        with torch.device(DEV):
            patches = tfem.Patches(torch.tensor(centers, dtype=torch.float32), torch.tensor(radius, dtype=torch.float32))
            basis = tfem.PatchesBasis(patches, tfem.ElementTri(1, 4))
        residual_case(label, basis, 4 * len(centers), 6, 2, 4, flush if levels == 10 else None, rows)
        if levels == 10:
            residual_case(label + " (source evaluated in-kernel)", basis, 4 * len(centers), 6, 2, 4, flush, rows, analytic=True)
        del basis, patches

    # ---- C4: two fractures, 6-point quadrature, fp64 ----------------------------------------------------
    torch.set_default_dtype(torch.float64)
    nx, ny = max(int(1024 * args.scale) // 8 * 8, 8), max(int(256 * args.scale), 2)
    meshes, data = meshgen.two_fracture_network(nx, ny)
    with torch.device(DEV):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 4))
    for path in ("two_pass", "tiled"):  # element kernel + scatter | one launch of the tiled kernel (the default at this size)
        basis.residual_path = path
        residual_case(f"C4 two fractures {nx}x{ny} fp64 ({path})", basis, 2 * 2 * nx * ny, 6, 3, 8, flush, rows)
    del basis, mesh

    # ---- C4 producer: u_NN and grad u_NN at the 6.3 M quadrature points, MLP 3 -> 25 x 7 -> 1 ReLU x BC, and the parameter
    # gradients of a loss of both (example_fracture_vpinns.py:30-46,259): fused forward-mode kernel + hand-written adjoint
    # against the reference's autograd route (torch double backward)
    class Bc3(torch.nn.Module):
        def forward(self, x):
            return (x[..., :1] ** 2 - 1.0) * x[..., 1:2] * (x[..., 1:2] - 1.0) * (x[..., 2:3] ** 2 - 1.0)

    torch.manual_seed(0)
    net = tfem.FeedForwardNeuralNetwork(3, 1, 6, 25, activation_function=torch.nn.ReLU(), boundary_condition_modifier=Bc3()).to(device=DEV, dtype=torch.float64)
    meshes, data = meshgen.two_fracture_network(nx, ny)
    with torch.device(DEV):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 4))
    points = basis.integration_points
    n_pts = points.numel() // 3

    def mlp_forward():
        return net.value_and_gradient(points)

    def mlp_step():
        value, gradient = net.value_and_gradient(points)
        loss = (value**2).sum() + (gradient**2).sum()
        return torch.autograd.grad(loss, list(net.parameters()))

    for path in ("auto", "torch"):
        net.gradient_path = path
        label = "fused kernels" if path == "auto" else "torch autograd (reference route)"
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        t_f = timed(mlp_forward, 3, None)
        t_s = timed(mlp_step, 3, None)
        rows.append({"case": f"C4 MLP 3->25x7->1 ReLU at {n_pts} points, {label}", "elements": n_pts, "us_value_and_gradient": round(t_f, 1),
                     "us_with_parameter_gradients": round(t_s, 1), "peak_extra_GB": round((torch.cuda.max_memory_allocated() - base) / 1e9, 2)})
    del basis, mesh, net, points

    # ---- C5: seven fractures, stiffness + load (generic two-pass path) and the jump estimator ------------
    nx, ny = max(int(1024 * args.scale) // 8 * 8, 8), max(int(586 * args.scale), 2)
    meshes, data = meshgen.seven_fracture_network(nx, ny)
    n_el = 7 * 2 * nx * ny
    with torch.device(DEV):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 3))
        edges = tfem.InteriorEdgesFractureBasis(mesh, tfem.ElementLine(1, 2))
        basis2 = tfem.FractureBasis(mesh, tfem.ElementTri(1, 2))
    pat = basis.pattern
    load_form = forms.Load(rhs3)  # built once: the samples of f at the quadrature points are cached on the form
    algorithmic = 12 * n_el + 8 * 3 * pat.n_dof + 8 * pat.nnz + 8 * pat.n_dof + 8 * 4 * n_el  # conn, coords, values, load, f at 4 points
    for path in ("tiled", "two_pass"):
        t = timed(lambda: basis.assemble(forms.Stiffness(), load_form, layout="values", path=path), 10, flush)
        rows.append({"case": f"C5 seven fractures {nx}x{ny} fp64: K + load to CSR ({path})", "elements": n_el, "us": round(t, 1),
                     "GB/s": round(algorithmic / t / 1e3, 1), "elements/s": round(n_el / t * 1e6)})
    n_edge = int(np.prod(mesh["interior_edges", "cells"].shape[:-1]))
    u = torch.randn(7 * (nx + 1) * (ny + 1), 1, device=DEV)
    h_e = mesh["interior_edges", "length"].unsqueeze(-2)
    n_e = mesh["interior_edges", "normals_3d"].unsqueeze(-2)

    def jump():
        _, grad = basis2.interpolate(edges, u)
        return edges.integrate_functional(forms.Jump(grad), n_e, h_e)

    t = timed(jump, 10, flush)
    # unique streams: per edge cells 8 + x_q 48 + values 32 (w) + gradients 48 (w) + 48 (r) + normal 24 + h 8 + dx 16
    # + eta 8 = 240 B; per cell conn 12 + J^-1 48 + first vertex 24 = 84 B (shared by its edges, served by L2); u once
    bytes_jump = 240 * n_edge + 84 * n_el + 8 * u.numel()
    rows.append({"case": "C5 jump estimator (interpolate to edges + eta_E)", "elements": n_edge, "us": round(t, 1),
                 "GB/s": round(bytes_jump / t / 1e3, 1)})

    for row in rows:
        print(json.dumps(row))


if __name__ == "__main__":
    main()
