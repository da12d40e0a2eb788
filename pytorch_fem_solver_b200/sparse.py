"""Iterative solve on the assembled CSR system.

The reference solves `A[inner, inner] x = b[inner]` with a dense `torch.linalg.solve`
(basis/abstract_basis.py:177-195), which cannot exist once the matrix is only representable as
CSR (config 2: 2.1 M unknowns).  Here the same reduced system is solved by Jacobi-preconditioned
conjugate gradients on the full-length vectors with the non-interior rows and columns masked out;
one iteration is three hand-written kernels (`tfem_cg_iteration`: SpMV + p.Ap, the vector updates
+ r.z, the new direction) with every dot product reduced in a fixed order and the scalars kept on
the device; the host looks at the residual only every `check_every` iterations.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops


@dataclass
class CgInfo:
    iterations: int
    relative_residual: float
    converged: bool


def csr_diagonal(crow: torch.Tensor, col: torch.Tensor, val: torch.Tensor) -> torch.Tensor:
    """Diagonal of a CSR matrix (zeros where a row stores no diagonal entry)."""
    n = crow.shape[0] - 1
    counts = (crow[1:] - crow[:-1]).long()
    row_of = torch.repeat_interleave(torch.arange(n, device=col.device), counts)
    on_diag = col.long() == row_of
    diag = torch.zeros(n, dtype=val.dtype, device=val.device)
    diag[row_of[on_diag]] = val[on_diag]
    return diag


def cg(
    crow: torch.Tensor,
    col: torch.Tensor,
    val: torch.Tensor,
    rhs: torch.Tensor,
    keep: Optional[torch.Tensor] = None,
    x0: Optional[torch.Tensor] = None,
    rtol: float = 1e-10,
    max_iterations: Optional[int] = None,
    check_every: int = 50,
    use_graph: bool = True,
    fused: bool = True,
):
    """Solve `(M A M) x = M rhs` for symmetric positive definite `M A M`, `M = diag(keep)`.

    crow/col int32, val/rhs float; keep bool/uint8 of length n or None (every row).  Returns
    `(x, CgInfo)`; x is zero where keep is 0.  On CUDA one iteration is captured in a CUDA graph
    and replayed (`use_graph=False` keeps the eager loop); `fused=False` uses torch vector
    updates around `tfem_csr_spmv` instead of the three-kernel `tfem_cg_iteration`."""
    n = crow.shape[0] - 1
    b = rhs.reshape(-1).to(device=val.device, dtype=val.dtype)
    if b.shape[0] != n:
        raise ValueError("right-hand side length does not match the matrix")
    keep8 = None
    if keep is not None:
        keep8 = keep.reshape(-1).to(device=val.device, dtype=torch.uint8).contiguous()
        b = b * keep8.to(val.dtype)
    diag = csr_diagonal(crow, col, val)
    inv_diag = torch.where(diag != 0, 1.0 / diag, torch.ones_like(diag))
    x = torch.zeros_like(b) if x0 is None else x0.reshape(-1).to(device=val.device, dtype=val.dtype).clone()
    if keep8 is not None:
        x = x * keep8.to(val.dtype)
    r = b - ops.csr_spmv(crow, col, val, x, keep8) if x0 is not None else b.clone()
    z = r * inv_diag
    p = z.clone()
    ap = torch.empty_like(p)
    rz = torch.dot(r, z).reshape(1)
    b_norm = float(torch.linalg.vector_norm(b))
    if b_norm == 0.0:
        return x, CgInfo(0, 0.0, True)
    # Conjugate gradients converge in at most n steps in exact arithmetic; 2 n + 100 leaves room for rounding
    # without turning a breakdown into millions of graph replays
    limit = max_iterations if max_iterations is not None else 2 * n + 100
    initial = float(torch.linalg.vector_norm(r)) / b_norm
    if initial <= rtol or limit <= 0:  # e.g. a converged x0: nothing to iterate
        return x, CgInfo(0, initial, initial <= rtol)
    tiny = torch.finfo(val.dtype).tiny
    spmv = getattr(ops, "csr_spmv_raw", None) if val.is_cuda else None

    def iteration():  # every tensor it touches is preallocated: one CUDA graph replays it
        if spmv is not None:
            spmv(crow, col, val, p, keep8, ap)
        else:
            ap.copy_(ops.csr_spmv(crow, col, val, p, keep8))
        alpha = rz / torch.dot(p, ap).clamp_min(tiny)
        x.addcmul_(p, alpha)
        r.addcmul_(ap, alpha, value=-1.0)
        torch.mul(r, inv_diag, out=z)
        rz_new = torch.dot(r, z).reshape(1)
        p.mul_(rz_new / rz.clamp_min(tiny)).add_(z)
        rz.copy_(rz_new)

    fused_step = getattr(ops, "cg_iteration_raw", None) if (val.is_cuda and fused) else None
    if fused_step is not None:
        # three launches per iteration (tfem_cg_iteration): SpMV + p.Ap | x, r, z updates + r.z | new direction
        blocks = 8 * torch.cuda.get_device_properties(val.device).multi_processor_count
        partial = torch.zeros(2 * blocks, dtype=val.dtype, device=val.device)
        scal = torch.cat([rz, rz]).contiguous()
        inv_diag = inv_diag.contiguous()

        def iteration():  # noqa: F811
            fused_step(crow, col, val, keep8, inv_diag, x, r, z, p, ap, partial, scal)

    step = iteration
    iterations = 0
    if val.is_cuda and use_graph:
        # the loop is launch-bound (a dozen small kernels per iteration): capture one iteration
        side = torch.cuda.Stream(device=val.device)
        side.wait_stream(torch.cuda.current_stream(val.device))
        with torch.cuda.stream(side):
            warm = min(3, limit)  # warm-up outside capture (allocator, lazy module loading); they are real iterations
            for _ in range(warm):
                iteration()
            iterations = warm
        torch.cuda.current_stream(val.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):  # records the launches, does not run them
            iteration()
        step = graph.replay
    rel = float("inf")
    best, stalled = initial, 0
    while iterations < limit:
        step()
        iterations += 1
        if iterations % check_every == 0 or iterations >= limit:
            rel = float(torch.linalg.vector_norm(r)) / b_norm  # the only host synchronisation
            if rel <= rtol:
                break
            # breakdown: a non-SPD or singular reduced system gives p.Ap <= 0 and then inf / NaN, or a residual
            # that no longer decreases; stop and report instead of replaying to the iteration limit
            if rel != rel or rel == float("inf"):
                break
            if rel < 0.999 * best:
                best, stalled = rel, 0
            else:
                stalled += 1
                if stalled >= 20:
                    break
    if rel == float("inf") and iterations > 0:
        rel = float(torch.linalg.vector_norm(r)) / b_norm
    converged = rel == rel and rel <= rtol
    return x, CgInfo(iterations, rel, converged)


class _CgSolve(torch.autograd.Function):
    """x = (M A M)^-1 M b through `cg`, differentiable w.r.t. b: for symmetric A the adjoint is one more solve,
    b_bar = (M A M)^-1 M x_bar (the matrix is a constant of the optimisation, as the Gram matrix of
    examples/example_weak.py:84-86 is)."""

    @staticmethod
    def forward(ctx, b, crow, col, val, keep, rtol, max_iterations):
        x, info = cg(crow, col, val, b.detach(), keep, rtol=rtol, max_iterations=max_iterations)
        if not info.converged:
            raise RuntimeError(f"conjugate gradients stopped at relative residual {info.relative_residual:.3e} after {info.iterations} iterations")
        ctx.save_for_backward(crow, col, val)
        ctx.keep, ctx.rtol, ctx.max_iterations, ctx.shape = keep, rtol, max_iterations, b.shape
        ctx.info = info
        return x.reshape(b.shape)

    @staticmethod
    def backward(ctx, x_bar):
        crow, col, val = ctx.saved_tensors
        if not bool(x_bar.any()):
            return torch.zeros(ctx.shape, dtype=x_bar.dtype, device=x_bar.device), None, None, None, None, None, None
        b_bar, info = cg(crow, col, val, x_bar.detach(), ctx.keep, rtol=ctx.rtol, max_iterations=ctx.max_iterations)
        if not info.converged:
            raise RuntimeError(f"adjoint conjugate gradients stopped at relative residual {info.relative_residual:.3e}")
        return b_bar.reshape(ctx.shape), None, None, None, None, None, None


def solve(matrix: torch.Tensor, rhs: torch.Tensor, keep: Optional[torch.Tensor] = None, rtol: float = 1e-10,
          max_iterations: Optional[int] = None) -> torch.Tensor:
    """`A^-1 rhs` for a symmetric positive definite CSR matrix (rows / columns with keep == 0 masked out), usable inside
    a loss: gradients flow to `rhs` through a second conjugate-gradient solve."""
    return _CgSolve.apply(rhs, matrix.crow_indices().to(torch.int32), matrix.col_indices().to(torch.int32), matrix.values(), keep,
                          rtol, max_iterations)


class _QuadraticForm(torch.autograd.Function):
    """r^T A^-1 r with ONE solve for value and gradient: d/dr = 2 A^-1 r for symmetric A."""

    @staticmethod
    def forward(ctx, r, crow, col, val, keep, rtol, max_iterations):
        x, info = cg(crow, col, val, r.detach(), keep, rtol=rtol, max_iterations=max_iterations)
        if not info.converged:
            raise RuntimeError(f"conjugate gradients stopped at relative residual {info.relative_residual:.3e} after {info.iterations} iterations")
        x = x.reshape(r.shape)
        ctx.save_for_backward(x)
        masked = r.detach() if keep is None else r.detach() * keep.reshape(r.shape).to(r.dtype)
        return (masked * x).sum()

    @staticmethod
    def backward(ctx, loss_bar):
        (x,) = ctx.saved_tensors
        return 2.0 * loss_bar * x, None, None, None, None, None, None


def inverse_quadratic_form(matrix: torch.Tensor, residual: torch.Tensor, keep: Optional[torch.Tensor] = None, rtol: float = 1e-10,
                           max_iterations: Optional[int] = None) -> torch.Tensor:
    """The robust-VPINN loss `r^T G^-1 r` of examples/example_weak.py:84-86,138 on a CSR Gram matrix that cannot be
    inverted densely: one conjugate-gradient solve gives the value and, since G is symmetric, the gradient 2 G^-1 r."""
    return _QuadraticForm.apply(residual, matrix.crow_indices().to(torch.int32), matrix.col_indices().to(torch.int32), matrix.values(),
                                keep, rtol, max_iterations)
