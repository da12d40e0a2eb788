"""TEST-ONLY stand-ins for the CUDA ops, backed by the numpy oracle.

The product has no CPU path: `pytorch_fem_solver_b200.ops` raises on CPU tensors.  So that the
`-m "not gpu"` suite can still exercise the HOST logic of the API classes (shapes, index maps,
batch flattening, result layouts, closures) on a box without a GPU, `install(monkeypatch)`
swaps the op entry points for oracle-backed functions for the duration of one test.  Nothing
outside `tests/` imports this module, and the `-m gpu` tests run the same checks against the
real kernels.
"""

from __future__ import annotations

import numpy as np
import torch

from oracle import fem_oracle as fo
from pytorch_fem_solver_b200 import ops
from tests.plan_emulator import emulate_tiled


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _frac(n_mesh, jac, inv, det, t):
    if jac is None and det is None:
        return None
    out = {}
    if jac is not None:
        out["jac"] = _np(jac).reshape(n_mesh, 3, 2)
    if inv is not None:
        out["inv"] = _np(inv).reshape(n_mesh, 2, 3)
    if det is not None:
        out["det"] = _np(det).reshape(n_mesh, 1, 1)
    if t is not None:
        out["t"] = _np(t).reshape(n_mesh, 3, 1)
    return out


def _batched(coords, conn, n_el_per_mesh, n_vert_per_mesh):
    n_mesh = conn.shape[0] // n_el_per_mesh
    return _np(coords).reshape(n_mesh, n_vert_per_mesh, 2), _np(conn).reshape(n_mesh, n_el_per_mesh, 3).astype(np.int64), n_mesh


def tri_geometry(coords, conn, n_el_per_mesh, n_vert_per_mesh, quad_order, frac_jac=None, frac_inv=None, frac_det=None, frac_t=None):
    if quad_order not in ops.TRI_NQ:
        raise NotImplementedError("Integration order not implemented")
    c, k, n_mesh = _batched(coords, conn, n_el_per_mesh, n_vert_per_mesh)
    geo = fo.tri_geometry(c, k, quad_order, _frac(n_mesh, frac_jac, frac_inv, frac_det, frac_t))
    n, q = conn.shape[0], ops.TRI_NQ[quad_order]
    d = 3 if frac_jac is not None else 2
    as_t = lambda a, shape: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, a.shape).reshape(shape))).to(coords.dtype)  # noqa: E731
    return (
        as_t(geo["inv_map_jacobian"], (n, 2, d)),
        as_t(geo["v_grad"], (n, 3, d)),
        as_t(geo["integration_points"], (n, q, d)),
        as_t(geo["dx"], (n, q)),
    )


def edge_geometry(edge_coords, n_edge_per_mesh, quad_order, frac_jac=None, frac_det=None, frac_t=None):
    if quad_order not in ops.LINE_NQ:
        raise NotImplementedError("Integration order not implemented")
    n_edge = edge_coords.shape[0]
    n_mesh = n_edge // n_edge_per_mesh
    x = _np(edge_coords).reshape(n_mesh, n_edge_per_mesh, 2, 2)
    frac = _frac(n_mesh, frac_jac, None, frac_det, frac_t)
    if frac is None:
        x = x[0]
    geo = fo.edge_geometry(x, 2, frac) if quad_order == 2 else _edge_order3(x, frac)
    q = ops.LINE_NQ[quad_order]
    d = 3 if frac_jac is not None else 2
    as_t = lambda a, shape: torch.from_numpy(np.ascontiguousarray(a.reshape(shape))).to(edge_coords.dtype)  # noqa: E731
    return (
        as_t(geo["inv_map_jacobian"], (n_edge,)),
        as_t(geo["v_grad"], (n_edge, 2)),
        as_t(geo["integration_points"], (n_edge, q, d)),
        as_t(geo["dx"], (n_edge, q)),
    )


def _edge_order3(x, frac):  # pragma: no cover - the reference cannot run order 3 (see oracle note)
    raise NotImplementedError


def quad_reduce(integrand, dx):
    f = integrand
    if f.shape[0] == 1 and dx.shape[0] != 1:
        f = f.expand(dx.shape[0], *f.shape[1:])
    return (f * dx.unsqueeze(-1)).sum(1)


def scatter(local, seg, perm, inverse):
    flat = local.reshape(-1)[perm.long()]
    n_out = seg.shape[0] - 1
    counts = (seg[1:] - seg[:-1]).long()
    owner = torch.repeat_interleave(torch.arange(n_out), counts)
    return torch.zeros(n_out, dtype=local.dtype).index_add_(0, owner, flat)


def gather(src, idx):
    return src[idx.long()]


def unpack_add_(dst, idx, buf):
    dst[idx.long()] += buf


def local_forms(coords, conn, n_el_per_mesh, n_vert_per_mesh, quad_order, alpha, beta, want_matrix, source_kind,
                source_p, f_q=None, frac_jac=None, frac_inv=None, frac_det=None, frac_t=None):
    c, k, n_mesh = _batched(coords, conn, n_el_per_mesh, n_vert_per_mesh)
    geo = fo.tri_geometry(c, k, quad_order, _frac(n_mesh, frac_jac, frac_inv, frac_det, frac_t))
    n = conn.shape[0]
    mat = torch.empty((0, 3, 3), dtype=coords.dtype)
    vec = torch.empty((0, 3), dtype=coords.dtype)
    if want_matrix:
        integrand = alpha * fo.form_stiffness(geo) + beta * fo.form_mass(geo)
        mat = torch.from_numpy(fo.quad_reduce(integrand, geo["dx"]).reshape(n, 3, 3)).to(coords.dtype)
    if source_kind != ops.SRC_NONE:
        f = _source_values(geo, source_kind, source_p, f_q)
        vec = torch.from_numpy(fo.quad_reduce(fo.form_load(geo, f), geo["dx"]).reshape(n, 3)).to(coords.dtype)
    return mat, vec


def _source_values(geo, kind, p, f_q):
    pts = geo["integration_points"]
    if kind == ops.SRC_SAMPLED:
        return _np(f_q).reshape(pts.shape[:-1] + (1,))
    if kind == ops.SRC_CONST:
        return np.full(pts.shape[:-1] + (1,), p[0])
    if kind == ops.SRC_SINSIN:
        return fo.source_sinsin(pts, p[0], p[1], p[2])
    return np.zeros(pts.shape[:-1] + (1,))


def assemble_csr_tiled(plan_struct, coords, quad_order, alpha, beta, source_kind, source_p, csr_val, load, f_q=None,
                       n_el_per_mesh=0, frac_metric=None):
    """The tiled kernel's phases walked on the CPU (tests/plan_emulator.py) over the oracle's local matrices."""
    plan = assemble_csr_tiled.current_plan
    lay = assemble_csr_tiled.current_layout
    mat, vec_local = local_forms(lay.coords, lay.conn, lay.n_el_per_mesh, lay.n_vert_per_mesh, quad_order, alpha, beta, True,
                                 source_kind if load is not None else ops.SRC_CONST, source_p if load is not None else [0.0] * 4,
                                 f_q, *lay.frac_args())
    n = lay.conn.shape[0]
    if frac_metric is not None:  # the metric handed to the kernel must be J_f^+ J_f^+^T and det J_f
        inv, det = _np(lay.frac[1]), _np(lay.frac[2])
        expect = np.stack([(inv @ inv.transpose(0, 2, 1))[:, 0, 0], (inv @ inv.transpose(0, 2, 1))[:, 0, 1],
                           (inv @ inv.transpose(0, 2, 1))[:, 1, 1], det.reshape(-1)], axis=1)
        np.testing.assert_allclose(_np(frac_metric), expect, rtol=1e-13, atol=1e-15)
    offsets = (np.arange(n) // lay.n_el_per_mesh) * lay.n_vert_per_mesh
    geom_conn = _np(lay.conn).astype(np.int64) + offsets[:, None]
    nnz, n_dof = assemble_csr_tiled.current_sizes
    vals, vec = emulate_tiled(plan, None, _np(mat), _np(vec_local), geom_conn, nnz, n_dof)
    if csr_val is not None:
        csr_val.copy_(torch.from_numpy(vals))
    if load is not None:
        load.copy_(torch.from_numpy(vec))


class _WeakResidual(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_u, meta):
        (coords, conn, dof_conn, lin_seg, lin_perm, n_el_per_mesh, n_vert_per_mesh, quad_order, source_kind, source_p,
         f_q, fj, fi, fd, ft) = meta
        c, k, n_mesh = _batched(coords, conn, n_el_per_mesh, n_vert_per_mesh)
        geo = fo.tri_geometry(c, k, quad_order, _frac(n_mesh, fj, fi, fd, ft))
        f = _source_values(geo, source_kind, source_p, f_q)
        g = _np(grad_u).reshape(geo["integration_points"].shape)
        local = fo.quad_reduce(fo.form_weak_residual(geo, f, g), geo["dx"]).reshape(-1, 3)
        n_dof = lin_seg.shape[0] - 1
        r = fo.scatter_linear(local, _np(dof_conn).astype(np.int64), n_dof).reshape(-1)
        ctx.geo, ctx.dof_conn, ctx.shape = geo, _np(dof_conn).astype(np.int64), grad_u.shape
        return torch.from_numpy(r).to(grad_u.dtype)

    @staticmethod
    def backward(ctx, r_bar):
        geo = ctx.geo
        lead = geo["v_grad"].shape[:-3]
        g_bar = fo.weak_residual_backward(geo, ctx.dof_conn.reshape(lead + (3,)), _np(r_bar))
        return torch.from_numpy(np.ascontiguousarray(g_bar)).reshape(ctx.shape).to(r_bar.dtype), None


def weak_residual(grad_u, coords, conn, dof_conn, lin_seg, lin_perm, n_el_per_mesh, n_vert_per_mesh, quad_order,
                  source_kind, source_p, f_q=None, frac_jac=None, frac_inv=None, frac_det=None, frac_t=None):
    meta = (coords, conn, dof_conn, lin_seg, lin_perm, n_el_per_mesh, n_vert_per_mesh, quad_order, source_kind,
            source_p, f_q, frac_jac, frac_inv, frac_det, frac_t)
    return _WeakResidual.apply(grad_u, meta)


def interp_cells(u, dof_conn, v_grad, quad_order, seg=None, perm=None, inverse=None):
    nodes, _ = fo.tri_quadrature(quad_order)
    bar = fo.tri_barycentric(nodes)  # (q,3,1)
    nodal = _np(u)[_np(dof_conn).astype(np.int64)]  # (N,3)
    val = (nodal[:, None, :, None] * bar).sum(-2).reshape(nodal.shape[0], -1)
    grad = (nodal[:, :, None] * _np(v_grad)).sum(-2)
    return torch.from_numpy(val).to(u.dtype), torch.from_numpy(grad).to(u.dtype)


def interp_edges(u, edge_cells, conn, first_vertex, inv_jac, x_q, n_edge_per_mesh, n_el_per_mesh, seg=None, perm=None, inverse=None):
    n_edge, n_q, d = x_q.shape
    n_mesh = n_edge // n_edge_per_mesh
    cells = _np(edge_cells).astype(np.int64).reshape(n_mesh, n_edge_per_mesh, 2)
    k = _np(conn).astype(np.int64).reshape(n_mesh, n_el_per_mesh, 3)
    first = _np(first_vertex).reshape(n_mesh, n_el_per_mesh, 1, d)
    inv = _np(inv_jac).reshape(n_mesh, n_el_per_mesh, 1, 2, d)
    pts = _np(x_q).reshape(n_mesh, n_edge_per_mesh, 1, n_q, d)
    val, grad = fo.interpolate_edges(pts, cells, k, first, inv, _np(u))
    return (
        torch.from_numpy(np.ascontiguousarray(val)).reshape(n_edge, 2, n_q).to(u.dtype),
        torch.from_numpy(np.ascontiguousarray(grad)).reshape(n_edge, 2, d).to(u.dtype),
    )


def edge_jump(grad_edges, normals, h_e, dx):
    plus = (grad_edges[:, 0] * normals).sum(-1)
    minus = (grad_edges[:, 1] * -normals).sum(-1)
    return (h_e * (plus + minus) ** 2).unsqueeze(-1).mul(dx).sum(-1)


def csr_spmv(crow, col, val, x, keep=None):
    """Plain numpy restatement of y = A x with optional row mask (test stand-in for tfem_csr_spmv)."""
    c, j, v = crow.numpy().astype(np.int64), col.numpy().astype(np.int64), val.numpy()
    prod = v * x.numpy()[j]
    y = np.add.reduceat(np.concatenate([prod, [0.0]]), np.minimum(c[:-1], prod.size)) * (c[1:] > c[:-1])
    if keep is not None:
        y = y * (keep.numpy() != 0)
    return torch.from_numpy(np.asarray(y, dtype=v.dtype))


def install(monkeypatch):
    """Route the op entry points to the oracle for one test (CPU host-logic checks only)."""
    from pytorch_fem_solver_b200.basis import abstract_basis

    monkeypatch.setattr(ops, "place_mesh", lambda mesh: mesh)
    for name in ("tri_geometry", "edge_geometry", "quad_reduce", "scatter", "gather", "unpack_add_", "local_forms",
                 "weak_residual", "interp_cells", "interp_edges", "edge_jump", "assemble_csr_tiled", "csr_spmv"):
        monkeypatch.setattr(ops, name, globals()[name])

    original = abstract_basis.AbstractBasis.tile_plan

    def tile_plan(self, *args, **kwargs):
        plan = original(self, *args, **kwargs)
        assemble_csr_tiled.current_plan = plan
        assemble_csr_tiled.current_layout = self._layout
        assemble_csr_tiled.current_sizes = (self.pattern.nnz, self.pattern.n_dof)
        return plan

    monkeypatch.setattr(abstract_basis.AbstractBasis, "tile_plan", tile_plan)
