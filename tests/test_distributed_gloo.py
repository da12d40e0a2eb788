"""Multi-rank interface exchange on CPU: world_size 2 and 3 over gloo.

Each rank assembles its element range with the ORACLE (this is a host-logic test of ownership,
ghost columns, key agreement and the deterministic owner-side sum; pack/unpack are plain torch
indexing here, CUDA kernels in production) and the owned rows are compared with the single-mesh
oracle assembly."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fem_oracle as fo
from pytorch_fem_solver_b200 import csr, distributed, meshgen


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_mesh(case):
    if case == "strips":
        nx, ny, world = 6, 3, None
        return None
    if case == "delaunay":
        return meshgen.delaunay_unit_square(80, seed=3)
    return meshgen.permute_mesh(meshgen.structured_rectangle(9, 7, jitter=0.2, topology=False), seed=5)


def _worker(rank, world, port, case):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if case == "strips":
            nx, ny = 6, 3
            parts = [distributed.strip_mesh(nx, ny, r, world, jitter=0.25) for r in range(world)]
            n_global = parts[0][2]
            coords = np.zeros((n_global, 2))
            conns = []
            for mesh, offset, _ in parts:
                coords[offset : offset + mesh["vertices"].shape[0]] = mesh["vertices"]
                conns.append(mesh["triangles"].astype(np.int64) + offset)
            conn = np.concatenate(conns)
            bounds = np.cumsum([0] + [c.shape[0] for c in conns])
        else:
            mesh = _global_mesh(case)
            coords, conn = mesh["vertices"], mesh["triangles"].astype(np.int64)
            n_global = coords.shape[0]
            bounds = np.linspace(0, conn.shape[0], world + 1).astype(int)
        mine = conn[bounds[rank] : bounds[rank + 1]]

        plan = distributed.InterfacePlan(torch.from_numpy(mine), n_global, rank, world)
        pattern = csr.build_pattern(plan.dof_conn, plan.n_local, plan.extra_keys)
        plan.bind(pattern)
        l2g = plan.local_to_global.numpy()
        dof_conn = plan.dof_conn.numpy().astype(np.int64)
        geo = fo.tri_geometry(coords[l2g], dof_conn, 3)
        local = fo.quad_reduce(fo.form_stiffness_mass(geo), geo["dx"]).reshape(-1)
        f_q = fo.source_sinsin(geo["integration_points"])
        lvec = fo.quad_reduce(fo.form_load(geo, f_q), geo["dx"]).reshape(-1)
        owner = torch.repeat_interleave(torch.arange(pattern.nnz), (pattern.seg[1:] - pattern.seg[:-1]).long())
        values = torch.zeros(pattern.nnz, dtype=torch.float64).index_add_(0, owner, torch.from_numpy(local)[pattern.perm.long()])
        load = torch.from_numpy(fo.scatter_linear(lvec, dof_conn, plan.n_local).reshape(-1).copy())

        ops = distributed.ExchangeOps(
            pack=lambda src, idx: src[idx.long()],
            unpack_add=lambda dst, idx, buf: dst.index_add_(0, idx.long(), buf),
        )
        before = values.clone()
        load_before = load.clone()
        distributed.InterfaceExchange(plan, ops)(values, load)
        # the single-buffer / all-gather variant used by StripAssembly.step must give the same bits
        buffer = torch.cat([before, load_before])
        fused = distributed.FusedExchange(plan, pattern.nnz, torch.float64, "cpu", ops)
        fused(buffer)
        assert torch.equal(buffer[: pattern.nnz], values) and torch.equal(buffer[pattern.nnz :], load)
        again = before.clone()
        distributed.InterfaceExchange(plan, ops)(again, load.clone())
        assert torch.equal(values, again), "owner-side sums must be bitwise reproducible"

        # reference: the whole mesh assembled at once
        g_geo = fo.tri_geometry(coords, conn, 3)
        g_crow, g_col, g_vals = fo.scatter_bilinear_csr(fo.quad_reduce(fo.form_stiffness_mass(g_geo), g_geo["dx"]), conn, n_global)
        g_load = fo.scatter_linear(fo.quad_reduce(fo.form_load(g_geo, fo.source_sinsin(g_geo["integration_points"])), g_geo["dx"]), conn, n_global).reshape(-1)
        crow, col = pattern.crow.numpy(), pattern.col.numpy()
        owned = plan.owned_rows.numpy()
        checked = 0
        for i in np.nonzero(owned)[0]:
            g = l2g[i]
            mine_cols = l2g[col[crow[i] : crow[i + 1]]]
            ref_cols = g_col[g_crow[g] : g_crow[g + 1]]
            assert np.array_equal(mine_cols, ref_cols), f"row {g}: pattern differs"
            np.testing.assert_allclose(values.numpy()[crow[i] : crow[i + 1]], g_vals[g_crow[g] : g_crow[g + 1]], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(load.numpy()[i], g_load[g], rtol=1e-12, atol=1e-15)
            checked += 1
        # every touched global row is owned by exactly one rank
        count = torch.zeros(n_global, dtype=torch.int64)
        count[torch.from_numpy(l2g[owned])] = 1
        dist.all_reduce(count)
        touched_any = torch.zeros(n_global, dtype=torch.int64)
        touched_any[torch.from_numpy(np.unique(conn))] = 1
        assert torch.equal(count, touched_any)
        assert checked > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case,world", [("strips", 2), ("strips", 3), ("delaunay", 2), ("permuted", 3)])
def test_interface_exchange(case, world):
    mp.spawn(_worker, args=(world, _free_port(), case), nprocs=world, join=True)
