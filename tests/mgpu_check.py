"""Multi-GPU parity check, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py [nx ny]

Every rank assembles its strip with the tiled CUDA kernel, interface rows are exchanged (NCCL / NVLink peer memory),
and each rank compares its OWNED rows with the oracle's assembly of the whole (N-strip) mesh.
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import fem_oracle as fo  # noqa: E402
from pytorch_fem_solver_b200 import distributed, forms, ops  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    torch.set_default_dtype(torch.float64)

    asm = distributed.StripAssembly(nx, ny, rank, world, device, 3, rows_per_tile=64)
    assert asm.interface_tiles.tile_list.numel() + asm.interior_tiles.tile_list.numel() == asm.full_plan.n_tiles
    basis, pat = asm.basis, asm.basis.pattern
    src = forms.SinSinSource()
    values = torch.empty(pat.nnz, dtype=torch.float64, device=device)
    load = torch.empty(pat.n_dof, dtype=torch.float64, device=device)
    plan = basis.tile_plan(64)
    for _ in range(2):  # twice: the exchange must be repeatable
        ops.assemble_csr_tiled(plan.c_struct(), basis._layout.coords, 3, 1.0, 1.0, src.kind, src.params, values, load)
        asm.exchange(values, load)
    torch.cuda.synchronize()
    # the overlapped path (interface tiles, exchange on a side stream, interior tiles) must give the same bits
    for _ in range(2):
        asm.step()
    torch.cuda.synchronize()
    assert torch.equal(asm.values, values) and torch.equal(asm.load, load), "overlapped step differs from the serial one"
    # the same steps fed from and drained to pinned host memory, two in flight
    pipeline = distributed.StripHostPipeline(asm, depth=2)
    coords_host = basis._layout.coords.cpu().pin_memory()
    values_host = torch.empty(pat.nnz, dtype=torch.float64).pin_memory()
    load_host = torch.empty((pat.n_dof, 1), dtype=torch.float64).pin_memory()
    for _ in range(3):
        pipeline.step(coords_host, values_host, load_host)
    pipeline.synchronize()
    assert torch.equal(values_host, values.cpu()) and torch.equal(load_host.reshape(-1), load.cpu()), "host pipeline differs"

    # oracle on the whole mesh (small sizes only)
    parts = [distributed.strip_mesh(nx, ny, r, world) for r in range(world)]
    n_global = parts[0][2]
    coords = np.zeros((n_global, 2))
    conns = []
    for mesh, offset, _ in parts:
        coords[offset : offset + mesh["vertices"].shape[0]] = mesh["vertices"]
        conns.append(mesh["triangles"].astype(np.int64) + offset)
    conn = np.concatenate(conns)
    geo = fo.tri_geometry(coords, conn, 3)
    g_crow, g_col, g_vals = fo.scatter_bilinear_csr(fo.quad_reduce(fo.form_stiffness_mass(geo), geo["dx"]), conn, n_global)
    g_load = fo.scatter_linear(fo.quad_reduce(fo.form_load(geo, fo.source_sinsin(geo["integration_points"])), geo["dx"]), conn, n_global).reshape(-1)

    l2g = asm.plan.local_to_global.cpu().numpy()
    crow, col = pat.crow.cpu().numpy(), pat.col.cpu().numpy()
    vals, vec = values.cpu().numpy(), load.cpu().numpy()
    owned = asm.plan.owned_rows.cpu().numpy()
    worst = 0.0
    scale = np.abs(g_vals).max()
    for i in np.nonzero(owned)[0]:
        g = l2g[i]
        assert np.array_equal(l2g[col[crow[i] : crow[i + 1]]], g_col[g_crow[g] : g_crow[g + 1]]), f"row {g}: pattern"
        worst = max(worst, np.abs(vals[crow[i] : crow[i + 1]] - g_vals[g_crow[g] : g_crow[g + 1]]).max() / scale)
        worst = max(worst, abs(vec[i] - g_load[g]) / np.abs(g_load).max())
    assert worst < 1e-12, worst
    count = torch.zeros(n_global, dtype=torch.int64, device=device)
    count[torch.from_numpy(l2g[owned]).to(device)] = 1
    dist.all_reduce(count)
    assert int(count.sum()) == n_global and int(count.max()) == 1
    print(f"rank {rank}/{world}: {int(owned.sum())} owned rows match the oracle (max rel err {worst:.2e}), "
          f"{asm.exchange.bytes_sent // 2} interface bytes per assembly", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
