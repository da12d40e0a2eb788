"""Multi-GPU parity check, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py [--mode weak|strong|delaunay] [--nx NX --ny NY] [--rows-per-tile R]

  weak      every rank owns an nx x ny strip of a mesh `world` times taller (bench.py's default multi-GPU run)
  strong    the ONE nx x ny mesh (BASELINE config 2 at 2048 x 1024) cut into `world` element ranges
  delaunay  an unstructured mesh cut into element ranges (the mesh-agnostic partitioned assembler)

Every rank assembles its elements with the tiled CUDA kernel, interface rows are exchanged (NVLink peer memory /
NCCL), and each rank compares ALL its OWNED rows (pattern, values, load) with the CPU restatement of the reference
on the whole mesh (weak mode at large sizes: on its own and the adjacent strips, which is everything its rows see).
Also checked, bitwise: the overlapped one-launch step == the serial pack/exchange/add step == the host pipeline, and
a step entered with the main stream deliberately delayed on rank 0 (the receive-only owner).
"""

import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.torch_cpu_port import reference_assembly_cpu  # noqa: E402
from pytorch_fem_solver_b200 import distributed, forms, meshgen, ops  # noqa: E402


def reference_rows(coords, conn):
    """(sorted global keys row*n+col, values, load) of the whole-mesh CPU assembly."""
    n = coords.shape[0]
    matrix, load = reference_assembly_cpu(torch.from_numpy(coords), torch.from_numpy(conn.astype(np.int32)), 3)
    crow, col = matrix.crow_indices().numpy(), matrix.col_indices().numpy()
    rows = np.repeat(np.arange(n), np.diff(crow))
    return rows * n + col, matrix.values().numpy(), load.numpy().reshape(-1), np.diff(crow)



def _say(message: str, flush: bool = True) -> None:
    """One write per report line: `print` issues the text and the newline separately, and under torchrun the ranks'
    lines then run into each other."""
    import sys

    sys.stdout.write(message + "\n")
    sys.stdout.flush()

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="weak", choices=["weak", "strong", "delaunay"])
    ap.add_argument("--nx", type=int, default=96)
    ap.add_argument("--ny", type=int, default=40)
    ap.add_argument("--rows-per-tile", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    torch.set_default_dtype(torch.float64)
    nx, ny = args.nx, args.ny

    # ---- this rank's assembler and the part of the global mesh its owned rows can see ------------------------
    if args.mode == "weak":
        asm = distributed.StripAssembly(nx, ny, rank, world, device, 3, rows_per_tile=args.rows_per_tile)
        near = [r for r in (rank - 1, rank, rank + 1) if 0 <= r < world] if world * nx * ny > 2_000_000 else list(range(world))
        parts = {r: distributed.strip_mesh(nx, ny, r, world) for r in near}
        n_global = parts[rank][2]
        lo_v = min(parts[r][1] for r in near)
        hi_v = max(parts[r][1] + parts[r][0]["vertices"].shape[0] for r in near)
        coords = np.zeros((hi_v - lo_v, 2))
        conns = []
        for r in near:
            mesh, offset, _ = parts[r]
            coords[offset - lo_v : offset - lo_v + mesh["vertices"].shape[0]] = mesh["vertices"]
            conns.append(mesh["triangles"].astype(np.int64) + offset - lo_v)
        conn, shift = np.concatenate(conns), lo_v  # reference numbering = global id - shift
        complete = np.ones(hi_v - lo_v, dtype=bool)  # rows of the reference whose every element is included
        if near[0] > 0:
            complete[: nx + 1] = False
        if near[-1] < world - 1:
            complete[-(nx + 1):] = False
    else:
        if args.mode == "strong":
            mesh = meshgen.structured_rectangle(nx, ny, jitter=0.25, seed=1234, topology=False)
        else:
            mesh = meshgen.delaunay_unit_square(nx * ny, seed=11)
        asm = distributed.PartitionedAssembly.from_mesh(mesh, rank, world, device, quad_order=3, rows_per_tile=args.rows_per_tile)
        coords, conn, shift = mesh["vertices"], mesh["triangles"].astype(np.int64), 0
        n_global = coords.shape[0]
        complete = np.ones(n_global, dtype=bool)
    assert asm.interface_tiles.tile_list.numel() + asm.interior_tiles.tile_list.numel() == asm.full_plan.n_tiles
    basis, pat = asm.basis, asm.basis.pattern
    src = forms.SinSinSource()
    values = torch.empty(pat.nnz, dtype=torch.float64, device=device)
    load = torch.empty(pat.n_dof, dtype=torch.float64, device=device)
    plan = basis.tile_plan(args.rows_per_tile)

    # ---- serial reference path: whole local assembly, then pack / isend-irecv / add ---------------------------
    for _ in range(2):  # twice: the exchange must be repeatable
        ops.assemble_csr_tiled(plan.c_struct(), basis._layout.coords, 3, 1.0, 1.0, src.kind, src.params, values, load)
        asm.exchange(values, load)
    torch.cuda.synchronize()
    # the overlapped path (interface tiles first, exchange on a side stream beside the interior tiles): same bits
    for _ in range(2):
        asm.step()
    torch.cuda.synchronize()
    assert torch.equal(asm.values, values) and torch.equal(asm.load, load), "overlapped step differs from the serial one"
    # ... also when the assembly launch is late: the owner's add must still come after its own interface tiles
    for delayed in range(min(world, 2)):
        asm.buffer.fill_(float("nan"))
        dist.barrier()
        if rank == delayed:
            torch.cuda._sleep(int(2e7))  # ~10 ms on the main stream before this rank's launch
        asm.step()
        torch.cuda.synchronize()
        assert torch.equal(asm.values, values) and torch.equal(asm.load, load), f"step with rank {delayed} delayed differs"
    # the step replayed from CUDA graphs (one per receive-buffer half): same bits again
    if asm.capture():
        for _ in range(3):
            asm.buffer.fill_(float("nan"))
            asm.replay()
            torch.cuda.synchronize()
            assert torch.equal(asm.values, values) and torch.equal(asm.load, load), "graph replay differs from the serial step"
        graph_note = "graphs ok"
    else:
        graph_note = "graphs unavailable"
    # the same steps fed from and drained to pinned host memory, two in flight
    pipeline = distributed.StripHostPipeline(asm, depth=2)
    coords_host = basis._layout.coords.cpu().pin_memory()
    values_host = torch.empty(pat.nnz, dtype=torch.float64).pin_memory()
    load_host = torch.empty((pat.n_dof, 1), dtype=torch.float64).pin_memory()
    for _ in range(3):
        pipeline.step(coords_host, values_host, load_host)
    pipeline.synchronize()
    assert torch.equal(values_host, values.cpu()) and torch.equal(load_host.reshape(-1), load.cpu()), "host pipeline differs"

    # ---- every owned row against the CPU restatement on the whole mesh ----------------------------------------
    ref_keys, ref_vals, ref_load, ref_len = reference_rows(coords, conn)
    n_ref = coords.shape[0]
    l2g = asm.plan.local_to_global.cpu().numpy() - shift
    crow, col = pat.crow.cpu().numpy().astype(np.int64), pat.col.cpu().numpy().astype(np.int64)
    owned = asm.plan.owned_rows.cpu().numpy()
    assert complete[l2g[owned]].all(), "an owned row is not fully covered by the reference sub-mesh"
    row_of = np.repeat(np.arange(pat.n_dof), np.diff(crow))
    mine = owned[row_of]
    keys = l2g[row_of[mine]] * n_ref + l2g[col[mine]]
    assert np.array_equal(np.diff(crow)[owned], ref_len[l2g[owned]]), "row lengths differ from the reference pattern"
    pos = np.searchsorted(ref_keys, keys)
    assert np.array_equal(ref_keys[pos], keys), "pattern differs from the reference"
    vals, vec = values.cpu().numpy(), load.cpu().numpy()
    err_m = np.abs(vals[mine] - ref_vals[pos]).max() / np.abs(ref_vals).max()
    err_l = np.abs(vec[owned] - ref_load[l2g[owned]]).max() / np.abs(ref_load).max()
    assert err_m < 1e-12 and err_l < 1e-12, (err_m, err_l)
    count = torch.zeros(n_global, dtype=torch.int64, device=device)
    count[asm.plan.local_to_global[asm.plan.owned_rows]] = 1
    dist.all_reduce(count)
    assert int(count.sum()) == n_global and int(count.max()) == 1, "every global row must be owned exactly once"
    _say(f"rank {rank}/{world} [{args.mode} {nx}x{ny}]: {int(owned.sum())} owned rows, {int(mine.sum())} entries match the CPU "
          f"restatement (matrix {err_m:.2e}, load {err_l:.2e}); {asm.full_plan.n_tiles} tiles, {asm.full_plan.n_templates} templates, "
          f"{asm.n_interface_tiles} interface tiles, {asm.exchange.bytes_sent // 2} interface bytes per assembly; {graph_note}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
