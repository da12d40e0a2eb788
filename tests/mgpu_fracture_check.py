"""Multi-GPU parity of the element-partitioned seven-fracture network (BASELINE config 5), under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tests/mgpu_fracture_check.py [--nx 1024 --ny 586]

Every rank assembles its range of the network's element list (stiffness + load of a 3-D source) and exchanges the
interface rows; ALL its owned rows are then compared with the single-GPU assembly of the WHOLE network (the path
tests/test_full_size_gpu.py pins to the oracle), computed on the same device."""

import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pytorch_fem_solver_b200 as tfem  # noqa: E402
from pytorch_fem_solver_b200 import distributed, forms, meshgen  # noqa: E402
from tests.api_checks import rhs3  # noqa: E402



def _say(message: str, flush: bool = True) -> None:
    """One write per report line: `print` issues the text and the newline separately, and under torchrun the ranks'
    lines then run into each other."""
    import sys

    sys.stdout.write(message + "\n")
    sys.stdout.flush()

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=64)
    ap.add_argument("--ny", type=int, default=32)
    ap.add_argument("--rows-per-tile", type=int, default=192)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    torch.set_default_dtype(torch.float64)
    meshes, data = meshgen.seven_fracture_network(args.nx, args.ny)
    with torch.device(device):
        mesh = tfem.FracturesTri(meshes, torch.tensor(data))
        basis = tfem.FractureBasis(mesh, tfem.ElementTri(1, 3))
    bilinear, load_form = forms.Stiffness(), forms.Load(rhs3)
    ref_values, ref_load = basis.assemble(bilinear, load_form, layout="values", path="tiled")
    pat = basis.pattern

    asm = distributed.PartitionedFractureAssembly(basis, rank, world, rows_per_tile=args.rows_per_tile)
    for _ in range(2):  # twice: the exchange must be repeatable
        values, load = asm.step(bilinear, load_form.source)
    torch.cuda.synchronize()
    lp = asm.pattern
    l2g = asm.plan.local_to_global
    owned = asm.plan.owned_rows
    crow = lp.crow.long()
    row_of = torch.repeat_interleave(torch.arange(lp.n_dof, device=device), crow[1:] - crow[:-1])
    mine = owned[row_of]
    keys = l2g[row_of[mine]] * pat.n_dof + l2g[lp.col.long()[mine]]
    pos = torch.searchsorted(pat.keys, keys)
    assert bool((pat.keys[pos.clamp_max(pat.nnz - 1)] == keys).all()), "pattern differs from the whole-network pattern"
    ref_len = (pat.crow[1:] - pat.crow[:-1]).long()
    assert torch.equal((crow[1:] - crow[:-1])[owned], ref_len[l2g[owned]]), "row lengths differ"
    err_m = float((values[mine] - ref_values[pos]).abs().max() / ref_values.abs().max())
    err_l = float((load[owned] - ref_load.reshape(-1)[l2g[owned]]).abs().max() / ref_load.abs().max())
    assert err_m < 1e-12 and err_l < 1e-12, (err_m, err_l)
    count = torch.zeros(pat.n_dof, dtype=torch.int64, device=device)
    count[l2g[owned]] = 1
    dist.all_reduce(count)
    assert int(count.sum()) == pat.n_dof and int(count.max()) == 1, "every global row must be owned exactly once"
    _say(f"rank {rank}/{world} [7 fractures {args.nx}x{args.ny}, elements {asm.lo}:{asm.hi} of {basis._layout.n_total}]: "
          f"{int(owned.sum())} owned rows, {int(mine.sum())} entries match the whole-network assembly (matrix {err_m:.2e}, load {err_l:.2e}); "
          f"{asm.tile_plan.n_tiles} tiles, metric rows of {asm.metric_unit} elements, {asm.exchange.bytes_sent // 2} interface bytes per assembly",
          flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
