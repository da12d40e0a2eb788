#!/usr/bin/env python
"""Where does the multi-GPU end-to-end step go?  Every rank copies the CSR values + load of config 2 (134 MB) device ->
pinned host, (a) one rank at a time, (b) all ranks at once; the same host -> device for the coordinates (33.6 MB).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_scaling.py
Measurement helper (run under gpurun --gpus N)."""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
dev = torch.device("cuda", local)
n_out, n_in = 134316048 // 8, 33603600 // 8
d_out = torch.randn(n_out, dtype=torch.float64, device=dev)
h_out = torch.empty(n_out, dtype=torch.float64).pin_memory()
d_in = torch.empty(n_in, dtype=torch.float64, device=dev)
h_in = torch.randn(n_in, dtype=torch.float64).pin_memory()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, repeats=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(repeats):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / repeats


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def h2d():
    d_in.copy_(h_in, non_blocking=True)


alone = [0.0, 0.0]
for turn in range(world):
    barrier()
    if turn == rank:
        alone = [timed(d2h), timed(h2d)]
barrier()
together = [timed(d2h), timed(h2d)]
barrier()
both = timed(lambda: (d2h(), h2d()))  # full duplex, all ranks
rows = [None] * world
mine = (rank, alone, together, both)
if world > 1:
    dist.all_gather_object(rows, mine)
else:
    rows = [mine]
if rank == 0:
    gb_out, gb_in = n_out * 8 / 1e9, n_in * 8 / 1e9
    for r, a, t, b in rows:
        print(f"rank {r}: D2H alone {gb_out / a[0]:5.1f} GB/s, all ranks at once {gb_out / t[0]:5.1f} GB/s | H2D alone {gb_in / a[1]:5.1f}, at once {gb_in / t[1]:5.1f} GB/s | "
              f"both directions at once {b * 1e3:.2f} ms per step")
    print(f"aggregate D2H with all {world} ranks copying: {sum(gb_out / t[0] for _, _, t, _ in rows):.1f} GB/s; slowest rank's full-duplex step {max(b for *_, b in rows) * 1e3:.2f} ms")
