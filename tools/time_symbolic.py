#!/usr/bin/env python
"""One-time symbolic phase of config 2 on the GPU: native tfem_csr_symbolic against the torch program, tile plan.
Measurement helper (run under gpurun)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytorch_fem_solver_b200 import csr, meshgen  # noqa: E402

mesh = meshgen.structured_rectangle(2048, 1024, jitter=0.25, seed=1234, topology=False)
coords = torch.from_numpy(mesh["vertices"]).cuda()
conn = torch.from_numpy(mesh["triangles"]).cuda()


def timed(label, fn, repeats=3):
    best = 1e9
    for _ in range(repeats):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{label}: {best * 1e3:.1f} ms", flush=True)
    return out


pat = timed("csr pattern, native (tfem_csr_symbolic)", lambda: csr.build_pattern(conn, coords.shape[0]))
os.environ["TFEM_SYMBOLIC"] = "torch"
ref = timed("csr pattern, torch sort/unique/cumsum", lambda: csr.build_pattern(conn, coords.shape[0]))
os.environ["TFEM_SYMBOLIC"] = "native"
same = all(torch.equal(getattr(pat, n), getattr(ref, n)) for n in ("crow", "col", "seg", "perm", "lin_seg", "lin_perm", "keys"))
print("bit-identical:", same, "nnz", pat.nnz)
timed("tile plan (336 rows per tile)", lambda: csr.build_tile_plan(conn, conn, pat, coords, 336, "auto"))
if "--profile" in sys.argv:
    import cProfile
    import pstats

    pr = cProfile.Profile()
    pr.enable()
    csr.build_tile_plan(conn, conn, pat, coords, 336, "auto")
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
