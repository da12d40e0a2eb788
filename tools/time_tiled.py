#!/usr/bin/env python
"""Kernel-only timing of tfem_tri_p1_assemble_csr_f64 on BASELINE config 2 for several library builds,
tile shapes and consumer-thread counts in ONE process (mesh, pattern and plans are built once).

    python tools/time_tiled.py --libs default,variants/libX.so --shapes 21x16,16x16 --consumers 384,256

Prints one line per combination: mean / min microseconds over `--steps` launches (L2 flushed in between),
the roofline fraction against MEASURED_PEAKS.json, and -- for builds with -DTFEM_DEBUG_TIMING -- the
per-phase cycle table.  Measurement helper, not part of the product path."""
import argparse
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from pytorch_fem_solver_b200 import _lib, csr, forms, meshgen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default="default")
    ap.add_argument("--shapes", default="21x16")
    ap.add_argument("--consumers", default="384")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--ny", type=int, default=1024)
    ap.add_argument("--check", action="store_true", help="compare every build's output with the first one")
    ap.add_argument("--reserve", type=int, default=0, help="CTA slots left free (148 = one CTA per SM)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    mesh = meshgen.structured_rectangle(args.nx, args.ny, jitter=0.25, seed=1234, topology=False)
    coords = torch.from_numpy(mesh["vertices"]).to(dev)
    conn = torch.from_numpy(mesh["triangles"]).to(dev)
    pat = csr.build_pattern(conn, coords.shape[0])
    n_el, n_v, nnz = conn.shape[0], coords.shape[0], pat.nnz
    algorithmic = 12 * n_el + 16 * n_v + 8 * nnz + 8 * n_v
    peak = 6538.6
    if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")):
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    values = torch.empty(nnz, dtype=torch.float64, device=dev)
    load = torch.empty(n_v, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    form = _lib.Bilinear(1.0, 1.0)
    src = forms.SinSinSource()
    source = _lib.make_source(src.kind, src.params)
    stream = torch.cuda.current_stream(dev).cuda_stream
    reference = None
    plans = {}
    for shape in args.shapes.split(","):
        bx, by = (int(v) for v in shape.split("x"))
        plans[shape] = csr.build_tile_plan(conn, conn, pat, coords, bx * by, "auto", tile_shape=(bx, by))
        p = plans[shape]
        print(f"# plan {shape}: tiles {p.n_tiles} templates {p.n_templates} halo {p.halo_factor:.4f} max_elem {p.max_elem} "
              f"max_vert {p.max_vert} index_bytes {p.index_bytes}", flush=True)
    for lib_path in args.libs.split(","):
        path = _lib.LIB_PATH if lib_path == "default" else os.path.join(REPO, lib_path)
        lib = ctypes.CDLL(path)
        fn = lib.tfem_tri_p1_assemble_csr_f64
        fn.argtypes = _lib._TYPED["tfem_tri_p1_assemble_csr"]
        fn.restype = ctypes.c_int
        lib.tfem_set_device(0)
        timing = getattr(lib, "tfem_debug_timing_read", None)
        for shape, plan in plans.items():
            for consumers in (int(c) for c in args.consumers.split(",")):
                plan.consumer_threads = consumers
                plan.reserve_ctas = args.reserve
                struct = plan.c_struct()

                def launch():
                    status = fn(struct, coords.data_ptr(), 3, form, source, values.data_ptr(), load.data_ptr(), stream)
                    if status != 0:
                        raise RuntimeError(f"status {status}")

                try:
                    for _ in range(3):
                        launch()
                    torch.cuda.synchronize()
                except RuntimeError as error:
                    print(f"{lib_path:32s} {shape:7s} c={consumers}: {error}", flush=True)
                    continue
                if timing is not None:
                    table = (ctypes.c_longlong * 16)()
                    ctas = ctypes.c_int()
                    timing(table, ctypes.byref(ctas))
                times = []
                for i in range(args.steps):
                    flush.fill_(float(i))
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    launch()
                    e.record()
                    e.synchronize()
                    times.append(s.elapsed_time(e) * 1e3)
                mean, best = sum(times) / len(times), min(times)
                note = ""
                if args.check:
                    out = torch.cat([values, load]).clone()
                    if reference is None:
                        reference = out
                    else:
                        scale = float(reference.abs().max())
                        note = f" maxdiff/scale {float((out - reference).abs().max()) / scale:.2e}"
                print(f"{lib_path:32s} {shape:7s} c={consumers}: mean {mean:7.2f} us  min {best:7.2f} us  frac {algorithmic / (mean * 1e-6) / 1e9 / peak:.3f}{note}",
                      flush=True)
                if timing is not None:
                    table = (ctypes.c_longlong * 16)()
                    ctas = ctypes.c_int()
                    timing(table, ctypes.byref(ctas))
                    prod = [table[i] / max(ctas.value, 1) for i in range(8)]  # the CTA counter grows by one per CTA per launch
                    cons = [table[8 + i] / max(ctas.value, 1) for i in range(8)]
                    print("    producer cycles/CTA/launch: wait done %.0f | wait inst %.0f | issue gather %.0f | template %.0f | base+arrive %.0f" % tuple(prod[:5]))
                    if consumers >= 1000:  # role-specialised kernel: thread 0 of the integration group, thread 0 of the reduction group
                        print("    integration thread 0: wait full %.0f | wait table free %.0f | B %.0f   reduction thread 0: wait B %.0f | C %.0f"
                              % (cons[0], cons[1], cons[2], cons[4], cons[5]))
                    else:
                        print("    consumer thread 0 cycles/CTA/launch: wait full %.0f | header %.0f | B + barrier %.0f | wait TC %.0f | (unused) %.0f | C %.0f | end barrier %.0f"
                              % tuple(cons[:7]))


if __name__ == "__main__":
    main()
