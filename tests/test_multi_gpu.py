"""Multi-GPU parity as part of `pytest -m gpu`: runs tests/mgpu_check.py under torchrun on every visible GPU
(skipped below 2 GPUs).  Each rank compares ALL its owned rows with the CPU restatement of the reference, and the
overlapped / delayed / graph-replayed / host-pipelined steps with the serial step, bitwise.  The script's report is
kept in gpurun_out/ (copied to profiles/ by tools/run_scale.sh at round end)."""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode, nx, ny, rows_per_tile, port):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "tests", "mgpu_check.py"), "--mode", mode, "--nx", str(nx), "--ny", str(ny),
           "--rows-per-tile", str(rows_per_tile)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    import re

    # (one entry per rank even if two ranks' lines ran into each other on the shared pipe)
    report = [piece for piece in re.split(r"(?=rank \d+/\d+ \[)", out.stdout) if piece.startswith("rank ")]
    report = [piece.strip() for piece in report]
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", f"mgpu_check_{mode}_{nx}x{ny}_n{n}.txt"), "w") as fh:
        fh.write("\n".join(report) + "\n")
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert len(report) == n


@pytest.mark.parametrize("mode,nx,ny,rows_per_tile", [("weak", 96, 40, 64), ("strong", 96, 80, 64), ("delaunay", 60, 50, 48)])
def test_partitioned_assembly_small(mode, nx, ny, rows_per_tile):
    _run(mode, nx, ny, rows_per_tile, 29531)


@pytest.mark.parametrize("mode", ["weak", "strong"])
def test_partitioned_assembly_bench_size(mode):
    """BASELINE config 2: 2048 x 1024 per rank (weak) and the one 2048 x 1024 mesh cut into ranges (strong), with
    the tile plan bench.py uses."""
    _run(mode, 2048, 1024, 336, 29532)
