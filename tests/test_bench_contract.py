"""bench.py's output contract, as far as it can be exercised without a GPU: the reference arm prints exactly
ONE JSON line on stdout with the keys the driver reads."""

import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    proc = subprocess.run(
        [sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--nx", "32", "--ny", "32", "--steps", "2", "--warmup", "1"],
        capture_output=True, text=True, env=env, timeout=300, cwd=REPO,
    )
    assert proc.returncode == 0, proc.stderr[-2000:]
    return proc.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = [ln for ln in run_reference().splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config"):
        assert key in line, key
    assert line["unit"] == "elements/s" and line["dtype"] == "f64" and "workload" in line["config"]
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    baseline = line["cpu_baseline"]
    assert baseline["kind"] in ("port", "reference") and baseline["cores"] >= 1 and baseline["value"] == line["value"]
    assert "sample" in baseline


def test_reference_arm_runs_on_rank_zero_only():
    # under torchrun the other ranks exit 0 without work and without output
    assert run_reference({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""
