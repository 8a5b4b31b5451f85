"""bench.py's JSON-line contract, checked on the arm that runs without a GPU (`--impl reference`: the oracle port on the
host cores) and on the helpers the GPU arm shares with it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=300, env={**os.environ, **(env or {})})
    assert out.returncode == 0, out.stderr[-2000:]
    return [line for line in out.stdout.splitlines() if line.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run("--impl", "reference", "--rays", "64", "--steps", "1", "--warmup", "0", "--log2-hashmap-size", "12")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("train rays/s (fwd+bwd)") and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks print nothing and exit 0."""
    assert _run("--impl", "reference", "--rays", "64", "--steps", "1", "--warmup", "0", "--gpus", "2",
                env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_algorithmic_bytes_follow_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    # main grid L=16, F=2, fp32 rows
    assert bench.algorithmic_bytes_per_point(16) == 12 + 16 * 8 * 2 * 4 + 16 * 2 * 4 == 1164
    assert bench.algorithmic_bytes_per_point(16, bwd=True, dx=True) == 12 + 128 + 1024 + 12 == 1176
    assert bench.algorithmic_bytes_per_point(5) == 372
    peak, src = bench.measured_peaks()
    assert peak > 1000 and src
