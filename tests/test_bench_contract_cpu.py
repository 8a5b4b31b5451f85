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
    peak, tflops, src = bench.measured_peaks()
    assert peak > 1000 and tflops > 100 and src


def test_kernel_rooflines_from_launch_tags():
    """units x per-unit algorithmic bytes / event time, per launch tag; the dominant kernel is picked among ALL tags."""
    sys.path.insert(0, ROOT)
    import bench
    peaks = (6500.0, 1600.0, "test")
    l2 = {"gather8": 240.0, "red_v2": 160.0}
    n = 4096 * 48
    r = bench.kernel_roofline("tn_hash_encode_bwd[L16,T2^19,dx]", 4, 0.4, 4 * n, peaks, l2)
    assert r["bound"] == "hbm" and r["bytes_per_unit"] == 1176
    assert abs(r["achieved"] - 1176 * n / 0.1e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / 6500.0) < 1e-12
    floor = n * 128 / 240e9 + n * 128 / 160e9
    assert abs(r["l2"]["floor_ms_per_launch"] - floor * 1e3) < 1e-9 and abs(r["l2"]["frac"] - floor / 0.1e-3) < 1e-9
    f = bench.kernel_roofline("tn_hash_encode_fwd[L16,T2^19]", 1, 0.05, n, peaks, l2)
    assert f["bytes_per_unit"] == 1164 and "l2" in f
    p = bench.kernel_roofline("tn_prop_density_bwd[L5,S256,dx]", 2, 0.4, 2 * 4096 * 256, peaks)
    assert p["bytes_per_unit"] == 12 + 4 + 320 + 12
    m = bench.kernel_roofline("tn_mlp_tc_fwd[32-64x1-16]", 1, 0.03, n, peaks)
    assert m["bound"] == "tensor" and m["flops_per_unit"] == 2 * (32 * 64 + 64 * 16)
    assert bench.kernel_roofline("tn_render_fwd", 6, 0.1, 0, peaks) is None
    table = {"tn_hash_encode_bwd[L16,T2^19,dx]": (4, 0.4, 4 * n), "tn_prop_density_bwd[L5,S256,dx]": (2, 0.5, 2 * 4096 * 256),
             "tn_render_fwd": (6, 0.9, 0)}
    roofs = bench.top_rooflines(table, peaks, l2, False)
    assert roofs[0]["kernel"].startswith("tn_prop_density_bwd") and len(roofs) == 2
