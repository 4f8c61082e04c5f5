"""bench.py contract (no GPU): the reference arm runs on the host cores alone and prints ONE JSON line with the
keys the driver reads; rank != 0 prints nothing; the config table names BASELINE.json's workloads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*argv, env=None):
    e = dict(os.environ, FENIX_BENCH_BUDGET_S="1.0", **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          env=e, timeout=300, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    res = run_bench("--impl", "reference", "--config", "c2", "--rows", "70000", "--steps", "1", "--warmup", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in out, key
    assert out["impl"] == "reference" and out["metric"] == "knn_qps" and out["unit"] == "queries/s"
    assert out["vs_baseline"] is None and out["higher_is_better"] is True
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1 and out["cpu_baseline"]["value"] > 0
    assert out["cpu_baseline"]["best_case"]["value"] > out["cpu_baseline"]["value"]
    assert out["e2e"] == {"value": out["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in out["config"] and "model" not in out["config"]


def test_reference_arm_other_ranks_stay_silent():
    res = run_bench("--impl", "reference", "--gpus", "2", "--config", "c2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_config_table_covers_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.CONFIGS["c2"]["n"] == 1_000_000 and bench.CONFIGS["c2"]["d"] == 128 and bench.CONFIGS["c2"]["k"] == 100
    assert bench.CONFIGS["c3"]["n"] == 10_000_000 and bench.CONFIGS["c3"]["d"] == 768 and bench.CONFIGS["c3"]["q"] == 4096
    assert bench.CONFIGS["c4"]["n"] == 100_000_000 and bench.CONFIGS["c4"]["d"] == 96
    assert {f"c5_{b}" for b in (1, 2, 4, 8, 16, 32, 64)} <= set(bench.CONFIGS)
