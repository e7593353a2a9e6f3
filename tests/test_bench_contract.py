"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout
with the keys the driver reads (workloads shrunk through the PCR_BENCH_TEST_N test hook)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(extra_env, *args):
    env = dict(os.environ, PCR_BENCH_TEST_N="30000", REF_BUDGET_S="2", **extra_env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = run_reference({}, "--steps", "3", "--warmup", "3")
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ICP iterations/sec" and d["unit"] == "iterations/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["extrapolated"] is False and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["measured"]["scan_points"] == d["config"]["scan_points_measured"] == 30000
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c2") and d["vs_baseline"] is None


def test_reference_arm_under_torchrun_only_rank0_prints():
    out = run_reference({"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "3", "--warmup", "3")
    assert out.strip() == ""
    out = run_reference({"WORLD_SIZE": "2", "RANK": "0", "LOCAL_RANK": "0"}, "--gpus", "2", "--steps", "3", "--warmup", "3")
    d = json.loads(out.strip())
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["config"]["workload"].startswith("c5")
    # a bounded sample is labelled as such: ms_per_step is the MEASURED sample time, value the extrapolation
    assert d["extrapolated"] == (d["measured"]["scan_points"] < d["config"]["scan_points"] or d["measured"]["target_points"] < d["config"]["target_points"])
    assert abs(d["ms_per_step"] - d["measured"]["ms_per_step"]) < 1e-9
