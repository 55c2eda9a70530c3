"""bench.py's JSON contract, checked where it can run without a GPU: the reference arm (`--impl reference`, the CPU
restatement timed on a bounded sample) must print ONE JSON line with the agreed keys; ranks other than 0 print
nothing and exit 0; the GPU arm's argument surface is the driver's (`--gpus/--steps/--warmup`)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, env=env,
                          cwd=str(ROOT), timeout=300)


def test_reference_arm_prints_the_contract_line():
    r = _run(None, "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--cpu-sample-rows", "20000")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["metric"] == json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 1 and j["vs_baseline"] is None
    # ms_per_step is the step that was really timed (one search over the 20000-row sample); `value` is that rate
    # scaled linearly to the 100M rows of the metric
    assert j["value"] > 0 and abs(j["ms_per_step_scaled_to_all_rows"] * j["value"] - 1e3) < 1e-6 * 1e3
    assert abs(j["ms_per_step"] * 100_000_000 / 20_000 - j["ms_per_step_scaled_to_all_rows"]) < 1e-6 * j["ms_per_step_scaled_to_all_rows"]
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == j["value"] and "20000x512" in cb["sample"]
    assert cb["optimistic_all_threads_cores"] >= 1 and cb["optimistic_all_threads_value"] > 0
    assert j["config"]["workload"].startswith("100000000x512") and j["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
             "--warmup", "1", "--cpu-sample-rows", "20000")
    assert r.returncode == 0 and r.stdout.strip() == ""
