"""The CPU arm of bench.py (`--impl reference`: the oracle restatement of helper_functions.py:166-176 timed on the host
cores) at a tiny size: it must run without a GPU and print ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "192", "--N", "8",
           "--steps", "3", "--warmup", "1", "--ref-budget-s", "30"] + extra
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run([])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "node_prune_steps_per_sec" and d["higher_is_better"] is True
    assert d["unit"] == "node-prune-steps/s" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] >= 2 and d["measured"] is True and d["extrapolated"] is False
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-9 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"]
    assert set(cb["per_kernel"]) == {"linear", "quadratic"}           # Gaussian: the reference cannot run the config
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["n_samples"] == 192 and d["config"]["n_nodes"] == 8 and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run([], env={"RANK": "1", "WORLD_SIZE": "2"}) == []
