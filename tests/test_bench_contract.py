"""bench.py contract checks that need no GPU: the reference arm (CPU port of the reference algorithm) prints one JSON
line with the agreed keys, and under torchrun only rank 0 runs it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CELL = "sweep:2:0.2"  # config-2 shape, 2 barcodes, 0.2x coverage: ~40 k reads per iteration, a second of CPU work


def run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", CELL,
                           "--steps", "1", "--warmup", "1", "--gpus", "2" if "RANK" in env_extra else "1"],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line():
    p = run({})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("simulated reads placed per second") and d["value"] > 0 and d["iterations_per_s"] > 0
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    p = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""
