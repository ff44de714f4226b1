"""The reference arm of bench.py (the oracle port on the host cores) prints the JSON line of the bench contract;
checked on a coarse grid so that it runs in seconds.  The GPU arm is exercised on the GPU box by the driver."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout.strip().splitlines()


def test_reference_arm_json_line():
    lines = _run("--impl", "reference", "--nlat", "46", "--nlon", "90", "--steps", "1", "--warmup", "1")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["higher_is_better"] is True and d["unit"] == "time steps/s" and d["value"] > 0
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None
    cpu = d["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] == d["value"] and cpu["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["config"]) == {"workload", "l2", "parallelism"}  # no per-arm keys: both arms print the same config


def test_reference_arm_other_ranks_exit_quietly():
    """under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output"""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    assert _run("--impl", "reference", "--gpus", "2", "--nlat", "46", "--nlon", "90", "--steps", "1", env=env) == []
