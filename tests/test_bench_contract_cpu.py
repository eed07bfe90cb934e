"""bench.py's reference arm runs on CPU: its JSON line carries the contract's keys, the same `config` dictionary the
GPU arm prints (the driver compares them), and a CPU baseline that names what was sampled.  The GPU arm itself is
checked where there is a GPU (tests/test_gpu_bench.py)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line(built):
    d = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--scale", "0.125", "--settle", "1")
    assert d["impl"] == "reference" and d["metric"].startswith("particle-steps/s") and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "cells" in cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cfg = d["config"]
    assert cfg["workload"] == "uniform3d" and "model" not in cfg and "load" in cfg
    assert cfg["particles_per_gpu"] == cfg["ppc"] * cfg["cells_per_gpu"][0] * cfg["cells_per_gpu"][1] * cfg["cells_per_gpu"][2]


def test_reference_arm_other_ranks_exit_quietly(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device(built):
    """no CPU fallback: without a CUDA device the product arm stops with a message, it does not time the oracle"""
    import ctypes
    from osiris_b200 import lib
    if lib.load().osgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--scale", "0.125"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
