"""The product arm of bench.py at 1/512 of the workload: the JSON line carries every key of the contract and the
numbers hang together (roofline fraction = achieved / peak, e2e moves the bytes it says, launches were counted)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_small(built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--scale", "0.125", "--steps", "4", "--warmup", "3",
                          "--settle", "2", "--e2e-steps", "1"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "clocks", "gpu_launches", "e2e", "cpu_baseline"):
        assert k in d, k
    assert "impl" not in d and d["dtype"] == "f32" and d["steps"] == 4 and d["warmup"] == 3
    cfg = d["config"]
    npart = cfg["particles_per_gpu"]
    assert npart == 32 ** 3 * 8 and abs(d["value"] - npart / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    assert r["kernel_ms"] <= d["ms_per_step"]
    assert d["gpu_launches"] >= 4 * 8                      # at least 8 kernels of ours per step
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= npart * 40 and e["d2h_bytes_per_step"] >= npart * 40
    assert e["value"] < d["value"]                          # host arrays in and out every step cannot be faster
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] and cb["cores"] >= 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
