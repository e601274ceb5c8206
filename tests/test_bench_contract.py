"""bench.py prints ONE JSON line with the keys the driver reads (contract in the task statement)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    d = _run(["--impl", "reference", "--rows", "512", "--steps", "2", "--warmup", "1", "--ref-pivots", "4"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "pivot_update_hbm_GBps" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "pivots_per_s": d["pivots_per_s"],
                        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - d["pivots_per_s"] * 16 * 512 * 512 / 1e9) < 1e-6 * d["value"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--rows", "256",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--rows", "4096", "--pivots", "32", "--steps", "2", "--warmup", "3", "--no-secondary", "--cpu-pivots", "4"])
    assert BASE_KEYS | {"roofline", "clocks", "pivots_per_s", "lookahead"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["metric"] == "pivot_update_hbm_GBps"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["algorithmic_bytes_per_launch"] == 16 * 4096 * 4096
    assert d["e2e"]["h2d_bytes_per_step"] > 8 * 4095 * 4095 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] >= 3 * 32 * 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["lookahead"]["K=8"]["pivots_per_s"] > d["pivots_per_s"]
