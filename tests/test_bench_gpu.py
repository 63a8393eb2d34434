"""bench.py end to end on a small batch: the JSON line carries every key of the contract (the driver runs the full-size
line after the tests; this catches a broken bench before it costs the round's numbers)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload,extra", [("2d_static_dense", []), ("3d_dynamic_dense", ["--mode", "step"]),
                                            ("1d_dynamic", ["--timing", "stream"])])
def test_bench_line_contract_small(workload, extra):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--envs", "16384", "--steps", "6", "--warmup", "3",
           "--no-workloads", "--no-cpu-baseline", "--e2e-steps", "4"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "repeats", "region_ms", "pcie"):
        assert k in d, k
    assert d["metric"] == "env-steps/sec" and d["steps"] == 6 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["repeats"] >= 30
    assert d["value"] > 0 and abs(d["value"] - 16384 * 6 / (d["region_ms"]["median"] * 1e-3)) < 1e-6 * d["value"]
    assert abs(d["ms_per_step"] * 6 - d["region_ms"]["median"]) < 1e-9
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["config"]["workload"] == workload and d["gpu_launches"] == d["gpu_launches_per_region"] * d["repeats"]
    for k in ("e2e", "e2e_f32", "e2e_i16", "e2e_record", "e2e_record_mapped") + (() if workload.startswith("1d") else ("e2e_bits_staged",)):
        assert d[k]["value"] > 0 and d[k]["d2h_bytes_per_step"] > 0 and d[k]["h2d_bytes_per_step"] == 16384
    assert d["e2e"]["d2h_bytes_per_step"] == 16384 * {"1": 16, "2": 16, "3": 32}[workload[0]]
    assert d["e2e_record"]["d2h_bytes_per_step"] == 16384 * (16 if workload.startswith("1d") else 56)
    assert d["other_mode"]["value"] > 0 and d["clocks"]["samples"] >= 1
