"""Host-side logic of bench.py that needs no GPU: the launch plan that times EXACTLY K steps, the workload table and
the reference arm's JSON contract on a tiny sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("K,KL", [(20, 64), (20, 16), (8192, 1), (100, 16), (3, 16), (1, 1), (4097, 64), (64, 64), (65, 64)])
def test_launch_plan_times_exactly_k_steps(K, KL):
    launches = bench.plan_launches(K, KL)
    assert sum(launches) == K and all(0 < k <= KL for k in launches)
    assert len(launches) == -(-K // KL) and all(k == KL for k in launches[:-1])
    if K <= KL:
        assert launches == [K]                       # the driver's --steps 20: ONE dmp_rollout launch per timed region


def test_workload_table_matches_baseline_configs():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert "env-steps/sec" in base["metric"] and bench.METRIC == "env-steps/sec"
    w = bench.WORKLOADS
    assert w["1d_dynamic"][4] == 65536 and w["2d_static_dense"][4] == 1048576 and w["3d_static_dense"][4] == 262144
    # algorithmic bytes per env-step, fp32 observations, step mode
    assert (w["1d_static_step"][5], w["1d_dynamic"][5]) == (64, 67)
    assert (w["2d_static_dense"][5], w["2d_dynamic_dense"][5]) == (262, 268)
    assert (w["3d_static_dense"][5], w["3d_dynamic_dense"][5]) == (330, 336)


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the staged reference classes -- the oracle port where nothing is staged -- on the
    host cores) on a tiny sample: one JSON line with the contract's keys; non-zero ranks print nothing."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "30", "--warmup", "3",
                          "--workload", "2d_static_dense", "--ref-envs-per-proc", "8"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["higher_is_better"] is True
    assert d["steps"] == 30 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    if d["cpu_baseline"]["kind"] == "reference":
        assert "UNMODIFIED reference classes" in d["cpu_baseline"]["sample"] and d["cpu_baseline_port"]["value"] > 0
    assert d["config"]["workload"] == "2d_static_dense" and d["config"]["total_envs"] == 1048576
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "30", "--warmup", "3"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_byte_formulas_follow_survey_8d():
    """B_alg per env-step (fp32 observations), and the same with this repository's 3D nibble layout."""
    B = bench.Bench
    assert B.b_alg(2, 262, 51, 1, True) == 261 and B.b_alg(2, 262, 51, 1, False) == 262
    assert B.b_alg(2, 262, 51, 20, True) == 204 + 5 + (12 + 8 + 100) / 20 == 215.0
    assert B.b_alg(1, 67, 7, 64, True) == 28 + 5 + (10 + 8 + 120) / 64
    assert B.b_alg(3, 330, 51, 1, True, mean_len=25.0) == 329 + 64.0
    assert B.b_alg(3, 330, 51, 20, True, mean_len=25.0) == 204 + 5 + (12 + 8 + 1600) / 20 + 64.0
    assert B.b_layout(2, 262, 51, 20, True) == B.b_alg(2, 262, 51, 20, True)
    assert B.b_layout(3, 330, 51, 1, True, mean_len=26.0) == 330 - 100 + 25.5 - 1 + 16.0
    assert B.b_layout(3, 330, 51, 20, True, mean_len=26.0) == 204 + 5 + (12 + 8 + 416) / 20 + 16.0
