"""bench.py contract checks that need no GPU: the reference arm (CPU restatement of the path, the one place outside
tests/ where the oracle may be executed) prints ONE JSON line with the keys the driver reads, rank 0 only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nx", "41",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("GMRF posterior (mean+marginal var) solves/sec")
    assert d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["parity_check"]["mean_residual"] < 1e-10 and d["parity_check"]["var_positive"]


def test_reference_arm_other_ranks_exit_quietly():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_committed_traffic_summary_feeds_the_roofline():
    """roofline.traffic of the GPU arm comes from the committed ncu pass (profiles/r01_gemm_traffic.json)."""
    sys.path.insert(0, ROOT)
    import bench

    t = bench.ncu_traffic("k_gemm<NT> (DMMA)")
    assert t is not None and t > 1e6
    assert bench.ncu_traffic("no such kernel") is None
