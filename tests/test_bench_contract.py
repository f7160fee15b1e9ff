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


def test_reference_arm_is_independent_of_the_product_library():
    """The CPU arm orders, analyses and factorises with oracle code only: libgmrfb.so must not be mapped by it, and it
    uses all host cores even when the launcher exported OMP_NUM_THREADS=1 (torchrun does)."""
    code = ("import sys, runpy, json\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--nx', '41', '--steps', '1', '--warmup', '1']\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "print('MAPS', 'libgmrfb' in maps, 'libsupernodal' in maps)\n") % os.path.join(ROOT, "bench.py")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "MAPS False True" in out.stdout
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)


def test_both_arms_print_the_same_config():
    """`config` describes the workload only; arm-specific facts (nnz(L), ordering, lanes) live under `detail`."""
    sys.path.insert(0, ROOT)
    import bench

    c = bench.workload_config(1001, 1002001, 18997999)
    assert set(c) == {"workload", "n", "nnz_Q", "obs_frac", "q_eps", "corr_range", "l2_policy"}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(') == 2  # the reference line and the product line


def test_traffic_is_only_reported_for_a_matching_capture(tmp_path, monkeypatch):
    """roofline.traffic comes from a committed ncu pass that records its own configuration; a capture of another mesh,
    ordering or launch count is not reported (stale DRAM bytes next to fresh TFLOP/s)."""
    sys.path.insert(0, ROOT)
    import bench

    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "r02_gemm_traffic.json").write_text(json.dumps({
        "config": {"nx": 1001, "ordering": "nd", "nd_cover": "1"},
        "kernels": {"k_gemm<NT> (DMMA)": {"launches": 141, "bytes_per_launch": 1.0e8}}}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.delenv("GMRFB_ND_COVER", raising=False)
    t, src = bench.ncu_traffic("k_gemm<NT> (DMMA)", 1001, "nd", 141)
    assert t == 1.0e8 and "r02_gemm_traffic.json" in src
    assert bench.ncu_traffic("k_gemm<NT> (DMMA)", 1001, "nd", 157)[0] is None     # other launch mix
    assert bench.ncu_traffic("k_gemm<NT> (DMMA)", 601, "nd", 141)[0] is None      # other mesh
    assert bench.ncu_traffic("k_gemm<NT> (DMMA)", 1001, "amd", 141)[0] is None    # other ordering
    assert bench.ncu_traffic("no such kernel", 1001, "nd", 141)[0] is None
