#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native GMRF linear-algebra hot path.

Metric (BASELINE.json): GMRF posterior (mean + marginal variance) solves/sec.
One *step* = one posterior solve of BASELINE config 4 (synthetic 2-D Matern SPDE GMRF on a ~1M-node P1 mesh,
10 % of nodes observed): numeric supernodal Cholesky of Q_post (symbolic analysis reused, as the reference does
with `perm=p`) + posterior mean (forward + backward solve) + marginal variances by Takahashi selected inversion.

    python bench.py --gpus N --steps K --warmup W          # product arm (CUDA, one process per GPU)
    python bench.py --impl reference ...                    # CPU arm: the oracle port on the host cores

`value`   : whole-job solves/s with the precision values and right-hand side resident in HBM (device entry points).
`e2e`     : the same metric through the host C-ABI calls (pinned host buffers, H2D/D2H inside the timed region).
`roofline`: the dominant kernel of the step, timed with CUDA events inside this run (profiled replay of the same
            step), against the measured FP64 tensor peak (cuBLAS DGEMM, profiles/r01_fp64_probe.json) or the
            measured HBM copy bandwidth (MEASURED_PEAKS.json).
N > 1     : independent posterior problems (the reference's dataset loop, scripts/darcy/solve_darcy_gmrf-fem.jl:210)
            are sharded one per rank — no data-path collective; weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "GMRF posterior (mean+marginal var) solves/sec"
UNIT = "solves/s"
FP64_PEAK_TFLOPS_FALLBACK = 35.5  # cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/r01_fp64_probe.json)


def load_peaks():
    hbm, fp64, src = 6650.0, FP64_PEAK_TFLOPS_FALLBACK, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f)["hbm_gbs"])
            src = "measured"
    except Exception:  # noqa: BLE001
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_probe.json")) as f:
            fp64 = float(json.load(f)["dgemm_8192_tflops"])
    except Exception:  # noqa: BLE001
        pass
    return hbm, fp64, src


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower() == "active":
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def build_problem(nx, seed):
    pkg = entry.load_pkg()
    prob = pkg.workloads.matern_posterior(nx, obs_frac=0.1, q_eps=1e2, corr_range=0.05, seed=seed)
    return prob


# --------------------------------------------------------------------------------------------- CPU arm ----
class CpuPosterior:
    """The reference path on the host cores: supernodal multifrontal Cholesky with BLAS-3 supernodes
    (oracle/supernodal_chol.c, CHOLMOD's algorithm class; OpenBLAS through scipy, all host threads) on the SAME
    workload, ordering and supernode partition as the GPU arm.  One step = numeric factorisation (symbolic reused,
    as `perm=p` does in the reference) + posterior mean (forward + backward sweep) + Takahashi variances."""

    def __init__(self, nx, seed=0):
        pkg = entry.load_pkg()
        orc = entry.load_oracle()
        self.prob = build_problem(nx, seed)
        Qp = self.prob["Qpost"]
        sym = pkg.Symbolic(Qp, coords=self.prob["nodes"], host_only=True)  # integer analysis only, no GPU involved
        p, ipost = sym.p, sym.ipost
        perm_int = np.empty(len(p), np.int64)
        perm_int[ipost] = p
        sptr = sym.super_ptr
        rows = [sym.super_rows(s) for s in range(len(sptr) - 1)]
        self.info = sym.info
        self.F = orc.SupernodalCholesky(Qp, perm_int, sptr, rows)
        self.n = Qp.shape[0]

    def step(self):
        t0 = time.perf_counter()
        self.F.refactor(self.prob["Qpost"].data)
        t1 = time.perf_counter()
        x = self.F.solve(self.prob["rhs"])
        t2 = time.perf_counter()
        v = self.F.selinv_diag()
        t3 = time.perf_counter()
        return dict(factor_s=t1 - t0, solve_s=t2 - t1, selinv_s=t3 - t2, total_s=t3 - t0), x, v


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count() or 1])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def cpu_sample_desc(nx, s):
    return (f"full workload: one posterior solve of the {nx}x{nx} mesh (n={nx * nx}) by oracle/supernodal_chol.c "
            f"(supernodal multifrontal, OpenBLAS): factor {s['factor_s']:.2f}s, solves {s['solve_s']:.2f}s, "
            f"selinv {s['selinv_s']:.2f}s; a restatement of CHOLMOD's supernodal algorithm class, not CHOLMOD")


def cpu_baseline(nx):
    cp = CpuPosterior(nx)
    cp.step()  # warm-up call, as the reference does before each timed call
    s, _, _ = cp.step()
    return {"value": 1.0 / s["total_s"], "unit": UNIT, "cores": cpu_threads(), "kind": "port",
            "sample": cpu_sample_desc(nx, s), "sample_seconds": s["total_s"]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx = args.nx
    cp = CpuPosterior(nx)
    times, last = [], None
    for it in range(args.warmup + args.steps):
        s, x, v = cp.step()
        if it >= args.warmup:
            times.append(s["total_s"])
            last = s
    per = sum(times) / len(times)
    resid = float(np.linalg.norm(cp.prob["Qpost"] @ x - cp.prob["rhs"]) / np.linalg.norm(cp.prob["rhs"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": 1.0 / per, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config 4: 2-D Matern SPDE GMRF posterior, {nx}x{nx} P1 mesh (n={nx * nx}), "
                               "numeric supernodal Cholesky + mean + selected-inversion variances per step",
                   "n": nx * nx, "nnz_L": int(cp.info.nnz_L), "factor_flops": cp.info.flops,
                   "note": "CPU arm = oracle port on the host cores (CHOLMOD/Julia are not installed in this image); "
                           "full workload per step, no sampling"},
        "cpu_baseline": {"value": 1.0 / per, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                         "sample": cpu_sample_desc(nx, last)},
        "e2e": {"value": 1.0 / per, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "parity_check": {"mean_residual": resid, "var_positive": bool(np.all(v > 0))},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm ----
class Lane:
    """One posterior problem in flight on the GPU: its own context (CUDA stream), factor handle and buffers."""

    def __init__(self, pkg, torch, dev, local, nx, seed, perm=None, ordering="nd"):
        self.prob = build_problem(nx, seed)
        Qp = self.prob["Qpost"]
        self.n = Qp.shape[0]
        self.ctx = pkg.Context(local)
        self.ext = torch.cuda.ExternalStream(self.ctx.stream, device=dev)
        # the first lane orders the pattern (library nested dissection); the others reuse its permutation, exactly as
        # the reference passes `perm=p` for every further problem (scripts/darcy/solve_darcy_gmrf-fem.jl:169,174)
        if perm is None:
            self.sym = pkg.Symbolic(Qp, coords=self.prob["nodes"] if ordering == "nd" else None,
                                    ordering={"ndgraph": "nd"}.get(ordering, ordering), ctx=self.ctx)
        else:
            self.sym = pkg.Symbolic(Qp, perm=perm, ctx=self.ctx)
        self.fac = pkg.CholeskyFactor(self.sym)
        self.nz_host = torch.from_numpy(np.ascontiguousarray(Qp.data)).pin_memory()
        self.rhs_host = torch.from_numpy(np.ascontiguousarray(self.prob["rhs"])).pin_memory()
        self.d_nz = self.nz_host.to(dev)
        self.d_rhs = self.rhs_host.to(dev)
        self.d_x = torch.empty_like(self.d_rhs)
        self.d_var = torch.empty_like(self.d_rhs)
        self.x_host = np.empty(self.n)
        self.end = torch.cuda.Event(enable_timing=True)
        self.torch, self.local = torch, local
        self.xs = self.v = None

    def steps_device(self, k):
        """k posterior solves, everything resident in HBM: numeric factor + mean + selected-inversion variances."""
        torch = self.torch
        torch.cuda.set_device(self.local)
        for _ in range(k):
            with torch.cuda.stream(self.ext):
                self.d_x.copy_(self.d_rhs, non_blocking=True)
            self.fac.factorize_dev(self.d_nz.data_ptr())
            self.fac.solve_dev(self.d_x.data_ptr(), 1)
            self.fac.var_selinv_dev(self.d_var.data_ptr())

    def steps_e2e(self, k):
        """The same through the host C-ABI calls: pinned host inputs, host outputs (H2D/D2H inside)."""
        self.torch.cuda.set_device(self.local)
        for _ in range(k):
            self.fac.factorize(self.nz_host.numpy())
            self.x_host[:] = self.rhs_host.numpy()
            self.xs = self.fac.solve(self.x_host)
            self.v = self.fac.var_selinv()


def run_gpu_arm(args):
    import threading

    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg = entry.load_pkg()
    hbm_peak, fp64_peak, peak_src = load_peaks()

    nx, B = args.nx, max(1, args.inflight)
    t_setup = time.perf_counter()
    # rank r, lane b solves problem r*B + b: independent posterior problems on one sparsity pattern
    lanes = [Lane(pkg, torch, dev, local, nx, rank * B, ordering=args.ordering)]
    for b in range(1, B):
        lanes.append(Lane(pkg, torch, dev, local, nx, rank * B + b, perm=lanes[0].sym.p))
    t_setup = time.perf_counter() - t_setup
    L0 = lanes[0]
    n, Qp, info = L0.n, L0.prob["Qpost"], L0.sym.info
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(which, k, use):
        """Run k steps on each lane of `use` concurrently (one host thread per lane); device time from a start event
        every lane's stream waits on to an end event that waits on every lane's stream."""
        barrier()
        cur = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for L in use:
            L.ext.wait_event(e0)
        th = [threading.Thread(target=getattr(L, which), args=(k,)) for L in use]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for L in use:
            L.end.record(L.ext)
            cur.wait_event(L.end)
        e1.record(cur)
        barrier()
        return e0.elapsed_time(e1)

    timed("steps_device", args.warmup, lanes)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sum(L.ctx.launch_count for L in lanes)
    ms = timed("steps_device", args.steps, lanes)
    launches = sum(L.ctx.launch_count for L in lanes) - l0
    clocks = sampler.stop() if rank == 0 else None
    # latency of a single posterior solve with nothing else in flight
    ms_single = timed("steps_device", args.steps, lanes[:1]) / args.steps

    # end-to-end through the host C-ABI calls (pinned host inputs, host outputs)
    ke = max(1, min(args.steps, 3))
    timed("steps_e2e", 1, lanes)
    ms_e2e = timed("steps_e2e", ke, lanes) / ke

    if dist is not None:
        t = torch.tensor([ms, ms_e2e, ms_single], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_single = float(t[0]), float(t[1]), float(t[2])

    # parity spot check of what was timed (size-independent property: residual of the mean, and var > 0), every lane
    ok, resid, varpos = True, 0.0, True
    for L in lanes:
        mean = L.d_x.cpu().numpy()
        r = float(np.linalg.norm(L.prob["Qpost"] @ mean - L.prob["rhs"]) / np.linalg.norm(L.prob["rhs"]))
        var = L.d_var.cpu().numpy()
        resid = max(resid, r)
        varpos = varpos and bool(np.all(var > 0))
        same = float(np.max(np.abs(L.xs - mean))) <= 1e-12 * float(np.max(np.abs(mean))) + 1e-300
        same = same and float(np.max(np.abs(L.v - var))) <= 1e-12 * float(np.max(np.abs(var)))
        ok = ok and r < 1e-9 and varpos and same

    # per-kernel profile of one more step (CUDA events around every launch on the library's stream)
    roofline = None
    prof_rows = []
    if rank == 0:
        ctx = L0.ctx
        # cudaProfilerStart/Stop bracket exactly this replayed step: `ncu --profile-from-start off` captures the
        # launches of one posterior solve without any launch counting (no effect outside a profiler)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        ctx.profile_begin()
        L0.steps_device(1)
        prof = ctx.profile_end()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        tot = sum(p["ms"] for p in prof)
        for p in sorted(prof, key=lambda q: -q["ms"]):
            row = dict(name=p["name"], launches=p["launches"], ms=round(p["ms"], 3), share=round(p["ms"] / tot, 4))
            if p["flops"] > 0 and p["ms"] > 0:
                row["tflops"] = round(p["flops"] / p["ms"] * 1e-9, 3)
            if p["bytes"] > 0 and p["ms"] > 0:
                row["gbs"] = round(p["bytes"] / p["ms"] * 1e-6, 1)
            prof_rows.append(row)
        top = max(prof, key=lambda q: q["ms"])
        traffic = ncu_traffic(top["name"])
        if top["flops"] > 0 and "gemm" in top["name"]:
            ach = top["flops"] / top["ms"] * 1e-9
            roofline = {"bound": "tensor", "kernel": top["name"], "achieved": ach, "peak": fp64_peak,
                        "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic,
                        "peak_source": "cuBLAS DGEMM 8192^3 measured on this pool (profiles/r01_fp64_probe.json); "
                                       "FP64 tensor (DMMA) issue peak 37.1 TFLOP/s; no f64 kind exists for tcgen05",
                        "launches_per_step": top["launches"], "ms_per_step": top["ms"],
                        "flops_per_step": top["flops"],
                        "note": "achieved = algorithmic flops of all launches of this kernel in one posterior solve / "
                                "their summed CUDA-event durations (profiled replay of the same step, one problem in flight)"}
        else:
            ach = top["bytes"] / top["ms"] * 1e-6 if top["ms"] > 0 else 0.0
            roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({peak_src})",
                        "launches_per_step": top["launches"], "ms_per_step": top["ms"]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(nx)
    per_step = ms / args.steps
    line = {
        "metric": METRIC, "value": world * B * 1e3 / per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config 4: 2-D Matern SPDE GMRF posterior, {nx}x{nx} P1 mesh (n={n}), numeric "
                               "supernodal Cholesky + mean + selected-inversion variances per posterior solve; "
                               f"one step = one batch of {B} independent posterior problems per GPU (same pattern, "
                               "different values; the reference's dataset loop), each on its own CUDA stream",
                   "n": n, "nnz_Q": int(Qp.nnz), "nnz_L": int(info.nnz_L), "factor_flops": info.flops,
                   "nsuper": int(info.nsuper), "levels": int(info.nlevels), "max_front": int(info.max_front),
                   "front_arena_gb": info.front_bytes / 1e9, "ordering": {"nd": "library nested dissection (geometric, minimum-vertex-cover separators"
                                      + (")" if os.environ.get("GMRFB_ND_COVER", "1") != "0" else " off: plain boundary layers)"),
                                "ndgraph": "library nested dissection (graph bisection, no coordinates)",
                                "nd_amd": "library nested dissection (graph) with halo-AMD leaves",
                                "amd": "library approximate minimum degree"}[args.ordering],
                   "problems_per_gpu_in_flight": B, "solves_per_step": B,
                   "single_solve_latency_ms": ms_single,
                   "l2_policy": "working set (front arenas, 20 GB per problem) >> 126 MB L2; no flush needed",
                   "setup_s_outside_timing": round(t_setup, 2)},
        "e2e": {"value": world * B * 1e3 / ms_e2e, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(B * (L0.nz_host.numel() * 8 + n * 8)), "d2h_bytes_per_step": int(B * 2 * n * 8)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernel_profile": prof_rows,
        "parity_check": {"mean_residual": resid, "var_positive": varpos, "ok": bool(ok)},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if not ok:
        sys.exit("bench: parity spot check failed")


def ncu_traffic(kernel_name):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernel, averaged over all its
    launches of one posterior solve, from the committed ncu pass over the same bench step
    (profiles/r01_gemm_traffic.json, written by tools/summarize_traffic.py from tools/gpu_evidence.sh); None when
    no capture of that kernel is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")) as f:
            d = json.load(f)
        e = d["kernels"].get(kernel_name)
        return e["bytes_per_launch"] if e else None
    except Exception:  # noqa: BLE001
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=1001, help="mesh nodes per side (1001 -> 1,002,001 nodes)")
    ap.add_argument("--nx-sample", dest="nx_sample", type=int, default=0, help="(unused; kept for old command lines)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ordering", choices=["nd", "ndgraph", "nd_amd", "amd"], default="nd",
                    help="fill-reducing ordering computed once by the library and reused as perm=p (default: geometric "
                         "nested dissection, the configuration every committed profile was measured with)")
    ap.add_argument("--inflight", type=int, default=4,
                    help="independent posterior problems in flight per GPU (one CUDA stream each)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
