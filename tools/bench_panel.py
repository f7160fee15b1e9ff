#!/usr/bin/env python
"""Panel (multi-right-hand-side) solves: time and HBM fraction of batched solves / RBMC-50 on configs 3 and 4.

    python tools/bench_panel.py [--nx 601] [--nrhs 50 64] [--out profiles/rXX_panel.json]

Bytes per sweep pair: 2 * (8 nnz(L) + 16 n nrhs).  Device-resident timing (CUDA events on the library's
stream through torch's ExternalStream)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, nargs="+", default=[601])
    ap.add_argument("--nrhs", type=int, nargs="+", default=[1, 4, 50, 64])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--dump", default="")
    args = ap.parse_args()
    import torch

    pkg = entry.load_pkg()
    W = pkg.workloads
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        hbm = 6536.7
    dev = torch.device("cuda", 0)
    ctx = pkg.Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    res = []
    for nx in args.nx:
        prob = W.matern_posterior(nx, obs_frac=0.1, q_eps=1e2, corr_range=0.05, seed=0)
        Q = prob["Qpost"]
        n = Q.shape[0]
        sym = pkg.Symbolic(Q, coords=prob["nodes"], ctx=ctx)
        fac = pkg.CholeskyFactor(sym).factorize(Q.data)
        nnzL = int(sym.info.nnz_L_stored)
        Qd = pkg.SparseMatrix(Q, ctx=ctx)
        for nrhs in args.nrhs:
            X = torch.randn((nrhs, n), dtype=torch.float64, device=dev)  # column-major n x nrhs
            X0 = X.clone()

            def run(mode):
                with torch.cuda.stream(ext):
                    X.copy_(X0, non_blocking=True)
                fac.solve_dev(X.data_ptr(), nrhs, mode=mode)

            out = {"nx": nx, "n": n, "nnz_L_stored": nnzL, "nrhs": nrhs}
            for name, mode, sweeps in (("solve_A", pkg._lib.SOLVE_A, 2), ("sample_UP", pkg._lib.SOLVE_UP, 1)):
                run(mode)
                ctx.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ext)
                for _ in range(args.reps):
                    run(mode)
                e1.record(ext)
                ctx.sync()
                ms = e0.elapsed_time(e1) / args.reps
                byts = sweeps * (8.0 * nnzL + 16.0 * n * nrhs)
                out[name] = {"ms": ms, "gbs": byts / ms * 1e-6, "hbm_frac": byts / ms * 1e-6 / hbm,
                             "ms_per_rhs": ms / nrhs}
            # parity spot check: residual of the batch
            run(pkg._lib.SOLVE_A)
            ctx.sync()
            Xh = X.cpu().numpy().T
            B = X0.cpu().numpy().T
            out["residual"] = float(np.linalg.norm(Q @ Xh - B) / np.linalg.norm(B))
            res.append(out)
            print(json.dumps(out), flush=True)
        # RBMC-50 as the host interface runs it (normals drawn on the device)
        Z = torch.randn((50, n), dtype=torch.float64, device=dev)
        fac.var_rbmc(Qd, Z)
        t = time.perf_counter()
        for _ in range(args.reps):
            v = fac.var_rbmc(Qd, Z)
        dt = (time.perf_counter() - t) / args.reps
        vex = fac.var_selinv()
        out = {"nx": nx, "rbmc50_ms": dt * 1e3, "median_rel_err_vs_takahashi": float(np.median(np.abs(v - vex) / vex))}
        if args.dump:
            os.environ["GMRFB_PROFILE_DUMP"] = args.dump + f".rbmc_{nx}.csv"
            ctx.profile_begin()
            fac.var_rbmc(Qd, Z)
            prof = ctx.profile_end()
            out["rbmc50_profile"] = sorted(([p["name"], p["launches"], round(p["ms"], 3)] for p in prof), key=lambda r: -r[2])
        res.append(out)
        print(json.dumps(out), flush=True)
        del fac, sym
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
