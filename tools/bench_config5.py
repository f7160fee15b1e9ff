#!/usr/bin/env python
"""Config 5 through the sparse path on one GPU: the 256 x 256-node heat space-time precision (implicit Euler,
src/spdes/shallow_water.jl:198-230) for the largest numbers of time steps whose supernodal factor fits one B200, ordered
by nested dissection in space-time (3-D coordinates).  Factor + posterior-mean solve + marginal variances (Takahashi
selected inversion while the second arena fits, RBMC-50 panel sweeps otherwise).

    python tools/bench_config5.py --nx 256 --steps 8 16 --out profiles/r02_config5_sparse.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=256)
ap.add_argument("--steps", type=int, nargs="+", default=[8, 16])
ap.add_argument("--out", default="")
ap.add_argument("--selinv-max-gb", type=float, default=80.0, help="run the selected inversion only below this arena size")
args = ap.parse_args()
import torch  # noqa: E402

pkg = entry.load_pkg()
W = pkg.workloads
ctx = pkg.Context(0)
dev = torch.device("cuda", 0)
res = []
for N in args.steps:
    st = W.heat_spacetime_sparse(args.nx, N)
    A = st["A"]
    n = A.shape[0]
    t = time.perf_counter()
    sym = pkg.Symbolic(A, coords=st["coords"], ctx=ctx)
    t_an = time.perf_counter() - t
    info = sym.info
    out = {"nx": args.nx, "n_steps": N, "n": n, "nnz_A": int(A.nnz), "nnz_L": int(info.nnz_L), "factor_flops": info.flops,
           "max_front": int(info.max_front), "levels": int(info.nlevels), "nsuper": int(info.nsuper),
           "front_arena_gb": info.front_bytes / 1e9, "analyze_s": t_an}
    fac = pkg.CholeskyFactor(sym)
    d_nz = torch.from_numpy(np.ascontiguousarray(A.data)).to(dev)
    fac.factorize_dev(d_nz.data_ptr())
    ctx.sync()
    ts = []
    for _ in range(2):
        t = time.perf_counter()
        fac.factorize_dev(d_nz.data_ptr())
        ctx.sync()
        ts.append(time.perf_counter() - t)
    out["factor_s"] = min(ts)
    out["factor_tflops"] = info.flops / min(ts) * 1e-12
    rhs = np.random.default_rng(0).standard_normal(n)
    x = fac.solve(rhs)
    t = time.perf_counter()
    x = fac.solve(rhs)
    out["mean_solve_s"] = time.perf_counter() - t
    out["mean_residual"] = float(np.linalg.norm(A @ x - rhs) / np.linalg.norm(rhs))
    Qd = pkg.SparseMatrix(A, ctx=ctx)
    Z = torch.randn((50, n), dtype=torch.float64, device=dev)
    v_rb = fac.var_rbmc(Qd, Z)
    t = time.perf_counter()
    v_rb = fac.var_rbmc(Qd, Z)
    out["rbmc50_s"] = time.perf_counter() - t
    if info.front_bytes / 1e9 <= args.selinv_max_gb:
        t = time.perf_counter()
        v = fac.var_selinv()
        out["selinv_first_s"] = time.perf_counter() - t
        t = time.perf_counter()
        v = fac.var_selinv()
        out["selinv_s"] = time.perf_counter() - t
        out["var_positive"] = bool(np.all(v > 0))
        out["rbmc_median_rel_err_vs_selinv"] = float(np.median(np.abs(v_rb - v) / v))
    else:
        out["selinv_s"] = None
        out["selinv_note"] = "the inverse arena (same size as the factor arena) does not fit next to the factor on one GPU"
    res.append(out)
    print(json.dumps(out), flush=True)
    del fac, sym, Qd, d_nz, Z
    import gc

    gc.collect()
    pkg.pool_trim(0)
    torch.cuda.empty_cache()
if args.out:
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
