#!/usr/bin/env python
"""Config 3 (Darcy dataset loop) on the element the reference's script uses: quadratic triangles with
QuadratureRule{RefTriangle}(3) (src/utils.jl:20-38, element_order = 2), Matern prior of smoothness 2 (alpha = 3,
scripts/darcy/solve_darcy_gmrf-fem.jl:97), coefficient looked up at every quadrature point.  Times the device assembly
(gmrfb_fem2d_*), the prior construction, and the posterior factor / mean / RBMC-50 on the resulting pattern; the NumPy
element loops of oracle/fem_oracle.py are timed beside the assembly.

    python tools/bench_fem2d.py --nel 300 --out profiles/r02_fem2d_darcy_p2.json
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
ap.add_argument("--nel", type=int, default=300, help="cells per side (N_xy of the script)")
ap.add_argument("--problems", type=int, default=3)
ap.add_argument("--out", default="")
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()

pkg, orc = entry.load_pkg(), entry.load_oracle()
W = pkg.workloads
ctx = pkg.Context(0)
res = {"config": "3 on quadratic triangles: Darcy stiffness with per-quadrature-point coefficient lookup, Matern alpha = 3 prior",
       "cells_per_side": args.nel}


def timed(fn, reps=1):
    ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        out = fn()
    ctx.sync()
    return out, (time.perf_counter() - t) / reps


nodes, tris = W.structured_mesh(args.nel + 1, args.nel + 1, seed=0)
nodes, elems = W.quadratic_mesh(nodes, tris)
n = nodes.shape[0]
res.update(n=int(n), elements=int(elems.shape[0]))
x, y = nodes[:, 0], nodes[:, 1]
bnd = (x == 0) | (x == 1) | (y == 0) | (y == 1)
fem, res["mesh_analysis_s"] = timed(lambda: pkg.FEMLagrange(nodes, elems, ctx=ctx))
res.update(fem.info)
g = 241
xc = yc = np.linspace(0, 1, g)
_, res["coeff_grid_lookup_s"] = timed(lambda: fem.set_coeff_grid(xc, yc))
grids = [W.darcy_problem(nx=9, seed=s)["coeff_grid"] for s in range(args.problems)]
(G, f), res["stiffness_first_s"] = timed(lambda: fem.stiffness(grids[0], prescribed=bnd))
_, res["stiffness_s"] = timed(lambda: fem.stiffness(grids[1 % len(grids)], prescribed=bnd, load=False), reps=5)
if not args.no_cpu:
    t = time.perf_counter()
    Gref, fref = orc.fem.assemble_darcy_lagrange(nodes, elems, 2, xc, yc, grids[0].T, prescribed=bnd)
    res["stiffness_numpy_element_loops_s"] = time.perf_counter() - t
    (G, f), _ = timed(lambda: fem.stiffness(grids[0], prescribed=bnd))
    Gd = G.to_scipy()
    res["stiffness_max_rel_vs_element_loops"] = float(abs(Gd - Gref).max() / abs(Gref).max())
    res["load_max_rel_vs_element_loops"] = float(np.abs(f - fref).max() / np.abs(fref).max())

try:
    # prior: Matern, range 1 / sqrt(300), smoothness 2 -> alpha = 3 (first call builds the two product plans)
    kappa = np.sqrt(8.0 * 2) / (1.0 / np.sqrt(300.0))
    ratio = 1.0 / (8.0 * np.pi * kappa**4)
    Q, res["matern_alpha3_first_s"] = timed(lambda: fem.matern_precision(kappa, ratio, alpha=3))
    Q, res["matern_alpha3_s"] = timed(lambda: fem.matern_precision(kappa, ratio, alpha=3), reps=3)
    res["nnz_Q"] = Q.dims()[2]

    # posterior of one problem: Q + A' 1e8 A (A = stiffness with identity rows), factor, mean, RBMC-50
    plan, res["posterior_plan_s"] = timed(lambda: pkg.PosteriorPrecision(Q, G))
    Apost, res["posterior_assembly_s"] = timed(lambda: plan.compute(1e8), reps=3)
    pat = Apost.to_scipy()
    res["nnz_Qpost"] = int(pat.nnz)
    sym, res["symbolic_s"] = timed(lambda: pkg.Symbolic(pat, coords=nodes, ctx=ctx))
    res.update(nnz_L=int(sym.info.nnz_L), factor_flops=float(sym.info.flops))
    fac = pkg.CholeskyFactor(sym)
    _, res["factor_first_s"] = timed(lambda: fac.factorize_dev(Apost.values_dev()))
    _, res["factor_s"] = timed(lambda: fac.factorize_dev(Apost.values_dev()), reps=3)
    res["factor_tflops"] = res["factor_flops"] / res["factor_s"] * 1e-12
    rhs = 1e8 * G.matvec(f, trans=True)
    mean, res["mean_s"] = timed(lambda: fac.solve(rhs), reps=3)
    r = pat @ mean - rhs
    res["mean_residual"] = float(np.linalg.norm(r) / np.linalg.norm(rhs))
    import torch

    # the normals live on the device, one sample per row, as in the dataset loop of tools/bench_configs.py (a NumPy
    # array would add a 144 MB host transpose and a pageable upload to every call)
    Z = torch.randn((50, n), dtype=torch.float64, device=f"cuda:{ctx.device}", generator=torch.Generator(f"cuda:{ctx.device}").manual_seed(0))
    _, res["rbmc50_first_s"] = timed(lambda: fac.var_rbmc(Apost, Z))
    v, res["rbmc50_s"] = timed(lambda: fac.var_rbmc(Apost, Z), reps=3)
    res["var_positive"] = bool(np.all(v > 0))
except Exception as e:  # noqa: BLE001 - keep what was measured
    res["error"] = f"{type(e).__name__}: {e}"
line = json.dumps(res)
print(line)
if args.out:
    open(args.out, "w").write(json.dumps(res, indent=1) + "\n")
