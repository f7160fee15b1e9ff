#!/usr/bin/env python
"""One pass over every kernel family of libgmrfb at small sizes, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
    compute-sanitizer --tool initcheck python tools/sanitize_smoke.py     (GMRFB_POOL=0 recommended)

Covers: GPU symbolic kernels, fused small-front kernels, blocked large-front path (POTRF64 / apply-inverse / grouped DMMA
GEMM / extend-add), level-scheduled solves, panel solves, selected inversion (gather / recursive-doubling inverse),
RBMC, posterior-precision assembly, FEM assembly, block-tridiagonal factor (serial and look-ahead schedules) / solves /
block selected inversion, device Gauss-Newton.  Exits non-zero if a result is wrong."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_pkg()
W = pkg.workloads
ctx = pkg.Context(0)
nx = int(os.environ.get("SAN_NX", "90"))
ok = True


def check(name, val, tol):
    global ok
    good = bool(val < tol)
    ok = ok and good
    print(f"{name:44s} {val:.3e}  {'ok' if good else 'FAIL'}", flush=True)


prob = W.matern_posterior(nx, obs_frac=0.2, q_eps=1e2, corr_range=0.15, seed=1)
Q, n = prob["Qpost"], prob["Qpost"].shape[0]
sym = pkg.Symbolic(Q, coords=prob["nodes"], ctx=ctx)  # GPU symbolic kernels
print("max front", sym.info.max_front, "levels", sym.info.nlevels, flush=True)
fac = pkg.CholeskyFactor(sym).factorize(Q.data)
rng = np.random.default_rng(0)
b = rng.standard_normal(n)
x = fac.solve(b)
check("mean residual", np.linalg.norm(Q @ x - b) / np.linalg.norm(b), 1e-11)
x = fac.solve(b)  # graph replay
check("mean residual (graph replay)", np.linalg.norm(Q @ x - b) / np.linalg.norm(b), 1e-11)
B = rng.standard_normal((n, 19))
X = fac.solve(B)  # panel path
check("panel residual (19 rhs)", np.linalg.norm(Q @ X - B) / np.linalg.norm(B), 1e-11)
S = fac.UP_solve(B)
check("samples: P'L^-T z consistency", np.linalg.norm(fac.PtL_solve(Q @ S) - B) / np.linalg.norm(B), 1e-9)
v = fac.var_selinv()
e = np.zeros(n)
e[n // 3] = 1.0
check("selinv diag entry vs solve", abs(v[n // 3] - fac.solve(e)[n // 3]) / v[n // 3], 1e-9)
Qd = pkg.SparseMatrix(Q, ctx=ctx)
vr = fac.var_rbmc(Qd, rng.standard_normal((n, 12)))
check("rbmc median rel err", float(np.median(np.abs(vr - v) / v)), 0.6)
xr = fac.solve(b, refine=Qd, max_iter=1)
check("refined residual", np.linalg.norm(Q @ xr - b) / np.linalg.norm(b), 1e-12)
# conditioning on the device + FEM assembly
nodes, tris = W.structured_mesh(31, 31, seed=0)
fem = pkg.FEMP1(nodes, tris, ctx=ctx)
m, G = W.p1_mass_stiffness(nodes, tris)
check("fem stiffness", abs(fem.assemble().to_scipy() - G).max() / abs(G).max(), 1e-12)
Qm = W.matern_precision(nodes, tris, 0.2)
kappa = np.sqrt(8.0) / 0.2
check("fem matern prior", abs(fem.matern_precision(kappa, 1.0 / (4 * np.pi * kappa**2)).to_scipy() - Qm).max() / abs(Qm).max(), 1e-12)
p2 = W.matern_posterior(31, obs_frac=0.3, seed=2)
xg = pkg.GMRF(np.zeros(31 * 31), p2["Q"], pkg.CholeskySolverBlueprint(ctx=ctx))
xc = pkg.condition_on_observations(xg, p2["A"], p2["q_eps"], p2["y"])
mc = pkg.mean(xc)
check("conditioned mean residual", np.linalg.norm(p2["Qpost"] @ mc - p2["rhs"]) / np.linalg.norm(p2["rhs"]), 1e-10)
# block tridiagonal: serial and look-ahead schedules
for la in ("0", "1"):
    os.environ["GMRFB_BTD_LOOKAHEAD"] = la
    D, Bs = W.random_btd(200, 4, seed=3)
    F = pkg.tridiagonal_cholesky_dense(D, Bs, ctx=ctx)
    A = W.btd_to_sparse(D, Bs)
    R = rng.standard_normal((800, 3))
    Xb = pkg.ldiv(F, R)
    check(f"btd residual (lookahead={la})", np.linalg.norm(A @ Xb - R) / np.linalg.norm(R), 1e-11)
    vb = F.selinv_diag()
    check(f"btd selinv vs sparse (lookahead={la})", np.max(np.abs(vb - pkg.cholesky(A, ctx=ctx).var_selinv()) / vb), 1e-8)
# device Gauss-Newton (bilinear Burgers residual, scripts/solve_burger.jl:143-180)
P = W.burgers_spacetime(16, 5)
dgn = pkg.DeviceGaussNewton(P["mu"], P["Q"], P["L"], P["A"], P["D"], P["c"], 1e4, P["y"], P["mu"],
                            solver_bp=pkg.GNCholeskySolverBlueprint(ctx=ctx))
xd = dgn.optimize()
gno = pkg.GaussNewtonOptimizer(P["mu"], P["Q"], P["f_and_J"], 1e4, P["y"], P["mu"], solver_bp=pkg.GNCholeskySolverBlueprint(ctx=ctx))
xh = pkg.optimize(gno)
check("device GN vs host-driven loop", np.linalg.norm(xd - xh) / np.linalg.norm(xh), 1e-9)
print("SANITIZE_SMOKE", "PASS" if ok else "FAIL", flush=True)
sys.exit(0 if ok else 1)
