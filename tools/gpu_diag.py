"""GPU diagnostic: run each stage of the hot path at several sizes and print errors vs the oracle without
stopping at the first failure (used while bringing kernels up; results in gpurun_out/diag.txt)."""
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as g

pkg = g.load_pkg()
orc = g.load_oracle()
W = pkg.workloads
ctx = pkg.default_context()


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def stage(name, fn):
    try:
        t = time.time()
        r = fn()
        print(f"[ok ] {name}: {r}  ({time.time() - t:.2f}s)", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[ERR] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc(limit=3)


for b, N in [(8, 2), (64, 2), (100, 3), (200, 3), (300, 2)]:
    D, Bs = W.random_btd(b, N, seed=b)
    A = W.btd_to_sparse(D, Bs)
    Fo = orc.tridiagonal_cholesky(A, N)
    box = {}

    def f_factor():
        box["F"] = pkg.tridiagonal_cholesky_dense(D, Bs, ctx=ctx)
        ch, cs = box["F"].chos, box["F"].Cs
        return "L errs " + " ".join(f"{rel(ch[i].L, Fo.chos[i]):.1e}" for i in range(N)) + " C errs " + " ".join(
            f"{rel(cs[i], Fo.Cs[i]):.1e}" for i in range(N - 1))

    stage(f"btd factor b={b} N={N}", f_factor)
    rhs = np.random.default_rng(0).standard_normal((b * N, 2))
    if "F" in box:
        stage("  btd fwd", lambda: f"{rel(pkg.forward_solve(box['F'], rhs), np.stack([orc.btd_forward_solve(Fo, rhs[:, k]) for k in range(2)], 1)):.2e}")
        stage("  btd bwd", lambda: f"{rel(pkg.backward_solve(box['F'], rhs), np.stack([orc.btd_backward_solve(Fo, rhs[:, k]) for k in range(2)], 1)):.2e}")
        stage("  btd ldiv", lambda: f"{rel(pkg.ldiv(box['F'], rhs), np.stack([orc.btd_ldiv(Fo, rhs[:, k]) for k in range(2)], 1)):.2e}")
        stage("  btd selinv", lambda: f"{rel(box['F'].selinv_diag(), orc.btd_selinv_diag(Fo)):.2e}")

for nx in (5, 12, 30, 70):
    prob = W.matern_posterior(nx, obs_frac=0.2, q_eps=1e2, corr_range=0.15, seed=nx)
    Q = prob["Qpost"]
    n = Q.shape[0]
    box = {}

    def f_fac():
        box["sym"] = pkg.Symbolic(Q, ctx=ctx)
        i = box["sym"].info
        box["fac"] = pkg.CholeskyFactor(box["sym"]).factorize(Q.data)
        box["ref"] = orc.SparseCholesky(Q, box["sym"].p)
        return f"n={n} nsuper={i.nsuper} levels={i.nlevels} maxfront={i.max_front} diagL err {rel(box['fac'].diagL(), box['ref'].diagL()):.2e}"

    stage(f"sparse factor nx={nx}", f_fac)
    if "fac" in box:
        fac, ref = box["fac"], box["ref"]
        Bm = np.random.default_rng(1).standard_normal((n, 5))
        stage("  L values", lambda: f"{abs(fac.L - ref.L()).max() / abs(ref.L()).max():.2e}")
        stage("  PtL", lambda: f"{rel(fac.PtL_solve(Bm), ref.solve_PtL(Bm)):.2e}")
        stage("  UP", lambda: f"{rel(fac.UP_solve(Bm), ref.solve_UP(Bm)):.2e}")
        stage("  solve", lambda: f"{rel(fac.solve(Bm), ref.solve(Bm)):.2e}")
        stage("  selinv", lambda: f"{np.max(np.abs(fac.var_selinv() - ref.selinv_diag()) / ref.selinv_diag()):.2e}")
        stage("  rbmc", lambda: f"{rel(fac.var_rbmc(pkg.SparseMatrix(Q, ctx=ctx), Bm), orc.rbmc_variance(ref, Q, Bm)):.2e}")
print("launches", ctx.launch_count)
