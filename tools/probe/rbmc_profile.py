"""Where the time of an RBMC-50 variance estimate goes (tuning aid): wall vs kernel time of gmrfb_var_rbmc."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
pkg = g.load_pkg(); W = pkg.workloads; ctx = pkg.default_context()
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 601
P = W.matern_posterior(nx, obs_frac=0.1, q_eps=1e2, corr_range=0.05, seed=0)
Qp = P["Qpost"]; n = Qp.shape[0]
sym = pkg.Symbolic(Qp, coords=P["nodes"], ctx=ctx)
fac = pkg.CholeskyFactor(sym).factorize(Qp.data)
Qd = pkg.SparseMatrix(Qp, ctx=ctx)
t = time.perf_counter(); Z = np.random.default_rng(0).standard_normal((n, 50)); t_rng = time.perf_counter() - t
Zf = np.asfortranarray(Z)
fac.var_rbmc(Qd, Zf)
for rep in range(2):
    ctx.profile_begin(); ctx.sync(); t = time.perf_counter()
    v = fac.var_rbmc(Qd, Zf)
    ctx.sync(); dt = time.perf_counter() - t
    prof = ctx.profile_end()
print(f"n={n} nnz_L={sym.info.nnz_L:.3e}: rng {t_rng*1e3:.0f} ms, var_rbmc wall {dt*1e3:.1f} ms, kernels {sum(p['ms'] for p in prof):.1f} ms")
for p in sorted(prof, key=lambda p: -p["ms"])[:8]:
    print(f"   {p['name']:24s} n={p['launches']:5d} ms={p['ms']:8.2f}")
t = time.perf_counter(); x = fac.solve(P["rhs"]); ctx.sync(); print("single solve wall ms", (time.perf_counter() - t) * 1e3)
