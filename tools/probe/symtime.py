import sys,time
sys.path.insert(0,".")
import __graft_entry__ as e
pkg=e.load_pkg(); W=pkg.workloads
prob=W.matern_posterior(1001, obs_frac=0.1, q_eps=1e2, corr_range=0.05, seed=0)
ctx=pkg.Context(0)
for rep in range(5):
    t=time.time(); sym=pkg.Symbolic(prob["Qpost"], coords=prob["nodes"], ctx=ctx); print("analyze (GPU phases) s", time.time()-t, flush=True)
    if rep==2: del sym
