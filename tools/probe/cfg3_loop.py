"""Per-problem section times of the Darcy dataset loop (config 3), printed per problem; optional cProfile of one
condition_on_observations call.  GMRFB_GRAPHS=0/1 compares the graph policy."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g

pkg = g.load_pkg()
W = pkg.workloads
ctx = pkg.default_context()
nx = int(os.environ.get("NX", "601"))
nprob = int(os.environ.get("NPROB", "7"))
P0 = W.darcy_problem(nx, seed=0)
n = P0["Q"].shape[0]
bp = pkg.CholeskySolverBlueprint(var_strategy=pkg.RBMCStrategy(50), coords=P0["nodes"])
x = pkg.GMRF(np.zeros(n), P0["Q"], bp)
xc = pkg.condition_on_observations(x, P0["A"], P0["q_eps"], P0["y"], solver_blueprint=bp)
p = xc.solver_ref.value.precision_chol.p
bp2 = pkg.CholeskySolverBlueprint(var_strategy=pkg.RBMCStrategy(50), perm=p)
nodes, tris = W.structured_mesh(nx, nx, seed=0)
fem = pkg.FEMP1(nodes, tris, ctx=ctx)
gsz = P0["coeff_grid"].shape[0]
fem.set_coeff_grid(np.linspace(0, 1, gsz), np.linspace(0, 1, gsz))
xb, yb = nodes[:, 0], nodes[:, 1]
bnd = (xb == 0) | (xb == 1) | (yb == 0) | (yb == 1)
rng = np.random.default_rng(1)
probs = [W.darcy_problem(nx, seed=k) for k in range(nprob)]
for k, Pk in enumerate(probs):
    ctx.sync()
    Ad = fem.assemble(Pk["coeff_grid"], prescribed=bnd)
    ctx.sync()
    t0 = time.perf_counter()
    if k == nprob - 1:
        pr = cProfile.Profile()
        pr.enable()
    xk = pkg.condition_on_observations(x, Ad, Pk["q_eps"], Pk["y"], solver_blueprint=bp2)
    ctx.sync()
    if k == nprob - 1:
        pr.disable()
    t1 = time.perf_counter()
    m = pkg.mean(xk)
    t2 = time.perf_counter()
    s = pkg.rand(rng, xk)
    t3 = time.perf_counter()
    sd = pkg.std(xk)
    t4 = time.perf_counter()
    print(f"problem {k}: conditioning {1e3*(t1-t0):.1f} ms  mean {1e3*(t2-t1):.1f}  sample {1e3*(t3-t2):.1f}  std {1e3*(t4-t3):.1f}", flush=True)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
