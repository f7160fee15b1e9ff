"""Gauss-Newton tangent assembly on the device against the restated host element loops, at the sizes of the reference's
scripts:
  * config 1 (_research/elliptic_chen24.jl:118-171): 201 x 201 P1 nodes (n = 40 401), f_and_J = stiffness + cubic term
    by 3-point quadrature, rows of the boundary dofs skipped;
  * config 2' (scripts/burgers/solve_burgers_gmrf-fem.jl): 800 periodic quadratic lines (1 600 dofs) x 101 time steps
    (n = 161 600), f_and_J = J_static + dt J_adv over all steps.
Each: time of one f_and_J evaluation (device kernel path vs oracle/fem_oracle.py on the host), agreement of the two, and
the whole Gauss-Newton loop (GaussNewtonOptimizer) driven by the device tangent.  One JSON line per case on stdout.

    python tools/bench_fem_tangents.py [--small]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
args = ap.parse_args()
pkg = g.load_pkg()
orc = g.load_oracle()
fo = orc.fem
W = pkg.workloads
ctx = pkg.default_context()


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def best_of(fn, reps):
    fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ctx.sync()
        ts.append(time.perf_counter() - t)
    return out, min(ts)


def elliptic():
    nx = 41 if args.small else 201
    P = W.elliptic_problem(nx)
    nodes, tris = W.structured_mesh(nx, nx, seed=0)
    n = P["n"]
    x, y = nodes[:, 0], nodes[:, 1]
    bnd = (x == 0) | (x == 1) | (y == 0) | (y == 1)
    out = {"case": "config 1 tangent: stiffness + cubic term, P1, 3-point rule (_research/elliptic_chen24.jl:180-285)",
           "mesh": f"{nx}x{nx}", "n": n}
    Jd = fo.assemble_stiffness_skipped_rows_p1(nodes, tris, bnd)
    gl = np.where(bnd, 0.0, P["y"])
    t = time.perf_counter()
    fem = pkg.FEMP1(nodes, tris, ctx=ctx)
    out["mesh_analysis_s"] = time.perf_counter() - t
    u = P["u_true"] + 0.1 * np.random.default_rng(0).standard_normal(n)
    (fd, Jdev), out["device_f_and_J_s"] = best_of(lambda: fem.assemble_cubic(u, prescribed=bnd), 20)

    def host():
        Jc, fc = fo.assemble_cubic_p1(nodes, tris, u, bnd, 2)
        return Jd @ u + fc, (Jd + Jc).tocsc()

    (fh, Jh), out["host_numpy_f_and_J_s"] = best_of(host, 3)
    out["J_max_rel_diff"] = float(abs(Jdev.to_scipy() - Jh).max() / abs(Jh).max())
    out["f_max_rel_diff"] = float(np.abs(fd - fh).max() / np.abs(fh).max())
    # device-resident iterate and residual (what a device-side loop passes): no host transfer at all
    import torch
    ud = torch.as_tensor(u, device="cuda")
    fdev = torch.empty(n, dtype=torch.float64, device="cuda")
    _, out["device_f_and_J_s (device vectors)"] = best_of(lambda: fem.assemble_cubic(ud, prescribed=bnd, out=fdev), 20)
    xc = pkg.condition_on_observations(pkg.GMRF(np.zeros(n), P["Q"], pkg.CholeskySolverBlueprint(coords=nodes, ctx=ctx)),
                                       P["A_bnd"], 1e8, P["y_bnd"])
    p = xc.solver_ref[()].precision_chol.p

    def run(fj):
        gno = pkg.GaussNewtonOptimizer(pkg.mean(xc), pkg.precision_map(xc), fj, 1e6, np.zeros(n), pkg.mean(xc),
                                       solver_bp=pkg.GNCholeskySolverBlueprint(p, ctx=ctx), max_steps=10)
        t0 = time.perf_counter()
        xs = pkg.optimize(gno)
        ctx.sync()
        return xs, gno.n_steps, time.perf_counter() - t0

    def fj_dev(w):
        f, J = fem.assemble_cubic(np.ascontiguousarray(w), prescribed=bnd)
        return f - gl, J

    def fj_host(w):
        f, J = (lambda Jc_fc: (Jd @ w + Jc_fc[1], (Jd + Jc_fc[0]).tocsc()))(fo.assemble_cubic_p1(nodes, tris, w, bnd, 2))
        return f - gl, J

    run(fj_dev)
    xd, sd, td = run(fj_dev)
    xh, sh, th = run(fj_host)
    out["gauss_newton"] = {"steps_device_tangent": sd, "optimize_s_device_tangent": td, "steps_host_tangent": sh,
                           "optimize_s_host_tangent": th, "rel_diff_of_solutions": rel(xd, xh),
                           "rel_err_vs_manufactured_solution": rel(xd, P["u_true"])}
    return out


def burgers():
    ne, nt = (100, 11) if args.small else (800, 101)
    order, dt, nu = 2, 0.01, 0.01
    xe, el = W.periodic_line_mesh(ne, order)
    ns = int(el.max()) + 1
    n = ns * nt
    out = {"case": "config 2' tangent: space-time Burgers f_and_J (scripts/burgers/solve_burgers_gmrf-fem.jl:115-142)",
           "elements": ne, "order": order, "dofs_per_step": ns, "steps": nt, "n": n}
    t = time.perf_counter()
    fem = pkg.FEM1D(xe, el, order=order, ctx=ctx)
    out["mesh_analysis_s"] = time.perf_counter() - t
    rng = np.random.default_rng(0)
    xs = np.zeros(ns)
    xs[el.ravel()] = xe.ravel() % 1.0
    u0 = np.zeros(ns)
    for k in range(1, 5):
        u0 += rng.standard_normal() / k * np.sin(2 * np.pi * k * xs) + rng.standard_normal() / k * np.cos(2 * np.pi * k * xs)
    w = np.tile(u0, nt) + 0.01 * rng.standard_normal(n)
    (fd, Jdev), out["device_f_and_J_s (first call builds the space-time pattern)"] = best_of(
        lambda: fem.spacetime_tangent(w, nt, dt, nu), 20)
    (fh, Jh), out["host_numpy_f_and_J_s"] = best_of(lambda: fo.burgers_spacetime_tangent(xe, el, w, nt, dt, nu, order), 2)
    out["nnz_J"] = int(Jh.nnz)
    out["J_max_rel_diff"] = float(abs(Jdev.to_scipy() - Jh).max() / abs(Jh).max())
    out["f_max_rel_diff"] = float(np.abs(fd - fh).max() / np.abs(fh).max())
    import torch
    wd = torch.as_tensor(w, device="cuda")
    fdev = torch.empty((nt - 1) * ns, dtype=torch.float64, device="cuda")
    _, out["device_f_and_J_s (device vectors)"] = best_of(lambda: fem.spacetime_tangent(wd, nt, dt, nu, out=fdev), 20)
    # prior: implicit-Euler diffusion state-space model from the FEM matrices (src/spdes/shallow_water.jl:198-228),
    # initial condition observed with Q_eps = 1e8 (scripts/burgers/solve_burgers_gmrf-fem.jl:144,161)
    M, G = (A.to_scipy() for A in fem.mass_stiffness())
    ml = np.asarray(M.sum(axis=1)).ravel()
    Gs = (M + dt * nu * G).tocsc()
    binv = sp.diags(1.0 / (dt * ml))
    kappa = 8.0
    K0 = (kappa**2 * M + G).tocsc()
    Q0 = (K0.T @ sp.diags(1.0 / ml) @ K0).tocsc()
    Id = sp.identity(ns, format="csc")
    T = sp.identity(nt, format="csc")
    first = sp.csc_matrix(([1.0], ([0], [0])), shape=(nt, nt))
    last = sp.csc_matrix(([1.0], ([nt - 1], [nt - 1])), shape=(nt, nt))
    sub = sp.csc_matrix((np.ones(nt - 1), (np.arange(1, nt), np.arange(nt - 1))), shape=(nt, nt))
    off = -(Gs.T @ binv @ M)
    Q = (sp.kron(first, Q0 + 1e8 * Id) + sp.kron(T - first, Gs.T @ binv @ Gs) + sp.kron(T - last, M.T @ binv @ M)
         + sp.kron(sub, off) + sp.kron(sub.T, off.T)).tocsc()
    Q.sort_indices()
    mu = np.tile(u0, nt)
    noise = 1e8

    def run(fj, steps):
        gno = pkg.GaussNewtonOptimizer(mu, Q, fj, noise, np.zeros((nt - 1) * ns), mu,
                                       solver_bp=pkg.GNCholeskySolverBlueprint(ctx=ctx), max_steps=steps)
        t0 = time.perf_counter()
        xk = pkg.optimize(gno)
        ctx.sync()
        return xk, gno, time.perf_counter() - t0

    fj_dev = lambda v: fem.spacetime_tangent(np.ascontiguousarray(v), nt, dt, nu)  # noqa: E731
    fj_host = lambda v: fo.burgers_spacetime_tangent(xe, el, v, nt, dt, nu, order)  # noqa: E731
    run(fj_dev, 2)
    xd, gd, td = run(fj_dev, 20)
    xh, gh, th = run(fj_host, 20)
    fx, _ = fj_dev(xd)
    out["gauss_newton"] = {"steps_device_tangent": gd.n_steps, "optimize_s_device_tangent (incl. symbolic analysis)": td,
                           "steps_host_tangent": gh.n_steps, "optimize_s_host_tangent": th,
                           "rel_diff_of_solutions": rel(xd, xh), "residual_norm_first_last": [gd.r_obs_norm_history[0],
                                                                                              gd.r_obs_norm_history[-1]],
                           "final_pde_residual_inf": float(np.abs(fx).max())}
    return out


for fn in (elliptic, burgers):
    t = time.perf_counter()
    res = fn()
    res["wall_s"] = time.perf_counter() - t
    print(json.dumps(res), flush=True)
