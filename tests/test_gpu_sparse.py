"""Parity of the CUDA sparse path (through the C ABI) against the CPU oracle on identical seeded inputs.
Tolerances are the north star's: means/samples 1e-10 relative, marginal variances 1e-8 relative."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

TOL_SOLVE = 1e-10
TOL_VAR = 1e-8


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.fixture(scope="module")
def problems(W):
    out = {}
    for nx in (7, 24, 61, 130):
        out[nx] = W.matern_posterior(nx, obs_frac=0.2, q_eps=1e2, corr_range=0.15, seed=nx)
    return out


@pytest.mark.parametrize("nx", [7, 24, 61, 130])
@pytest.mark.parametrize("mode", ["nd", "geo", "given", "amd"])
def test_factor_solve_sample(pkg, orc, ctx, problems, nx, mode):
    prob = problems[nx]
    Q = prob["Qpost"]
    n = Q.shape[0]
    rng = np.random.default_rng(nx)
    if mode == "nd":
        sym = pkg.Symbolic(Q, ctx=ctx)
    elif mode == "geo":
        sym = pkg.Symbolic(Q, ctx=ctx, coords=prob["nodes"])
    elif mode == "amd":
        # approximate minimum degree: a deep, unbalanced supernodal tree (many levels with few fronts each)
        sym = pkg.Symbolic(Q, ctx=ctx, ordering="amd")
    else:
        # "perm = p reused" path of the reference: ordering from a first analysis, then ORDER_GIVEN
        p0 = pkg.Symbolic(Q, host_only=True).p
        sym = pkg.Symbolic(Q, perm=p0, ctx=ctx)
        assert np.array_equal(sym.p, p0)
    fac = pkg.CholeskyFactor(sym).factorize(Q.data)
    assert fac.issuccess()
    ref = orc.SparseCholesky(Q, sym.p)
    # pattern-level parity is bit exact
    assert np.array_equal(sym.parent, ref.parent) and np.array_equal(sym.colcount, ref.colcount)
    assert sym.info.nnz_L == ref.nnz
    B = rng.standard_normal((n, 5))
    assert rel(fac.solve(B), ref.solve(B)) < TOL_SOLVE            # F \ b (posterior mean)
    assert rel(fac.solve(B[:, 0]), ref.solve(B[:, 0])) < TOL_SOLVE
    assert rel(fac.PtL_solve(B), ref.solve_PtL(B)) < TOL_SOLVE    # F.PtL \ b
    assert rel(fac.UP_solve(B), ref.solve_UP(B)) < TOL_SOLVE      # F.UP \ z (samples under the same z)
    mu = rng.standard_normal(n)
    assert rel(fac.sample(B, mean=mu), ref.solve_UP(B) + mu[:, None]) < TOL_SOLVE
    np.testing.assert_allclose(fac.diagL(), ref.diagL(), rtol=1e-11)
    assert abs(fac.logdet() - ref.logdet()) < 1e-10 * abs(ref.logdet())
    # residual check, independent of the oracle
    x = fac.solve(prob["rhs"])
    assert np.linalg.norm(Q @ x - prob["rhs"]) < 1e-11 * np.linalg.norm(prob["rhs"]) * 1e2


@pytest.mark.parametrize("nx", [7, 24, 61])
def test_factor_values(pkg, orc, ctx, problems, nx):
    Q = problems[nx]["Qpost"]
    sym = pkg.Symbolic(Q, ctx=ctx)
    fac = pkg.CholeskyFactor(sym).factorize(Q.data)
    L = fac.L
    Lref = orc.SparseCholesky(Q, sym.p).L()
    d = (L - Lref)
    assert abs(d).max() < 1e-11 * abs(Lref).max()
    # same structural pattern up to exact zeros created by relaxed supernodes (dropped by get_L)
    assert (abs(Lref) > 0).nnz >= L.nnz * 0.999


@pytest.mark.parametrize("nx", [7, 24, 61, 130])
def test_selected_inversion(pkg, orc, ctx, problems, nx):
    _selected_inversion(pkg, orc, ctx, problems, nx, "nd")


def test_inplace_solve_and_caller_owned_outputs(pkg, orc, ctx, problems):
    """``solve_inplace`` (ldiv!: the C ABI overwrites the caller's buffer) and ``var_selinv(out=...)`` with page-locked
    host buffers give the same numbers as the allocating forms, for a vector and for a column-major panel."""
    import torch
    P = problems[24]
    n = P["Qpost"].shape[0]
    sym = pkg.Symbolic(P["Qpost"], coords=P["nodes"], ctx=ctx)
    fac = pkg.CholeskyFactor(sym).factorize(P["Qpost"].data)
    x_ref = fac.solve(P["rhs"])
    xp = torch.empty(n, dtype=torch.float64).pin_memory()
    xp.copy_(torch.from_numpy(P["rhs"]))
    out = fac.solve_inplace(xp.numpy())
    assert out.ctypes.data == xp.numpy().ctypes.data and np.array_equal(out, x_ref)
    Bm = np.asfortranarray(np.random.default_rng(0).standard_normal((n, 7)))
    X_ref = fac.solve(Bm)
    fac.solve_inplace(Bm)
    assert np.array_equal(Bm, X_ref)
    vp = torch.empty(n, dtype=torch.float64).pin_memory()
    v = fac.var_selinv(out=vp.numpy())
    assert np.array_equal(v, fac.var_selinv())
    with pytest.raises(AssertionError):
        fac.solve_inplace(np.ascontiguousarray(Bm))  # row-major panel: not what the ABI reads


@pytest.mark.parametrize("nx", [24, 130])
def test_selected_inversion_amd_ordering(pkg, orc, ctx, problems, nx):
    _selected_inversion(pkg, orc, ctx, problems, nx, "amd")


def _selected_inversion(pkg, orc, ctx, problems, nx, ordering):
    Q = problems[nx]["Qpost"]
    sym = pkg.Symbolic(Q, ctx=ctx, ordering=ordering)
    fac = pkg.CholeskyFactor(sym).factorize(Q.data)
    v = fac.var_selinv()
    ref = orc.SparseCholesky(Q, sym.p)
    vref = ref.selinv_diag()
    assert np.max(np.abs(v - vref) / np.abs(vref)) < TOL_VAR
    if nx <= 24:
        np.testing.assert_allclose(v, orc.dense_inverse_diag(Q), rtol=TOL_VAR)
    # off-diagonal selected entries on the pattern of Q
    C = sp.coo_matrix(sp.tril(Q))
    pick = np.random.default_rng(0).choice(C.nnz, size=min(200, C.nnz), replace=False)
    vals = fac.selinv_entries(C.row[pick], C.col[pick])
    Z = ref.selinv()  # permuted ordering, lower
    ip = np.empty(Q.shape[0], np.int64)
    ip[sym.p] = np.arange(Q.shape[0])
    a, b = ip[C.row[pick]], ip[C.col[pick]]
    want = np.asarray(Z[np.maximum(a, b), np.minimum(a, b)]).ravel()
    np.testing.assert_allclose(vals, want, rtol=1e-8, atol=1e-14)


def test_refactorize_same_pattern(pkg, orc, ctx, problems):
    """Gauss-Newton pattern reuse: new values, same symbolic handle and factor storage."""
    prob = problems[24]
    Q = prob["Qpost"]
    sym = pkg.Symbolic(Q, ctx=ctx)
    fac = pkg.CholeskyFactor(sym)
    b = np.random.default_rng(1).standard_normal(Q.shape[0])
    for scale in (1.0, 3.5, 0.25):
        Q2 = Q.copy()
        Q2.data = Q.data * scale
        Q2 = (Q2 + sp.identity(Q.shape[0]) * 0).tocsc()
        fac.factorize(Q2.data)
        ref = orc.SparseCholesky(Q2, sym.p)
        assert rel(fac.solve(b), ref.solve(b)) < TOL_SOLVE
        assert rel(fac.var_selinv(), ref.selinv_diag()) < TOL_VAR


def test_not_positive_definite(pkg, ctx, problems):
    Q = problems[24]["Qpost"].copy()
    sym = pkg.Symbolic(Q, ctx=ctx)
    fac = pkg.CholeskyFactor(sym)
    bad = Q.data.copy()
    bad[Q.indptr[100]:Q.indptr[101]][Q.indices[Q.indptr[100]:Q.indptr[101]] == 100] = -1.0
    with pytest.raises(pkg.NotPositiveDefinite):
        fac.factorize(bad)
    fac.factorize(bad, check=False)  # Julia check=false: no throw, queryable flag
    assert not fac.issuccess() and fac.info.status == 2 and fac.info.fail_column >= 0
    with pytest.raises(pkg.GmrfbError):
        fac.solve(np.ones(Q.shape[0]))
    fac.factorize(Q.data)  # recovers
    assert fac.issuccess()


def test_edge_matrices(pkg, orc, ctx):
    for A in (sp.identity(1, format="csc") * 4.0,
              sp.diags(np.arange(1.0, 40.0)).tocsc(),
              sp.diags([-np.ones(299), 2.5 * np.ones(300), -np.ones(299)], [-1, 0, 1], format="csc")):
        sym = pkg.Symbolic(A, ctx=ctx)
        fac = pkg.CholeskyFactor(sym).factorize(A.data)
        b = np.random.default_rng(0).standard_normal(A.shape[0])
        ref = orc.SparseCholesky(A, sym.p)
        assert rel(fac.solve(b), ref.solve(b)) < TOL_SOLVE
        np.testing.assert_allclose(fac.var_selinv(), ref.selinv_diag(), rtol=TOL_VAR)
    # dense SPD matrix as a sparse one: a single big supernode (front order 300)
    R = np.random.default_rng(1).standard_normal((300, 300))
    A = sp.csc_matrix(R @ R.T + 300 * np.eye(300))
    sym = pkg.Symbolic(A, ctx=ctx)
    fac = pkg.CholeskyFactor(sym).factorize(A.data)
    b = np.ones(300)
    assert rel(fac.solve(b), np.linalg.solve(A.toarray(), b)) < TOL_SOLVE
    np.testing.assert_allclose(fac.var_selinv(), np.diag(np.linalg.inv(A.toarray())), rtol=TOL_VAR)


def test_rbmc_matches_oracle_same_z(pkg, orc, ctx, problems):
    prob = problems[24]
    Q = prob["Qpost"]
    n = Q.shape[0]
    sym = pkg.Symbolic(Q, ctx=ctx)
    fac = pkg.CholeskyFactor(sym).factorize(Q.data)
    Z = np.random.default_rng(5).standard_normal((n, 50))  # RBMCStrategy(50)
    v = fac.var_rbmc(pkg.SparseMatrix(Q, ctx=ctx), Z)
    vref = orc.rbmc_variance(orc.SparseCholesky(Q, sym.p), Q, Z)
    np.testing.assert_allclose(v, vref, rtol=1e-9)


def test_rbmc_and_samples_sharded_over_emulated_ranks(pkg, orc, ctx, problems):
    """SURVEY.md 8(e) row 2 through the C ABI: each (emulated) rank factorises redundantly and handles its share of the
    50 sample columns; the sample-count-weighted sum of the per-rank estimates (what the all-reduce of
    dist.rbmc_variance_sharded adds up) equals the single-GPU estimate and the oracle's."""
    prob = problems[24]
    Q = prob["Qpost"]
    n = Q.shape[0]
    sym = pkg.Symbolic(Q, ctx=ctx)
    Qd = pkg.SparseMatrix(Q, ctx=ctx)
    Z = np.random.default_rng(5).standard_normal((n, 50))
    world = 8
    bounds = pkg.dist.sample_bounds(50, world)
    acc = np.zeros(n)
    cols = []
    for r in range(world):
        fac = pkg.CholeskyFactor(sym).factorize(Q.data)  # one factor per rank
        lo, hi = bounds[r]
        acc += fac.var_rbmc(Qd, np.asfortranarray(Z[:, lo:hi])) * ((hi - lo) / 50)
        cols.append(pkg.dist.rand_sharded(fac, Z, r, world, mean=prob["rhs"]))
    one = pkg.dist.rbmc_variance_sharded(pkg.CholeskyFactor(sym).factorize(Q.data), Qd, Z, 0, 1)
    ref = orc.SparseCholesky(Q, sym.p)
    np.testing.assert_allclose(acc, one, rtol=1e-12)
    np.testing.assert_allclose(acc, orc.rbmc_variance(ref, Q, Z), rtol=1e-9)
    X = np.hstack(cols)
    Xref = ref.solve_UP(Z) + prob["rhs"][:, None]
    assert np.max(np.abs(X - Xref)) <= 1e-10 * np.max(np.abs(Xref))


def test_spmv_sqmahal_postprec(pkg, orc, ctx, problems):
    prob = problems[24]
    Q, A = prob["Q"], prob["A"]
    n = Q.shape[0]
    rng = np.random.default_rng(2)
    Qd, Ad = pkg.SparseMatrix(Q, ctx=ctx), pkg.SparseMatrix(A, ctx=ctx)
    x = rng.standard_normal(n)
    r = rng.standard_normal(A.shape[0])
    np.testing.assert_allclose(Qd.matvec(x), Q @ x, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Ad.matvec(x), A @ x, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Ad.matvec(r, trans=True), A.T @ r, rtol=1e-12, atol=1e-12)
    y0 = rng.standard_normal(n)
    np.testing.assert_allclose(Qd.matvec(x, alpha=2.0, beta=-0.5, y=y0), 2 * (Q @ x) - 0.5 * y0, rtol=1e-12, atol=1e-12)
    mu = rng.standard_normal(n)
    assert abs(Qd.sqmahal(x, mu) - (x - mu) @ (Q @ (x - mu))) < 1e-11 * abs((x - mu) @ (Q @ (x - mu)))
    plan = pkg.PosteriorPrecision(Qd, Ad)
    for qe in (1e2, 7.0):
        got = plan.compute(qe).to_scipy()
        want = orc.posterior_precision(Q, A, qe)
        assert abs(got - want).max() < 1e-12 * abs(want).max()
    w = rng.random(A.shape[0]) + 0.5
    got = plan.compute(w).to_scipy()
    assert abs(got - orc.posterior_precision(Q, A, w)).max() < 1e-12 * abs(got).max()
    # a general (non-selection) observation operator: the Darcy case A = stiffness-like matrix
    G = (Q.copy())
    Gd = pkg.SparseMatrix(G, ctx=ctx)
    plan2 = pkg.PosteriorPrecision(Qd, Gd)
    got = plan2.compute(3.0).to_scipy()
    want = orc.posterior_precision(Q, G, 3.0)
    assert abs(got - want).max() < 1e-11 * abs(want).max()


def test_gmrf_interface(pkg, orc, ctx, problems):
    """condition_on_observations / mean / std / rand / sqmahal with the reference's blueprint spelling
    (scripts/darcy/solve_darcy_gmrf-fem.jl:100,165-192)."""
    prob = problems[24]
    Q, A, y, qe = prob["Q"], prob["A"], prob["y"], prob["q_eps"]
    n = Q.shape[0]
    x = pkg.GMRF(np.zeros(n), Q, pkg.CholeskySolverBlueprint(ctx=ctx))
    xc = pkg.condition_on_observations(x, A, qe, y)
    p = xc.solver_ref[()].precision_chol.p
    assert sorted(p.tolist()) == list(range(n))
    # second conditioning with the permutation reused (cbp2 = perm p) and RBMC variances
    bp2 = pkg.CholeskySolverBlueprint(var_strategy=pkg.RBMCStrategy(50, rng=np.random.default_rng(7)), perm=p, ctx=ctx)
    xc2 = pkg.condition_on_observations(x, A, qe, y, solver_blueprint=bp2)
    assert np.array_equal(xc2.solver_ref[()].precision_chol.p, p)
    Qp = orc.posterior_precision(Q, A, qe)
    ref = orc.SparseCholesky(Qp, p)
    mref = orc.posterior_mean(ref, Q, A, qe, y, np.zeros(n))
    assert rel(pkg.mean(xc), mref) < TOL_SOLVE and rel(pkg.mean(xc2), mref) < TOL_SOLVE
    assert rel(pkg.std(xc), np.sqrt(ref.selinv_diag())) < TOL_VAR
    Z = np.random.default_rng(7).standard_normal((50, n)).T  # the host mirror draws one sample per contiguous row
    assert rel(pkg.std(xc2), np.sqrt(orc.rbmc_variance(ref, Qp, Z))) < 1e-9
    # without an rng the normals are drawn on the device: a Monte-Carlo estimate of the exact (Takahashi) variances
    bp3 = pkg.CholeskySolverBlueprint(var_strategy=pkg.RBMCStrategy(4000), perm=p, ctx=ctx)
    xc3 = pkg.condition_on_observations(x, A, qe, y, solver_blueprint=bp3)
    assert np.max(np.abs(pkg.var(xc3) - ref.selinv_diag()) / ref.selinv_diag()) < 0.15
    z = np.random.default_rng(11).standard_normal(n)
    s = pkg.rand(np.random.default_rng(11), xc)
    assert rel(s, mref + ref.solve_UP(z)) < TOL_SOLVE
    v = np.random.default_rng(3).standard_normal(n)
    want = (v - mref) @ (Qp @ (v - mref))
    assert abs(pkg.sqmahal(xc, v) - want) < 1e-9 * abs(want)
    assert xc.solver_ref[()].precision_chol.nnz >= ref.nnz


def test_gauss_newton(pkg, orc, ctx, W):
    """GaussNewtonOptimizer/optimize against the restated loop of scripts/solve_burger.jl:143-180."""
    P = W.burgers_like_problem(48)
    n = P["n"]
    noise = 1e4
    p0 = pkg.Symbolic(orc.posterior_precision(P["Q"], P["f_and_J"](P["mu"])[1], noise), host_only=True).p
    gno = pkg.GaussNewtonOptimizer(P["mu"], P["Q"], P["f_and_J"], noise, P["y"], P["mu"],
                                   solver_bp=pkg.GNCholeskySolverBlueprint(p0, ctx=ctx))
    xg = pkg.optimize(gno)
    # oracle loop
    x = P["mu"].copy()
    Qx = P["Q"] @ P["mu"]

    def obj(x):
        r = P["y"] - P["f"](x)
        d = P["mu"] - x
        return d @ (P["Q"] @ d) + noise * (r @ r)

    last, cur, steps = np.inf, obj(x), 0
    while abs(last - cur) / abs(cur) > 1e-4 and steps < 20:
        fx, J = P["f_and_J"](x)
        x = orc.gauss_newton_step(P["Q"], J, noise, x, Qx, P["y"] - fx, p0)
        last, cur = cur, obj(x)
        steps += 1
    assert steps == gno.n_steps and steps >= 2
    assert rel(xg, x) < 1e-9


def test_dataset_loop_reuses_workspace(pkg, orc, ctx, W):
    """The Darcy dataset loop (scripts/darcy/solve_darcy_gmrf-fem.jl:210): one prior, one mesh, a new coefficient
    field per problem.  Later calls of condition_on_observations reuse the device matrices, the fixed-pattern plan and
    the symbolic analysis and only upload A's values; every posterior must still match the oracle and earlier
    posteriors must stay valid (each owns its factor)."""
    P0 = W.darcy_problem(17, seed=0, q_eps=1e4)
    n = P0["Q"].shape[0]
    x = pkg.GMRF(np.zeros(n), P0["Q"], pkg.CholeskySolverBlueprint(coords=P0["nodes"], ctx=ctx))
    first = pkg.condition_on_observations(x, P0["A"], P0["q_eps"], P0["y"])
    p = first.solver_ref[()].precision_chol.p
    bp = pkg.CholeskySolverBlueprint(perm=p, ctx=ctx)
    posts = []
    for k in range(3):
        Pk = W.darcy_problem(17, seed=k, q_eps=1e4)
        posts.append((Pk, pkg.condition_on_observations(x, Pk["A"], Pk["q_eps"], Pk["y"], solver_blueprint=bp)))
    assert x._cond_ws is not None and bp._sym_cache is not None
    syms = {id(xc.solver_ref[()].precision_chol.sym) for _, xc in posts}
    assert len(syms) == 1  # one symbolic analysis for the whole loop
    for Pk, xc in posts:  # checked after the loop: earlier posteriors were not overwritten by later ones
        Qp = orc.posterior_precision(Pk["Q"], Pk["A"], Pk["q_eps"])
        ref = orc.SparseCholesky(Qp, p)
        mref = orc.posterior_mean(ref, Pk["Q"], Pk["A"], Pk["q_eps"], Pk["y"], np.zeros(n))
        assert abs(pkg.precision_map(xc) - Qp).max() < 1e-12 * abs(Qp).max()
        assert rel(pkg.mean(xc), mref) < 1e-9
        assert rel(pkg.var(xc), ref.selinv_diag()) < TOL_VAR


def test_dataset_loop_frees_factors_and_replays_graphs(pkg, orc, ctx, W):
    """The dataset loop as the reference script writes it (``x_k = condition_on_observations(...)`` rebinding one
    name): the factor of the previous problem is released by reference counting alone (no GMRF <-> solver cycle left
    to the cyclic collector), its buffers come back from the pool for the next factor, and that factor replays the
    CUDA graphs captured for its predecessor (cache shared through the symbolic handle, keyed by every buffer
    address).  Every posterior must still match the oracle - a stale pointer in a replayed graph would show here."""
    import gc
    import weakref
    P0 = W.darcy_problem(23, seed=0, q_eps=1e4)
    n = P0["Q"].shape[0]
    x = pkg.GMRF(np.zeros(n), P0["Q"], pkg.CholeskySolverBlueprint(coords=P0["nodes"], ctx=ctx))
    first = pkg.condition_on_observations(x, P0["A"], P0["q_eps"], P0["y"])
    p = first.solver_ref[()].precision_chol.p
    bp = pkg.CholeskySolverBlueprint(var_strategy=pkg.RBMCStrategy(600, rng=np.random.default_rng(5)), perm=p, ctx=ctx)
    gc.collect()
    gc.disable()
    try:
        xk, prev = None, None
        for k in range(7):
            Pk = W.darcy_problem(23, seed=k, q_eps=1e4)
            xk = pkg.condition_on_observations(x, Pk["A"], Pk["q_eps"], Pk["y"], solver_blueprint=bp)
            if prev is not None:
                assert prev() is None  # the previous factor went with its GMRF, without the cyclic collector
            prev = weakref.ref(xk.solver_ref[()].precision_chol)
            Qp = orc.posterior_precision(Pk["Q"], Pk["A"], Pk["q_eps"])
            ref = orc.SparseCholesky(Qp, p)
            mref = orc.posterior_mean(ref, Pk["Q"], Pk["A"], Pk["q_eps"], Pk["y"], np.zeros(n))
            assert rel(pkg.mean(xk), mref) < 1e-9
            vref = ref.selinv_diag()
            assert np.median(np.abs(pkg.var(xk) - vref) / vref) < 0.1  # RBMC-600 (panel sweeps, graphed)
            assert rel(xk.solver_ref[()].precision_chol.var_selinv(), vref) < TOL_VAR
    finally:
        gc.enable()


def test_device_gauss_newton_matches_host_loop(pkg, orc, ctx, W):
    """gmrfb_gn_*: the whole Gauss-Newton iteration of scripts/solve_burger.jl:143-180 on the device for the bilinear
    Burgers residual, against (a) the host-driven GaussNewtonOptimizer and (b) the oracle's restated loop."""
    P = W.burgers_spacetime(24, 6)
    noise = 1e4
    n = P["Q"].shape[0]
    pat = orc.posterior_precision(P["Q"], P["f_and_J"](P["mu"])[1], noise)
    p0 = pkg.Symbolic(pat, host_only=True).p
    dgn = pkg.DeviceGaussNewton(P["mu"], P["Q"], P["L"], P["A"], P["D"], P["c"], noise, P["y"], P["mu"],
                                solver_bp=pkg.GNCholeskySolverBlueprint(p0, ctx=ctx))
    xd = dgn.optimize()
    gno = pkg.GaussNewtonOptimizer(P["mu"], P["Q"], P["f_and_J"], noise, P["y"], P["mu"],
                                   solver_bp=pkg.GNCholeskySolverBlueprint(p0, ctx=ctx))
    xh = pkg.optimize(gno)
    assert dgn.n_steps == gno.n_steps and dgn.n_steps >= 2
    assert rel(xd, xh) < 1e-10
    assert rel(np.array(dgn.obj_history[1:]), np.array(gno.obj_history)) < 1e-10
    # oracle loop
    x, Qx = P["mu"].copy(), P["Q"] @ P["mu"]
    for _ in range(dgn.n_steps):
        fx, J = P["f_and_J"](x)
        x = orc.gauss_newton_step(P["Q"], J, noise, x, Qx, P["y"] - fx, p0)
    assert rel(xd, x) < 1e-9
    # state read after the loop: tangent and posterior precision of the last linearisation
    assert abs(dgn.Q_mat - gno.Q_mat).max() < 1e-10 * abs(gno.Q_mat).max()
    assert abs(dgn.Jk - gno.Jk).max() < 1e-12 * abs(gno.Jk).max()


def test_device_gauss_newton_cubic_term(pkg, orc, ctx, W):
    """The elliptic problem -lap u + u^3 = g (_research/elliptic_chen24.jl): f(u) = K u + m .* u.^3 through the cubic
    term of gmrfb_gn_*, against the host-driven optimiser on the same prior."""
    P = W.elliptic_problem(15)
    n = P["n"]
    Qc = (P["Q"] + 1e6 * (P["A_bnd"].T @ P["A_bnd"])).tocsc()  # prior conditioned on the boundary data
    mu = np.zeros(n)
    noise = 1e10
    p0 = pkg.Symbolic(orc.posterior_precision(Qc, P["f_and_J"](mu)[1], noise), host_only=True).p
    dgn = pkg.DeviceGaussNewton(mu, Qc, P["K"], None, None, 0.0, noise, P["y"], mu, cubic=P["m"],
                                solver_bp=pkg.GNCholeskySolverBlueprint(p0, ctx=ctx), max_steps=10, rel_tol=1e-8)
    xd = dgn.optimize()
    gno = pkg.GaussNewtonOptimizer(mu, Qc, P["f_and_J"], noise, P["y"], mu,
                                   solver_bp=pkg.GNCholeskySolverBlueprint(p0, ctx=ctx), max_steps=10, rel_tol=1e-8)
    xh = pkg.optimize(gno)
    assert dgn.n_steps == gno.n_steps and dgn.n_steps >= 2
    assert rel(xd, xh) < 1e-9
    assert rel(xd, P["u_true"]) < 1e-2  # the collocation solution of the manufactured problem


def test_metrics_on_device(pkg, orc, ctx):
    """src/metrics.jl:3-13 (rmse, max_err, rel_err), with and without an evaluation matrix."""
    import scipy.sparse as sp

    rng = np.random.default_rng(0)
    n, m = 5000, 1777
    x, truth = rng.standard_normal(n), rng.standard_normal(m)
    E = sp.random(m, n, density=0.002, random_state=1, format="csc")
    got = pkg.metrics(x, truth, E=pkg.SparseMatrix(E, ctx=ctx))
    want = orc.metrics(E @ x, truth)
    assert np.allclose(got, want, rtol=1e-12)
    t2 = x + 1e-3 * rng.standard_normal(n)
    assert np.allclose(pkg.metrics(x, t2, ctx=ctx), orc.metrics(x, t2), rtol=1e-12)
