"""The oracle pinned against dense LAPACK (the reference ships no golden vectors: parity unpinned)."""
import numpy as np
import pytest


@pytest.mark.parametrize("nx,seed", [(6, 0), (17, 1), (30, 2)])
def test_sparse_oracle_matches_lapack(orc, W, nx, seed):
    prob = W.matern_posterior(nx, obs_frac=0.3, corr_range=0.3, seed=seed)
    Q = prob["Qpost"]
    n = Q.shape[0]
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n)
    ch = orc.SparseCholesky(Q, perm)
    dch = orc.DenseCholesky(Q, perm)
    B = rng.standard_normal((n, 3))
    for name in ("solve", "solve_UP", "solve_PtL"):
        a, b = getattr(ch, name)(B), getattr(dch, name)(B)
        assert np.linalg.norm(a - b) <= 1e-11 * np.linalg.norm(b), name
    assert abs(ch.logdet() - dch.logdet()) <= 1e-9 * abs(dch.logdet())
    np.testing.assert_allclose(ch.diagL(), dch.diagL(), rtol=1e-11)
    np.testing.assert_allclose(ch.selinv_diag(), orc.dense_inverse_diag(Q), rtol=1e-9)
    # L pattern/values against dense Cholesky of the permuted matrix
    Ld = np.linalg.cholesky(Q.toarray()[np.ix_(perm, perm)])
    np.testing.assert_allclose(ch.L().toarray(), Ld, atol=1e-11 * np.abs(Ld).max())
    # column counts count exactly the structural nonzeros
    assert ch.nnz == int(ch.colcount.sum())


def test_sample_covariance_convention(orc, W):
    """x = F.UP \\ z has covariance Q^{-1} (src/tridiagonal_cholesky.jl:20-22): check P'L^{-T} algebraically."""
    prob = W.matern_posterior(8, obs_frac=0.5, corr_range=0.4)
    Q = prob["Qpost"]
    n = Q.shape[0]
    perm = np.random.default_rng(3).permutation(n)
    ch = orc.SparseCholesky(Q, perm)
    M = ch.solve_UP(np.eye(n))  # M = P' L^{-T}
    np.testing.assert_allclose(M @ M.T, np.linalg.inv(Q.toarray()), rtol=1e-9, atol=1e-12)


def test_not_spd_raises(orc):
    import scipy.sparse as sp

    A = sp.csc_matrix(np.array([[1.0, 2.0], [2.0, 1.0]]))
    with pytest.raises(np.linalg.LinAlgError):
        orc.SparseCholesky(A, np.arange(2))


@pytest.mark.parametrize("b,N", [(1, 1), (5, 4), (16, 7)])
def test_btd_oracle_matches_lapack(orc, W, b, N):
    D, Bs = W.random_btd(b, N, seed=b + N)
    A = W.btd_to_sparse(D, Bs)
    Ad = A.toarray()
    F = orc.tridiagonal_cholesky(A, N)
    assert F.N == b * N and len(F.chos) == N and len(F.Cs) == N - 1
    # the block factor is the Cholesky factor of the assembled matrix
    Lfull = np.linalg.cholesky(Ad)
    for i in range(N):
        np.testing.assert_allclose(F.chos[i], Lfull[i * b:(i + 1) * b, i * b:(i + 1) * b], atol=1e-12)
        if i > 0:
            np.testing.assert_allclose(F.Cs[i - 1], Lfull[i * b:(i + 1) * b, (i - 1) * b:i * b], atol=1e-12)
    rhs = np.random.default_rng(0).standard_normal(b * N)
    np.testing.assert_allclose(orc.btd_ldiv(F, rhs), np.linalg.solve(Ad, rhs), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(orc.btd_forward_solve(F, rhs), np.linalg.solve(Lfull, rhs), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(orc.btd_backward_solve(F, rhs), np.linalg.solve(Lfull.T, rhs), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(orc.btd_selinv_diag(F), np.diag(np.linalg.inv(Ad)), rtol=1e-9)
    assert abs(orc.btd_logdet(F) - np.linalg.slogdet(Ad)[1]) < 1e-9 * max(1.0, abs(orc.btd_logdet(F)))


def test_btd_remainder_rows_ignored(orc, W):
    """b = n div N_blocks; trailing rows are dropped (src/tridiagonal_cholesky.jl:66)."""
    import scipy.sparse as sp

    D, Bs = W.random_btd(4, 3, seed=5)
    A = W.btd_to_sparse(D, Bs)
    Apad = sp.block_diag([A, sp.identity(2) * 7.0]).tocsc()  # 14 rows, 3 blocks -> b = 4, 2 rows ignored
    F = orc.tridiagonal_cholesky(Apad, 3)
    F0 = orc.tridiagonal_cholesky(A, 3)
    for a, b in zip(F.chos, F0.chos):
        np.testing.assert_array_equal(a, b)


def test_rbmc_converges_to_exact(orc, W):
    prob = W.matern_posterior(10, obs_frac=0.5, corr_range=0.3)
    Q = prob["Qpost"]
    n = Q.shape[0]
    ch = orc.SparseCholesky(Q, np.arange(n))
    Z = np.random.default_rng(0).standard_normal((n, 4000))
    v = orc.rbmc_variance(ch, Q, Z)
    np.testing.assert_allclose(v, orc.dense_inverse_diag(Q), rtol=0.1)


def test_posterior_and_gauss_newton_restatement(orc, W):
    prob = W.matern_posterior(9, obs_frac=0.4, corr_range=0.3)
    Qp = orc.posterior_precision(prob["Q"], prob["A"], prob["q_eps"])
    assert abs(Qp - prob["Qpost"]).max() < 1e-9 * abs(Qp).max()
    n = Qp.shape[0]
    ch = orc.SparseCholesky(Qp, np.arange(n))
    mu = orc.posterior_mean(ch, prob["Q"], prob["A"], prob["q_eps"], prob["y"], np.zeros(n))
    ref = np.linalg.solve(Qp.toarray(), prob["rhs"])
    np.testing.assert_allclose(mu, ref, rtol=1e-9, atol=1e-12)
    # linear f => one GN step lands on the posterior mean (scripts/solve_burger.jl:143-149)
    A = prob["A"]
    x0 = np.zeros(n)
    x1 = orc.gauss_newton_step(prob["Q"], A, prob["q_eps"], x0, prob["Q"] @ np.zeros(n), prob["y"] - A @ x0,
                               np.arange(n))
    np.testing.assert_allclose(x1, ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("nx", [9, 40, 97])
def test_oracle_ordering_and_supernodal_symbolic(orc, pkg, W, nx):
    """The CPU baseline's own analysis (oracle/sn_symbolic.c + the NumPy nested dissection): a valid permutation; column
    counts / nnz(L) / flops equal to the scalar oracle's and, bit-exactly, to the product's host analysis of the same
    permutation (two independent implementations of the same integer algorithms); supernodes that tile the columns,
    contain their children's rows, and factor to the scalar oracle's numbers."""
    prob = W.matern_posterior(nx, obs_frac=0.2, corr_range=0.15, seed=nx)
    Q = prob["Qpost"]
    n = Q.shape[0]
    p = orc.nested_dissection(Q, prob["nodes"], leaf=16)
    assert np.array_equal(np.sort(p), np.arange(n))
    sy = orc.supernodal_symbolic(Q, p)
    ref = orc.SparseCholesky(Q, sy["perm_int"])
    assert sy["nnz_L"] == ref.nnz == int(ref.colcount.sum())
    assert np.all(ref.parent[:-1] > np.arange(n - 1)) or n == 1  # postordered: parents follow their children
    sym = pkg.Symbolic(Q, perm=p, host_only=True)
    assert sym.info.nnz_L == sy["nnz_L"] and sym.info.flops == sy["flops"]
    sptr, rptr, rows = sy["sptr"], sy["rptr"], sy["rows"]
    assert sptr[0] == 0 and sptr[-1] == n and np.all(np.diff(sptr) > 0)
    Lpat = ref.L().tocsc()
    for s in range(len(sptr) - 1):
        r = rows[rptr[s]:rptr[s + 1]]
        w = sptr[s + 1] - sptr[s]
        assert np.array_equal(r[:w], np.arange(sptr[s], sptr[s + 1])) and np.all(np.diff(r[w:]) > 0)
        # the structure of every column of the supernode is contained in the supernode's row list
        for j in range(sptr[s], sptr[s + 1]):
            assert set(Lpat.indices[Lpat.indptr[j]:Lpat.indptr[j + 1]]) <= set(r.tolist())
    F = orc.SupernodalCholesky.analyze(Q, perm=p)
    b = np.random.default_rng(0).standard_normal(n)
    assert np.linalg.norm(F.solve(b) - orc.SparseCholesky(Q, p).solve(b)) < 1e-11 * np.linalg.norm(b)
    np.testing.assert_allclose(F.selinv_diag(), orc.SparseCholesky(Q, p).selinv_diag(), rtol=1e-9)


def test_oracle_nd_separators_separate(orc, W):
    """After removing the top separator (the last vertices of the ordering that form one etree chain) the mesh falls
    apart: a structural check of the vertex-cover separators, via the fill they produce against a random ordering."""
    prob = W.matern_posterior(60, obs_frac=0.2, corr_range=0.15, seed=1)
    Q = prob["Qpost"]
    p = orc.nested_dissection(Q, prob["nodes"], leaf=32)
    nd = orc.supernodal_symbolic(Q, p)["nnz_L"]
    rnd = orc.supernodal_symbolic(Q, np.random.default_rng(0).permutation(Q.shape[0]))["nnz_L"]
    nat = orc.supernodal_symbolic(Q, np.arange(Q.shape[0]))["nnz_L"]
    assert nd < 0.65 * nat and nd < 0.2 * rnd  # 60 x 60 mesh: banded (natural) fill is n^1.5, nested dissection n log n


# ------------------------------------------------------------------ FEM tangent restatements (oracle/fem_oracle.py) --
def test_fem_oracle_tangents_are_derivatives_of_residuals(orc, W):
    """The restated element loops are pinned by what they must satisfy independently of any implementation: the
    tangent matrix is the derivative of the residual vector (central differences), constants integrate exactly and
    the quadrature rules integrate their polynomial degree."""
    fo = orc.fem
    rng = np.random.default_rng(0)
    for deg, pw in ((1, 1), (2, 2), (4, 4)):  # int over the reference triangle of l0^a l1^b = a! b! 2 / (a+b+2)! x area
        lam, wq = fo.tri_quadrature(deg)
        assert abs(wq.sum() - 1) < 1e-14
        from math import factorial
        for a in range(pw + 1):
            b = pw - a
            exact = factorial(a) * factorial(b) * 2 / factorial(a + b + 2)
            assert abs(np.sum(wq * lam[:, 0]**a * lam[:, 1]**b) - exact) < 1e-14
    nodes, tris = W.structured_mesh(9, 9, seed=1)
    n = nodes.shape[0]
    w = rng.standard_normal(n)
    bnd = (nodes[:, 0] == 0) | (nodes[:, 0] == 1)
    for deg in (1, 2, 4):
        J, f = fo.assemble_cubic_p1(nodes, tris, w, bnd, deg)
        for k in (5, 40):
            e = np.zeros(n)
            e[k] = 1e-6
            fd = (fo.assemble_cubic_p1(nodes, tris, w + e, bnd, deg)[1] - fo.assemble_cubic_p1(nodes, tris, w - e, bnd, deg)[1]) / 2e-6
            assert abs(fd - J[:, k].toarray().ravel()).max() < 1e-8
        assert abs(J[np.flatnonzero(bnd)]).max() == 0 and np.all(f[bnd] == 0)
    m, K = W.p1_mass_stiffness(nodes, tris)
    J4, f4 = fo.assemble_cubic_p1(nodes, tris, np.full(n, 2.0), None, 4)
    assert abs(f4 - 8 * m).max() < 1e-15 and abs(np.asarray(J4.sum(1)).ravel() - 12 * m).max() < 1e-15
    assert abs(fo.assemble_stiffness_skipped_rows_p1(nodes, tris) - K).max() < 1e-12
    for order in (1, 2):
        xe, el = W.periodic_line_mesh(11, order)
        ns = int(el.max()) + 1
        u = rng.standard_normal(ns)
        G, v = fo.assemble_burgers_advection(xe, el, u, order)
        e = np.zeros(ns)
        e[3] = 1e-6
        fd = (fo.assemble_burgers_advection(xe, el, u + e, order)[1] - fo.assemble_burgers_advection(xe, el, u - e, order)[1]) / 2e-6
        assert abs(fd - G[:, 3].toarray().ravel()).max() < 1e-8
        assert abs(v.sum()) < 1e-12  # int u u_x over a ring = 0 (exact: the rule integrates the integrand's degree)
        M, Gs = fo.assemble_mass_stiffness_1d(xe, el, order)
        assert abs(M.sum() - 1.0) < 1e-14 and abs(Gs @ np.ones(ns)).max() < 1e-11
        Ml, _ = fo.assemble_mass_stiffness_1d(xe, el, order, lumping=True)
        assert abs(Ml.diagonal() - np.asarray(M.sum(1)).ravel()).max() < 1e-15
        nt = 4
        w = rng.standard_normal(nt * ns)
        f, J = fo.burgers_spacetime_tangent(xe, el, w, nt, 0.01, 0.02, order)
        e = np.zeros(nt * ns)
        e[ns + 2] = 1e-6
        fd = (fo.burgers_spacetime_tangent(xe, el, w + e, nt, 0.01, 0.02, order)[0] - fo.burgers_spacetime_tangent(xe, el, w - e, nt, 0.01, 0.02, order)[0]) / 2e-6
        assert abs(fd - J[:, ns + 2].toarray().ravel()).max() < 1e-8


def test_fem_oracle_lagrange_triangles(orc, W):
    """The order-1 / order-2 triangle restatements (tri_cellvalues and the *_lagrange assemblies) pinned by what any
    correct implementation must satisfy: agreement with the independent P1 formulas, exactness of the P2 space on
    quadratics (energies and masses against closed-form integrals), partition of unity on curved (isoparametric)
    elements, the Dunavant degree-3 rule integrating cubics, tangent = derivative of the residual, and the mass
    lumpings conserving the element mass."""
    from math import factorial

    fo = orc.fem
    lam, wq = fo.tri_quadrature(3)
    assert abs(wq.sum() - 1) < 1e-14 and wq[0] < 0
    for a in range(4):
        for b in range(4 - a):
            exact = factorial(a) * factorial(b) * 2 / factorial(a + b + 2)
            assert abs(np.sum(wq * lam[:, 0]**a * lam[:, 1]**b) - exact) < 1e-14
    # shape functions: Kronecker property at the six nodes, partition of unity, gradients sum to zero
    nodes_ref = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [.5, .5, 0], [0, .5, .5], [.5, 0, .5]])
    N, dN = fo.tri_shapes(2, nodes_ref)
    assert abs(N - np.eye(6)).max() < 1e-15
    Nq, dNq = fo.tri_shapes(2, lam)
    assert abs(Nq.sum(1) - 1).max() < 1e-15
    assert abs((dNq[:, :, 1] - dNq[:, :, 0]).sum(1)).max() < 1e-14 and abs((dNq[:, :, 2] - dNq[:, :, 0]).sum(1)).max() < 1e-14

    nodes, tris = W.structured_mesh(9, 7, jitter=0.2, seed=1)
    n6, e6 = W.quadratic_mesh(nodes, tris)
    assert n6.shape[0] == (2 * 9 - 1) * (2 * 7 - 1) and e6.shape == (tris.shape[0], 6)
    assert abs(n6[e6[:, 3]] - 0.5 * (n6[e6[:, 0]] + n6[e6[:, 1]])).max() < 1e-15    # Ferrite's edge numbering
    assert abs(n6[e6[:, 4]] - 0.5 * (n6[e6[:, 1]] + n6[e6[:, 2]])).max() < 1e-15
    assert abs(n6[e6[:, 5]] - 0.5 * (n6[e6[:, 2]] + n6[e6[:, 0]])).max() < 1e-15
    # order 1 through the general code == the independent P1 formulas
    m, Kp = W.p1_mass_stiffness(nodes, tris)
    G1, f1 = fo.assemble_darcy_lagrange(nodes, tris, 1)
    assert abs(G1 - Kp).max() < 1e-13 and abs(f1 - m).max() < 1e-15
    w1 = np.sin(3 * nodes[:, 0]) + nodes[:, 1]
    Jc, fc = fo.assemble_cubic_lagrange(nodes, tris, 1, w1, degree=2)
    Jp, fp = fo.assemble_cubic_p1(nodes, tris, w1, None, 2)
    assert abs(Jc - Jp).max() < 1e-15 and abs(fc - fp).max() < 1e-15
    assert abs(fo.assemble_mass_lagrange(nodes, tris, 1, 1) - m).max() < 1e-15
    # P2 reproduces quadratics: u = x^2 + xy/2 - 2y^2 + x  =>  int |grad u|^2 = 59/6, and the load integrates to the area
    x, y = n6[:, 0], n6[:, 1]
    u = x * x + 0.5 * x * y - 2 * y * y + x
    G2, f2 = fo.assemble_darcy_lagrange(n6, e6, 2)
    assert abs(u @ (G2 @ u) - 59 / 6) < 1e-12 and abs(f2.sum() - 1) < 1e-13 and abs(G2 @ np.ones(n6.shape[0])).max() < 1e-11
    ul = 1 + 2 * x - y                                  # int ul^2 over the unit square = 8/3 (degree-4 integrand)
    M4 = fo.assemble_mass_lagrange(n6, e6, 2, 0, degree=4)
    assert abs(ul @ (M4 @ ul) - 8 / 3) < 1e-12
    for deg in (3, 4):
        for kind in (1, 2):
            ml = fo.assemble_mass_lagrange(n6, e6, 2, kind, degree=deg)
            assert abs(ml.sum() - 1) < 1e-12
        assert fo.assemble_mass_lagrange(n6, e6, 2, 2, degree=deg).min() > 0
    assert abs(fo.assemble_mass_lagrange(n6, e6, 2, 1, degree=4)[:nodes.shape[0]]).max() < 1e-15  # vertex row sums vanish
    # curved elements: area and constants are still exact, the stiffness annihilates constants
    n6c, e6c = W.quadratic_mesh(nodes, tris, curve=0.08, seed=3)
    assert abs(n6c - n6).max() > 1e-3
    Gc, fcu = fo.assemble_darcy_lagrange(n6c, e6c, 2, degree=4)
    assert abs(fcu.sum() - 1) < 1e-12 and abs(Gc @ np.ones(n6c.shape[0])).max() < 1e-11
    # coefficient lookup: first minimum on ties, per quadrature point
    ix, iy = fo.get_xy_idcs(np.array([[0.25, 0.5], [0.0, 1.0]]), [0.0, 0.5, 1.0], [1.0, 0.0])
    assert list(ix) == [0, 0] and list(iy) == [0, 0]
    cm = np.array([[1.0, 2.0], [3.0, 4.0]])
    Gq, _ = fo.assemble_darcy_lagrange(n6, e6, 2, [0.0, 1.0], [0.0, 1.0], cm)
    assert abs(Gq @ np.ones(n6.shape[0])).max() < 1e-11 and abs(Gq - Gq.T).max() < 1e-12
    lo = fo.assemble_darcy_lagrange(n6, e6, 2)[0]
    q = np.random.default_rng(0).standard_normal(n6.shape[0])
    assert 1.0 * (q @ (lo @ q)) < q @ (Gq @ q) < 4.0 * (q @ (lo @ q))
    # tangent = derivative of the residual, rows of prescribed dofs skipped
    bnd = (x == 0) | (x == 1)
    w = np.sin(3 * x) + y
    J, f = fo.assemble_cubic_lagrange(n6, e6, 2, w, bnd, stiffness_scale=0.7)
    d = np.cos(5 * x) * y
    fpl = fo.assemble_cubic_lagrange(n6, e6, 2, w + 1e-6 * d, bnd, stiffness_scale=0.7)[1]
    fmi = fo.assemble_cubic_lagrange(n6, e6, 2, w - 1e-6 * d, bnd, stiffness_scale=0.7)[1]
    assert abs((fpl - fmi) / 2e-6 - J @ d).max() < 1e-7 * abs(J @ d).max()
    assert abs(J[np.flatnonzero(bnd)]).max() == 0 and np.all(f[bnd] == 0)
    # Matern powers: symmetric positive definite (alpha 3 without prescribed dofs: overwriting G[dof, dof] = 1 makes K
    # indefinite, which K' Mt^-1 K tolerates and the odd power does not)
    Q2 = fo.matern_precision_lagrange(n6, e6, 2, 3.0, 0.5, alpha=2, prescribed=bnd)
    Q3 = fo.matern_precision_lagrange(n6, e6, 2, 3.0, 0.5, alpha=3)
    for Q in (Q2, Q3):
        assert abs(Q - Q.T).max() < 1e-9 * abs(Q).max()
        assert np.linalg.eigvalsh(((Q + Q.T) / 2).toarray()).min() > 0


@pytest.mark.parametrize("nx", [151, 301])
def test_oracles_against_superlu_beyond_dense_sizes(orc, W, nx):
    """Tier B of SURVEY.md section 8(c): at sizes dense LAPACK cannot reach (n = 22 801 / 90 601) both CPU oracles - the
    scalar up-looking Cholesky and the supernodal baseline of `bench.py --impl reference` - are compared with SciPy's
    SuperLU, an independent sparse direct solver (LU with partial pivoting and COLAMD, no code or ordering in common):
    posterior mean, log-determinant (sum of log |U_ii|: the L of SuperLU has a unit diagonal) and marginal variances at
    probe nodes (columns of the inverse by unit-vector solves against the Takahashi recurrences)."""
    import scipy.sparse.linalg as spla

    prob = W.matern_posterior(nx, obs_frac=0.1, q_eps=1e2, corr_range=0.1, seed=nx)
    Q = prob["Qpost"].tocsc()
    n = Q.shape[0]
    rng = np.random.default_rng(nx)
    lu = spla.splu(Q)
    B = rng.standard_normal((n, 2))
    Xlu = lu.solve(B)
    logdet_lu = float(np.sum(np.log(np.abs(lu.U.diagonal()))))
    probes = rng.choice(n, size=12, replace=False)
    E = np.zeros((n, probes.size))
    E[probes, np.arange(probes.size)] = 1.0
    var_lu = lu.solve(E)[probes, np.arange(probes.size)]
    F = orc.SupernodalCholesky.analyze(Q, coords=prob["nodes"])
    Xs = F.solve(B)
    assert np.linalg.norm(Xs - Xlu) <= 1e-10 * np.linalg.norm(Xlu)
    assert abs(F.logdet() - logdet_lu) <= 1e-10 * abs(logdet_lu)
    np.testing.assert_allclose(F.selinv_diag()[probes], var_lu, rtol=1e-9)
    if nx <= 151:  # the scalar oracle's Takahashi recurrence is a Python-driven column loop
        p = orc.nested_dissection(Q, prob["nodes"], leaf=32)
        ch = orc.SparseCholesky(Q, p)
        assert np.linalg.norm(ch.solve(B) - Xlu) <= 1e-10 * np.linalg.norm(Xlu)
        assert abs(ch.logdet() - logdet_lu) <= 1e-10 * abs(logdet_lu)
        np.testing.assert_allclose(ch.selinv_diag()[probes], var_lu, rtol=1e-9)
