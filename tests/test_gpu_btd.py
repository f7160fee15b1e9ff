"""Parity of the CUDA block-tridiagonal Cholesky against the restatement of src/tridiagonal_cholesky.jl."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.mark.parametrize("b,N", [(1, 1), (1, 5), (7, 3), (64, 4), (65, 3), (130, 5), (257, 3), (400, 2), (600, 3)])
def test_btd_dense_blocks(pkg, orc, ctx, W, b, N):
    D, Bs = W.random_btd(b, N, seed=b * 31 + N)
    A = W.btd_to_sparse(D, Bs)
    F = pkg.tridiagonal_cholesky_dense(D, Bs, ctx=ctx)
    Fo = orc.tridiagonal_cholesky(A, N)
    assert F.N == b * N
    chos, Cs = F.chos, F.Cs
    for i in range(N):
        assert rel(chos[i].L, Fo.chos[i]) < 1e-12
        if i > 0:
            assert rel(Cs[i - 1], Fo.Cs[i - 1]) < 1e-12
    rng = np.random.default_rng(b + N)
    # 1..8 right-hand sides take the bandwidth-bound streaming kernels (one template instance per 1/2/4/8), more
    # than 8 the tile engine
    for nrhs in (1, 2, 3, 5, 8, 11):
        rhs = rng.standard_normal((b * N, nrhs))
        want_f = np.stack([orc.btd_forward_solve(Fo, rhs[:, k]) for k in range(nrhs)], 1)
        want_b = np.stack([orc.btd_backward_solve(Fo, rhs[:, k]) for k in range(nrhs)], 1)
        want_a = np.stack([orc.btd_ldiv(Fo, rhs[:, k]) for k in range(nrhs)], 1)
        assert rel(pkg.forward_solve(F, rhs), want_f) < 1e-10
        assert rel(pkg.backward_solve(F, rhs), want_b) < 1e-10
        assert rel(pkg.ldiv(F, rhs), want_a) < 1e-10
    v = rng.standard_normal(b * N)
    assert rel(pkg.ldiv(F, v), orc.btd_ldiv(Fo, v)) < 1e-10
    assert abs(F.logdet() - orc.btd_logdet(Fo)) < 1e-10 * max(1.0, abs(orc.btd_logdet(Fo)))
    assert np.max(np.abs(F.selinv_diag() - orc.btd_selinv_diag(Fo)) / np.abs(orc.btd_selinv_diag(Fo))) < 1e-8
    assert F.info.status == 0 and F.info.b == b and F.info.nblocks == N


def test_btd_from_sparse_heat_equation(pkg, orc, ctx, W):
    """tridiagonal_cholesky(A::SparseMatrixCSC, N_blocks) on a space-time heat GMRF (config 5 in miniature),
    cross-checked with the sparse supernodal path on the same matrix."""
    hs = W.heat_spacetime(9, 6, dt=1e-2)
    A, N = hs["A"], hs["N"]
    F = pkg.tridiagonal_cholesky(A, N, ctx=ctx)
    Fo = orc.tridiagonal_cholesky(A, N)
    for a, c in zip(F.chos, Fo.chos):
        assert rel(a.L, c) < 1e-11
    rhs = np.random.default_rng(0).standard_normal(A.shape[0])
    x = pkg.ldiv(F, rhs)
    assert rel(x, orc.btd_ldiv(Fo, rhs)) < 1e-10
    assert np.linalg.norm(A @ x - rhs) < 1e-9 * np.linalg.norm(rhs)
    fac = pkg.cholesky(A, ctx=ctx)
    assert rel(fac.solve(rhs), x) < 1e-9
    np.testing.assert_allclose(fac.var_selinv(), F.selinv_diag(), rtol=1e-8)
    # dense-block entry gives the same factor
    F2 = pkg.tridiagonal_cholesky_dense(hs["D"], hs["B"], ctx=ctx)
    assert rel(pkg.ldiv(F2, rhs), x) < 1e-12


def test_btd_remainder_and_offband_ignored(pkg, orc, ctx, W):
    import scipy.sparse as sp

    D, Bs = W.random_btd(12, 3, seed=9)
    A = W.btd_to_sparse(D, Bs).tolil()
    A[30, 2] = A[2, 30] = 5.0  # outside the block tridiagonal: silently ignored (:73-76 read two blocks only)
    Apad = sp.block_diag([A.tocsc(), sp.identity(2) * 3.0]).tocsc()  # 38 rows / 3 blocks -> b = 12, remainder 2
    F = pkg.tridiagonal_cholesky(Apad, 3, ctx=ctx)
    Fo = orc.tridiagonal_cholesky(W.btd_to_sparse(D, Bs), 3)
    assert F.b == 12
    for a, c in zip(F.chos, Fo.chos):
        assert rel(a.L, c) < 1e-12


def test_btd_not_positive_definite(pkg, ctx, W):
    D, Bs = W.random_btd(20, 4, seed=2)
    D[:, :, 2] -= 50.0 * np.eye(20)
    with pytest.raises(pkg.NotPositiveDefinite) as ei:
        pkg.tridiagonal_cholesky_dense(D, Bs, ctx=ctx)
    assert "block 2" in str(ei.value)


@pytest.mark.parametrize("b,N,P", [(48, 9, 2), (70, 11, 3), (33, 16, 4), (64, 5, 1)])
def test_time_sharded_emulated_ranks(pkg, orc, ctx, W, b, N, P):
    """The time-sharded factor/solve with P ranks driven from one process on one GPU (phases called rank by rank,
    the all-gather emulated by concatenation) against the sequential oracle."""
    import torch

    D, Bs = W.random_btd(b, N, seed=b + N + P)
    A = W.btd_to_sparse(D, Bs)
    Fo = orc.tridiagonal_cholesky(A, N)
    bounds = pkg.dist.slab_bounds(N, P)
    ranks = []
    for r, (lo, hi) in enumerate(bounds):
        Dl, Bl = pkg.dist.local_blocks(D, Bs, lo, hi)
        ranks.append(pkg.dist.TimeShardedCholesky(Dl, Bl, r, P, ctx=ctx, auto_exchange=False))
    gathered = torch.cat([t.iface() for t in ranks])
    for t in ranks:
        t.reduce(gathered)
    rhs = np.random.default_rng(1).standard_normal((b * N, 3))
    sends = torch.cat([t.solve_begin(rhs[lo * b:hi * b]) for t, (lo, hi) in zip(ranks, bounds)])
    X = np.vstack([t.solve_end(sends) for t in ranks])
    want = np.stack([orc.btd_ldiv(Fo, rhs[:, k]) for k in range(3)], 1)
    assert rel(X, want) < 1e-10
    import ctypes as C
    tot = 0.0
    red = 0.0
    for t in ranks:
        loc, rd = C.c_double(), C.c_double()
        pkg._lib.check(pkg._lib.lib().gmrfb_btd_dist_logdet(t.h, C.byref(loc), C.byref(rd)), ctx.h)
        tot += loc.value
        red = rd.value
    assert abs(tot + red - orc.btd_logdet(Fo)) < 1e-9 * max(1.0, abs(orc.btd_logdet(Fo)))


def test_btd_solve_ill_conditioned_residual(pkg, orc, ctx, W):
    """The solves apply W_i = L_i^{-1} with one refinement step against L_i: on an ill-conditioned chain (Burgers
    posterior, Q_eps = 1e8) the residual must stay at the level of the substitution-based oracle."""
    P = W.burgers_spacetime(96, 7)
    fx, J = P["f_and_J"](P["mu"])
    A = (P["Q"] + 1e8 * (J.T @ J)).tocsc()
    A.sort_indices()
    F = pkg.tridiagonal_cholesky(A, 7, ctx=ctx)
    Fo = orc.tridiagonal_cholesky(A, 7)
    rhs = np.random.default_rng(5).standard_normal((A.shape[0], 3))
    x = pkg.ldiv(F, rhs)
    xo = np.stack([orc.btd_ldiv(Fo, rhs[:, k]) for k in range(3)], 1)
    res = np.linalg.norm(A @ x - rhs) / np.linalg.norm(rhs)
    res_o = np.linalg.norm(A @ xo - rhs) / np.linalg.norm(rhs)
    assert res < 10 * res_o + 1e-13


def test_btd_ssm_blocks_match_dense_entry(pkg, orc, ctx, W):
    """gmrfb_btd_factor_ssm: the implicit-Euler heat prior (config 5 in miniature) from its four distinct blocks."""
    hs = W.heat_spacetime(7, 6, dt=1e-2)
    D, Bs, N = hs["D"], hs["B"], hs["N"]
    assert all(np.array_equal(D[:, :, t], D[:, :, 1]) for t in range(1, N - 1))
    F = pkg.tridiagonal_cholesky_ssm(D[:, :, 0], D[:, :, 1], D[:, :, N - 1], Bs[:, :, 0], N, ctx=ctx)
    Fo = orc.tridiagonal_cholesky(hs["A"], N)
    rhs = np.random.default_rng(3).standard_normal(hs["A"].shape[0])
    assert rel(pkg.ldiv(F, rhs), orc.btd_ldiv(Fo, rhs)) < 1e-10
    assert abs(F.logdet() - orc.btd_logdet(Fo)) < 1e-10 * abs(orc.btd_logdet(Fo))
