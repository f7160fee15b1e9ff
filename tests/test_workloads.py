"""Host-side generators of the configurations (diffeqgmrfs.jl_b200/workloads.py): structure the CUDA path relies on."""
import numpy as np
import scipy.sparse as sp


def _chol_ok(A):
    np.linalg.cholesky(A.toarray())


def test_burgers_spacetime_is_block_tridiagonal_spd_with_fixed_tangent_pattern(W):
    nx, nt = 24, 6
    P = W.burgers_spacetime(nx, nt)
    Q = P["Q"]
    assert abs(Q - Q.T).max() == 0
    _chol_ok(Q)
    f0, J0 = P["f_and_J"](P["mu"])
    f1, J1 = P["f_and_J"](P["mu"] + 0.1 * np.random.default_rng(0).standard_normal(nx * nt))
    assert np.array_equal(J0.indptr, J1.indptr) and np.array_equal(J0.indices, J1.indices)  # refactorisable pattern
    A = (Q + 1e8 * (J1.T @ J1)).tocoo()
    assert np.all(abs(A.row // nx - A.col // nx) <= 1)  # what tridiagonal_cholesky reads is the whole matrix
    # the tangent is the derivative of the residual
    w = P["mu"] + 0.05 * np.random.default_rng(1).standard_normal(nx * nt)
    fx, J = P["f_and_J"](w)
    d = 1e-6 * np.random.default_rng(2).standard_normal(nx * nt)
    assert np.linalg.norm(P["f_and_J"](w + d)[0] - fx - J @ d) < 1e-5 * np.linalg.norm(J @ d)


def test_elliptic_manufactured_solution_and_tangent(W):
    P = W.elliptic_problem(21)
    fx, J = P["f_and_J"](P["u_true"])
    assert np.allclose(fx, P["y"], rtol=0, atol=1e-12 * abs(P["y"]).max())
    u = P["u_true"] + 0.01
    d = 1e-6 * np.random.default_rng(0).standard_normal(P["n"])
    f1, J1 = P["f_and_J"](u)
    assert np.linalg.norm(P["f_and_J"](u + d)[0] - f1 - J1 @ d) < 1e-5 * np.linalg.norm(J1 @ d)
    assert P["A_bnd"].shape == (80, 441) and np.array_equal(J.indices, J1.indices)


def test_darcy_dataset_shares_one_pattern(W):
    a, b = W.darcy_problem(15, seed=0), W.darcy_problem(15, seed=1)
    assert np.array_equal(a["A"].indptr, b["A"].indptr) and np.array_equal(a["A"].indices, b["A"].indices)
    assert abs(a["A"] - b["A"]).max() > 0 and set(np.unique(a["coeff_grid"])) == {3.0, 12.0}
    Qp = (a["Q"] + 1e4 * (a["A"].T @ a["A"])).tocsc()
    _chol_ok(Qp)
