#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz.

The reference (timweiland/DiffEqGMRFs.jl) pins no numerical results and cannot run here (no Julia / CHOLMOD), so
these vectors are produced by an implementation that shares no code with either the CUDA library or the C oracle:
dense LAPACK through numpy/scipy (`cholesky`, `solve_triangular`, `inv`) plus a boolean dense symbolic
elimination for the elimination tree and column counts.  For SPD A and a fixed permutation the factor is unique,
so these are the values CHOLMOD / `src/tridiagonal_cholesky.jl` produce up to rounding.

    python tests/golden/make_golden.py        # rewrites the .npz files (deterministic seeds)
"""
import os
import sys

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def dense_symbolic(A, perm):
    """Pattern of chol(A[perm, perm]) by boolean right-looking elimination -> (parent, colcount)."""
    P = (np.asarray(A.todense())[np.ix_(perm, perm)] != 0)
    n = P.shape[0]
    Lp = np.tril(P)
    parent = np.full(n, -1, np.int64)
    cc = np.zeros(n, np.int64)
    for j in range(n):
        rows = np.nonzero(Lp[j + 1:, j])[0] + j + 1
        cc[j] = 1 + rows.size
        if rows.size:
            parent[j] = rows[0]
            Lp[np.ix_(rows, rows)] |= True
            Lp[:, :] = np.tril(Lp)
    return parent, cc


def sparse_case(nx, seed):
    W = entry.load_pkg().workloads
    pkg = entry.load_pkg()
    prob = W.matern_posterior(nx, obs_frac=0.25, q_eps=1e2, corr_range=0.2, seed=seed)
    Q = sp.csc_matrix(prob["Qpost"])
    Q.sort_indices()
    n = Q.shape[0]
    rng = np.random.default_rng(seed)
    # the permutation is an INPUT of the fixture (the reference always passes `perm=p`); taken once from the library's
    # host-side nested dissection and frozen here
    perm = pkg.Symbolic(Q, coords=prob["nodes"], host_only=True).p.copy()
    Ad = np.asarray(Q.todense())
    L = np.linalg.cholesky(Ad[np.ix_(perm, perm)])
    rhs = prob["rhs"]
    z = rng.standard_normal((n, 3))
    mean = np.empty(n)
    y = sla.solve_triangular(L, rhs[perm], lower=True)
    mean[perm] = sla.solve_triangular(L, y, lower=True, trans="T")
    ptl = sla.solve_triangular(L, rhs[perm], lower=True)
    up = np.empty_like(z)
    up[perm] = sla.solve_triangular(L, z, lower=True, trans="T")
    var = np.diag(np.linalg.inv(Ad)).copy()
    parent, cc = dense_symbolic(Q, perm)
    d = Q.diagonal()
    T = Q @ up - d[:, None] * up
    rbmc = 1.0 / d + np.mean(T * T, axis=1) / d ** 2
    return dict(n=n, colptr=Q.indptr.astype(np.int64), rowval=Q.indices.astype(np.int64), nzval=Q.data, perm=perm,
                rhs=rhs, z=z, mean=mean, ptl=ptl, up=up, var=var, rbmc=rbmc, parent=parent, colcount=cc,
                diagL=np.diag(L).copy(), logdet=2 * np.sum(np.log(np.diag(L))))


def btd_case(b, N, seed, extra_rows=0):
    """Block-tridiagonal SPD matrix; factor/solve per src/tridiagonal_cholesky.jl:65-82 semantics via dense LAPACK on
    the assembled matrix (block (i,i) of chol(A) is L_i, block (i+1,i) is C_i)."""
    W = entry.load_pkg().workloads
    D, Bs = W.random_btd(b, N, seed=seed)
    A = sp.csc_matrix(W.btd_to_sparse(D, Bs))
    if extra_rows:
        # trailing rows that `b = n ÷ N_blocks` drops (:66) and an entry outside the block tridiagonal that is ignored
        n = A.shape[0]
        A = sp.bmat([[A, None], [None, 7.0 * sp.identity(extra_rows)]], format="lil")
        if N > 2:
            A[0, 2 * b] = A[2 * b, 0] = 0.123
        A = sp.csc_matrix(A)
    A.sort_indices()
    n = b * N
    At = np.zeros((n, n))
    for i in range(N):
        At[i * b:(i + 1) * b, i * b:(i + 1) * b] = D[:, :, i]
        if i > 0:
            At[i * b:(i + 1) * b, (i - 1) * b:i * b] = Bs[:, :, i - 1]
            At[(i - 1) * b:i * b, i * b:(i + 1) * b] = Bs[:, :, i - 1].T
    L = np.linalg.cholesky(At)
    rng = np.random.default_rng(seed + 100)
    rhs = rng.standard_normal((n, 2))
    fwd = sla.solve_triangular(L, rhs, lower=True)
    bwd = sla.solve_triangular(L, rhs, lower=True, trans="T")
    sol = np.linalg.solve(At, rhs)
    Ls = np.stack([L[i * b:(i + 1) * b, i * b:(i + 1) * b] for i in range(N)], axis=2)
    Cs = np.stack([L[(i + 1) * b:(i + 2) * b, i * b:(i + 1) * b] for i in range(N - 1)], axis=2) if N > 1 else np.zeros((b, b, 0))
    return dict(b=b, N=N, n_total=A.shape[0], colptr=A.indptr.astype(np.int64), rowval=A.indices.astype(np.int64),
                nzval=A.data, rhs=rhs, fwd=fwd, bwd=bwd, sol=sol, Ls=Ls, Cs=Cs, var=np.diag(np.linalg.inv(At)).copy(),
                logdet=2 * np.sum(np.log(np.diag(L))))


def main():
    np.savez_compressed(os.path.join(HERE, "sparse_nx9.npz"), **sparse_case(9, 3))
    np.savez_compressed(os.path.join(HERE, "sparse_nx21.npz"), **sparse_case(21, 5))
    np.savez_compressed(os.path.join(HERE, "btd_b6_N5.npz"), **btd_case(6, 5, 11))
    np.savez_compressed(os.path.join(HERE, "btd_b70_N3_ragged.npz"), **btd_case(70, 3, 12, extra_rows=2))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
