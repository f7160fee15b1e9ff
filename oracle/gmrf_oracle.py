"""oracle/gmrf_oracle.py — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's precision-matrix linear-algebra path.  Nothing under
``diffeqgmrfs.jl_b200/`` imports this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

PARITY UNPINNED.  The reference (timweiland/DiffEqGMRFs.jl) pins no numerical results
(test/runtests.jl:5-10 is Aqua lint only), Julia/CHOLMOD/GaussianMarkovRandomFields.jl are not installed
here, and the only hot-path arithmetic whose source is in the reference is src/tridiagonal_cholesky.jl.
The oracle therefore has three tiers, each pinned against something independent:

* Tier A (dense, n <~ 8000): LAPACK through numpy/scipy — ``cho_factor``, triangular solves, ``inv``.
* Tier B (sparse, mid size): ``sparse_chol.c`` — CHOLMOD's published simplicial algorithm
  (Liu etree, row-subtree column counts, up-looking Cholesky, Takahashi recurrences); pinned against
  Tier A in tests/test_oracle.py and, like the supernodal baseline (``supernodal_chol.c``), against SciPy's
  SuperLU (an independent sparse direct solver: LU, COLAMD) at n = 22 801 / 90 601, beyond dense sizes.
* Block-tridiagonal: a line-by-line restatement of src/tridiagonal_cholesky.jl:65-82 (factor) and of the
  *intended* semantics of :24-63 (solves; the three defects documented in SURVEY.md §8a T4/T5/T7 are not
  reproduced), pinned against Tier A on the assembled matrix.

For SPD Q and a fixed permutation, L is unique; means, ``P'L^{-T}z`` samples and ``diag(Q^{-1})`` are
therefore implementation independent up to rounding, which is what makes these tiers a valid oracle.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load_fem():
    """oracle/fem_oracle.py (element-loop restatements of the reference's tangent assemblies) as ``orc.fem``."""
    import importlib.util
    import sys

    if "gmrf_fem_oracle" in sys.modules:
        return sys.modules["gmrf_fem_oracle"]
    spec = importlib.util.spec_from_file_location("gmrf_fem_oracle", os.path.join(_HERE, "fem_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gmrf_fem_oracle"] = mod
    spec.loader.exec_module(mod)
    return mod


fem = _load_fem()
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile sparse_chol.c into oracle/liboracle.so (gcc only)."""
    src = os.path.join(_HERE, "sparse_chol.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB_PATH, src, "-lm"])
    build_supernodal(force)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        _lib.orc_symbolic.argtypes = [ctypes.c_int64, i64p, i64p, i64p, i64p, i64p]
        _lib.orc_symbolic.restype = ctypes.c_int
        _lib.orc_cholesky.argtypes = [ctypes.c_int64, i64p, i64p, f64p, i64p, i64p, i64p, i64p, f64p]
        _lib.orc_cholesky.restype = ctypes.c_int
        _lib.orc_lsolve.argtypes = [ctypes.c_int64, i64p, i64p, f64p, f64p, ctypes.c_int64]
        _lib.orc_ltsolve.argtypes = [ctypes.c_int64, i64p, i64p, f64p, f64p, ctypes.c_int64]
        _lib.orc_selinv.argtypes = [ctypes.c_int64, i64p, i64p, f64p, f64p]
        _lib.orc_selinv.restype = ctypes.c_int
    return _lib


def _csc(A):
    A = sp.csc_matrix(A)
    A.sort_indices()
    return A


# ----------------------------------------------------------------------------------------------- Tier B --
class SparseCholesky:
    """P A P' = L L' with a given permutation (new->old, 0-based) — the CHOLMOD factor views the reference
    uses: ``F \\ b``, ``F.PtL \\ b``, ``F.UP \\ z`` (src/tridiagonal_cholesky.jl:20-22,39-41), ``F.p``,
    ``nnz``, ``diag(F.L)``."""

    def __init__(self, A, perm):
        lib = _load()
        A = _csc(A)
        n = A.shape[0]
        self.n = n
        self.perm = np.ascontiguousarray(perm, dtype=np.int64)
        Ap = A.indptr.astype(np.int64)
        Ai = A.indices.astype(np.int64)
        Ax = A.data.astype(np.float64)
        self.parent = np.empty(n, np.int64)
        self.colcount = np.empty(n, np.int64)
        rc = lib.orc_symbolic(n, Ap, Ai, self.perm, self.parent, self.colcount)
        if rc != 0:
            raise MemoryError("orc_symbolic failed")
        self.Lp = np.zeros(n + 1, np.int64)
        np.cumsum(self.colcount, out=self.Lp[1:])
        self.Li = np.empty(int(self.Lp[-1]), np.int64)
        self.Lx = np.empty(int(self.Lp[-1]), np.float64)
        rc = lib.orc_cholesky(n, Ap, Ai, Ax, self.perm, self.parent, self.Lp, self.Li, self.Lx)
        if rc != 0:
            raise np.linalg.LinAlgError(f"not positive definite at permuted column {rc - 1}")

    @property
    def nnz(self):
        return int(self.Lp[-1])

    @property
    def flops(self):
        return float(np.sum(self.colcount.astype(np.float64) ** 2))

    def L(self):
        return sp.csc_matrix((self.Lx, self.Li, self.Lp), shape=(self.n, self.n))

    def diagL(self):
        return self.Lx[self.Lp[:-1]].copy()

    def logdet(self):
        return 2.0 * float(np.sum(np.log(self.diagL())))

    def _cols(self, B):
        B = np.asarray(B, dtype=np.float64)
        one = B.ndim == 1
        X = np.ascontiguousarray(B.reshape(self.n, -1).T).copy()  # rows = right-hand sides
        return X, one

    def solve_PtL(self, B):
        """L^{-1} P b  (``F.PtL \\ b``)."""
        X, one = self._cols(B)
        X = np.ascontiguousarray(X[:, self.perm])
        _load().orc_lsolve(self.n, self.Lp, self.Li, self.Lx, X, X.shape[0])
        return X[0] if one else X.T.copy()

    def solve_UP(self, Z):
        """P' L^{-T} z  (``F.UP \\ z``) — a N(0, A^{-1}) sample for z ~ N(0, I)."""
        X, one = self._cols(Z)
        _load().orc_ltsolve(self.n, self.Lp, self.Li, self.Lx, X, X.shape[0])
        out = np.empty_like(X)
        out[:, self.perm] = X
        return out[0] if one else out.T.copy()

    def solve(self, B):
        """A^{-1} b  (``F \\ b``)."""
        X, one = self._cols(B)
        X = np.ascontiguousarray(X[:, self.perm])
        lib = _load()
        lib.orc_lsolve(self.n, self.Lp, self.Li, self.Lx, X, X.shape[0])
        lib.orc_ltsolve(self.n, self.Lp, self.Li, self.Lx, X, X.shape[0])
        out = np.empty_like(X)
        out[:, self.perm] = X
        return out[0] if one else out.T.copy()

    def selinv_diag(self):
        """diag(A^{-1}) in the original ordering by the Takahashi recurrences on L's pattern."""
        Zx = np.zeros_like(self.Lx)
        rc = _load().orc_selinv(self.n, self.Lp, self.Li, self.Lx, Zx)
        if rc != 0:
            raise RuntimeError("orc_selinv failed")
        d = np.empty(self.n)
        d[self.perm] = Zx[self.Lp[:-1]]
        return d

    def selinv(self):
        """Selected inverse on the pattern of L (permuted ordering), as a CSC lower-triangular matrix."""
        Zx = np.zeros_like(self.Lx)
        rc = _load().orc_selinv(self.n, self.Lp, self.Li, self.Lx, Zx)
        if rc != 0:
            raise RuntimeError("orc_selinv failed")
        return sp.csc_matrix((Zx, self.Li, self.Lp), shape=(self.n, self.n))


def symbolic(A, perm):
    """(parent, colcount) of P A P' for a given perm (0-based, new->old); root parent = -1."""
    lib = _load()
    A = _csc(A)
    n = A.shape[0]
    parent = np.empty(n, np.int64)
    colcount = np.empty(n, np.int64)
    rc = lib.orc_symbolic(n, A.indptr.astype(np.int64), A.indices.astype(np.int64),
                          np.ascontiguousarray(perm, dtype=np.int64), parent, colcount)
    if rc != 0:
        raise MemoryError
    return parent, colcount


# ----------------------------------------------------------------------------------------------- Tier A --
class DenseCholesky:
    """LAPACK restatement of the same factor views for small n."""

    def __init__(self, A, perm):
        A = np.asarray(A.todense() if sp.issparse(A) else A, dtype=np.float64)
        self.perm = np.asarray(perm, dtype=np.int64)
        self.n = A.shape[0]
        self.Lmat = np.linalg.cholesky(A[np.ix_(self.perm, self.perm)])

    def solve_PtL(self, B):
        B = np.asarray(B, dtype=np.float64)
        return sla.solve_triangular(self.Lmat, B[self.perm], lower=True)

    def solve_UP(self, Z):
        X = sla.solve_triangular(self.Lmat, np.asarray(Z, dtype=np.float64), lower=True, trans="T")
        out = np.empty_like(X)
        out[self.perm] = X
        return out

    def solve(self, B):
        B = np.asarray(B, dtype=np.float64)
        Y = sla.solve_triangular(self.Lmat, B[self.perm], lower=True)
        X = sla.solve_triangular(self.Lmat, Y, lower=True, trans="T")
        out = np.empty_like(X)
        out[self.perm] = X
        return out

    def logdet(self):
        return 2.0 * float(np.sum(np.log(np.diag(self.Lmat))))

    def diagL(self):
        return np.diag(self.Lmat).copy()


def dense_inverse_diag(A):
    A = np.asarray(A.todense() if sp.issparse(A) else A, dtype=np.float64)
    return np.diag(np.linalg.inv(A)).copy()


# ------------------------------------------------------------------------------------ GMRF-level helpers --
def posterior_precision(Q, A, qeps):
    """Q + A' diag(qeps) A — the assembly inside condition_on_observations
    (scripts/darcy/solve_darcy_gmrf-fem.jl:165-167) / ``Q + noise*J'*J`` (scripts/solve_burger.jl:145)."""
    A = sp.csc_matrix(A)
    W = sp.diags(np.broadcast_to(np.asarray(qeps, dtype=np.float64), (A.shape[0],)))
    return _csc(sp.csc_matrix(Q) + A.T @ W @ A)


def posterior_mean(chol, Q, A, qeps, y, mu):
    """mu + Qpost^{-1} A' Q_eps (y - A mu): standard Gaussian conditioning (GMRF.jl ``mean`` of a
    conditioned GMRF; formula not in the reference tree — 'vs. restated oracle')."""
    A = sp.csc_matrix(A)
    w = np.broadcast_to(np.asarray(qeps, dtype=np.float64), (A.shape[0],))
    return mu + chol.solve(A.T @ (w * (y - A @ mu)))


def rbmc_variance(chol, Q, Z):
    """Rao-Blackwellised Monte-Carlo marginal variances, RBMCStrategy(N)
    (scripts/darcy/solve_darcy_gmrf-fem.jl:100,174,192); estimator of Siden et al. (2018):
    var_i = 1/Q_ii + mean_k (sum_{j != i} Q_ij x_j^(k))^2 / Q_ii^2 with x^(k) = F.UP \\ z^(k)."""
    Q = sp.csr_matrix(Q)
    X = chol.solve_UP(Z)
    d = Q.diagonal()
    T = Q @ X - d[:, None] * X
    return 1.0 / d + np.mean(T * T, axis=1) / d ** 2


def gauss_newton_step(Q, J, noise, x, Qx_prior, obs_diff, perm, chol_cls=SparseCholesky):
    """One step of scripts/solve_burger.jl:143-149:
    A = Q + noise J'J;  rhs = Q x_prior + noise J'(J x + obs_diff);  x+ = A \\ rhs (fixed perm)."""
    J = sp.csc_matrix(J)
    A = _csc(sp.csc_matrix(Q) + noise * (J.T @ J))
    rhs = Qx_prior + noise * (J.T @ (J @ x + obs_diff))
    return chol_cls(A, perm).solve(rhs)


def metrics(pred, truth):
    """src/metrics.jl:3-13."""
    pred = np.asarray(pred)
    truth = np.asarray(truth)
    rmse = float(np.sqrt(np.mean((pred - truth) ** 2)))
    max_err = float(np.max(np.abs(pred - truth)))
    rel_err = float(np.linalg.norm(pred - truth) / np.linalg.norm(truth))
    return rmse, max_err, rel_err


# ------------------------------------------------------------------------------------ block tridiagonal --
class TridiagonalCholeskyFactor:
    """src/tridiagonal_cholesky.jl:5-9 — N (total rows), chos (dense lower factors), Cs (sub-diagonal blocks;
    Cs[k] belongs to block row k+1)."""

    def __init__(self, N, chos, Cs):
        self.N = N
        self.chos = chos
        self.Cs = Cs


def tridiagonal_cholesky(A, n_blocks):
    """src/tridiagonal_cholesky.jl:65-82, line by line (dense LAPACK blocks)."""
    A = sp.csc_matrix(A)
    b = A.shape[0] // n_blocks                                   # :66 (remainder rows ignored)
    chos = [np.linalg.cholesky(A[:b, :b].toarray())]              # :67
    Cs = []
    for i in range(1, n_blocks):                                  # :70
        r0, r1 = i * b, (i + 1) * b
        B = A[r0:r1, (i - 1) * b:i * b].toarray()                 # :73
        C = sla.solve_triangular(chos[-1], B.T, lower=True).T     # :74  C = B L^{-T}
        Cs.append(C)
        D = A[r0:r1, r0:r1].toarray()                             # :76
        chos.append(np.linalg.cholesky(D - C @ C.T))              # :77
    return TridiagonalCholeskyFactor(A.shape[0], chos, Cs)


def btd_forward_solve(F, b):
    """Intended semantics of src/tridiagonal_cholesky.jl:43-52: x_1 = L_1^{-1} b_1; x_i = L_i^{-1}(b_i - C_i x_{i-1})."""
    nb = len(F.chos)
    bs = F.chos[0].shape[0]
    b = np.asarray(b, dtype=np.float64)
    x = b[: nb * bs].copy()
    for i in range(nb):
        rhs = x[i * bs:(i + 1) * bs]
        if i > 0:
            rhs = rhs - F.Cs[i - 1] @ x[(i - 1) * bs:i * bs]
        x[i * bs:(i + 1) * bs] = sla.solve_triangular(F.chos[i], rhs, lower=True)
    return x


def btd_backward_solve(F, b):
    """Intended semantics of :24-33: x_N = L_N^{-T} b_N; x_i = L_i^{-T}(b_i - C_{i+1}' x_{i+1})."""
    nb = len(F.chos)
    bs = F.chos[0].shape[0]
    b = np.asarray(b, dtype=np.float64)
    x = b[: nb * bs].copy()
    for i in range(nb - 1, -1, -1):
        rhs = x[i * bs:(i + 1) * bs]
        if i < nb - 1:
            rhs = rhs - F.Cs[i].T @ x[(i + 1) * bs:(i + 2) * bs]
        x[i * bs:(i + 1) * bs] = sla.solve_triangular(F.chos[i], rhs, lower=True, trans="T")
    return x


def btd_ldiv(F, b):
    """Intended semantics of :54-63: A^{-1} b = backward(forward(b))."""
    return btd_backward_solve(F, btd_forward_solve(F, b))


def btd_logdet(F):
    return 2.0 * float(sum(np.sum(np.log(np.diag(L))) for L in F.chos))


def btd_selinv_diag(F):
    """diag(A^{-1}) by the block Takahashi recursion S_N = L_N^{-T}L_N^{-1},
    S_i = L_i^{-T}(I + C_{i+1}' S_{i+1} C_{i+1}) L_i^{-1}."""
    nb = len(F.chos)
    bs = F.chos[0].shape[0]
    out = np.empty(nb * bs)
    S = None
    for i in range(nb - 1, -1, -1):
        H = np.eye(bs)
        if i < nb - 1:
            C = F.Cs[i]
            H = H + C.T @ S @ C
        Li = sla.solve_triangular(F.chos[i], np.eye(bs), lower=True)
        S = Li.T @ H @ Li
        out[i * bs:(i + 1) * bs] = np.diag(S)
    return out


# ------------------------------------------------------------------- supernodal CPU baseline (BLAS-3) --
_SN_PATH = os.path.join(_HERE, "libsupernodal.so")
_sn = None


def _capsule_ptr(mod, name):
    cap = mod.__pyx_capi__[name]
    ctypes.pythonapi.PyCapsule_GetName.restype = ctypes.c_char_p
    ctypes.pythonapi.PyCapsule_GetName.argtypes = [ctypes.py_object]
    ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
    ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    return ctypes.pythonapi.PyCapsule_GetPointer(cap, ctypes.pythonapi.PyCapsule_GetName(cap))


def build_supernodal(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "supernodal_chol.c"), os.path.join(_HERE, "sn_symbolic.c")]
    if force or not os.path.exists(_SN_PATH) or os.path.getmtime(_SN_PATH) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SN_PATH, *srcs, "-lm"])
    return _SN_PATH


def _load_sn():
    """libsupernodal.so with OpenBLAS/LAPACK entry points taken from scipy's Cython capsules."""
    global _sn
    if _sn is None:
        import scipy.linalg.cython_blas as cb
        import scipy.linalg.cython_lapack as cl

        build_supernodal()
        L = ctypes.CDLL(_SN_PATH)
        i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.sn_set_blas.argtypes = [ctypes.c_void_p] * 9
        L.sn_set_blas(*[_capsule_ptr(cb, k) for k in ("dgemm", "dsyrk", "dtrsm", "dtrsv", "dgemv", "dsymm")],
                      *[_capsule_ptr(cl, k) for k in ("dpotrf", "dtrtri", "dlauum")])
        I = ctypes.c_int64
        L.sn_factor.argtypes = [I, I, i64p, i64p, i64p, i64p, i64p, i64p, f64p, i64p, f64p]
        L.sn_factor.restype = I
        L.sn_solve.argtypes = [I, I, i64p, i64p, i64p, i64p, f64p, f64p, I, ctypes.c_int, ctypes.c_int]
        L.sn_selinv.argtypes = [I, I, i64p, i64p, i64p, i64p, i64p, f64p, f64p]
        L.sn_selinv.restype = I
        L.orc_sn_symbolic.argtypes = [I, i64p, i64p, i64p, i64p]
        L.orc_sn_symbolic.restype = I
        L.orc_sn_symbolic_nrows.restype = I
        L.orc_sn_symbolic_nnzL.restype = I
        L.orc_sn_symbolic_flops.restype = ctypes.c_double
        L.orc_sn_symbolic_fetch.argtypes = [i64p, i64p, i64p]
        _sn = L
    return _sn


def _min_vertex_cover(u, v):
    """Minimum vertex cover of the bipartite graph with edges (u[k], v[k]) (u: left labels, v: right labels)."""
    from scipy.sparse.csgraph import maximum_bipartite_matching

    if u.size == 0:
        return np.zeros(0, np.int64)
    Lset, li = np.unique(u, return_inverse=True)
    Rset, ri = np.unique(v, return_inverse=True)
    B = sp.csr_matrix((np.ones(li.size, np.int8), (li, ri)), shape=(Lset.size, Rset.size))
    B.sum_duplicates()
    match_l = maximum_bipartite_matching(B, perm_type="column")  # for every left vertex its right partner or -1
    match_r = np.full(Rset.size, -1, np.int64)
    ok = match_l >= 0
    match_r[match_l[ok]] = np.flatnonzero(ok)
    # alternating reachability from the unmatched left vertices: left -> right along any edge, right -> left along
    # matching edges
    vis_l = np.zeros(Lset.size, bool)
    vis_r = np.zeros(Rset.size, bool)
    frontier = np.flatnonzero(~ok)
    vis_l[frontier] = True
    indptr, indices = B.indptr, B.indices
    while frontier.size:
        cnt = indptr[frontier + 1] - indptr[frontier]
        if cnt.sum() == 0:
            break
        pos = np.repeat(indptr[frontier], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        nr = np.unique(indices[pos])
        nr = nr[~vis_r[nr]]
        vis_r[nr] = True
        nl = match_r[nr]
        nl = nl[nl >= 0]
        nl = nl[~vis_l[nl]]
        vis_l[nl] = True
        frontier = nl
    return np.concatenate([Lset[~vis_l], Rset[vis_r]]).astype(np.int64)


def nested_dissection(A, coords, leaf: int = 64):
    """Geometric nested dissection (new->old permutation) for the CPU baseline, independent of the product's
    orderings: level-synchronous recursive coordinate bisection at the median of the wider axis; the separator of a
    cut is a minimum vertex cover of the cut edges (maximum bipartite matching + Koenig).  Leaf regions first, then the
    separators from the deepest level up to the top one (the kind of ordering CHOLMOD obtains from METIS when the
    reference calls `cholesky(A)` on a mesh matrix)."""
    A = _csc(A)
    n = A.shape[0]
    coords = np.asarray(coords, dtype=np.float64)
    C = A.tocoo()
    ei, ej = C.row[C.row != C.col], C.col[C.row != C.col]
    region = np.zeros(n, np.int64)      # current region of every vertex still inside a region
    sep_level = np.full(n, -1, np.int64)  # level at which a vertex became a separator vertex (-1: none yet)
    active = np.ones(n, bool)
    level = 0
    while True:
        idx = np.flatnonzero(active)
        if idx.size == 0:
            break
        reg = region[idx]
        sizes = np.bincount(reg)
        big = sizes[reg] > leaf
        if not big.any():
            break
        idx, reg = idx[big], reg[big]
        # wider axis per region, median split
        nreg = int(reg.max()) + 1
        ext = np.zeros((nreg, coords.shape[1]))
        for a in range(coords.shape[1]):
            lo = np.full(nreg, np.inf)
            hi = np.full(nreg, -np.inf)
            np.minimum.at(lo, reg, coords[idx, a])
            np.maximum.at(hi, reg, coords[idx, a])
            ext[:, a] = hi - lo
        axis = np.argmax(ext, axis=1)
        key = coords[idx, axis[reg]]
        order = np.lexsort((key, reg))
        ridx, rreg = idx[order], reg[order]
        start = np.flatnonzero(np.r_[True, rreg[1:] != rreg[:-1]])
        cnt = np.diff(np.r_[start, rreg.size])
        rank = np.arange(rreg.size) - np.repeat(start, cnt)
        side = np.zeros(n, np.int8)  # 1: low half, 2: high half (of a region being split at this level)
        low = rank < np.repeat(cnt // 2, cnt)
        side[ridx[low]] = 1
        side[ridx[~low]] = 2
        # separator: a minimum vertex cover of the cut edges (low side <-> high side inside one region), by maximum
        # bipartite matching + Koenig's theorem; all regions of the level at once (their cut graphs are disjoint)
        m = (side[ei] == 1) & (side[ej] == 2) & (region[ei] == region[ej]) & active[ei] & active[ej]
        sep = _min_vertex_cover(ei[m], ej[m])
        sep_level[sep] = level
        active[sep] = False
        side[sep] = 0
        # children regions: 2 * region + (0 | 1) for the vertices of split regions; unsplit regions become inactive leaves
        split = side > 0
        region[split] = 2 * region[split] + (side[split] == 2)
        small = active & ~split & (sep_level < 0)
        small_idx = np.flatnonzero(small)
        # leaves keep their region id but must not collide with the renumbered ones: tag them with the level
        region[small_idx] = -(region[small_idx] * 64 + level) - 1
        active[small_idx] = False
        # compress region ids of the active vertices
        act = np.flatnonzero(active)
        if act.size:
            _, region[act] = np.unique(region[act], return_inverse=True)
        level += 1
    # elimination order: leaves (grouped by region), then separators by descending level (grouped by coordinates order)
    lvl_key = np.where(sep_level < 0, level + 1, sep_level)
    perm = np.lexsort((region, -lvl_key))
    return perm.astype(np.int64)


def supernodal_symbolic(A, perm):
    """Postordered permutation, supernode partition and supernodal row structures for `perm` (oracle/sn_symbolic.c):
    -> dict(perm_int, sptr, rptr, rows, nnz_L, flops)."""
    L = _load_sn()
    A = _csc(A)
    n = A.shape[0]
    Ap, Ai = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    perm = np.ascontiguousarray(perm, dtype=np.int64)
    perm_int = np.empty(n, np.int64)
    ns = L.orc_sn_symbolic(n, Ap, Ai, perm, perm_int)
    if ns < 0:
        raise MemoryError(f"orc_sn_symbolic failed ({ns})")
    sptr, rptr = np.empty(ns + 1, np.int64), np.empty(ns + 1, np.int64)
    rows = np.empty(max(int(L.orc_sn_symbolic_nrows()), 1), np.int64)
    nnzL, flops = int(L.orc_sn_symbolic_nnzL()), float(L.orc_sn_symbolic_flops())
    L.orc_sn_symbolic_fetch(sptr, rptr, rows)
    return dict(perm_int=perm_int, sptr=sptr, rptr=rptr, rows=rows[: int(rptr[-1])], nnz_L=nnzL, flops=flops)


class SupernodalCholesky:
    """CPU supernodal multifrontal Cholesky with BLAS-3 supernodes (CHOLMOD's algorithm class; the timed CPU
    baseline of bench.py).  The supernode partition is an input: ``perm_int`` (internal new->old, 0-based, an etree
    postorder of the user's perm), ``sptr`` (first internal column of each supernode) and per-supernode row lists
    (own columns first, then the below rows ascending) in the internal numbering."""

    def __init__(self, A, perm_int, sptr, rows_list, rptr=None):
        """``rows_list``: per-supernode row lists, or (with ``rptr``) the flat row array of `supernodal_symbolic`."""
        A = _csc(A)
        self.n = n = A.shape[0]
        self.perm = np.ascontiguousarray(perm_int, dtype=np.int64)
        self.sptr = np.ascontiguousarray(sptr, dtype=np.int64)
        self.ns = ns = len(self.sptr) - 1
        if rptr is not None:
            self.rptr = np.ascontiguousarray(rptr, dtype=np.int64)
            self.rows = np.ascontiguousarray(rows_list, dtype=np.int64)
        else:
            self.rptr = np.zeros(ns + 1, np.int64)
            np.cumsum([len(r) for r in rows_list], out=self.rptr[1:])
            self.rows = np.ascontiguousarray(np.concatenate(rows_list) if ns else np.zeros(0), dtype=np.int64)
        sc = np.diff(self.sptr)
        d = np.diff(self.rptr)
        snode = np.repeat(np.arange(ns, dtype=np.int64), sc)
        self.sparent = np.full(ns, -1, np.int64)
        has = d > sc
        self.sparent[has] = snode[self.rows[(self.rptr[:-1] + sc)[has]]]
        self.loff = np.zeros(ns + 1, np.int64)
        np.cumsum(d * sc, out=self.loff[1:])
        self.Lx = np.zeros(int(self.loff[-1]))
        # tril(P A P') once, with the map from A's stored values to its entries (numeric refactorisations reuse it)
        tag = sp.csc_matrix((np.arange(1, A.nnz + 1, dtype=np.float64), A.indices, A.indptr), shape=A.shape)
        C = sp.tril(tag[self.perm][:, self.perm], format="csc")
        C.sort_indices()
        self._Cp = C.indptr.astype(np.int64)
        self._Ci = C.indices.astype(np.int64)
        self._vmap = (C.data - 1).astype(np.int64)
        self.refactor(A)

    def refactor(self, A):
        """Numeric factorisation of a matrix with the analysed pattern (A: scipy sparse, or its CSC value array)."""
        vals = _csc(A).data if sp.issparse(A) else np.asarray(A, dtype=np.float64)
        Cx = np.ascontiguousarray(vals[self._vmap], dtype=np.float64)
        rc = _load_sn().sn_factor(self.n, self.ns, self.sptr, self.rptr, self.rows, self.sparent, self._Cp, self._Ci, Cx,
                                  self.loff, self.Lx)
        if rc != 0:
            raise np.linalg.LinAlgError(f"supernodal factorisation failed ({rc})")
        return self

    def _sweep(self, B, fwd, bwd, perm_in, perm_out):
        B = np.asarray(B, dtype=np.float64)
        one = B.ndim == 1
        X = np.asfortranarray(B.reshape(self.n, -1).copy())
        if perm_in:
            X = np.asfortranarray(X[self.perm])
        Xf = np.ascontiguousarray(X.T)  # rows = right-hand sides = column-major n x nrhs
        _load_sn().sn_solve(self.n, self.ns, self.sptr, self.rptr, self.rows, self.loff, self.Lx, Xf, Xf.shape[0],
                            int(fwd), int(bwd))
        X = Xf.T
        if perm_out:
            out = np.empty_like(X)
            out[self.perm] = X
            X = out
        return X[:, 0].copy() if one else np.array(X)

    def solve(self, B):
        return self._sweep(B, True, True, True, True)

    def selinv_diag(self):
        z = np.zeros(self.n)
        rc = _load_sn().sn_selinv(self.n, self.ns, self.sptr, self.rptr, self.rows, self.sparent, self.loff, self.Lx, z)
        if rc != 0:
            raise MemoryError("sn_selinv failed")
        out = np.empty(self.n)
        out[self.perm] = z
        return out

    @classmethod
    def analyze(cls, A, coords=None, perm=None, leaf=64):
        """Order (own nested dissection, or a given perm), analyse (oracle/sn_symbolic.c) and factorise: the CPU
        baseline path without any of the product's host code."""
        if perm is None:
            perm = nested_dissection(A, coords, leaf)
        sy = supernodal_symbolic(A, perm)
        F = cls(A, sy["perm_int"], sy["sptr"], sy["rows"], rptr=sy["rptr"])
        F.nnz_L, F.flops = sy["nnz_L"], sy["flops"]
        return F

    def logdet(self):
        sc = np.diff(self.sptr)
        d = np.diff(self.rptr)
        tot = 0.0
        for s in range(self.ns):
            P = self.Lx[self.loff[s]:self.loff[s + 1]].reshape(sc[s], d[s])  # column-major d x sc == C-order sc x d
            tot += float(np.sum(np.log(np.diag(P[:, :sc[s]]))))
        return 2.0 * tot
