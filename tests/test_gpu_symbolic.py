"""The CUDA phases of the symbolic analysis (csrc/symbolic_gpu.cu: pattern validation + adjacency, permuted / internal
adjacency, relmap, amap) against the host loops they replace: every integer output bit for bit, for every ordering."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def same_analysis(a, b):
    assert np.array_equal(a.p, b.p) and np.array_equal(a.parent, b.parent) and np.array_equal(a.colcount, b.colcount)
    assert np.array_equal(a.super_ptr, b.super_ptr) and np.array_equal(a.ipost, b.ipost)
    ia, ib = a.info, b.info
    for f in ("n", "nnz_lower_A", "nnz_L", "nnz_L_stored", "flops", "nsuper", "nlevels", "max_front", "front_bytes"):
        assert getattr(ia, f) == getattr(ib, f), f
    for s in range(0, ia.nsuper, max(1, ia.nsuper // 200)):
        assert np.array_equal(a.super_rows(s), b.super_rows(s))
    am_a, rm_a = a.maps()
    am_b, rm_b = b.maps()
    assert np.array_equal(am_a, am_b) and np.array_equal(rm_a, rm_b)


@pytest.mark.parametrize("nx", [3, 24, 130])
@pytest.mark.parametrize("ordering", ["nd", "geo", "amd", "natural", "given"])
def test_gpu_analysis_equals_host_analysis(pkg, ctx, W, nx, ordering):
    prob = W.matern_posterior(nx, obs_frac=0.2, corr_range=0.15, seed=nx)
    Q = prob["Qpost"]
    kw = {}
    if ordering == "geo":
        kw = dict(coords=prob["nodes"])
    elif ordering == "given":
        kw = dict(perm=np.random.default_rng(nx).permutation(Q.shape[0]))
    elif ordering != "nd":
        kw = dict(ordering=ordering)
    same_analysis(pkg.Symbolic(Q, ctx=ctx, **kw), pkg.Symbolic(Q, host_only=True, **kw))


def test_gpu_analysis_long_rows_and_unsymmetric_patterns(pkg, ctx):
    """An arrow matrix (one row longer than the per-thread sort limit: sorted on the host side of the GPU path) and a
    structurally unsymmetric pattern (falls back to the host symmetrisation): same results as host_only."""
    n = 2000
    rng = np.random.default_rng(0)
    A = sp.random(n, n, density=2.0 / n, random_state=1, format="lil")
    A = (A + A.T).tolil()
    A[0, :] = 1.0
    A[:, 0] = 1.0
    A.setdiag(10.0 + np.arange(n))
    A = sp.csc_matrix(A)
    A.sort_indices()
    same_analysis(pkg.Symbolic(A, ctx=ctx, ordering="amd"), pkg.Symbolic(A, host_only=True, ordering="amd"))
    same_analysis(pkg.Symbolic(A, ctx=ctx, ordering="natural"), pkg.Symbolic(A, host_only=True, ordering="natural"))
    U = sp.csc_matrix(sp.triu(A) + sp.identity(n))  # only one triangle stored under STORAGE_FULL: unsymmetric pattern
    U.sort_indices()
    same_analysis(pkg.Symbolic(U, ctx=ctx, ordering="amd"), pkg.Symbolic(U, host_only=True, ordering="amd"))
    # the numeric path on the GPU-analysed arrow matrix
    fac = pkg.CholeskyFactor(pkg.Symbolic(A, ctx=ctx, ordering="amd")).factorize(A.data)
    b = rng.standard_normal(n)
    assert np.linalg.norm(A @ fac.solve(b) - b) < 1e-10 * np.linalg.norm(b)


def test_gpu_analysis_rejects_malformed_patterns(pkg, ctx):
    A = sp.identity(5, format="csc") * 2.0
    bad = A.copy()
    bad.indices = bad.indices.copy()
    bad.indices[2] = 7  # out of range
    with pytest.raises(pkg.GmrfbError):
        _raw_analyze(pkg, ctx, bad)


def _raw_analyze(pkg, ctx, A):
    import ctypes as C

    B = pkg._lib
    opts = B.AnalyzeOpts()
    opts.ordering_kind = B.ORDER_NATURAL
    _, cp = B.i64(A.indptr)
    _, ri = B.i64(A.indices)
    h = C.c_void_p()
    st = B.lib().gmrfb_analyze(ctx.h, A.shape[0], cp, ri, None, C.byref(opts), C.byref(h))
    B.check(st, ctx.h)
