"""Host-side symbolic analysis of the library against the oracle: with the same permutation the elimination
tree, column counts and nnz(L) must be bit-exact (SURVEY.md §7 'hard parts')."""
import numpy as np
import pytest
import scipy.sparse as sp


def _check(pkg, orc, A, perm=None, **kw):
    sym = pkg.Symbolic(A, perm=perm, host_only=True, **kw)
    n = A.shape[0]
    p = sym.p
    assert sorted(p.tolist()) == list(range(n))
    if perm is not None:
        assert np.array_equal(p, perm)  # Julia: F.p == perm
    parent, cc = orc.symbolic(A, p)
    assert np.array_equal(sym.parent, parent)
    assert np.array_equal(sym.colcount, cc)
    info = sym.info
    assert info.nnz_L == int(cc.sum())
    assert info.flops == float(np.sum(cc.astype(float) ** 2))
    assert info.nnz_L_stored >= info.nnz_L
    # supernodes partition the columns; row structures contain exactly the columns first
    sptr = sym.super_ptr
    assert sptr[0] == 0 and sptr[-1] == n and np.all(np.diff(sptr) > 0)
    ipost = sym.ipost
    assert sorted(ipost.tolist()) == list(range(n))
    # ipost is an etree postorder: every parent comes after its children
    for k in range(n):
        if parent[k] >= 0:
            assert ipost[parent[k]] > ipost[k]
    return sym


@pytest.mark.parametrize("nx", [2, 5, 13, 40])
def test_nd_ordering_mesh(pkg, orc, W, nx):
    prob = W.matern_posterior(nx, obs_frac=0.2, corr_range=0.3)
    _check(pkg, orc, prob["Qpost"])
    _check(pkg, orc, prob["Qpost"], coords=prob["nodes"])


def test_given_and_natural_perm(pkg, orc, W):
    prob = W.matern_posterior(12, obs_frac=0.2, corr_range=0.3)
    Q = prob["Qpost"]
    n = Q.shape[0]
    _check(pkg, orc, Q, perm=np.random.default_rng(0).permutation(n))
    _check(pkg, orc, Q, perm=np.arange(n)[::-1].copy())
    _check(pkg, orc, Q, ordering="natural")


def test_edge_patterns(pkg, orc):
    _check(pkg, orc, sp.identity(1, format="csc"))
    _check(pkg, orc, sp.identity(50, format="csc"))  # fully disconnected
    # two disconnected cliques + isolated vertices
    blocks = [sp.csc_matrix(np.ones((7, 7))), sp.identity(3), sp.csc_matrix(np.ones((20, 20)))]
    _check(pkg, orc, sp.block_diag(blocks, format="csc"))
    # arrow matrix (dense last row/col)
    n = 30
    A = sp.lil_matrix((n, n))
    A.setdiag(1.0)
    A[n - 1, :] = 1.0
    A[:, n - 1] = 1.0
    _check(pkg, orc, A.tocsc())
    # 1-D chain and a random sparse symmetric pattern
    chain = sp.diags([np.ones(99), np.ones(100), np.ones(99)], [-1, 0, 1], format="csc")
    _check(pkg, orc, chain)
    R = sp.random(200, 200, density=0.02, random_state=1, format="csc")
    _check(pkg, orc, (R + R.T + sp.identity(200)).tocsc())


def test_nd_fill_quality(pkg, W):
    """Nested dissection must beat the natural (banded) ordering clearly on a 2-D mesh."""
    prob = W.matern_posterior(60, obs_frac=0.1, corr_range=0.2)
    nd = pkg.Symbolic(prob["Qpost"], host_only=True).info
    geo = pkg.Symbolic(prob["Qpost"], host_only=True, coords=prob["nodes"]).info
    nat = pkg.Symbolic(prob["Qpost"], host_only=True, ordering="natural").info
    assert nd.flops < 0.5 * nat.flops and geo.flops < 0.5 * nat.flops


def test_invalid_inputs(pkg):
    A = sp.identity(4, format="csc")
    with pytest.raises(pkg.GmrfbError):
        pkg.Symbolic(A, perm=np.array([0, 1, 1, 3]), host_only=True)  # not a permutation
    with pytest.raises(ValueError):
        pkg.Symbolic(sp.csc_matrix(np.ones((3, 4))), host_only=True)


# ------------------------------------------------------------------------ approximate minimum degree ----
def _brute_force_min_degree_fill(A):
    """Exact minimum-degree elimination on the dense pattern (ties: smallest index); returns nnz(L)."""
    n = A.shape[0]
    adj = [set(np.nonzero(A[:, j].toarray().ravel())[0].tolist()) - {j} for j in range(n)]
    alive = set(range(n))
    nnz = 0
    while alive:
        v = min(alive, key=lambda i: (len(adj[i]), i))
        nb = adj[v]
        nnz += len(nb) + 1
        for u in nb:
            adj[u] |= nb
            adj[u] -= {u, v}
        alive.remove(v)
    return nnz


@pytest.mark.parametrize("nx", [2, 5, 13, 40])
def test_amd_ordering_mesh(pkg, orc, W, nx):
    """ORDER_AMD (what `cholesky(A)` without `perm` asks CHOLMOD for): a valid permutation whose etree / column counts
    match the oracle's symbolic analysis of the same permutation."""
    prob = W.matern_posterior(nx, obs_frac=0.2, corr_range=0.3)
    _check(pkg, orc, prob["Qpost"], ordering="amd")


def test_amd_edge_patterns(pkg, orc):
    _check(pkg, orc, sp.identity(1, format="csc"), ordering="amd")
    _check(pkg, orc, sp.identity(50, format="csc"), ordering="amd")
    blocks = [sp.csc_matrix(np.ones((7, 7))), sp.identity(3), sp.csc_matrix(np.ones((20, 20)))]
    sym = _check(pkg, orc, sp.block_diag(blocks, format="csc"), ordering="amd")
    assert sym.info.nnz_L == 7 * 8 // 2 + 3 + 20 * 21 // 2  # cliques stay cliques: no fill
    n = 30
    A = sp.lil_matrix((n, n))
    A.setdiag(1.0)
    A[0, :] = 1.0
    A[:, 0] = 1.0  # arrow pointing the wrong way: the natural ordering fills completely, minimum degree not at all
    sym = _check(pkg, orc, A.tocsc(), ordering="amd")
    assert sym.info.nnz_L == 2 * n - 1
    chain = sp.diags([np.ones(99), np.ones(100), np.ones(99)], [-1, 0, 1], format="csc")
    sym = _check(pkg, orc, chain, ordering="amd")
    assert sym.info.nnz_L == 199  # a tree has a perfect elimination ordering and minimum degree finds it
    R = sp.random(200, 200, density=0.02, random_state=1, format="csc")
    _check(pkg, orc, (R + R.T + sp.identity(200)).tocsc(), ordering="amd")


def test_amd_fill_quality(pkg, W):
    """Fill of the approximate-degree ordering stays close to exact minimum degree on small problems and beats the
    natural ordering by a wide margin on a mesh."""
    rng = np.random.default_rng(3)
    for n, dens in ((40, 0.08), (80, 0.05), (120, 0.03)):
        R = sp.random(n, n, density=dens, random_state=int(rng.integers(1 << 30)), format="csc")
        A = (R + R.T + sp.identity(n)).tocsc()
        amd = pkg.Symbolic(A, host_only=True, ordering="amd").info.nnz_L
        exact = _brute_force_min_degree_fill(A)
        assert amd <= 1.25 * exact + 10, (n, amd, exact)
    prob = W.matern_posterior(60, obs_frac=0.1, corr_range=0.2)
    amd = pkg.Symbolic(prob["Qpost"], host_only=True, ordering="amd").info
    nat = pkg.Symbolic(prob["Qpost"], host_only=True, ordering="natural").info
    nd = pkg.Symbolic(prob["Qpost"], host_only=True).info
    assert amd.flops < 0.5 * nat.flops
    assert amd.nnz_L < 1.3 * nd.nnz_L


# ------------------------------------------------- nested dissection with halo-AMD leaves (ORDER_ND_AMD) ----
@pytest.mark.parametrize("nx", [5, 13, 40, 70])
def test_nd_amd_ordering_mesh(pkg, orc, W, nx):
    prob = W.matern_posterior(nx, obs_frac=0.2, corr_range=0.3)
    _check(pkg, orc, prob["Qpost"], ordering="nd_amd")
    _check(pkg, orc, prob["Qpost"], ordering="nd_amd", coords=prob["nodes"])
    _check(pkg, orc, prob["Qpost"], ordering="nd_amd", nd_leaf=30)  # several leaves with halos even on small meshes


def test_nd_amd_edge_patterns_and_fill(pkg, orc, W):
    _check(pkg, orc, sp.identity(1, format="csc"), ordering="nd_amd")
    _check(pkg, orc, sp.identity(50, format="csc"), ordering="nd_amd")
    blocks = [sp.csc_matrix(np.ones((7, 7))), sp.identity(3), sp.csc_matrix(np.ones((20, 20)))]
    _check(pkg, orc, sp.block_diag(blocks, format="csc"), ordering="nd_amd", nd_leaf=8)
    R = sp.random(300, 300, density=0.02, random_state=2, format="csc")
    _check(pkg, orc, (R + R.T + sp.identity(300)).tocsc(), ordering="nd_amd", nd_leaf=40)
    # ordering the leaves by minimum degree must not lose against leaves left in BFS order
    prob = W.matern_posterior(90, obs_frac=0.1, corr_range=0.2)
    nd = pkg.Symbolic(prob["Qpost"], host_only=True, nd_leaf=200).info
    nda = pkg.Symbolic(prob["Qpost"], host_only=True, ordering="nd_amd", nd_leaf=200).info
    assert nda.nnz_L <= nd.nnz_L and nda.flops <= nd.flops


# ------------------------------------------------------------- separator refinement of the nested dissection ----
def test_nd_regular_coordinates_with_ties(pkg, orc, W):
    """Structured meshes: whole grid lines share a coordinate; the cut is snapped to a gap between distinct values."""
    for nx in (5, 13, 40):
        prob = W.matern_posterior(nx, obs_frac=0.2, corr_range=0.3)
        g = np.linspace(0.0, 1.0, nx)
        X, Y = np.meshgrid(g, g)
        _check(pkg, orc, prob["Qpost"], coords=np.column_stack([X.ravel(), Y.ravel()]))
        _check(pkg, orc, prob["Qpost"], coords=np.column_stack([X.ravel(), np.zeros(nx * nx)]))  # degenerate axis


def test_vertex_cover_separators_reduce_fill():
    """GMRFB_ND_COVER is read once per process, so the two variants run in subprocesses: on a jittered mesh the
    minimum-vertex-cover separators must need clearly fewer flops than the plain boundary layers, for the geometric
    and no more for the graph dissection."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json; sys.path.insert(0, %r)\n"
        "import __graft_entry__ as g\n"
        "pkg = g.load_pkg(); prob = pkg.workloads.matern_posterior(90, obs_frac=0.1, corr_range=0.2)\n"
        "geo = pkg.Symbolic(prob['Qpost'], host_only=True, coords=prob['nodes']).info\n"
        "gra = pkg.Symbolic(prob['Qpost'], host_only=True).info\n"
        "print(json.dumps([geo.flops, geo.nnz_L, gra.flops, gra.nnz_L]))\n" % root
    )
    out = {}
    for flag in ("0", "1"):
        env = dict(os.environ, GMRFB_ND_COVER=flag)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        out[flag] = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["1"][0] < 0.9 * out["0"][0] and out["1"][1] < out["0"][1]  # geometric: at least 10 % fewer flops
    assert out["1"][2] <= 1.02 * out["0"][2]                               # graph: no worse


def test_amd_dense_rows_are_ordered_last(pkg, orc):
    """A vertex adjacent to (nearly) everything is removed before the minimum-degree elimination and ordered last
    (SuiteSparse AMD's dense-row rule): no quadratic rescans, and the arrow matrix still factors without fill."""
    import time

    n = 30000
    rows = np.concatenate([np.zeros(n - 1, int), np.arange(1, n)])
    cols = np.concatenate([np.arange(1, n), np.zeros(n - 1, int)])
    A = (sp.coo_matrix((np.ones(2 * (n - 1)), (rows, cols)), shape=(n, n)) + sp.identity(n) * n).tocsc()
    t0 = time.time()
    sym = pkg.Symbolic(A, host_only=True, ordering="amd")
    assert time.time() - t0 < 5.0
    assert sym.info.nnz_L == 2 * n - 1 and sym.p[-1] == 0
    # a hub on top of a mesh: same etree / counts as the oracle, hub last
    m = 40
    T = sp.diags([np.ones(m - 1), np.ones(m - 1)], [-1, 1])
    G = sp.kron(sp.identity(m), T) + sp.kron(T, sp.identity(m))
    N = m * m + 1
    H = sp.lil_matrix((N, N))
    H[:m * m, :m * m] = G
    H[m * m, :m * m] = 1.0
    H[:m * m, m * m] = 1.0
    H.setdiag(10.0)
    sym = _check(pkg, orc, H.tocsc(), ordering="amd")
    assert sym.p[-1] == m * m


def test_nested_dissection_is_independent_of_the_worker_threads(pkg, W, monkeypatch):
    """The dissection of disjoint vertex ranges runs on worker threads (GMRFB_ND_THREADS): every bisection reads and
    writes the state of its own range only, so the permutation is the same for any number of workers - geometric and
    graph bisection, with and without halo-AMD leaves (meshes above the 50 000-vertex threshold of the parallel path)."""
    prob = W.matern_posterior(230, obs_frac=0.1, q_eps=1e2, corr_range=0.05, seed=3)
    Q, nodes = prob["Qpost"], prob["nodes"]
    assert Q.shape[0] > 50000
    for ordering, coords in (("nd", nodes), ("nd", None), ("nd_amd", None)):
        ref = None
        for threads in (1, 3, 8):
            monkeypatch.setenv("GMRFB_ND_THREADS", str(threads))
            s = pkg.Symbolic(Q, coords=coords, ordering=ordering, host_only=True)
            p = np.asarray(s.p).copy()
            assert np.array_equal(np.sort(p), np.arange(Q.shape[0]))
            if ref is None:
                ref = (p, s.info.nnz_L)
            assert np.array_equal(p, ref[0]) and s.info.nnz_L == ref[1]
