"""Lagrange triangles of order 1 / 2 on the device (gmrfb_fem2d_*, gmrfb_spgemm_*) against the restated element loops of
oracle/fem_oracle.py: assemble_darcy_diff_matrix (src/problems/darcy.jl:5-63) with the coefficient looked up at every
quadrature point, the element-lumped mass and the Matern powers of src/spdes/shallow_water.jl:115,172-190, and
f_and_J of _research/elliptic_chen24.jl:180-285 - on the quadratic elements the reference's scripts actually use
(src/utils.jl:20-38, element_order = 2)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


# The restated element loops compute the Jacobian of an element of size h from absolute node coordinates, as Ferrite's
# reinit! does: a relative rounding error of about eps / h (2e-13 on the 300 x 300-cell mesh) that the device kernel
# avoids by taking coordinates relative to the element's first vertex.  Tolerances below are against that oracle.
TOL = 2e-11


def relmat(A, B):
    return abs(A - B).max() / abs(B).max()


def same_pattern(A, B):
    return np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)


def _mesh(W, nx, order, curve=0.0, seed=0):
    nodes, tris = W.structured_mesh(nx, nx, seed=seed + nx)
    return (nodes, tris) if order == 1 else W.quadratic_mesh(nodes, tris, curve=curve, seed=seed)


def _boundary(nodes):
    x, y = nodes[:, 0], nodes[:, 1]
    return (x == 0) | (x == 1) | (y == 0) | (y == 1)


@pytest.mark.parametrize("order,nx,degree,curve", [(1, 5, 0, 0.0), (1, 40, 2, 0.0), (2, 4, 0, 0.0), (2, 33, 0, 0.0),
                                                   (2, 33, 4, 0.06), (2, 90, 3, 0.0), (1, 33, 4, 0.0), (2, 9, 2, 0.05)])
def test_unit_stiffness_load_and_mass(pkg, orc, ctx, W, order, nx, degree, curve):
    nodes, elems = _mesh(W, nx, order, curve)
    fem = pkg.FEMLagrange(nodes, elems, quad_degree=degree, ctx=ctx)
    deg = degree or order + 1
    assert fem.info["nodes_per_element"] == (3 if order == 1 else 6)
    assert fem.info["nquad"] == {1: 1, 2: 3, 3: 4, 4: 6}[deg]
    Gref, fref = orc.fem.assemble_darcy_lagrange(nodes, elems, order, beta=2.5, degree=deg)
    G, f = fem.stiffness(beta=2.5)
    Gd = G.to_scipy()
    assert same_pattern(Gd, Gref)
    assert relmat(Gd, Gref) < TOL
    np.testing.assert_allclose(f, fref, rtol=0, atol=TOL * np.abs(fref).max())
    # mass: consistent and the three lumpings
    Mref = orc.fem.assemble_mass_lagrange(nodes, elems, order, 0, degree=deg)
    M, _ = fem.mass(0)
    assert relmat(M.to_scipy(), Mref) < TOL
    for kind in (1, 2):
        mref = orc.fem.assemble_mass_lagrange(nodes, elems, order, kind, degree=deg)
        Ml, ml = fem.mass(kind)
        np.testing.assert_allclose(ml, mref, rtol=0, atol=TOL * np.abs(mref).max())
        Mld = Ml.to_scipy()
        np.testing.assert_allclose(Mld.diagonal(), ml, rtol=0, atol=0)
        assert abs(Mld - sp.diags(ml)).max() == 0.0
    _, m3 = fem.mass(3)
    np.testing.assert_array_equal(m3, fem.mass(1 if order == 1 else 2)[1])
    # SpMV through the row-wise copy of the device matrix
    u = np.random.default_rng(nx).standard_normal(nodes.shape[0])
    np.testing.assert_allclose(G.matvec(u), Gref @ u, rtol=0, atol=TOL * np.abs(Gref @ u).max())


def test_order1_general_path_equals_p1_path(pkg, ctx, W):
    """Order 1 through the general kernels against the specialised P1 assembler (same mesh, unit coefficient)."""
    nodes, tris = W.structured_mesh(57, 57, seed=5)
    g1 = pkg.FEMP1(nodes, tris, ctx=ctx)
    g2 = pkg.FEMLagrange(nodes, tris, ctx=ctx)
    A, B = g1.assemble().to_scipy(), g2.stiffness()[0].to_scipy()
    assert same_pattern(A, B) and relmat(B, A) < TOL
    np.testing.assert_allclose(g2.mass(1)[1], g1.mass, rtol=TOL)
    u = np.sin(3 * nodes[:, 0]) + nodes[:, 1] ** 2
    f1, J1 = g1.assemble_cubic(u, quad_degree=2, stiffness_scale=0.7)
    f2, J2 = g2.assemble_cubic(u, stiffness_scale=0.7)  # order + 1 = 2
    assert relmat(J2.to_scipy(), J1.to_scipy()) < TOL
    np.testing.assert_allclose(f2, f1, rtol=0, atol=TOL * np.abs(f1).max())


@pytest.mark.parametrize("order,nx,seed", [(2, 21, 0), (2, 61, 3), (1, 61, 1)])
def test_darcy_stiffness_per_quadrature_point_lookup(pkg, orc, ctx, W, order, nx, seed):
    """Config 3's observation operator on quadratic triangles: the two-level coefficient of a 241 x 241 grid looked up
    at every quadrature point (elements straddling a jump get both levels), identity rows on the boundary."""
    nodes, elems = _mesh(W, nx, order)
    g = 241
    coeff_grid = W.darcy_problem(nx=9, seed=seed)["coeff_grid"]          # (gy, gx), levels 3 / 12
    xc = yc = np.linspace(0, 1, g)
    bnd = _boundary(nodes)
    # the reference indexes coeff_mat[x_idx, y_idx]: coeff_mat = coeff_grid.T
    Gref, fref = orc.fem.assemble_darcy_lagrange(nodes, elems, order, xc, yc, coeff_grid.T, beta=1.0, prescribed=bnd)
    fem = pkg.FEMLagrange(nodes, elems, ctx=ctx)
    fem.set_coeff_grid(xc, yc)
    G, f = fem.stiffness(coeff_grid, prescribed=bnd)
    Gd = G.to_scipy()
    assert abs(Gd - Gref).max() < TOL * abs(Gref).max()
    np.testing.assert_allclose(f, fref, rtol=0, atol=TOL * np.abs(fref).max())
    assert np.all(f[bnd] == 0.0)
    # the lookup really is per quadrature point: a per-element (centroid) coefficient gives a different matrix
    if order == 2:
        N, grad, dO, xq = orc.fem.tri_cellvalues(nodes, elems, order)
        ix, iy = orc.fem.get_xy_idcs(xq.reshape(-1, 2), xc, yc)
        cq = coeff_grid.T[ix, iy].reshape(xq.shape[:2])
        assert np.any(cq.min(axis=1) != cq.max(axis=1))
    # next coefficient field of the dataset loop on the same handle, as a device tensor
    import torch

    cg2 = W.darcy_problem(nx=9, seed=seed + 7)["coeff_grid"]
    G2ref, _ = orc.fem.assemble_darcy_lagrange(nodes, elems, order, xc, yc, cg2.T, prescribed=bnd)
    G2, _ = fem.stiffness(torch.from_numpy(np.ascontiguousarray(cg2)).to(f"cuda:{ctx.device}"), prescribed=bnd, load=False)
    assert abs(G2.to_scipy() - G2ref).max() < TOL * abs(G2ref).max()


def test_nearest_index_ties_and_unsorted_axes(pkg, orc, ctx, W):
    """`argmin(abs.(coords .- x))` takes the first minimum: quadrature points exactly between two grid lines, and a
    descending axis (the general search)."""
    nodes, tris = W.structured_mesh(9, 9, jitter=0.0)
    nodes6, elems6 = W.quadratic_mesh(nodes, tris)
    rng = np.random.default_rng(2)
    for xc, yc in ((np.linspace(0, 1, 13), np.linspace(0, 1, 25)), (np.linspace(1, 0, 17), np.linspace(0, 1, 6))):
        cm = rng.uniform(1, 5, size=(xc.size, yc.size))                 # coeff_mat[x_idx, y_idx]
        Gref, _ = orc.fem.assemble_darcy_lagrange(nodes6, elems6, 2, xc, yc, cm, degree=2)
        fem = pkg.FEMLagrange(nodes6, elems6, quad_degree=2, ctx=ctx)
        fem.set_coeff_grid(xc, yc)
        G, _ = fem.stiffness(np.ascontiguousarray(cm.T), load=False)
        assert abs(G.to_scipy() - Gref).max() < TOL * abs(Gref).max()


@pytest.mark.parametrize("order,nx,scale,with_bc,curve", [(2, 6, 1.0, True, 0.0), (2, 40, 0.0, False, 0.0),
                                                          (2, 40, 2.5, True, 0.05), (1, 40, 1.0, True, 0.0)])
def test_cubic_tangent_on_lagrange_triangles(pkg, orc, ctx, W, order, nx, scale, with_bc, curve):
    nodes, elems = _mesh(W, nx, order, curve)
    n = nodes.shape[0]
    u = np.random.default_rng(nx).standard_normal(n)
    bnd = _boundary(nodes) if with_bc else None
    Jref, fref = orc.fem.assemble_cubic_lagrange(nodes, elems, order, u, bnd, stiffness_scale=scale)
    fem = pkg.FEMLagrange(nodes, elems, ctx=ctx)
    f, J = fem.assemble_cubic(u, prescribed=bnd, stiffness_scale=scale)
    Jg = J.to_scipy()
    assert abs(Jg - Jref).max() < TOL * abs(Jref).max()
    np.testing.assert_allclose(f, fref, rtol=0, atol=TOL * np.abs(fref).max())
    if with_bc:
        assert abs(Jg[np.flatnonzero(bnd)]).max() == 0.0 and np.all(f[bnd] == 0.0)
    import torch

    ud = torch.as_tensor(u, device="cuda")
    fd = torch.empty(n, dtype=torch.float64, device="cuda")
    fem.assemble_cubic(ud, prescribed=bnd, stiffness_scale=scale, out=fd)
    assert np.array_equal(fd.cpu().numpy(), f)


@pytest.mark.parametrize("order,nx,alpha,with_bc", [(2, 7, 2, False), (2, 25, 3, False), (1, 30, 3, False), (2, 25, 2, True),
                                                    (2, 12, 3, True)])
def test_matern_powers_on_device(pkg, orc, ctx, W, order, nx, alpha, with_bc):
    nodes, elems = _mesh(W, nx, order)
    kappa, ratio = np.sqrt(8.0) / 0.2, 0.37
    bnd = _boundary(nodes) if with_bc else None
    Qref = orc.fem.matern_precision_lagrange(nodes, elems, order, kappa, ratio, alpha=alpha, prescribed=bnd)
    fem = pkg.FEMLagrange(nodes, elems, ctx=ctx)
    Q = fem.matern_precision(kappa, ratio, alpha=alpha, prescribed=bnd).to_scipy()
    assert abs(Q - Qref).max() < TOL * abs(Qref).max()
    assert abs(Q - Q.T).max() < 1e-12 * abs(Q).max()
    # the prior factorises and its marginal variances are positive (what `discretize` hands to the solver); the odd
    # power with overwritten diagonal entries of prescribed dofs is indefinite by construction (K is), so not that one
    if not (alpha == 3 and with_bc):
        F = pkg.cholesky(sp.csc_matrix((Q + Q.T) * 0.5), ctx=ctx)
        assert F.issuccess() and np.all(F.var_selinv() > 0)
    # second call on the same handle re-uses both plans
    Q2 = fem.matern_precision(kappa * 1.3, ratio, alpha=alpha, prescribed=bnd).to_scipy()
    Q2ref = orc.fem.matern_precision_lagrange(nodes, elems, order, kappa * 1.3, ratio, alpha=alpha, prescribed=bnd)
    assert abs(Q2 - Q2ref).max() < TOL * abs(Q2ref).max()


def test_sparse_product_plan(pkg, ctx):
    rng = np.random.default_rng(0)
    A = sp.random(70, 50, density=0.08, random_state=1, format="csc")
    Bm = sp.random(50, 90, density=0.1, random_state=2, format="csc")
    A.sort_indices(), Bm.sort_indices()
    w = rng.uniform(0.5, 2.0, 50)
    Ad, Bd = pkg.SparseMatrix(A, ctx=ctx), pkg.SparseMatrix(Bm, ctx=ctx)
    plan = pkg.SparseProduct(Ad, Bd)
    Cref = (A @ sp.diags(w) @ Bm).tocsc()
    C = plan.compute(alpha=-1.5, w=w).to_scipy()
    assert abs(C + 1.5 * Cref).max() < 1e-14 * abs(Cref).max()
    # structural pattern (no numerical cancellation dropped), columns sorted
    S = ((abs(A) > 0).astype(np.float64) @ (abs(Bm) > 0).astype(np.float64)).tocsc()
    S.sort_indices()
    assert same_pattern(C, S)
    C1 = plan.compute().to_scipy()
    assert abs(C1 - (A @ Bm)).max() < 1e-14 * abs(Cref).max()


def test_sparse_product_with_an_empty_factor(pkg, ctx):
    A = sp.random(70, 50, density=0.08, random_state=1, format="csc")
    A.sort_indices()
    E = sp.csc_matrix((50, 90))
    C0 = pkg.SparseProduct(pkg.SparseMatrix(A, ctx=ctx), pkg.SparseMatrix(E, ctx=ctx)).compute()
    assert C0.dims() == (70, 90, 0)


def test_elliptic_gauss_newton_on_quadratic_triangles(pkg, orc, ctx, W):
    """Config 1 on the element the script uses (P2, QuadratureRule(3)): the Gauss-Newton loop with the device tangent
    reproduces the loop driven by the restated element loops."""
    nodes, elems = W.quadratic_mesh(*W.structured_mesh(13, 13, seed=2))
    n = nodes.shape[0]
    bnd = _boundary(nodes)
    x, y = nodes[:, 0], nodes[:, 1]
    truth = np.sin(np.pi * x) * np.sin(np.pi * y)
    fem = pkg.FEMLagrange(nodes, elems, ctx=ctx)
    kappa = np.sqrt(8.0) / 0.3
    Q = fem.matern_precision(kappa, 1.0 / (4 * np.pi * kappa**2), alpha=2).to_scipy()
    Q = sp.csc_matrix((Q + Q.T) * 0.5) + 1e4 * sp.diags(bnd.astype(np.float64))   # boundary pinned to 0 by the prior
    # manufactured load: -lap u + u^3 with the oracle's consistent mass
    M = orc.fem.assemble_mass_lagrange(nodes, elems, 2, 0)
    G0, _ = orc.fem.assemble_darcy_lagrange(nodes, elems, 2)
    load = M @ (2 * np.pi**2 * truth + truth**3)
    load[bnd] = 0.0

    def f_and_J_dev(u):
        f, J = fem.assemble_cubic(u, prescribed=bnd)
        return f - load, J

    def f_and_J_ref(u):
        J, f = orc.fem.assemble_cubic_lagrange(nodes, elems, 2, u, bnd, stiffness_scale=1.0)
        return f - load, J

    outs = []
    for fj in (f_and_J_ref, f_and_J_dev):
        gno = pkg.GaussNewtonOptimizer(np.zeros(n), Q, fj, 1e6, np.zeros(n), np.zeros(n), max_steps=6,
                                       solver_bp=pkg.GNCholeskySolverBlueprint(ctx=ctx))
        outs.append((pkg.optimize(gno), gno.n_steps))
    assert outs[0][1] == outs[1][1] >= 2
    assert np.linalg.norm(outs[1][0] - outs[0][0]) < 1e-8 * np.linalg.norm(outs[0][0])
    # and the loop solves the PDE: P2 on a 12 x 12 mesh resolves sin(pi x) sin(pi y) to a few percent at worst
    assert np.linalg.norm(outs[0][0] - truth) < 0.1 * np.linalg.norm(truth)


def test_darcy_dataset_loop_on_quadratic_triangles(pkg, orc, ctx, W):
    """scripts/darcy/solve_darcy_gmrf-fem.jl:93-100,165-196 on the script's own discretisation (element_order = 2): Matern
    prior of smoothness 2 built on the device, then per problem of the dataset loop the stiffness of a new coefficient
    field (device), condition_on_observations with the device matrix, posterior mean, a sample and RBMC variances -
    against the oracle's sparse Cholesky on the matrices of the restated element loops."""
    nodes, elems = W.quadratic_mesh(*W.structured_mesh(21, 21, seed=0))
    n = nodes.shape[0]
    bnd = _boundary(nodes)
    xc = yc = np.linspace(0, 1, 241)
    fem = pkg.FEMLagrange(nodes, elems, ctx=ctx)
    fem.set_coeff_grid(xc, yc)
    kappa = np.sqrt(8.0 * 2) / 0.25
    ratio = 1.0 / (8.0 * np.pi * kappa**4)
    Qd = fem.matern_precision(kappa, ratio, alpha=3)
    Qref = orc.fem.matern_precision_lagrange(nodes, elems, 2, kappa, ratio, alpha=3)
    Qref = sp.csc_matrix((Qref + Qref.T) * 0.5)
    Qdev = Qd.to_scipy()
    bp = pkg.CholeskySolverBlueprint(var_strategy=pkg.RBMCStrategy(40, rng=np.random.default_rng(1)), coords=nodes, ctx=ctx)
    x = pkg.GMRF(np.zeros(n), sp.csc_matrix((Qdev + Qdev.T) * 0.5), bp)
    q_eps = 1e4
    for seed in (0, 1):
        cg = W.darcy_problem(nx=9, seed=seed)["coeff_grid"]
        Ad, y = fem.stiffness(cg, prescribed=bnd)
        xc_dev = pkg.condition_on_observations(x, Ad, q_eps, y)
        m_dev = pkg.mean(xc_dev)
        Aref, yref = orc.fem.assemble_darcy_lagrange(nodes, elems, 2, xc, yc, cg.T, prescribed=bnd)
        Qpost = orc.posterior_precision(Qref, Aref, q_eps)
        sym = xc_dev.solver_ref.value.precision_chol.sym
        ch = orc.SparseCholesky(Qpost, sym.p)
        m_ref = orc.posterior_mean(ch, Qref, Aref, q_eps, yref, np.zeros(n))
        assert np.linalg.norm(m_dev - m_ref) < 1e-7 * np.linalg.norm(m_ref)
        v_dev = pkg.var(xc_dev)
        assert np.all(v_dev > 0)
        # RBMC(40) against the exact marginal variances of the oracle: a Monte-Carlo estimate, so a loose band
        v_ref = ch.selinv_diag()
        assert np.median(np.abs(v_dev - v_ref) / v_ref) < 0.35
        s = pkg.rand(np.random.default_rng(seed), xc_dev)
        assert s.shape == (n,) and np.all(np.isfinite(s))


def test_invalid_arguments_are_reported(pkg, ctx, W):
    """Error behaviour at the boundary: status codes with a message, never a crash."""
    nodes, tris = W.structured_mesh(5, 5, seed=0)
    n6, e6 = W.quadratic_mesh(nodes, tris)
    with pytest.raises(pkg.GmrfbError):
        pkg.FEMLagrange(n6, e6 * 0 + n6.shape[0], ctx=ctx)  # node index out of range
    with pytest.raises(pkg.GmrfbError):
        pkg.FEMLagrange(n6, e6, quad_degree=7, ctx=ctx)
    flat = n6.copy()
    flat[:, 1] = 0.0                                   # every element degenerate: det J = 0
    with pytest.raises(pkg.GmrfbError):
        pkg.FEMLagrange(flat, e6, ctx=ctx)
    lonely = np.concatenate([n6, [[2.0, 2.0]]], axis=0)  # a node that belongs to no element
    with pytest.raises(pkg.GmrfbError):
        pkg.FEMLagrange(lonely, e6, ctx=ctx)
    fem = pkg.FEMLagrange(n6, e6, ctx=ctx)
    fem._grid = (3, 3)
    with pytest.raises(pkg.GmrfbError):                # coefficient grid before gmrfb_fem2d_set_coeff_grid
        fem.stiffness(np.ones((3, 3)))
    with pytest.raises(pkg.GmrfbError):
        fem.matern_precision(-1.0, 1.0)
    with pytest.raises(pkg.GmrfbError):
        fem.matern_precision(1.0, 1.0, alpha=4)
    G, f = fem.stiffness()                             # the handle is still usable afterwards
    assert abs(f.sum() - 1.0) < 1e-12
