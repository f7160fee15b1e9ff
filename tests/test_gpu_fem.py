"""Device FEM assembly (gmrfb_fem_*) against the SciPy assembly of the workload generators (which restate
src/problems/darcy.jl:5-63 and src/spdes/shallow_water.jl:177-194 for P1 triangles)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def relmat(A, B):
    return abs(A - B).max() / abs(B).max()


@pytest.mark.parametrize("nx", [5, 33, 120])
def test_mass_and_unit_stiffness(pkg, ctx, W, nx):
    nodes, tris = W.structured_mesh(nx, nx, seed=nx)
    m, G = W.p1_mass_stiffness(nodes, tris)
    fem = pkg.FEMP1(nodes, tris, ctx=ctx)
    np.testing.assert_allclose(fem.mass, m, rtol=1e-13)
    Gd = fem.assemble().to_scipy()
    assert np.array_equal(Gd.indptr, G.indptr) and np.array_equal(Gd.indices, G.indices)  # same pattern, bit exact
    assert relmat(Gd, G) < 1e-13


@pytest.mark.parametrize("nx,seed", [(41, 0), (97, 3)])
def test_darcy_stiffness_matches_scipy_assembly(pkg, ctx, W, nx, seed):
    """Config 3's observation operator: stiffness of the looked-up coefficient with identity rows on the boundary."""
    prob = W.darcy_problem(nx=nx, seed=seed)
    nodes, tris = W.structured_mesh(nx, nx, seed=0)
    g = prob["coeff_grid"].shape[0]
    fem = pkg.FEMP1(nodes, tris, ctx=ctx)
    fem.set_coeff_grid(np.linspace(0, 1, g), np.linspace(0, 1, g))
    x, y = nodes[:, 0], nodes[:, 1]
    bnd = (x == 0) | (x == 1) | (y == 0) | (y == 1)
    A = fem.assemble(prob["coeff_grid"], prescribed=bnd).to_scipy()
    Aref = prob["A"]
    # the SciPy reference drops nothing structurally, but its boundary rows hold explicit zeros too: compare densely
    assert abs(A - Aref).max() < 1e-12 * abs(Aref).max()
    np.testing.assert_allclose(fem.mass * (1.0 - bnd), prob["y"], rtol=1e-13, atol=0)
    # a second coefficient field on the same handle (the dataset loop), passed as a device tensor
    import torch

    prob2 = W.darcy_problem(nx=nx, seed=seed + 1)
    cg = torch.from_numpy(np.ascontiguousarray(prob2["coeff_grid"])).to(f"cuda:{ctx.device}")
    A2 = fem.assemble(cg, prescribed=bnd).to_scipy()
    assert abs(A2 - prob2["A"]).max() < 1e-12 * abs(prob2["A"]).max()


@pytest.mark.parametrize("nx", [9, 64])
def test_matern_precision_on_device(pkg, ctx, W, nx):
    nodes, tris = W.structured_mesh(nx, nx, seed=1)
    corr_range = 0.2
    Q = W.matern_precision(nodes, tris, corr_range)
    kappa = np.sqrt(8.0) / corr_range
    ratio = 1.0 / (4.0 * np.pi * kappa**2)
    fem = pkg.FEMP1(nodes, tris, ctx=ctx)
    Qd = fem.matern_precision(kappa, ratio).to_scipy()
    assert abs(Qd - Q).max() < 1e-12 * abs(Q).max()
    # prescribed dofs as in src/spdes/shallow_water.jl:178-181: Mt_ii = 1e-2, G_ii = 1
    presc = np.zeros(nodes.shape[0], bool)
    presc[[0, nx - 1, nx * nx - 1]] = True
    m, G = W.p1_mass_stiffness(nodes, tris)
    G = G.tolil()
    mt = m.copy()
    for d in np.flatnonzero(presc):
        G[d, d] = 1.0
        mt[d] = 1e-2
    K = (kappa**2 * sp.diags(mt) + G.tocsc()).tocsc()
    Qp = (ratio * (K.T @ sp.diags(1.0 / mt) @ K)).tocsc()
    Qpd = fem.matern_precision(kappa, ratio, prescribed=presc, prescribed_mass=1e-2).to_scipy()
    assert abs(Qpd - Qp).max() < 1e-12 * abs(Qp).max()


def test_dataset_loop_with_device_assembly(pkg, orc, ctx, W):
    """scripts/darcy/solve_darcy_gmrf-fem.jl:176-196 with the per-problem stiffness assembled on the device and handed to
    condition_on_observations as a device matrix: same posterior mean as the host-assembled path."""
    nx = 61
    nodes, tris = W.structured_mesh(nx, nx, seed=0)
    fem = pkg.FEMP1(nodes, tris, ctx=ctx)
    fem.set_coeff_grid(np.linspace(0, 1, 241), np.linspace(0, 1, 241))
    x_, y_ = nodes[:, 0], nodes[:, 1]
    bnd = (x_ == 0) | (x_ == 1) | (y_ == 0) | (y_ == 1)
    prob0 = W.darcy_problem(nx=nx, seed=0, q_eps=1e4)
    bp = pkg.CholeskySolverBlueprint(coords=nodes, ctx=ctx)
    x = pkg.GMRF(np.zeros(nx * nx), prob0["Q"], bp)
    for seed in (0, 1, 2):
        prob = W.darcy_problem(nx=nx, seed=seed, q_eps=1e4)
        Ad = fem.assemble(prob["coeff_grid"], prescribed=bnd)
        xc_dev = pkg.condition_on_observations(x, Ad, prob["q_eps"], prob["y"])
        m_dev = pkg.mean(xc_dev)
        Qpost = orc.posterior_precision(prob["Q"], prob["A"], prob["q_eps"])
        sym = xc_dev.solver_ref.value.precision_chol.sym
        ch = orc.SparseCholesky(Qpost, sym.p)
        m_ref = orc.posterior_mean(ch, prob["Q"], prob["A"], prob["q_eps"], prob["y"], np.zeros(nx * nx))
        assert np.linalg.norm(m_dev - m_ref) < 1e-9 * np.linalg.norm(m_ref)


# ------------------------------------------------------------------ nonlinear tangents --
@pytest.mark.parametrize("nx,degree,scale,with_bc", [(7, 2, 1.0, True), (40, 1, 0.0, False), (40, 4, 1.0, True),
                                                     (97, 2, 2.5, True)])
def test_cubic_tangent_matches_element_loop(pkg, orc, ctx, W, nx, degree, scale, with_bc):
    """gmrfb_fem_assemble_cubic against the restated element loops of _research/elliptic_chen24.jl:180-285
    (assemble_J_diff_and_f + assemble_J_cube + f_and_J): J = s J_diff + J_cube, f = s J_diff u + f_cube, rows of
    prescribed dofs skipped."""
    nodes, tris = W.structured_mesh(nx, nx, seed=nx)
    n = nodes.shape[0]
    rng = np.random.default_rng(nx)
    u = rng.standard_normal(n)
    x, y = nodes[:, 0], nodes[:, 1]
    bnd = ((x == 0) | (x == 1) | (y == 0) | (y == 1)) if with_bc else None
    Jc, fc = orc.fem.assemble_cubic_p1(nodes, tris, u, bnd, degree)
    Jd = orc.fem.assemble_stiffness_skipped_rows_p1(nodes, tris, bnd)
    Jref = (scale * Jd + Jc).tocsc()
    fref = scale * (Jd @ u) + fc
    fem = pkg.FEMP1(nodes, tris, ctx=ctx)
    f, J = fem.assemble_cubic(u, prescribed=bnd, quad_degree=degree, stiffness_scale=scale)
    Jg = J.to_scipy()
    assert abs(Jg - Jref).max() < 1e-13 * abs(Jref).max()
    np.testing.assert_allclose(f, fref, rtol=0, atol=1e-13 * np.abs(fref).max())
    if with_bc:
        assert abs(Jg[np.flatnonzero(bnd)]).max() == 0.0 and np.all(f[bnd] == 0.0)
    # the device matrix feeds SpMV and the posterior-precision plan like any other (row-wise copy kept in sync)
    np.testing.assert_allclose(J.matvec(u), Jref @ u, rtol=0, atol=1e-12 * np.abs(Jref @ u).max())
    np.testing.assert_allclose(J.matvec(u, trans=True), Jref.T @ u, rtol=0, atol=1e-12 * np.abs(Jref.T @ u).max())
    # device tensors in and out
    import torch
    ud = torch.as_tensor(u, device="cuda")
    fd = torch.empty(n, dtype=torch.float64, device="cuda")
    fem.assemble_cubic(ud, prescribed=bnd, quad_degree=degree, stiffness_scale=scale, out=fd)
    assert np.array_equal(fd.cpu().numpy(), f)  # same kernels, same summation order: bit-identical


def _bounded_line_mesh(ne, order, rng):
    """Non-uniform bounded mesh of [0, 1]; quadratic mid nodes off-centre (isoparametric Jacobian varies)."""
    xv = np.concatenate([[0.0], np.sort(rng.random(ne - 1)), [1.0]])
    left, right = np.arange(ne), np.arange(1, ne + 1)
    if order == 1:
        return xv, np.stack([left, right], axis=1)
    xm = xv[:-1] + (0.5 + 0.1 * (rng.random(ne) - 0.5)) * np.diff(xv)
    return np.concatenate([xv, xm]), np.stack([left, right, ne + 1 + left], axis=1)


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("periodic", [True, False])
def test_fem1d_matrices_match_element_loops(pkg, orc, ctx, W, order, periodic):
    """gmrfb_fem1d_mass_stiffness / gmrfb_fem1d_advection against the restated loops of src/problems/burgers.jl:5-98."""
    rng = np.random.default_rng(10 * order + periodic)
    ne = 37
    if periodic:
        x, el = W.periodic_line_mesh(ne, order)
        presc = None
    else:
        x, el = _bounded_line_mesh(ne, order, rng)
        presc = np.zeros(int(el.max()) + 1, bool)
        presc[[0, ne]] = True  # Dirichlet ends
    n = int(el.max()) + 1
    fem = pkg.FEM1D(x, el, order=order, ctx=ctx)
    assert fem.n == n
    for lump in (False, True):
        Mr, Gr = orc.fem.assemble_mass_stiffness_1d(x, el, order, None, lump, presc)
        M, G = fem.mass_stiffness(lumping=lump, prescribed=presc)
        assert abs(M.to_scipy() - Mr).max() < 1e-14 * abs(Mr).max()
        assert abs(G.to_scipy() - Gr).max() < 1e-13 * abs(Gr).max()
    u = rng.standard_normal(n)
    Ar, vr = orc.fem.assemble_burgers_advection(x, el, u, order, None, presc)
    A, v = fem.advection(u, prescribed=presc)
    assert abs(A.to_scipy() - Ar).max() < 1e-13 * abs(Ar).max()
    np.testing.assert_allclose(v, vr, rtol=0, atol=1e-13 * np.abs(vr).max())
    if not periodic:
        assert np.all(v[presc] == 0.0)
    # more quadrature points than the default: still the same as the loop with that rule
    fem4 = pkg.FEM1D(x, el, order=order, nquad=4, ctx=ctx)
    Ar4, vr4 = orc.fem.assemble_burgers_advection(x, el, u, order, 4, presc)
    A4, v4 = fem4.advection(u, prescribed=presc)
    assert abs(A4.to_scipy() - Ar4).max() < 1e-13 * abs(Ar4).max()
    np.testing.assert_allclose(v4, vr4, rtol=0, atol=1e-13 * np.abs(vr4).max())


@pytest.mark.parametrize("order,ne,nt", [(1, 16, 2), (2, 25, 7), (2, 200, 21)])
def test_fem1d_spacetime_tangent(pkg, orc, ctx, W, order, ne, nt):
    """f_and_J of scripts/burgers/solve_burgers_gmrf-fem.jl:115-142 in one kernel against the restated per-step loop."""
    x, el = W.periodic_line_mesh(ne, order)
    n = int(el.max()) + 1
    rng = np.random.default_rng(ne)
    w = rng.standard_normal(nt * n)
    dt, nu = 0.01, 0.02
    fr, Jr = orc.fem.burgers_spacetime_tangent(x, el, w, nt, dt, nu, order)
    fem = pkg.FEM1D.periodic_unit_interval(ne, order=order, ctx=ctx)
    f, J = fem.spacetime_tangent(w, nt, dt, nu)
    Jg = J.to_scipy()
    assert Jg.shape == Jr.shape == ((nt - 1) * n, nt * n)
    assert abs(Jg - Jr).max() < 1e-13 * abs(Jr).max()
    np.testing.assert_allclose(f, fr, rtol=0, atol=1e-13 * np.abs(fr).max())
    # a second iterate on the same handle (the Gauss-Newton loop), passed and returned as device tensors
    import torch
    w2 = rng.standard_normal(nt * n)
    fr2, Jr2 = orc.fem.burgers_spacetime_tangent(x, el, w2, nt, dt, nu, order)
    fd = torch.empty((nt - 1) * n, dtype=torch.float64, device="cuda")
    _, J2 = fem.spacetime_tangent(torch.as_tensor(w2, device="cuda"), nt, dt, nu, out=fd)
    assert J2.h.value == J.h.value  # fixed pattern, same device matrix
    assert abs(J2.to_scipy() - Jr2).max() < 1e-13 * abs(Jr2).max()
    np.testing.assert_allclose(fd.cpu().numpy(), fr2, rtol=0, atol=1e-13 * np.abs(fr2).max())
    np.testing.assert_allclose(J2.matvec(w2), Jr2 @ w2, rtol=0, atol=1e-12 * np.abs(Jr2 @ w2).max())


def test_gauss_newton_with_device_tangents(pkg, orc, ctx, W):
    """The Gauss-Newton drivers of config 1 (_research/elliptic_chen24.jl:148-161) and of the Burgers FEM script
    (scripts/burgers/solve_burgers_gmrf-fem.jl:172-182) with ``f_and_J`` assembled on the device: same iterates as the
    same loop fed by the restated host element loops."""
    # elliptic: -lap u + u^3 = g, Dirichlet rows skipped, prior = Matern
    P = W.elliptic_problem(25)
    nodes, tris = W.structured_mesh(25, 25, seed=0)
    n = P["n"]
    x, y = nodes[:, 0], nodes[:, 1]
    bnd = (x == 0) | (x == 1) | (y == 0) | (y == 1)
    Jd = orc.fem.assemble_stiffness_skipped_rows_p1(nodes, tris, bnd)
    g = np.where(bnd, 0.0, P["y"])
    fem = pkg.FEMP1(nodes, tris, ctx=ctx)

    def f_and_J_host(u):
        Jc, fc = orc.fem.assemble_cubic_p1(nodes, tris, u, bnd, 2)
        return Jd @ u + fc - g, (Jd + Jc).tocsc()

    def f_and_J_dev(u):
        f, J = fem.assemble_cubic(np.ascontiguousarray(u), prescribed=bnd, quad_degree=2, stiffness_scale=1.0)
        return f - g, J

    xc = pkg.condition_on_observations(pkg.GMRF(np.zeros(n), P["Q"], pkg.CholeskySolverBlueprint(coords=nodes, ctx=ctx)),
                                       P["A_bnd"], 1e8, P["y_bnd"])
    p = xc.solver_ref[()].precision_chol.p
    out = []
    for fj in (f_and_J_host, f_and_J_dev):
        gno = pkg.GaussNewtonOptimizer(pkg.mean(xc), pkg.precision_map(xc), fj, 1e6, np.zeros(n), pkg.mean(xc),
                                       solver_bp=pkg.GNCholeskySolverBlueprint(p, ctx=ctx), max_steps=6)
        out.append((pkg.optimize(gno), gno.n_steps))
    assert out[0][1] == out[1][1] >= 2
    assert rel(out[1][0], out[0][0]) < 1e-9

    # Burgers: periodic quadratic lines, implicit-Euler residual over nt steps
    ne, nt, dt, nu = 24, 6, 0.02, 0.05
    xe, el = W.periodic_line_mesh(ne, 2)
    ns = int(el.max()) + 1
    f1 = pkg.FEM1D(xe, el, order=2, ctx=ctx)
    M, G = orc.fem.assemble_mass_stiffness_1d(xe, el, 2)
    nodes_x = np.zeros(ns)
    nodes_x[el.ravel()] = xe.ravel() % 1.0
    u0 = np.sin(2 * np.pi * nodes_x)
    import scipy.sparse as sps
    Qs = (M + 0.05 * G).tocsc()
    Q = (sps.kron(sps.identity(nt), Qs) + 1e4 * sps.kron(sps.csc_matrix(([1.0], ([0], [0])), shape=(nt, nt)),
                                                          sps.identity(ns))).tocsc()
    mu = np.tile(u0, nt)

    def fj_host(w):
        return orc.fem.burgers_spacetime_tangent(xe, el, w, nt, dt, nu, 2)

    def fj_dev(w):
        return f1.spacetime_tangent(np.ascontiguousarray(w), nt, dt, nu)

    res = []
    for fj in (fj_host, fj_dev):
        gno = pkg.GaussNewtonOptimizer(mu, Q, fj, 1e6, np.zeros((nt - 1) * ns), mu,
                                       solver_bp=pkg.GNCholeskySolverBlueprint(ctx=ctx), max_steps=8)
        res.append((pkg.optimize(gno), gno.n_steps))
    assert res[0][1] == res[1][1] >= 2
    assert rel(res[1][0], res[0][0]) < 1e-9
