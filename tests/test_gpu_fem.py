"""Device FEM assembly (gmrfb_fem_*) against the SciPy assembly of the workload generators (which restate
src/problems/darcy.jl:5-63 and src/spdes/shallow_water.jl:177-194 for P1 triangles)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


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
