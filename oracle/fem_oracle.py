"""TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu_baseline) - never imported by the product.

CPU restatement (NumPy, cell loops vectorised over the cells, quadrature/basis loops explicit) of the reference's
nonlinear FEM tangent assemblies - the producers of the Gauss-Newton Jacobians:

  * ``assemble_J_cube``                    _research/elliptic_chen24.jl:231-278   (2-D, cubic reaction term)
  * ``assemble_J_diff_and_f``              _research/elliptic_chen24.jl:180-228   (2-D, stiffness with skipped rows)
  * ``assemble_burgers_advection_matrix``  src/problems/burgers.jl:5-59           (1-D, u u_x and its tangent)
  * ``assemble_burgers_mass_diffusion_matrices`` src/problems/burgers.jl:61-98    (1-D, mass and stiffness)
  * the space-time tangent  J = J_static + dt J_adv,  f = J_static w + dt f_adv
                                           scripts/burgers/solve_burgers_gmrf-fem.jl:115-142

parity unpinned: the reference cannot run here (no Julia, no Ferrite) and holds no fixtures for these functions; the
loops below follow its source line by line with Lagrange elements and Gauss rules written out (Ferrite's reference
shapes: triangle P1 with the 3-point degree-2 rule for ``QuadratureRule{RefTriangle}(2)``; lines of order 1 / 2 with
Gauss-Legendre rules, quadratic lines numbered (left, right, middle) as Ferrite's ``QuadraticLine``).
"""
import numpy as np
import scipy.sparse as sp


# ------------------------------------------------------------------------------------------ quadrature --
def tri_quadrature(degree: int):
    """Barycentric points (nq x 3) and weights (sum 1; d Omega = weight * area) of symmetric triangle rules
    (Strang-Fix / Dunavant) exact to `degree` in {1, 2, 4}."""
    if degree == 1:
        return np.array([[1 / 3, 1 / 3, 1 / 3]]), np.array([1.0])
    if degree == 2:
        a, b = 1 / 6, 2 / 3
        return np.array([[b, a, a], [a, b, a], [a, a, b]]), np.full(3, 1 / 3)
    if degree == 4:
        a1, w1 = 0.445948490915965, 0.223381589678011
        a2, w2 = 0.091576213509771, 0.109951743655322
        pts = []
        for a in (a1, a2):
            b = 1 - 2 * a
            pts += [[b, a, a], [a, b, a], [a, a, b]]
        return np.array(pts), np.array([w1] * 3 + [w2] * 3)
    raise ValueError("degree must be 1, 2 or 4")


def line_quadrature(npts: int):
    """Gauss-Legendre points on [-1, 1] and weights (sum 2)."""
    return np.polynomial.legendre.leggauss(npts)


def line_shapes(order: int, xi):
    """Lagrange shape functions and d/dxi on [-1, 1]; order 2 numbered (left, right, middle)."""
    xi = np.asarray(xi, dtype=np.float64)
    if order == 1:
        return np.stack([(1 - xi) / 2, (1 + xi) / 2]), np.stack([np.full_like(xi, -0.5), np.full_like(xi, 0.5)])
    if order == 2:
        return (np.stack([xi * (xi - 1) / 2, xi * (xi + 1) / 2, 1 - xi * xi]),
                np.stack([xi - 0.5, xi + 0.5, -2 * xi]))
    raise ValueError("order must be 1 or 2")


def _coo(n_rows, n_cols, rows, cols, vals):
    A = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(n_rows, n_cols)).tocsc()
    A.sort_indices()
    return A


# ------------------------------------------------------------------------------------------ 2-D, P1 triangles --
def _tri_geometry(nodes, tris):
    p0, p1, p2 = nodes[tris[:, 0]], nodes[tris[:, 1]], nodes[tris[:, 2]]
    a2 = (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p1[:, 1] - p0[:, 1]) * (p2[:, 0] - p0[:, 0])
    ex = np.stack([p2[:, 0] - p1[:, 0], p0[:, 0] - p2[:, 0], p1[:, 0] - p0[:, 0]], axis=1)
    ey = np.stack([p2[:, 1] - p1[:, 1], p0[:, 1] - p2[:, 1], p1[:, 1] - p0[:, 1]], axis=1)
    gx, gy = -ey / a2[:, None], ex / a2[:, None]
    return 0.5 * np.abs(a2), gx, gy


def assemble_cubic_p1(nodes, tris, w, prescribed=None, degree=2):
    """(J_cube, f_cube) of _research/elliptic_chen24.jl:231-278 on P1 triangles:
    Je[i, j] += 3 phi_i u_q^2 phi_j dOmega, ve[i] += phi_i u_q^3 dOmega, rows of prescribed dofs skipped (:259-261)."""
    n = nodes.shape[0]
    area, _, _ = _tri_geometry(nodes, tris)
    lam, wq = tri_quadrature(degree)
    wc = w[tris]  # cells x 3
    Je = np.zeros((tris.shape[0], 3, 3))
    ve = np.zeros((tris.shape[0], 3))
    for q in range(len(wq)):
        dO = wq[q] * area
        cur_u = wc @ lam[q]
        for i in range(3):
            for j in range(3):
                Je[:, i, j] += 3 * lam[q, i] * cur_u**2 * lam[q, j] * dO
            ve[:, i] += lam[q, i] * cur_u**3 * dO
    if prescribed is not None:
        skip = np.asarray(prescribed, dtype=bool)[tris]
        Je[skip] = 0.0
        ve[skip] = 0.0
    rows = np.repeat(tris[:, :, None], 3, axis=2)
    cols = np.repeat(tris[:, None, :], 3, axis=1)
    f = np.zeros(n)
    np.add.at(f, tris.ravel(), ve.ravel())
    return _coo(n, n, rows, cols, Je), f


def assemble_stiffness_skipped_rows_p1(nodes, tris, prescribed=None):
    """J_diff of _research/elliptic_chen24.jl:180-228 (grad u . grad du, rows of prescribed dofs skipped :207-209)."""
    n = nodes.shape[0]
    area, gx, gy = _tri_geometry(nodes, tris)
    Je = (gx[:, :, None] * gx[:, None, :] + gy[:, :, None] * gy[:, None, :]) * area[:, None, None]
    if prescribed is not None:
        Je[np.asarray(prescribed, dtype=bool)[tris]] = 0.0
    rows = np.repeat(tris[:, :, None], 3, axis=2)
    cols = np.repeat(tris[:, None, :], 3, axis=1)
    return _coo(n, n, rows, cols, Je)


# ------------------------------------------------------------------------------------------ 1-D lines --
def _line_cellvalues(x, elems, order, nquad):
    xi, wq = line_quadrature(nquad)
    N, dN = line_shapes(order, xi)              # (npe, nq)
    xc = x if x.ndim == 2 else x[elems]         # cells x npe (per-element coordinates: periodic ring)
    jac = xc @ dN                               # cells x nq: dx/dxi
    return N, dN, jac, wq


def assemble_burgers_advection(x, elems, w, order=1, nquad=None, prescribed=None):
    """(G, v) of src/problems/burgers.jl:5-59: Ge[i, j] += phi_i (phi_j u_x + u phi_j') dOmega,
    ve[i] += phi_i u u_x dOmega; rows and columns of prescribed dofs zeroed afterwards (:53-57)."""
    n = int(elems.max()) + 1
    npe = order + 1
    nquad = nquad or order + 1
    N, dN, jac, wq = _line_cellvalues(x, elems, order, nquad)
    wc = w[elems]
    Ge = np.zeros((elems.shape[0], npe, npe))
    ve = np.zeros((elems.shape[0], npe))
    for q in range(len(wq)):
        dO = wq[q] * jac[:, q]
        cur_u = wc @ N[:, q]
        grad = dN[:, q][None, :] / jac[:, q][:, None]  # cells x npe: d phi_k / dx
        cur_du = np.sum(grad * wc, axis=1)
        for i in range(npe):
            for j in range(npe):
                Ge[:, i, j] += N[i, q] * (N[j, q] * cur_du + cur_u * grad[:, j]) * dO
            ve[:, i] += N[i, q] * cur_u * cur_du * dO
    rows = np.repeat(elems[:, :, None], npe, axis=2)
    cols = np.repeat(elems[:, None, :], npe, axis=1)
    G = _coo(n, n, rows, cols, Ge)
    v = np.zeros(n)
    np.add.at(v, elems.ravel(), ve.ravel())
    if prescribed is not None:
        keep = sp.diags((~np.asarray(prescribed, dtype=bool)).astype(np.float64))
        G = (keep @ G @ keep).tocsc()
        v[np.asarray(prescribed, dtype=bool)] = 0.0
    return G, v


def assemble_mass_stiffness_1d(x, elems, order=1, nquad=None, lumping=False, prescribed=None):
    """(M, G) of src/problems/burgers.jl:61-98 (consistent mass, or row-sum lumped; stiffness), rows and columns of
    prescribed dofs zeroed (:88-93)."""
    n = int(elems.max()) + 1
    npe = order + 1
    nquad = nquad or order + 1
    N, dN, jac, wq = _line_cellvalues(x, elems, order, nquad)
    Me = np.zeros((elems.shape[0], npe, npe))
    Ge = np.zeros((elems.shape[0], npe, npe))
    for q in range(len(wq)):
        dO = wq[q] * jac[:, q]
        grad = dN[:, q][None, :] / jac[:, q][:, None]
        for i in range(npe):
            for j in range(npe):
                Me[:, i, j] += N[i, q] * N[j, q] * dO
                Ge[:, i, j] += grad[:, i] * grad[:, j] * dO
    rows = np.repeat(elems[:, :, None], npe, axis=2)
    cols = np.repeat(elems[:, None, :], npe, axis=1)
    M, G = _coo(n, n, rows, cols, Me), _coo(n, n, rows, cols, Ge)
    if prescribed is not None:
        keep = sp.diags((~np.asarray(prescribed, dtype=bool)).astype(np.float64))
        M, G = (keep @ M @ keep).tocsc(), (keep @ G @ keep).tocsc()
    if lumping:
        M = sp.diags(np.asarray(M.sum(axis=1)).ravel()).tocsc()
    return M, G


def burgers_spacetime_tangent(x, elems, w, nt, dt, nu, order=1, nquad=None, prescribed=None):
    """(f, J) of scripts/burgers/solve_burgers_gmrf-fem.jl:115-142: J_static = M_{t+1} - M_t + dt nu G_{t+1} (rows =
    steps 2..nt), J = J_static + dt blockdiag(J_adv(u_t), t = 2..nt) placed in the columns of step t,
    f = J_static w + dt f_adv.  w is time-major (step t = entries [t n, (t+1) n))."""
    n = int(elems.max()) + 1
    M, G = assemble_mass_stiffness_1d(x, elems, order, nquad, False, prescribed)
    Enext = sp.csc_matrix((np.ones(nt - 1), (np.arange(nt - 1), np.arange(1, nt))), shape=(nt - 1, nt))
    Eprev = sp.csc_matrix((np.ones(nt - 1), (np.arange(nt - 1), np.arange(nt - 1))), shape=(nt - 1, nt))
    J_static = (sp.kron(Enext, M + dt * nu * G) - sp.kron(Eprev, M)).tocsc()
    blocks, vs = [], []
    for t in range(1, nt):
        Gt, vt = assemble_burgers_advection(x, elems, w[t * n:(t + 1) * n], order, nquad, prescribed)
        blocks.append(Gt)
        vs.append(vt)
    J_adv = sp.hstack([sp.csc_matrix(((nt - 1) * n, n)), sp.block_diag(blocks, format="csc")]).tocsc()
    f = J_static @ w + dt * np.concatenate(vs)
    J = (J_static + dt * J_adv).tocsc()
    J.sort_indices()
    return f, J
