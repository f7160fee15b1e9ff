"""TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu_baseline) - never imported by the product.

CPU restatement (NumPy, cell loops vectorised over the cells, quadrature/basis loops explicit) of the reference's
nonlinear FEM tangent assemblies - the producers of the Gauss-Newton Jacobians:

  * ``assemble_J_cube``                    _research/elliptic_chen24.jl:231-278   (2-D, cubic reaction term)
  * ``assemble_J_diff_and_f``              _research/elliptic_chen24.jl:180-228   (2-D, stiffness with skipped rows)
  * ``assemble_burgers_advection_matrix``  src/problems/burgers.jl:5-59           (1-D, u u_x and its tangent)
  * ``assemble_burgers_mass_diffusion_matrices`` src/problems/burgers.jl:61-98    (1-D, mass and stiffness)
  * the space-time tangent  J = J_static + dt J_adv,  f = J_static w + dt f_adv
                                           scripts/burgers/solve_burgers_gmrf-fem.jl:115-142
  * ``assemble_darcy_diff_matrix``         src/problems/darcy.jl:5-63             (2-D, coefficient looked up at every
                                           quadrature point by nearest grid index, src/datasets/darcy.jl:30-34)
  * Lagrange triangles of order 1 AND 2 with the cell values Ferrite computes for them (isoparametric geometry,
    ``QuadratureRule{RefTriangle}(order + 1)``, src/utils.jl:29-31, _research/elliptic_chen24.jl:118-122): the
    ``*_lagrange`` functions below; the lumped element mass of src/spdes/shallow_water.jl:115 (``lump_matrix``)

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
    (Strang-Fix / Dunavant) exact to `degree` in {1, 2, 3, 4}; degree 3 is Dunavant's 4-point rule with the negative
    centroid weight (what ``QuadratureRule{RefTriangle}(3)`` selects for quadratic triangles [RECALL])."""
    if degree == 1:
        return np.array([[1 / 3, 1 / 3, 1 / 3]]), np.array([1.0])
    if degree == 2:
        a, b = 1 / 6, 2 / 3
        return np.array([[b, a, a], [a, b, a], [a, a, b]]), np.full(3, 1 / 3)
    if degree == 3:
        return (np.array([[1 / 3, 1 / 3, 1 / 3], [0.6, 0.2, 0.2], [0.2, 0.6, 0.2], [0.2, 0.2, 0.6]]),
                np.array([-27 / 48, 25 / 48, 25 / 48, 25 / 48]))
    if degree == 4:
        a1, w1 = 0.445948490915965, 0.223381589678011
        a2, w2 = 0.091576213509771, 0.109951743655322
        pts = []
        for a in (a1, a2):
            b = 1 - 2 * a
            pts += [[b, a, a], [a, b, a], [a, a, b]]
        return np.array(pts), np.array([w1] * 3 + [w2] * 3)
    raise ValueError("degree must be 1, 2, 3 or 4")


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


# ------------------------------------------------------------------------------------------ 2-D, Lagrange triangles of order 1 / 2 --
def tri_shapes(order: int, lam):
    """Lagrange shape functions on the triangle at the barycentric points `lam` (nq x 3): (N, dN) with N[q, a] and
    dN[q, a, b] = dN_a / dl_b (the three barycentric coordinates taken as independent variables).  Order 2 is numbered
    as Ferrite's QuadraticTriangle: vertices 0, 1, 2, then the midpoints of the edges (0,1), (1,2), (2,0)."""
    lam = np.asarray(lam, dtype=np.float64)
    nq = lam.shape[0]
    if order == 1:
        return lam.copy(), np.broadcast_to(np.eye(3), (nq, 3, 3)).copy()
    if order != 2:
        raise ValueError("order must be 1 or 2")
    N = np.zeros((nq, 6))
    dN = np.zeros((nq, 6, 3))
    for v in range(3):
        N[:, v] = lam[:, v] * (2 * lam[:, v] - 1)
        dN[:, v, v] = 4 * lam[:, v] - 1
    for e, (a, b) in enumerate(((0, 1), (1, 2), (2, 0))):
        N[:, 3 + e] = 4 * lam[:, a] * lam[:, b]
        dN[:, 3 + e, a] = 4 * lam[:, b]
        dN[:, 3 + e, b] = 4 * lam[:, a]
    return N, dN


def tri_cellvalues(nodes, elems, order, degree=None):
    """What ``reinit!(cellvalues, cell)`` provides for every cell: N[q, a], physical gradients grad[c, q, a, :],
    dOmega[c, q] (``getdetJdV``) and the quadrature points xq[c, q, :] (``spatial_coordinate``), with the geometry
    interpolated by the same Lagrange basis (isoparametric).  Reference coordinates xi = l1, eta = l2."""
    lam, wq = tri_quadrature(degree or order + 1)
    N, dNl = tri_shapes(order, lam)
    dref = np.stack([dNl[:, :, 1] - dNl[:, :, 0], dNl[:, :, 2] - dNl[:, :, 0]], axis=2)  # nq x npe x (xi, eta)
    X = nodes[elems]                                        # cells x npe x 2
    Jm = np.einsum("cad,qar->cqdr", X, dref)                # J[d, r] = d x_d / d ref_r
    det = Jm[..., 0, 0] * Jm[..., 1, 1] - Jm[..., 0, 1] * Jm[..., 1, 0]
    inv = np.empty_like(Jm)
    inv[..., 0, 0], inv[..., 0, 1] = Jm[..., 1, 1], -Jm[..., 0, 1]
    inv[..., 1, 0], inv[..., 1, 1] = -Jm[..., 1, 0], Jm[..., 0, 0]
    inv /= det[..., None, None]
    grad = np.einsum("qar,cqrd->cqad", dref, inv)           # [dN/dx, dN/dy] = [dN/dxi, dN/deta] J^-1
    dO = 0.5 * np.abs(det) * wq[None, :]
    xq = np.einsum("qa,cad->cqd", N, X)
    return N, grad, dO, xq


def _scatter(n, elems, Ae):
    npe = elems.shape[1]
    rows = np.repeat(elems[:, :, None], npe, axis=2)
    cols = np.repeat(elems[:, None, :], npe, axis=1)
    return _coo(n, n, rows, cols, Ae)


def get_xy_idcs(points, x_coords, y_coords):
    """src/datasets/darcy.jl:30-34: nearest grid index per axis, first minimum on ties (``argmin``)."""
    ix = np.argmin(np.abs(np.asarray(x_coords)[None, :] - points[:, 0:1]), axis=1)
    iy = np.argmin(np.abs(np.asarray(y_coords)[None, :] - points[:, 1:2]), axis=1)
    return ix, iy


def assemble_darcy_lagrange(nodes, elems, order, x_coords=None, y_coords=None, coeff_mat=None, beta=1.0,
                            prescribed=None, degree=None):
    """(G, f) of src/problems/darcy.jl:5-63: Ge[i, j] += coeff(x_q) grad phi_i . grad phi_j dOmega,
    fe[i] += beta phi_i dOmega, the coefficient ``coeff_mat[x_idx, y_idx]`` looked up at every quadrature point (:39).
    Dirichlet rows (``prescribed``) become identity rows with a zero load, as the host mirror does for P1."""
    n = nodes.shape[0]
    N, grad, dO, xq = tri_cellvalues(nodes, elems, order, degree)
    nc, nq, npe = grad.shape[:3]
    if coeff_mat is None:
        cq = np.ones((nc, nq))
    else:
        ix, iy = get_xy_idcs(xq.reshape(-1, 2), x_coords, y_coords)
        cq = np.asarray(coeff_mat)[ix, iy].reshape(nc, nq)
    Ge = np.zeros((nc, npe, npe))
    fe = np.zeros((nc, npe))
    for q in range(nq):
        for i in range(npe):
            fe[:, i] += beta * N[q, i] * dO[:, q]
            for j in range(npe):
                Ge[:, i, j] += cq[:, q] * np.sum(grad[:, q, i] * grad[:, q, j], axis=1) * dO[:, q]
    G = _scatter(n, elems, Ge)
    f = np.zeros(n)
    np.add.at(f, elems.ravel(), fe.ravel())
    if prescribed is not None:
        p = np.asarray(prescribed, dtype=bool)
        G = (sp.diags((~p).astype(np.float64)) @ G + sp.diags(p.astype(np.float64))).tocsc()
        G.sort_indices()
        f[p] = 0.0
    return G, f


def assemble_mass_lagrange(nodes, elems, order, lumping=0, degree=None):
    """Mass matrix: consistent (lumping 0), or lumped per element as src/spdes/shallow_water.jl:115 does with
    ``lump_matrix(me, ip)`` - row sums (lumping 1: first-order elements) or the diagonal scaled to the element's total
    mass (lumping 2: ``diag(me) * sum(me) / sum(diag(me))``, higher orders, where the row sums of the vertex functions
    vanish) [RECALL: GaussianMarkovRandomFields.jl, not in the reference tree].  Lumped: returns the diagonal."""
    n = nodes.shape[0]
    N, _, dO, _ = tri_cellvalues(nodes, elems, order, degree)
    nc, nq = dO.shape
    npe = N.shape[1]
    Me = np.zeros((nc, npe, npe))
    for q in range(nq):
        for i in range(npe):
            for j in range(npe):
                Me[:, i, j] += N[q, i] * N[q, j] * dO[:, q]
    if lumping == 0:
        return _scatter(n, elems, Me)
    if lumping == 1:
        ml = Me.sum(axis=2)
    else:
        d = np.einsum("cii->ci", Me)
        ml = d * (Me.sum(axis=(1, 2)) / d.sum(axis=1))[:, None]
    m = np.zeros(n)
    np.add.at(m, elems.ravel(), ml.ravel())
    return m


def assemble_cubic_lagrange(nodes, elems, order, w, prescribed=None, degree=None, stiffness_scale=0.0):
    """(J, f) with J = s J_diff + J_cube and f = s J_diff w + f_cube: ``assemble_J_cube``
    (_research/elliptic_chen24.jl:231-278) and ``assemble_J_diff_and_f`` without its load (:180-228), rows of
    prescribed dofs skipped (:207-209, :259-261)."""
    n = nodes.shape[0]
    N, grad, dO, _ = tri_cellvalues(nodes, elems, order, degree)
    nc, nq, npe = grad.shape[:3]
    wc = w[elems]
    Je = np.zeros((nc, npe, npe))
    Jd = np.zeros((nc, npe, npe))
    ve = np.zeros((nc, npe))
    for q in range(nq):
        cur_u = wc @ N[q]
        for i in range(npe):
            for j in range(npe):
                Je[:, i, j] += 3 * N[q, i] * cur_u**2 * N[q, j] * dO[:, q]
                Jd[:, i, j] += np.sum(grad[:, q, i] * grad[:, q, j], axis=1) * dO[:, q]
            ve[:, i] += N[q, i] * cur_u**3 * dO[:, q]
    if prescribed is not None:
        skip = np.asarray(prescribed, dtype=bool)[elems]
        Je[skip] = 0.0
        Jd[skip] = 0.0
        ve[skip] = 0.0
    J_diff = _scatter(n, elems, Jd)
    f = np.zeros(n)
    np.add.at(f, elems.ravel(), ve.ravel())
    # one matrix on the pattern `allocate_matrix(dh)` gives both parts (a sparse sum would drop entries that happen to
    # be exactly zero and change the pattern from one Gauss-Newton iterate to the next)
    J = _scatter(n, elems, stiffness_scale * Jd + Je)
    return J, stiffness_scale * (J_diff @ w) + f


def matern_precision_lagrange(nodes, elems, order, kappa, ratio, alpha=2, prescribed=None, prescribed_mass=1e-2,
                              degree=None):
    """src/spdes/shallow_water.jl:172-190: Mt = lumped mass with Mt[dof, dof] = 1e-2 and G[dof, dof] = 1 for
    prescribed dofs, K = kappa^2 Mt + G, Q = ratio K' Mt^-1 K (alpha 2); alpha 3 is the commented-out line :186,
    Q = ratio K Mt^-1 K Mt^-1 K (what a Matern field of smoothness 2 in two dimensions needs,
    scripts/darcy/solve_darcy_gmrf-fem.jl:97)."""
    G, _ = assemble_darcy_lagrange(nodes, elems, order, degree=degree)
    m = assemble_mass_lagrange(nodes, elems, order, lumping=1 if order == 1 else 2, degree=degree)
    G = G.tolil()
    if prescribed is not None:
        for dof in np.flatnonzero(np.asarray(prescribed, dtype=bool)):
            G[dof, dof] = 1.0
            m[dof] = prescribed_mass
    K = (kappa**2 * sp.diags(m) + G.tocsc()).tocsc()
    Mi = sp.diags(1.0 / m)
    Q = ratio * (K.T @ Mi @ K) if alpha == 2 else ratio * (K @ Mi @ K @ Mi @ K)
    Q = Q.tocsc()
    Q.sort_indices()
    return Q
