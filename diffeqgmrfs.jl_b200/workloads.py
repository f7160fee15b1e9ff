"""Synthetic workload generators for the configurations of BASELINE.json.

The reference assembles its matrices with Ferrite/Gmsh (src/utils.jl:20-49) from datasets that are not shipped;
these generators produce matrices of the same structure on structured (optionally jittered) triangulations with
deterministic seeds, so CPU oracle and GPU library see identical inputs.  Pure NumPy/SciPy, host side only.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def structured_mesh(nx: int, ny: int | None = None, jitter: float = 0.2, seed: int = 0):
    """(nx x ny) nodes on the unit square, each cell split into two P1 triangles along its (1,1) diagonal.
    Interior nodes are jittered by `jitter`*h (seeded) so that triangles are generic and the stiffness
    matrix has the full 7-point pattern of an unstructured P1 mesh (=> 19-point pattern for K M^-1 K)."""
    ny = ny or nx
    xs, ys = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny), indexing="xy")
    nodes = np.stack([xs.ravel(), ys.ravel()], axis=1)
    if jitter > 0 and nx > 2 and ny > 2:
        rng = np.random.default_rng(seed)
        h = np.array([1.0 / (nx - 1), 1.0 / (ny - 1)])
        jit = (rng.random(nodes.shape) - 0.5) * jitter * h
        interior = ((xs > 0) & (xs < 1) & (ys > 0) & (ys < 1)).ravel()
        nodes[interior] += jit[interior]
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    v00 = (j * nx + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + nx, v00 + nx + 1
    tris = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=0)
    return nodes, tris


def quadratic_mesh(nodes, tris, curve: float = 0.0, seed: int = 0):
    """Six-node (quadratic) triangles from a P1 mesh: one new node per edge, elements numbered as Ferrite's
    QuadraticTriangle (vertices 0, 1, 2, then the nodes on the edges (0,1), (1,2), (2,0)) - the meshes of
    `generate_grid(QuadraticTriangle, ...)` (_research/elliptic_chen24.jl:119-120) and of Gmsh with element_order = 2
    (src/utils.jl:20-31).  `curve` > 0 moves the interior edge nodes off the edge midpoints by up to curve * edge
    length (seeded): genuinely isoparametric elements for the tests.  Returns (nodes6, elems6)."""
    tris = np.asarray(tris, dtype=np.int64)
    n = nodes.shape[0]
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]], axis=0)
    lo, hi = e.min(axis=1), e.max(axis=1)
    key = lo * n + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    a, b = lo[first], hi[first]
    mid = 0.5 * (nodes[a] + nodes[b])
    if curve > 0:
        counts = np.bincount(inv, minlength=uniq.size)
        d = nodes[b] - nodes[a]
        nrm = np.stack([-d[:, 1], d[:, 0]], axis=1)
        rng = np.random.default_rng(seed)
        mid = mid + (counts == 2)[:, None] * (rng.random(uniq.size)[:, None] - 0.5) * 2 * curve * nrm
    T = tris.shape[0]
    eid = n + inv.reshape(3, T).T
    return np.concatenate([nodes, mid], axis=0), np.concatenate([tris, eid], axis=1)


def periodic_line_mesh(n_elems: int, order: int = 2):
    """Periodic mesh of the unit interval with `n_elems` Lagrange lines of `order` 1 or 2
    (periodic_unit_interval_discretization, src/utils.jl:42-49, with the periodic constraint condensed: the last
    element closes the ring, node 0 stands for x = 0 and x = 1).  Returns (xe, elems): quadratic elements are numbered
    (left, right, middle), the `n_elems` vertex nodes come first; xe holds the coordinates per element node
    (E x (order + 1)) because the shared node 0 has coordinate 1 as the right end of the last element."""
    h = 1.0 / n_elems
    left = np.arange(n_elems)
    right = (left + 1) % n_elems
    if order == 1:
        return np.stack([left * h, (left + 1) * h], axis=1), np.stack([left, right], axis=1)
    mid = n_elems + left
    return np.stack([left * h, (left + 1) * h, (left + 0.5) * h], axis=1), np.stack([left, right, mid], axis=1)


def p1_mass_stiffness(nodes, tris, coeff=None):
    """Lumped mass vector and stiffness matrix of P1 elements; `coeff` is an optional per-triangle diffusion
    coefficient (the piecewise-constant Darcy coefficient, src/problems/darcy.jl:39)."""
    p0, p1, p2 = nodes[tris[:, 0]], nodes[tris[:, 1]], nodes[tris[:, 2]]
    d1, d2 = p1 - p0, p2 - p0
    area2 = d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]
    area = 0.5 * np.abs(area2)
    # gradients of the hat functions: g_k = rot90(edge opposite to k) / (2 area)
    e = np.stack([p2 - p1, p0 - p2, p1 - p0], axis=1)  # (T, 3, 2)
    g = np.stack([-e[:, :, 1], e[:, :, 0]], axis=2) / area2[:, None, None]
    w = area if coeff is None else area * coeff
    Kloc = np.einsum("tid,tjd->tij", g, g) * w[:, None, None]
    n = nodes.shape[0]
    I = np.repeat(tris, 3, axis=1).ravel()
    J = np.tile(tris, (1, 3)).ravel()
    G = sp.coo_matrix((Kloc.ravel(), (I, J)), shape=(n, n)).tocsc()
    m = np.bincount(tris.ravel(), weights=np.repeat(area / 3.0, 3), minlength=n)
    G.sum_duplicates()
    G.sort_indices()
    return m, G


def matern_precision(nodes, tris, corr_range: float, sigma2: float = 1.0):
    """Matern (smoothness 1, alpha = 2) precision  Q = ratio * K' M^-1 K,  K = kappa^2 M + G, with lumped M —
    the formula of the in-repo specimen src/spdes/shallow_water.jl:177-190 (there with nu = 2 constants)."""
    m, G = p1_mass_stiffness(nodes, tris)
    nu = 1.0
    kappa = np.sqrt(8.0 * nu) / corr_range
    K = (kappa**2 * sp.diags(m) + G).tocsc()
    ratio = 1.0 / (4.0 * np.pi * kappa ** (2 * nu)) / sigma2
    Q = (ratio * (K.T @ sp.diags(1.0 / m) @ K)).tocsc()
    Q = ((Q + Q.T) * 0.5).tocsc()
    Q.sort_indices()
    return Q


def selection_observations(n: int, frac: float, seed: int = 0):
    """Point observations of a seeded Bernoulli(frac) subset of the nodes: A (m x n selection), y ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    idx = np.flatnonzero(rng.random(n) < frac)
    A = sp.csc_matrix((np.ones(idx.size), (np.arange(idx.size), idx)), shape=(idx.size, n))
    y = rng.standard_normal(idx.size)
    return A, y


def matern_posterior(nx: int, obs_frac: float = 0.1, q_eps: float = 1e2, corr_range: float = 0.05, seed: int = 0):
    """Config 4 of BASELINE.json: 2-D Matern SPDE GMRF on an nx x nx P1 mesh with point observations.
    Returns dict(Q, A, y, q_eps, Qpost, rhs, nodes)."""
    nodes, tris = structured_mesh(nx, nx, seed=seed)
    Q = matern_precision(nodes, tris, corr_range)
    A, y = selection_observations(Q.shape[0], obs_frac, seed=seed + 1)
    Qpost = (Q + q_eps * (A.T @ A)).tocsc()
    Qpost.sort_indices()
    rhs = q_eps * (A.T @ y)  # prior mean 0: posterior mean = Qpost^-1 A' Q_eps y
    return dict(Q=Q, A=A, y=y, q_eps=q_eps, Qpost=Qpost, rhs=rhs, nodes=nodes)


def heat_spacetime(nx: int, n_steps: int, dt: float = 1e-3, diffusivity: float = 1.0, tau: float = 1.0,
                   corr_range: float = 0.2, seed: int = 0):
    """Config 5 of BASELINE.json: implicit-Euler heat-equation space-time GMRF (ingredients of
    src/spdes/shallow_water.jl:198-228): residual  G x_{t+1} - M x_t ~ N(0, dt tau^2 I),  G = M + dt*kappa*K_stiff,
    initial state x_1 ~ Matern.  Returns the dense diagonal blocks D[b,b,N] and sub-diagonal blocks B[b,b,N-1]
    of the block-tridiagonal joint precision, and the same matrix assembled as sparse CSC."""
    nodes, tris = structured_mesh(nx, nx, seed=seed)
    m, Kst = p1_mass_stiffness(nodes, tris)
    M = sp.diags(m).tocsc()
    G = (M + dt * diffusivity * Kst).tocsc()
    Q0 = matern_precision(nodes, tris, corr_range)
    binv = 1.0 / (dt * tau**2)
    GtG = (binv * (G.T @ G)).tocsc()
    MM = (binv * (M.T @ M)).tocsc()
    off = (-binv * (G.T @ M)).tocsc()  # block (t+1, t)
    b = nodes.shape[0]
    N = n_steps
    diag_blocks = []
    for t in range(N):
        Dt = sp.csc_matrix((b, b))
        if t == 0:
            Dt = Dt + Q0
        else:
            Dt = Dt + GtG
        if t < N - 1:
            Dt = Dt + MM
        diag_blocks.append(Dt.tocsc())
    rows = []
    for t in range(N):
        row = [None] * N
        row[t] = diag_blocks[t]
        if t > 0:
            row[t - 1] = off
        if t < N - 1:
            row[t + 1] = off.T
        rows.append(row)
    A = sp.bmat(rows, format="csc")
    A.sort_indices()
    D = np.stack([d.toarray() for d in diag_blocks], axis=2)
    Bsub = np.stack([off.toarray()] * (N - 1), axis=2) if N > 1 else np.zeros((b, b, 0))
    return dict(A=A, D=np.asfortranarray(D), B=np.asfortranarray(Bsub), b=b, N=N, nodes=nodes)


def heat_spacetime_sparse(nx: int, n_steps: int, dt: float = 1e-3, diffusivity: float = 1.0, tau: float = 1.0,
                          corr_range: float = 0.2, seed: int = 0):
    """Config 5 without dense blocks: the same implicit-Euler heat space-time precision as `heat_spacetime`
    (src/spdes/shallow_water.jl:198-230), assembled only as a sparse matrix (time-major unknowns, block t = rows
    [t b, (t+1) b)), with 3-D coordinates (x, y, t h) for nested dissection in space-time (h = mesh width, so that a
    time step is as long as a mesh edge).  Returns dict(A, coords, b, N, nodes)."""
    nodes, tris = structured_mesh(nx, nx, seed=seed)
    m, Kst = p1_mass_stiffness(nodes, tris)
    M = sp.diags(m).tocsc()
    G = (M + dt * diffusivity * Kst).tocsc()
    Q0 = matern_precision(nodes, tris, corr_range)
    binv = 1.0 / (dt * tau**2)
    GtG = (binv * (G.T @ G)).tocsc()
    MM = (binv * (M.T @ M)).tocsc()
    off = (-binv * (G.T @ M)).tocsc()  # block (t+1, t)
    b, N = nodes.shape[0], n_steps
    T = sp.identity(N, format="csc")
    first = sp.csc_matrix(([1.0], ([0], [0])), shape=(N, N))
    last = sp.csc_matrix(([1.0], ([N - 1], [N - 1])), shape=(N, N))
    sub = sp.csc_matrix((np.ones(N - 1), (np.arange(1, N), np.arange(N - 1))), shape=(N, N))
    A = (sp.kron(first, Q0) + sp.kron(T - first, GtG) + sp.kron(T - last, MM) + sp.kron(sub, off)
         + sp.kron(sub.T, off.T)).tocsc()
    A.sort_indices()
    h = 1.0 / (nx - 1)
    coords = np.column_stack([np.tile(nodes[:, 0], N), np.tile(nodes[:, 1], N), np.repeat(np.arange(N) * h, b)])
    return dict(A=A, coords=coords, b=b, N=N, nodes=nodes)


def random_btd(b: int, N: int, seed: int = 0, coupling: float = 0.4):
    """Generic SPD block-tridiagonal test matrix with dense blocks (diagonally dominant by construction)."""
    rng = np.random.default_rng(seed)
    D = np.empty((b, b, N), order="F")
    Bs = np.empty((b, b, max(N - 1, 0)), order="F")
    for i in range(N):
        R = rng.standard_normal((b, b)) / np.sqrt(b)
        D[:, :, i] = R @ R.T + 2.0 * np.eye(b)
    for i in range(N - 1):
        Bs[:, :, i] = coupling * rng.standard_normal((b, b)) / np.sqrt(b)
    return D, Bs


def btd_to_sparse(D, Bs):
    b, _, N = D.shape
    rows = []
    for t in range(N):
        row = [None] * N
        row[t] = sp.csc_matrix(D[:, :, t])
        if t > 0:
            row[t - 1] = sp.csc_matrix(Bs[:, :, t - 1])
        if t < N - 1:
            row[t + 1] = sp.csc_matrix(Bs[:, :, t].T)
        rows.append(row)
    A = sp.bmat(rows, format="csc")
    A.sort_indices()
    return A


def burgers_like_problem(nx: int = 64, seed: int = 0, dt: float = 0.01, nu: float = 0.02):
    """A 1-D periodic, single-time-step Burgers-type collocation problem with the residual and Gauss-Newton
    tangent of scripts/solve_burger.jl:127-134:
        f(w) = A1 w - A0 w + dt (A1 w) .* (D w) - dt nu D2 w,
        J(w) = J_static + dt (Diag(D w) A1 + Diag(A1 w) D).
    Unknown w = [u_t ; u_{t+1}] (2 nx); A0/A1 select the two time levels; D, D2 are periodic central differences
    acting on u_{t+1}.  Returns dict(Q, f_and_J, y, mu, n)."""
    rng = np.random.default_rng(seed)
    h = 1.0 / nx
    I = sp.identity(nx, format="csc")
    Z = sp.csc_matrix((nx, nx))
    S = sp.csc_matrix((np.ones(nx), (np.arange(nx), (np.arange(nx) + 1) % nx)), shape=(nx, nx))
    D1 = (S - S.T) / (2 * h)
    D2m = (S + S.T - 2 * I) / h**2
    A0 = sp.hstack([I, Z]).tocsc()
    A1 = sp.hstack([Z, I]).tocsc()
    Dx = sp.hstack([Z, D1]).tocsc()
    Dxx = sp.hstack([Z, D2m]).tocsc()
    J_static = (A1 - A0 - dt * nu * Dxx).tocsc()

    def f(w):
        return A1 @ w - A0 @ w + dt * (A1 @ w) * (Dx @ w) - dt * nu * (Dxx @ w)

    def f_and_J(w):
        J = J_static + dt * (sp.diags(Dx @ w) @ A1 + sp.diags(A1 @ w) @ Dx)
        # keep a fixed pattern across iterations: union with the static pattern (explicit zeros kept)
        P = (abs(J_static) + abs(A1) + abs(Dx)).tocsc()
        P.data[:] = 0.0
        J = (J + P).tocsc()
        J.sort_indices()
        return f(w), J

    x = np.arange(nx) * h
    u0 = np.sin(2 * np.pi * x) + 0.3 * np.cos(4 * np.pi * x)
    # prior: smooth (second-difference) penalty on both time levels + a tight initial condition
    L2 = (D2m.T @ D2m) * h**3 + 1e-2 * I
    Q = sp.block_diag([L2 + 1e4 * I, L2]).tocsc()
    Q.sort_indices()
    mu = np.concatenate([u0, u0])
    y = np.zeros(nx)
    return dict(Q=Q, f_and_J=f_and_J, f=f, y=y, mu=mu, n=2 * nx)


# ----------------------------------------------------------------------------- configs 1-3 at full size ----
def elliptic_problem(nx: int = 201, corr_range: float = 0.1, seed: int = 0):
    """Config 1 of BASELINE.json (_research/elliptic_chen24.jl:118-171): nonlinear elliptic PDE  -lap u + u^3 = g  on
    the unit square, GMRF prior (Matern, range 0.1, smoothness 1), Dirichlet data as point observations of the
    boundary nodes, Gauss-Newton on the collocation residual.  P1 stand-in for the P2 Ferrite assembly on the same
    grid (nx = 201 => n = 40 401 as N_el_xy = 100 with P2).  Manufactured solution as in :60-67.
    Returns dict(Q, nodes, A_bnd, y_bnd, f_and_J, y, u_true, n)."""
    nodes, tris = structured_mesh(nx, nx, seed=seed)
    m, K = p1_mass_stiffness(nodes, tris)
    Q = matern_precision(nodes, tris, corr_range)
    x, y = nodes[:, 0], nodes[:, 1]
    u_true = np.sin(np.pi * x) * np.sin(np.pi * y) + 2.0 * np.sin(4 * np.pi * x) * np.sin(4 * np.pi * y)
    g = K @ u_true + m * u_true**3  # discrete right-hand side of the manufactured solution
    bnd = np.flatnonzero((x == 0) | (x == 1) | (y == 0) | (y == 1))
    A_bnd = sp.csc_matrix((np.ones(bnd.size), (np.arange(bnd.size), bnd)), shape=(bnd.size, nodes.shape[0]))
    Kc = K.tocsc()
    Kc.sort_indices()
    diag_pos = np.array([Kc.indptr[j] + np.searchsorted(Kc.indices[Kc.indptr[j]:Kc.indptr[j + 1]], j)
                         for j in range(Kc.shape[0])])

    def f_and_J(u):
        J = Kc.copy()  # fixed pattern: the cubic term only touches the diagonal (lumped mass)
        J.data[diag_pos] += 3.0 * m * u**2
        return Kc @ u + m * u**3, J

    # K and m: the same residual as f(u) = K u + m .* u.^3 for the device Gauss-Newton driver
    return dict(Q=Q, nodes=nodes, A_bnd=A_bnd, y_bnd=u_true[bnd], f_and_J=f_and_J, y=g, u_true=u_true,
                n=nodes.shape[0], K=Kc, m=m)


def burgers_spacetime(nx: int = 4095, nt: int = 201, dt: float = 0.01, nu: float = 0.01, tau: float = 1.0,
                      q_ic: float = 1e8, seed: int = 0):
    """Config 2 of BASELINE.json (scripts/solve_burger.jl): 1-D periodic viscous Burgers in space-time.  Prior: implicit
    Euler diffusion state-space model (block-tridiagonal precision, block size b = nx, N = nt blocks; ingredients of
    src/spdes/shallow_water.jl:198-228) conditioned on the initial condition at step 1 with Q_eps = 1e8 (:98).
    Observation model of the Gauss-Newton loop: collocation residual of every step with the tangent of :127-134,
        f(w)_t = M (u_{t+1} - u_t) + dt M (u_{t+1} .* D u_{t+1}) + dt nu K u_{t+1},      y = 0.
    Unknown w is time-major (block t = rows [t nx, (t+1) nx)).  Returns dict(Q, mu, f_and_J, y, coords, b, N, u0)."""
    rng = np.random.default_rng(seed)
    h = 1.0 / nx
    I = sp.identity(nx, format="csc")
    S = sp.csc_matrix((np.ones(nx), (np.arange(nx), (np.arange(nx) + 1) % nx)), shape=(nx, nx))
    D1 = ((S - S.T) / (2 * h)).tocsc()
    Kx = ((2 * I - S - S.T) / h).tocsc()
    M = (h * I).tocsc()
    G = (M + dt * nu * Kx).tocsc()
    kappa = 8.0
    K0 = (kappa**2 * M + Kx).tocsc()
    Q0 = (K0.T @ K0 / h).tocsc()
    binv = 1.0 / (dt * tau**2) / h
    GtG, MM, off = binv * (G.T @ G), binv * (M.T @ M), -binv * (G.T @ M)
    Tdiag = sp.identity(nt, format="csc")
    first = sp.csc_matrix(([1.0], ([0], [0])), shape=(nt, nt))
    last = sp.csc_matrix(([1.0], ([nt - 1], [nt - 1])), shape=(nt, nt))
    sub = sp.csc_matrix((np.ones(nt - 1), (np.arange(1, nt), np.arange(nt - 1))), shape=(nt, nt))
    Q = (sp.kron(first, Q0 + q_ic * I) + sp.kron(Tdiag - first, GtG) + sp.kron(Tdiag - last, MM) + sp.kron(sub, off)
         + sp.kron(sub.T, off.T)).tocsc()
    Q.sort_indices()
    xs = np.arange(nx) * h
    u0 = np.zeros(nx)
    for k in range(1, 5):  # smooth random Fourier series
        u0 += rng.standard_normal() / k * np.sin(2 * np.pi * k * xs) + rng.standard_normal() / k * np.cos(2 * np.pi * k * xs)
    mu = np.tile(u0, nt)
    # residual operator pieces: rows = steps 1..nt-1
    Enext = sp.csc_matrix((np.ones(nt - 1), (np.arange(nt - 1), np.arange(1, nt))), shape=(nt - 1, nt))
    Eprev = sp.csc_matrix((np.ones(nt - 1), (np.arange(nt - 1), np.arange(nt - 1))), shape=(nt - 1, nt))
    A_next = sp.kron(Enext, I).tocsc()
    L_static = (sp.kron(Enext, G) - sp.kron(Eprev, M)).tocsc()
    Dn = sp.kron(Enext, D1).tocsc()
    P = (abs(L_static) + abs(A_next) + abs(Dn)).tocsc()
    P.data[:] = 0.0

    def f_and_J(w):
        un, dun = A_next @ w, Dn @ w
        fx = L_static @ w + dt * h * un * dun
        J = (L_static + dt * h * (sp.diags(dun) @ A_next + sp.diags(un) @ Dn) + P).tocsc()
        J.sort_indices()
        return fx, J

    tt, xx = np.meshgrid(np.arange(nt) / max(nt - 1, 1), xs, indexing="ij")
    coords = np.stack([xx.ravel(), tt.ravel()], axis=1)
    # the same residual in the bilinear form f(w) = L w + c (A w).*(D w) the device Gauss-Newton driver takes
    return dict(Q=Q, mu=mu, f_and_J=f_and_J, y=np.zeros(nx * (nt - 1)), coords=coords, b=nx, N=nt, u0=u0,
                L=L_static, A=A_next, D=Dn, c=dt * h)


def darcy_problem(nx: int = 601, seed: int = 0, q_eps: float = 1e8):
    """Config 3 of BASELINE.json (scripts/darcy/solve_darcy_gmrf-fem.jl): Darcy flow  -div(a grad u) = 1  with a
    piecewise-constant two-level coefficient (3 / 12) looked up by nearest index on a 241 x 241 grid
    (src/datasets/darcy.jl:30-34, src/problems/darcy.jl:39) — here thresholded from a seeded smooth random field since
    the dataset is not shipped.  Prior: Matern with range 1/sqrt(300) (:97-98); observation operator A = stiffness of
    the coefficient, y = load vector, Q_eps = 1e8.  Returns dict(Q, A, y, q_eps, nodes, coeff_grid)."""
    rng = np.random.default_rng(seed)
    nodes, tris = structured_mesh(nx, nx, seed=0)  # one mesh (one pattern) for every problem of the dataset loop
    g = 241
    kx = np.fft.fftfreq(g)[:, None]
    ky = np.fft.fftfreq(g)[None, :]
    spec = np.exp(-((kx**2 + ky**2) * (g / 6.0) ** 2))
    field = np.real(np.fft.ifft2(np.fft.fft2(rng.standard_normal((g, g))) * spec))
    coeff_grid = np.where(field > 0, 12.0, 3.0)
    cent = nodes[tris].mean(axis=1)
    ij = np.clip(np.rint(cent * (g - 1)).astype(int), 0, g - 1)
    coeff = coeff_grid[ij[:, 1], ij[:, 0]]
    m, A = p1_mass_stiffness(nodes, tris, coeff=coeff)
    # homogeneous Dirichlet rows on the boundary (the reference applies them through Ferrite's constraint handler)
    x, yy = nodes[:, 0], nodes[:, 1]
    bnd = ((x == 0) | (x == 1) | (yy == 0) | (yy == 1)).astype(np.float64)
    A = (sp.diags(1.0 - bnd) @ A + sp.diags(bnd)).tocsc()
    A.sort_indices()
    y = m * (1.0 - bnd)
    Q = matern_precision(nodes, tris, 1.0 / np.sqrt(300.0))
    return dict(Q=Q, A=A, y=y, q_eps=q_eps, nodes=nodes, coeff_grid=coeff_grid)
