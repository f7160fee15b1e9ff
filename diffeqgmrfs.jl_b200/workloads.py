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
