"""World-size-2 (and 3) gloo runs of the time-sharded block-tridiagonal orchestration on CPU.

The product's numeric phases are CUDA only; here the orchestration code of diffeqgmrfs.jl_b200/dist.py
(slab partition, all-gather protocol, packing order of the interface blocks) runs under torch.distributed/gloo with
a NumPy slab backend that restates the per-rank algebra of csrc/btd.cu (gmrfb_btd_dist_*), and the assembled
result is compared with the sequential oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpySlab:
    """Per-rank algebra of the time-sharded factor (mirrors gmrfb_btd_dist_create/reduce/solve_*), on the CPU."""

    def __init__(self, D, Bl, rank, world):
        import torch

        self.torch = torch
        self.rank, self.P = rank, world
        self.b, _, self.nloc = D.shape
        b = self.b
        self.has_sep, self.has_spike = rank < world - 1, rank > 0
        self.ni = self.nloc - 1 if self.has_sep else self.nloc
        # interior factor: L_1 = chol(D_1); C_i = B_i L_{i-1}^{-T}; L_i = chol(D_i - C_i C_i')
        self.L, self.C = [np.linalg.cholesky(D[:, :, 0])], [None]
        for i in range(1, self.ni):
            Ci = sla.solve_triangular(self.L[-1], Bl[:, :, i].T, lower=True).T
            self.C.append(Ci)
            self.L.append(np.linalg.cholesky(D[:, :, i] - Ci @ Ci.T))
        Q = np.zeros((b, b))
        R = np.zeros((b, b))
        i0 = np.zeros((b, b))
        if self.has_spike:
            self.W = [sla.solve_triangular(self.L[0], Bl[:, :, 0], lower=True).T]  # E_l' L_1^{-T}
            for i in range(1, self.ni):
                self.W.append(sla.solve_triangular(self.L[i], (-self.W[-1] @ self.C[i].T).T, lower=True).T)
            Q = sum(w @ w.T for w in self.W)
        if self.has_sep:
            self.V = sla.solve_triangular(self.L[-1], Bl[:, :, self.nloc - 1].T, lower=True).T  # E_r L_ni^{-T}
            i0 = D[:, :, self.nloc - 1] - self.V @ self.V.T
            if self.has_spike:
                R = self.V @ self.W[-1].T
        self._iface = np.concatenate([i0.ravel(order="F"), Q.ravel(order="F"), R.ravel(order="F")])

    def iface(self):
        return self.torch.from_numpy(self._iface.copy())

    def reduce(self, gathered):
        b, P = self.b, self.P
        if P == 1:
            return
        g = gathered.numpy().reshape(P, 3, b, b).transpose(0, 1, 3, 2)  # column-major blocks
        self.RL, self.RC = [], [None]
        for r in range(P - 1):
            Dh = g[r, 0] - g[r + 1, 1]
            Dh = np.tril(Dh) + np.tril(Dh, -1).T  # the device keeps lower triangles only
            if r > 0:
                Cr = sla.solve_triangular(self.RL[-1], (-g[r, 2]).T, lower=True).T
                self.RC.append(Cr)
                Dh = Dh - Cr @ Cr.T
            self.RL.append(np.linalg.cholesky(Dh))

    def solve_begin(self, X):
        b = self.b
        X = np.asarray(X, dtype=np.float64).reshape(b * self.nloc, -1)
        self.nrhs = X.shape[1]
        self.y = [None] * self.ni
        for i in range(self.ni):
            rhs = X[i * b:(i + 1) * b].copy()
            if i > 0:
                rhs -= self.C[i] @ self.y[i - 1]
            self.y[i] = sla.solve_triangular(self.L[i], rhs, lower=True)
        s0 = np.zeros((b, self.nrhs))
        u = np.zeros((b, self.nrhs))
        if self.has_sep:
            s0 = X[self.ni * b:] - self.V @ self.y[-1]
        if self.has_spike:
            u = sum(w @ y for w, y in zip(self.W, self.y))
        return self.torch.from_numpy(np.concatenate([s0.ravel(), u.ravel()]))

    def solve_end(self, gathered):
        b, P = self.b, self.P
        xs = None
        if P > 1:
            g = gathered.numpy().reshape(P, 2, b, self.nrhs)
            bh = [g[r, 0] - g[r + 1, 1] for r in range(P - 1)]
            z = []
            for r in range(P - 1):
                rhs = bh[r] - (self.RC[r] @ z[r - 1] if r > 0 else 0)
                z.append(sla.solve_triangular(self.RL[r], rhs, lower=True))
            xs = [None] * (P - 1)
            for r in range(P - 2, -1, -1):
                rhs = z[r] - (self.RC[r + 1].T @ xs[r + 1] if r < P - 2 else 0)
                xs[r] = sla.solve_triangular(self.RL[r], rhs, lower=True, trans="T")
        y = [v.copy() for v in self.y]
        if self.has_spike:
            for i in range(self.ni):
                y[i] -= self.W[i].T @ xs[self.rank - 1]
        if self.has_sep:
            y[-1] -= self.V.T @ xs[self.rank]
        x = [None] * self.ni
        for i in range(self.ni - 1, -1, -1):
            rhs = y[i] - (self.C[i + 1].T @ x[i + 1] if i < self.ni - 1 else 0)
            x[i] = sla.solve_triangular(self.L[i], rhs, lower=True, trans="T")
        if self.has_sep:
            x.append(xs[self.rank])
        return np.vstack(x)

    def logdet_parts(self):
        loc = 2.0 * sum(np.sum(np.log(np.diag(L))) for L in self.L)
        red = 2.0 * sum(np.sum(np.log(np.diag(L))) for L in self.RL) if self.P > 1 else 0.0
        return float(loc), float(red)


def _worker(rank, world, port, b, N, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import __graft_entry__ as g

    pkg = g.load_pkg()
    orc = g.load_oracle()
    W = pkg.workloads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D, Bs = W.random_btd(b, N, seed=7)
    lo, hi = pkg.dist.slab_bounds(N, world)[rank]
    Dl, Bl = pkg.dist.local_blocks(D, Bs, lo, hi)
    ts = pkg.dist.TimeShardedCholesky(None, None, rank, world, slab=NumpySlab(Dl, Bl, rank, world))
    rhs = np.random.default_rng(3).standard_normal((b * N, 2))
    x_local = ts.solve(rhs[lo * b:hi * b])
    logdet = ts.logdet()
    Fo = orc.tridiagonal_cholesky(W.btd_to_sparse(D, Bs), N)
    want = np.stack([orc.btd_ldiv(Fo, rhs[:, k]) for k in range(2)], 1)[lo * b:hi * b]
    err = np.linalg.norm(x_local - want) / np.linalg.norm(want)
    lerr = abs(logdet - orc.btd_logdet(Fo)) / abs(orc.btd_logdet(Fo))
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{err} {lerr}\n")
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,b,N", [(2, 6, 7), (3, 5, 10)])
def test_time_sharded_orchestration_gloo(tmp_path, world, b, N):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), b, N, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err, lerr = map(float, open(tmp_path / f"rank{r}.txt").read().split())
        assert err < 1e-10 and lerr < 1e-12, (r, err, lerr)


class NumpyFactor:
    """CPU stand-in with the two methods the sample-sharded helpers call (restated RBMC / sampling of the oracle)."""

    def __init__(self, orc, Q, perm):
        self.orc, self.Q = orc, Q
        self.chol = orc.SparseCholesky(Q, perm)

    def var_rbmc(self, Q, Z):
        return self.orc.rbmc_variance(self.chol, Q, Z)

    def sample(self, Z, mean=None):
        X = self.chol.solve_UP(Z)
        return X if mean is None else X + np.asarray(mean)[:, None]


def _worker_rbmc(rank, world, port, nsamp, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import __graft_entry__ as g

    pkg = g.load_pkg()
    orc = g.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = pkg.workloads.matern_posterior(9, obs_frac=0.3, corr_range=0.3)
    Q = prob["Qpost"]
    n = Q.shape[0]
    perm = np.random.default_rng(5).permutation(n)
    fac = NumpyFactor(orc, Q, perm)
    Z = np.random.default_rng(11).standard_normal((n, nsamp))  # the same seeded stream on every rank
    var = pkg.dist.rbmc_variance_sharded(fac, Q, Z, rank, world)
    want = orc.rbmc_variance(fac.chol, Q, Z)
    mean = np.linspace(0.0, 1.0, n)
    X = pkg.dist.rand_sharded(fac, Z, rank, world, mean=mean, gather=True)
    Xw = fac.sample(Z, mean=mean)
    lo, hi = pkg.dist.sample_bounds(nsamp, world)[rank]
    Xl = pkg.dist.rand_sharded(fac, Z, rank, world, mean=mean)
    e1 = float(np.max(np.abs(var - want) / want))
    e2 = float(np.max(np.abs(X - Xw)))
    e3 = float(np.max(np.abs(Xl - Xw[:, lo:hi]))) if hi > lo else 0.0
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{e1} {e2} {e3}\n")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nsamp", [(2, 7), (3, 2)])
def test_sample_sharded_rbmc_gloo(tmp_path, world, nsamp):
    """sample columns split across ranks (ragged, and one rank empty in the second case), one
    all-reduce of the n-vector; the result must equal the single-process estimate over all columns."""
    import torch.multiprocessing as mp

    mp.spawn(_worker_rbmc, args=(world, _free_port(), nsamp, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        e1, e2, e3 = map(float, open(tmp_path / f"rank{r}.txt").read().split())
        assert e1 < 1e-12 and e2 < 1e-13 and e3 < 1e-13, (r, e1, e2, e3)


def test_slab_bounds(pkg=None):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    d = g.load_pkg().dist
    assert d.slab_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert d.slab_bounds(1024, 8)[-1] == (896, 1024)
    with pytest.raises(ValueError):
        d.slab_bounds(3, 3)
    # load-balanced split: rank 0 (no spike: 7/3 b^3 per block instead of 19/3) takes 19/7 times the blocks of the others
    for n, w in ((256, 2), (256, 8), (64, 4), (16, 3)):
        bb = d.slab_bounds(n, w, first_weight=19.0 / 7.0)
        sizes = [hi - lo for lo, hi in bb]
        assert bb[0][0] == 0 and bb[-1][1] == n and all(a[1] == b[0] for a, b in zip(bb, bb[1:]))
        assert min(sizes) >= 2 and sizes[0] == max(sizes) and max(sizes[1:]) - min(sizes[1:]) == 0
        if n >= 64:
            assert abs(sizes[0] / sizes[1] - 19.0 / 7.0) < 0.5
    assert d.slab_bounds(5, 1, first_weight=3.0) == [(0, 5)]
    assert d.sample_bounds(50, 8) == [(0, 7), (7, 14), (14, 20), (20, 26), (26, 32), (32, 38), (38, 44), (44, 50)]
    assert d.sample_bounds(2, 3) == [(0, 1), (1, 2), (2, 2)]
