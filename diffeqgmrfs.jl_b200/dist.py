"""Time-sharded block-tridiagonal Cholesky across ranks (one process per GPU).

The space-time precision of an implicit-Euler GMRF is block tridiagonal along the time axis
(src/spdes/shallow_water.jl:219-230; the factor of src/tridiagonal_cholesky.jl:65-82).  Rank r owns a contiguous
slab of time blocks; ranks 0..P-2 use their last block as a separator.  All dense work is local CUDA
(libgmrfb, gmrfb_btd_dist_*); the only exchange is an all-gather of three b x b blocks per rank for the factor and
two b x nrhs panels per rank for a solve, done here with torch.distributed (NCCL over NVLink on GPUs).

`slab_bounds(N, P)` gives the block partition; `TimeShardedCholesky` runs the phases.  The phases are also exposed
individually (`iface`, `reduce`, `solve_begin`, `solve_end`) so that P ranks can be driven from one process in
tests.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib as B
from .solver import Context, default_context


def slab_bounds(n_blocks: int, world: int, first_weight: float = 1.0):
    """Contiguous slabs [lo, hi) per rank (every rank but the last needs >= 2 blocks).  With ``first_weight`` = 1 the
    sizes are as equal as possible.  Rank 0 has no spike to eliminate — it does 7/3 b^3 flops per block where the other
    ranks do 19/3 b^3 — so a ``first_weight`` > 1 gives it that many times the blocks of the others and balances the local
    phases (flop ratio 19/7 = 2.7; measured time ratio per block at b = 4096 on B200: 2.15, which is what bench.py uses)."""
    if world == 1:
        return [(0, n_blocks)]
    if first_weight == 1.0:
        base, rem = divmod(n_blocks, world)
        sizes = [base + (1 if r < rem else 0) for r in range(world)]
    else:
        per = n_blocks / (first_weight + world - 1)
        sizes = [max(2, int(round(per))) for _ in range(world)]
        sizes[0] = n_blocks - sum(sizes[1:])
    out, lo = [], 0
    for r in range(world):
        out.append((lo, lo + sizes[r]))
        lo += sizes[r]
    if any(hi - lo < 2 for lo, hi in out[:-1]) or out[-1][1] - out[-1][0] < 1 or out[-1][1] != n_blocks:
        raise ValueError("too few time blocks for this many ranks")
    return out


def local_blocks(D, Bsub, lo, hi):
    """Slice the global dense blocks (D[b,b,N], Bsub[b,b,N-1]) into the rank-local arrays of the C ABI."""
    b = D.shape[0]
    Dl = np.asfortranarray(D[:, :, lo:hi])
    Bl = np.zeros((b, b, hi - lo), order="F")
    for k in range(lo, hi):
        if k > 0:
            Bl[:, :, k - lo] = Bsub[:, :, k - 1]
    return Dl, Bl


class TimeShardedCholesky:
    """One rank of the time-sharded factor.  `slab` may be a backend object implementing the four phases
    (`iface`, `reduce`, `solve_begin`, `solve_end`, `logdet_parts`) on CPU tensors — used by the gloo tests to run
    this orchestration code without a GPU; by default the phases are the CUDA library's."""

    def __init__(self, D_local, B_local, rank: int, world: int, ctx: Context | None = None, group=None,
                 auto_exchange: bool = True, slab=None):
        import torch

        self.torch = torch
        self.rank, self.world, self.group = rank, world, group
        self._slab = slab
        if slab is not None:
            self.b, self.nloc = slab.b, slab.nloc
            if auto_exchange:
                self.reduce(self._allgather(self.iface()))
            return
        self.ctx = ctx or default_context()
        if torch.is_tensor(D_local):
            # device-resident blocks: CUDA float64 tensors of shape (nloc, b, b), C-contiguous, element [k, j, i] =
            # entry (i, j) of block k (each block column-major) - the memory layout of the NumPy form below
            assert D_local.is_cuda and B_local.is_cuda and D_local.is_contiguous() and B_local.is_contiguous()
            assert D_local.dtype == torch.float64 and B_local.dtype == torch.float64
            torch.cuda.current_stream(D_local.device).synchronize()  # the blocks may still be in the making on torch's stream
            self.nloc, self.b, _ = D_local.shape
            dp, bp = C.cast(C.c_void_p(D_local.data_ptr()), B._F64P), C.cast(C.c_void_p(B_local.data_ptr()), B._F64P)
        else:
            D_local = np.asfortranarray(D_local, dtype=np.float64)
            B_local = np.asfortranarray(B_local, dtype=np.float64)
            self.b, _, self.nloc = D_local.shape
            dp, bp = D_local.ctypes.data_as(B._F64P), B_local.ctypes.data_as(B._F64P)
        h = C.c_void_p()
        st = B.lib().gmrfb_btd_dist_create(self.ctx.h, rank, world, self.b, self.nloc, dp, bp, C.byref(h))
        B.check(st, self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_btd_dist_destroy, h)
        self.dev = torch.device("cuda", self.ctx.device)
        if auto_exchange:
            self.reduce(self._allgather(self.iface()))

    # ---- phases ----
    def iface(self):
        if self._slab is not None:
            return self._slab.iface()
        cnt = int(B.lib().gmrfb_btd_dist_iface_count(self.h))
        send = self.torch.empty(cnt, dtype=self.torch.float64, device=self.dev)
        self._sync_torch()
        B.check(B.lib().gmrfb_btd_dist_get_iface(self.h, C.c_void_p(send.data_ptr())), self.ctx.h)
        return send

    def reduce(self, gathered):
        if self._slab is not None:
            return self._slab.reduce(gathered)
        self._sync_torch()
        ptr = C.c_void_p(gathered.data_ptr()) if gathered is not None else None
        B.check(B.lib().gmrfb_btd_dist_reduce(self.h, ptr), self.ctx.h)

    def solve_begin(self, X_local):
        if self._slab is not None:
            return self._slab.solve_begin(X_local)
        X = np.asfortranarray(np.asarray(X_local, dtype=np.float64).reshape(self.b * self.nloc, -1))
        self._nrhs = X.shape[1]
        cnt = int(B.lib().gmrfb_btd_dist_solve_count(self.h, self._nrhs))
        # torch.empty, not zeros: a fill kernel queued on torch's stream is NOT ordered against the library's own
        # (non-blocking) stream and could land after the library has written the buffer; the library clears it itself
        send = self.torch.empty(max(cnt, 1), dtype=self.torch.float64, device=self.dev)[:cnt]
        self._sync_torch()
        B.check(B.lib().gmrfb_btd_dist_solve_begin(self.h, X.ctypes.data_as(B._F64P), X.shape[0], self._nrhs,
                                                   C.c_void_p(send.data_ptr())), self.ctx.h)
        return send

    def solve_end(self, gathered):
        if self._slab is not None:
            return self._slab.solve_end(gathered)
        self._sync_torch()
        n = self.b * self.nloc
        X = np.empty((n, self._nrhs), order="F")
        ptr = C.c_void_p(gathered.data_ptr()) if gathered is not None else None
        B.check(B.lib().gmrfb_btd_dist_solve_end(self.h, ptr, X.ctypes.data_as(B._F64P), n, self._nrhs), self.ctx.h)
        return X

    # ---- collective wrappers ----
    def _sync_torch(self):
        """The library runs on its own non-blocking CUDA stream: work torch (or NCCL) has queued on ITS stream for a buffer
        that is about to be handed to the library must have finished first.  (The other direction is covered by the C ABI:
        every call returns with its outputs complete.)"""
        if self._slab is None:
            self.torch.cuda.current_stream(self.dev).synchronize()

    def _allgather(self, send):
        if self.world == 1:
            return send
        import torch.distributed as dist

        out = self.torch.empty(self.world * send.numel(), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(out, send, group=self.group)
        # the collective is ordered on torch's current stream only: wait for it before the library reads `out`
        self._sync_torch()
        return out

    def solve(self, X_local):
        """This rank's rows of A^{-1} X (intended semantics of ldiv, src/tridiagonal_cholesky.jl:54-63)."""
        one = np.asarray(X_local).ndim == 1
        X = self.solve_end(self._allgather(self.solve_begin(X_local)))
        return X[:, 0].copy() if one else X

    def logdet_parts(self):
        if self._slab is not None:
            return self._slab.logdet_parts()
        loc, red = C.c_double(), C.c_double()
        B.check(B.lib().gmrfb_btd_dist_logdet(self.h, C.byref(loc), C.byref(red)), self.ctx.h)
        return loc.value, red.value

    def logdet(self):
        loc, red = self.logdet_parts()
        total_local = loc
        if self.world > 1:
            import torch.distributed as dist

            dev = self.dev if self._slab is None else "cpu"
            t = self.torch.tensor([loc], dtype=self.torch.float64, device=dev)
            dist.all_reduce(t, group=self.group)
            total_local = float(t.item())
        return total_local + red


# ------------------------------------------------------------------- sample-sharded RBMC variances / samples --
# SURVEY.md §8(e), row "multi-RHS samples / RBMC": every rank holds the same factor (factorised redundantly: one
# numeric factorisation is cheaper than shipping 8 nnz(L) bytes), the N sample columns are split across ranks, and the
# only exchange is one all-reduce of an n-vector.  The Rao-Blackwellised estimator is an average over samples,
#   var_i = 1/Q_ii + mean_k t_ik^2 / Q_ii^2,
# so the global estimate is the sample-count-weighted average of the per-rank estimates: no new device code is needed,
# each rank calls gmrfb_var_rbmc on its columns (scripts/darcy/solve_darcy_gmrf-fem.jl:192, RBMCStrategy(50)).
def sample_bounds(n_samples: int, world: int):
    """Contiguous sample-column ranges [lo, hi) per rank, sizes as equal as possible (ranks may be empty)."""
    base, rem = divmod(n_samples, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def rbmc_variance_sharded(factor, Q, Z, rank: int, world: int, group=None, device=None):
    """RBMC marginal variances with the sample columns of ``Z`` (n x N, the same array on every rank: the reference
    threads one seeded generator through ``RBMCStrategy(N; rng)``) split across ranks.  ``factor`` needs a
    ``var_rbmc(Q, Z_cols)`` method (``CholeskyFactor``; a NumPy stand-in in the gloo tests).  Returns the full n-vector
    on every rank; with ``world == 1`` it is exactly ``factor.var_rbmc(Q, Z)``."""
    import torch

    Z = np.asarray(Z)
    n, N = Z.shape
    lo, hi = sample_bounds(N, world)[rank]
    part = np.zeros(n)
    if hi > lo:
        part = np.asarray(factor.var_rbmc(Q, np.asfortranarray(Z[:, lo:hi]))) * ((hi - lo) / N)
    if world == 1:
        return part
    import torch.distributed as dist

    t = torch.from_numpy(part)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def rand_sharded(factor, Z, rank: int, world: int, mean=None, gather=False, group=None, device=None):
    """Posterior samples mean + P' L^{-T} z for this rank's columns of ``Z`` (n x N); no collective unless
    ``gather`` asks for all N columns on every rank (scripts/darcy/solve_darcy_gmrf-fem.jl:191 in a sample loop)."""
    import torch

    Z = np.asarray(Z)
    n, N = Z.shape
    bounds = sample_bounds(N, world)
    lo, hi = bounds[rank]
    X = np.zeros((n, 0))
    if hi > lo:
        X = np.asarray(factor.sample(np.asfortranarray(Z[:, lo:hi]), mean=mean)).reshape(n, hi - lo)
    if not gather or world == 1:
        return X
    import torch.distributed as dist

    # ragged column counts: gather equal-sized (padded) panels and trim
    cmax = max(h - l for l, h in bounds)
    mine = torch.zeros((cmax, n), dtype=torch.float64)
    mine[:hi - lo] = torch.from_numpy(np.ascontiguousarray(X.T))
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in bounds]
    dist.all_gather(parts, mine, group=group)
    return torch.cat([p[:h - l] for p, (l, h) in zip(parts, bounds)]).cpu().numpy().T
