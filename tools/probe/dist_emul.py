"""Emulated-rank time-sharded solve on one GPU: residual for host / device inputs and several nrhs (debug aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
pkg = g.load_pkg()
ctx = pkg.Context(0)
dev = torch.device("cuda", 0)
for b, N, P in ((1024, 16, 2), (256, 8, 2), (1024, 16, 4)):
    rng = np.random.default_rng(0)
    R = rng.standard_normal((b, b)) / np.sqrt(b)
    Dblk = R @ R.T + 2.0 * np.eye(b)
    Bblk = 0.4 * R
    bounds = pkg.dist.slab_bounds(N, P)
    for mode in ("device", "host"):
        for nrhs in (3, 8, 16):
            ranks = []
            for r, (lo, hi) in enumerate(bounds):
                nloc = hi - lo
                if mode == "device":
                    Dl = torch.from_numpy(np.ascontiguousarray(Dblk.T)).to(dev).unsqueeze(0).repeat(nloc, 1, 1).contiguous()
                    Bl = torch.from_numpy(np.ascontiguousarray(Bblk.T)).to(dev).unsqueeze(0).repeat(nloc, 1, 1).contiguous()
                else:
                    Dl = np.asfortranarray(np.repeat(Dblk[:, :, None], nloc, axis=2))
                    Bl = np.asfortranarray(np.repeat(Bblk[:, :, None], nloc, axis=2))
                ranks.append(pkg.dist.TimeShardedCholesky(Dl, Bl, r, P, ctx=ctx, auto_exchange=False))
            gathered = torch.cat([t.iface() for t in ranks])
            for t in ranks:
                t.reduce(gathered)
            rhs = np.random.default_rng(1).standard_normal((b * N, nrhs))
            sends = torch.cat([t.solve_begin(rhs[lo * b:hi * b]) for t, (lo, hi) in zip(ranks, bounds)])
            X = np.vstack([t.solve_end(sends) for t in ranks])
            res = 0.0
            worst = -1
            for k in range(N):
                rr = Dblk @ X[k * b:(k + 1) * b]
                if k > 0:
                    rr += Bblk @ X[(k - 1) * b:k * b]
                if k < N - 1:
                    rr += Bblk.T @ X[(k + 1) * b:(k + 2) * b]
                e = float(np.linalg.norm(rr - rhs[k * b:(k + 1) * b]) / np.linalg.norm(rhs[k * b:(k + 1) * b]))
                if e > res:
                    res, worst = e, k
            print(f"b={b} N={N} P={P} input={mode} nrhs={nrhs}: max rel residual {res:.3e} at block {worst}", flush=True)
            del ranks
