"""Timing breakdown of one rank's phases of the time-sharded block-tridiagonal factor on a single GPU (tuning aid)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
pkg = g.load_pkg(); ctx = pkg.Context(0)
b = int(sys.argv[1]); dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
R = rng.standard_normal((b, b)) / np.sqrt(b)
Dblk = R @ R.T + 2.0 * np.eye(b); Bblk = 0.4 * R
for nloc, rank, world in [(32, 1, 2), (16, 1, 4), (16, 0, 4), (16, 3, 4)]:
    Dl = torch.from_numpy(np.ascontiguousarray(Dblk.T)).to(dev).unsqueeze(0).repeat(nloc, 1, 1).contiguous()
    Bl = torch.from_numpy(np.ascontiguousarray(Bblk.T)).to(dev).unsqueeze(0).repeat(nloc, 1, 1).contiguous()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ts = pkg.dist.TimeShardedCholesky(Dl, Bl, rank, world, ctx=ctx, auto_exchange=False)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        send = ts.iface()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        allb = send.repeat(world)
        try:
            ts.reduce(allb)  # the repeated interface blocks need not form an SPD reduced system: timing only
        except Exception:
            pass
        torch.cuda.synchronize(); t3 = time.perf_counter()
        if rep == 0: del ts
    print(f"b={b} nloc={nloc} rank={rank}/{world}: create {t1-t0:.3f}s iface {t2-t1:.3f}s reduce {t3-t2:.3f}s", flush=True)
    ctx.profile_begin()
    ts2 = pkg.dist.TimeShardedCholesky(Dl, Bl, rank, world, ctx=ctx, auto_exchange=False)
    prof = ctx.profile_end()
    tot = sum(p["ms"] for p in prof)
    print("   kernels %.1f ms:" % tot, [(p["name"], p["launches"], round(p["ms"], 1)) for p in sorted(prof, key=lambda p: -p["ms"])[:5]])
    del ts, ts2, Dl, Bl
