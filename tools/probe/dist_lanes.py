"""Local phase of the time-sharded block-tridiagonal factor on ONE GPU (rank 1 of 3: has a spike and a separator):
one-stream order (GMRFB_BTD_DIST_SERIAL=1) against the two-lane schedule (factor chain of block i + 1 on the first
stream, W_i and the spike step of block i on the second).  Usage: python tools/probe/dist_lanes.py [b] [nloc]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
pkg = g.load_pkg(); ctx = pkg.Context(0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nloc = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rng = np.random.default_rng(0)
R = rng.standard_normal((b, b)) / np.sqrt(b)
dev = torch.device("cuda", 0)
Dl = torch.from_numpy(np.ascontiguousarray((R @ R.T + 2 * np.eye(b)).T)).to(dev).unsqueeze(0).repeat(nloc, 1, 1).contiguous()
Bl = torch.from_numpy(np.ascontiguousarray((0.4 * R).T)).to(dev).unsqueeze(0).repeat(nloc, 1, 1).contiguous()
ifaces = {}
for mode in ("serial", "lanes", "serial", "lanes"):
    if mode == "serial":
        os.environ["GMRFB_BTD_DIST_SERIAL"] = "1"
    else:
        os.environ.pop("GMRFB_BTD_DIST_SERIAL", None)
    torch.cuda.synchronize(); t = time.perf_counter()
    ts = pkg.dist.TimeShardedCholesky(Dl, Bl, 1, 3, ctx=ctx, auto_exchange=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    ifaces[mode] = ts.iface().cpu().numpy().copy()
    flops = (19.0 / 3.0) * b**3 * (nloc - 1)
    print(f"b={b} nloc={nloc} {mode:6s}: local factor + spikes {dt*1e3:8.1f} ms  ({flops/dt*1e-12:.1f} TFLOP/s on 19/3 b^3 per block)")
    del ts
d = np.max(np.abs(ifaces["serial"] - ifaces["lanes"])) / np.max(np.abs(ifaces["serial"]))
print("max rel difference of the interface blocks, serial vs lanes:", d)
