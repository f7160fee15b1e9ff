"""Block-tridiagonal factor: serial schedule against the look-ahead schedule (GMRFB_BTD_LOOKAHEAD), wall time of the
device work (four-block input, gmrfb_btd_factor_ssm: no host->device traffic besides 4 b^2 doubles)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, nargs="+", default=[2048, 4096])
ap.add_argument("--N", type=int, default=24)
args = ap.parse_args()
pkg = g.load_pkg()
ctx = pkg.default_context()
for b in args.b:
    rng = np.random.default_rng(0)
    R = rng.standard_normal((b, b)) / np.sqrt(b)
    D = R @ R.T + 2.0 * np.eye(b)
    Bs = 0.4 * R
    rhs = rng.standard_normal((b * args.N, 3))
    dev = torch.device("cuda", 0)
    Dd = torch.from_numpy(np.ascontiguousarray(D.T)).to(dev)
    Bd = torch.from_numpy(np.ascontiguousarray(Bs.T)).to(dev)
    ref = None
    for la in ("0", "1"):
        os.environ["GMRFB_BTD_LOOKAHEAD"] = la
        ts = []
        for rep in range(3):
            ctx.sync()
            t = time.perf_counter()
            F = pkg.tridiagonal_cholesky_ssm(Dd, Dd, Dd, Bd, args.N, ctx=ctx)
            ctx.sync()
            ts.append(time.perf_counter() - t)
            if rep < 2:
                del F
        x = pkg.ldiv(F, rhs)
        if ref is None:
            ref = x
        flops = F.info.flops
        print(json.dumps({"b": b, "N": args.N, "lookahead": la, "factor_s": min(ts), "tflops": flops / min(ts) * 1e-12,
                          "logdet": F.logdet(), "max_diff_vs_serial": float(np.max(np.abs(x - ref)))}), flush=True)
        del F
