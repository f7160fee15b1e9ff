"""Per-kernel-kind time split of the serial block-tridiagonal factor (GMRFB_BTD_LOOKAHEAD=0, per-launch events)."""
import json
import os
import sys

import numpy as np
import torch

os.environ["GMRFB_BTD_LOOKAHEAD"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g

pkg = g.load_pkg()
ctx = pkg.default_context()
b, N = int(os.environ.get("B", "4096")), int(os.environ.get("NBLK", "6"))
rng = np.random.default_rng(0)
R = rng.standard_normal((b, b)) / np.sqrt(b)
D = R @ R.T + 2.0 * np.eye(b)
dev = torch.device("cuda", 0)
Dd = torch.from_numpy(np.ascontiguousarray(D.T)).to(dev)
Bd = torch.from_numpy(np.ascontiguousarray((0.4 * R).T)).to(dev)
F = pkg.tridiagonal_cholesky_ssm(Dd, Dd, Dd, Bd, N, ctx=ctx)
del F
ctx.sync()
ctx.profile_begin()
F = pkg.tridiagonal_cholesky_ssm(Dd, Dd, Dd, Bd, N, ctx=ctx)
ctx.sync()
recs = ctx.profile_end()
tot = sum(r["ms"] for r in recs)
print("b", b, "N", N, "total ms", tot, "per block", tot / N, "flops", F.info.flops, "TF (sum of kernel times)",
      F.info.flops / tot * 1e-9)
for r in sorted(recs, key=lambda r: -r["ms"]):
    print("%-28s n=%5d ms=%8.3f share=%5.1f%% TF=%6.2f avg_us=%7.1f" % (r["name"], r["launches"], r["ms"], 100 * r["ms"] / tot,
                                                                       r["flops"] / max(r["ms"], 1e-9) * 1e-9,
                                                                       1e3 * r["ms"] / max(r["launches"], 1)))
