"""Block-tridiagonal factor / solve / selected-inversion timing with the per-kernel profile (tuning aid)."""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=2048)
ap.add_argument("--N", type=int, default=8)
ap.add_argument("--nrhs", type=int, default=1)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--selinv", action="store_true")
args = ap.parse_args()
pkg = g.load_pkg()
W = pkg.workloads
ctx = pkg.default_context()
b, N = args.b, args.N
rng = np.random.default_rng(0)
# cheap SPD blocks: diagonally dominant
D = np.empty((b, b, N), order="F")
Bs = np.empty((b, b, N - 1), order="F")
R = rng.standard_normal((b, b)) / np.sqrt(b)
base = R @ R.T + 2.0 * np.eye(b)
for i in range(N):
    D[:, :, i] = base
for i in range(N - 1):
    Bs[:, :, i] = 0.4 * R
out = {}
for rep in range(args.reps):
    ctx.profile_begin() if rep == args.reps - 1 else None
    t = time.perf_counter()
    F = pkg.tridiagonal_cholesky_dense(D, Bs, ctx=ctx)
    ctx.sync()
    dt = time.perf_counter() - t
    if rep == args.reps - 1:
        prof = ctx.profile_end()
        out["factor_profile"] = sorted(prof, key=lambda p: -p["ms"])
    if rep < args.reps - 1:
        del F
flops = F.info.flops
ktot = sum(p["ms"] for p in out["factor_profile"])
print(f"b={b} N={N}: factor wall {dt*1e3:.1f} ms (incl. H2D of {D.nbytes/1e9+Bs.nbytes/1e9:.2f} GB); kernels {ktot:.1f} ms -> {flops/ktot*1e-9:.2f} TFLOP/s")
for p in out["factor_profile"]:
    print(f"   {p['name']:24s} n={p['launches']:5d} ms={p['ms']:9.3f} TF={p['flops']/p['ms']*1e-9 if p['ms']>0 else 0:7.2f}")
rhs = rng.standard_normal((b * N, args.nrhs))
x = pkg.ldiv(F, rhs)
ctx.profile_begin()
t = time.perf_counter()
x = pkg.ldiv(F, rhs)
dt = time.perf_counter() - t
prof = ctx.profile_end()
ktot = sum(p["ms"] for p in prof)
print(f"solve nrhs={args.nrhs}: wall {dt*1e3:.1f} ms kernels {ktot:.2f} ms -> {24.0*b*b*N/ktot*1e-6:.1f} GB/s of factor bytes")
for p in sorted(prof, key=lambda p: -p["ms"]):
    print(f"   {p['name']:24s} n={p['launches']:5d} ms={p['ms']:9.3f}")
if args.selinv:
    ctx.profile_begin()
    v = F.selinv_diag()
    prof = ctx.profile_end()
    ktot = sum(p["ms"] for p in prof)
    fl = sum(p["flops"] for p in prof)
    print(f"selinv: kernels {ktot:.1f} ms -> {fl/ktot*1e-9:.2f} TFLOP/s")
    for p in sorted(prof, key=lambda p: -p["ms"]):
        print(f"   {p['name']:24s} n={p['launches']:5d} ms={p['ms']:9.3f} TF={p['flops']/p['ms']*1e-9 if p['ms']>0 else 0:7.2f}")
