"""Time-sharded block-tridiagonal factor + solve across ranks (launch with torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_btd_dist.py --b 2048 --N 64

Every rank builds the same synthetic SPD block-tridiagonal system, keeps its slab, factors and solves; rank 0
additionally factors the whole system on its own GPU (sequential algorithm) when it fits, and the residual of the
distributed solution is checked on every rank's rows.  Times are CUDA-event times, max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=1024)
ap.add_argument("--N", type=int, default=64)
ap.add_argument("--nrhs", type=int, default=8)
ap.add_argument("--seq", action="store_true", help="also time the sequential factor on rank 0")
ap.add_argument("--first-weight", type=float, default=2.15,
                help="rank 0 runs the plain recurrence (7/3 b^3 per block, the others 19/3): it takes this many times the blocks")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = g.load_pkg()
ctx = pkg.Context(local)
b, N = args.b, args.N
rng = np.random.default_rng(0)
R = rng.standard_normal((b, b)) / np.sqrt(b)
Dblk = R @ R.T + 2.0 * np.eye(b)
Bblk = 0.4 * R
bounds = pkg.dist.slab_bounds(N, world, first_weight=args.first_weight)
lo, hi = bounds[rank]
nloc = hi - lo
# device-resident slab (element [k, j, i] = entry (i, j) of block k): the factor timing below is GPU work only
dev = torch.device("cuda", local)
Dl = torch.from_numpy(np.ascontiguousarray(Dblk.T)).to(dev).unsqueeze(0).repeat(nloc, 1, 1).contiguous()
Bl = torch.from_numpy(np.ascontiguousarray(Bblk.T)).to(dev).unsqueeze(0).repeat(nloc, 1, 1).contiguous()
rhs = np.random.default_rng(1).standard_normal((b * N, args.nrhs))


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn):
    sync()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return out, float(dt.item())


phases = {}
for rep in range(2):
    sync()
    if rep == 1:
        ctx.profile_begin()  # CUDA events around every launch: GPU time of the local phase without allocations
    t0 = time.perf_counter()
    ts = pkg.dist.TimeShardedCholesky(Dl, Bl, rank, world, ctx=ctx, auto_exchange=False)  # local factor + spikes
    torch.cuda.synchronize()
    t1l = time.perf_counter()
    kern_ms = sum(p["ms"] for p in ctx.profile_end()) if rep == 1 else 0.0
    if world > 1:
        dist.barrier()  # so that the exchange timer below does not include waiting for the slowest rank's local phase
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    send = ts.iface()
    gathered = ts._allgather(send)                                                        # 3 b^2 doubles per rank
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ts.reduce(gathered)                                                                   # reduced (P-1)-block chain
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    tt = torch.tensor([t3 - t0, t1l - t0, t2 - t1, t3 - t2, kern_ms * 1e-3], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_fac = float(tt[0])
    phases = {"local_factor_and_spikes_s": float(tt[1]), "local_kernels_s": float(tt[4]), "exchange_s": float(tt[2]),
              "reduced_system_s": float(tt[3]),
              "note": "wall times, max over ranks; local phase = kernels + device allocations of the slab's factor, "
                      "inverse and spike buffers (several GB per rank: concurrent cudaMalloc/cudaFree of the ranks "
                      "serialise in the driver)"}
    x, t_sol = timed(lambda: ts.solve(rhs[lo * b:hi * b]))
    if rep == 0:
        del ts, send, gathered
# residual of this rank's rows: needs neighbours' solution rows -> gather the full solution (small: b*N*nrhs)
xt = torch.from_numpy(np.ascontiguousarray(x)).to(f"cuda:{local}")
if world > 1:
    sizes = [(h - l) * b for l, h in bounds]
    parts = [torch.empty((s, args.nrhs), dtype=torch.float64, device=f"cuda:{local}") for s in sizes]
    dist.all_gather(parts, xt)
    xfull = torch.cat(parts).cpu().numpy()
else:
    xfull = x
res = 0.0
for k in range(lo, hi):
    r = Dblk @ xfull[k * b:(k + 1) * b]
    if k > 0:
        r += Bblk @ xfull[(k - 1) * b:k * b]
    if k < N - 1:
        r += Bblk.T @ xfull[(k + 1) * b:(k + 2) * b]
    res = max(res, float(np.linalg.norm(r - rhs[k * b:(k + 1) * b]) / np.linalg.norm(rhs[k * b:(k + 1) * b])))
rt = torch.tensor([res], device=f"cuda:{local}", dtype=torch.float64)
if world > 1:
    dist.all_reduce(rt, op=dist.ReduceOp.MAX)
logdet = ts.logdet()
seq = None
if args.seq and rank == 0:
    Dall = torch.from_numpy(np.ascontiguousarray(Dblk.T)).to(dev).unsqueeze(0).repeat(N, 1, 1).contiguous()
    Ball = torch.from_numpy(np.ascontiguousarray(Bblk.T)).to(dev).unsqueeze(0).repeat(max(N - 1, 1), 1, 1).contiguous()
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        F = pkg.tridiagonal_cholesky_dense(Dall, Ball, ctx=ctx)
        ctx.sync()
        t_seq = time.perf_counter() - t
        if rep == 0:
            del F
    seq = {"factor_s": t_seq, "logdet": F.logdet()}
if rank == 0:
    flops_seq = (N - 1) * 7.0 / 3.0 * b**3 + b**3 / 3.0
    print(json.dumps({"b": b, "N": N, "world": world, "nrhs": args.nrhs, "factor_s": t_fac, "solve_s": t_sol,
                      "factor_incl_h2d": False, "factor_phases": phases, "seq_equiv_tflops": flops_seq / t_fac * 1e-12,
                      "max_rel_residual": float(rt.item()), "logdet": logdet, "sequential": seq,
                      "blocks_per_rank": [h - l for l, h in bounds]}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
