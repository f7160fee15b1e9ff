#!/usr/bin/env python
"""Config 5 (256 x 256 nodes x N_t implicit-Euler steps): computed memory / flop budget of a subtree-to-GPU mapping.

Host analysis only (no GPU): the library's own ordering (space-time nested dissection) and supernodal symbolic analysis
through tools/plancheck, then for P GPUs the supernodal elimination tree is cut into a TOP part (the separators shared by
several GPUs, to be factorised as distributed dense fronts) and P local subtree sets by proportional mapping (largest
subtree first, split until there are >= 4 P pieces, pieces assigned to the least-loaded GPU by flops).  Reported per
GPU: flops, memory of (a) the shipped layout (every d x d front resident), (b) a panel layout (d x s panels resident +
the peak of the update-matrix stack along a postorder traversal).  This replaces the n^(4/3) extrapolation of DESIGN.md
section 5 by computed numbers.

    python tools/config5_budget.py --steps 16 32 64 --gpus 8 --out profiles/r02_config5_budget.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "plancheck"))
import __graft_entry__ as entry  # noqa: E402
import plancheck as pc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=256)
ap.add_argument("--steps", type=int, nargs="+", default=[16, 32])
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--out", default="")
args = ap.parse_args()
W = entry.load_pkg().workloads
res = []
for N in args.steps:
    st = W.heat_spacetime_sparse(args.nx, N)
    t = time.time()
    P = pc.Plans(st["A"], ordering="nd", coords=st["coords"])
    t_an = time.time() - t
    ns = P.nsuper
    d = (P.rptr[1:] - P.rptr[:-1]).astype(np.float64)
    s = (P.sptr[1:] - P.sptr[:-1]).astype(np.float64)
    r = d - s
    flops = s**3 / 3 + s * s * r + s * r * (r + 1)          # partial factorisation of a d x d front with s pivots
    full = d * d                                            # doubles, shipped layout
    panel = d * s
    upd = r * r
    parent = P.sparent.astype(np.int64)
    # subtree sums (children have smaller indices: postordered)
    sub_flops, sub_full, sub_panel = flops.copy(), full.copy(), panel.copy()
    for j in range(ns):
        p = parent[j]
        if p >= 0:
            sub_flops[p] += sub_flops[j]
            sub_full[p] += sub_full[j]
            sub_panel[p] += sub_panel[j]
    kids = [[] for _ in range(ns)]
    roots = []
    for j in range(ns):
        (kids[parent[j]] if parent[j] >= 0 else roots).append(j)
    # proportional mapping: split the heaviest piece until there are >= 4 P pieces
    pieces, top = list(roots), []
    while len(pieces) < 4 * args.gpus:
        j = max(pieces, key=lambda q: sub_flops[q])
        if not kids[j]:
            break
        pieces.remove(j)
        top.append(j)
        pieces += kids[j]
    load = np.zeros(args.gpus)
    mem_full = np.zeros(args.gpus)
    mem_panel = np.zeros(args.gpus)
    owner = {}
    for j in sorted(pieces, key=lambda q: -sub_flops[q]):
        g = int(np.argmin(load))
        owner[j] = g
        load[g] += sub_flops[j]
        mem_full[g] += sub_full[j]
        mem_panel[g] += sub_panel[j]

    # peak of the update-matrix stack inside one piece (postorder: a front's update stays until its parent is assembled)
    def stack_peak(root):
        peak, cur = 0.0, 0.0
        order, stk = [], [(root, 0)]
        while stk:  # iterative postorder
            v, i = stk.pop()
            if i < len(kids[v]):
                stk.append((v, i + 1))
                stk.append((kids[v][i], 0))
            else:
                order.append(v)
        for v in order:
            cur += full[v]              # the front is assembled (children's updates still live)
            peak = max(peak, cur)
            cur -= full[v]
            cur -= sum(upd[c] for c in kids[v])
            cur += upd[v]
        return peak

    peak_stack = np.zeros(args.gpus)
    for j, g in owner.items():
        peak_stack[g] = max(peak_stack[g], stack_peak(j))
    # the shipped schedule runs level by level (all fronts of a tree level in one grouped launch): with a level-lifetime
    # stack a GPU holds the full fronts of the current level of its pieces + the updates still waiting for their parents
    height = np.zeros(ns, dtype=np.int64)
    for j in range(ns):
        if parent[j] >= 0:
            height[parent[j]] = max(height[parent[j]], height[j] + 1)
    piece_of = np.full(ns, -1, dtype=np.int64)          # GPU that owns supernode j (-1: top)
    for j in range(ns - 1, -1, -1):
        if j in owner:
            piece_of[j] = owner[j]
        elif parent[j] >= 0 and piece_of[parent[j]] >= 0:
            piece_of[j] = piece_of[parent[j]]
    nlev = int(height.max()) + 1
    lev_front = np.zeros((args.gpus, nlev))
    lev_upd = np.zeros((args.gpus, nlev + 1))           # updates live while level in (height[j], height[parent]]
    for j in range(ns):
        g = piece_of[j]
        if g < 0:
            continue
        lev_front[g, height[j]] += full[j]
        hp = height[parent[j]] if parent[j] >= 0 and piece_of[parent[j]] >= 0 else nlev - 1
        lev_upd[g, height[j] + 1:hp + 1] += upd[j]
    level_peak = (lev_front + lev_upd[:, :nlev]).max(axis=1)
    top = np.array(sorted(top))
    out = dict(nx=args.nx, n_steps=N, n=int(P.n), nsuper=int(ns), analyze_s=round(t_an, 1), nnz_L=float(panel.sum() - (s * (s - 1) / 2).sum()),
               factor_flops=float(flops.sum()), front_arena_gb=float(full.sum() * 8e-9), panel_gb=float(panel.sum() * 8e-9),
               max_front=int(d.max()), gpus=args.gpus, pieces=len(pieces), top_supernodes=int(top.size),
               top_flops=float(flops[top].sum()), top_flops_frac=float(flops[top].sum() / flops.sum()),
               top_fronts_gb=float(full[top].sum() * 8e-9), top_panels_gb=float(panel[top].sum() * 8e-9),
               top_max_front=int(d[top].max()) if top.size else 0,
               per_gpu_flops=[float(x) for x in load], per_gpu_full_gb=[float(x * 8e-9) for x in mem_full],
               per_gpu_panel_gb=[float(x * 8e-9) for x in mem_panel], per_gpu_stack_peak_gb=[float(x * 8e-9) for x in peak_stack],
               per_gpu_level_stack_peak_gb=[float(x * 8e-9) for x in level_peak])
    out["fits_shipped_layout"] = bool(max(out["per_gpu_full_gb"]) + out["top_fronts_gb"] / args.gpus < 170)
    out["fits_panel_layout"] = bool(max(a + b for a, b in zip(out["per_gpu_panel_gb"], out["per_gpu_stack_peak_gb"])) +
                                    out["top_panels_gb"] / args.gpus < 170)
    out["fits_panel_layout_level_schedule"] = bool(max(a + b for a, b in zip(out["per_gpu_panel_gb"], out["per_gpu_level_stack_peak_gb"])) +
                                                   out["top_panels_gb"] / args.gpus < 170)
    out["seconds_at_25_tflops_per_gpu"] = float(max(load) / 25e12 + flops[top].sum() / (25e12 * args.gpus))
    print(json.dumps(out), flush=True)
    res.append(out)
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
