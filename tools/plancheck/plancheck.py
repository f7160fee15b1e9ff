"""Static launch plans of libgmrfb on the CPU: hazard check and NumPy interpretation.

The numeric phases of the library are static lists of launches (csrc/plan.cpp); the tasks of ONE launch run concurrently
on the GPU and the kernels use no atomics ("one child rank per launch"), so the plans are correct only if, inside every
launch, no two tasks write the same arena entry and no task reads an entry another task of that launch writes.
compute-sanitizer's racecheck is refused on the GPU pool (profiles/r02_sanitizer.md); this tool checks that property on
the host, at task granularity, from the same Task records the kernels consume - and interprets the plans with NumPy
(every launch kind restated from its kernel in csrc/kernels.cu) so that the model of reads and writes is itself checked:
the interpreted factor must satisfy L L' = P A P', the interpreted selected inverse must equal inv(P A P') on the
pattern of L.  TEST INFRASTRUCTURE (tests/test_plan_interpreter.py); nothing here is used by the product.

    python tools/plancheck/plancheck.py            # a few meshes, prints launch / task / hazard statistics
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "diffeqgmrfs.jl_b200", "csrc")
SO = "/tmp/libplancheck.so"

# launch kinds / flags of csrc/tasks.hpp
(LK_GEMM_NT, LK_GEMM_NN, LK_GEMM_TN, LK_POTRF, LK_TRSM_RLT, LK_TRSM_RLN, LK_EXTEND_ADD, LK_GATHER_SYM, LK_SET_IDENTITY,
 LK_TRANSPOSE, LK_SCALE, LK_DIAG_OUT, LK_SYMMETRIZE, LK_FRONT_FACTOR_SMALL, LK_FRONT_SELINV_SMALL, LK_GEMM_TT) = range(16)
LK_MR_FWD_SMALL, LK_MR_BWD_SMALL, LK_MR_ASSEMBLE, LK_MR_GATHER, LK_ZERO_FRONT = 32, 33, 34, 35, 36
TF_TRI, TF_NEG, TF_KLOW, TF_B_DINV, TF_NOFACTOR, TF_BLOW, TF_BUPP = 1 << 8, 1 << 9, 1 << 10, 1 << 11, 1 << 12, 1 << 13, 1 << 14
DINV_ARENA = 4          # pseudo arena index of the 64 x 64 inverse-block slots
ZDIAG_ARENA = 5         # pseudo arena index of the diag(Z) output vector
DINV_SLOT = 64 * 64
SMALL_FRONT_MAX = 152
KIND_NAME = {0: "GEMM_NT", 1: "GEMM_NN", 2: "GEMM_TN", 3: "POTRF", 4: "TRSM_RLT", 5: "TRSM_RLN", 6: "EXTEND_ADD",
             7: "GATHER_SYM", 8: "SET_IDENTITY", 9: "TRANSPOSE", 10: "SCALE", 11: "DIAG_OUT", 12: "SYMMETRIZE",
             13: "FRONT_FACTOR_SMALL", 14: "FRONT_SELINV_SMALL", 15: "GEMM_TT", 32: "MR_FWD_SMALL", 33: "MR_BWD_SMALL",
             34: "MR_ASSEMBLE", 35: "MR_GATHER", 36: "ZERO_FRONT"}

TASK = np.dtype([("a", "<i8"), ("b", "<i8"), ("c", "<i8"), ("M", "<i4"), ("N", "<i4"), ("K", "<i4"), ("lda", "<i4"),
                 ("ldb", "<i4"), ("ldc", "<i4"), ("tile0", "<i4"), ("flags", "<i4"), ("aux0", "<i4"), ("aux1", "<i4"),
                 ("alpha", "<f8"), ("beta", "<f8")])
LAUNCH = np.dtype([("kind", "<i4"), ("task0", "<i4"), ("ntasks", "<i4"), ("grid", "<i4"), ("flops", "<f8"),
                   ("bytes", "<f8"), ("smem", "<i4"), ("cfg", "<i4"), ("wait_ev", "<i2"), ("rec_ev", "<i2"),
                   ("_pad", "<i4")])


def build():
    src = [os.path.join(ROOT, "tools", "plancheck", "plancheck.cpp"), os.path.join(CSRC, "symbolic.cpp"),
           os.path.join(CSRC, "plan.cpp")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) > os.path.getmtime(f) for f in src + [os.path.join(CSRC, "tasks.hpp"),
                                                                                              os.path.join(CSRC, "plan.hpp")]):
        return
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I", CSRC, "-I", "/usr/local/cuda/include"] + src +
                          ["-o", SO, "-lpthread"])


def _lib():
    build()
    L = C.CDLL(SO)
    P = C.c_void_p
    L.pc_create.restype = P
    L.pc_create.argtypes = [C.c_int64, P, P, P, C.c_int, C.c_int, P]
    L.pc_error.restype = C.c_char_p
    L.pc_error.argtypes = [P]
    L.pc_destroy.argtypes = [P]
    L.pc_sizes.argtypes = [P, P]
    L.pc_arrays.argtypes = [P] * 12
    L.pc_set_wide.restype = C.c_int64
    L.pc_set_wide.argtypes = [P, C.c_int, C.c_int, P]
    L.pc_get_wide.argtypes = [P, P, P, P]
    L.pc_build.argtypes = [P, C.c_int, C.c_int, P]
    L.pc_get.argtypes = [P, C.c_int, P, P, P]
    L.pc_build_mr.argtypes = [P, C.c_int, C.c_int, P]
    assert L.pc_sizeof_task() == TASK.itemsize and L.pc_sizeof_launch() == LAUNCH.itemsize
    return L


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Plans:
    """Symbolic analysis + plans of a symmetric CSC matrix `A` (both triangles stored), through the library's host code."""

    ORDER = {"given": 0, "natural": 1, "nd": 2, "amd": 3, "nd_amd": 4}

    def __init__(self, A, ordering="nd", perm=None, coords=None, wide_min=128, wide_max=4096):
        self.L = _lib()
        A = sp.csc_matrix(A)
        A.sort_indices()
        self.A = A
        n = A.shape[0]
        cp, ri = A.indptr.astype(np.int64), A.indices.astype(np.int64)
        pm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int64)
        co = None if coords is None else np.ascontiguousarray(coords, dtype=np.float64)
        kind = 0 if perm is not None else self.ORDER[ordering]
        self.h = self.L.pc_create(n, _p(cp), _p(ri), _p(pm), kind, 0 if co is None else co.shape[1], _p(co))
        err = self.L.pc_error(self.h).decode()
        if err:
            raise ValueError(err)
        sz = np.zeros(6, dtype=np.int64)
        self.L.pc_sizes(self.h, _p(sz))
        self.n, self.nsuper, self.arena, nrows, self.nnzA, nci = (int(v) for v in sz)
        self.sptr = np.zeros(self.nsuper + 1, dtype=np.int32)
        self.rptr = np.zeros(self.nsuper + 1, dtype=np.int64)
        self.rows = np.zeros(nrows, dtype=np.int32)
        self.relmap = np.zeros(nrows, dtype=np.int32)
        self.ld = np.zeros(self.nsuper, dtype=np.int32)
        self.foff = np.zeros(self.nsuper, dtype=np.int64)
        self.sparent = np.zeros(self.nsuper, dtype=np.int32)
        self.child_ptr = np.zeros(self.nsuper + 1, dtype=np.int32)
        self.child_idx = np.zeros(max(nci, 1), dtype=np.int32)
        self.amap = np.zeros(self.nnzA, dtype=np.int64)
        self.perm = np.zeros(self.n, dtype=np.int32)
        self.L.pc_arrays(self.h, _p(self.sptr), _p(self.rptr), _p(self.rows), _p(self.relmap), _p(self.ld), _p(self.foff),
                         _p(self.sparent), _p(self.child_ptr), _p(self.child_idx), _p(self.amap), _p(self.perm))
        out = np.zeros(1, dtype=np.int64)
        nw = int(self.L.pc_set_wide(self.h, wide_min, wide_max, _p(out)))
        self.wide_doubles = int(out[0])
        self.wide = np.zeros(nw, dtype=np.int32)
        self.wide_off = np.zeros(nw, dtype=np.int64)
        self.wide_ld = np.zeros(nw, dtype=np.int32)
        if nw:
            self.L.pc_get_wide(self.h, _p(self.wide), _p(self.wide_off), _p(self.wide_ld))
        self.plans = {}

    def __del__(self):
        try:
            self.L.pc_destroy(self.h)
        except Exception:  # noqa: BLE001
            pass

    def d(self, s):
        return int(self.rptr[s + 1] - self.rptr[s])

    def sc(self, s):
        return int(self.sptr[s + 1] - self.sptr[s])

    def plan(self, which, use_wide=False):
        """which: 'zero' | 'factor' | 'selinv' | 'wide' -> dict(tasks, launches, scratch, dinv, winv_slot)."""
        key = (which, use_wide)
        if key in self.plans:
            return self.plans[key]
        idx = {"zero": 0, "factor": 1, "selinv": 2, "wide": 3}[which]
        out = np.zeros(5, dtype=np.int64)
        self.L.pc_build(self.h, idx, int(use_wide), _p(out))
        nt, nl = int(out[0]), int(out[1])
        tasks = np.zeros(max(nt, 1), dtype=TASK)
        launches = np.zeros(max(nl, 1), dtype=LAUNCH)
        winv = np.full(self.nsuper, -1, dtype=np.int64)
        self.L.pc_get(self.h, idx, _p(tasks), _p(launches), _p(winv))
        self.plans[key] = dict(tasks=tasks[:nt], launches=launches[:nl], scratch=int(out[2]), dinv=int(out[3]),
                               kept=int(out[4]), winv_slot=winv)
        return self.plans[key]


def _mr_plans(self, nr, ldk):
    """Panel sweeps for nr right-hand sides (leading dimension ldk): (forward plan, backward plan); also sets
    self.uoff (first update-panel row of every supernode), self.mr = (nr, ldk), self.urows."""
    out = np.zeros(6, dtype=np.int64)
    self.L.pc_build_mr(self.h, int(nr), int(ldk), _p(out))
    res = []
    for which, (nt, nl) in ((4, (int(out[0]), int(out[1]))), (5, (int(out[2]), int(out[3])))):
        tasks = np.zeros(max(nt, 1), dtype=TASK)
        launches = np.zeros(max(nl, 1), dtype=LAUNCH)
        self.L.pc_get(self.h, which, _p(tasks), _p(launches), None)
        res.append(dict(tasks=tasks[:nt], launches=launches[:nl], scratch=0, dinv=int(out[4]), kept=0, winv_slot=None))
    self.plans.pop(("factor", False), None)  # rebuilt by the C side (identical)
    self.uoff = np.concatenate([[0], np.cumsum([self.d(s) - self.sc(s) for s in range(self.nsuper)])]).astype(np.int64)
    self.urows = int(out[5])
    self.mr = (int(nr), int(ldk))
    return res


Plans.mr_plans = _mr_plans


# --------------------------------------------------------------------------------------------- state ----
class State:
    """The arenas the plans address: 0 fronts, 1 inverse fronts, 2 work scratch, 3 kept wide inverses (+ TRTRI scratch),
    4 the 64 x 64 inverse-block slots, 5 diag(Z)."""

    def __init__(self, P: Plans, poison=True):
        need_work = max([P.plan(w, uw)["scratch"] for w, uw in (("factor", False), ("selinv", False), ("selinv", True))] + [1])
        need_dinv = max([P.plan(w, uw)["dinv"] for w, uw in (("factor", False), ("selinv", False), ("selinv", True),
                                                              ("wide", False))] + [DINV_SLOT])
        fill = np.nan if poison else 0.0   # poison: anything read before it is written shows up as NaN in the results
        self.ar = [np.full(P.arena, fill), np.full(P.arena, fill), np.full(need_work, fill),
                   np.full(max(2 * P.wide_doubles, 1), fill), np.full(need_dinv, fill), np.full(P.n, fill)]


def _arena(t, shift):
    return (int(t["flags"]) >> shift) & 3


def _view(arr, off, ld, M, N):
    return np.lib.stride_tricks.as_strided(arr[int(off):], shape=(int(M), int(N)), strides=(8, 8 * int(ld)))


def _idx(off, ld, rows, cols):
    """flat arena indices of the entries (rows x cols) of a column-major matrix at `off`."""
    return (int(off) + np.asarray(rows, dtype=np.int64)[:, None] + np.asarray(cols, dtype=np.int64)[None, :] * int(ld)).ravel()


def _tri_idx(off, ld, M, N):
    """entries (i, j) with i >= j of an M x N matrix (what a TF_TRI result touches)."""
    i, j = np.meshgrid(np.arange(M), np.arange(N), indexing="ij")
    m = i >= j
    return int(off) + i[m].astype(np.int64) + j[m].astype(np.int64) * int(ld)


# ------------------------------------------------------------------------------- per-task semantics ----
def task_access(P: Plans, kind, t):
    """(reads, writes): lists of (arena, flat index array) a task touches - restated from the kernels."""
    f = int(t["flags"])
    M, N, K = int(t["M"]), int(t["N"]), int(t["K"])
    aa, ab, ac = _arena(t, 0), _arena(t, 2), _arena(t, 4)
    R, Wr = [], []
    if kind in (LK_GEMM_NT, LK_GEMM_NN, LK_GEMM_TN, LK_GEMM_TT):
        ta = kind in (LK_GEMM_TN, LK_GEMM_TT)          # A stored K x M
        tb = kind in (LK_GEMM_NT, LK_GEMM_TT)          # B stored N x K
        ar_, ac_ = (K, M) if ta else (M, K)
        br_, bc_ = (N, K) if tb else (K, N)
        if f & TF_KLOW:                                  # A (K x M) lower triangular: only k >= m is read
            R.append((aa, _tri_idx(t["a"], t["lda"], ar_, ac_)))
        else:
            R.append((aa, _idx(t["a"], t["lda"], range(ar_), range(ac_))))
        R.append((ab, _idx(t["b"], t["ldb"], range(br_), range(bc_))))
        cidx = _tri_idx(t["c"], t["ldc"], M, N) if f & TF_TRI else _idx(t["c"], t["ldc"], range(M), range(N))
        if float(t["beta"]) != 0.0:
            R.append((ac, cidx))
        Wr.append((ac, cidx))
    elif kind == LK_POTRF:
        R.append((aa, _tri_idx(t["a"], t["lda"], M, M)))
        if not f & TF_NOFACTOR:
            Wr.append((aa, _tri_idx(t["a"], t["lda"], M, M)))
        if f & TF_B_DINV:
            Wr.append((DINV_ARENA, _idx(t["b"], 64, range(64), range(64))))
        else:
            Wr.append((ab, _idx(t["b"], t["ldb"], range(M), range(M))))
    elif kind in (LK_TRSM_RLT, LK_TRSM_RLN):
        if f & TF_B_DINV:
            R.append((DINV_ARENA, _tri_idx(t["b"], 64, 64, 64)))
        else:
            R.append((ab, _tri_idx(t["b"], t["ldb"], N, N)))
        x = _idx(t["c"], t["ldc"], range(M), range(N))
        R.append((ac, x))
        Wr.append((ac, x))
    elif kind == LK_EXTEND_ADD:
        rel = P.relmap[(int(t["aux1"]) << 32 | (int(t["aux0"]) & 0xffffffff)):][:M].astype(np.int64)
        R.append((aa, _tri_idx(t["a"], t["lda"], M, M)))
        i, j = np.meshgrid(np.arange(M), np.arange(M), indexing="ij")
        m = i >= j
        p = int(t["c"]) + rel[i[m]] + rel[j[m]] * int(t["ldc"])
        R.append((ac, p))
        Wr.append((ac, p))
    elif kind == LK_GATHER_SYM:
        rel = P.relmap[(int(t["aux1"]) << 32 | (int(t["aux0"]) & 0xffffffff)):][:M].astype(np.int64)
        i, j = np.meshgrid(np.arange(M), np.arange(M), indexing="ij")
        a, b = np.maximum(rel[i], rel[j]), np.minimum(rel[i], rel[j])
        R.append((aa, (int(t["a"]) + a + b * int(t["lda"])).ravel()))
        Wr.append((ac, _idx(t["c"], t["ldc"], range(M), range(M))))
    elif kind in (LK_SET_IDENTITY, LK_SCALE):
        x = _idx(t["c"], t["ldc"], range(M), range(N))
        if kind == LK_SCALE:
            R.append((ac, x))
        Wr.append((ac, x))
    elif kind == LK_DIAG_OUT:
        R.append((ac, int(t["c"]) + np.arange(M, dtype=np.int64) * (int(t["ldc"]) + 1)))
        Wr.append((ZDIAG_ARENA, (int(t["aux1"]) << 32 | (int(t["aux0"]) & 0xffffffff)) + np.arange(M, dtype=np.int64)))
    elif kind == LK_ZERO_FRONT:
        if f & TF_TRI:  # whole 64 x 64 tiles that meet the lower triangle
            i, j = np.meshgrid(np.arange(M), np.arange(N), indexing="ij")
            m = (i // 64) >= (j // 64)
            Wr.append((ac, int(t["c"]) + i[m].astype(np.int64) + j[m].astype(np.int64) * int(t["ldc"])))
        else:
            Wr.append((ac, _idx(t["c"], t["ldc"], range(M), range(N))))
    elif kind == LK_FRONT_FACTOR_SMALL:
        s = int(t["aux0"])
        d, sc, ld, fo = P.d(s), P.sc(s), int(P.ld[s]), int(P.foff[s])
        R.append((0, _tri_idx(fo, ld, d, sc)))                       # the assembled panel
        for c in P.child_idx[P.child_ptr[s]:P.child_ptr[s + 1]]:
            dc, scc = P.d(c), P.sc(c)
            R.append((0, _tri_idx(int(P.foff[c]) + scc * int(P.ld[c]) + scc, int(P.ld[c]), dc - scc, dc - scc)))
        Wr.append((0, _tri_idx(fo, ld, d, d)))                       # L and the update matrix
    elif kind == LK_FRONT_SELINV_SMALL:
        s = int(t["aux0"])
        d, sc, ld, fo = P.d(s), P.sc(s), int(P.ld[s]), int(P.foff[s])
        R.append((0, _tri_idx(fo, ld, d, sc)))
        r = d - sc
        if r > 0:
            p = int(P.sparent[s])
            rel = P.relmap[int(P.rptr[s]) + sc:int(P.rptr[s]) + d].astype(np.int64)
            i, j = np.meshgrid(np.arange(r), np.arange(r), indexing="ij")
            m = i >= j
            R.append((1, int(P.foff[p]) + rel[i[m]] + rel[j[m]] * int(P.ld[p])))
        Wr.append((1, _tri_idx(fo, ld, d, d)))
        Wr.append((ZDIAG_ARENA, int(P.sptr[s]) + np.arange(sc, dtype=np.int64)))
    elif kind in (LK_MR_FWD_SMALL, LK_MR_BWD_SMALL, LK_MR_ASSEMBLE, LK_MR_GATHER):
        s = int(t["aux0"])
        nr, ldk = P.mr
        d, sc = P.d(s), P.sc(s)
        q = np.arange(nr, dtype=np.int64)

        def panel(first_col, cols):   # entries (q, first_col + cols) of a node-major panel
            return (q[:, None] + (int(first_col) + np.asarray(cols, dtype=np.int64))[None, :] * ldk).ravel()

        xj = panel(P.sptr[s], np.arange(sc))
        uj = panel(P.uoff[s], np.arange(d - sc))
        below = P.rows[int(P.rptr[s]) + sc:int(P.rptr[s]) + d].astype(np.int64)
        kids = P.child_idx[P.child_ptr[s]:P.child_ptr[s + 1]]
        if kind in (LK_MR_FWD_SMALL, LK_MR_BWD_SMALL):
            R.append((0, _tri_idx(P.foff[s], P.ld[s], d, sc)))
        if kind == LK_MR_FWD_SMALL:
            R.append((1, xj))
            for c in kids:
                R.append((2, panel(P.uoff[c], np.arange(P.d(c) - P.sc(c)))))
            Wr.append((1, xj))
            Wr.append((2, uj))
        elif kind == LK_MR_ASSEMBLE:
            for c in kids:
                R.append((2, panel(P.uoff[c], np.arange(P.d(c) - P.sc(c)))))
                rel = P.relmap[int(P.rptr[c]) + P.sc(c):int(P.rptr[c]) + P.d(c)].astype(np.int64)
                into = panel(P.sptr[s], rel[rel < sc])
                R.append((1, into))
                Wr.append((1, into))
            Wr.append((2, uj))
        elif kind == LK_MR_GATHER:
            R.append((1, panel(0, below)))
            Wr.append((2, uj))
        else:
            R.append((1, xj))
            R.append((1, panel(0, below)))
            Wr.append((1, xj))
    else:
        raise NotImplementedError(f"launch kind {kind}")
    return R, Wr


def hazards(P: Plans, plan):
    """Check every launch of `plan`: distinct tasks write disjoint entries and no task reads what another task of the same
    launch writes.  Returns (number of launches, number of tasks, list of violations)."""
    sizes = {}
    bad = []
    ntasks = 0
    for li, L in enumerate(plan["launches"]):
        kind = int(L["kind"])
        writer = {}
        acc = []
        for k in range(int(L["ntasks"])):
            t = plan["tasks"][int(L["task0"]) + k]
            R, Wr = task_access(P, kind, t)
            acc.append((R, Wr))
            ntasks += 1
            for a, ix in Wr:
                w = writer.setdefault(a, {})
                ix = np.unique(ix)
                key = ix.tobytes()
                del key
                arr = w.setdefault("idx", [])
                arr.append((k, ix))
        # write / write
        for a, w in writer.items():
            allix = np.concatenate([ix for _, ix in w["idx"]])
            owner = np.concatenate([np.full(ix.size, k, dtype=np.int32) for k, ix in w["idx"]])
            order = np.argsort(allix, kind="stable")
            sx, so = allix[order], owner[order]
            dup = np.flatnonzero((sx[1:] == sx[:-1]) & (so[1:] != so[:-1]))
            if dup.size:
                bad.append(("write/write", li, KIND_NAME.get(kind, kind), a, int(so[dup[0]]), int(so[dup[0] + 1]), int(sx[dup[0]])))
            sizes[a] = (sx, so)
        # read / write between different tasks
        for k, (R, _) in enumerate(acc):
            for a, ix in R:
                if a not in sizes or a not in writer:
                    continue
                sx, so = sizes[a]
                ix = np.unique(ix)
                if sx.size == 0 or ix.size == 0:
                    continue
                pos = np.searchsorted(sx, ix)
                pos[pos >= sx.size] = sx.size - 1
                hit = (sx[pos] == ix) & (so[pos] != k)
                # an entry may be written by several positions (same owner); check every owner at that index
                if hit.any():
                    bad.append(("read/write", li, KIND_NAME.get(kind, kind), a, k, int(so[pos[np.flatnonzero(hit)[0]]]),
                                int(ix[np.flatnonzero(hit)[0]])))
        sizes.clear()
    return len(plan["launches"]), ntasks, bad


# -------------------------------------------------------------------------------------- interpreter ----
def _run_task(P: Plans, S: State, kind, t):
    f = int(t["flags"])
    M, N, K = int(t["M"]), int(t["N"]), int(t["K"])
    ar = S.ar
    aa, ab, ac = _arena(t, 0), _arena(t, 2), _arena(t, 4)
    if kind in (LK_GEMM_NT, LK_GEMM_NN, LK_GEMM_TN, LK_GEMM_TT):
        ta = kind in (LK_GEMM_TN, LK_GEMM_TT)
        tb = kind in (LK_GEMM_NT, LK_GEMM_TT)
        A = _view(ar[aa], t["a"], t["lda"], K if ta else M, M if ta else K)
        B = _view(ar[ab], t["b"], t["ldb"], N if tb else K, K if tb else N)
        A = np.tril(A) if f & TF_KLOW else np.array(A)
        opA = A.T if ta else A
        opB = B.T if tb else B
        Cv = _view(ar[ac], t["c"], t["ldc"], M, N)
        res = float(t["alpha"]) * (opA @ opB)
        if float(t["beta"]) != 0.0:
            res = res + float(t["beta"]) * Cv
        if f & TF_TRI:
            i, j = np.meshgrid(np.arange(M), np.arange(N), indexing="ij")
            m = i >= j
            Cv[m] = res[m]
        else:
            Cv[:, :] = res
    elif kind == LK_POTRF:
        Av = _view(ar[aa], t["a"], t["lda"], M, M)
        Lm = np.tril(Av)
        if not f & TF_NOFACTOR:
            full = Lm + np.tril(Lm, -1).T
            Lm = np.linalg.cholesky(full)
            i, j = np.tril_indices(M)
            Av[i, j] = Lm[i, j]
        Wm = np.linalg.solve(Lm, np.eye(M))
        Wm = np.tril(Wm)
        if f & TF_B_DINV:
            slot = _view(ar[DINV_ARENA], t["b"], 64, 64, 64)
            slot[:, :] = np.eye(64)
            slot[:M, :M] = Wm
        else:
            _view(ar[ab], t["b"], t["ldb"], M, M)[:, :] = Wm
    elif kind in (LK_TRSM_RLT, LK_TRSM_RLN):
        Wm = np.tril(_view(ar[DINV_ARENA], t["b"], 64, 64, 64)[:N, :N]) if f & TF_B_DINV else \
            np.tril(_view(ar[ab], t["b"], t["ldb"], N, N))
        X = _view(ar[ac], t["c"], t["ldc"], M, N)
        sgn = -1.0 if f & TF_NEG else 1.0
        X[:, :] = sgn * (X @ (Wm.T if kind == LK_TRSM_RLT else Wm))
    elif kind == LK_EXTEND_ADD:
        rel = P.relmap[(int(t["aux1"]) << 32 | (int(t["aux0"]) & 0xffffffff)):][:M].astype(np.int64)
        U = _view(ar[aa], t["a"], t["lda"], M, M)
        i, j = np.tril_indices(M)
        np.add.at(ar[ac], int(t["c"]) + rel[i] + rel[j] * int(t["ldc"]), U[i, j])
    elif kind == LK_GATHER_SYM:
        rel = P.relmap[(int(t["aux1"]) << 32 | (int(t["aux0"]) & 0xffffffff)):][:M].astype(np.int64)
        i, j = np.meshgrid(np.arange(M), np.arange(M), indexing="ij")
        a, b = np.maximum(rel[i], rel[j]), np.minimum(rel[i], rel[j])
        _view(ar[ac], t["c"], t["ldc"], M, M)[:, :] = ar[aa][int(t["a"]) + a + b * int(t["lda"])]
    elif kind == LK_SET_IDENTITY:
        _view(ar[ac], t["c"], t["ldc"], M, N)[:, :] = np.eye(M, N)
    elif kind == LK_SCALE:
        _view(ar[ac], t["c"], t["ldc"], M, N)[:, :] *= float(t["alpha"])
    elif kind == LK_DIAG_OUT:
        o = int(t["aux1"]) << 32 | (int(t["aux0"]) & 0xffffffff)
        ar[ZDIAG_ARENA][o:o + M] = np.diag(_view(ar[ac], t["c"], t["ldc"], M, M))
    elif kind == LK_ZERO_FRONT:
        Cv = _view(ar[ac], t["c"], t["ldc"], M, N)
        if f & TF_TRI:
            i, j = np.meshgrid(np.arange(M), np.arange(N), indexing="ij")
            Cv[(i // 64) >= (j // 64)] = 0.0
        else:
            Cv[:, :] = 0.0
    elif kind == LK_FRONT_FACTOR_SMALL:
        s = int(t["aux0"])
        d, sc, ld, fo = P.d(s), P.sc(s), int(P.ld[s]), int(P.foff[s])
        Fv = _view(ar[0], fo, ld, d, d)
        Fm = np.zeros((d, d))
        Fm[:, :sc] = np.tril(Fv)[:, :sc]                    # only the panel is staged; the rest starts from zero
        for c in P.child_idx[P.child_ptr[s]:P.child_ptr[s + 1]]:
            dc, scc = P.d(c), P.sc(c)
            rc = dc - scc
            if rc <= 0:
                continue
            U = np.tril(_view(ar[0], int(P.foff[c]) + scc * int(P.ld[c]) + scc, int(P.ld[c]), rc, rc))
            rel = P.relmap[int(P.rptr[c]) + scc:int(P.rptr[c]) + dc].astype(np.int64)
            Fm[np.ix_(rel, rel)] += U                       # rel is increasing: lower stays lower
        F11 = Fm[:sc, :sc] + np.tril(Fm[:sc, :sc], -1).T
        L11 = np.linalg.cholesky(F11)
        L21 = np.linalg.solve(L11, Fm[sc:, :sc].T).T
        U22 = np.tril(Fm[sc:, sc:]) - np.tril(L21 @ L21.T)
        out = np.zeros((d, d))
        out[:sc, :sc] = L11
        out[sc:, :sc] = L21
        out[sc:, sc:] = U22
        i, j = np.tril_indices(d)
        Fv[i, j] = out[i, j]
    elif kind == LK_FRONT_SELINV_SMALL:
        s = int(t["aux0"])
        d, sc, ld, fo = P.d(s), P.sc(s), int(P.ld[s]), int(P.foff[s])
        r = d - sc
        Lv = np.tril(_view(ar[0], fo, ld, d, d))
        L11, L21 = Lv[:sc, :sc], Lv[sc:, :sc]
        W = np.linalg.solve(L11, np.eye(sc))
        Z = np.zeros((d, d))
        if r > 0:
            p = int(P.sparent[s])
            rel = P.relmap[int(P.rptr[s]) + sc:int(P.rptr[s]) + d].astype(np.int64)
            Zp = _view(ar[1], int(P.foff[p]), int(P.ld[p]), P.d(p), P.d(p))
            Zrr = np.tril(Zp)[np.ix_(rel, rel)]
            Zrr = Zrr + np.tril(Zrr, -1).T
            Y = L21 @ W
            Zrc = -Zrr @ Y
            Z[sc:, sc:] = Zrr
            Z[sc:, :sc] = Zrc
            Z[:sc, :sc] = W.T @ W - Y.T @ Zrc
        else:
            Z[:sc, :sc] = W.T @ W
        Zv = _view(ar[1], fo, ld, d, d)
        i, j = np.tril_indices(d)
        Zv[i, j] = Z[i, j]
        ar[ZDIAG_ARENA][int(P.sptr[s]):int(P.sptr[s]) + sc] = np.diag(Z)[:sc]
    elif kind in (LK_MR_FWD_SMALL, LK_MR_BWD_SMALL, LK_MR_ASSEMBLE, LK_MR_GATHER):
        s = int(t["aux0"])
        nr, ldk = P.mr
        d, sc = P.d(s), P.sc(s)
        r = d - sc
        X = _view(ar[1], 0, ldk, nr, P.n)                     # node-major panel: X[q, node]
        U = _view(ar[2], 0, ldk, nr, max(P.urows, 1))
        c0, u0 = int(P.sptr[s]), int(P.uoff[s])
        below = P.rows[int(P.rptr[s]) + sc:int(P.rptr[s]) + d].astype(np.int64)
        kids = P.child_idx[P.child_ptr[s]:P.child_ptr[s + 1]]
        Lv = np.tril(_view(ar[0], P.foff[s], P.ld[s], d, d))[:, :sc] if kind in (LK_MR_FWD_SMALL, LK_MR_BWD_SMALL) else None

        def add_children(v):                                   # v: nr x d working panel of J
            for c in kids:
                rc = P.d(c) - P.sc(c)
                rel = P.relmap[int(P.rptr[c]) + P.sc(c):int(P.rptr[c]) + P.d(c)].astype(np.int64)
                v[:, rel] += U[:, int(P.uoff[c]):int(P.uoff[c]) + rc]

        if kind == LK_MR_FWD_SMALL:
            v = np.zeros((nr, d))
            v[:, :sc] = X[:, c0:c0 + sc]
            add_children(v)
            y = np.linalg.solve(Lv[:sc], v[:, :sc].T).T
            X[:, c0:c0 + sc] = y
            U[:, u0:u0 + r] = v[:, sc:] - y @ Lv[sc:].T
        elif kind == LK_MR_ASSEMBLE:
            v = np.zeros((nr, d))
            v[:, :sc] = X[:, c0:c0 + sc]
            add_children(v)
            X[:, c0:c0 + sc] = v[:, :sc]
            U[:, u0:u0 + r] = v[:, sc:]
        elif kind == LK_MR_GATHER:
            U[:, u0:u0 + r] = X[:, below]
        else:
            rhs = X[:, c0:c0 + sc] - X[:, below] @ Lv[sc:]
            X[:, c0:c0 + sc] = np.linalg.solve(Lv[:sc].T, rhs.T).T
    else:
        raise NotImplementedError(f"launch kind {kind}")


def run(P: Plans, S: State, plan):
    for L in plan["launches"]:
        for k in range(int(L["ntasks"])):
            _run_task(P, S, int(L["kind"]), plan["tasks"][int(L["task0"]) + k])


def scatter_values(P: Plans, S: State):
    """k_scatter_values: arena[amap[k]] = nzval[k] for the stored entries of the analysed triangle."""
    m = P.amap >= 0
    S.ar[0][P.amap[m]] = P.A.data[m]


def factor_matrix(P: Plans, S: State):
    """L (internal ordering) as a sparse matrix, read from the factor panels of the fronts."""
    rows, cols, vals = [], [], []
    for s in range(P.nsuper):
        d, sc = P.d(s), P.sc(s)
        Fv = _view(S.ar[0], P.foff[s], P.ld[s], d, sc)
        rr = P.rows[int(P.rptr[s]):int(P.rptr[s]) + d]
        for j in range(sc):
            rows.append(rr[j:])
            cols.append(np.full(d - j, int(P.sptr[s]) + j))
            vals.append(np.array(Fv[j:, j]))
    return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(P.n, P.n))


def selected_inverse_entries(P: Plans, S: State):
    """(rows, cols, values) of the selected inverse on the pattern of L (lower part, internal ordering)."""
    rows, cols, vals = [], [], []
    for s in range(P.nsuper):
        d, sc = P.d(s), P.sc(s)
        Zv = _view(S.ar[1], P.foff[s], P.ld[s], d, sc)
        rr = P.rows[int(P.rptr[s]):int(P.rptr[s]) + d]
        for j in range(sc):
            rows.append(rr[j:])
            cols.append(np.full(d - j, int(P.sptr[s]) + j))
            vals.append(np.array(Zv[j:, j]))
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def check_matrix(A, ordering="nd", coords=None, perm=None, use_wide=False, wide_min=128, verbose=False):
    """Hazard check + interpretation of the zero / factor / (wide inverse) / selected-inversion plans of `A`.
    Returns a dict of statistics and errors."""
    P = Plans(A, ordering=ordering, coords=coords, perm=perm, wide_min=wide_min)
    out = dict(n=P.n, nsuper=P.nsuper, n_wide=int(P.wide.size), max_front=max(P.d(s) for s in range(P.nsuper)), violations=[])
    names = ["zero", "factor"] + (["wide"] if use_wide and P.wide.size else []) + ["selinv"]
    for w in names:
        pl = P.plan(w, use_wide and w == "selinv" and P.wide.size > 0)
        nl, nt, bad = hazards(P, pl)
        out[w] = dict(launches=nl, tasks=nt)
        out["violations"] += [(w,) + b for b in bad]
    S = State(P, poison=True)
    run(P, S, P.plan("zero"))
    scatter_values(P, S)
    run(P, S, P.plan("factor"))
    Lm = factor_matrix(P, S)
    Ap = P.A[P.perm][:, P.perm]
    out["factor_err"] = float(abs(Lm @ Lm.T - Ap).max() / abs(Ap).max())
    if use_wide and P.wide.size:
        run(P, S, P.plan("wide"))
        werr = 0.0
        for i, s in enumerate(P.wide):
            sc = P.sc(s)
            Wm = np.tril(_view(S.ar[3], P.wide_off[i], P.wide_ld[i], sc, sc))
            L11 = np.tril(_view(S.ar[0], P.foff[s], P.ld[s], sc, sc))
            werr = max(werr, float(abs(Wm @ L11 - np.eye(sc)).max()))
        out["wide_err"] = werr
    run(P, S, P.plan("selinv", use_wide and P.wide.size > 0))
    r, c, v = selected_inverse_entries(P, S)
    Zd = np.linalg.inv(Ap.toarray())
    out["selinv_err"] = float(np.max(np.abs(v - Zd[r, c])) / np.max(np.abs(Zd)))
    out["zdiag_err"] = float(np.max(np.abs(S.ar[ZDIAG_ARENA] - np.diag(Zd))) / np.max(np.abs(Zd)))
    if verbose:
        print(out)
    return out


def check_hazards_only(A, ordering="nd", coords=None, perm=None, nr=50, wide_min=128):
    """Hazard check of every plan of `A` (zero, factor, wide inverses, selected inversion with the kept inverses, panel
    sweeps for nr right-hand sides) without interpreting them: for patterns whose dense inverse is out of reach."""
    import time
    P = Plans(A, ordering=ordering, coords=coords, perm=perm, wide_min=wide_min)
    out = dict(n=P.n, nsuper=P.nsuper, n_wide=int(P.wide.size), max_front=max(P.d(s) for s in range(P.nsuper)),
               arena_doubles=P.arena, violations=[])
    t0 = time.time()
    uw = P.wide.size > 0
    for w in ["zero", "factor"] + (["wide"] if uw else []) + ["selinv"]:
        nl, nt, bad = hazards(P, P.plan(w, uw and w == "selinv"))
        out[w] = dict(launches=nl, tasks=nt)
        out["violations"] += [(w,) + b for b in bad]
    fwd, bwd = P.mr_plans(nr, (nr + 1) & ~1)
    for name, pl in (("panel_fwd", fwd), ("panel_bwd", bwd)):
        nl, nt, bad = hazards(P, pl)
        out[name] = dict(launches=nl, tasks=nt)
        out["violations"] += [(name,) + b for b in bad]
    out["seconds"] = round(time.time() - t0, 1)
    return out


def check_panel_solves(A, nr=7, ordering="nd", coords=None, perm=None):
    """Hazard check + interpretation of the panel (multi-right-hand-side) sweeps: zero, scatter, factor, then forward
    and backward sweep of a node-major panel of nr right-hand sides; the result must solve (P A P') X' = B'."""
    P = Plans(A, ordering=ordering, coords=coords, perm=perm)
    ldk = (nr + 1) & ~1
    fwd, bwd = P.mr_plans(nr, ldk)
    out = dict(n=P.n, nsuper=P.nsuper, violations=[])
    for name, pl in (("fwd", fwd), ("bwd", bwd)):
        nl, nt, bad = hazards(P, pl)
        out[name] = dict(launches=nl, tasks=nt)
        out["violations"] += [(name,) + b for b in bad]
    S = State(P, poison=True)
    S.ar[4] = np.full(max(fwd["dinv"], P.plan("factor")["dinv"], DINV_SLOT), np.nan)
    run(P, S, P.plan("zero"))
    scatter_values(P, S)
    run(P, S, P.plan("factor"))
    rng = np.random.default_rng(0)
    Bm = rng.standard_normal((nr, P.n))
    S.ar[1] = np.full(ldk * P.n, np.nan)
    S.ar[2] = np.full(ldk * max(P.urows, 1), np.nan)
    _view(S.ar[1], 0, ldk, nr, P.n)[:, :] = Bm
    run(P, S, fwd)
    Ap = P.A[P.perm][:, P.perm].toarray()
    Lc = np.linalg.cholesky(Ap)
    Y = _view(S.ar[1], 0, ldk, nr, P.n)
    out["fwd_err"] = float(np.abs(Y - np.linalg.solve(Lc, Bm.T).T).max() / np.abs(Bm).max())
    run(P, S, bwd)
    Xs = _view(S.ar[1], 0, ldk, nr, P.n)
    ref = np.linalg.solve(Ap, Bm.T).T
    out["solve_err"] = float(np.abs(Xs - ref).max() / np.abs(ref).max())
    return out


# ------------------------------------------------- block-tridiagonal look-ahead schedule (csrc/btd.cu) ----
# btd_run_factor runs three plans per block on three streams (POTRF_i on the context's stream, TRSM_{i+1} and the rank-512
# SYRK_{i+1} on two further streams) ordered only by per-panel events.  Launches of different streams run concurrently,
# so the hazard to exclude is between LAUNCHES: any two launches that are not ordered by stream order + events must not
# touch the same entry unless both only read it.  The host issue sequence of btd_run_factor is replayed here (same
# event bookkeeping as cudaEventRecord / cudaStreamWaitEvent: a wait refers to the most recent record at the time of the
# call), giving a happens-before relation; the access sets come from the same Task records as above.
class BtdPlans:
    """The three look-ahead plans of a block-tridiagonal factor with blocks of order b (plan.cpp builders)."""

    def __init__(self, b):
        self.L = _lib()
        P = C.c_void_p
        self.L.pcb_create.restype = P
        self.L.pcb_create.argtypes = [C.c_int]
        self.L.pcb_destroy.argtypes = [P]
        self.L.pcb_sizes.argtypes = [P, P]
        self.L.pcb_get.argtypes = [P, C.c_int, P, P]
        h = self.L.pcb_create(int(b))
        sz = np.zeros(20, dtype=np.int64)
        self.L.pcb_sizes(h, _p(sz))
        self.b, self.ld, self.slot = int(b), int(sz[0]), int(sz[1])
        self.nblk = (self.b + 63) // 64
        self.bs = self.ld * self.b
        self.plans = []  # 0 POTRF_i, 1 TRSM_{i+1}, 2 SYRK_{i+1}; time-sharded lane: 3 W_i = L_i^-1, 4 / 5 spike first / step
        for w in range(6):
            nt, nl, dinv = (int(v) for v in sz[2 + 3 * w:5 + 3 * w])
            tasks = np.zeros(max(nt, 1), dtype=TASK)
            launches = np.zeros(max(nl, 1), dtype=LAUNCH)
            self.L.pcb_get(h, w, _p(tasks), _p(launches))
            self.plans.append(dict(tasks=tasks[:nt], launches=launches[:nl], dinv=dinv))
        self.L.pcb_destroy(h)
        self.setsz = self.nblk * DINV_SLOT


def _shift_task(BP, t, bases, dset=0):
    """the Task as the kernel sees it once Arenas holds `bases` (arena index -> start of that arena in one flat array that
    stands for all device buffers) and inverse set `dset`: absolute offsets, every operand in arena 0."""
    t = t.copy()
    f = int(t["flags"])
    t["a"] += bases[_arena(t, 0)]
    t["c"] += bases[_arena(t, 4)]
    t["b"] += dset * BP.setsz if f & TF_B_DINV else bases[_arena(t, 2)]
    t["flags"] = f & ~0x3f
    return t


LK_ZERO_UPPER = 100  # k_zero_upper of btd_winv_block (not a plan launch): strict upper triangle of a b x b block <- 0


def _node_access(nd):
    if nd["kind"] == LK_ZERO_UPPER:
        off, ld, b = nd["tasks"][0]
        i, j = np.triu_indices(b, 1)
        return [], [(0, off + i.astype(np.int64) + j.astype(np.int64) * ld)]
    R, Wr = [], []
    for t in nd["tasks"]:
        r, w = task_access(None, nd["kind"], t)
        R += r
        Wr += w
    return R, Wr


def _node_run(nd, S):
    if nd["kind"] == LK_ZERO_UPPER:
        off, ld, b = nd["tasks"][0]
        i, j = np.triu_indices(b, 1)
        S.ar[0][off + i.astype(np.int64) + j.astype(np.int64) * ld] = 0.0
        return
    for t in nd["tasks"]:
        _run_task(None, S, nd["kind"], t)


def btd_layout(BP, nblocks):
    """one flat array for every device buffer of the (time-sharded) factor: the factor slots, W (nblocks blocks), the
    scratch of the recursive doubling, the spike blocks S (nblocks blocks), the work arena [T | E_l | Q]."""
    o, lay = 0, {}
    for name, sz in (("arena", BP.slot * nblocks), ("W", BP.bs * nblocks), ("scratch", BP.bs), ("S", BP.bs * nblocks),
                     ("work", 3 * BP.bs)):
        lay[name] = o
        o += sz
    lay["total"] = o
    return lay


def btd_schedule(BP, nblocks, drop_trsm_waits=False, drop_potrf_wait=False, syrk_wait_shift=0, lane=False,
                 drop_lane_wait=False):
    """Replay of the host side of btd_run_factor (look-ahead branch).  Returns the launches in host issue order as dicts
    (stream, block, which, kind, tasks (absolute offsets), preds) plus the issues found while replaying (a wait on an
    event that was never recorded).  The keyword arguments remove / misplace synchronisation (negative controls)."""
    nodes, issues = [], []
    last = {0: None, 1: None, 2: None, 3: None}  # stream -> index of its last launch
    pending = {0: [], 1: [], 2: [], 3: []}       # stream -> launches its next launch must wait for
    events = {}
    lay = btd_layout(BP, nblocks)

    def record(ev, st):
        events[ev] = last[st]

    def wait(st, ev):
        if ev not in events:
            issues.append(("wait on an event never recorded", ev))
        elif events[ev] is not None:
            pending[st].append(events[ev])

    def push(st, block, which, li, kind, tasks, grid=1, flops=0.0):
        preds = ([last[st]] if last[st] is not None else []) + pending[st]
        pending[st] = []
        nodes.append(dict(stream=st, block=block, which=which, li=li, kind=kind, tasks=tasks, preds=preds, grid=grid,
                          flops=flops))
        last[st] = len(nodes) - 1

    def run_plan(which, st, block, dset, wait_evs, rec_evs):
        pl = BP.plans[which]
        slot_i = lay["arena"] + block * BP.slot
        bases = {0: [slot_i] * 4, 1: [slot_i] * 4, 2: [slot_i] * 4,
                 3: [slot_i, lay["W"] + block * BP.bs, lay["scratch"], 0],
                 4: [slot_i, lay["S"] + block * BP.bs, lay["work"], lay["W"] + block * BP.bs]}[min(which, 4)]
        for li, L in enumerate(pl["launches"]):
            wv = int(L["wait_ev"])
            if which == 2 and wv >= 0:
                wv = max(0, wv - syrk_wait_shift)
            if wv >= 0 and wait_evs is not None:
                wait(st, (wait_evs, wv))
            tasks = [_shift_task(BP, pl["tasks"][int(L["task0"]) + k], bases, dset) for k in range(int(L["ntasks"]))]
            push(st, block, which, li, int(L["kind"]), tasks, int(L["grid"]), float(L["flops"]))
            if int(L["rec_ev"]) >= 0 and rec_evs is not None:
                record((rec_evs, int(L["rec_ev"])), st)

    record("done", 0)
    wait(1, "done")
    wait(2, "done")
    for i in range(nblocks):
        s = i & 1
        if i > 0:
            run_plan(1, 1, i, 1 - s, None if drop_trsm_waits else ("ev", 1 - s), ("cev", s))
            run_plan(2, 2, i, 1 - s, ("cev", s), None)
            record("done", 2)
            if not drop_potrf_wait:
                wait(0, "done")
        run_plan(0, 0, i, s, None, ("ev", s))
        if lane:
            # gmrfb_btd_dist_factor's after_block: W_i = L_i^-1 and the spike step of block i on a further stream, ordered
            # behind POTRF_i by one event; the next blocks' chains go on meanwhile
            record("lane", 0)
            if not drop_lane_wait:
                wait(3, "lane")
            if BP.b > 1:
                push(3, i, 3, -1, LK_ZERO_UPPER, [(lay["arena"] + i * BP.slot, BP.ld, BP.b)])
            run_plan(3, 3, i, 0, None, None)
            run_plan(4 if i == 0 else 5, 3, i, 0, None, None)
    # the host synchronises the context's stream (then, with the lane, that stream too)
    return nodes, issues


def _happens_before(nodes):
    n = len(nodes)
    hb = np.zeros((n, n), dtype=bool)           # hb[k, j]: launch j completes before launch k starts
    for k, nd in enumerate(nodes):
        for p in nd["preds"]:
            hb[k] |= hb[p]
            hb[k, p] = True
    return hb


def btd_races(BP, nodes):
    """Launch pairs that touch a common entry (at least one writing) without being ordered.  One forward pass (each
    access against the last earlier writer of the entry) and one backward pass (each read against the next later writer)
    over the host order find every such pair up to transitivity."""
    hb = _happens_before(nodes)
    size = {0: 0, DINV_ARENA: 2 * BP.setsz}
    acc = []
    for nd in nodes:
        R, Wr = {}, {}
        r, w = _node_access(nd)
        for a, ix in r:
            R.setdefault(a, []).append(ix)
        for a, ix in w:
            Wr.setdefault(a, []).append(ix)
        R = {a: np.unique(np.concatenate(v)) for a, v in R.items()}
        Wr = {a: np.unique(np.concatenate(v)) for a, v in Wr.items()}
        for d in (R, Wr):
            for a, ix in d.items():
                size[a] = max(size.get(a, 0), int(ix.max()) + 1 if ix.size else 0)
        acc.append((R, Wr))
    bad = []
    name = {0: "POTRF", 1: "TRSM", 2: "SYRK", 3: "WINV", 4: "SPIKE", 5: "SPIKE"}

    def label(k):
        nd = nodes[k]
        return f"{name[nd['which']]}_{nd['block']}[{nd['li']}:{KIND_NAME.get(nd['kind'], 'ZERO_UPPER')}]"

    lastw = {a: np.full(s, -1, dtype=np.int64) for a, s in size.items()}
    for k, (R, Wr) in enumerate(acc):
        for what, d in (("read after write", R), ("write after write", Wr)):
            for a, ix in d.items():
                for w in np.unique(lastw[a][ix]):
                    if w >= 0 and w != k and not hb[k, w]:
                        bad.append((what, label(int(w)), label(k), a))
        for a, ix in Wr.items():
            lastw[a][ix] = k
    nextw = {a: np.full(s, -1, dtype=np.int64) for a, s in size.items()}
    for k in range(len(nodes) - 1, -1, -1):
        R, Wr = acc[k]
        for a, ix in R.items():
            for w in np.unique(nextw[a][ix]):
                if w >= 0 and w != k and not hb[w, k]:
                    bad.append(("write after read", label(k), label(int(w)), a))
        for a, ix in Wr.items():
            nextw[a][ix] = k
    # completion: the last launch of the context's stream (what the host waits for) is after every other launch
    sink = max(k for k, nd in enumerate(nodes) if nd["stream"] == 0)
    for k in range(len(nodes)):
        if k != sink and nodes[k]["stream"] != 3 and not hb[sink, k]:
            bad.append(("not ordered before the end of the context's stream", label(k), label(sink), -1))
    return bad, hb


def btd_simulate(nodes, t_potrf=22.7e-6, t_apply=8.8e-6, t_launch=5e-6, peak=33e12, sms=148, serial=False):
    """Fluid model of the schedule on `sms` SMs: a launch becomes ready when its predecessors have finished, asks for
    min(grid, sms) SMs and, given all of them, takes t_launch + flops / (peak * share of the machine) (GEMM kinds) or a
    fixed latency (64 x 64 POTRF + inverse; apply-inverse) - the per-launch figures of profiles/r02_launch_list.md.
    The launches of the context's stream (highest priority) are served first, the others share what is left in
    proportion to their demand.  Returns (makespan, length of the critical path with unlimited SMs, SM-seconds of work /
    sms).  serial=True: one stream, launches back to back (the schedule without look-ahead)."""
    n = len(nodes)
    dem = np.array([min(max(nd["grid"], 1), sms) for nd in nodes], dtype=float)
    ideal = np.empty(n)
    for k, nd in enumerate(nodes):
        if nd["kind"] == LK_POTRF:
            ideal[k] = t_potrf
        elif nd["kind"] in (LK_TRSM_RLT, LK_TRSM_RLN, LK_ZERO_UPPER, LK_SET_IDENTITY):
            ideal[k] = t_apply
        else:
            ideal[k] = t_launch + nd["flops"] / (peak * dem[k] / sms)
    work = float((ideal * dem).sum() / sms)
    if serial:
        return float(ideal.sum()), float(ideal.sum()), work
    preds = [sorted(set(nd["preds"])) for nd in nodes]
    cp = np.zeros(n)
    for k in range(n):
        cp[k] = ideal[k] + max([cp[p] for p in preds[k]], default=0.0)
    succ = [[] for _ in range(n)]
    left = np.array([len(p) for p in preds])
    for k, pp in enumerate(preds):
        for p in pp:
            succ[p].append(k)
    remaining = ideal.copy()          # seconds of work at full demand
    running = [k for k in range(n) if left[k] == 0]
    t = 0.0
    while running:
        free = float(sms)
        rate = {}
        hi = [k for k in running if nodes[k]["stream"] == 0]
        lo = [k for k in running if nodes[k]["stream"] != 0]
        for k in hi:
            a = min(dem[k], free)
            free -= a
            rate[k] = a / dem[k]
        tot = sum(dem[k] for k in lo)
        for k in lo:
            rate[k] = min(1.0, free / tot) if tot > 0 else 0.0
        dt = min(remaining[k] / rate[k] for k in running if rate[k] > 0)
        t += dt
        done = []
        for k in running:
            remaining[k] -= dt * rate[k]
            if remaining[k] <= 1e-15:
                done.append(k)
        for k in done:
            running.remove(k)
            for q in succ[k]:
                left[q] -= 1
                if left[q] == 0:
                    running.append(q)
    return t, float(cp.max()), work


class _BtdState:
    def __init__(self, arena, dinv):
        self.ar = [arena, None, None, None, dinv, None]


def check_btd_lookahead(b, nblocks=4, seed=0, orders=3, lane=False, **mutations):
    """Hazard check of the look-ahead schedule of a b x b, nblocks-block factor + interpretation of the launches in host
    order and in `orders` random topological orders of the happens-before relation, on arenas whose unused parts are
    NaN: every order must give the block Cholesky factor of src/tridiagonal_cholesky.jl:70-80."""
    BP = BtdPlans(b)
    out = dict(b=b, ld=BP.ld, nblocks=nblocks, violations=[])
    for w, nm in enumerate(("potrf", "trsm", "syrk") + (("winv", "spike_first", "spike_step") if lane else ())):
        nl, nt, bad = hazards(None, BP.plans[w])
        out[nm] = dict(launches=nl, tasks=nt)
        out["violations"] += [(nm,) + x for x in bad]
    nodes, issues = btd_schedule(BP, nblocks, lane=lane, **mutations)
    lay = btd_layout(BP, nblocks)
    out["violations"] += issues
    races, hb = btd_races(BP, nodes)
    out["violations"] += races
    out["launches"] = len(nodes)
    # pairs of launches that may overlap on the GPU (neither ordered before the other)
    free = ~(hb | hb.T)
    np.fill_diagonal(free, False)
    out["concurrent_pairs"] = int(free.sum() // 2)
    # interpretation
    rng = np.random.default_rng(seed)
    ld, slot = BP.ld, BP.slot
    D, Bs = [], []
    for i in range(nblocks):
        G = rng.standard_normal((b, b))
        D.append(G @ G.T / b + 4.0 * np.eye(b))
        Bs.append(rng.standard_normal((b, b)) / np.sqrt(b))
    Lr, Cr = [np.linalg.cholesky(D[0])], [None]
    for i in range(1, nblocks):
        Ci = np.linalg.solve(Lr[i - 1], Bs[i].T).T
        Cr.append(Ci)
        Lr.append(np.linalg.cholesky(D[i] - Ci @ Ci.T))
    # spike recurrence of the time-sharded factor: S_1 = E_l' W_1', S_i = -S_{i-1} C_i' W_i', Q = sum S_i S_i'
    El = Bs[0]
    Wr_ = [np.linalg.inv(Lm) for Lm in Lr]
    Sr = [El.T @ Wr_[0].T]
    for i in range(1, nblocks):
        Sr.append(-(Sr[i - 1] @ Cr[i].T) @ Wr_[i].T)
    Qr = sum(Sm @ Sm.T for Sm in Sr)
    errs = []
    n = len(nodes)
    for o in range(orders + 1):
        if o == 0:
            order = list(range(n))
        else:  # random topological order: repeatedly pick any launch whose predecessors have all run
            indeg = np.array([len(set(nd["preds"])) for nd in nodes])
            succ = [[] for _ in range(n)]
            for k, nd in enumerate(nodes):
                for p in set(nd["preds"]):
                    succ[p].append(k)
            ready = [k for k in range(n) if indeg[k] == 0]
            order = []
            while ready:
                k = ready.pop(int(rng.integers(len(ready))))
                order.append(k)
                for q in succ[k]:
                    indeg[q] -= 1
                    if indeg[q] == 0:
                        ready.append(q)
            assert len(order) == n
        arena = np.full(lay["total"], np.nan)
        if lane:  # what gmrfb_btd_dist_factor queues on the context's stream before the factor: W <- 0, Q <- 0, E_l
            arena[lay["W"]:lay["W"] + BP.bs * nblocks] = 0.0
            arena[lay["work"] + 2 * BP.bs:lay["work"] + 3 * BP.bs] = 0.0
            _view(arena, lay["work"] + BP.bs, ld, b, b)[:, :] = El
        for i in range(nblocks):
            Dv = _view(arena, i * slot, ld, b, b)
            il, jl = np.tril_indices(b)
            Dv[il, jl] = D[i][il, jl]                    # only the lower triangle is supplied
            if i > 0:
                _view(arena, i * slot + ld * b, ld, b, b)[:, :] = Bs[i]
        S = _BtdState(arena, np.full(2 * BP.setsz, np.nan))
        for k in order:
            _node_run(nodes[k], S)
        e = 0.0
        for i in range(nblocks):
            e = max(e, float(np.abs(np.tril(_view(arena, i * slot, ld, b, b)) - Lr[i]).max()))
            if i > 0:
                e = max(e, float(np.abs(_view(arena, i * slot + ld * b, ld, b, b) - Cr[i]).max()))
            if lane:
                e = max(e, float(np.abs(_view(arena, lay["W"] + i * BP.bs, ld, b, b) - Wr_[i]).max() / np.abs(Wr_[i]).max()))
                e = max(e, float(np.abs(_view(arena, lay["S"] + i * BP.bs, ld, b, b) - Sr[i]).max() / np.abs(Sr[i]).max()))
        if lane:
            Qv = np.tril(_view(arena, lay["work"] + 2 * BP.bs, ld, b, b))
            e = max(e, float(np.abs(Qv - np.tril(Qr)).max() / np.abs(Qr).max()))
        errs.append(e)
    out["factor_err"] = errs
    return out


def main():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    W = g.load_pkg().workloads
    for nx, ordering, uw in ((12, "nd", False), (40, "nd", False), (58, "nd", True), (45, "amd", False)):
        prob = W.matern_posterior(nx, obs_frac=0.2, q_eps=1e2, corr_range=0.2, seed=nx)
        res = check_matrix(prob["Qpost"], ordering=ordering, coords=prob["nodes"] if ordering == "nd" else None,
                           use_wide=uw, wide_min=65)
        print(nx, ordering, {k: v for k, v in res.items() if k != "violations"}, "violations:", res["violations"][:3])
    for b, nb, lane in ((320, 4, False), (700, 4, True), (1100, 3, False)):
        res = check_btd_lookahead(b, nblocks=nb, orders=1, lane=lane)
        print("btd look-ahead", {k: v for k, v in res.items() if k != "violations"}, "violations:", res["violations"][:3])


if __name__ == "__main__":
    main()
