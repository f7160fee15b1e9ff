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


def main():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    W = g.load_pkg().workloads
    for nx, ordering, uw in ((12, "nd", False), (40, "nd", False), (58, "nd", True), (45, "amd", False)):
        prob = W.matern_posterior(nx, obs_frac=0.2, q_eps=1e2, corr_range=0.2, seed=nx)
        res = check_matrix(prob["Qpost"], ordering=ordering, coords=prob["nodes"] if ordering == "nd" else None,
                           use_wide=uw, wide_min=65)
        print(nx, ordering, {k: v for k, v in res.items() if k != "violations"}, "violations:", res["violations"][:3])


if __name__ == "__main__":
    main()
