"""Drive tools/probe/fem2d_emul.cpp (host emulation of the fem2d / spgemm kernels) against oracle/fem_oracle.py with the
same cases as tests/test_gpu_fem2d.py.  Usage: python tools/probe/fem2d_emul.py  (builds /tmp/libfem2d_emul.so)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg, orc = g.load_pkg(), g.load_oracle()
W, fo = pkg.workloads, orc.fem
so = "/tmp/libfem2d_emul.so"
subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I", os.path.join(ROOT, "diffeqgmrfs.jl_b200", "csrc"),
                       os.path.join(ROOT, "tools", "probe", "fem2d_emul.cpp"), "-o", so])
L = C.CDLL(so)
P = C.c_void_p
L.emul_create.restype = P
L.emul_create.argtypes = [C.c_int, C.c_int, C.c_int64, P, C.c_int64, P]
L.emul_nnz.restype = C.c_int64
L.emul_nnz.argtypes = [P]
L.emul_nq.argtypes = [P]
L.emul_pattern.argtypes = [P, P, P]
L.emul_set_grid.argtypes = [P, C.c_int64, P, C.c_int64, P]
L.emul_stiffness.argtypes = [P, P, P, C.c_double, P, P]
L.emul_mass.argtypes = [P, C.c_int, P, P]
L.emul_matern_k.argtypes = [P, P, P, C.c_double, C.c_double, C.c_double, P, P]
L.emul_cubic.argtypes = [P, P, C.c_double, P, P, P]
L.emul_spgemm.restype = C.c_int64
L.emul_spgemm.argtypes = [C.c_int64, C.c_int64, C.c_int64, P, P, P, P, P, P, P, C.c_double, P, P, P]
L.emul_destroy.argtypes = [P]


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Emul:
    def __init__(self, nodes, elems, order, degree=0):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.elems = np.ascontiguousarray(elems, dtype=np.int64)
        self.n = self.nodes.shape[0]
        self.h = L.emul_create(order, degree, self.n, ptr(self.nodes), self.elems.shape[0], ptr(self.elems))
        assert self.h
        self.nnz = L.emul_nnz(self.h)
        self.colptr = np.empty(self.n + 1, dtype=np.int64)
        self.rowval = np.empty(self.nnz, dtype=np.int64)
        L.emul_pattern(self.h, ptr(self.colptr), ptr(self.rowval))

    def mat(self, vals):
        return sp.csc_matrix((vals, self.rowval, self.colptr), shape=(self.n, self.n))

    def set_grid(self, xc, yc):
        xc, yc = np.ascontiguousarray(xc, dtype=np.float64), np.ascontiguousarray(yc, dtype=np.float64)
        L.emul_set_grid(self.h, xc.size, ptr(xc), yc.size, ptr(yc))

    def stiffness(self, coeff=None, presc=None, beta=1.0):
        out, f = np.empty(self.nnz), np.empty(self.n)
        c = None if coeff is None else np.ascontiguousarray(coeff, dtype=np.float64)
        p = None if presc is None else np.ascontiguousarray(presc, dtype=np.uint8)
        L.emul_stiffness(self.h, ptr(c), ptr(p), beta, ptr(out), ptr(f))
        return self.mat(out), f

    def mass(self, lumping):
        out, ml = np.empty(self.nnz), np.empty(self.n)
        L.emul_mass(self.h, lumping, ptr(out), ptr(ml))
        return self.mat(out), ml

    def matern(self, kappa, ratio, alpha, presc=None, pm=1e-2, order=2):
        _, ml = self.mass(1 if order == 1 else 2)
        kv, w = np.empty(self.nnz), np.empty(self.n)
        p = None if presc is None else np.ascontiguousarray(presc, dtype=np.uint8)
        L.emul_matern_k(self.h, ptr(ml), ptr(p), pm, kappa**2, ratio if alpha == 2 else 1.0, ptr(kv), ptr(w))
        K = self.mat(kv)
        Q2 = (K.T @ sp.diags(w) @ K).tocsc()          # the postprec plan's job (tested on the GPU since round 1)
        if alpha == 2:
            return Q2
        Q2.sort_indices()
        return spgemm(Q2, K, w, ratio)

    def cubic(self, u, s, presc=None):
        Jv, f = np.empty(self.nnz), np.empty(self.n)
        p = None if presc is None else np.ascontiguousarray(presc, dtype=np.uint8)
        u = np.ascontiguousarray(u, dtype=np.float64)
        L.emul_cubic(self.h, ptr(u), s, ptr(p), ptr(Jv), ptr(f))
        return self.mat(Jv), f


def spgemm(A, Bm, w=None, alpha=1.0):
    A, Bm = A.tocsc(), Bm.tocsc()
    A.sort_indices(), Bm.sort_indices()
    ac, ar, av = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float64)
    bc, br, bv = Bm.indptr.astype(np.int64), Bm.indices.astype(np.int32), Bm.data.astype(np.float64)
    wv = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
    m, k, n = A.shape[0], A.shape[1], Bm.shape[1]
    nnz = L.emul_spgemm(m, k, n, ptr(ac), ptr(ar), ptr(av), ptr(bc), ptr(br), ptr(bv), ptr(wv), alpha, None, None, None)
    cp, cr, cv = np.empty(n + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32), np.empty(nnz)
    L.emul_spgemm(m, k, n, ptr(ac), ptr(ar), ptr(av), ptr(bc), ptr(br), ptr(bv), ptr(wv), alpha, ptr(cp), ptr(cr), ptr(cv))
    return sp.csc_matrix((cv, cr, cp), shape=(m, n))


def relmat(A, B):
    return abs(A - B).max() / abs(B).max()


def mesh(nx, order, curve=0.0, seed=0):
    nodes, tris = W.structured_mesh(nx, nx, seed=seed + nx)
    return (nodes, tris) if order == 1 else W.quadratic_mesh(nodes, tris, curve=curve, seed=seed)


def boundary(nodes):
    x, y = nodes[:, 0], nodes[:, 1]
    return (x == 0) | (x == 1) | (y == 0) | (y == 1)


# posterior-precision plan: same result for any number of builder threads, equal to Q + A' W A
L.emul_postprec.restype = C.c_int64
L.emul_postprec.argtypes = [C.c_int64, C.c_int64, P, P, P, P, P, P, C.c_int64, P, C.c_int, P, P, P]


def postprec(Q, A, w, threads):
    Q, A = Q.tocsc(), A.tocsc()
    Q.sort_indices(), A.sort_indices()
    qc, qr, qv = Q.indptr.astype(np.int64), Q.indices.astype(np.int32), Q.data.astype(np.float64)
    ac, ar, av = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float64)
    n, m = Q.shape[0], A.shape[0]
    wv = np.ascontiguousarray(w, dtype=np.float64)
    nnz = L.emul_postprec(n, m, ptr(qc), ptr(qr), ptr(qv), ptr(ac), ptr(ar), ptr(av), A.nnz, ptr(wv), threads, None, None, None)
    cp, cr, cv = np.empty(n + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32), np.empty(nnz)
    L.emul_postprec(n, m, ptr(qc), ptr(qr), ptr(qv), ptr(ac), ptr(ar), ptr(av), A.nnz, ptr(wv), threads, ptr(cp), ptr(cr), ptr(cv))
    return sp.csc_matrix((cv, cr, cp), shape=(n, n))



def main():
    worst = 0.0
    for order, nx, degree, curve in [(1, 5, 0, 0.0), (1, 40, 2, 0.0), (2, 4, 0, 0.0), (2, 33, 0, 0.0), (2, 33, 4, 0.06),
                                     (2, 60, 3, 0.0), (1, 33, 4, 0.0), (2, 9, 2, 0.05)]:
        nodes, elems = mesh(nx, order, curve)
        deg = degree or order + 1
        E = Emul(nodes, elems, order, degree)
        Gref, fref = fo.assemble_darcy_lagrange(nodes, elems, order, beta=2.5, degree=deg)
        G, f = E.stiffness(beta=2.5)
        assert np.array_equal(G.indptr, Gref.indptr) and np.array_equal(G.indices, Gref.indices), "pattern"
        e = [relmat(G, Gref), abs(f - fref).max() / abs(fref).max()]
        e.append(relmat(E.mass(0)[0], fo.assemble_mass_lagrange(nodes, elems, order, 0, degree=deg)))
        for kind in (1, 2):
            mref = fo.assemble_mass_lagrange(nodes, elems, order, kind, degree=deg)
            Ml, ml = E.mass(kind)
            e.append(abs(ml - mref).max() / abs(mref).max())
            assert abs(Ml - sp.diags(ml)).max() == 0
        print("unit", order, nx, deg, curve, ["%.1e" % v for v in e])
        worst = max(worst, max(e))

    for order, nx, seed in [(2, 21, 0), (2, 61, 3), (1, 61, 1)]:
        nodes, elems = mesh(nx, order)
        cg = W.darcy_problem(nx=9, seed=seed)["coeff_grid"]
        xc = yc = np.linspace(0, 1, 241)
        bnd = boundary(nodes)
        Gref, fref = fo.assemble_darcy_lagrange(nodes, elems, order, xc, yc, cg.T, beta=1.0, prescribed=bnd)
        E = Emul(nodes, elems, order)
        E.set_grid(xc, yc)
        G, f = E.stiffness(cg, presc=bnd)            # row-major (gy, gx) = entry ix + iy gx
        e = [abs(G - Gref).max() / abs(Gref).max(), abs(f - fref).max() / abs(fref).max()]
        print("darcy", order, nx, ["%.1e" % v for v in e])
        worst = max(worst, max(e))

    nodes, tris = W.structured_mesh(9, 9, jitter=0.0)
    n6, e6 = W.quadratic_mesh(nodes, tris)
    rng = np.random.default_rng(2)
    for xc, yc in ((np.linspace(0, 1, 13), np.linspace(0, 1, 25)), (np.linspace(1, 0, 17), np.linspace(0, 1, 6))):
        cm = rng.uniform(1, 5, size=(xc.size, yc.size))
        Gref, _ = fo.assemble_darcy_lagrange(n6, e6, 2, xc, yc, cm, degree=2)
        E = Emul(n6, e6, 2, 2)
        E.set_grid(xc, yc)
        G, _ = E.stiffness(np.ascontiguousarray(cm.T))
        e = abs(G - Gref).max() / abs(Gref).max()
        print("ties", "%.1e" % e)
        worst = max(worst, e)

    for order, nx, scale, with_bc, curve in [(2, 6, 1.0, True, 0.0), (2, 40, 0.0, False, 0.0), (2, 40, 2.5, True, 0.05),
                                             (1, 40, 1.0, True, 0.0)]:
        nodes, elems = mesh(nx, order, curve)
        u = np.random.default_rng(nx).standard_normal(nodes.shape[0])
        bnd = boundary(nodes) if with_bc else None
        Jref, fref = fo.assemble_cubic_lagrange(nodes, elems, order, u, bnd, stiffness_scale=scale)
        J, f = Emul(nodes, elems, order).cubic(u, scale, bnd)
        e = [abs(J - Jref).max() / abs(Jref).max(), abs(f - fref).max() / abs(fref).max()]
        print("cubic", order, nx, ["%.1e" % v for v in e])
        worst = max(worst, max(e))

    for order, nx, alpha, with_bc in [(2, 7, 2, False), (2, 25, 3, False), (1, 30, 3, False), (2, 25, 2, True), (2, 12, 3, True)]:
        nodes, elems = mesh(nx, order)
        kappa, ratio = np.sqrt(8.0) / 0.2, 0.37
        bnd = boundary(nodes) if with_bc else None
        Qref = fo.matern_precision_lagrange(nodes, elems, order, kappa, ratio, alpha=alpha, prescribed=bnd)
        Q = Emul(nodes, elems, order).matern(kappa, ratio, alpha, bnd, order=order)
        e = abs(Q - Qref).max() / abs(Qref).max()
        print("matern", order, nx, alpha, with_bc, "%.1e" % e)
        worst = max(worst, e)

    A = sp.random(70, 50, density=0.08, random_state=1, format="csc")
    Bm = sp.random(50, 90, density=0.1, random_state=2, format="csc")
    w = rng.uniform(0.5, 2.0, 50)
    Cm = spgemm(A, Bm, w, -1.5)
    Cref = (A @ sp.diags(w) @ Bm).tocsc()
    S = ((abs(A) > 0).astype(np.float64) @ (abs(Bm) > 0).astype(np.float64)).tocsc()
    S.sort_indices()
    assert np.array_equal(Cm.indptr, S.indptr) and np.array_equal(Cm.indices, S.indices)
    e = abs(Cm + 1.5 * Cref).max() / abs(Cref).max()
    print("spgemm", "%.1e" % e, "empty:", spgemm(A, sp.csc_matrix((50, 90))).nnz)
    worst = max(worst, e)
    for (mq, nq_, dens) in ((60, 40, 0.1), (500, 700, 0.01), (1, 1, 1.0)):
        Aobs = sp.random(mq, nq_, density=dens, random_state=3, format="csc")
        Qp = sp.random(nq_, nq_, density=dens, random_state=4, format="csc")
        Qp = (Qp + Qp.T + sp.identity(nq_)).tocsc()
        wobs = rng.uniform(0.5, 2.0, mq)
        ref = (Qp + Aobs.T @ sp.diags(wobs) @ Aobs).tocsc()
        outs = [postprec(Qp, Aobs, wobs, t) for t in (1, 2, 3, 7)]
        for o in outs[1:]:
            assert np.array_equal(o.indptr, outs[0].indptr) and np.array_equal(o.indices, outs[0].indices)
            assert np.array_equal(o.data, outs[0].data)
        e = abs(outs[0] - ref).max() / abs(ref).max()
        print("postprec", mq, nq_, "%.1e" % e, "nnz", outs[0].nnz)
        worst = max(worst, e)
    nodes_, tris_ = W.structured_mesh(40, 40, seed=1)
    n6_, e6_ = W.quadratic_mesh(nodes_, tris_)
    G_, _ = fo.assemble_darcy_lagrange(n6_, e6_, 2)
    Q_ = fo.matern_precision_lagrange(n6_, e6_, 2, 5.0, 0.1, alpha=2)
    w_ = rng.uniform(0.5, 2.0, n6_.shape[0])
    o1, o5 = postprec(Q_, G_, w_, 1), postprec(Q_, G_, w_, 5)
    assert np.array_equal(o1.indices, o5.indices) and np.array_equal(o1.data, o5.data)
    ref = (Q_ + G_.T @ sp.diags(w_) @ G_).tocsc()
    e = abs(o1 - ref).max() / abs(ref).max()
    print("postprec P2 mesh", "%.1e" % e)
    worst = max(worst, e)
    # sparse-product pattern: the same for any number of builder threads (n >= 20000 takes the threaded path)
    Kbig = fo.assemble_darcy_lagrange(*W.quadratic_mesh(*W.structured_mesh(80, 80, seed=2)), 2)[0]
    pats = []
    for th in ("1", "3", "8"):
        os.environ["GMRFB_HOST_THREADS"] = th
        Cb = spgemm(Kbig, Kbig)
        pats.append(Cb)
    os.environ.pop("GMRFB_HOST_THREADS")
    for o in pats[1:]:
        assert np.array_equal(o.indptr, pats[0].indptr) and np.array_equal(o.indices, pats[0].indices) and np.array_equal(o.data, pats[0].data)
    e = abs(pats[0] - Kbig @ Kbig).max() / abs(Kbig @ Kbig).max()
    print("spgemm threaded pattern", Kbig.shape[0], "%.1e" % e)
    worst = max(worst, e)
    print("worst relative difference", "%.2e" % worst)
    assert worst < 1e-12
    print("OK")


if __name__ == "__main__":
    main()
