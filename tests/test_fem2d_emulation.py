"""CPU tier: the kernels of the Lagrange-triangle assembler and of the sparse-product / posterior-precision plans, run
thread by thread on the host (tools/probe/fem2d_emul.cpp compiles the kernel headers as plain C++ with the CUDA
qualifiers defined away) against the oracle.  No GPU and no libgmrfb call involved: this pins the kernel arithmetic, the
reference tables, the contribution lists and the coefficient lookup on a box without a GPU; the GPU tier
(tests/test_gpu_fem2d.py) runs the same cases through the C ABI."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    spec = importlib.util.spec_from_file_location("fem2d_emul", os.path.join(ROOT, "tools", "probe", "fem2d_emul.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # builds /tmp/libfem2d_emul.so with g++
    return mod


def rel(A, B):
    return abs(A - B).max() / abs(B).max()


@pytest.mark.parametrize("order,nx,degree,curve", [(1, 9, 0, 0.0), (2, 9, 0, 0.0), (2, 12, 4, 0.06), (2, 7, 2, 0.05)])
def test_kernels_thread_by_thread(emul, order, nx, degree, curve):
    fo, W = emul.fo, emul.W
    nodes, elems = emul.mesh(nx, order, curve)
    deg = degree or order + 1
    bnd = emul.boundary(nodes)
    E = emul.Emul(nodes, elems, order, degree)
    # unit stiffness, load, masses
    Gref, fref = fo.assemble_darcy_lagrange(nodes, elems, order, beta=2.5, degree=deg)
    G, f = E.stiffness(beta=2.5)
    assert np.array_equal(G.indptr, Gref.indptr) and np.array_equal(G.indices, Gref.indices)
    assert rel(G, Gref) < 1e-12 and np.abs(f - fref).max() < 1e-12 * np.abs(fref).max()
    assert rel(E.mass(0)[0], fo.assemble_mass_lagrange(nodes, elems, order, 0, degree=deg)) < 1e-12
    for kind in (1, 2):
        mref = fo.assemble_mass_lagrange(nodes, elems, order, kind, degree=deg)
        assert np.abs(E.mass(kind)[1] - mref).max() < 1e-12 * np.abs(mref).max()
    # coefficient per quadrature point, identity rows
    xc, yc = np.linspace(0, 1, 17), np.linspace(0, 1, 9)
    cm = np.random.default_rng(nx).uniform(1, 5, size=(xc.size, yc.size))
    Gref, fref = fo.assemble_darcy_lagrange(nodes, elems, order, xc, yc, cm, prescribed=bnd, degree=deg)
    E.set_grid(xc, yc)
    G, f = E.stiffness(np.ascontiguousarray(cm.T), presc=bnd)
    assert abs(G - Gref).max() < 1e-12 * abs(Gref).max() and np.all(f[bnd] == 0)
    # cubic tangent
    u = np.random.default_rng(1).standard_normal(nodes.shape[0])
    Jref, fref = fo.assemble_cubic_lagrange(nodes, elems, order, u, bnd, degree=deg, stiffness_scale=0.7)
    J, f = E.cubic(u, 0.7, bnd)
    assert abs(J - Jref).max() < 1e-12 * abs(Jref).max() and np.abs(f - fref).max() < 1e-12 * np.abs(fref).max()
    # Matern powers (the third through the sparse-product kernel)
    for alpha in (2, 3):
        Qref = fo.matern_precision_lagrange(nodes, elems, order, 9.0, 0.37, alpha=alpha, degree=deg)
        Q = E.matern(9.0, 0.37, alpha, order=order) if degree == 0 else None
        if Q is not None:
            assert abs(Q - Qref).max() < 1e-11 * abs(Qref).max()


def test_plan_builders_do_not_depend_on_the_worker_threads(emul, monkeypatch):
    rng = np.random.default_rng(0)
    A = sp.random(300, 400, density=0.02, random_state=3, format="csc")
    Q = sp.random(400, 400, density=0.02, random_state=4, format="csc")
    Q = (Q + Q.T + sp.identity(400)).tocsc()
    w = rng.uniform(0.5, 2.0, 300)
    outs = [emul.postprec(Q, A, w, t) for t in (1, 2, 5)]
    for o in outs[1:]:
        assert np.array_equal(o.indptr, outs[0].indptr) and np.array_equal(o.indices, outs[0].indices)
        assert np.array_equal(o.data, outs[0].data)
    ref = (Q + A.T @ sp.diags(w) @ A).tocsc()
    assert rel(outs[0], ref) < 1e-14
    K = emul.fo.assemble_darcy_lagrange(*emul.W.quadratic_mesh(*emul.W.structured_mesh(75, 75, seed=2)), 2)[0]
    assert K.shape[0] >= 20000  # the threaded path of the product pattern
    pats = []
    for th in ("1", "4"):
        monkeypatch.setenv("GMRFB_HOST_THREADS", th)
        pats.append(emul.spgemm(K, K))
    assert np.array_equal(pats[0].indices, pats[1].indices) and np.array_equal(pats[0].data, pats[1].data)
    assert rel(pats[0], K @ K) < 1e-14
