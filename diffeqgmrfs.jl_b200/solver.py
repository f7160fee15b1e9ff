"""Host-side mirror of the reference's solver/blueprint interface on top of libgmrfb (ctypes).

The reference selects its linear-algebra backend by passing GaussianMarkovRandomFields.jl *blueprints*
(``CholeskySolverBlueprint(var_strategy=RBMCStrategy(50), perm=p)``, ``GNCholeskySolverBlueprint(p)``) to
``condition_on_observations`` / ``GaussNewtonOptimizer`` and then calls ``mean``, ``std``, ``rand``,
``sqmahal`` on the resulting GMRF (e.g. scripts/darcy/solve_darcy_gmrf-fem.jl:100,165-192).
Julia is not installed in this image, so the host side above the C ABI is written in Python with the same
names, argument meaning and error behaviour; ``julia/GMRFB200.jl`` holds the equivalent ``ccall`` shim.

Everything numeric happens in CUDA kernels behind ``include/gmrfb.h``; this module only marshals arrays.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import scipy.sparse as sp

from . import _lib as B


# ----------------------------------------------------------------------------------------------- context --
class Context:
    """A device + stream (gmrfb_ctx).  Raises if no sm_100-class GPU is visible — there is no CPU fallback."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        st = B.lib().gmrfb_ctx_create(device, C.byref(h))
        B.check(st, None)
        self.h = h
        self.device = device
        self._fin = weakref.finalize(self, B.lib().gmrfb_ctx_destroy, h)

    def sync(self):
        B.check(B.lib().gmrfb_ctx_sync(self.h), self.h)

    def profile_begin(self):
        B.check(B.lib().gmrfb_ctx_profile_begin(self.h), self.h)

    def profile_end(self):
        """-> list of dicts(name, kind, launches, ms, flops, bytes), one per kernel kind."""
        ent = (B.ProfileEntry * 64)()
        cnt = C.c_int32()
        B.check(B.lib().gmrfb_ctx_profile_end(self.h, ent, 64, C.byref(cnt)), self.h)
        return [dict(name=e.name.decode(), kind=e.kind, launches=e.launches, ms=e.ms, flops=e.flops, bytes=e.bytes)
                for e in ent[:cnt.value]]

    @property
    def stream(self) -> int:
        return int(B.lib().gmrfb_ctx_stream(self.h))

    @property
    def launch_count(self) -> int:
        return int(B.lib().gmrfb_ctx_launch_count(self.h))


def pool_trim(keep_bytes: int = 0) -> int:
    """Give cached device buffers back to the driver (gmrfb_pool_trim); returns the bytes still cached."""
    out = C.c_int64()
    B.check(B.lib().gmrfb_pool_trim(int(keep_bytes), C.byref(out)), None)
    return out.value


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _torch_ready(t):
    """A CUDA tensor about to be handed to the library: the library runs on its own non-blocking stream, which is not
    ordered against torch's, so whatever torch has queued to produce the tensor must have finished first."""
    import torch

    torch.cuda.current_stream(t.device).synchronize()


def _csc(A):
    A = sp.csc_matrix(A)
    if not A.has_sorted_indices:
        A = A.copy()
        A.sort_indices()
    A.sum_duplicates()
    return A


# --------------------------------------------------------------------------------------- sparse matrices --
class SparseMatrix:
    """Device-resident CSC matrix (gmrfb_spm)."""

    def __init__(self, A, ctx: Context | None = None, _handle=None, _owner=None):
        self.ctx = ctx or default_context()
        if _handle is not None:
            self.h = _handle
            self._owner = _owner
            return
        A = _csc(A)
        self.shape = A.shape
        _, cp = B.i64(A.indptr)
        _, ri = B.i64(A.indices)
        _, nz = B.f64(A.data)
        h = C.c_void_p()
        st = B.lib().gmrfb_spm_create(self.ctx.h, A.shape[0], A.shape[1], cp, ri, nz, 0, C.byref(h))
        B.check(st, self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_spm_destroy, h)

    def dims(self):
        m, n, z = C.c_int64(), C.c_int64(), C.c_int64()
        B.check(B.lib().gmrfb_spm_dims(self.h, C.byref(m), C.byref(n), C.byref(z)), self.ctx.h)
        return m.value, n.value, z.value

    def set_values(self, data):
        _, nz = B.f64(data)
        B.check(B.lib().gmrfb_spm_set_values(self.h, nz), self.ctx.h)

    def to_scipy(self):
        m, n, z = self.dims()
        cp = np.empty(n + 1, np.int64)
        ri = np.empty(z, np.int64)
        nz = np.empty(z, np.float64)
        B.check(B.lib().gmrfb_spm_get(self.h, 0, cp.ctypes.data_as(B._I64P), ri.ctypes.data_as(B._I64P),
                                      nz.ctypes.data_as(B._F64P)), self.ctx.h)
        return sp.csc_matrix((nz, ri, cp), shape=(m, n))

    def values_dev(self) -> int:
        return int(B.lib().gmrfb_spm_values_dev(self.h) or 0)

    def values_host(self):
        """The nonzero values only (the pattern of a fixed-pattern result is fetched once by its owner)."""
        _, _, z = self.dims()
        nz = np.empty(z, np.float64)
        B.check(B.lib().gmrfb_spm_get(self.h, 0, None, None, nz.ctypes.data_as(B._F64P)), self.ctx.h)
        return nz

    def matvec(self, x, trans=False, alpha=1.0, beta=0.0, y=None):
        m, n, _ = self.dims()
        x, xp = B.f64(x)
        out = np.zeros(n if trans else m) if y is None else np.array(y, dtype=np.float64)
        B.check(B.lib().gmrfb_spmv(self.h, int(trans), alpha, xp, beta, out.ctypes.data_as(B._F64P)), self.ctx.h)
        return out

    def sqmahal(self, v, mu=None):
        _, vp = B.f64(v)
        mup = None
        if mu is not None:
            mu, mup = B.f64(mu)
        out = C.c_double()
        B.check(B.lib().gmrfb_sqmahal(self.h, mup, vp, C.byref(out)), self.ctx.h)
        return out.value


class PosteriorPrecision:
    """Fixed-pattern plan for Q + A' diag(q_eps) A (gmrfb_postprec): symbolic once, numeric per call."""

    def __init__(self, Q: SparseMatrix, A: SparseMatrix):
        self.ctx = Q.ctx
        self.Q, self.A = Q, A
        h = C.c_void_p()
        B.check(B.lib().gmrfb_postprec_create(self.ctx.h, Q.h, A.h, C.byref(h)), self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_postprec_destroy, h)

    def compute(self, q_eps) -> SparseMatrix:
        out = C.c_void_p()
        if np.isscalar(q_eps):
            st = B.lib().gmrfb_postprec_compute(self.h, float(q_eps), None, C.byref(out))
        else:
            _, wp = B.f64(q_eps)
            st = B.lib().gmrfb_postprec_compute(self.h, 0.0, wp, C.byref(out))
        B.check(st, self.ctx.h)
        return SparseMatrix(None, ctx=self.ctx, _handle=out, _owner=self)


class SparseProduct:
    """Fixed-pattern plan for ``alpha * A * diag(w) * B`` (gmrfb_spgemm): symbolic once, numeric per call."""

    def __init__(self, A: SparseMatrix, Bm: SparseMatrix):
        self.ctx = A.ctx
        self.A, self.B = A, Bm
        h = C.c_void_p()
        B.check(B.lib().gmrfb_spgemm_create(self.ctx.h, A.h, Bm.h, C.byref(h)), self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_spgemm_destroy, h)

    def compute(self, alpha=1.0, w=None) -> SparseMatrix:
        out = C.c_void_p()
        wp = None
        if w is not None:
            self._w_keep, wp = B.f64(w)
        B.check(B.lib().gmrfb_spgemm_compute(self.h, float(alpha), wp, C.byref(out)), self.ctx.h)
        return SparseMatrix(None, ctx=self.ctx, _handle=out, _owner=self)


class FEMP1:
    """P1 finite-element assembly on the device (gmrfb_fem): the Darcy stiffness of a new coefficient field
    (``assemble_darcy_diff_matrix``, src/problems/darcy.jl:5-63, with the nearest-index coefficient lookup of
    src/datasets/darcy.jl:30-34) and the Matern prior ``ratio * K' Mt^-1 K`` (src/spdes/shallow_water.jl:177-194), on a
    pattern analysed once per mesh.  ``nodes``: n x 2, ``tris``: T x 3 (0-based)."""

    def __init__(self, nodes, tris, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self._nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        tris, tp = B.i64(np.ascontiguousarray(tris).reshape(-1))
        self.n = self._nodes.shape[0]
        h = C.c_void_p()
        B.check(B.lib().gmrfb_fem_create(self.ctx.h, self.n, self._nodes.ctypes.data_as(B._F64P), tris.size // 3, tp, 0,
                                         C.byref(h)), self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_fem_destroy, h)
        self._presc_keep = None

    @property
    def mass(self):
        """Lumped mass vector = load vector of f = 1."""
        out = np.empty(self.n)
        B.check(B.lib().gmrfb_fem_get_mass(self.h, out.ctypes.data_as(B._F64P)), self.ctx.h)
        return out

    def set_coeff_grid(self, x_coords, y_coords):
        x, xp = B.f64(x_coords)
        y, yp = B.f64(y_coords)
        B.check(B.lib().gmrfb_fem_set_coeff_grid(self.h, x.size, xp, y.size, yp), self.ctx.h)
        self._grid = (x.size, y.size)

    def _presc(self, prescribed):
        if prescribed is None:
            return None
        self._presc_keep = np.ascontiguousarray(np.asarray(prescribed) != 0, dtype=np.uint8)
        assert self._presc_keep.size == self.n
        return C.c_void_p(self._presc_keep.ctypes.data)

    def assemble(self, coeff_grid=None, prescribed=None) -> SparseMatrix:
        """Stiffness matrix of the coefficient grid ``coeff_grid[iy, ix]`` (NumPy, shape (gy, gx): element (iy, ix) is
        ``coeff_mat[ix, iy]`` of the reference's column-major array) or a CUDA float64 tensor of the same layout; rows
        of ``prescribed`` dofs become identity rows.  The returned device matrix is owned by this object (fixed
        pattern, values of the last call)."""
        cp = None
        if coeff_grid is not None:
            if hasattr(coeff_grid, "data_ptr"):
                assert coeff_grid.is_cuda and coeff_grid.is_contiguous() and str(coeff_grid.dtype) == "torch.float64"
                assert tuple(coeff_grid.shape) == (self._grid[1], self._grid[0])
                _torch_ready(coeff_grid)
                cp = C.c_void_p(coeff_grid.data_ptr())
                self._coeff_keep = coeff_grid
            else:
                self._coeff_keep = np.ascontiguousarray(coeff_grid, dtype=np.float64)
                assert self._coeff_keep.shape == (self._grid[1], self._grid[0])
                cp = C.c_void_p(self._coeff_keep.ctypes.data)
        out = C.c_void_p()
        B.check(B.lib().gmrfb_fem_assemble(self.h, cp, self._presc(prescribed), C.byref(out)), self.ctx.h)
        M = SparseMatrix(None, ctx=self.ctx, _handle=out, _owner=self)
        M.shape = (self.n, self.n)
        return M

    def matern_precision(self, kappa, ratio, prescribed=None, prescribed_mass=1e-2) -> SparseMatrix:
        """``ratio * K' Mt^-1 K`` with ``K = kappa^2 Mt + G`` (device matrix owned by this object)."""
        out = C.c_void_p()
        B.check(B.lib().gmrfb_fem_matern_precision(self.h, float(kappa), float(ratio), self._presc(prescribed),
                                                   float(prescribed_mass), C.byref(out)), self.ctx.h)
        M = SparseMatrix(None, ctx=self.ctx, _handle=out, _owner=self)
        M.shape = (self.n, self.n)
        return M


    def assemble_cubic(self, u, prescribed=None, quad_degree=2, stiffness_scale=1.0, out=None):
        """Gauss-Newton tangent and residual of ``-lap u + u^3`` at the iterate ``u`` (``f_and_J`` of
        _research/elliptic_chen24.jl:280-285 without the static load vector): returns ``(f, J)`` with
        ``J = s G + 3 int u^2 phi_i phi_j`` (device matrix owned by this object) and ``f = s G u + int u^3 phi_i``,
        rows of ``prescribed`` dofs skipped.  ``u`` / ``out``: NumPy arrays or CUDA float64 tensors."""
        up, _keep_u = _vec_ptr(u, self.n)
        if out is None:
            out = np.empty(self.n)
        fp, _keep_f = _vec_ptr(out, self.n)
        J = C.c_void_p()
        B.check(B.lib().gmrfb_fem_assemble_cubic(self.h, up, int(quad_degree), float(stiffness_scale),
                                                 self._presc(prescribed), C.byref(J), fp), self.ctx.h)
        M = SparseMatrix(None, ctx=self.ctx, _handle=J, _owner=self)
        M.shape = (self.n, self.n)
        return out, M


class FEMLagrange:
    """Lagrange triangles of order 1 or 2 on the device (gmrfb_fem2d) with the cell values Ferrite computes for them
    (isoparametric geometry, ``QuadratureRule{RefTriangle}(order + 1)``): the discretisations of src/utils.jl:20-38
    (Darcy, ``element_order=2``) and _research/elliptic_chen24.jl:118-122.  ``nodes``: n x 2; ``elems``: E x 3 or E x 6
    (0-based; six-node elements numbered as Ferrite's QuadraticTriangle)."""

    def __init__(self, nodes, elems, order=None, quad_degree=0, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self._nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        elems = np.ascontiguousarray(elems)
        self.order = int(order) if order is not None else {3: 1, 6: 2}[elems.shape[1]]
        assert elems.shape[1] == (3 if self.order == 1 else 6)
        el, ep = B.i64(elems.reshape(-1))
        self.n = self._nodes.shape[0]
        h = C.c_void_p()
        B.check(B.lib().gmrfb_fem2d_create(self.ctx.h, self.order, self.n, self._nodes.ctypes.data_as(B._F64P),
                                           elems.shape[0], ep, 0, int(quad_degree), C.byref(h)), self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_fem2d_destroy, h)
        self._presc_keep = None
        self._grid = None

    def _presc(self, prescribed):
        if prescribed is None:
            return None
        self._presc_keep = np.ascontiguousarray(np.asarray(prescribed) != 0, dtype=np.uint8)
        assert self._presc_keep.size == self.n
        return C.c_void_p(self._presc_keep.ctypes.data)

    def _wrap(self, h):
        M = SparseMatrix(None, ctx=self.ctx, _handle=h, _owner=self)
        M.shape = (self.n, self.n)
        return M

    @property
    def info(self):
        o, npe, nq, nnz = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        B.check(B.lib().gmrfb_fem2d_info(self.h, C.byref(o), C.byref(npe), C.byref(nq), C.byref(nnz)), self.ctx.h)
        return dict(order=o.value, nodes_per_element=npe.value, nquad=nq.value, nnz=nnz.value)

    def set_coeff_grid(self, x_coords, y_coords):
        x, xp = B.f64(x_coords)
        y, yp = B.f64(y_coords)
        B.check(B.lib().gmrfb_fem2d_set_coeff_grid(self.h, x.size, xp, y.size, yp), self.ctx.h)
        self._grid = (x.size, y.size)

    def stiffness(self, coeff_grid=None, prescribed=None, beta=1.0, load=True):
        """``assemble_darcy_diff_matrix`` (src/problems/darcy.jl:5-63): returns ``(G, f)``; ``coeff_grid[iy, ix]``
        (NumPy, shape (gy, gx), or a CUDA float64 tensor of that layout) is ``coeff_mat[ix, iy]`` of the reference,
        looked up at every quadrature point; rows of ``prescribed`` dofs become identity rows with a zero load."""
        cp = None
        if coeff_grid is not None:
            if hasattr(coeff_grid, "data_ptr"):
                assert coeff_grid.is_cuda and coeff_grid.is_contiguous() and str(coeff_grid.dtype) == "torch.float64"
                assert tuple(coeff_grid.shape) == (self._grid[1], self._grid[0])
                _torch_ready(coeff_grid)
                cp = C.c_void_p(coeff_grid.data_ptr())
                self._coeff_keep = coeff_grid
            else:
                self._coeff_keep = np.ascontiguousarray(coeff_grid, dtype=np.float64)
                assert self._coeff_keep.shape == (self._grid[1], self._grid[0])
                cp = C.c_void_p(self._coeff_keep.ctypes.data)
        f = np.empty(self.n) if load else None
        out = C.c_void_p()
        B.check(B.lib().gmrfb_fem2d_stiffness(self.h, cp, self._presc(prescribed), float(beta), C.byref(out),
                                              C.c_void_p(f.ctypes.data) if load else None), self.ctx.h)
        return self._wrap(out), f

    def mass(self, lumping=0):
        """Mass matrix (``lumping`` 0 consistent, 1 row sums, 2 scaled element diagonals, 3 = ``lump_matrix`` for this
        order); returns ``(M, lumped vector or None)``."""
        out = C.c_void_p()
        ml = np.empty(self.n) if lumping else None
        B.check(B.lib().gmrfb_fem2d_mass(self.h, int(lumping), C.byref(out), ml.ctypes.data_as(B._F64P) if lumping else None),
                self.ctx.h)
        return self._wrap(out), ml

    def matern_precision(self, kappa, ratio, alpha=2, prescribed=None, prescribed_mass=1e-2) -> SparseMatrix:
        """``ratio * K' Mt^-1 K`` (alpha 2) or ``ratio * K Mt^-1 K Mt^-1 K`` (alpha 3), ``K = kappa^2 Mt + G``
        (src/spdes/shallow_water.jl:172-190); device matrix owned by this object."""
        out = C.c_void_p()
        B.check(B.lib().gmrfb_fem2d_matern_precision(self.h, float(kappa), float(ratio), int(alpha), self._presc(prescribed),
                                                     float(prescribed_mass), C.byref(out)), self.ctx.h)
        return self._wrap(out)

    def assemble_cubic(self, u, prescribed=None, stiffness_scale=1.0, out=None):
        """``(f, J)`` of ``f_and_J`` (_research/elliptic_chen24.jl:280-285) without the static load vector, with the
        quadrature rule of the handle; see FEMP1.assemble_cubic."""
        up, _keep_u = _vec_ptr(u, self.n)
        if out is None:
            out = np.empty(self.n)
        fp, _keep_f = _vec_ptr(out, self.n)
        J = C.c_void_p()
        B.check(B.lib().gmrfb_fem2d_assemble_cubic(self.h, up, float(stiffness_scale), self._presc(prescribed), C.byref(J),
                                                   fp), self.ctx.h)
        return out, self._wrap(J)


def _vec_ptr(v, n):
    """(pointer, keep-alive) of a length-n float64 vector: NumPy (host) or a contiguous CUDA tensor (device)."""
    if hasattr(v, "data_ptr"):
        assert v.is_cuda and v.is_contiguous() and str(v.dtype) == "torch.float64" and v.numel() == n
        _torch_ready(v)
        return C.c_void_p(v.data_ptr()), v
    assert isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags.c_contiguous and v.size == n, \
        "expected a contiguous float64 vector of length %d" % n
    return C.c_void_p(v.ctypes.data), v


class FEM1D:
    """Lagrange line elements of order 1 or 2 on the device (gmrfb_fem1d) for the Burgers Gauss-Newton loop:
    ``assemble_burgers_mass_diffusion_matrices`` / ``assemble_burgers_advection_matrix`` (src/problems/burgers.jl:5-98)
    and the space-time ``f_and_J`` of scripts/burgers/solve_burgers_gmrf-fem.jl:115-142 in one kernel.
    ``elems``: E x (order + 1) node ids (0-based; quadratic: left, right, middle); ``x``: either the n node
    coordinates or, for a periodic mesh whose last element closes the ring, the E x (order + 1) coordinates of the
    element nodes."""

    def __init__(self, x, elems, order=1, nquad=0, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        elems = np.ascontiguousarray(elems)
        assert elems.ndim == 2 and elems.shape[1] == order + 1
        x = np.asarray(x, dtype=np.float64)
        xe = x if x.ndim == 2 else x[elems]
        assert xe.shape == elems.shape
        xe, xp = B.f64(np.ascontiguousarray(xe).reshape(-1))
        el, ep = B.i64(elems.reshape(-1))
        self.n, self.order = int(elems.max()) + 1, order
        h = C.c_void_p()
        B.check(B.lib().gmrfb_fem1d_create(self.ctx.h, self.n, elems.shape[0], ep, xp, int(order), 0, int(nquad),
                                           C.byref(h)), self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_fem1d_destroy, h)
        self._presc_keep = None

    @classmethod
    def periodic_unit_interval(cls, n_elems, order=2, **kw):
        """``periodic_unit_interval_discretization(N_x; element_order)`` (src/utils.jl:42-49) with the periodic
        constraint condensed into the connectivity: the last element closes the ring (no slave dof)."""
        from .workloads import periodic_line_mesh

        x, el = periodic_line_mesh(n_elems, order)
        return cls(x, el, order=order, **kw)

    def _presc(self, prescribed):
        if prescribed is None:
            return None
        self._presc_keep = np.ascontiguousarray(np.asarray(prescribed) != 0, dtype=np.uint8)
        assert self._presc_keep.size == self.n
        return C.c_void_p(self._presc_keep.ctypes.data)

    def _wrap(self, h, shape):
        M = SparseMatrix(None, ctx=self.ctx, _handle=h, _owner=self)
        M.shape = shape
        return M

    def mass_stiffness(self, lumping=False, prescribed=None):
        M, G = C.c_void_p(), C.c_void_p()
        B.check(B.lib().gmrfb_fem1d_mass_stiffness(self.h, int(bool(lumping)), self._presc(prescribed), C.byref(M),
                                                   C.byref(G)), self.ctx.h)
        return self._wrap(M, (self.n, self.n)), self._wrap(G, (self.n, self.n))

    def advection(self, u, prescribed=None, out=None):
        """``(G_adv, v)`` of ``assemble_burgers_advection_matrix(disc, u)``."""
        up, _ku = _vec_ptr(u, self.n)
        if out is None:
            out = np.empty(self.n)
        vp, _kv = _vec_ptr(out, self.n)
        A = C.c_void_p()
        B.check(B.lib().gmrfb_fem1d_advection(self.h, up, self._presc(prescribed), C.byref(A), vp), self.ctx.h)
        return self._wrap(A, (self.n, self.n)), out

    def spacetime_tangent(self, w, nt, dt, nu, prescribed=None, out=None):
        """``(f, J)`` of the script's ``f_and_J(w)``: ``w`` time-major with ``nt`` steps; J is a device matrix
        ((nt-1) n x nt n, fixed pattern, values of the last call) owned by this object."""
        wp, _kw = _vec_ptr(w, nt * self.n)
        if out is None:
            out = np.empty((nt - 1) * self.n)
        fp, _kf = _vec_ptr(out, (nt - 1) * self.n)
        J = C.c_void_p()
        B.check(B.lib().gmrfb_fem1d_spacetime_tangent(self.h, int(nt), float(dt), float(nu), wp, self._presc(prescribed),
                                                      C.byref(J), fp), self.ctx.h)
        return out, self._wrap(J, ((nt - 1) * self.n, nt * self.n))


# ------------------------------------------------------------------------------------ symbolic + numeric --
class Symbolic:
    """Symbolic analysis of one sparsity pattern (gmrfb_sym).  ``perm`` is 0-based new->old (Julia's
    ``perm=p`` is 1-based; the Julia shim passes base=1)."""

    def __init__(self, A, perm=None, ordering="nd", coords=None, ctx: Context | None = None, host_only=False,
                 nd_leaf=0, relax_small=0, relax_zeros=0.0, storage=B.STORAGE_FULL):
        A = _csc(A)
        if A.shape[0] != A.shape[1]:
            raise ValueError("matrix must be square")
        self.ctx = None if host_only else (ctx or default_context())
        self.n = A.shape[0]
        self.nnz = A.nnz
        opts = B.AnalyzeOpts()
        if perm is not None:
            opts.ordering_kind = B.ORDER_GIVEN
        else:
            opts.ordering_kind = {"nd": B.ORDER_ND, "natural": B.ORDER_NATURAL, "amd": B.ORDER_AMD, "nd_amd": B.ORDER_ND_AMD}[ordering]
        opts.storage = storage
        opts.base = 0
        self._coords = None
        if coords is not None:
            self._coords = np.ascontiguousarray(coords, dtype=np.float64)
            opts.coord_dim = self._coords.shape[1]
            opts.coords = self._coords.ctypes.data_as(B._F64P)
        opts.nd_leaf = nd_leaf
        opts.relax_small = relax_small
        opts.relax_zeros = relax_zeros
        _, cp = B.i64(A.indptr)
        _, ri = B.i64(A.indices)
        pp = None
        if perm is not None:
            _, pp = B.i64(perm)
        h = C.c_void_p()
        ch = self.ctx.h if self.ctx else None
        st = B.lib().gmrfb_analyze(ch, self.n, cp, ri, pp, C.byref(opts), C.byref(h))
        B.check(st, ch)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_sym_destroy, h)
        self._cache = None

    @property
    def info(self) -> B.SymInfo:
        info = B.SymInfo()
        B.check(B.lib().gmrfb_sym_get_info(self.h, C.byref(info)), None)
        return info

    def _get(self):
        if self._cache is None:
            n = self.n
            ns = self.info.nsuper
            perm, parent, cc, ip = (np.empty(n, np.int64) for _ in range(4))
            sptr = np.empty(ns + 1, np.int64)
            B.check(B.lib().gmrfb_sym_get(self.h, *(a.ctypes.data_as(B._I64P) for a in (perm, parent, cc, sptr, ip))),
                    None)
            self._cache = dict(perm=perm, parent=parent, colcount=cc, super_ptr=sptr, ipost=ip)
        return self._cache

    p = property(lambda self: self._get()["perm"])
    parent = property(lambda self: self._get()["parent"])
    colcount = property(lambda self: self._get()["colcount"])
    super_ptr = property(lambda self: self._get()["super_ptr"])
    ipost = property(lambda self: self._get()["ipost"])

    def maps(self):
        """(amap, relmap) of the analysis (gmrfb_sym_get_maps; diagnostic: GPU vs host analysis, bit for bit)."""
        na, nr = C.c_int64(), C.c_int64()
        B.check(B.lib().gmrfb_sym_get_maps(self.h, None, None, C.byref(na), C.byref(nr)), None)
        amap, relmap = np.empty(na.value, np.int64), np.empty(nr.value, np.int64)
        B.check(B.lib().gmrfb_sym_get_maps(self.h, amap.ctypes.data_as(B._I64P), relmap.ctypes.data_as(B._I64P), None, None),
                None)
        return amap, relmap

    def super_rows(self, s):
        cnt = C.c_int64()
        B.check(B.lib().gmrfb_sym_get_super_rows(self.h, s, None, 0, C.byref(cnt)), None)
        rows = np.empty(cnt.value, np.int64)
        B.check(B.lib().gmrfb_sym_get_super_rows(self.h, s, rows.ctypes.data_as(B._I64P), cnt.value, C.byref(cnt)), None)
        return rows


class CholeskyFactor:
    """Numeric supernodal factor (gmrfb_fac) with the CHOLMOD-factor surface the reference touches:
    ``F \\ b`` -> ``solve``; ``F.PtL \\ b`` -> ``PtL_solve``; ``F.UP \\ z`` -> ``UP_solve``; ``F.p``; ``nnz``;
    ``F.L``; ``issuccess``."""

    def __init__(self, sym: Symbolic):
        if sym.ctx is None:
            raise B.GmrfbError(B.ERR_STATE, "symbolic handle is host-only")
        self.sym = sym
        self.ctx = sym.ctx
        h = C.c_void_p()
        B.check(B.lib().gmrfb_fac_create(sym.h, C.byref(h)), self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_fac_destroy, h)
        self.success = False

    def factorize(self, nzval, check=True):
        nzval, p = B.f64(nzval)
        if nzval.size != self.sym.nnz:
            raise ValueError("nzval length does not match the analysed pattern")
        st = B.lib().gmrfb_factorize(self.h, p)
        self.success = st == B.OK
        if st == B.ERR_NOT_SPD and not check:
            return self
        B.check(st, self.ctx.h)
        return self

    def factorize_dev(self, dptr: int, check=True):
        st = B.lib().gmrfb_factorize_dev(self.h, C.c_void_p(dptr))
        self.success = st == B.OK
        if st == B.ERR_NOT_SPD and not check:
            return self
        B.check(st, self.ctx.h)
        return self

    def issuccess(self):
        return self.success

    @property
    def info(self) -> B.FacInfo:
        info = B.FacInfo()
        B.check(B.lib().gmrfb_fac_get_info(self.h, C.byref(info)), self.ctx.h)
        return info

    @property
    def p(self):
        return self.sym.p

    @property
    def nnz(self):
        return int(self.info.nnz_L)

    def logdet(self):
        return float(self.info.logdet)

    def diagL(self):
        out = np.empty(self.sym.n)
        B.check(B.lib().gmrfb_fac_diag(self.h, out.ctypes.data_as(B._F64P)), self.ctx.h)
        return out

    @property
    def L(self):
        n = self.sym.n
        cp = np.empty(n + 1, np.int64)
        B.check(B.lib().gmrfb_fac_get_L(self.h, 0, 1, cp.ctypes.data_as(B._I64P), None, None), self.ctx.h)
        ri = np.empty(cp[-1], np.int64)
        nz = np.empty(cp[-1], np.float64)
        B.check(B.lib().gmrfb_fac_get_L(self.h, 0, 1, cp.ctypes.data_as(B._I64P), ri.ctypes.data_as(B._I64P),
                                        nz.ctypes.data_as(B._F64P)), self.ctx.h)
        return sp.csc_matrix((nz, ri, cp), shape=(n, n))

    def _solve(self, mode, b):
        b = np.asarray(b, dtype=np.float64)
        one = b.ndim == 1
        X = np.asfortranarray(b.reshape(self.sym.n, -1).copy(order="F"))
        B.check(B.lib().gmrfb_solve(self.h, mode, X.ctypes.data_as(B._F64P), self.sym.n, X.shape[1]), self.ctx.h)
        return X[:, 0].copy() if one else X

    def solve(self, b, refine: "SparseMatrix | None" = None, max_iter: int = 1, return_residual=False):
        """``F \\ b``.  With ``refine=Q`` (the precision as a device matrix) up to ``max_iter`` steps of iterative
        refinement x += F \\ (b - Q x) follow the factor solve (gmrfb_solve_refined): for precisions with a large
        conditioning term (Q_eps = 1e8, scripts/solve_burger.jl:98,140) one step brings the residual down to that of
        a backward-stable substitution."""
        if refine is None:
            return self._solve(B.SOLVE_A, b)
        b = np.asarray(b, dtype=np.float64)
        one = b.ndim == 1
        X = np.asfortranarray(b.reshape(self.sym.n, -1).copy(order="F"))
        res = np.empty(X.shape[1])
        B.check(B.lib().gmrfb_solve_refined(self.h, refine.h, X.ctypes.data_as(B._F64P), self.sym.n, X.shape[1],
                                            int(max_iter), res.ctypes.data_as(B._F64P)), self.ctx.h)
        out = X[:, 0].copy() if one else X
        return (out, res) if return_residual else out

    def solve_inplace(self, X, mode=B.SOLVE_A):
        """``ldiv!(F, X)``: X (length n, or n x k column-major) is a caller-owned float64 array that is overwritten by
        the solution - exactly what the C ABI does, with no temporary on the host.  Page-locked arrays (e.g.
        ``torch.empty(n, dtype=torch.float64).pin_memory().numpy()``) make both transfers asynchronous DMA."""
        assert isinstance(X, np.ndarray) and X.dtype == np.float64 and X.shape[0] == self.sym.n
        assert X.ndim == 1 or X.flags.f_contiguous, "X must be a vector or a column-major matrix"
        assert X.flags.writeable and (X.ndim > 1 or X.flags.c_contiguous)
        nrhs = 1 if X.ndim == 1 else X.shape[1]
        B.check(B.lib().gmrfb_solve(self.h, mode, X.ctypes.data_as(B._F64P), self.sym.n, nrhs), self.ctx.h)
        return X

    def PtL_solve(self, b):
        return self._solve(B.SOLVE_PTL, b)

    def UP_solve(self, z):
        return self._solve(B.SOLVE_UP, z)

    def L_solve(self, b):
        return self._solve(B.SOLVE_L, b)

    def Lt_solve(self, b):
        return self._solve(B.SOLVE_LT, b)

    def solve_dev(self, dptr: int, nrhs=1, mode=B.SOLVE_A):
        B.check(B.lib().gmrfb_solve_dev(self.h, mode, C.c_void_p(dptr), self.sym.n, nrhs), self.ctx.h)

    def sample(self, Z, mean=None):
        Z = np.asarray(Z, dtype=np.float64)
        one = Z.ndim == 1
        Zf = np.asfortranarray(Z.reshape(self.sym.n, -1))
        X = np.empty_like(Zf, order="F")
        mp = None
        if mean is not None:
            mean, mp = B.f64(mean)
        B.check(B.lib().gmrfb_sample(self.h, mp, Zf.ctypes.data_as(B._F64P), self.sym.n, X.ctypes.data_as(B._F64P),
                                     self.sym.n, Zf.shape[1]), self.ctx.h)
        return X[:, 0].copy() if one else X

    def var_selinv(self, out=None):
        """diag(Q^-1) by Takahashi selected inversion; ``out``: optional caller-owned (e.g. page-locked) result array."""
        if out is None:
            out = np.empty(self.sym.n)
        assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.size == self.sym.n and out.flags.c_contiguous
        B.check(B.lib().gmrfb_var_selinv(self.h, out.ctypes.data_as(B._F64P)), self.ctx.h)
        return out

    def var_selinv_dev(self, dptr: int):
        B.check(B.lib().gmrfb_var_selinv_dev(self.h, C.c_void_p(dptr)), self.ctx.h)

    def var_rbmc(self, Q: SparseMatrix, Z):
        """Z: n x nsamp standard normals - a NumPy array (any order) or a CUDA float64 tensor of shape (nsamp, n)
        (one sample per contiguous row, i.e. the column-major n x nsamp matrix the C ABI expects)."""
        out = np.empty(self.sym.n)
        if hasattr(Z, "data_ptr"):
            assert Z.is_cuda and Z.is_contiguous() and Z.shape[1] == self.sym.n and str(Z.dtype) == "torch.float64"
            _torch_ready(Z)
            zp, ns = C.cast(C.c_void_p(Z.data_ptr()), B._F64P), Z.shape[0]
        else:
            Zf = np.asfortranarray(np.asarray(Z, dtype=np.float64).reshape(self.sym.n, -1))
            zp, ns = Zf.ctypes.data_as(B._F64P), Zf.shape[1]
        B.check(B.lib().gmrfb_var_rbmc(self.h, Q.h, zp, self.sym.n, ns, out.ctypes.data_as(B._F64P)), self.ctx.h)
        return out

    def var_rbmc_dev(self, Q: SparseMatrix, Z, out_dptr: int):
        """RBMC variances of the CUDA tensor ``Z`` (nsamp, n) written to device memory at ``out_dptr`` (n doubles);
        queued on the context's stream (no synchronisation, no host copy)."""
        assert Z.is_cuda and Z.is_contiguous() and Z.shape[1] == self.sym.n and str(Z.dtype) == "torch.float64"
        _torch_ready(Z)
        zp = C.cast(C.c_void_p(Z.data_ptr()), B._F64P)
        B.check(B.lib().gmrfb_var_rbmc_dev(self.h, Q.h, zp, self.sym.n, Z.shape[0], C.c_void_p(out_dptr)), self.ctx.h)

    def selinv_entries(self, rows, cols):
        rows, rp = B.i64(rows)
        cols, cp = B.i64(cols)
        out = np.empty(rows.size)
        B.check(B.lib().gmrfb_selinv_entries(self.h, 0, rows.size, rp, cp, out.ctypes.data_as(B._F64P)), self.ctx.h)
        return out


def cholesky(A, perm=None, check=True, ctx=None, coords=None, ordering="nd") -> CholeskyFactor:
    """``cholesky(Symmetric(A); perm=perm, check=check)`` (scripts/solve_burger.jl:147,
    scripts/darcy/solve_darcy_fem.jl:93).  ``perm`` is 0-based here; without it the library orders the matrix
    (``ordering="nd"`` nested dissection, ``"amd"`` approximate minimum degree, ``"natural"``)."""
    A = _csc(A)
    sym = Symbolic(A, perm=perm, ctx=ctx, coords=coords, ordering=ordering)
    return CholeskyFactor(sym).factorize(A.data, check=check)


# ------------------------------------------------------------------------------- blueprints and the GMRF --
class RBMCStrategy:
    """``RBMCStrategy(N; rng)`` (scripts/darcy/solve_darcy_gmrf-fem.jl:100).  With an ``rng`` the normals come from
    that host generator (reproducible against the oracle); without one they are drawn on the device."""

    def __init__(self, n_samples: int, rng=None):
        self.n_samples = int(n_samples)
        self.rng = rng

    def normals(self, n, device_index):
        """n x n_samples standard normals in the layout ``CholeskyFactor.var_rbmc`` takes."""
        if self.rng is None:
            import torch

            return torch.randn((self.n_samples, n), dtype=torch.float64, device=torch.device("cuda", device_index))
        return self.rng.standard_normal((self.n_samples, n)).T  # column-major n x n_samples without a copy


class TakahashiStrategy:
    """Exact marginal variances by selected inversion (north-star capability)."""


class CholeskySolverBlueprint:
    """``CholeskySolverBlueprint(; var_strategy, perm)`` (scripts/darcy/solve_darcy_gmrf-fem.jl:100,174)."""

    def __init__(self, var_strategy=None, perm=None, coords=None, ctx=None):
        self.var_strategy = var_strategy if var_strategy is not None else TakahashiStrategy()
        self.perm = perm
        self.coords = coords
        self.ctx = ctx
        self._sym_cache = None  # (indptr, indices, Symbolic) of the last pattern analysed through this blueprint

    def symbolic_for(self, Q) -> "Symbolic":
        """Symbolic analysis for Q's pattern.  A blueprint that is reused over a dataset loop (the reference passes
        ``perm=p`` for exactly that, scripts/darcy/solve_darcy_gmrf-fem.jl:169-174) keeps the analysis of the last
        pattern, so further matrices with the same pattern only pay the numeric factorisation."""
        c = self._sym_cache
        if c is not None and ((c[0] is Q.indptr and c[1] is Q.indices) or (
                c[0].size == Q.indptr.size and c[1].size == Q.indices.size
                and np.array_equal(c[0], Q.indptr) and np.array_equal(c[1], Q.indices))):
            return c[2]
        sym = Symbolic(Q, perm=self.perm, coords=self.coords, ctx=self.ctx)
        self._sym_cache = (Q.indptr, Q.indices, sym)  # pattern arrays are treated as immutable
        return sym


class GNCholeskySolverBlueprint(CholeskySolverBlueprint):
    """``GNCholeskySolverBlueprint(p)`` (scripts/burgers/solve_burgers_gmrf-fem.jl:170)."""

    def __init__(self, perm=None, **kw):
        super().__init__(perm=perm, **kw)


class _Ref:
    """Julia ``Ref``: ``x.solver_ref[]`` is spelled ``x.solver_ref[()]`` / ``x.solver_ref.value`` here."""

    def __init__(self, value):
        self.value = value

    def __getitem__(self, _):
        return self.value


class CholeskySolver:
    """What ``x.solver_ref[]`` points to: holds ``precision_chol`` (with ``.p``, ``.L``, ``nnz``)."""

    def __init__(self, gmrf, blueprint: CholeskySolverBlueprint, symbolic: Symbolic | None = None, values_dev: int = 0,
                 precision_dev: "SparseMatrix | None" = None):
        # no back-reference to the GMRF: a GMRF <-> solver cycle would leave the factor's GBs of device memory to the
        # cyclic garbage collector instead of freeing them (into the buffer pool) when the GMRF goes out of scope
        self.n, self._Q = gmrf.n, gmrf.precision
        self._prior_mean, self._information = gmrf.prior_mean, gmrf.information
        self.blueprint = blueprint
        self._precision_dev = precision_dev  # the precision as a device matrix, when the caller already has one
        Q = gmrf.precision
        sym = symbolic or blueprint.symbolic_for(Q)
        if values_dev:  # the values already sit in device memory (fixed-pattern assembly): no host round trip
            self.precision_chol = CholeskyFactor(sym).factorize_dev(values_dev)
        else:
            self.precision_chol = CholeskyFactor(sym).factorize(Q.data)
        self._mean = None
        self._var = None

    def compute_mean(self):
        if self._mean is None:
            if self._information is None:
                self._mean = self._prior_mean
            else:
                self._mean = self._prior_mean + self.precision_chol.solve(self._information)
        return self._mean

    def compute_variance(self):
        if self._var is None:
            vs = self.blueprint.var_strategy
            if isinstance(vs, RBMCStrategy):
                Z = vs.normals(self.n, self.precision_chol.ctx.device)
                # the posterior precision assembled on the device is used where it lies (no upload, no row-wise copy)
                # (only while it still holds THIS posterior: a later conditioning on the same prior overwrites its values)
                pd = self._precision_dev
                Qd = pd[0] if (pd is not None and pd[1].generation == pd[2]) else \
                    SparseMatrix(self._Q, ctx=self.precision_chol.ctx)
                self._var = self.precision_chol.var_rbmc(Qd, Z)
            else:
                self._var = self.precision_chol.var_selinv()
        return self._var

    def compute_rand(self, rng):
        z = rng.standard_normal(self.n)
        return self.precision_chol.sample(z, mean=self.compute_mean())


class GMRF:
    """``GMRF(mean, precision, solver_blueprint)`` (_research/elliptic_chen24.jl:166).  ``information`` carries
    the lazy right-hand side of a conditioned GMRF: mean = prior_mean + Q^{-1} information."""

    def __init__(self, mean, precision, solver_blueprint=None, information=None, _symbolic=None, _values_dev=0,
                 _precision_dev=None):
        self.precision = _csc(precision)
        self.n = self.precision.shape[0]
        self.prior_mean = np.asarray(mean, dtype=np.float64)
        self.information = information
        bp = solver_blueprint or CholeskySolverBlueprint()
        self.solver_ref = _Ref(CholeskySolver(self, bp, _symbolic, _values_dev, _precision_dev))
        self._cond_ws = None

    def __len__(self):
        return self.n


def precision_map(x: GMRF):
    return x.precision


def to_matrix(Q):
    return Q


def mean(x: GMRF):
    return x.solver_ref.value.compute_mean()


def var(x: GMRF):
    return x.solver_ref.value.compute_variance()


def std(x: GMRF):
    return np.sqrt(var(x))


def rand(rng, x: GMRF):
    return x.solver_ref.value.compute_rand(rng)


def sqmahal(x: GMRF, v):
    ch = x.solver_ref.value.precision_chol
    return SparseMatrix(x.precision, ctx=ch.ctx).sqmahal(v, mu=mean(x))


class _ConditioningWorkspace:
    """Device-resident state of ``condition_on_observations`` for one prior and one observation pattern: Q and A on
    the device, the fixed-pattern plan for Q + A' Q_eps A and the pattern of its result.  A dataset loop (same mesh,
    new coefficients: scripts/darcy/solve_darcy_gmrf-fem.jl:210) then only uploads A's values per problem."""

    def __init__(self, x, A, ctx):
        self.Qd = SparseMatrix(x.precision, ctx=ctx)
        if isinstance(A, SparseMatrix):  # assembled on the device (FEMP1.assemble): no upload, now or later
            self.Ad, self.dev_handle = A, A.h.value
            self.shape = tuple(A.dims()[:2])
            self.A_indptr = self.A_indices = None
        else:
            self.A_indptr, self.A_indices, self.shape = A.indptr, A.indices, A.shape
            self.Ad, self.dev_handle = SparseMatrix(A, ctx=ctx), None
        self.plan = PosteriorPrecision(self.Qd, self.Ad)
        self.pattern = None  # (indptr, indices) of the posterior precision, fetched with the first result
        self.generation = 0  # bumped by every numeric assembly: the plan's result matrix is overwritten in place

    def matches(self, A):
        if isinstance(A, SparseMatrix):
            return self.dev_handle is not None and A.h.value == self.dev_handle
        if self.dev_handle is not None:
            return False
        return A.shape == self.shape and A.indptr.size == self.A_indptr.size and A.indices.size == self.A_indices.size \
            and np.array_equal(A.indptr, self.A_indptr) and np.array_equal(A.indices, self.A_indices)


def condition_on_observations(x: GMRF, A, Q_eps, y, solver_blueprint=None) -> GMRF:
    """``condition_on_observations(x, A, Q_eps, y; solver_blueprint)``
    (scripts/darcy/solve_darcy_gmrf-fem.jl:165-167,188-189): posterior precision Q + A'Q_eps A (device
    SpGEMM on a fixed pattern), posterior mean mu + Qpost^{-1} A'Q_eps (y - A mu).  The values of the posterior
    precision go from the assembly kernel straight into the numeric factorisation; the host copy that
    ``precision_map`` returns is downloaded once per call."""
    bp = solver_blueprint or x.solver_ref.value.blueprint
    ctx = x.solver_ref.value.precision_chol.ctx
    on_device = isinstance(A, SparseMatrix)  # e.g. the stiffness matrix straight from FEMP1.assemble
    if not on_device:
        A = _csc(A)
    mu = mean(x)
    ws = x._cond_ws
    if ws is None or not ws.matches(A):
        ws = x._cond_ws = _ConditioningWorkspace(x, A, ctx)
    elif not on_device:
        ws.Ad.set_values(A.data)
    Apost = ws.plan.compute(Q_eps)
    ws.generation += 1
    if ws.pattern is None:
        P = Apost.to_scipy()
        ws.pattern = (P.indptr, P.indices)
        vals = P.data
    else:
        vals = Apost.values_host()
    Qpost = sp.csc_matrix((vals, ws.pattern[1], ws.pattern[0]), shape=x.precision.shape)
    Qpost.has_sorted_indices = True
    Qpost.has_canonical_format = True
    w = np.broadcast_to(np.asarray(Q_eps, dtype=np.float64), (ws.shape[0],))
    resid = np.asarray(y, dtype=np.float64) - ws.Ad.matvec(mu)
    info = ws.Ad.matvec(w * resid, trans=True)
    return GMRF(mu, Qpost, bp, information=info, _values_dev=Apost.values_dev(),
                _precision_dev=(Apost, ws, ws.generation))


# ------------------------------------------------------------------------------------------ Gauss-Newton --
class GaussNewtonOptimizer:
    """``GaussNewtonOptimizer(mu, Q, f_and_J, noise, y, x0; solver_bp, stopping_criterion)``
    (scripts/burgers/solve_burgers_gmrf-fem.jl:172-182).  Each step is scripts/solve_burger.jl:143-149 with the
    pattern analysed once and only the numeric factorisation repeated."""

    def __init__(self, mu, Q, f_and_J, noise, y, x0, solver_bp=None, max_steps=20, rel_tol=1e-4):
        self.mu = np.asarray(mu, dtype=np.float64)
        self.Q_prior = _csc(Q)
        self.f_and_J = f_and_J
        self.noise = float(noise)
        self.y = np.asarray(y, dtype=np.float64)
        self.xk = np.asarray(x0, dtype=np.float64).copy()
        self.bp = solver_bp or GNCholeskySolverBlueprint()
        self.max_steps, self.rel_tol = max_steps, rel_tol
        self.r_obs_norm_history = []
        self.obj_history = []
        self.Jk = None
        self.Q_mat = None
        self._sym = None
        self._fac = None
        self._plan = None

    def _objective(self, x, fx):
        d = self.mu - x
        r = self.y - fx
        return float(d @ (self.Q_prior @ d) + self.noise * (r @ r)), r

    def step(self):
        ctx = self.bp.ctx or default_context()
        fx, J = self.f_and_J(self.xk)
        # f_and_J may assemble its tangent on the device (FEMP1.assemble_cubic, FEM1D.spacetime_tangent): the matrix is
        # then a fixed-pattern SparseMatrix whose values were just overwritten in place - no upload, no host SpGEMM
        on_device = isinstance(J, SparseMatrix)
        if not on_device:
            J = _csc(J)
        if self._plan is None:
            self._Qd = SparseMatrix(self.Q_prior, ctx=ctx)
            self._Jd = J if on_device else SparseMatrix(J, ctx=ctx)
            self._plan = PosteriorPrecision(self._Qd, self._Jd)
        elif on_device:
            if J.h.value != self._Jd.h.value:
                raise ValueError("f_and_J must return the same device matrix (fixed pattern) at every iteration")
        else:
            if J.nnz != self._Jd.dims()[2]:
                raise ValueError("f_and_J must return a tangent with the same sparsity pattern at every iteration "
                                 f"({J.nnz} stored entries now, {self._Jd.dims()[2]} at the first step)")
            self._Jd.set_values(J.data)
        Apost = self._plan.compute(self.noise)
        if self._sym is None:
            pat = Apost.to_scipy()
            self._sym = Symbolic(pat, perm=self.bp.perm, coords=self.bp.coords, ctx=ctx)
            self._fac = CholeskyFactor(self._sym)
        self._fac.factorize_dev(Apost.values_dev())
        fx = np.asarray(fx, dtype=np.float64)
        if on_device:
            lin = self._Jd.matvec(self.xk) + (self.y - fx)
            rhs = self.Q_prior @ self.mu + self.noise * self._Jd.matvec(lin, trans=True)
        else:
            rhs = self.Q_prior @ self.mu + self.noise * (J.T @ (J @ self.xk + (self.y - fx)))
        self.xk = self._fac.solve(rhs)
        self.Jk = J  # a device tangent holds the values of the LAST f_and_J evaluation
        self._Apost = Apost
        return self.xk


def optimize(gno: GaussNewtonOptimizer):
    """``optimize(gno)``: iterate until the relative objective change < rel_tol or max_steps
    (stopping rule of scripts/solve_burger.jl:171-180)."""
    fx, _ = gno.f_and_J(gno.xk)
    obj, r = gno._objective(gno.xk, fx)
    last = np.inf
    steps = 0
    while (abs(last - obj) / abs(obj) > gno.rel_tol) and steps < gno.max_steps:
        gno.step()
        fx, _ = gno.f_and_J(gno.xk)
        last = obj
        obj, r = gno._objective(gno.xk, fx)
        gno.obj_history.append(obj)
        gno.r_obs_norm_history.append(float(np.linalg.norm(r)))
        steps += 1
    gno.Q_mat = gno._Apost.to_scipy() if steps > 0 else gno.Q_prior
    gno.n_steps = steps
    return gno.xk


def metrics(pred_or_x, truth, E: SparseMatrix | None = None, ctx=None):
    """``rmse``, ``max_err``, ``rel_err`` (src/metrics.jl:3-13) on the device; with ``E`` the prediction is ``E * x``
    (scripts/darcy/solve_darcy_gmrf-fem.jl:190-196)."""
    ctx = ctx or (E.ctx if E is not None else default_context())
    _, xp = B.f64(pred_or_x)
    t, tp = B.f64(truth)
    out = (C.c_double * 3)()
    B.check(B.lib().gmrfb_metrics(ctx.h, E.h if E is not None else None, xp, tp, t.size, out), ctx.h)
    return out[0], out[1], out[2]


class DeviceGaussNewton:
    """Gauss-Newton with the whole iteration on the device (gmrfb_gn_*) for a bilinear collocation residual
    ``f(w) = L w + c (A w).*(D w)`` - the Burgers residual of scripts/solve_burger.jl:127-134 with L = A1 - A0 - dt nu D2,
    A = A1, c = dt.  Same update, objective and stopping rule as ``GaussNewtonOptimizer`` / ``optimize``
    (scripts/solve_burger.jl:143-180), but residual, tangent, ``Q + noise J'J``, refactorisation and solve never leave
    the GPU; fields read afterwards as in scripts/burgers/solve_burgers_gmrf-fem.jl:184-188: ``xk``, ``Jk``, ``Q_mat``,
    ``obj_history``, ``n_steps``."""

    def __init__(self, mu, Q, L, A, D, c, noise, y, x0, solver_bp=None, max_steps=20, rel_tol=1e-4, cubic=None):
        """``cubic``: optional vector e adding ``e .* w.^3`` to the residual (the elliptic problem of
        _research/elliptic_chen24.jl: L = stiffness, e = lumped mass, A = D = 0)."""
        self.bp = solver_bp or GNCholeskySolverBlueprint()
        self.ctx = self.bp.ctx or default_context()
        Q = _csc(Q)
        L = _csc(L)
        A = _csc(A) if A is not None else sp.csc_matrix(L.shape)
        D = _csc(D) if D is not None else sp.csc_matrix(L.shape)
        if not (L.shape == A.shape == D.shape) or L.shape[1] != Q.shape[0]:
            raise ValueError("L, A, D must share one shape (m x n) with n = size of Q")
        m, n = L.shape
        U = (abs(L) + abs(A) + abs(D)).tocsc()  # union pattern
        if cubic is not None:
            U = (U + sp.identity(n, format="csc")).tocsc()  # the cubic term lives on the diagonal
        U.sort_indices()
        ucol = np.repeat(np.arange(n, dtype=np.int64), np.diff(U.indptr))
        ukey = ucol * m + U.indices

        def aligned(M):
            mcol = np.repeat(np.arange(n, dtype=np.int64), np.diff(M.indptr))
            pos = np.searchsorted(ukey, mcol * m + M.indices)
            out = np.zeros(U.nnz)
            out[pos] = M.data
            return out

        self._Qd = SparseMatrix(Q, ctx=self.ctx)
        self.xk = np.ascontiguousarray(x0, dtype=np.float64).copy()
        self.max_steps, self.rel_tol = int(max_steps), float(rel_tol)
        self._pattern = (U.indptr.astype(np.int64), U.indices.astype(np.int64), (m, n))
        opts = B.AnalyzeOpts()
        opts.ordering_kind = B.ORDER_ND
        opts.storage = B.STORAGE_FULL
        opts.base = 0
        self._coords = None
        if self.bp.coords is not None:
            self._coords = np.ascontiguousarray(self.bp.coords, dtype=np.float64)
            opts.coord_dim = self._coords.shape[1]
            opts.coords = self._coords.ctypes.data_as(B._F64P)
        pp = None
        if self.bp.perm is not None:
            _, pp = B.i64(self.bp.perm)
        _, cp = B.i64(self._pattern[0])
        _, ri = B.i64(self._pattern[1])
        _, lp = B.f64(aligned(L))
        _, ap = B.f64(aligned(A))
        _, dp = B.f64(aligned(D))
        _, yp = B.f64(y)
        _, mp = B.f64(mu)
        ep = None
        if cubic is not None:
            _, ep = B.f64(cubic)
        h = C.c_void_p()
        B.check(B.lib().gmrfb_gn_create(self.ctx.h, self._Qd.h, m, cp, ri, lp, ap, dp, ep, 0, float(c), float(noise), yp, mp, pp,
                                        C.byref(opts), C.byref(h)), self.ctx.h)
        self.h = h
        self._fin = weakref.finalize(self, B.lib().gmrfb_gn_destroy, h)
        self.obj_history, self.n_steps = [], 0

    def optimize(self):
        hist = np.zeros(self.max_steps + 1)
        steps = C.c_int32()
        B.check(B.lib().gmrfb_gn_optimize(self.h, self.xk.ctypes.data_as(B._F64P), self.max_steps, self.rel_tol,
                                          C.byref(steps), hist.ctypes.data_as(B._F64P)), self.ctx.h)
        self.n_steps = steps.value
        self.obj_history = hist[: self.n_steps + 1].tolist()
        return self.xk

    def _view(self, which):
        out = [C.c_void_p() for _ in range(4)]
        B.check(B.lib().gmrfb_gn_get(self.h, *(C.byref(o) for o in out)), self.ctx.h)
        return out[which]

    @property
    def Jk(self):
        """Tangent at the last linearisation point."""
        return SparseMatrix(None, ctx=self.ctx, _handle=self._view(2), _owner=self).to_scipy()

    @property
    def Q_mat(self):
        """Q + noise J_k' J_k of the last step (the posterior precision the reference builds at :184-186)."""
        return SparseMatrix(None, ctx=self.ctx, _handle=self._view(3), _owner=self).to_scipy()


# ------------------------------------------------------------------------------- block tridiagonal factor --
class _DenseChol:
    """Stand-in for LinearAlgebra.Cholesky: ``.L`` is the lower factor."""

    def __init__(self, L):
        self.L = L


class TridiagonalCholeskyFactor:
    """``TridiagonalCholeskyFactor{T}`` (src/tridiagonal_cholesky.jl:5-9): ``N`` total rows, ``chos`` the
    per-block factors, ``Cs`` the sub-diagonal blocks (``Cs[k]`` belongs to block row k+1).  The factor lives
    on the device; blocks are copied out on access."""

    def __init__(self, handle, ctx, N, b, nblocks):
        self.h, self.ctx, self.N, self.b, self.nblocks = handle, ctx, N, b, nblocks
        self._fin = weakref.finalize(self, B.lib().gmrfb_btd_destroy, handle)

    def _block(self, i, which):
        out = np.empty((self.b, self.b), order="F")
        B.check(B.lib().gmrfb_btd_get_block(self.h, i, which, out.ctypes.data_as(B._F64P), self.b), self.ctx.h)
        return out

    @property
    def chos(self):
        return [_DenseChol(self._block(i, B.BTD_BLOCK_L)) for i in range(self.nblocks)]

    @property
    def Cs(self):
        return [self._block(i, B.BTD_BLOCK_C) for i in range(self.nblocks - 1)]

    @property
    def info(self) -> B.BtdInfo:
        info = B.BtdInfo()
        B.check(B.lib().gmrfb_btd_get_info(self.h, C.byref(info)), self.ctx.h)
        return info

    def _solve(self, mode, b):
        b = np.asarray(b, dtype=np.float64)
        n = self.b * self.nblocks
        if b.shape[0] != n:
            # the factor covers N_blocks * (size(A, 1) ÷ N_blocks) rows (src/tridiagonal_cholesky.jl:66); the reference's
            # chunked solves fail on any other length (`make_chunks` puts the remainder rows into the last chunk)
            raise ValueError(f"right-hand side has {b.shape[0]} rows, the block factor covers {n} "
                             f"({self.nblocks} blocks of {self.b})")
        one = b.ndim == 1
        X = np.asfortranarray(b.reshape(n, -1).copy(order="F"))
        B.check(B.lib().gmrfb_btd_solve(self.h, mode, X.ctypes.data_as(B._F64P), n, X.shape[1]), self.ctx.h)
        return X[:, 0].copy() if one else X

    def logdet(self):
        out = C.c_double()
        B.check(B.lib().gmrfb_btd_logdet(self.h, C.byref(out)), self.ctx.h)
        return out.value

    def selinv_diag(self):
        out = np.empty(self.b * self.nblocks)
        B.check(B.lib().gmrfb_btd_selinv_diag(self.h, out.ctypes.data_as(B._F64P)), self.ctx.h)
        return out


def tridiagonal_cholesky(A, N_blocks, ctx=None) -> TridiagonalCholeskyFactor:
    """``tridiagonal_cholesky(A::SparseMatrixCSC, N_blocks)`` (src/tridiagonal_cholesky.jl:65-82)."""
    ctx = ctx or default_context()
    A = _csc(A)
    _, cp = B.i64(A.indptr)
    _, ri = B.i64(A.indices)
    _, nz = B.f64(A.data)
    h = C.c_void_p()
    st = B.lib().gmrfb_btd_factor(ctx.h, A.shape[0], cp, ri, nz, 0, N_blocks, C.byref(h))
    if st != B.OK and h.value:
        B.lib().gmrfb_btd_destroy(h)
    B.check(st, ctx.h)
    return TridiagonalCholeskyFactor(h, ctx, A.shape[0], A.shape[0] // N_blocks, N_blocks)


def tridiagonal_cholesky_dense(D, Bsub, ctx=None) -> TridiagonalCholeskyFactor:
    """Dense-block entry: D[b,b,N] diagonal blocks, Bsub[b,b,N-1] sub-diagonal blocks (Fortran order), or the same
    memory as CUDA float64 tensors of shape (N, b, b) / (N-1, b, b) (element [k, j, i] = entry (i, j) of block k)."""
    ctx = ctx or default_context()
    if hasattr(D, "data_ptr"):  # device-resident blocks (torch CUDA tensors)
        assert D.is_cuda and D.is_contiguous() and str(D.dtype) == "torch.float64"
        _torch_ready(D)
        N, b, _ = D.shape
        Dp = C.cast(C.c_void_p(D.data_ptr()), B._F64P)
        if N > 1:
            assert Bsub.is_cuda and Bsub.is_contiguous() and str(Bsub.dtype) == "torch.float64" \
                and tuple(Bsub.shape) == (N - 1, b, b) and Bsub.device == D.device
        Bp = C.cast(C.c_void_p(Bsub.data_ptr()), B._F64P) if N > 1 else None
    else:
        D = np.asfortranarray(D, dtype=np.float64)
        b, _, N = D.shape
        Dp = D.ctypes.data_as(B._F64P)
        Bp = None
        if N > 1:
            Bsub = np.asfortranarray(Bsub, dtype=np.float64)
            Bp = Bsub.ctypes.data_as(B._F64P)
    h = C.c_void_p()
    st = B.lib().gmrfb_btd_factor_dense(ctx.h, b, N, Dp, Bp, C.byref(h))
    if st != B.OK and h.value:
        B.lib().gmrfb_btd_destroy(h)
    B.check(st, ctx.h)
    return TridiagonalCholeskyFactor(h, ctx, b * N, b, N)


def tridiagonal_cholesky_ssm(D_first, D_mid, D_last, B_sub, N_blocks, ctx=None) -> TridiagonalCholeskyFactor:
    """Block-tridiagonal factor of a constant-mesh implicit-Euler state-space prior given by its four distinct blocks
    (gmrfb_btd_factor_ssm; ingredients of src/spdes/shallow_water.jl:198-228)."""
    ctx = ctx or default_context()
    if hasattr(D_first, "data_ptr"):
        # device-resident blocks: CUDA float64 tensors of shape (b, b), element [j, i] = entry (i, j) (column-major)
        blocks = [D_first, D_mid, D_last, B_sub]
        for M in blocks:
            assert M is None or (M.is_cuda and M.is_contiguous() and str(M.dtype) == "torch.float64")
        _torch_ready(blocks[0])
        b = blocks[0].shape[0]
        ptrs = [C.cast(C.c_void_p(M.data_ptr()), B._F64P) if M is not None else None for M in blocks]
    else:
        blocks = [np.asfortranarray(M, dtype=np.float64) if M is not None else None for M in (D_first, D_mid, D_last, B_sub)]
        b = blocks[0].shape[0]
        ptrs = [M.ctypes.data_as(B._F64P) if M is not None else None for M in blocks]
    h = C.c_void_p()
    st = B.lib().gmrfb_btd_factor_ssm(ctx.h, b, int(N_blocks), *ptrs, C.byref(h))
    if st != B.OK and h.value:
        B.lib().gmrfb_btd_destroy(h)
    B.check(st, ctx.h)
    return TridiagonalCholeskyFactor(h, ctx, b * int(N_blocks), b, int(N_blocks))


def forward_solve(L, b):
    """``forward_solve`` (src/tridiagonal_cholesky.jl:35-52): block factor -> L^{-1} b; sparse factor -> F.PtL \\ b."""
    if isinstance(L, TridiagonalCholeskyFactor):
        return L._solve(B.BTD_SOLVE_FWD, b)
    return L.PtL_solve(b)


def backward_solve(L, b):
    """``backward_solve`` (src/tridiagonal_cholesky.jl:16-33): block factor -> L^{-T} b; sparse factor -> F.UP \\ b."""
    if isinstance(L, TridiagonalCholeskyFactor):
        return L._solve(B.BTD_SOLVE_BWD, b)
    return L.UP_solve(b)


def ldiv(L: TridiagonalCholeskyFactor, b):
    """``ldiv(L, b)`` (src/tridiagonal_cholesky.jl:60-63), intended semantics A^{-1} b."""
    return L._solve(B.BTD_SOLVE_A, b)


def ldiv_(y, L: TridiagonalCholeskyFactor, b):
    """``ldiv!(y, L, b)`` (:54-58)."""
    y[...] = ldiv(L, b)
    return y
