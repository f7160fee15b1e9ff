"""ctypes binding of libgmrfb.so — exactly the symbols declared in include/gmrfb.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (csrc/Makefile, nvcc, sm_100a).
There is no CPU fallback: if the library is missing, or no B200-class GPU is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgmrfb.so")

OK, ERR_INVALID, ERR_NOT_SPD, ERR_ALLOC, ERR_CUDA, ERR_COMM, ERR_STATE, ERR_INTERNAL = range(8)
ORDER_GIVEN, ORDER_NATURAL, ORDER_ND, ORDER_AMD, ORDER_ND_AMD = 0, 1, 2, 3, 4
STORAGE_FULL, STORAGE_LOWER, STORAGE_UPPER = 0, 1, 2
SOLVE_A, SOLVE_PTL, SOLVE_UP, SOLVE_L, SOLVE_LT = range(5)
BTD_BLOCK_L, BTD_BLOCK_C = 0, 1
BTD_SOLVE_A, BTD_SOLVE_FWD, BTD_SOLVE_BWD = 0, 1, 2


class GmrfbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libgmrfb status {status}: {msg}")
        self.status = status


class NotPositiveDefinite(GmrfbError):
    """Julia: PosDefException (or the `check=false` flag, scripts/solve_burger.jl:147)."""


class AnalyzeOpts(C.Structure):
    _fields_ = [("ordering_kind", C.c_int32), ("storage", C.c_int32), ("base", C.c_int32),
                ("coord_dim", C.c_int32), ("coords", C.POINTER(C.c_double)), ("nd_leaf", C.c_int32),
                ("relax_small", C.c_int32), ("relax_zeros", C.c_double)]


class SymInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("nnz_lower_A", C.c_int64), ("nnz_L", C.c_int64), ("nnz_L_stored", C.c_int64),
                ("flops", C.c_double), ("nsuper", C.c_int64), ("nlevels", C.c_int64), ("max_front", C.c_int64),
                ("front_bytes", C.c_int64)]


class FacInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("fail_column", C.c_int64), ("logdet", C.c_double), ("nnz_L", C.c_int64)]


class ProfileEntry(C.Structure):
    _fields_ = [("kind", C.c_int32), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double), ("name", C.c_char * 32)]


class BtdInfo(C.Structure):
    _fields_ = [("b", C.c_int64), ("nblocks", C.c_int64), ("status", C.c_int32), ("fail_block", C.c_int64),
                ("flops", C.c_double)]


_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
_F64P = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/gmrfb.h one to one
SIGNATURES = {
    "gmrfb_version": (C.c_int32, []),
    "gmrfb_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(_P)]),
    "gmrfb_ctx_destroy": (C.c_int32, [_P]),
    "gmrfb_last_error": (C.c_char_p, [_P]),
    "gmrfb_ctx_sync": (C.c_int32, [_P]),
    "gmrfb_ctx_stream": (C.c_uint64, [_P]),
    "gmrfb_ctx_launch_count": (C.c_int64, [_P]),
    "gmrfb_ctx_profile_begin": (C.c_int32, [_P]),
    "gmrfb_ctx_profile_end": (C.c_int32, [_P, C.POINTER(ProfileEntry), C.c_int32, C.POINTER(C.c_int32)]),
    "gmrfb_analyze": (C.c_int32, [_P, C.c_int64, _I64P, _I64P, _I64P, C.POINTER(AnalyzeOpts), C.POINTER(_P)]),
    "gmrfb_sym_destroy": (C.c_int32, [_P]),
    "gmrfb_sym_get_info": (C.c_int32, [_P, C.POINTER(SymInfo)]),
    "gmrfb_sym_get": (C.c_int32, [_P, _I64P, _I64P, _I64P, _I64P, _I64P]),
    "gmrfb_sym_get_super_rows": (C.c_int32, [_P, C.c_int64, _I64P, C.c_int64, _I64P]),
    "gmrfb_sym_get_maps": (C.c_int32, [_P, _I64P, _I64P, _I64P, _I64P]),
    "gmrfb_fac_create": (C.c_int32, [_P, C.POINTER(_P)]),
    "gmrfb_fac_destroy": (C.c_int32, [_P]),
    "gmrfb_factorize": (C.c_int32, [_P, _F64P]),
    "gmrfb_factorize_dev": (C.c_int32, [_P, _P]),
    "gmrfb_fac_get_info": (C.c_int32, [_P, C.POINTER(FacInfo)]),
    "gmrfb_fac_diag": (C.c_int32, [_P, _F64P]),
    "gmrfb_fac_get_L": (C.c_int32, [_P, C.c_int32, C.c_int32, _I64P, _I64P, _F64P]),
    "gmrfb_solve": (C.c_int32, [_P, C.c_int32, _F64P, C.c_int64, C.c_int64]),
    "gmrfb_solve_dev": (C.c_int32, [_P, C.c_int32, _P, C.c_int64, C.c_int64]),
    "gmrfb_solve_refined": (C.c_int32, [_P, _P, _F64P, C.c_int64, C.c_int64, C.c_int32, _F64P]),
    "gmrfb_pool_trim": (C.c_int32, [C.c_int64, _I64P]),
    "gmrfb_sample": (C.c_int32, [_P, _F64P, _F64P, C.c_int64, _F64P, C.c_int64, C.c_int64]),
    "gmrfb_var_selinv": (C.c_int32, [_P, _F64P]),
    "gmrfb_var_selinv_dev": (C.c_int32, [_P, _P]),
    "gmrfb_var_rbmc": (C.c_int32, [_P, _P, _F64P, C.c_int64, C.c_int64, _F64P]),
    "gmrfb_var_rbmc_dev": (C.c_int32, [_P, _P, _F64P, C.c_int64, C.c_int64, _P]),
    "gmrfb_selinv_entries": (C.c_int32, [_P, C.c_int32, C.c_int64, _I64P, _I64P, _F64P]),
    "gmrfb_spm_create": (C.c_int32, [_P, C.c_int64, C.c_int64, _I64P, _I64P, _F64P, C.c_int32, C.POINTER(_P)]),
    "gmrfb_spm_set_values": (C.c_int32, [_P, _F64P]),
    "gmrfb_spm_destroy": (C.c_int32, [_P]),
    "gmrfb_spmv": (C.c_int32, [_P, C.c_int32, C.c_double, _F64P, C.c_double, _F64P]),
    "gmrfb_sqmahal": (C.c_int32, [_P, _F64P, _F64P, _F64P]),
    "gmrfb_postprec_create": (C.c_int32, [_P, _P, _P, C.POINTER(_P)]),
    "gmrfb_postprec_destroy": (C.c_int32, [_P]),
    "gmrfb_postprec_compute": (C.c_int32, [_P, C.c_double, _F64P, C.POINTER(_P)]),
    "gmrfb_metrics": (C.c_int32, [_P, _P, _F64P, _F64P, C.c_int64, _F64P]),
    "gmrfb_postprec_result": (C.c_int32, [_P, C.POINTER(_P)]),
    "gmrfb_fem_create": (C.c_int32, [_P, C.c_int64, _F64P, C.c_int64, _I64P, C.c_int32, C.POINTER(_P)]),
    "gmrfb_fem_destroy": (C.c_int32, [_P]),
    "gmrfb_fem_get_mass": (C.c_int32, [_P, _F64P]),
    "gmrfb_fem_set_coeff_grid": (C.c_int32, [_P, C.c_int64, _F64P, C.c_int64, _F64P]),
    "gmrfb_fem_assemble": (C.c_int32, [_P, _P, _P, C.POINTER(_P)]),
    "gmrfb_fem_matern_precision": (C.c_int32, [_P, C.c_double, C.c_double, _P, C.c_double, C.POINTER(_P)]),
    "gmrfb_fem_assemble_cubic": (C.c_int32, [_P, _P, C.c_int32, C.c_double, _P, C.POINTER(_P), _P]),
    "gmrfb_spgemm_create": (C.c_int32, [_P, _P, _P, C.POINTER(_P)]),
    "gmrfb_spgemm_destroy": (C.c_int32, [_P]),
    "gmrfb_spgemm_compute": (C.c_int32, [_P, C.c_double, _F64P, C.POINTER(_P)]),
    "gmrfb_fem2d_create": (C.c_int32, [_P, C.c_int32, C.c_int64, _F64P, C.c_int64, _I64P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "gmrfb_fem2d_destroy": (C.c_int32, [_P]),
    "gmrfb_fem2d_info": (C.c_int32, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), _I64P]),
    "gmrfb_fem2d_set_coeff_grid": (C.c_int32, [_P, C.c_int64, _F64P, C.c_int64, _F64P]),
    "gmrfb_fem2d_stiffness": (C.c_int32, [_P, _P, _P, C.c_double, C.POINTER(_P), _P]),
    "gmrfb_fem2d_mass": (C.c_int32, [_P, C.c_int32, C.POINTER(_P), _F64P]),
    "gmrfb_fem2d_matern_precision": (C.c_int32, [_P, C.c_double, C.c_double, C.c_int32, _P, C.c_double, C.POINTER(_P)]),
    "gmrfb_fem2d_assemble_cubic": (C.c_int32, [_P, _P, C.c_double, _P, C.POINTER(_P), _P]),
    "gmrfb_fem1d_create": (C.c_int32, [_P, C.c_int64, C.c_int64, _I64P, _F64P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "gmrfb_fem1d_destroy": (C.c_int32, [_P]),
    "gmrfb_fem1d_mass_stiffness": (C.c_int32, [_P, C.c_int32, _P, C.POINTER(_P), C.POINTER(_P)]),
    "gmrfb_fem1d_advection": (C.c_int32, [_P, _P, _P, C.POINTER(_P), _P]),
    "gmrfb_fem1d_spacetime_tangent": (C.c_int32, [_P, C.c_int64, C.c_double, C.c_double, _P, _P, C.POINTER(_P), _P]),
    "gmrfb_gn_create": (C.c_int32, [_P, _P, C.c_int64, _I64P, _I64P, _F64P, _F64P, _F64P, _F64P, C.c_int32, C.c_double, C.c_double,
                                    _F64P, _F64P, _I64P, C.POINTER(AnalyzeOpts), C.POINTER(_P)]),
    "gmrfb_gn_optimize": (C.c_int32, [_P, _F64P, C.c_int32, C.c_double, C.POINTER(C.c_int32), _F64P]),
    "gmrfb_gn_get": (C.c_int32, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "gmrfb_gn_destroy": (C.c_int32, [_P]),
    "gmrfb_spm_dims": (C.c_int32, [_P, _I64P, _I64P, _I64P]),
    "gmrfb_spm_get": (C.c_int32, [_P, C.c_int32, _I64P, _I64P, _F64P]),
    "gmrfb_spm_values_dev": (_P, [_P]),
    "gmrfb_btd_factor": (C.c_int32, [_P, C.c_int64, _I64P, _I64P, _F64P, C.c_int32, C.c_int64, C.POINTER(_P)]),
    "gmrfb_btd_factor_dense": (C.c_int32, [_P, C.c_int64, C.c_int64, _F64P, _F64P, C.POINTER(_P)]),
    "gmrfb_btd_factor_ssm": (C.c_int32, [_P, C.c_int64, C.c_int64, _F64P, _F64P, _F64P, _F64P, C.POINTER(_P)]),
    "gmrfb_btd_destroy": (C.c_int32, [_P]),
    "gmrfb_btd_get_block": (C.c_int32, [_P, C.c_int64, C.c_int32, _F64P, C.c_int64]),
    "gmrfb_btd_solve": (C.c_int32, [_P, C.c_int32, _F64P, C.c_int64, C.c_int64]),
    "gmrfb_btd_logdet": (C.c_int32, [_P, _F64P]),
    "gmrfb_btd_selinv_diag": (C.c_int32, [_P, _F64P]),
    "gmrfb_btd_get_info": (C.c_int32, [_P, C.POINTER(BtdInfo)]),
    "gmrfb_btd_dist_create": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _F64P, _F64P, C.POINTER(_P)]),
    "gmrfb_btd_dist_iface_count": (C.c_int64, [_P]),
    "gmrfb_btd_dist_get_iface": (C.c_int32, [_P, _P]),
    "gmrfb_btd_dist_reduce": (C.c_int32, [_P, _P]),
    "gmrfb_btd_dist_solve_count": (C.c_int64, [_P, C.c_int64]),
    "gmrfb_btd_dist_solve_begin": (C.c_int32, [_P, _F64P, C.c_int64, C.c_int64, _P]),
    "gmrfb_btd_dist_solve_end": (C.c_int32, [_P, _P, _F64P, C.c_int64, C.c_int64]),
    "gmrfb_btd_dist_logdet": (C.c_int32, [_P, _F64P, _F64P]),
    "gmrfb_btd_dist_destroy": (C.c_int32, [_P]),
}

_lib = None


def lib():
    """Load libgmrfb.so (once).  Raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  libgmrfb has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_I64P)


def f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_F64P)


def check(status, ctx_handle):
    if status == OK:
        return
    msg = lib().gmrfb_last_error(ctx_handle)
    msg = msg.decode() if msg else ""
    if status == ERR_NOT_SPD:
        raise NotPositiveDefinite(status, msg)
    raise GmrfbError(status, msg)
