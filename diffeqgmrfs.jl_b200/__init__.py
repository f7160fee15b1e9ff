"""diffeqgmrfs.jl_b200 — B200-native (sm_100a) drop-in for the precision-matrix linear-algebra hot path of
timweiland/DiffEqGMRFs.jl: supernodal sparse Cholesky, triangular solves, samples, marginal variances and the
dense block-tridiagonal Cholesky, as CUDA kernels behind the C ABI of include/gmrfb.h.

`csrc/` holds the CUDA kernels and the C ABI (built in-tree into libgmrfb.so); `solver.py` mirrors the
reference's blueprint/GMRF interface on top of it; `workloads.py` generates the synthetic configurations.
There is no CPU fallback: importing works anywhere, but every numeric call needs the built library and a GPU.
"""
from . import _lib, dist, workloads  # noqa: F401
from .solver import (  # noqa: F401
    CholeskyFactor, CholeskySolverBlueprint, Context, DeviceGaussNewton, FEM1D, FEMLagrange, FEMP1, GMRF, GNCholeskySolverBlueprint, GaussNewtonOptimizer,
    PosteriorPrecision, RBMCStrategy, SparseMatrix, SparseProduct, Symbolic, TakahashiStrategy, TridiagonalCholeskyFactor,
    backward_solve, cholesky, condition_on_observations, default_context, forward_solve, ldiv, ldiv_, mean, metrics,
    optimize, pool_trim, precision_map, rand, sqmahal, std, to_matrix, tridiagonal_cholesky, tridiagonal_cholesky_dense, tridiagonal_cholesky_ssm, var,
)
from ._lib import GmrfbError, NotPositiveDefinite  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
