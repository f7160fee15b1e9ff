# GMRFB200GMRFExt — package extension that makes the B200 library a GaussianMarkovRandomFields.jl solver backend, so that
# the reference's scripts switch backends by changing one token:
#
#     scripts/darcy/solve_darcy_gmrf-fem.jl:100   cbp  = CholeskySolverBlueprint(var_strategy=RBMCStrategy(100))
#                                            ->   cbp  = B200CholeskySolverBlueprint(var_strategy=RBMCStrategy(100))
#     scripts/darcy/solve_darcy_gmrf-fem.jl:174   cbp2 = CholeskySolverBlueprint(var_strategy=RBMCStrategy(50; rng=rng), perm=p)
#                                            ->   cbp2 = B200CholeskySolverBlueprint(var_strategy=RBMCStrategy(50; rng=rng), perm=p)
#     scripts/burgers/solve_burgers_gmrf-fem.jl:170  GNCholeskySolverBlueprint(p) -> B200GNCholeskySolverBlueprint(p)
#
# after which `condition_on_observations(x, A, Q_ϵ, ys; solver_blueprint = cbp)` (:165-167), `mean`, `std`, `rand`,
# `x_cond.solver_ref[].precision_chol.p` and `nnz(x_cond.solver_ref[].precision_chol)` (:169-170) run on the GPU.
#
# NOT EXECUTED IN THE BUILD CONTAINER (no Julia, and GaussianMarkovRandomFields.jl — unregistered, installed from its
# GitHub HEAD by the reference, README.md:17-19 — is not available offline).  The solver protocol below
# (`construct_solver`, `compute_mean`, `compute_variance`, `compute_rand!`, `AbstractSolver`, `AbstractSolverBlueprint`,
# `precision_map`, `to_matrix`, `information_vector`) is written from the call sites in the reference's scripts and from
# recollection of that package (SURVEY.md appendix A, label [RECALL]); a maintainer with the package at hand adjusts the
# imported names if its HEAD differs.  The arithmetic is all behind `GMRFB200` (ccall -> libgmrfb), exercised here through
# the Python mirror of the same glue (`diffeqgmrfs.jl_b200/solver.py`: CholeskySolver.compute_mean / compute_variance /
# compute_rand, tests/test_gpu_sparse.py::test_gmrf_interface).
module GMRFB200GMRFExt

using GMRFB200
using GaussianMarkovRandomFields
using LinearAlgebra, SparseArrays, Random

import GaussianMarkovRandomFields: construct_solver, compute_mean, compute_variance, compute_rand!

const GMRFs = GaussianMarkovRandomFields

"What `x.solver_ref[]` returns for the B200 blueprints: the fields the reference's scripts touch are kept."
struct B200CholeskySolver{G} <: GMRFs.AbstractSolver
    gmrf::G
    precision_chol::GMRFB200.B200Factor      # `.p`, `nnz(·)`, `\`, `.PtL`, `.UP`
    var_strategy::Any
    Q::SparseMatrixCSC{Float64,Int64}        # host copy of the precision (RBMC accumulates against it on the device)
    mean_cache::Base.RefValue{Union{Nothing,Vector{Float64}}}
end

_precision(x) = SparseMatrixCSC{Float64,Int64}(sparse(GMRFs.to_matrix(GMRFs.precision_map(x))))

function _construct(bp_perm, var_strategy, x)
    Q = _precision(x)
    F = GMRFB200.b200_cholesky(Q; perm = bp_perm)       # perm === nothing: the library orders (nested dissection)
    return B200CholeskySolver(x, F, var_strategy, Q, Ref{Union{Nothing,Vector{Float64}}}(nothing))
end

construct_solver(bp::GMRFB200.B200CholeskySolverBlueprint, x::GMRFs.AbstractGMRF) = _construct(bp.perm, bp.var_strategy, x)
construct_solver(bp::GMRFB200.B200GNCholeskySolverBlueprint, x::GMRFs.AbstractGMRF) = _construct(bp.perm, :takahashi, x)

# posterior mean of a (possibly conditioned) GMRF: prior mean + Q⁻¹ (information vector), one factor solve
function compute_mean(s::B200CholeskySolver)
    if s.mean_cache[] === nothing
        μ = Vector{Float64}(mean(s.gmrf.prior))                     # [RECALL] field / accessor names of a conditioned GMRF
        rhs = Vector{Float64}(GMRFs.information_vector(s.gmrf))     # A' Q_ϵ (y - A μ)
        s.mean_cache[] = μ .+ (s.precision_chol \ rhs)
    end
    return s.mean_cache[]
end

# marginal variances: RBMCStrategy(N; rng) (scripts/darcy/solve_darcy_gmrf-fem.jl:100,174,192) or Takahashi
function compute_variance(s::B200CholeskySolver)
    vs = s.var_strategy
    if vs isa GMRFs.RBMCStrategy
        n = size(s.Q, 1)
        Z = randn(vs.rng, n, vs.n_samples)                          # the caller's generator: "same seed" = same z
        return GMRFB200.var_rbmc(s.precision_chol, s.Q, Z)          # all N samples in one backward sweep over L
    end
    return GMRFB200.var_selinv(s.precision_chol)
end

# one posterior sample x = μ + P' L⁻ᵀ z  (`rand(rng, x_cond)`, scripts/darcy/solve_darcy_gmrf-fem.jl:191)
function compute_rand!(s::B200CholeskySolver, rng::Random.AbstractRNG, out::AbstractVector)
    z = randn(rng, length(out))
    out .= GMRFB200.sample(s.precision_chol, z; mean = compute_mean(s))
    return out
end

end # module
