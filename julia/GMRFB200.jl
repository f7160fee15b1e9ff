# GMRFB200.jl — thin `ccall` shim over libgmrfb (include/gmrfb.h).
#
# NOT EXECUTED IN THE BUILD CONTAINER (Julia is not installed there); kept trivially thin on purpose:
# pointer passing, `GC.@preserve`, status -> exception.  The Python ctypes host
# (`diffeqgmrfs.jl_b200/solver.py`) binds the same symbols with the same argument meaning and is what the
# parity tests drive.  INTEGRATION.md shows how the reference's scripts select these types.
#
# Reference surface mirrored (file:line under timweiland/DiffEqGMRFs.jl):
#   cholesky(Symmetric(A); perm, check)            scripts/solve_burger.jl:147, scripts/darcy/solve_darcy_fem.jl:93
#   F \ b, F.PtL \ b, F.UP \ z, F.p, nnz(F), F.L    src/tridiagonal_cholesky.jl:20-22,39-41; scripts/darcy/solve_darcy_gmrf-fem.jl:169-170
#   tridiagonal_cholesky / forward_solve / ...      src/tridiagonal_cholesky.jl:5-82
#   CholeskySolverBlueprint(; var_strategy, perm)   scripts/darcy/solve_darcy_gmrf-fem.jl:100,174
module GMRFB200

using LinearAlgebra, SparseArrays

export B200Context, B200Factor, b200_cholesky, b200_tridiagonal_cholesky, B200TridiagonalCholeskyFactor,
       B200CholeskySolverBlueprint, B200GNCholeskySolverBlueprint, forward_solve, backward_solve, ldiv, ldiv!,
       var_selinv, var_rbmc, logdet_factor, issuccess, B200SparseMatrix, to_sparse, B200FEMP1, B200FEM1D, lumped_mass,
       set_coeff_grid!, assemble_darcy, matern_precision, assemble_cubic, assemble_burgers_mass_diffusion_matrices,
       assemble_burgers_advection_matrix, burgers_f_and_J, B200FEMLagrange, assemble_mass, B200SparseProduct, product!

const libgmrfb = get(ENV, "GMRFB_LIB", joinpath(@__DIR__, "..", "diffeqgmrfs.jl_b200", "libgmrfb.so"))

const GMRFB_OK = Int32(0)
const GMRFB_ERR_NOT_SPD = Int32(2)
const ORDER_GIVEN, ORDER_NATURAL, ORDER_ND, ORDER_AMD, ORDER_ND_AMD = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const SOLVE_A, SOLVE_PTL, SOLVE_UP, SOLVE_L, SOLVE_LT = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const BTD_SOLVE_A, BTD_SOLVE_FWD, BTD_SOLVE_BWD = Int32(0), Int32(1), Int32(2)

struct GmrfbError <: Exception
    status::Int32
    msg::String
end

# mirrors gmrfb_analyze_opts
struct AnalyzeOpts
    ordering_kind::Int32
    storage::Int32
    base::Int32
    coord_dim::Int32
    coords::Ptr{Float64}
    nd_leaf::Int32
    relax_small::Int32
    relax_zeros::Float64
end

# mirrors gmrfb_sym_info / gmrfb_fac_info / gmrfb_btd_info
struct SymInfo
    n::Int64; nnz_lower_A::Int64; nnz_L::Int64; nnz_L_stored::Int64; flops::Float64
    nsuper::Int64; nlevels::Int64; max_front::Int64; front_bytes::Int64
end
struct FacInfo
    status::Int32; fail_column::Int64; logdet::Float64; nnz_L::Int64
end
struct BtdInfo
    b::Int64; nblocks::Int64; status::Int32; fail_block::Int64; flops::Float64
end

mutable struct B200Context
    h::Ptr{Cvoid}
    function B200Context(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:gmrfb_ctx_create, libgmrfb), Int32, (Int32, Ref{Ptr{Cvoid}}), device, out)
        st == GMRFB_OK || throw(GmrfbError(st, unsafe_string(ccall((:gmrfb_last_error, libgmrfb), Cstring, (Ptr{Cvoid},), C_NULL))))
        ctx = new(out[])
        finalizer(c -> ccall((:gmrfb_ctx_destroy, libgmrfb), Int32, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

const _default_ctx = Ref{Union{Nothing,B200Context}}(nothing)
default_context() = (_default_ctx[] === nothing && (_default_ctx[] = B200Context(0)); _default_ctx[]::B200Context)

function _check(ctx::B200Context, st::Int32)
    st == GMRFB_OK && return
    msg = unsafe_string(ccall((:gmrfb_last_error, libgmrfb), Cstring, (Ptr{Cvoid},), ctx.h))
    st == GMRFB_ERR_NOT_SPD && throw(PosDefException(1))   # what stdlib `cholesky` throws with check=true
    throw(GmrfbError(st, msg))
end

# ------------------------------------------------------------------------------------------ sparse factor --
"Numeric supernodal factor with the CHOLMOD.Factor surface the reference touches (`.p`, `\\`, `.PtL`, `.UP`, `nnz`)."
mutable struct B200Factor
    ctx::B200Context
    sym::Ptr{Cvoid}
    fac::Ptr{Cvoid}
    n::Int
    p::Vector{Int64}          # 1-based, new -> old: identical to the `perm` passed in when one is given
    success::Bool
end

function _destroy(F::B200Factor)
    ccall((:gmrfb_fac_destroy, libgmrfb), Int32, (Ptr{Cvoid},), F.fac)
    ccall((:gmrfb_sym_destroy, libgmrfb), Int32, (Ptr{Cvoid},), F.sym)
end

"""
    b200_cholesky(A; perm=nothing, check=true, coords=nothing, ordering=:nd, ctx=default_context())

Drop-in for `cholesky(Symmetric(A); perm, check)` on a `SparseMatrixCSC{Float64,Int64}` holding both triangles.
Julia's 1-based `colptr`/`rowval`/`perm` are passed untouched (`base = 1`).  Without `perm` the library orders the
matrix itself: `ordering = :nd` (nested dissection, geometric when `coords` is given; the faster factor on the GPU) or
`ordering = :amd` (approximate minimum degree, the kind of ordering `cholesky(A)` gets from CHOLMOD).
"""
function b200_cholesky(A::SparseMatrixCSC{Float64,Int64}; perm::Union{Nothing,Vector{Int64}} = nothing,
                       check::Bool = true, coords::Union{Nothing,Matrix{Float64}} = nothing,
                       ordering::Symbol = :nd, ctx::B200Context = default_context())
    n = size(A, 1)
    cdim = coords === nothing ? Int32(0) : Int32(size(coords, 1))      # coords is dim x n (node-major in memory)
    cptr = coords === nothing ? Ptr{Float64}(C_NULL) : pointer(coords)
    kind = perm !== nothing ? ORDER_GIVEN : ordering === :amd ? ORDER_AMD : ORDER_ND
    opts = Ref(AnalyzeOpts(kind, Int32(0), Int32(1), cdim, cptr, 0, 0, 0.0))
    sym = Ref{Ptr{Cvoid}}(C_NULL)
    fac = Ref{Ptr{Cvoid}}(C_NULL)
    pptr = perm === nothing ? Ptr{Int64}(C_NULL) : pointer(perm)
    GC.@preserve A perm coords begin
        _check(ctx, ccall((:gmrfb_analyze, libgmrfb), Int32,
                         (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{AnalyzeOpts}, Ref{Ptr{Cvoid}}),
                         ctx.h, n, A.colptr, A.rowval, pptr, opts, sym))
        _check(ctx, ccall((:gmrfb_fac_create, libgmrfb), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), sym[], fac))
        p = Vector{Int64}(undef, n)
        _check(ctx, ccall((:gmrfb_sym_get, libgmrfb), Int32,
                         (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                         sym[], p, C_NULL, C_NULL, C_NULL, C_NULL))
        F = B200Factor(ctx, sym[], fac[], n, p, false)
        finalizer(_destroy, F)
        refactorize!(F, A.nzval; check = check)
        return F
    end
end

"Numeric refactorisation on the analysed pattern (the Gauss-Newton loop, scripts/solve_burger.jl:171-180)."
function refactorize!(F::B200Factor, nzval::Vector{Float64}; check::Bool = true)
    st = GC.@preserve nzval ccall((:gmrfb_factorize, libgmrfb), Int32, (Ptr{Cvoid}, Ptr{Float64}), F.fac, nzval)
    F.success = st == GMRFB_OK
    (st == GMRFB_ERR_NOT_SPD && !check) && return F        # `check=false`: failure is a queryable flag
    _check(F.ctx, st)
    return F
end

LinearAlgebra.issuccess(F::B200Factor) = F.success

function _info(F::B200Factor)
    info = Ref(FacInfo(0, 0, 0.0, 0))
    _check(F.ctx, ccall((:gmrfb_fac_get_info, libgmrfb), Int32, (Ptr{Cvoid}, Ref{FacInfo}), F.fac, info))
    return info[]
end
SparseArrays.nnz(F::B200Factor) = Int(_info(F).nnz_L)                  # scripts/darcy/solve_darcy_gmrf-fem.jl:170
logdet_factor(F::B200Factor) = _info(F).logdet
LinearAlgebra.logdet(F::B200Factor) = _info(F).logdet

function _solve!(F::B200Factor, mode::Int32, X::StridedVecOrMat{Float64})
    nrhs = size(X, 2)
    ldx = X isa AbstractVector ? length(X) : stride(X, 2)
    GC.@preserve X _check(F.ctx, ccall((:gmrfb_solve, libgmrfb), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Int64),
                                      F.fac, mode, X, ldx, nrhs))
    return X
end
Base.:\(F::B200Factor, b::StridedVecOrMat{Float64}) = _solve!(F, SOLVE_A, copy(b))     # scripts/solve_burger.jl:148

"`F.PtL \\ b` and `F.UP \\ z` (src/tridiagonal_cholesky.jl:20-22,39-41) as lazy views."
struct FactorView
    F::B200Factor
    mode::Int32
end
Base.:\(V::FactorView, b::StridedVecOrMat{Float64}) = _solve!(V.F, V.mode, copy(b))
function Base.getproperty(F::B200Factor, s::Symbol)
    s === :PtL && return FactorView(F, SOLVE_PTL)
    s === :UP && return FactorView(F, SOLVE_UP)
    s === :L && return _sparse_L(F)                                    # scripts/burgers/solve_burgers_gmrf-collocation.jl:209
    return getfield(F, s)
end

function _sparse_L(F::B200Factor)
    n = getfield(F, :n); fac = getfield(F, :fac); ctx = getfield(F, :ctx)
    colptr = Vector{Int64}(undef, n + 1)
    _check(ctx, ccall((:gmrfb_fac_get_L, libgmrfb), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                     fac, 1, 1, colptr, C_NULL, C_NULL))
    nz = colptr[end] - 1
    rowval = Vector{Int64}(undef, nz); nzval = Vector{Float64}(undef, nz)
    _check(ctx, ccall((:gmrfb_fac_get_L, libgmrfb), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                     fac, 1, 1, colptr, rowval, nzval))
    return SparseMatrixCSC(n, n, colptr, rowval, nzval)
end

"x = mean + P' L^{-T} z for host-supplied standard normals z (`rand(rng, x)`, scripts/darcy/solve_darcy_gmrf-fem.jl:191)."
function sample(F::B200Factor, Z::StridedVecOrMat{Float64}; mean::Union{Nothing,Vector{Float64}} = nothing)
    X = similar(Z)
    mptr = mean === nothing ? Ptr{Float64}(C_NULL) : pointer(mean)
    GC.@preserve Z X mean _check(F.ctx, ccall((:gmrfb_sample, libgmrfb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64),
        F.fac, mptr, Z, size(Z, 1), X, size(X, 1), size(Z, 2)))
    return X
end

"diag(Q^{-1}) by Takahashi selected inversion."
function var_selinv(F::B200Factor)
    v = Vector{Float64}(undef, F.n)
    GC.@preserve v _check(F.ctx, ccall((:gmrfb_var_selinv, libgmrfb), Int32, (Ptr{Cvoid}, Ptr{Float64}), F.fac, v))
    return v
end

"RBMCStrategy(N) variances (scripts/darcy/solve_darcy_gmrf-fem.jl:100,174,192); Z = n x N standard normals."
function var_rbmc(F::B200Factor, Q::SparseMatrixCSC{Float64,Int64}, Z::Matrix{Float64})
    ctx = F.ctx
    spm = Ref{Ptr{Cvoid}}(C_NULL)
    v = Vector{Float64}(undef, F.n)
    GC.@preserve Q Z v begin
        _check(ctx, ccall((:gmrfb_spm_create, libgmrfb), Int32,
                         (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
                         ctx.h, size(Q, 1), size(Q, 2), Q.colptr, Q.rowval, Q.nzval, 1, spm))
        try
            _check(ctx, ccall((:gmrfb_var_rbmc, libgmrfb), Int32,
                             (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}),
                             F.fac, spm[], Z, size(Z, 1), size(Z, 2), v))
        finally
            ccall((:gmrfb_spm_destroy, libgmrfb), Int32, (Ptr{Cvoid},), spm[])
        end
    end
    return v
end

# --------------------------------------------------------------------------------- blueprints (GMRF.jl seam) --
# GaussianMarkovRandomFields.jl constructs its solver from a blueprint via `construct_solver(bp, gmrf)`; a package
# extension adds methods for these two types that build a `B200Factor` (see INTEGRATION.md for the glue).
Base.@kwdef struct B200CholeskySolverBlueprint
    var_strategy::Any = :takahashi          # or RBMCStrategy(N; rng) from GaussianMarkovRandomFields
    perm::Union{Nothing,Vector{Int64}} = nothing
end
struct B200GNCholeskySolverBlueprint
    perm::Union{Nothing,Vector{Int64}}
end
B200GNCholeskySolverBlueprint() = B200GNCholeskySolverBlueprint(nothing)

# --------------------------------------------------------------------------- block-tridiagonal Cholesky --
"Device-resident counterpart of `TridiagonalCholeskyFactor{T}` (src/tridiagonal_cholesky.jl:5-9)."
mutable struct B200TridiagonalCholeskyFactor
    ctx::B200Context
    h::Ptr{Cvoid}
    N::Int          # total rows, as in the reference struct
    b::Int
    nblocks::Int
end

# ---------------------------------------------------------------------------------------------------------------
# Gauss-Newton on the device (gmrfb_gn_*): the loop of scripts/solve_burger.jl:143-180 for a bilinear collocation
# residual f(w) = L w + c (A w).*(D w)  (Burgers, :127-134: L = A1 - A0 - dt*nu*D2, A = A1, c = dt).
mutable struct B200GaussNewton
    ctx::B200Context
    h::Ptr{Cvoid}
    Q::Ptr{Cvoid}        # device copy of the prior precision (owned)
    n::Int
    obj_history::Vector{Float64}
    n_steps::Int
end

function _union_pattern(Ms::SparseMatrixCSC{Float64,Int64}...)
    U = sum(M -> SparseMatrixCSC(M.m, M.n, M.colptr, M.rowval, ones(length(M.nzval))), Ms)  # structural union
    vals = map(Ms) do M
        v = zeros(nnz(U))
        for j in 1:M.n, p in nzrange(M, j)
            q = searchsortedfirst(view(U.rowval, nzrange(U, j)), M.rowval[p]) + first(nzrange(U, j)) - 1
            v[q] = M.nzval[p]
        end
        v
    end
    return U, vals
end

function b200_gauss_newton(mu::Vector{Float64}, Q::SparseMatrixCSC{Float64,Int64}, L::SparseMatrixCSC{Float64,Int64},
                           A::SparseMatrixCSC{Float64,Int64}, D::SparseMatrixCSC{Float64,Int64}, c::Real, noise::Real,
                           y::Vector{Float64}; perm::Union{Nothing,Vector{Int64}} = nothing,
                           ctx::B200Context = default_context())
    U, (lv, av, dv) = _union_pattern(L, A, D)
    qh = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Q _check(ctx, ccall((:gmrfb_spm_create, libgmrfb), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
        ctx.h, size(Q, 1), size(Q, 2), Q.colptr, Q.rowval, Q.nzval, 1, qh))
    opts = Ref(AnalyzeOpts(perm === nothing ? ORDER_ND : ORDER_GIVEN, 0, 1, 0, C_NULL, 0, 0, 0.0))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve U lv av dv y mu perm _check(ctx, ccall((:gmrfb_gn_create, libgmrfb), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32,
         Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ref{AnalyzeOpts}, Ref{Ptr{Cvoid}}),
        ctx.h, qh[], size(U, 1), U.colptr, U.rowval, lv, av, dv, C_NULL #= cubic term e, if any =#, 1, Float64(c),
        Float64(noise), y, mu,
        perm === nothing ? C_NULL : pointer(perm), opts, out))
    gn = B200GaussNewton(ctx, out[], qh[], size(Q, 1), Float64[], 0)
    finalizer(gn) do g
        ccall((:gmrfb_gn_destroy, libgmrfb), Int32, (Ptr{Cvoid},), g.h)
        ccall((:gmrfb_spm_destroy, libgmrfb), Int32, (Ptr{Cvoid},), g.Q)
    end
    return gn
end

"`optimize(gno)` of scripts/burgers/solve_burgers_gmrf-fem.jl:182 - returns the final iterate."
function optimize!(gn::B200GaussNewton, x0::Vector{Float64}; max_steps::Integer = 20, rel_tol::Real = 1e-4)
    x = copy(x0)
    hist = zeros(max_steps + 1)
    steps = Ref{Int32}(0)
    GC.@preserve x hist _check(gn.ctx, ccall((:gmrfb_gn_optimize, libgmrfb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Ref{Int32}, Ptr{Float64}),
        gn.h, x, Int32(max_steps), Float64(rel_tol), steps, hist))
    gn.n_steps = steps[]
    gn.obj_history = hist[1:steps[]+1]
    return x
end

"`tridiagonal_cholesky(A::SparseMatrixCSC, N_blocks)` (src/tridiagonal_cholesky.jl:65-82)."
function b200_tridiagonal_cholesky(A::SparseMatrixCSC{Float64,Int64}, N_blocks::Integer; ctx::B200Context = default_context())
    out = Ref{Ptr{Cvoid}}(C_NULL)
    st = GC.@preserve A ccall((:gmrfb_btd_factor, libgmrfb), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Int64, Ref{Ptr{Cvoid}}),
        ctx.h, size(A, 1), A.colptr, A.rowval, A.nzval, 1, N_blocks, out)
    if st != GMRFB_OK && out[] != C_NULL
        ccall((:gmrfb_btd_destroy, libgmrfb), Int32, (Ptr{Cvoid},), out[])
    end
    _check(ctx, st)
    F = B200TridiagonalCholeskyFactor(ctx, out[], size(A, 1), size(A, 1) ÷ N_blocks, N_blocks)
    finalizer(f -> ccall((:gmrfb_btd_destroy, libgmrfb), Int32, (Ptr{Cvoid},), f.h), F)
    return F
end

function _block(F::B200TridiagonalCholeskyFactor, i::Integer, which::Integer)
    out = Matrix{Float64}(undef, F.b, F.b)
    GC.@preserve out _check(F.ctx, ccall((:gmrfb_btd_get_block, libgmrfb), Int32,
        (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Int64), F.h, i - 1, which, out, F.b))
    return out
end
"`chos[i].L` and `Cs[i]` of the reference struct, copied out of device memory on demand."
chos(F::B200TridiagonalCholeskyFactor) = [Cholesky(LowerTriangular(_block(F, i, 0))) for i in 1:F.nblocks]
Cs(F::B200TridiagonalCholeskyFactor) = [_block(F, i, 1) for i in 1:F.nblocks-1]

function _btd_solve(F::B200TridiagonalCholeskyFactor, mode::Int32, b::StridedVecOrMat{Float64})
    n = F.b * F.nblocks
    X = copy(b)
    ldx = X isa AbstractVector ? length(X) : stride(X, 2)
    GC.@preserve X _check(F.ctx, ccall((:gmrfb_btd_solve, libgmrfb), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Int64),
                                      F.h, mode, X, ldx, size(X, 2)))
    return X
end
# intended semantics of src/tridiagonal_cholesky.jl:24-63 (flat vectors; defects T4/T5/T7 of SURVEY.md §8a not reproduced)
forward_solve(F::B200TridiagonalCholeskyFactor, b) = _btd_solve(F, BTD_SOLVE_FWD, b)
backward_solve(F::B200TridiagonalCholeskyFactor, b) = _btd_solve(F, BTD_SOLVE_BWD, b)
forward_solve(F::B200Factor, b) = F.PtL \ b          # :39-41
backward_solve(F::B200Factor, b) = F.UP \ b          # :20-22
ldiv(F::B200TridiagonalCholeskyFactor, b) = _btd_solve(F, BTD_SOLVE_A, b)
# extends LinearAlgebra.ldiv! (a `using`-visible name that is not imported must be qualified to receive methods)
LinearAlgebra.ldiv!(y, F::B200TridiagonalCholeskyFactor, b) = (y .= ldiv(F, b); y)

function LinearAlgebra.logdet(F::B200TridiagonalCholeskyFactor)
    out = Ref(0.0)
    _check(F.ctx, ccall((:gmrfb_btd_logdet, libgmrfb), Int32, (Ptr{Cvoid}, Ref{Float64}), F.h, out))
    return out[]
end

# ------------------------------------------------------------- finite-element assembly on the device --
# The matrices the reference rebuilds inside its hot loops (SURVEY.md section 8(f) N2/N3).  Results are device matrices
# owned by the assembler handle (fixed pattern, values of the last call); `to_sparse` copies one back.
"Borrowed view of a device CSC matrix (gmrfb_spm) owned by `owner`."
struct B200SparseMatrix
    ctx::B200Context
    h::Ptr{Cvoid}
    owner::Any
end

function Base.size(M::B200SparseMatrix)
    m, n, z = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(M.ctx, ccall((:gmrfb_spm_dims, libgmrfb), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), M.h, m, n, z))
    return (Int(m[]), Int(n[]))
end

function to_sparse(M::B200SparseMatrix)
    m, n, z = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    _check(M.ctx, ccall((:gmrfb_spm_dims, libgmrfb), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), M.h, m, n, z))
    colptr, rowval, nzval = Vector{Int64}(undef, n[] + 1), Vector{Int64}(undef, z[]), Vector{Float64}(undef, z[])
    GC.@preserve colptr rowval nzval _check(M.ctx, ccall((:gmrfb_spm_get, libgmrfb), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), M.h, 1, colptr, rowval, nzval))
    return SparseMatrixCSC(Int(m[]), Int(n[]), colptr, rowval, nzval)
end

_mask(p::Nothing, n) = (C_NULL, nothing)
function _mask(p, n)
    m = zeros(UInt8, n)
    m[collect(p)] .= 0x01          # a collection of 1-based dof indices (ch.prescribed_dofs)
    return (pointer(m), m)
end

"P1 triangles: `nodes` 2 x n coordinates, `tris` 3 x T vertex indices (1-based)."
mutable struct B200FEMP1
    ctx::B200Context
    h::Ptr{Cvoid}
    n::Int
end

function B200FEMP1(nodes::Matrix{Float64}, tris::Matrix{Int64}; ctx::B200Context = default_context())
    size(nodes, 1) == 2 && size(tris, 1) == 3 || throw(ArgumentError("nodes must be 2 x n and tris 3 x T"))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve nodes tris _check(ctx, ccall((:gmrfb_fem_create, libgmrfb), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Int32, Ref{Ptr{Cvoid}}),
        ctx.h, size(nodes, 2), nodes, size(tris, 2), tris, 1, out))
    F = B200FEMP1(ctx, out[], size(nodes, 2))
    finalizer(f -> ccall((:gmrfb_fem_destroy, libgmrfb), Int32, (Ptr{Cvoid},), f.h), F)
    return F
end

"Lumped mass vector (load vector of f = 1)."
function lumped_mass(F::B200FEMP1)
    m = Vector{Float64}(undef, F.n)
    GC.@preserve m _check(F.ctx, ccall((:gmrfb_fem_get_mass, libgmrfb), Int32, (Ptr{Cvoid}, Ptr{Float64}), F.h, m))
    return m
end

function set_coeff_grid!(F::B200FEMP1, x_coords::Vector{Float64}, y_coords::Vector{Float64})
    GC.@preserve x_coords y_coords _check(F.ctx, ccall((:gmrfb_fem_set_coeff_grid, libgmrfb), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}), F.h, length(x_coords), x_coords, length(y_coords), y_coords))
    return F
end

"assemble_darcy_diff_matrix (src/problems/darcy.jl:5-63): `coeff_mat[ix, iy]` as in the reference, or nothing for a unit coefficient."
function assemble_darcy(F::B200FEMP1, coeff_mat::Union{Nothing,Matrix{Float64}} = nothing; prescribed = nothing)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    pm, keep = _mask(prescribed, F.n)
    cp = coeff_mat === nothing ? Ptr{Float64}(C_NULL) : pointer(coeff_mat)
    GC.@preserve coeff_mat keep _check(F.ctx, ccall((:gmrfb_fem_assemble, libgmrfb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Ref{Ptr{Cvoid}}), F.h, cp, pm, out))
    return B200SparseMatrix(F.ctx, out[], F)
end

"ratio * K' Mt^-1 K, K = kappa^2 Mt + G (src/spdes/shallow_water.jl:177-194)."
function matern_precision(F::B200FEMP1, kappa::Real, ratio::Real; prescribed = nothing, prescribed_mass::Real = 1e-2)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    pm, keep = _mask(prescribed, F.n)
    GC.@preserve keep _check(F.ctx, ccall((:gmrfb_fem_matern_precision, libgmrfb), Int32,
        (Ptr{Cvoid}, Float64, Float64, Ptr{UInt8}, Float64, Ref{Ptr{Cvoid}}),
        F.h, Float64(kappa), Float64(ratio), pm, Float64(prescribed_mass), out))
    return B200SparseMatrix(F.ctx, out[], F)
end

"f_and_J of _research/elliptic_chen24.jl:280-285 without the static load vector: (s G u + f_cube, s G + J_cube)."
function assemble_cubic(F::B200FEMP1, u::Vector{Float64}; prescribed = nothing, quad_degree::Integer = 2,
                        stiffness_scale::Real = 1.0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    f = Vector{Float64}(undef, F.n)
    pm, keep = _mask(prescribed, F.n)
    GC.@preserve u f keep _check(F.ctx, ccall((:gmrfb_fem_assemble_cubic, libgmrfb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Ptr{UInt8}, Ref{Ptr{Cvoid}}, Ptr{Float64}),
        F.h, u, Int32(quad_degree), Float64(stiffness_scale), pm, out, f))
    return f, B200SparseMatrix(F.ctx, out[], F)
end

"""
Lagrange triangles of order 1 or 2 with Ferrite's cell values (isoparametric geometry, `QuadratureRule{RefTriangle}(order + 1)`
unless `quad_degree` says otherwise): the discretisations of src/utils.jl:20-38 and _research/elliptic_chen24.jl:118-122.
`nodes` 2 x n, `elems` 3 x E or 6 x E node indices (1-based; six-node cells numbered as Ferrite's QuadraticTriangle).
"""
mutable struct B200FEMLagrange
    ctx::B200Context
    h::Ptr{Cvoid}
    n::Int
    order::Int
end

function B200FEMLagrange(nodes::Matrix{Float64}, elems::Matrix{Int64}; order::Integer = size(elems, 1) == 3 ? 1 : 2,
                         quad_degree::Integer = 0, ctx::B200Context = default_context())
    size(nodes, 1) == 2 && size(elems, 1) == (order == 1 ? 3 : 6) || throw(ArgumentError("nodes must be 2 x n, elems 3 x E or 6 x E"))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve nodes elems _check(ctx, ccall((:gmrfb_fem2d_create, libgmrfb), Int32,
        (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Int32, Int32, Ref{Ptr{Cvoid}}),
        ctx.h, Int32(order), size(nodes, 2), nodes, size(elems, 2), elems, 1, Int32(quad_degree), out))
    F = B200FEMLagrange(ctx, out[], size(nodes, 2), order)
    finalizer(f -> ccall((:gmrfb_fem2d_destroy, libgmrfb), Int32, (Ptr{Cvoid},), f.h), F)
    return F
end

function set_coeff_grid!(F::B200FEMLagrange, x_coords::Vector{Float64}, y_coords::Vector{Float64})
    GC.@preserve x_coords y_coords _check(F.ctx, ccall((:gmrfb_fem2d_set_coeff_grid, libgmrfb), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}), F.h, length(x_coords), x_coords, length(y_coords), y_coords))
    return F
end

"assemble_darcy_diff_matrix (src/problems/darcy.jl:5-63): (G, f); the coefficient `coeff_mat[ix, iy]` is looked up at every quadrature point."
function assemble_darcy(F::B200FEMLagrange, coeff_mat::Union{Nothing,Matrix{Float64}} = nothing; prescribed = nothing,
                        beta::Real = 1.0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    f = Vector{Float64}(undef, F.n)
    pm, keep = _mask(prescribed, F.n)
    cp = coeff_mat === nothing ? Ptr{Float64}(C_NULL) : pointer(coeff_mat)
    GC.@preserve coeff_mat keep f _check(F.ctx, ccall((:gmrfb_fem2d_stiffness, libgmrfb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Float64, Ref{Ptr{Cvoid}}, Ptr{Float64}), F.h, cp, pm, Float64(beta), out, f))
    return B200SparseMatrix(F.ctx, out[], F), f
end

"Mass matrix; lumping 0 consistent, 1 row sums, 2 scaled element diagonals, 3 = `lump_matrix(me, ip)` for this order: (M, lumped vector)."
function assemble_mass(F::B200FEMLagrange; lumping::Integer = 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    ml = Vector{Float64}(undef, F.n)
    GC.@preserve ml _check(F.ctx, ccall((:gmrfb_fem2d_mass, libgmrfb), Int32,
        (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}, Ptr{Float64}), F.h, Int32(lumping), out, ml))
    return B200SparseMatrix(F.ctx, out[], F), (lumping == 0 ? nothing : ml)
end

"ratio K' Mt^-1 K (alpha = 2) or ratio K Mt^-1 K Mt^-1 K (alpha = 3), K = kappa^2 Mt + G (src/spdes/shallow_water.jl:172-190)."
function matern_precision(F::B200FEMLagrange, kappa::Real, ratio::Real; alpha::Integer = 2, prescribed = nothing,
                          prescribed_mass::Real = 1e-2)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    pm, keep = _mask(prescribed, F.n)
    GC.@preserve keep _check(F.ctx, ccall((:gmrfb_fem2d_matern_precision, libgmrfb), Int32,
        (Ptr{Cvoid}, Float64, Float64, Int32, Ptr{UInt8}, Float64, Ref{Ptr{Cvoid}}),
        F.h, Float64(kappa), Float64(ratio), Int32(alpha), pm, Float64(prescribed_mass), out))
    return B200SparseMatrix(F.ctx, out[], F)
end

"f_and_J of _research/elliptic_chen24.jl:280-285 without the static load vector, with the rule of the handle: (f, J)."
function assemble_cubic(F::B200FEMLagrange, u::Vector{Float64}; prescribed = nothing, stiffness_scale::Real = 1.0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    f = Vector{Float64}(undef, F.n)
    pm, keep = _mask(prescribed, F.n)
    GC.@preserve u f keep _check(F.ctx, ccall((:gmrfb_fem2d_assemble_cubic, libgmrfb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Float64, Ptr{UInt8}, Ref{Ptr{Cvoid}}, Ptr{Float64}),
        F.h, u, Float64(stiffness_scale), pm, out, f))
    return f, B200SparseMatrix(F.ctx, out[], F)
end

"Fixed-pattern product plan  C = alpha A diag(w) B  (gmrfb_spgemm): symbolic once, `product!(plan; alpha, w)` per call."
mutable struct B200SparseProduct
    ctx::B200Context
    h::Ptr{Cvoid}
    A::B200SparseMatrix
    B::B200SparseMatrix
end

function B200SparseProduct(A::B200SparseMatrix, B::B200SparseMatrix)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _check(A.ctx, ccall((:gmrfb_spgemm_create, libgmrfb), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
        A.ctx.h, A.h, B.h, out))
    P = B200SparseProduct(A.ctx, out[], A, B)
    finalizer(p -> ccall((:gmrfb_spgemm_destroy, libgmrfb), Int32, (Ptr{Cvoid},), p.h), P)
    return P
end

function product!(P::B200SparseProduct; alpha::Real = 1.0, w::Union{Nothing,Vector{Float64}} = nothing)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    wp = w === nothing ? Ptr{Float64}(C_NULL) : pointer(w)
    GC.@preserve w _check(P.ctx, ccall((:gmrfb_spgemm_compute, libgmrfb), Int32,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ref{Ptr{Cvoid}}), P.h, Float64(alpha), wp, out))
    return B200SparseMatrix(P.ctx, out[], P)
end

"Lagrange lines of order 1 or 2: `elems` (order+1) x E node indices (1-based; quadratic: left, right, middle), `elem_x` their coordinates."
mutable struct B200FEM1D
    ctx::B200Context
    h::Ptr{Cvoid}
    n::Int
end

function B200FEM1D(elems::Matrix{Int64}, elem_x::Matrix{Float64}; order::Integer = size(elems, 1) - 1, nquad::Integer = 0,
                   ctx::B200Context = default_context())
    size(elems) == size(elem_x) && size(elems, 1) == order + 1 || throw(ArgumentError("elems and elem_x must be (order+1) x E"))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    n = Int(maximum(elems))
    GC.@preserve elems elem_x _check(ctx, ccall((:gmrfb_fem1d_create, libgmrfb), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
        ctx.h, n, size(elems, 2), elems, elem_x, Int32(order), 1, Int32(nquad), out))
    F = B200FEM1D(ctx, out[], n)
    finalizer(f -> ccall((:gmrfb_fem1d_destroy, libgmrfb), Int32, (Ptr{Cvoid},), f.h), F)
    return F
end

"assemble_burgers_mass_diffusion_matrices(disc; lumping) (src/problems/burgers.jl:61-98)."
function assemble_burgers_mass_diffusion_matrices(F::B200FEM1D; lumping::Bool = false, prescribed = nothing)
    M, G = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
    pm, keep = _mask(prescribed, F.n)
    GC.@preserve keep _check(F.ctx, ccall((:gmrfb_fem1d_mass_stiffness, libgmrfb), Int32,
        (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}), F.h, Int32(lumping), pm, M, G))
    return B200SparseMatrix(F.ctx, M[], F), B200SparseMatrix(F.ctx, G[], F)
end

"assemble_burgers_advection_matrix(disc, cur_weights) (src/problems/burgers.jl:5-59): (G_adv, v)."
function assemble_burgers_advection_matrix(F::B200FEM1D, u::Vector{Float64}; prescribed = nothing)
    A = Ref{Ptr{Cvoid}}(C_NULL)
    v = Vector{Float64}(undef, F.n)
    pm, keep = _mask(prescribed, F.n)
    GC.@preserve u v keep _check(F.ctx, ccall((:gmrfb_fem1d_advection, libgmrfb), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Ref{Ptr{Cvoid}}, Ptr{Float64}), F.h, u, pm, A, v))
    return B200SparseMatrix(F.ctx, A[], F), v
end

"f_and_J(w) of scripts/burgers/solve_burgers_gmrf-fem.jl:115-142 for `nt` time steps (w time-major): (f, J)."
function burgers_f_and_J(F::B200FEM1D, w::Vector{Float64}, nt::Integer, dt::Real, nu::Real; prescribed = nothing)
    length(w) == nt * F.n || throw(DimensionMismatch("w must hold nt * n values"))
    J = Ref{Ptr{Cvoid}}(C_NULL)
    f = Vector{Float64}(undef, (nt - 1) * F.n)
    pm, keep = _mask(prescribed, F.n)
    GC.@preserve w f keep _check(F.ctx, ccall((:gmrfb_fem1d_spacetime_tangent, libgmrfb), Int32,
        (Ptr{Cvoid}, Int64, Float64, Float64, Ptr{Float64}, Ptr{UInt8}, Ref{Ptr{Cvoid}}, Ptr{Float64}),
        F.h, Int64(nt), Float64(dt), Float64(nu), w, pm, J, f))
    return f, B200SparseMatrix(F.ctx, J[], F)
end

end # module
