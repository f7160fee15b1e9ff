/*
 * gmrfb.h — C ABI of libgmrfb, the B200-native (sm_100a) precision-matrix linear algebra
 * behind Gaussian-Markov-random-field PDE solvers.
 *
 * This is the drop-in boundary for the hot path of timweiland/DiffEqGMRFs.jl: every entry
 * point below names the reference call (file:line under /root/reference) it replaces.  The
 * reference has no FFI of its own (it is pure Julia on top of SparseArrays.CHOLMOD and
 * GaussianMarkovRandomFields.jl); a Julia `ccall` shim (julia/GMRFB200.jl) or the
 * Python ctypes host in `diffeqgmrfs.jl_b200/` binds exactly these symbols.
 *
 * Conventions
 *  - Every function returns a gmrfb_status (0 = ok).  Nothing throws or aborts across the ABI.
 *    `gmrfb_last_error(ctx)` returns a human-readable message for the last failure on that context.
 *  - All host arrays are caller-owned and only read/written during the call (calls are synchronous:
 *    when a call returns, its host outputs are complete).  All device memory is owned by the library
 *    behind opaque handles released by the matching *_destroy.
 *  - Sparse matrices are CSC with 64-bit indices (Julia `SparseMatrixCSC{Float64,Int64}`):
 *    colptr[n+1], rowval[nnz] sorted within a column, nzval[nnz]; `base` is 0 or 1 and applies to
 *    colptr, rowval and permutations alike, so Julia passes its arrays untouched.
 *  - Dense matrices are column-major Float64 with an explicit leading dimension.
 *  - There is no CPU fallback: every numeric entry point runs CUDA kernels on the context's device
 *    and fails with GMRFB_ERR_CUDA if that is impossible.
 *  - A context and every handle created from it (they share its stream, its buffer bookkeeping and the caches of
 *    captured CUDA graphs kept by the symbolic handles) are used by one host thread at a time; distinct contexts may
 *    be used concurrently (bench.py: one context per problem in flight).
 */
#ifndef GMRFB_H
#define GMRFB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t gmrfb_status;
enum {
  GMRFB_OK = 0,
  GMRFB_ERR_INVALID = 1,  /* bad argument (null pointer, size mismatch, unsorted/duplicate indices, bad perm) */
  GMRFB_ERR_NOT_SPD = 2,  /* a pivot was <= 0 or NaN; gmrfb_factor_info reports the failing column */
  GMRFB_ERR_ALLOC = 3,    /* host or device allocation failed */
  GMRFB_ERR_CUDA = 4,     /* CUDA runtime error, or no usable device */
  GMRFB_ERR_COMM = 5,     /* multi-GPU exchange failed */
  GMRFB_ERR_STATE = 6,    /* handle used in the wrong state (e.g. solve before a successful factorize) */
  GMRFB_ERR_INTERNAL = 7  /* an unexpected C++ exception was caught at the boundary (message: gmrfb_last_error) */
};

typedef struct gmrfb_ctx gmrfb_ctx; /* device + stream + scratch              */
typedef struct gmrfb_sym gmrfb_sym; /* symbolic analysis of one sparsity pattern */
typedef struct gmrfb_fac gmrfb_fac; /* numeric supernodal factor L (Q = P' L L' P) */
typedef struct gmrfb_btd gmrfb_btd; /* dense block-tridiagonal factor          */
typedef struct gmrfb_spm gmrfb_spm; /* device-resident sparse matrix (CSC)      */

/* ---------------------------------------------------------------- context ---- */

/* Library version (major*10000 + minor*100 + patch). */
int32_t gmrfb_version(void);

/* Create a context bound to CUDA device `device` (its own stream).  Fails loudly without a GPU. */
gmrfb_status gmrfb_ctx_create(int32_t device, gmrfb_ctx** out);
gmrfb_status gmrfb_ctx_destroy(gmrfb_ctx* ctx);
/* Message of the last error raised through `ctx` ("" if none).  Valid until the next call on ctx.
 * With ctx == NULL returns the last context-free error (e.g. from a failed gmrfb_ctx_create). */
const char* gmrfb_last_error(gmrfb_ctx* ctx);
/* Block until all work queued on the context's stream has finished. */
gmrfb_status gmrfb_ctx_sync(gmrfb_ctx* ctx);
/* The context's cudaStream_t, as an integer, so a host (torch, CUDA.jl) can order its own work
 * (event timing, NCCL exchanges) against the library's kernels. */
uint64_t gmrfb_ctx_stream(gmrfb_ctx* ctx);
/* Number of kernels this library has launched through `ctx` since creation (for benchmarks). */
int64_t gmrfb_ctx_launch_count(gmrfb_ctx* ctx);

/* The library caches released device buffers of >= 1 MB (a dataset loop that creates and destroys multi-GB factor handles
 * would otherwise pay cudaMalloc/cudaFree every iteration).  gmrfb_pool_trim frees cached buffers until at most
 * keep_bytes remain cached (0: everything) and reports what is still cached; destroying the last context trims the
 * cache completely, so a co-resident allocator (torch, CUDA.jl) gets the memory back.  GMRFB_POOL=0 disables the
 * cache, GMRFB_POOL_MAX_GB bounds it (default 24). */
gmrfb_status gmrfb_pool_trim(int64_t keep_bytes, int64_t* cached_bytes_out);

/* Per-kernel profiling for benchmarks: between profile_begin and profile_end every kernel launched through
 * `ctx` is bracketed by CUDA events on the context's stream; profile_end synchronises and returns one entry
 * per kernel kind with its launch count, summed device time and the algorithmic flops / bytes of those launches
 * (the numerators of the roofline; DESIGN.md states the per-unit figures). */
typedef struct gmrfb_profile_entry {
  int32_t kind;
  int64_t launches;
  double ms;
  double flops;
  double bytes;
  char name[32];
} gmrfb_profile_entry;
gmrfb_status gmrfb_ctx_profile_begin(gmrfb_ctx* ctx);
gmrfb_status gmrfb_ctx_profile_end(gmrfb_ctx* ctx, gmrfb_profile_entry* entries, int32_t cap, int32_t* count);

/* ------------------------------------------------------ symbolic analysis ---- */
/* Replaces the symbolic half of `cholesky(Symmetric(A); perm=p)` — CHOLMOD analyze / analyze_p:
 *   scripts/solve_burger.jl:147, scripts/darcy/solve_darcy_fem.jl:93 and every
 *   CholeskySolverBlueprint(perm=p) (scripts/darcy/solve_darcy_gmrf-fem.jl:100,174). */

enum { /* ordering_kind */
  GMRFB_ORDER_GIVEN = 0,   /* use `perm` exactly (Julia `perm=p`): the factor's .p equals perm */
  GMRFB_ORDER_NATURAL = 1, /* identity */
  GMRFB_ORDER_ND = 2,      /* library's nested dissection (graph BFS bisection; coordinates if supplied) */
  GMRFB_ORDER_AMD = 3,     /* approximate minimum degree on the quotient graph (Amestoy-Davis-Duff): what the
                              reference's `cholesky(A)` without `perm` asks CHOLMOD for; ties are this library's */
  GMRFB_ORDER_ND_AMD = 4   /* nested dissection down to leaves of nd_leaf (default 200) vertices, each leaf ordered by
                              halo-AMD (minimum degree aware of the surrounding separators)                          */
};
enum { /* storage */
  GMRFB_STORAGE_FULL = 0,  /* both triangles stored (Julia Symmetric(sparse) of an assembled matrix) */
  GMRFB_STORAGE_LOWER = 1, /* only entries with row >= col are stored                                */
  GMRFB_STORAGE_UPPER = 2  /* only entries with row <= col are stored                                */
};

typedef struct gmrfb_analyze_opts {
  int32_t ordering_kind;   /* GMRFB_ORDER_*                                                   */
  int32_t storage;         /* GMRFB_STORAGE_*                                                 */
  int32_t base;            /* 0 or 1: index base of colptr/rowval/perm                        */
  int32_t coord_dim;       /* 0, or 2/3 when `coords` is given (ND then bisects geometrically) */
  const double* coords;    /* optional n*coord_dim node coordinates, node-major; may be NULL  */
  int32_t nd_leaf;         /* ND leaf size (0 = default)                                      */
  int32_t relax_small;     /* supernode amalgamation: always merge when merged width <= this (0 = default) */
  double relax_zeros;      /* ... or when the fraction of explicit zeros stays below this (0 = default)   */
} gmrfb_analyze_opts;

/* Analyse the pattern of a symmetric n-by-n matrix.  `perm` (length n, `base`-based, new->old as in
 * Julia/CHOLMOD: row k of the permuted matrix is row perm[k] of A) is required for ORDER_GIVEN and
 * ignored otherwise.  Produces: the fill-reducing permutation, elimination tree, column counts,
 * supernode partition and every index map the numeric phases need (uploaded to the device once).
 * The data-parallel phases run as CUDA kernels on the context's device; the nested dissection runs on host worker
 * threads (environment: GMRFB_ND_THREADS, default the hardware threads, at most 16; the permutation does not depend on
 * it).  The host-side plan builders of gmrfb_postprec_create / gmrfb_spgemm_create use GMRFB_HOST_THREADS likewise. */
gmrfb_status gmrfb_analyze(gmrfb_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval,
                           const int64_t* perm, const gmrfb_analyze_opts* opts, gmrfb_sym** out);
gmrfb_status gmrfb_sym_destroy(gmrfb_sym* sym);

typedef struct gmrfb_sym_info {
  int64_t n;
  int64_t nnz_lower_A;  /* stored entries of tril(P A P')                                 */
  int64_t nnz_L;        /* Σ_j colcount[j]  (exact factor, before supernode relaxation)    */
  int64_t nnz_L_stored; /* entries of the supernodal storage incl. explicit zeros          */
  double flops;         /* Σ_j colcount[j]^2 (CHOLMOD's `fl` convention)                   */
  int64_t nsuper;       /* number of supernodes                                            */
  int64_t nlevels;      /* height of the supernodal elimination tree                       */
  int64_t max_front;    /* largest frontal matrix order                                    */
  int64_t front_bytes;  /* device bytes of the frontal arena (factor + update matrices)    */
} gmrfb_sym_info;
gmrfb_status gmrfb_sym_get_info(const gmrfb_sym* sym, gmrfb_sym_info* info);

/* Copy out symbolic results.  Any pointer may be NULL.  All index outputs are `base`-based.
 *   perm[n]      new->old, the `.p` field of a CHOLMOD factor (scripts/darcy/solve_darcy_gmrf-fem.jl:169)
 *   parent[n]    elimination tree of P A P' in that ordering (root: base-1, i.e. -1 / 0)
 *   colcount[n]  nnz of each column of L incl. the diagonal
 *   super_ptr[nsuper+1]  first column (in the library's internal postordered numbering) of each supernode
 *   ipost[n]     internal position of column k of the `perm` ordering (an etree postorder; identity-like
 *                relabelling that leaves L unchanged up to symmetric permutation) */
gmrfb_status gmrfb_sym_get(const gmrfb_sym* sym, int64_t* perm, int64_t* parent, int64_t* colcount,
                           int64_t* super_ptr, int64_t* ipost);
/* Row structure of supernode `s` (internal numbering, `base`-based): writes up to `cap` indices,
 * returns the full count in *nrows.  Columns of the supernode come first. */
gmrfb_status gmrfb_sym_get_super_rows(const gmrfb_sym* sym, int64_t s, int64_t* rows, int64_t cap,
                                      int64_t* nrows);

/* Index maps of the analysis (diagnostic / test entry point: the GPU and the host implementation of the analysis are
 * compared bit for bit through it).  amap[nnz]: arena slot of every stored matrix entry (-1: mirrored triangle);
 * relmap[total_rows]: for every below-row of every front its position in the parent's front (-1 for a front's own
 * columns and for roots); total_rows = sum of the front orders.  Call with NULL pointers to get the sizes. */
gmrfb_status gmrfb_sym_get_maps(const gmrfb_sym* sym, int64_t* amap, int64_t* relmap, int64_t* nnz_out,
                                int64_t* total_rows_out);

/* ------------------------------------------------- numeric factorisation ---- */
/* Replaces the numeric half of `cholesky(A; perm, check=false)` (scripts/solve_burger.jl:147) and the
 * factorisation inside condition_on_observations / GaussNewtonOptimizer for CholeskySolverBlueprint,
 * GNCholeskySolverBlueprint (scripts/darcy/solve_darcy_gmrf-fem.jl:188, scripts/burgers/solve_burgers_gmrf-fem.jl:170-182). */

/* Allocate the factor storage for `sym` (no values yet). */
gmrfb_status gmrfb_fac_create(gmrfb_sym* sym, gmrfb_fac** out);
gmrfb_status gmrfb_fac_destroy(gmrfb_fac* fac);
/* Numeric Cholesky from host values: nzval[nnz] in the order of the analysed colptr/rowval.
 * Re-callable with new values on the same pattern (the Gauss-Newton loop, scripts/solve_burger.jl:171-180).
 * Returns GMRFB_ERR_NOT_SPD (Julia: `check=false` leaves a queryable flag) if a pivot fails. */
gmrfb_status gmrfb_factorize(gmrfb_fac* fac, const double* nzval);
/* Same, values already on the device (device pointer, same order). */
gmrfb_status gmrfb_factorize_dev(gmrfb_fac* fac, const double* d_nzval);

typedef struct gmrfb_fac_info {
  int32_t status;       /* GMRFB_OK or GMRFB_ERR_NOT_SPD for the last factorisation          */
  int64_t fail_column;  /* first failing column (in `perm` order, 0-based) or -1              */
  double logdet;        /* log det Q = 2 Σ log L_jj                                           */
  int64_t nnz_L;        /* as `nnz(F)` (scripts/darcy/solve_darcy_gmrf-fem.jl:170): stored factor entries */
} gmrfb_fac_info;
gmrfb_status gmrfb_fac_get_info(gmrfb_fac* fac, gmrfb_fac_info* info);
/* diag(L) in the `perm` ordering — `diag(sparse(F.L))` (scripts/burgers/solve_burgers_gmrf-collocation.jl:209). */
gmrfb_status gmrfb_fac_diag(gmrfb_fac* fac, double* diagL);
/* L as CSC (lower, `perm` ordering, `base`-based, sorted rows, explicit supernodal zeros dropped when
 * drop_zeros != 0) — `sparse(F.L)`.  Call with colptr only to size the arrays (colptr[n] - base = nnz). */
gmrfb_status gmrfb_fac_get_L(gmrfb_fac* fac, int32_t base, int32_t drop_zeros, int64_t* colptr,
                             int64_t* rowval, double* nzval);

/* ------------------------------------------------------------------ solves --- */
/* X (n-by-nrhs, column-major, leading dimension ldx, host memory) is overwritten in place.
 * Modes follow the CHOLMOD factor views the reference uses (src/tridiagonal_cholesky.jl:20-22,39-41):
 *   A    : X <- Q^{-1} X            `F \ b`        (scripts/solve_burger.jl:148; posterior mean)
 *   PtL  : X <- L^{-1} P X          `F.PtL \ b`    (forward_solve)
 *   UP   : X <- P' L^{-T} X         `F.UP \ z`     (backward_solve; a N(0, Q^{-1}) sample for z ~ N(0,I))
 *   L, Lt: X <- L^{-1} X, L^{-T} X in the permuted ordering (no P).
 * Batches of up to 4 right-hand sides run the level-scheduled kernels; for supernodes with >= 128 columns these apply
 * the full inverse of the diagonal block that the factorisation kept (two bandwidth-bound products per supernode
 * instead of a chain of 64-column substitution steps) whenever the factorisation measured cond_1 of every such block
 * below 1e5 (environment: GMRFB_WIDE_COND_MAX; GMRFB_WIDE_INV=0 disables the inverses, GMRFB_WIDE_MIN the width).
 * Larger batches are swept as panels of up to 64 columns. */
enum { GMRFB_SOLVE_A = 0, GMRFB_SOLVE_PTL = 1, GMRFB_SOLVE_UP = 2, GMRFB_SOLVE_L = 3, GMRFB_SOLVE_LT = 4 };
gmrfb_status gmrfb_solve(gmrfb_fac* fac, int32_t mode, double* X, int64_t ldx, int64_t nrhs);
/* Device-pointer variant: d_X is n-by-nrhs column-major on the context's device. */
gmrfb_status gmrfb_solve_dev(gmrfb_fac* fac, int32_t mode, double* d_X, int64_t ldx, int64_t nrhs);

/* `F \ b` with iterative refinement against the matrix the factor was computed from:
 *   x_0 = Q^{-1} b (factor solve);  repeat up to max_iter times:  r = b - Q x,  x <- x + Q^{-1} r
 * (stops early when the relative residual ||r|| / ||b|| no longer decreases; the iterate with the smallest residual is
 * returned).  For precisions with a large conditioning term such as Q_eps = 1e8 (scripts/solve_burger.jl:98,140) one
 * step brings the residual of the solve down to that of a backward-stable substitution.  Q: the precision as a device
 * matrix (full symmetric storage); resid_out (nrhs doubles, may be NULL) receives the final relative residuals. */
gmrfb_status gmrfb_solve_refined(gmrfb_fac* fac, const gmrfb_spm* Q, double* X, int64_t ldx, int64_t nrhs,
                                 int32_t max_iter, double* resid_out);

/* Samples  X[:,k] = mean + P' L^{-T} Z[:,k]  — `rand(rng, x)` (scripts/darcy/solve_darcy_gmrf-fem.jl:191).
 * The host supplies the standard normals Z (n-by-nrhs) so "same seed" means "same z"; mean may be NULL. */
gmrfb_status gmrfb_sample(gmrfb_fac* fac, const double* mean, const double* Z, int64_t ldz, double* X,
                          int64_t ldx, int64_t nrhs);

/* ------------------------------------------------------ marginal variances --- */
/* diag(Q^{-1}) by Takahashi selected inversion on the supernodal factor (north-star capability;
 * GMRF.jl `var`/`std` with the default strategy).  var_out[n] in the original ordering. */
gmrfb_status gmrfb_var_selinv(gmrfb_fac* fac, double* var_out);
gmrfb_status gmrfb_var_selinv_dev(gmrfb_fac* fac, double* d_var_out);
/* Rao-Blackwellised Monte-Carlo variances, RBMCStrategy(N) (scripts/darcy/solve_darcy_gmrf-fem.jl:100,174,192):
 *   var_i = 1/Q_ii + mean_k ( Σ_{j != i} Q_ij x_j^(k) )^2 / Q_ii^2,   x^(k) = P' L^{-T} z^(k).
 * Q is the precision the factor was computed from; Z is n-by-nsamp standard normals supplied by the caller (host or
 * device memory: the caller owns the random stream, as the reference threads its MersenneTwister through). */
gmrfb_status gmrfb_var_rbmc(gmrfb_fac* fac, const gmrfb_spm* Q, const double* Z, int64_t ldz,
                            int64_t nsamp, double* var_out);
/* Same with the result left in device memory (n doubles; queued on the context's stream, not synchronised): the
 * sample-sharded estimate of several GPUs is then combined by one all-reduce without a host round trip. */
gmrfb_status gmrfb_var_rbmc_dev(gmrfb_fac* fac, const gmrfb_spm* Q, const double* Z, int64_t ldz, int64_t nsamp,
                                double* d_var_out);
/* Selected entries of Q^{-1}: for each k, out[k] = (Q^{-1})[rows[k], cols[k]] (`base`-based indices in
 * the original ordering).  Entries outside the filled pattern of L + L' yield GMRFB_ERR_INVALID. */
gmrfb_status gmrfb_selinv_entries(gmrfb_fac* fac, int32_t base, int64_t count, const int64_t* rows,
                                  const int64_t* cols, double* out);

/* --------------------------------------------------- sparse matrix helpers --- */
/* Device-resident CSC matrix (m-by-n, any shape) for SpMV and posterior-precision assembly. */
gmrfb_status gmrfb_spm_create(gmrfb_ctx* ctx, int64_t m, int64_t n, const int64_t* colptr,
                              const int64_t* rowval, const double* nzval, int32_t base, gmrfb_spm** out);
gmrfb_status gmrfb_spm_set_values(gmrfb_spm* A, const double* nzval);
gmrfb_status gmrfb_spm_destroy(gmrfb_spm* A);
/* y <- alpha * op(A) x + beta * y  (host vectors; trans != 0 selects A').  `Q*x`, `J*x`, `J'*r`
 * (scripts/solve_burger.jl:146,157). */
gmrfb_status gmrfb_spmv(const gmrfb_spm* A, int32_t trans, double alpha, const double* x, double beta,
                        double* y);
/* (v - mu)' Q (v - mu) — `sqmahal` (scripts/burgers/solve_burgers_gmrf-collocation.jl:262). mu may be NULL. */
gmrfb_status gmrfb_sqmahal(const gmrfb_spm* Q, const double* mu, const double* v, double* out);

/* Posterior precision  Qpost = Q + A' diag(qeps) A  — the assembly inside condition_on_observations
 * (scripts/darcy/solve_darcy_gmrf-fem.jl:165-167) and `Q + noise * J' * J` (scripts/solve_burger.jl:145).
 * The pattern is computed once (symbolic, host) and the values by a device kernel; re-running with new
 * values of A on the same pattern reuses the plan (Gauss-Newton). qeps_diag may be NULL (then qeps_scalar
 * is used for every observation).  The result is a new gmrfb_spm (full symmetric storage). */
typedef struct gmrfb_postprec gmrfb_postprec;
gmrfb_status gmrfb_postprec_create(gmrfb_ctx* ctx, const gmrfb_spm* Q, const gmrfb_spm* A,
                                   gmrfb_postprec** out);
gmrfb_status gmrfb_postprec_destroy(gmrfb_postprec* plan);
/* Numeric phase; `Qpost` is owned by the plan and valid until the plan is destroyed. */
gmrfb_status gmrfb_postprec_compute(gmrfb_postprec* plan, double qeps_scalar, const double* qeps_diag,
                                    const gmrfb_spm** Qpost);
/* Copy a device matrix's pattern/values back (any pointer may be NULL; sizes from gmrfb_spm_dims). */
gmrfb_status gmrfb_spm_dims(const gmrfb_spm* A, int64_t* m, int64_t* n, int64_t* nnz);
gmrfb_status gmrfb_spm_get(const gmrfb_spm* A, int32_t base, int64_t* colptr, int64_t* rowval,
                           double* nzval);
/* The plan's result matrix (pattern fixed at creation, values of the last gmrfb_postprec_compute); owned by the plan. */
gmrfb_status gmrfb_postprec_result(gmrfb_postprec* plan, const gmrfb_spm** Qpost);
/* Device pointer to the matrix values (for gmrfb_factorize_dev). */
const double* gmrfb_spm_values_dev(const gmrfb_spm* A);

/* Fixed-pattern sparse product  C = alpha * A * diag(w) * B  (w: A's column count doubles, host or device, or NULL for
 * the identity): the repeated products of a prior construction, e.g. the third Matern power
 * `ratio * K * Mt^-1 * K * Mt^-1 * K` (src/spdes/shallow_water.jl:186; scripts/darcy/solve_darcy_gmrf-fem.jl:97 asks
 * for smoothness 2).  The pattern of C is computed once (symbolic, host), the values by a device kernel with a fixed
 * summation order; A and B must outlive the plan, new values of A / B on the same patterns re-use it.  *C_out is
 * owned by the plan. */
typedef struct gmrfb_spgemm gmrfb_spgemm;
gmrfb_status gmrfb_spgemm_create(gmrfb_ctx* ctx, const gmrfb_spm* A, const gmrfb_spm* B, gmrfb_spgemm** out);
gmrfb_status gmrfb_spgemm_destroy(gmrfb_spgemm* plan);
gmrfb_status gmrfb_spgemm_compute(gmrfb_spgemm* plan, double alpha, const double* w_diag, const gmrfb_spm** C_out);

/* Evaluation metrics of src/metrics.jl:3-13 on the device: pred = E x (E = evaluation matrix, e.g. the 241 x 241 grid
 * of scripts/darcy/solve_darcy_gmrf-fem.jl:86-89,190-196; NULL: pred = x), out3 = { rmse, max_err, rel_err } of pred
 * against `truth` (ntruth values).  x and truth may be host or device pointers. */
gmrfb_status gmrfb_metrics(gmrfb_ctx* ctx, const gmrfb_spm* E, const double* x, const double* truth, int64_t ntruth,
                           double* out3);

/* ------------------------------------------- P1 finite-element assembly ------ */
/* The two matrices the reference rebuilds inside its hot loops, assembled on the device on a fixed pattern:
 *   assemble_darcy_diff_matrix(disc, x_coords, y_coords, coeff_mat)   src/problems/darcy.jl:5-63  -> gmrfb_fem_assemble
 *       (called per problem of the dataset loop through form_observations, scripts/darcy/solve_darcy_gmrf-fem.jl:104-137,178;
 *        the coefficient of an element is coeff_mat[get_xy_idcs(quadrature point)], src/datasets/darcy.jl:30-34)
 *   Q_matern = ratio * K_matern' * Mt^-1 * K_matern,  K_matern = kappa^2 Mt + G    src/spdes/shallow_water.jl:177-194
 *                                                                               -> gmrfb_fem_matern_precision
 * for linear (P1) triangles with the one-point rule (exact for piecewise-constant coefficients and P1 gradients).
 *   nodes : nnodes x 2 coordinates, node-major;  tris : ntri x 3 vertex indices (`base`-based), element-major.
 * The mesh is analysed once (pattern, element -> nonzero gather lists); every later call is a few kernels. */
typedef struct gmrfb_fem gmrfb_fem;
gmrfb_status gmrfb_fem_create(gmrfb_ctx* ctx, int64_t nnodes, const double* nodes, int64_t ntri, const int64_t* tris,
                              int32_t base, gmrfb_fem** out);
gmrfb_status gmrfb_fem_destroy(gmrfb_fem* fem);
/* lumped mass vector (integral of every hat function), nnodes doubles: the load vector of f = 1 (`fe[i] += beta * du * dOmega`
 * with beta = 1, src/problems/darcy.jl:46) */
gmrfb_status gmrfb_fem_get_mass(gmrfb_fem* fem, double* mass_out);
/* coefficient grid axes: gx values x_coords, gy values y_coords; element -> grid cell by nearest index per axis */
gmrfb_status gmrfb_fem_set_coeff_grid(gmrfb_fem* fem, int64_t gx, const double* x_coords, int64_t gy,
                                      const double* y_coords);
/* G = sum_T coeff(T) K_T.  coeff_grid: gx*gy doubles, entry ix + iy*gx = coeff_mat[ix, iy] (host or device memory), or
 * NULL for a unit coefficient; prescribed: nnodes bytes (non-zero = Dirichlet dof: its row becomes an identity row) or
 * NULL.  *G_out is a matrix owned by the handle (fixed pattern; values of the last call): hand it to
 * gmrfb_postprec_create once, then every gmrfb_fem_assemble + gmrfb_postprec_compute pair re-uses the plan. */
gmrfb_status gmrfb_fem_assemble(gmrfb_fem* fem, const double* coeff_grid, const uint8_t* prescribed,
                                const gmrfb_spm** G_out);
/* Q = ratio * K' Mt^-1 K with K = kappa^2 Mt + G (unit coefficient) and lumped Mt; for prescribed dofs (optional mask)
 * Mt_ii = prescribed_mass and G_ii = 1 as in src/spdes/shallow_water.jl:178-181.  *Q_out is owned by the handle. */
gmrfb_status gmrfb_fem_matern_precision(gmrfb_fem* fem, double kappa, double ratio, const uint8_t* prescribed,
                                        double prescribed_mass, const gmrfb_spm** Q_out);

/* Gauss-Newton tangent of the cubic reaction term -lap u + u^3 = g by quadrature of the current iterate
 * (_research/elliptic_chen24.jl: assemble_J_diff_and_f :180-228, assemble_J_cube :231-278, f_and_J :280-285):
 *     J = s G + J_cube,  J_cube[i,j] = sum_q 3 phi_i u_q^2 phi_j dOmega,
 *     f = s G u + f_cube,  f_cube[i] = sum_q phi_i u_q^3 dOmega      (the caller subtracts its static load vector),
 * s = stiffness_scale (0: the cubic part alone, as assemble_J_cube returns it).  Rows of prescribed dofs are skipped
 * (stay zero) in both parts, as the reference's `continue` does (:207-209,:259-261).  u: nnodes doubles, host or
 * device; quad_degree in {1, 2, 4}: symmetric triangle rule exact to that degree (2 = the 3-point rule of
 * QuadratureRule{RefTriangle}(2), :121).  *J_out is owned by the handle (pattern of the stiffness matrix, values of
 * the last call); f_out (nnodes doubles, host or device) may be NULL. */
gmrfb_status gmrfb_fem_assemble_cubic(gmrfb_fem* fem, const double* u, int32_t quad_degree, double stiffness_scale,
                                      const uint8_t* prescribed, const gmrfb_spm** J_out, double* f_out);

/* ------------------------------------------- Lagrange triangles of order 1 / 2 --- */
/* The discretisations the reference's scripts run: `uniform_unit_square_discretization(N; element_order = 2)` with
 * `QuadratureRule{RefTriangle}(element_order + 1)` (src/utils.jl:20-38) for the Darcy dataset loop, and
 * `generate_grid(QuadraticTriangle, ...)` with the same rule (_research/elliptic_chen24.jl:118-122) for the elliptic
 * Gauss-Newton solve.  Cell values as Ferrite computes them: isoparametric geometry, physical gradients and dOmega at
 * every quadrature point.
 *   order : 1 (3 nodes per element) or 2 (6 nodes: vertices 0, 1, 2, then the nodes on the edges (0,1), (1,2), (2,0) -
 *           Ferrite's QuadraticTriangle);  elems : nelem x nodes_per_element indices (`base`-based), element-major;
 *   quad_degree : 0 = order + 1, else 1, 2, 3 or 4 (symmetric rules of 1, 3, 4 and 6 points; 3 is Dunavant's rule with
 *           the negative centroid weight). */
typedef struct gmrfb_fem2d gmrfb_fem2d;
gmrfb_status gmrfb_fem2d_create(gmrfb_ctx* ctx, int32_t order, int64_t nnodes, const double* nodes, int64_t nelem,
                                const int64_t* elems, int32_t base, int32_t quad_degree, gmrfb_fem2d** out);
gmrfb_status gmrfb_fem2d_destroy(gmrfb_fem2d* fem);
gmrfb_status gmrfb_fem2d_info(gmrfb_fem2d* fem, int32_t* order, int32_t* nodes_per_element, int32_t* nquad, int64_t* nnz);
/* coefficient grid axes; every QUADRATURE POINT is mapped to its grid cell by nearest index per axis, first minimum on
 * ties (`coeff_mat[get_xy_idcs(x, x_coords, y_coords)...]`, src/problems/darcy.jl:33-39, src/datasets/darcy.jl:30-34) */
gmrfb_status gmrfb_fem2d_set_coeff_grid(gmrfb_fem2d* fem, int64_t gx, const double* x_coords, int64_t gy,
                                        const double* y_coords);
/* assemble_darcy_diff_matrix (src/problems/darcy.jl:5-63): G[i,j] = sum_q coeff(x_q) grad phi_i . grad phi_j dOmega and
 * the load f[i] = beta sum_q phi_i dOmega.  coeff_grid: gx*gy doubles, entry ix + iy*gx = coeff_mat[ix, iy] (host or
 * device), or NULL for a unit coefficient; prescribed: nnodes bytes (non-zero = Dirichlet dof: identity row, zero
 * load) or NULL; load_out: nnodes doubles (host or device) or NULL.  *G_out is owned by the handle (fixed pattern,
 * values of the last call). */
gmrfb_status gmrfb_fem2d_stiffness(gmrfb_fem2d* fem, const double* coeff_grid, const uint8_t* prescribed, double beta,
                                   const gmrfb_spm** G_out, double* load_out);
/* Mass matrix on the stiffness pattern.  lumping 0: consistent, sum_q phi_i phi_j dOmega; 1: row sums of every element
 * mass; 2: the diagonal of every element mass scaled to the element's total mass, diag(me) sum(me) / sum(diag(me));
 * 3: what `lump_matrix(me, ip)` (src/spdes/shallow_water.jl:115) does for this order - 1 for order 1, 2 for order 2,
 * where the row sums of the vertex functions vanish.  lumped_out (nnodes doubles, host) may be NULL. */
gmrfb_status gmrfb_fem2d_mass(gmrfb_fem2d* fem, int32_t lumping, const gmrfb_spm** M_out, double* lumped_out);
/* Matern prior (src/spdes/shallow_water.jl:172-190): K = kappa^2 Mt + G with the lumped mass of the order (above) and,
 * for prescribed dofs, Mt_ii = prescribed_mass, G_ii = 1;  alpha 2: Q = ratio K' Mt^-1 K (:187), alpha 3:
 * Q = ratio K Mt^-1 K Mt^-1 K (:186).  *Q_out is owned by the handle. */
gmrfb_status gmrfb_fem2d_matern_precision(gmrfb_fem2d* fem, double kappa, double ratio, int32_t alpha,
                                          const uint8_t* prescribed, double prescribed_mass, const gmrfb_spm** Q_out);
/* f_and_J of _research/elliptic_chen24.jl:180-285 with the rule the handle was created with: J = s J_diff + J_cube,
 * f = s J_diff u + f_cube (see gmrfb_fem_assemble_cubic); rows of prescribed dofs skipped. */
gmrfb_status gmrfb_fem2d_assemble_cubic(gmrfb_fem2d* fem, const double* u, double stiffness_scale,
                                        const uint8_t* prescribed, const gmrfb_spm** J_out, double* f_out);

/* ------------------------------------------- 1-D finite elements (Burgers) --- */
/* Lagrange line elements of order 1 or 2 (quadratic lines numbered left, right, middle as Ferrite's QuadraticLine;
 * periodic_unit_interval_discretization, src/utils.jl:42-49, uses order 2 with QuadratureRule{RefLine}(order + 1)):
 *   assemble_burgers_mass_diffusion_matrices(disc; lumping)   src/problems/burgers.jl:61-98  -> gmrfb_fem1d_mass_stiffness
 *   assemble_burgers_advection_matrix(disc, cur_weights)      src/problems/burgers.jl:5-59   -> gmrfb_fem1d_advection
 *   f_and_J of the Gauss-Newton loop (J_static + dt J_adv over all time steps)
 *                              scripts/burgers/solve_burgers_gmrf-fem.jl:115-142              -> gmrfb_fem1d_spacetime_tangent
 *   elems : nelem x (order + 1) node indices (`base`-based), element-major;  elem_x : the coordinates of those nodes,
 *   same layout (per element, so that the last element of a periodic mesh can close the ring: its right end has
 *   coordinate 1 while the node it shares with the first element has coordinate 0).
 *   nquad: Gauss-Legendre points per element (0: order + 1; at most 4).
 * prescribed (optional, nnodes bytes): rows and columns of those dofs are zero in every matrix and their vector
 * entries are zero (apply! followed by the explicit zeroing of :53-57,:88-93).  Returned matrices are owned by the
 * handle (fixed pattern, values of the last call). */
typedef struct gmrfb_fem1d gmrfb_fem1d;
gmrfb_status gmrfb_fem1d_create(gmrfb_ctx* ctx, int64_t nnodes, int64_t nelem, const int64_t* elems, const double* elem_x,
                                int32_t order, int32_t base, int32_t nquad, gmrfb_fem1d** out);
gmrfb_status gmrfb_fem1d_destroy(gmrfb_fem1d* fem);
/* M = int phi_i phi_j (lumping != 0: diagonal of its row sums), G = int phi_i' phi_j' */
gmrfb_status gmrfb_fem1d_mass_stiffness(gmrfb_fem1d* fem, int32_t lumping, const uint8_t* prescribed,
                                        const gmrfb_spm** M_out, const gmrfb_spm** G_out);
/* A[i,j] = sum_q phi_i (phi_j u_x + u phi_j') dOmega, v[i] = sum_q phi_i u u_x dOmega for the iterate u (nnodes
 * doubles, host or device); v_out (host or device) may be NULL */
gmrfb_status gmrfb_fem1d_advection(gmrfb_fem1d* fem, const double* u, const uint8_t* prescribed, const gmrfb_spm** A_out,
                                   double* v_out);
/* w: nt * nnodes doubles, time-major (step t = entries [t nnodes, (t+1) nnodes)), host or device.
 * J ((nt-1) nnodes x nt nnodes): block (t, t) = -M, block (t, t+1) = M + dt nu G + dt A(w_{t+1}), t = 0 .. nt-2;
 * f = J_static w + dt [v(w_1); ...; v(w_{nt-1})] ((nt-1) nnodes doubles, host or device, may be NULL).  One kernel
 * writes all values of J; hand *J_out to gmrfb_postprec_create once and re-use the plan at every iteration. */
gmrfb_status gmrfb_fem1d_spacetime_tangent(gmrfb_fem1d* fem, int64_t nt, double dt, double nu, const double* w,
                                           const uint8_t* prescribed, const gmrfb_spm** J_out, double* f_out);

/* ------------------------------------------- Gauss-Newton on the device ------ */
/* The explicit loop of scripts/solve_burger.jl:143-180 (packaged as GaussNewtonOptimizer / optimize in
 * scripts/burgers/solve_burgers_gmrf-fem.jl:172-182) for a bilinear collocation residual
 *     f(w) = L w + c (A w) .* (D w),        J(w) = L + c (diag(D w) A + diag(A w) D),
 * of which the Burgers residual f = A1 w - A0 w + dt (A1 w).*(D w) - dt nu D2 w (:127-134) is the instance
 * L = A1 - A0 - dt nu D2, A = A1, c = dt.  Per iteration, all on the device: residual and tangent, the
 * fixed-pattern assembly Q + noise J'J (:145), numeric refactorisation on the pattern analysed once, the solve
 * x+ = (Q + noise J'J)^{-1} (Q mu + noise J'(J x + y - f)) (:146-148) and the objective
 * (mu-x)'Q(mu-x) + noise |y-f|^2; the loop stops when its relative change is <= rel_tol or after max_steps (:171-180).
 * An optional elementwise cubic term, f += e .* w.^3 (J += diag(3 e w.^2); square systems whose pattern holds the
 * diagonal), covers the elliptic problem -lap u + u^3 = g of _research/elliptic_chen24.jl:231-285 (L = stiffness,
 * e = lumped mass).
 *   colptr/rowval : CSC union pattern of L, A, D (m-by-n, n = size of Q); lval/aval/dval are aligned to it (zeros where
 *                   a matrix has no entry)
 *   cubic         : e (length n) or NULL
 *   perm          : fill-reducing permutation to reuse (`opts->base`-based), or NULL to order by `opts` (may be NULL) */
typedef struct gmrfb_gn gmrfb_gn;
gmrfb_status gmrfb_gn_create(gmrfb_ctx* ctx, const gmrfb_spm* Q, int64_t m, const int64_t* colptr,
                             const int64_t* rowval, const double* lval, const double* aval, const double* dval,
                             const double* cubic, int32_t base, double c, double noise, const double* y, const double* mu,
                             const int64_t* perm, const gmrfb_analyze_opts* opts, gmrfb_gn** out);
/* x: in = starting point x0, out = final iterate; obj_hist (max_steps + 1 doubles, may be NULL) receives the objective
 * at x0 and after every step; *steps = iterations taken. */
gmrfb_status gmrfb_gn_optimize(gmrfb_gn* gn, double* x, int32_t max_steps, double rel_tol, int32_t* steps,
                               double* obj_hist);
/* Borrowed views of the optimiser's state after gmrfb_gn_optimize: the factor of the last Q + noise J'J (for
 * mean / sample / variance calls), its symbolic analysis, the last tangent J_k and the assembled matrix
 * (`Jₖ`, `Q_mat` in scripts/burgers/solve_burgers_gmrf-fem.jl:184-188).  Any of the outputs may be NULL. */
gmrfb_status gmrfb_gn_get(gmrfb_gn* gn, gmrfb_fac** fac, gmrfb_sym** sym, const gmrfb_spm** J, const gmrfb_spm** Qpost);
gmrfb_status gmrfb_gn_destroy(gmrfb_gn* gn);

/* ------------------------------------------- block-tridiagonal Cholesky ------ */
/* Replaces src/tridiagonal_cholesky.jl:
 *   tridiagonal_cholesky(A, N_blocks)            :65-82   -> gmrfb_btd_factor
 *   TridiagonalCholeskyFactor{N, chos, Cs}       :5-9     -> gmrfb_btd + gmrfb_btd_get_block
 *   forward_solve / backward_solve / ldiv!       :24-63   -> gmrfb_btd_solve (intended semantics)
 * A is block tridiagonal with N_blocks blocks of size b = n / N_blocks (integer division; trailing
 * n mod N_blocks rows are ignored exactly as at :66; entries outside the block tridiagonal are ignored).
 *   L_1 = chol(D_1);  C_i = B_i L_{i-1}^{-T};  L_i = chol(D_i - C_i C_i').                          */
gmrfb_status gmrfb_btd_factor(gmrfb_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval,
                              const double* nzval, int32_t base, int64_t nblocks, gmrfb_btd** out);
/* Dense-block entry: D is b-by-b-by-N (diagonal blocks, lower triangle read), B is b-by-b-by-(N-1)
 * (B[:,:,k] = block (k+2, k+1) of A, 1-based), both column-major and contiguous. */
gmrfb_status gmrfb_btd_factor_dense(gmrfb_ctx* ctx, int64_t b, int64_t nblocks, const double* D,
                                    const double* B, gmrfb_btd** out);
/* Constant-mesh implicit-Euler state-space prior built directly in block form (ingredients of
 * src/spdes/shallow_water.jl:198-228): the N-block matrix has diagonal blocks D_first, D_mid (blocks 2..N-1), D_last and
 * one sub-diagonal block B_sub everywhere.  Four b-by-b column-major blocks (host or device) instead of N. */
gmrfb_status gmrfb_btd_factor_ssm(gmrfb_ctx* ctx, int64_t b, int64_t nblocks, const double* D_first,
                                  const double* D_mid, const double* D_last, const double* B_sub, gmrfb_btd** out);
gmrfb_status gmrfb_btd_destroy(gmrfb_btd* f);
enum { GMRFB_BTD_BLOCK_L = 0, GMRFB_BTD_BLOCK_C = 1 };
/* Copy block i (0-based) out: which = L -> `chos[i+1].L` (b-by-b lower, upper zeroed);
 * which = C -> `Cs[i+1]` (the sub-diagonal block of L in block row i+2), i in [0, N-2]. */
gmrfb_status gmrfb_btd_get_block(gmrfb_btd* f, int64_t i, int32_t which, double* out, int64_t ldo);
enum { GMRFB_BTD_SOLVE_A = 0, GMRFB_BTD_SOLVE_FWD = 1, GMRFB_BTD_SOLVE_BWD = 2 };
/* X (b*N-by-nrhs, column-major, host) in place: FWD = L^{-1} X, BWD = L^{-T} X, A = A^{-1} X. */
gmrfb_status gmrfb_btd_solve(gmrfb_btd* f, int32_t mode, double* X, int64_t ldx, int64_t nrhs);
gmrfb_status gmrfb_btd_logdet(gmrfb_btd* f, double* logdet);
/* diag(A^{-1}) via the block Takahashi recursion  S_N = L_N^{-T} L_N^{-1},
 * S_i = L_i^{-T}(I + C_{i+1}' S_{i+1} C_{i+1}) L_i^{-1}. */
gmrfb_status gmrfb_btd_selinv_diag(gmrfb_btd* f, double* var_out);
typedef struct gmrfb_btd_info {
  int64_t b, nblocks;
  int32_t status;      /* GMRFB_OK or GMRFB_ERR_NOT_SPD      */
  int64_t fail_block;  /* first block whose POTRF failed, or -1 */
  double flops;        /* (N-1) * 7/3 b^3 + b^3/3             */
} gmrfb_btd_info;
gmrfb_status gmrfb_btd_get_info(gmrfb_btd* f, gmrfb_btd_info* info);

/* Time-sharded block-tridiagonal factor/solve across ranks (one process per GPU).  Rank r owns a contiguous slab
 * of nloc blocks; ranks 0..P-2 use their last block as a separator.  All numeric work is local; the only exchange
 * is an all-gather of three b-by-b blocks per rank (factor) and two b-by-nrhs panels per rank (solve), which the
 * host performs between the phased calls below (torch.distributed / NCCL all_gather_into_tensor over NVLink in
 * the Python host; NCCL.jl or MPI from Julia — see INTEGRATION.md).
 *   D_local : b-by-b-by-nloc diagonal blocks of the slab
 *   B_local : b-by-b-by-nloc, B_local[:,:,k] = block (slab block k, the block before it); ignored for k = 0 on rank 0
 * D_local / B_local (and D / Bsub of gmrfb_btd_factor_dense) may be host or device pointers (unified addressing). */
typedef struct gmrfb_btd_dist gmrfb_btd_dist;
gmrfb_status gmrfb_btd_dist_create(gmrfb_ctx* ctx, int32_t rank, int32_t nranks, int64_t b, int64_t nloc,
                                   const double* D_local, const double* B_local, gmrfb_btd_dist** out);
/* doubles each rank contributes to the factor exchange (3 b^2) */
int64_t gmrfb_btd_dist_iface_count(const gmrfb_btd_dist* h);
/* copy this rank's interface blocks into a caller-owned DEVICE buffer of iface_count doubles */
gmrfb_status gmrfb_btd_dist_get_iface(gmrfb_btd_dist* h, double* d_out);
/* d_all: the all-gathered interface blocks (DEVICE, rank-major, nranks * iface_count doubles): assembles and
 * factorises the (P-1)-block reduced Schur system redundantly on every rank */
gmrfb_status gmrfb_btd_dist_reduce(gmrfb_btd_dist* h, const double* d_all);
/* doubles each rank contributes to the solve exchange for nrhs right-hand sides */
int64_t gmrfb_btd_dist_solve_count(const gmrfb_btd_dist* h, int64_t nrhs);
/* phase 1: X_local (b*nloc-by-nrhs, host) -> local forward elimination; fills d_send (DEVICE, solve_count doubles) */
gmrfb_status gmrfb_btd_dist_solve_begin(gmrfb_btd_dist* h, const double* X_local, int64_t ldx, int64_t nrhs,
                                        double* d_send);
/* phase 2: d_all = all-gathered d_send of every rank (DEVICE); X_local <- this rank's rows of A^{-1} X */
gmrfb_status gmrfb_btd_dist_solve_end(gmrfb_btd_dist* h, const double* d_all, double* X_local, int64_t ldx,
                                      int64_t nrhs);
/* log det A = sum over ranks of *local_part + *reduced_part (the reduced part is identical on every rank) */
gmrfb_status gmrfb_btd_dist_logdet(gmrfb_btd_dist* h, double* local_part, double* reduced_part);
gmrfb_status gmrfb_btd_dist_destroy(gmrfb_btd_dist* h);

#ifdef __cplusplus
}
#endif
#endif /* GMRFB_H */
