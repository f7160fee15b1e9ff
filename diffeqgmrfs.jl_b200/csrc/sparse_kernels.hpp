// Host-visible interface of sparse_kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace gmrfb {

constexpr int SOLVE_NRC = 4;  // right-hand sides processed per pass of the level-scheduled solves (batches of up to 4)
constexpr int MR_MAX = 64;    // right-hand sides per pass of the panel solves (larger batches, solve_mr.cu)

// Per-supernode record used by the solve kernels (device copy of the symbolic structure).
struct SnodeDesc {
  int64_t foff;      // front offset in the frontal arena
  int64_t rows_off;  // offset of the front's row list / relmap
  int64_t uoff;      // offset (in rows) of the supernode's update vector
  int32_t ld, d, s, col0;
  int32_t child0, nchild;
};

cudaError_t sparse_kernels_init();
constexpr int SOLVE_SMALL_MAX = 160;  // supernodes with fronts up to this order are solved by one warp each (fused)
constexpr int SOLVE_FS_ROWS = 128;  // rows per CTA of the forward block step
constexpr int SOLVE_BR_ROWS = 512;  // rows per CTA of the backward R-part reduction
cudaError_t launch_perm_gather(const double* src, int64_t lds, double* dst, int64_t ldd, const int32_t* perm,
                               int64_t n, int nrhs, cudaStream_t st);
cudaError_t launch_perm_scatter(const double* src, int64_t lds, double* dst, int64_t ldd, const int32_t* perm,
                                int64_t n, int nrhs, const double* add, cudaStream_t st);
cudaError_t launch_perm_scatter_nodemajor(const double* src, int64_t lds, double* dst, int64_t ldk,
                                          const int32_t* perm, int64_t n, int k0, int nr, cudaStream_t st);
struct Task;
cudaError_t launch_fwd_small(const SnodeDesc* sd, const int32_t* list, int count, const int32_t* child_idx,
                             const int32_t* relmap, const double* F, double* w, double* ysol, int64_t ldx, double* uvec,
                             int nr, cudaStream_t st);
cudaError_t launch_bwd_small(const SnodeDesc* sd, const int32_t* list, int count, const int32_t* rows, const double* F,
                             const double* t, double* xsol, int64_t ldx, int nr, cudaStream_t st);
cudaError_t launch_fwd_assemble(const SnodeDesc* sd, const int32_t* list, int count, const int32_t* child_idx,
                                const int32_t* relmap, double* x, int64_t ldx, double* uvec, cudaStream_t st);
// dinv: the factor's stored inverses of 64x64 diagonal blocks (tasks whose `alpha` holds an offset >= 0 use them)
cudaError_t launch_fwd_step(const Task* tasks, int ntasks, int grid, const double* F, double* w, double* ysol,
                            int64_t ldx, double* uvec, int nr, const double* dinv, cudaStream_t st);
cudaError_t launch_bwd_rpart(const Task* tasks, int ntasks, int grid, const double* F, const int32_t* rows,
                             const double* x, int64_t ldx, double* partial, int nr, cudaStream_t st);
cudaError_t launch_bwd_step(const Task* tasks, int ntasks, int grid, const double* F, double* t, double* xsol,
                            int64_t ldx, const double* partial, int nr, const double* dinv, cudaStream_t st);
// wide supernodes (full inverse W_J kept by the factorisation): y_J = W_J x_J; u_J -= L21 y_J; x_J = W_J' (t_J - partials)
constexpr int SOLVE_WIDE_MIN = 128;  // supernodes with at least this many columns keep their full inverse (GMRFB_WIDE_MIN)
constexpr int SOLVE_WIDE_MAX = 4096; // ... and at most this many
constexpr int SOLVE_WG_ROWS = 32;    // rows per CTA of the row-oriented wide products
cudaError_t launch_wide_fwd(const Task* tasks, int ntasks, int grid_trmv, int grid_gemv, const double* F, const double* Wf,
                            const double* w, double* ysol, int64_t ldx, double* uvec, int nr, cudaStream_t st);
cudaError_t launch_wide_fwd_below(const Task* tasks, int ntasks, int grid, const double* F, const double* ysol, int64_t ldx,
                                  double* uvec, int nr, cudaStream_t st);
cudaError_t launch_wide_bwd(const Task* tasks, int ntasks, int grid, const double* Wf, const double* t, double* xsol,
                            int64_t ldx, const double* partial, int nr, cudaStream_t st);
cudaError_t launch_wide_norms(const Task* tasks, int ntasks, int grid, const double* F, const double* Wf, double* norms,
                              cudaStream_t st);
cudaError_t launch_spmv_rows(int64_t nrows, const int64_t* ptr, const int32_t* idx, const double* val,
                             const double* x, double* y, double alpha, double beta, cudaStream_t st);
cudaError_t launch_rbmc(int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, const double* X,
                        int64_t ldk, int nsamp, double* var, cudaStream_t st);
cudaError_t launch_gather_values(const double* src, const int64_t* map, int64_t nnz, double* dst, cudaStream_t st);
cudaError_t launch_postprec(int64_t nnz_out, const int64_t* qsrc, const double* Qval, const int64_t* pptr,
                            const int32_t* prow, const int64_t* pa, const int64_t* pb, const double* Aval,
                            const double* wdiag, double wscalar, double* out, cudaStream_t st);
cudaError_t launch_dot(const double* a, const double* b, int64_t n, double* out, cudaStream_t st);
cudaError_t launch_axpby(int64_t n, double a, const double* x, double b, const double* y, double* out,
                         cudaStream_t st);
cudaError_t launch_diag_L(const SnodeDesc* sd, int nsuper, const double* F, double* out, cudaStream_t st);

}  // namespace gmrfb
