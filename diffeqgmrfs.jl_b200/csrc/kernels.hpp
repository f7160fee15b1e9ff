// Host-visible interface of the CUDA kernels (kernels.cu, sparse_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "tasks.hpp"

namespace gmrfb {

// Up to four device arenas that Task offsets index into (selected by the TF_*_SHIFT bits of Task::flags).
struct Arenas {
  double* p[4];
  double* dinv = nullptr;  // scratch of inverted <=64x64 diagonal blocks (TF_B_DINV operands)
};

struct SnodeDesc;
struct LaunchAux {
  int* d_info = nullptr;             // POTRF failure column (atomicMin)
  const int32_t* d_relmap = nullptr; // relative indices for extend-add / gather
  double* d_out = nullptr;           // DIAG_OUT destination
  const SnodeDesc* d_snodes = nullptr;   // supernode records (fused small-front kernels)
  const int32_t* d_child_idx = nullptr;  // children lists
  const int32_t* d_sparent = nullptr;    // parent supernode of each supernode
  const int32_t* d_rows = nullptr;       // row lists of the fronts (panel solves, solve_mr.cu)
  int nr = 0, ldk = 0;                   // panel solves: right-hand sides in the panel and its leading dimension
};

cudaError_t kernels_init();
cudaError_t mr_kernels_init();
// panel-solve launch kinds (LK_MR_*), solve_mr.cu
cudaError_t run_mr_launch(const Launch& L, const Task* tasks, const Arenas& ar, const LaunchAux& aux, cudaStream_t st);
cudaError_t launch_mr_perm_in(const double* src, int64_t lds, double* X, int ldk, const int32_t* perm, int64_t n,
                              int nr, cudaStream_t st);
cudaError_t launch_mr_perm_out(const double* X, int ldk, double* dst, int64_t ldd, const int32_t* perm, int64_t n,
                               int nr, const double* add, cudaStream_t st);
cudaError_t launch_mr_perm_nodemajor(const double* X, int ldk, double* dst, int64_t ldd, const int32_t* perm,
                                     int64_t n, int q0, int nr, cudaStream_t st);
cudaError_t run_launch(const Launch& L, const Task* d_tasks, const Arenas& ar, const LaunchAux& aux,
                       cudaStream_t st);
cudaError_t launch_scatter_values(const double* d_nzval, const int64_t* d_amap, int64_t nnz, double* d_arena,
                                  cudaStream_t st);

}  // namespace gmrfb
