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
};

cudaError_t kernels_init();
cudaError_t run_launch(const Launch& L, const Task* d_tasks, const Arenas& ar, const LaunchAux& aux,
                       cudaStream_t st);
cudaError_t launch_scatter_values(const double* d_nzval, const int64_t* d_amap, int64_t nnz, double* d_arena,
                                  cudaStream_t st);

}  // namespace gmrfb
