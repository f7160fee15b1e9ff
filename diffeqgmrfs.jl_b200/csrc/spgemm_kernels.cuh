// Kernel and symbolic phase of the fixed-pattern sparse product (gmrfb_spgemm, spm.cu).  No CUDA runtime calls in here:
// also compiled as plain C++ by tools/probe/fem2d_emul.cpp (host emulation against SciPy).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace gmrfb {
namespace spgemm {

// pattern of C = A B (A: m x k, B: k x n, CSC with sorted columns): column j is the sorted union of the columns of A
// selected by the rows of B(:, j).  colnz[p] = column of the CSC position p of C.
inline void product_pattern(int64_t m, int64_t n, const int64_t* acolptr, const int32_t* arow, const int64_t* bcolptr,
                            const int32_t* brow, std::vector<int64_t>& ccolptr, std::vector<int64_t>& crow,
                            std::vector<int32_t>& colnz) {
  ccolptr.assign((size_t)n + 1, 0);
  crow.clear();
  colnz.clear();
  std::vector<int32_t> mark((size_t)m, -1), rows;
  for (int64_t j = 0; j < n; j++) {
    rows.clear();
    for (int64_t p = bcolptr[j]; p < bcolptr[j + 1]; p++) {
      const int32_t kk = brow[p];
      for (int64_t q = acolptr[kk]; q < acolptr[kk + 1]; q++) {
        const int32_t i = arow[q];
        if (mark[i] != (int32_t)j) {
          mark[i] = (int32_t)j;
          rows.push_back(i);
        }
      }
    }
    std::sort(rows.begin(), rows.end());
    for (int32_t i : rows) {
      crow.push_back(i);
      colnz.push_back((int32_t)j);
    }
    ccolptr[j + 1] = (int64_t)crow.size();
  }
}

__global__ void k_spgemm(int64_t nnzC, const int32_t* __restrict__ crow, const int32_t* __restrict__ ccol,
                         const int64_t* __restrict__ acolptr, const int32_t* __restrict__ arow,
                         const double* __restrict__ aval, const int64_t* __restrict__ bcolptr,
                         const int32_t* __restrict__ brow, const double* __restrict__ bval,
                         const double* __restrict__ w, double alpha, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnzC) return;
  const int32_t i = crow[k], j = ccol[k];
  double v = 0.0;
  for (int64_t p = bcolptr[j]; p < bcolptr[j + 1]; p++) {
    const int32_t kk = brow[p];
    int64_t lo = acolptr[kk], hi = acolptr[kk + 1];  // first position in [lo, hi) with arow >= i
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (arow[mid] < i) lo = mid + 1;
      else hi = mid;
    }
    if (lo < acolptr[kk + 1] && arow[lo] == i) v += aval[lo] * (w ? w[kk] : 1.0) * bval[p];
  }
  out[k] = alpha * v;
}
}  // namespace spgemm
}  // namespace gmrfb
