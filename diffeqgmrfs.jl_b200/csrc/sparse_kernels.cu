// Bandwidth-bound kernels of the sparse path: supernodal triangular solves (level-scheduled, multifrontal
// update vectors => deterministic, atomic-free), permutations, SpMV / SpMM, RBMC variance accumulation and the
// fixed-pattern posterior-precision assembly.
#include <cuda_runtime.h>

#include <cstdint>

#include "sparse_kernels.hpp"

namespace gmrfb {

constexpr int SOLVE_THREADS = 256;
constexpr int DLD = 65;  // padded leading dimension of the 64x64 diagonal block in shared memory

// ------------------------------------------------------------------------------------- permutations ----
// dst[k + q*ldd] = src[perm[k] + q*lds]   (gather rows through perm)
__global__ void k_perm_gather(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd,
                              const int32_t* __restrict__ perm, int64_t n, int nrhs) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t p = perm[k];
  for (int q = 0; q < nrhs; q++) dst[k + q * ldd] = src[p + q * lds];
}
// dst[perm[k] + q*ldd] = src[k + q*lds] (+ add[perm[k]] if add)
__global__ void k_perm_scatter(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd,
                               const int32_t* __restrict__ perm, int64_t n, int nrhs,
                               const double* __restrict__ add) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t p = perm[k];
  double a = add ? add[p] : 0.0;
  for (int q = 0; q < nrhs; q++) dst[p + q * ldd] = src[k + q * lds] + a;
}

// --------------------------------------------------------------------------------- forward solve ----
// One CTA per supernode J of the level.  w = [x_C ; 0] + sum_children u_c (mapped), then
//   y_C = L11^{-1} w_C,   u_J = w_R - L21 y_C   (u holds minus the accumulated updates).
template <int NRC>
__global__ void __launch_bounds__(SOLVE_THREADS) k_fwd_level(const SnodeDesc* __restrict__ sd,
                                                             const int32_t* __restrict__ list,
                                                             const int32_t* __restrict__ child_idx,
                                                             const int32_t* __restrict__ relmap,
                                                             const double* __restrict__ F, double* __restrict__ x,
                                                             int64_t ldx, double* __restrict__ uvec, int nr) {
  extern __shared__ __align__(16) double sm[];
  double* Ld = sm;
  double* invd = sm + 64 * DLD;
  double* w = invd + 64;
  const SnodeDesc D = sd[list[blockIdx.x]];
  const int d = D.d, s = D.s, r = d - s, ld = D.ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ Fj = F + D.foff;
  for (int i = tid; i < d; i += SOLVE_THREADS) {
#pragma unroll
    for (int q = 0; q < NRC; q++) w[i + q * d] = (i < s && q < nr) ? x[D.col0 + i + q * ldx] : 0.0;
  }
  __syncthreads();
  for (int ci = 0; ci < D.nchild; ci++) {
    const SnodeDesc C = sd[child_idx[D.child0 + ci]];
    const int rc = C.d - C.s;
    const double* __restrict__ uc = uvec + C.uoff * NRC;
    const int32_t* __restrict__ rel = relmap + C.rows_off + C.s;
    for (int i = tid; i < rc; i += SOLVE_THREADS) {
      const int p = rel[i];
#pragma unroll
      for (int q = 0; q < NRC; q++) w[p + q * d] += uc[i + q * rc];
    }
    __syncthreads();
  }
  for (int k0 = 0; k0 < s; k0 += 64) {
    const int nb = min(64, s - k0);
    for (int e = tid; e < nb * nb; e += SOLVE_THREADS) {
      int i = e % nb, j = e / nb;
      if (i >= j) Ld[j * DLD + i] = Fj[(k0 + i) + (int64_t)(k0 + j) * ld];
    }
    __syncthreads();
    if (tid < nb) invd[tid] = 1.0 / Ld[tid * DLD + tid];
    __syncthreads();
    if (warp < NRC && warp < nr) {
      double* wq = w + warp * d + k0;
      double v0 = (lane < nb) ? wq[lane] : 0.0, v1 = (lane + 32 < nb) ? wq[lane + 32] : 0.0;
      for (int c = 0; c < nb; c++) {
        double yc = __shfl_sync(0xffffffffu, (c < 32) ? v0 : v1, c & 31) * invd[c];
        if (c < 32) {
          if (lane == c) v0 = yc;
          if (lane > c && lane < nb) v0 -= Ld[c * DLD + lane] * yc;
        } else if (lane == c - 32) {
          v1 = yc;
        }
        if (lane + 32 > c && lane + 32 < nb) v1 -= Ld[c * DLD + lane + 32] * yc;
      }
      if (lane < nb) wq[lane] = v0;
      if (lane + 32 < nb) wq[lane + 32] = v1;
    }
    __syncthreads();
    // rows below the block: w[row] -= sum_c F[row, k0+c] y[c]
    for (int row = k0 + nb + tid; row < d; row += SOLVE_THREADS) {
      double acc[NRC];
#pragma unroll
      for (int q = 0; q < NRC; q++) acc[q] = 0.0;
      const double* __restrict__ fr = Fj + row + (int64_t)k0 * ld;
      int c = 0;
      for (; c + 8 <= nb; c += 8) {
        double f[8];
#pragma unroll
        for (int u = 0; u < 8; u++) f[u] = fr[(int64_t)(c + u) * ld];
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
          for (int q = 0; q < NRC; q++) acc[q] += f[u] * w[k0 + c + u + q * d];
      }
      for (; c < nb; c++) {
        double f = fr[(int64_t)c * ld];
#pragma unroll
        for (int q = 0; q < NRC; q++) acc[q] += f * w[k0 + c + q * d];
      }
#pragma unroll
      for (int q = 0; q < NRC; q++) w[row + q * d] -= acc[q];
    }
    __syncthreads();
  }
  for (int i = tid; i < s; i += SOLVE_THREADS)
    for (int q = 0; q < nr; q++) x[D.col0 + i + q * ldx] = w[i + q * d];
  double* __restrict__ uj = uvec + D.uoff * NRC;
  for (int i = tid; i < r; i += SOLVE_THREADS) {
#pragma unroll
    for (int q = 0; q < NRC; q++) uj[i + q * r] = w[s + i + q * d];
  }
}

// -------------------------------------------------------------------------------- backward solve ----
// x_C = L11^{-T} (y_C - L21' x_R), x_R gathered from the already final ancestors.
template <int NRC>
__global__ void __launch_bounds__(SOLVE_THREADS) k_bwd_level(const SnodeDesc* __restrict__ sd,
                                                             const int32_t* __restrict__ list,
                                                             const int32_t* __restrict__ rows,
                                                             const double* __restrict__ F, double* __restrict__ x,
                                                             int64_t ldx, int nr) {
  extern __shared__ __align__(16) double sm[];
  double* Ld = sm;
  double* invd = sm + 64 * DLD;
  double* w = invd + 64;
  const SnodeDesc D = sd[list[blockIdx.x]];
  const int d = D.d, s = D.s, ld = D.ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ Fj = F + D.foff;
  const int32_t* __restrict__ rw = rows + D.rows_off;
  for (int i = tid; i < d; i += SOLVE_THREADS) {
    const int64_t g = (i < s) ? (D.col0 + i) : rw[i];
#pragma unroll
    for (int q = 0; q < NRC; q++) w[i + q * d] = (q < nr) ? x[g + q * ldx] : 0.0;
  }
  __syncthreads();
  const int nblk = (s + 63) / 64;
  for (int kb = nblk - 1; kb >= 0; kb--) {
    const int k0 = kb * 64;
    const int nb = min(64, s - k0);
    for (int e = tid; e < nb * nb; e += SOLVE_THREADS) {
      int i = e % nb, j = e / nb;
      if (i >= j) Ld[j * DLD + i] = Fj[(k0 + i) + (int64_t)(k0 + j) * ld];
    }
    __syncthreads();
    if (tid < nb) invd[tid] = 1.0 / Ld[tid * DLD + tid];
    // t_c = w_c - sum_{row >= k0+nb} F[row, k0+c] w[row]: one warp per column, lanes stride the rows
    for (int c = warp; c < nb; c += SOLVE_THREADS / 32) {
      double acc[NRC];
#pragma unroll
      for (int q = 0; q < NRC; q++) acc[q] = 0.0;
      const double* __restrict__ fc = Fj + (int64_t)(k0 + c) * ld;
      for (int row = k0 + nb + lane; row < d; row += 32) {
        double f = fc[row];
#pragma unroll
        for (int q = 0; q < NRC; q++) acc[q] += f * w[row + q * d];
      }
#pragma unroll
      for (int q = 0; q < NRC; q++) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) w[k0 + c + q * d] -= v;
      }
    }
    __syncthreads();
    // diagonal block: L' x = t, column-oriented elimination from the last column
    if (warp < NRC && warp < nr) {
      double* wq = w + warp * d + k0;
      double v0 = (lane < nb) ? wq[lane] : 0.0, v1 = (lane + 32 < nb) ? wq[lane + 32] : 0.0;
      for (int c = nb - 1; c >= 0; c--) {
        double xc = __shfl_sync(0xffffffffu, (c < 32) ? v0 : v1, c & 31) * invd[c];
        if (c < 32) {
          if (lane == c) v0 = xc;
        } else {
          if (lane == c - 32) v1 = xc;
          if (lane + 32 < c) v1 -= Ld[(lane + 32) * DLD + c] * xc;
        }
        if (lane < c && lane < nb) v0 -= Ld[lane * DLD + c] * xc;
      }
      if (lane < nb) wq[lane] = v0;
      if (lane + 32 < nb) wq[lane + 32] = v1;
    }
    __syncthreads();
  }
  for (int i = tid; i < s; i += SOLVE_THREADS)
    for (int q = 0; q < nr; q++) x[D.col0 + i + q * ldx] = w[i + q * d];
}

// --------------------------------------------------------------------------------------- SpMV/SpMM ----
// y = alpha * G x + beta * y for a matrix stored by rows of G (ptr/idx/val): 8 lanes per row.
__global__ void k_spmv_rows(int64_t nrows, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                            const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                            double alpha, double beta) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  double acc = 0.0;
  if (row < nrows) {
    for (int64_t p = ptr[row] + sub; p < ptr[row + 1]; p += 8) acc += val[p] * x[idx[p]];
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (row < nrows && sub == 0) y[row] = alpha * acc + (beta != 0.0 ? beta * y[row] : 0.0);
}

// RBMC accumulation (Siden et al. 2018): for row i, with samples X node-major (X[k + j*ldk], k < nsamp):
//   var_i = 1/Q_ii + (1/nsamp) sum_k ( sum_{j != i} Q_ij x_j^(k) )^2 / Q_ii^2.   One warp per row.
__global__ void k_rbmc(int64_t n, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                       const double* __restrict__ val, const double* __restrict__ X, int64_t ldk, int nsamp,
                       double* __restrict__ var) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double qii = 0.0, sumsq = 0.0;
  for (int k0 = 0; k0 < nsamp; k0 += 32) {
    const int k = k0 + lane;
    double t = 0.0;
    for (int64_t p = ptr[row]; p < ptr[row + 1]; p++) {
      const int64_t j = idx[p];
      const double q = val[p];
      if (j == row) {
        qii = q;
      } else if (k < nsamp) {
        t += q * X[k + j * ldk];
      }
    }
    sumsq += t * t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
  if (lane == 0) var[row] = 1.0 / qii + sumsq / ((double)nsamp * qii * qii);
}

// dst[k + p*ldk] = src[p + k*lds]  (RHS-major -> node-major through a permutation: dst row = perm[p])
__global__ void k_perm_scatter_nodemajor(const double* __restrict__ src, int64_t lds, double* __restrict__ dst,
                                         int64_t ldk, const int32_t* __restrict__ perm, int64_t n, int k0, int nr) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int64_t o = perm[p];
  for (int q = 0; q < nr; q++) dst[(k0 + q) + o * ldk] = src[p + q * lds];
}

// values of a matrix re-laid out through an index map: dst[k] = src[map[k]]
__global__ void k_gather_values(const double* __restrict__ src, const int64_t* __restrict__ map, int64_t nnz,
                                double* __restrict__ dst) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) dst[k] = src[map[k]];
}

// Posterior precision values on a fixed pattern:
//   out[k] = Q-part + sum over precomputed products: out[k] = (qsrc[k] >= 0 ? Qval[qsrc[k]] : 0)
//            + sum_{t in [pptr[k], pptr[k+1])} w[prow[t]] * Aval[pa[t]] * Aval[pb[t]]
__global__ void k_postprec(int64_t nnz_out, const int64_t* __restrict__ qsrc, const double* __restrict__ Qval,
                           const int64_t* __restrict__ pptr, const int32_t* __restrict__ prow,
                           const int64_t* __restrict__ pa, const int64_t* __restrict__ pb,
                           const double* __restrict__ Aval, const double* __restrict__ wdiag, double wscalar,
                           double* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz_out) return;
  double v = qsrc[k] >= 0 ? Qval[qsrc[k]] : 0.0;
  for (int64_t t = pptr[k]; t < pptr[k + 1]; t++) {
    double w = wdiag ? wdiag[prow[t]] : wscalar;
    v += w * Aval[pa[t]] * Aval[pb[t]];
  }
  out[k] = v;
}

// sum_i a_i * b_i  ->  out (single double, atomically accumulated; out must be zeroed first)
__global__ void k_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* out) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += a[i] * b[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[i];
    atomicAdd(out, s);
  }
}

__global__ void k_axpby(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                        double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}

// log-determinant pieces: out[k] = F[diag of internal column k]
__global__ void k_diag_L(const SnodeDesc* __restrict__ sd, int nsuper, const double* __restrict__ F,
                         double* __restrict__ out) {
  int s = blockIdx.x;
  if (s >= nsuper) return;
  const SnodeDesc D = sd[s];
  for (int i = threadIdx.x; i < D.s; i += blockDim.x) out[D.col0 + i] = F[D.foff + (int64_t)i * D.ld + i];
}

// ------------------------------------------------------------------------------------ host wrappers ----
static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

cudaError_t launch_perm_gather(const double* src, int64_t lds, double* dst, int64_t ldd, const int32_t* perm,
                               int64_t n, int nrhs, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_perm_gather<<<blocks_for(n, 256), 256, 0, st>>>(src, lds, dst, ldd, perm, n, nrhs);
  return cudaGetLastError();
}
cudaError_t launch_perm_scatter(const double* src, int64_t lds, double* dst, int64_t ldd, const int32_t* perm,
                                int64_t n, int nrhs, const double* add, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_perm_scatter<<<blocks_for(n, 256), 256, 0, st>>>(src, lds, dst, ldd, perm, n, nrhs, add);
  return cudaGetLastError();
}
cudaError_t launch_perm_scatter_nodemajor(const double* src, int64_t lds, double* dst, int64_t ldk,
                                          const int32_t* perm, int64_t n, int k0, int nr, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_perm_scatter_nodemajor<<<blocks_for(n, 256), 256, 0, st>>>(src, lds, dst, ldk, perm, n, k0, nr);
  return cudaGetLastError();
}

size_t solve_smem_bytes(int maxd) { return (size_t)(64 * DLD + 64 + (size_t)maxd * SOLVE_NRC) * sizeof(double); }

cudaError_t sparse_kernels_init() {
  cudaError_t e = cudaFuncSetAttribute(k_fwd_level<SOLVE_NRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_bwd_level<SOLVE_NRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_fwd_level(const SnodeDesc* sd, const int32_t* list, int count, int maxd, const int32_t* child_idx,
                             const int32_t* relmap, const double* F, double* x, int64_t ldx, double* uvec, int nr,
                             cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  k_fwd_level<SOLVE_NRC><<<count, SOLVE_THREADS, solve_smem_bytes(maxd), st>>>(sd, list, child_idx, relmap, F, x, ldx,
                                                                               uvec, nr);
  return cudaGetLastError();
}
cudaError_t launch_bwd_level(const SnodeDesc* sd, const int32_t* list, int count, int maxd, const int32_t* rows,
                             const double* F, double* x, int64_t ldx, int nr, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  k_bwd_level<SOLVE_NRC><<<count, SOLVE_THREADS, solve_smem_bytes(maxd), st>>>(sd, list, rows, F, x, ldx, nr);
  return cudaGetLastError();
}

cudaError_t launch_spmv_rows(int64_t nrows, const int64_t* ptr, const int32_t* idx, const double* val,
                             const double* x, double* y, double alpha, double beta, cudaStream_t st) {
  if (nrows <= 0) return cudaSuccess;
  k_spmv_rows<<<blocks_for(nrows * 8, 256), 256, 0, st>>>(nrows, ptr, idx, val, x, y, alpha, beta);
  return cudaGetLastError();
}
cudaError_t launch_rbmc(int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, const double* X,
                        int64_t ldk, int nsamp, double* var, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_rbmc<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, ptr, idx, val, X, ldk, nsamp, var);
  return cudaGetLastError();
}
cudaError_t launch_gather_values(const double* src, const int64_t* map, int64_t nnz, double* dst, cudaStream_t st) {
  if (nnz <= 0) return cudaSuccess;
  k_gather_values<<<blocks_for(nnz, 256), 256, 0, st>>>(src, map, nnz, dst);
  return cudaGetLastError();
}
cudaError_t launch_postprec(int64_t nnz_out, const int64_t* qsrc, const double* Qval, const int64_t* pptr,
                            const int32_t* prow, const int64_t* pa, const int64_t* pb, const double* Aval,
                            const double* wdiag, double wscalar, double* out, cudaStream_t st) {
  if (nnz_out <= 0) return cudaSuccess;
  k_postprec<<<blocks_for(nnz_out, 256), 256, 0, st>>>(nnz_out, qsrc, Qval, pptr, prow, pa, pb, Aval, wdiag, wscalar,
                                                       out);
  return cudaGetLastError();
}
cudaError_t launch_dot(const double* a, const double* b, int64_t n, double* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), st);
  if (e != cudaSuccess) return e;
  if (n <= 0) return cudaSuccess;
  unsigned g = blocks_for(n, 256);
  if (g > 1024) g = 1024;
  k_dot<<<g, 256, 0, st>>>(a, b, n, out);
  return cudaGetLastError();
}
cudaError_t launch_axpby(int64_t n, double a, const double* x, double b, const double* y, double* out,
                         cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_axpby<<<blocks_for(n, 256), 256, 0, st>>>(n, a, x, b, y, out);
  return cudaGetLastError();
}
cudaError_t launch_diag_L(const SnodeDesc* sd, int nsuper, const double* F, double* out, cudaStream_t st) {
  if (nsuper <= 0) return cudaSuccess;
  k_diag_L<<<nsuper, 64, 0, st>>>(sd, nsuper, F, out);
  return cudaGetLastError();
}

}  // namespace gmrfb
