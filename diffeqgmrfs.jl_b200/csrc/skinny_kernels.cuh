// Bandwidth-bound "skinny" products for block sweeps with few right-hand sides (M = nrhs <= 8):
//
//   C (M x N, node-major: m contiguous) = beta*C + alpha * X (M x K) * op(B)
//     k_skinny_nt : B stored N x K (n contiguous), C[m,n] += X[m,k] B[n,k]     (forward sweeps:  . C_i', . W_i')
//     k_skinny_nn : B stored K x N (k contiguous), C[m,n] += X[m,k] B[k,n]     (backward sweeps: . C_{i+1}, . W_i)
//
// The work is reading B once (8 N K bytes, or half of it for a triangular W): every SM streams its share with eight
// independent 8-byte loads per thread in flight; the arithmetic (2 M flops per loaded value) is free.  The tile
// engine's GEMM would use N/64 CTAs with one 64-column stripe each for such a product and leave most of HBM idle.
//
//   nt: thread = one row n of B, CTA = 256 rows x one K slice; the K slices of a row chunk are combined by the last
//       CTA to finish (per-chunk arrival counter), always in slice order => bit-reproducible.
//   nn: warp = one column n of B (contiguous in k), lanes stride k, shuffle-tree reduction.
//
// Task: a = X (lda), b = B (ldb), c = C (ldc), M, N, K, alpha, beta, aux0 = number of K slices (nt),
//       flags TF_BUPP / TF_BLOW = B is lower triangular (skip the structurally zero part).
// Workspace (Arenas::dinv): [ 1024 arrival counters (as int32, zero when idle) | partial sums: slices x M x N ].
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"
#include "tasks.hpp"

namespace gmrfb {

constexpr int SK_ROWS = 256;       // rows of B per CTA (nt)
constexpr int SK_KC = 256;         // k values staged in shared memory at a time
constexpr int SK_COUNTERS = 1024;  // arrival counters at the head of the workspace (row chunks per launch <= 1024)
constexpr int SK_WS_HEAD = SK_COUNTERS / 2;  // doubles occupied by the counters

template <int NR>
__global__ void __launch_bounds__(SK_ROWS) k_skinny_nt(const Task* __restrict__ tasks, Arenas ar) {
  __shared__ double xs[SK_KC][NR];
  __shared__ int last;
  const Task T = tasks[0];
  const int M = T.M, N = T.N, K = T.K, KS = T.aux0;
  const double* __restrict__ X = ar.p[(T.flags >> TF_A_SHIFT) & 3] + T.a;
  const double* __restrict__ B = ar.p[(T.flags >> TF_B_SHIFT) & 3] + T.b;
  double* __restrict__ C = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  int* __restrict__ counters = reinterpret_cast<int*>(ar.dinv);
  double* __restrict__ part = ar.dinv + SK_WS_HEAD;
  const int tid = threadIdx.x, chunk = blockIdx.x / KS, ks = blockIdx.x - chunk * KS;
  const int n = chunk * SK_ROWS + tid;
  // K range of this slice; a lower-triangular B (N x K) is zero for k > n
  const int kend_all = (T.flags & TF_BUPP) ? min(K, chunk * SK_ROWS + SK_ROWS) : K;
  const int kper = ((kend_all + KS - 1) / KS + 7) & ~7;
  const int k0 = ks * kper, k1 = min(kend_all, k0 + kper);
  double acc[NR];
#pragma unroll
  for (int m = 0; m < NR; m++) acc[m] = 0.0;
  const double* __restrict__ bp = B + n;
  for (int kc = k0; kc < k1; kc += SK_KC) {
    const int kn = min(SK_KC, k1 - kc);
    __syncthreads();
    for (int e = tid; e < ((kn + 7) & ~7) * NR; e += SK_ROWS) {  // zero rows pad the last group of eight
      const int kk = e / NR, m = e - kk * NR;
      xs[kk][m] = (m < M && kk < kn) ? X[m + (int64_t)(kc + kk) * T.lda] : 0.0;
    }
    __syncthreads();
    if (n < N) {
      for (int kk = 0; kk < kn; kk += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (kk + u < kn) ? bp[(int64_t)(kc + kk + u) * T.ldb] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
          for (int m = 0; m < NR; m++) acc[m] += v[u] * xs[(kk + u) & (SK_KC - 1)][m];
      }
    }
  }
  if (KS == 1) {
    if (n < N)
#pragma unroll
      for (int m = 0; m < NR; m++)
        if (m < M) {
          double* c = C + m + (int64_t)n * T.ldc;
          *c = (T.beta != 0.0 ? T.beta * *c : 0.0) + T.alpha * acc[m];
        }
    return;
  }
  if (n < N)
#pragma unroll
    for (int m = 0; m < NR; m++) part[((int64_t)ks * NR + m) * N + n] = acc[m];
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(&counters[chunk], 1);
    last = (prev == KS - 1);
    if (last) counters[chunk] = 0;  // idle again for the next launch on this stream
  }
  __syncthreads();
  if (!last || n >= N) return;
  __threadfence();
#pragma unroll
  for (int m = 0; m < NR; m++) {
    if (m >= M) continue;
    double s = 0.0;
    for (int q = 0; q < KS; q++) s += __ldcg(&part[((int64_t)q * NR + m) * N + n]);
    double* c = C + m + (int64_t)n * T.ldc;
    *c = (T.beta != 0.0 ? T.beta * *c : 0.0) + T.alpha * s;
  }
}

template <int NR>
__global__ void __launch_bounds__(256) k_skinny_nn(const Task* __restrict__ tasks, Arenas ar) {
  __shared__ double xs[SK_KC][NR];
  const Task T = tasks[0];
  const int M = T.M, N = T.N, K = T.K;
  const double* __restrict__ X = ar.p[(T.flags >> TF_A_SHIFT) & 3] + T.a;
  const double* __restrict__ B = ar.p[(T.flags >> TF_B_SHIFT) & 3] + T.b;
  double* __restrict__ C = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x * 8 + warp;
  // a lower-triangular B (K x N) is zero for k < n: the CTA's eight columns start at its first column, rounded down
  const int kbeg = (T.flags & TF_BLOW) ? ((blockIdx.x * 8) & ~(SK_KC - 1)) : 0;
  double acc[NR];
#pragma unroll
  for (int m = 0; m < NR; m++) acc[m] = 0.0;
  const double* __restrict__ bp = B + (int64_t)min(n, N - 1) * T.ldb;
  for (int kc = kbeg; kc < K; kc += SK_KC) {
    const int kn = min(SK_KC, K - kc);
    __syncthreads();
    for (int e = tid; e < SK_KC * NR; e += 256) {  // rows past kn are zero
      const int kk = e / NR, m = e - kk * NR;
      xs[kk][m] = (m < M && kk < kn) ? X[m + (int64_t)(kc + kk) * T.lda] : 0.0;
    }
    __syncthreads();
    double v[SK_KC / 32];
#pragma unroll
    for (int u = 0; u < SK_KC / 32; u++) {
      const int kk = lane + 32 * u;
      v[u] = (kk < kn) ? bp[kc + kk] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < SK_KC / 32; u++)
#pragma unroll
      for (int m = 0; m < NR; m++) acc[m] += v[u] * xs[lane + 32 * u][m];
  }
#pragma unroll
  for (int m = 0; m < NR; m++) {
    double s = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && m < M && n < N) {
      double* c = C + m + (int64_t)n * T.ldc;
      *c = (T.beta != 0.0 ? T.beta * *c : 0.0) + T.alpha * s;
    }
  }
}

}  // namespace gmrfb
