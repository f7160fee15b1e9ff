// Dense tile engine for sm_100a: grouped (batched, variable-size) FP64 kernels driven by Task lists.
//
//  k_gemm<TA,TB>   128x128x16 tiles, 3-stage cp.async pipeline, FP64 tensor cores (DMMA m8n8k4 via
//                  mma.sync.aligned.m8n8k4.f64 — tcgen05 has no f64 kind; measured 37.1 TFLOP/s peak on B200,
//                  profiles/r01_fp64_probe.txt).  Used for every SYRK/GEMM of the multifrontal factorisation,
//                  selected inversion and the block-tridiagonal factor.
//  k_potrf64       Cholesky of one <=64x64 diagonal block per CTA in shared memory.
//  k_trsm_rlt/rln  X <- X L^{-T} / X L^{-1} with a <=64x64 triangle: one matrix row per thread in registers.
//  k_extend_add, k_gather_sym, ...  index-mapped assembly kernels of the multifrontal method.
//
// No CPU fallback exists for any of these; the host only builds task lists (plan.cpp) and launches.
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "kernels.hpp"
#include "sparse_kernels.hpp"
#include "tasks.hpp"

namespace gmrfb {

// ------------------------------------------------------------------------------------------ helpers ----
__device__ __forceinline__ int find_task(const Task* __restrict__ tasks, int ntasks, int cta) {
  int lo = 0, hi = ntasks - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (tasks[mid].tile0 <= cta)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 8 : 0;  // src-size 0 => the 8 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// lower-triangular tile index -> (row tile, col tile)
__device__ __forceinline__ void tri_decode(int t, int& ti, int& tj) {
  int r = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
  while ((r + 1) * (r + 2) / 2 <= t) r++;
  while (r * (r + 1) / 2 > t) r--;
  ti = r;
  tj = t - r * (r + 1) / 2;
}

// --------------------------------------------------------------------------------------------- GEMM ----
// C = beta*C + alpha*op(A)*op(B);  all column-major.
//   TA=false: A is M x K (m contiguous)     TA=true: A is K x M (k contiguous), used as A'
//   TB=false: B is N x K (n contiguous), used as B'   TB=true: B is K x N (k contiguous)
#ifndef GMRFB_GEMM_STAGES
#define GMRFB_GEMM_STAGES 3
#endif
constexpr int G_STAGES = GMRFB_GEMM_STAGES;
constexpr int G_LDNA = GEMM_BM + 4;  // A tile, [k][m] layout, +4 doubles: conflict-free 64-bit fragment loads
constexpr int G_LDNB = GEMM_BN + 4;  // B tile, [k][n] layout
constexpr int G_LDT = GEMM_BK + 4;   // [m][k] / [n][k] layouts
constexpr int G_A_STAGE_N = GEMM_BK * G_LDNA, G_A_STAGE_T = GEMM_BM * G_LDT;
constexpr int G_B_STAGE_N = GEMM_BK * G_LDNB, G_B_STAGE_T = GEMM_BN * G_LDT;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 2) k_gemm(const Task* __restrict__ tasks, int ntasks, Arenas ar) {
  extern __shared__ __align__(16) double smem[];
  constexpr int BM = GEMM_BM, BN = GEMM_BN, BK = GEMM_BK;
  constexpr int A_STAGE = TA ? G_A_STAGE_T : G_A_STAGE_N;
  constexpr int B_STAGE = TB ? G_B_STAGE_T : G_B_STAGE_N;
  double* As = smem;
  double* Bs = smem + G_STAGES * A_STAGE;

  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int local = blockIdx.x - T.tile0;
  const int M = T.M, N = T.N, K = T.K;
  const int ntn = (N + BN - 1) / BN;
  int tm, tn;
  if (T.flags & TF_TRI) {
    // row tile tm owns min(2*tm + 2, ntn) column tiles
    int rem = local;
    tm = 0;
    for (;;) {
      const int w = min(2 * tm + 2, ntn);
      if (rem < w) break;
      rem -= w;
      tm++;
    }
    tn = rem;
  } else {
    tm = local / ntn;
    tn = local % ntn;
  }
  const int m0 = tm * BM, n0 = tn * BN;
  const double* __restrict__ A = ar.p[(T.flags >> TF_A_SHIFT) & 3] + T.a;
  const double* __restrict__ B = ar.p[(T.flags >> TF_B_SHIFT) & 3] + T.b;
  double* __restrict__ C = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int lda = T.lda, ldb = T.ldb, ldc = T.ldc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Interleaved ownership of the 16 x 8 grid of 8x8 MMA sub-tiles: warp (wm, wn) owns row sub-tiles im*4 + wm and
  // column sub-tiles in*2 + wn.  Sub-tiles outside the problem (or above the diagonal of a triangular result) are
  // skipped with warp-uniform predicates, and the interleaving keeps the remaining work balanced across warps, so
  // partially filled tiles (small fronts in a batched launch) cost only their useful 8x8 blocks of DMMA issue.
  const int wm = warp & 3, wn = warp >> 2;
  unsigned active = 0;
#pragma unroll
  for (int im = 0; im < 4; im++)
#pragma unroll
    for (int in = 0; in < 4; in++) {
      const int r0 = m0 + (im * 4 + wm) * 8, c0 = n0 + (in * 2 + wn) * 8;
      bool on = (r0 < M) && (c0 < N);
      if ((T.flags & TF_TRI) && c0 > r0 + 7) on = false;
      if (on) active |= 1u << (im * 4 + in);
    }

  auto load_stage = [&](int stage, int k0) {
    double* as = As + stage * A_STAGE;
    double* bs = Bs + stage * B_STAGE;
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; i++) {
      int e = tid + i * 256;
      if (!TA) {
        int m = e & (BM - 1), kk = e >> 7;
        int gm = m0 + m, gk = k0 + kk;
        bool ok = (gm < M) && (gk < K);
        cp_async8(as + kk * G_LDNA + m, ok ? A + gm + (int64_t)gk * lda : A, ok);
      } else {
        int kk = e & (BK - 1), m = e / BK;
        int gm = m0 + m, gk = k0 + kk;
        bool ok = (gm < M) && (gk < K);
        cp_async8(as + m * G_LDT + kk, ok ? A + gk + (int64_t)gm * lda : A, ok);
      }
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; i++) {
      int e = tid + i * 256;
      if (!TB) {
        int n = e & (BN - 1), kk = e / BN;
        int gn = n0 + n, gk = k0 + kk;
        bool ok = (gn < N) && (gk < K);
        cp_async8(bs + kk * G_LDNB + n, ok ? B + gn + (int64_t)gk * ldb : B, ok);
      } else {
        int kk = e & (BK - 1), n = e / BK;
        int gn = n0 + n, gk = k0 + kk;
        bool ok = (gn < N) && (gk < K);
        cp_async8(bs + n * G_LDT + kk, ok ? B + gk + (int64_t)gn * ldb : B, ok);
      }
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nkt = (K + BK - 1) / BK;
  const int kt0 = (TA && (T.flags & TF_KLOW)) ? min(m0 / BK, nkt) : 0;  // A' lower triangular: rows k < m0 are zero
#pragma unroll
  for (int s = 0; s < G_STAGES - 1; s++) {
    if (kt0 + s < nkt) load_stage((kt0 + s) % G_STAGES, (kt0 + s) * BK);
    cp_async_commit();
  }
  const int lr = lane >> 2, lc = lane & 3;
  // Two instances of the main loop: full tiles run an unpredicated DMMA stream (no per-instruction predicate /
  // reconvergence overhead); partially filled tiles skip inactive 8x8 sub-tiles with warp-uniform predicates.
  auto main_loop = [&](auto full_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
    for (int kt = kt0; kt < nkt; kt++) {
      cp_async_wait<G_STAGES - 2>();
      __syncthreads();
      {
        int nk = kt + G_STAGES - 1;
        if (nk < nkt) load_stage(nk % G_STAGES, nk * BK);
        cp_async_commit();
      }
      const double* as = As + (kt % G_STAGES) * A_STAGE;
      const double* bs = Bs + (kt % G_STAGES) * B_STAGE;
#pragma unroll
      for (int kb = 0; kb < BK; kb += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int im = 0; im < 4; im++) {
          const int rr = (im * 4 + wm) * 8 + lr;
          af[im] = TA ? as[rr * G_LDT + kb + lc] : as[(kb + lc) * G_LDNA + rr];
        }
#pragma unroll
        for (int in = 0; in < 4; in++) {
          const int cc = (in * 2 + wn) * 8 + lr;
          bf[in] = TB ? bs[cc * G_LDT + kb + lc] : bs[(kb + lc) * G_LDNB + cc];
        }
#pragma unroll
        for (int im = 0; im < 4; im++)
#pragma unroll
          for (int in = 0; in < 4; in++)
            if (FULL || (active & (1u << (im * 4 + in)))) dmma884(acc[im][in][0], acc[im][in][1], af[im], bf[in]);
      }
    }
  };
  if (active == 0xffffu)
    main_loop(std::true_type{});
  else
    main_loop(std::false_type{});
  cp_async_wait<0>();

  const bool tri = (T.flags & TF_TRI) != 0;
  const double alpha = T.alpha, beta = T.beta;
  // Epilogue: C = beta*C + alpha*acc.  The read-modify-write is done one row sub-tile (8 values per thread) at a
  // time with all loads issued before the first store, so the global-load latency is paid once per group instead
  // of once per element (the compiler cannot reorder loads across stores to the same array by itself).
#pragma unroll
  for (int im = 0; im < 4; im++) {
    const int row = m0 + (im * 4 + wm) * 8 + lr;
    double cv[4][2];
    bool ok[4][2];
#pragma unroll
    for (int in = 0; in < 4; in++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int col = n0 + (in * 2 + wn) * 8 + 2 * lc + h;
        ok[in][h] = (active & (1u << (im * 4 + in))) && row < M && col < N && (!tri || row >= col);
        cv[in][h] = 0.0;
        if (ok[in][h] && beta != 0.0) cv[in][h] = C[row + (int64_t)col * ldc];
      }
#pragma unroll
    for (int in = 0; in < 4; in++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int col = n0 + (in * 2 + wn) * 8 + 2 * lc + h;
        if (ok[in][h]) C[row + (int64_t)col * ldc] = beta * cv[in][h] + alpha * acc[im][in][h];
      }
  }
}

// -------------------------------------------------------------------------------------------- POTRF ----
// In-place Cholesky of an n x n (n <= 64) diagonal block; only the lower triangle is read and written.
// The block sits in shared memory (identity padding up to 64).  Columns are processed in panels of 8: threads
// 0..63 own one matrix row each and keep their 8 panel entries in registers (the 8x8 diagonal block is kept
// symmetric-full so the pivot row broadcast yields l(k,c) directly), one rsqrt per column on the critical path;
// the rank-8 trailing update then runs on all 128 threads.  Small code (no full unrolling): it has to run at
// instruction-cache speed because it sits on the dependent chain of every blocked factorisation.
// A non-positive or NaN pivot records (aux0 + j) in *info (minimum over all failures) and poisons the block.
constexpr int P_LD = 65;
__global__ void __launch_bounds__(128) k_potrf64(const Task* __restrict__ tasks, int ntasks, Arenas ar,
                                                 int* __restrict__ info) {
  __shared__ double S[64 * P_LD];
  __shared__ double pr[8];
  const Task T = tasks[blockIdx.x];
  const int n = T.M, lda = T.lda;
  double* __restrict__ A = ar.p[(T.flags >> TF_A_SHIFT) & 3] + T.a;
  const int tid = threadIdx.x, i = tid & 63, half = tid >> 6;
  {
    double v[32];
#pragma unroll
    for (int u = 0; u < 32; u++) {
      const int k = half + 2 * u;
      v[u] = (i == k) ? 1.0 : 0.0;
      if (i < n && k <= i) v[u] = A[i + (int64_t)k * lda];
    }
#pragma unroll
    for (int u = 0; u < 32; u++) S[(half + 2 * u) * P_LD + i] = v[u];
  }
  __syncthreads();
  bool bad = false;
  for (int jb = 0; jb < 8; jb++) {
    const int j0 = jb * 8;
    if (j0 >= n) break;
    double p[8];
    if (half == 0) {
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const int col = j0 + c;
        p[c] = (i >= col) ? S[col * P_LD + i] : S[i * P_LD + col];  // mirror inside the diagonal block / above it unused
      }
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
      if (half == 0 && i == j0 + c) {
#pragma unroll
        for (int c2 = c; c2 < 8; c2++) pr[c2] = p[c2];
      }
      __syncthreads();
      if (half == 0) {
        const double d = pr[c];
        double dinv;
        if (d > 0.0) {
          dinv = rsqrt(d);
        } else {
          dinv = nan("");
          if (i == j0 + c && j0 + c < n) bad = true;
        }
        const double lic = (i == j0 + c) ? d * dinv : p[c] * dinv;
        p[c] = lic;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; c2++) p[c2] -= lic * (pr[c2] * dinv);
      }
      __syncthreads();
    }
    if (half == 0 && i >= j0) {
#pragma unroll
      for (int c = 0; c < 8; c++)
        if (i >= j0 + c) S[(j0 + c) * P_LD + i] = p[c];
    }
    __syncthreads();
    // rank-8 update of the trailing block: S(i,k) -= sum_c l(i,c) l(k,c), j0+8 <= k <= i; two threads per row
    if (half == 1) {
#pragma unroll
      for (int c = 0; c < 8; c++) p[c] = S[(j0 + c) * P_LD + i];
    }
    for (int k = j0 + 8 + half; k <= i; k += 2) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < 8; c++) acc += p[c] * S[(j0 + c) * P_LD + k];
      S[k * P_LD + i] -= acc;
    }
    __syncthreads();
  }
  if (bad) atomicMin(info, T.aux0 + i);
  {
#pragma unroll 8
    for (int u = 0; u < 32; u++) {
      const int k = half + 2 * u;
      if (i < n && k <= i) A[i + (int64_t)k * lda] = S[k * P_LD + i];
    }
  }
}

// --------------------------------------------------------------------------------------------- TRSM ----
// One CTA = TRSM_ROWS rows of X, one row per thread; the <=64x64 triangle sits in shared memory padded to 64x64
// with an identity.  The row is processed in 8-column blocks held in registers while the already solved part of
// the row is parked in shared memory, so the code stays small (instruction-cache resident) and every inner
// product is an unrolled 8x8 micro-kernel with broadcast reads of the triangle.
constexpr int T_LDX = 65;
template <bool TRANS>
__global__ void __launch_bounds__(TRSM_ROWS) k_trsm(const Task* __restrict__ tasks, int ntasks, Arenas ar) {
  extern __shared__ __align__(16) double tsm[];
  double* Ls = tsm;                 // 64 x 64, column-major
  double* invd = Ls + 64 * 64;      // 64
  double* Xs = invd + 64;           // TRSM_ROWS x T_LDX: row r of this CTA at Xs[r * T_LDX + c]
  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int M = T.M, N = T.N;
  const double* __restrict__ L = ar.p[(T.flags >> TF_B_SHIFT) & 3] + T.b;
  double* __restrict__ X = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int ldl = T.ldb, ldx = T.ldc;
  const int tid = threadIdx.x;
  {
    const int r = tid & 63, c0 = tid >> 6;
    double v[32];
#pragma unroll
    for (int u = 0; u < 32; u++) {
      const int c = c0 + 2 * u;
      v[u] = (r == c) ? 1.0 : 0.0;
      if (r < N && c < N && r >= c) v[u] = L[r + (int64_t)c * ldl];
    }
#pragma unroll
    for (int u = 0; u < 32; u++) Ls[(c0 + 2 * u) * 64 + r] = v[u];
  }
  const int row = (blockIdx.x - T.tile0) * TRSM_ROWS + tid;
  const bool live = row < M;
  double* xs = Xs + tid * T_LDX;
  {
#pragma unroll 16
    for (int j = 0; j < 64; j++) xs[j] = (live && j < N) ? X[row + (int64_t)j * ldx] : 0.0;
  }
  __syncthreads();
  if (tid < 64) invd[tid] = 1.0 / Ls[tid * 64 + tid];
  __syncthreads();
  if (TRANS) {
    // x L' = b:  blocks ascending.  x_J = (b_J - sum_{K<J} x_K L[J,K]') L[J,J]^{-T}
    for (int jb = 0; jb < 8; jb++) {
      const int j0 = jb * 8;
      if (j0 >= N) break;
      double xb[8];
#pragma unroll
      for (int c = 0; c < 8; c++) xb[c] = xs[j0 + c];
      for (int k = 0; k < j0; k++) {
        const double xk = xs[k];
#pragma unroll
        for (int c = 0; c < 8; c++) xb[c] -= xk * Ls[k * 64 + j0 + c];  // L[j0+c][k]
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        xb[c] *= invd[j0 + c];
#pragma unroll
        for (int c2 = c + 1; c2 < 8; c2++) xb[c2] -= xb[c] * Ls[(j0 + c) * 64 + j0 + c2];
      }
#pragma unroll
      for (int c = 0; c < 8; c++) xs[j0 + c] = xb[c];
    }
  } else {
    // x L = b:  blocks descending.  x_J = (b_J - sum_{K>J} x_K L[K,J]) L[J,J]^{-1}
    for (int jb = 7; jb >= 0; jb--) {
      const int j0 = jb * 8;
      if (j0 >= N) continue;
      double xb[8];
#pragma unroll
      for (int c = 0; c < 8; c++) xb[c] = xs[j0 + c];
      for (int k = 63; k >= j0 + 8; k--) {
        const double xk = xs[k];
#pragma unroll
        for (int c = 0; c < 8; c++) xb[c] -= xk * Ls[(j0 + c) * 64 + k];  // L[k][j0+c]
      }
#pragma unroll
      for (int c = 7; c >= 0; c--) {
        xb[c] *= invd[j0 + c];
#pragma unroll
        for (int c2 = 0; c2 < c; c2++) xb[c2] -= xb[c] * Ls[(j0 + c2) * 64 + j0 + c];  // L[j0+c][j0+c2]
      }
#pragma unroll
      for (int c = 0; c < 8; c++) xs[j0 + c] = xb[c];
    }
  }
  if (!live) return;
  const double sgn = (T.flags & TF_NEG) ? -1.0 : 1.0;
#pragma unroll 16
  for (int j = 0; j < 64; j++)
    if (j < N) X[row + (int64_t)j * ldx] = sgn * xs[j];
}

// ------------------------------------------------------------------------------ multifrontal assembly ----
// P[rel[i], rel[j]] += U[i, j] for i >= j (child update matrix into the parent front).  One CTA = one
// 64x64 tile of the lower triangle of U.  Children of one parent are issued in separate launches, so no two
// CTAs of a launch touch the same parent entry: the assembly is deterministic and atomic-free.
__global__ void __launch_bounds__(256) k_extend_add(const Task* __restrict__ tasks, int ntasks, Arenas ar,
                                                    const int32_t* __restrict__ relmap) {
  __shared__ int32_t ri[EA_TILE], rj[EA_TILE];
  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  int ti, tj;
  tri_decode(blockIdx.x - T.tile0, ti, tj);
  const int M = T.M;
  const double* __restrict__ U = ar.p[(T.flags >> TF_A_SHIFT) & 3] + T.a;
  double* __restrict__ P = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int32_t* rel = relmap + (((int64_t)T.aux1 << 32) | (uint32_t)T.aux0);
  const int tid = threadIdx.x;
  const int i0 = ti * EA_TILE, j0 = tj * EA_TILE;
  if (tid < EA_TILE) {
    ri[tid] = (i0 + tid < M) ? rel[i0 + tid] : 0;
  } else if (tid < 2 * EA_TILE) {
    int t = tid - EA_TILE;
    rj[t] = (j0 + t < M) ? rel[j0 + t] : 0;
  }
  __syncthreads();
  const int li = tid & 63;
  const int i = i0 + li;
  if (i >= M) return;
  const int64_t pr = ri[li];
  for (int lj = tid >> 6; lj < EA_TILE; lj += 4) {
    const int j = j0 + lj;
    if (j > i || j >= M) continue;
    P[pr + (int64_t)rj[lj] * T.ldc] += U[i + (int64_t)j * T.lda];
  }
}

// Zc[i, j] = Zp[rel[i], rel[j]] (Zp symmetric, lower triangle valid), full square written.
__global__ void __launch_bounds__(256) k_gather_sym(const Task* __restrict__ tasks, int ntasks, Arenas ar,
                                                    const int32_t* __restrict__ relmap) {
  __shared__ int32_t ri[EA_TILE], rj[EA_TILE];
  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int M = T.M;
  const int nt = (M + EA_TILE - 1) / EA_TILE;
  const int local = blockIdx.x - T.tile0;
  const int ti = local % nt, tj = local / nt;
  const double* __restrict__ Zp = ar.p[(T.flags >> TF_A_SHIFT) & 3] + T.a;
  double* __restrict__ Zc = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int32_t* rel = relmap + (((int64_t)T.aux1 << 32) | (uint32_t)T.aux0);
  const int tid = threadIdx.x;
  const int i0 = ti * EA_TILE, j0 = tj * EA_TILE;
  if (tid < EA_TILE) {
    ri[tid] = (i0 + tid < M) ? rel[i0 + tid] : 0;
  } else if (tid < 2 * EA_TILE) {
    int t = tid - EA_TILE;
    rj[t] = (j0 + t < M) ? rel[j0 + t] : 0;
  }
  __syncthreads();
  const int li = tid & 63;
  const int i = i0 + li;
  if (i >= M) return;
  const int64_t a = ri[li];
  for (int lj = tid >> 6; lj < EA_TILE; lj += 4) {
    const int j = j0 + lj;
    if (j >= M) continue;
    const int64_t b = rj[lj];
    const double v = (a >= b) ? Zp[a + b * T.lda] : Zp[b + a * T.lda];
    Zc[i + (int64_t)j * T.ldc] = v;
  }
}

// ------------------------------------------------------------------------ fused small-front kernels ----
// Fronts of order d <= SMALL_FRONT_MAX are processed by ONE CTA entirely in shared memory (the front is staged
// once, every operation of the multifrontal step runs on-chip, results are written once):
//   factor : load the assembled panel, extend-add the children's update matrices (fixed order), partial
//            Cholesky of the first s columns with the full trailing update, write L and the update matrix.
//   selinv : gather Z_RR from the parent's inverse front, then the dense Takahashi recurrence column by column
//            (Z_ij = (delta_ij / L_jj - sum_{k>j} L_kj Z_ik) / L_jj), write the inverse front and its diagonal.
// Task encoding: aux0 = supernode index.
__global__ void __launch_bounds__(256) k_front_factor_small(const Task* __restrict__ tasks, Arenas ar,
                                                            const SnodeDesc* __restrict__ sd,
                                                            const int32_t* __restrict__ child_idx,
                                                            const int32_t* __restrict__ relmap,
                                                            int* __restrict__ info) {
  extern __shared__ __align__(16) double S[];
  __shared__ int32_t rel[SMALL_FRONT_MAX];
  const SnodeDesc D = sd[tasks[blockIdx.x].aux0];
  const int d = D.d, s = D.s, ldg = D.ld;
  const int lds = d | 1;
  double* __restrict__ F = ar.p[0] + D.foff;
  const int tid = threadIdx.x, ti = tid & 63, tq = tid >> 6;
  // stage: panel columns from the arena, update-matrix part starts from zero
  for (int c = tq; c < d; c += 4)
    for (int i = ti; i < d; i += 64) S[c * lds + i] = (c < s && i >= c) ? F[(int64_t)c * ldg + i] : 0.0;
  __syncthreads();
  for (int ci = 0; ci < D.nchild; ci++) {
    const SnodeDesc C = sd[child_idx[D.child0 + ci]];
    const int rc = C.d - C.s;
    const double* __restrict__ U = ar.p[0] + C.foff + (int64_t)C.s * C.ld + C.s;
    const int32_t* __restrict__ rl = relmap + C.rows_off + C.s;
    if (rc <= SMALL_FRONT_MAX) {
      for (int i = tid; i < rc; i += 256) rel[i] = rl[i];
      __syncthreads();
      for (int j = tq; j < rc; j += 4) {
        const int pj = rel[j];
        for (int i = j + ti - (j & 63) + ((ti < (j & 63)) ? 64 : 0); i < rc; i += 64)
          S[pj * lds + rel[i]] += U[i + (int64_t)j * C.ld];
      }
    } else {
      // a child with a long boundary: its rows still all map inside this (small) front
      for (int j = tq; j < rc; j += 4) {
        const int pj = rl[j];
        for (int i = j + ti - (j & 63) + ((ti < (j & 63)) ? 64 : 0); i < rc; i += 64)
          S[pj * lds + rl[i]] += U[i + (int64_t)j * C.ld];
      }
    }
    __syncthreads();
  }
  for (int j = 0; j < s; j++) {
    const double dj = S[j * lds + j];
    __syncthreads();
    double ljj;
    if (dj > 0.0) {
      ljj = sqrt(dj);
    } else {
      ljj = nan("");
      if (tid == 0) atomicMin(info, D.col0 + j);
    }
    const double inv = 1.0 / ljj;
    for (int i = j + tid; i < d; i += 256) S[j * lds + i] = (i == j) ? ljj : S[j * lds + i] * inv;
    __syncthreads();
    for (int k = j + 1 + tq; k < d; k += 4) {
      const double lk = S[j * lds + k];
      for (int i = k + ti - (k & 63) + ((ti < (k & 63)) ? 64 : 0); i < d; i += 64) S[k * lds + i] -= S[j * lds + i] * lk;
    }
    __syncthreads();
  }
  for (int c = tq; c < d; c += 4)
    for (int i = c + ti - (c & 63) + ((ti < (c & 63)) ? 64 : 0); i < d; i += 64) F[(int64_t)c * ldg + i] = S[c * lds + i];
}

__global__ void __launch_bounds__(256) k_front_selinv_small(const Task* __restrict__ tasks, Arenas ar,
                                                            const SnodeDesc* __restrict__ sd,
                                                            const int32_t* __restrict__ relmap,
                                                            const int32_t* __restrict__ sparent,
                                                            double* __restrict__ zdiag) {
  extern __shared__ __align__(16) double Z[];
  __shared__ double lcol[SMALL_FRONT_MAX];
  __shared__ double part[4][SMALL_FRONT_MAX];
  __shared__ int32_t rel[SMALL_FRONT_MAX];
  __shared__ double red[8];
  const int sidx = tasks[blockIdx.x].aux0;
  const SnodeDesc D = sd[sidx];
  const int d = D.d, s = D.s, r = d - s, ldg = D.ld;
  const int lds = d | 1;
  const double* __restrict__ L = ar.p[0] + D.foff;
  double* __restrict__ Zg = ar.p[1] + D.foff;
  const int tid = threadIdx.x, ti = tid & 63, tq = tid >> 6, lane = tid & 31, warp = tid >> 5;
  if (r > 0) {
    const SnodeDesc P = sd[sparent[sidx]];
    const double* __restrict__ Zp = ar.p[1] + P.foff;
    const int32_t* __restrict__ rl = relmap + D.rows_off + s;
    for (int i = tid; i < r; i += 256) rel[i] = rl[i];
    __syncthreads();
    for (int j = tq; j < r; j += 4) {
      const int64_t b = rel[j];
      for (int i = ti; i < r; i += 64) {
        const int64_t a = rel[i];
        Z[(s + j) * lds + s + i] = (a >= b) ? Zp[a + b * P.ld] : Zp[b + a * P.ld];
      }
    }
  }
  __syncthreads();
  for (int j = s - 1; j >= 0; j--) {
    for (int i = j + tid; i < d; i += 256) lcol[i] = L[(int64_t)j * ldg + i];
    __syncthreads();
    const double inv = 1.0 / lcol[j];
    // partial sums over k = j+1+tq, step 4, for every row i > j
    for (int i = j + 1 + ti; i < d; i += 64) {
      double acc = 0.0;
      for (int k = j + 1 + tq; k < d; k += 4) acc += lcol[k] * Z[k * lds + i];
      part[tq][i] = acc;
    }
    __syncthreads();
    double dot = 0.0;
    for (int i = j + 1 + tid; i < d; i += 256) {
      const double z = -(((part[0][i] + part[1][i]) + part[2][i]) + part[3][i]) * inv;
      Z[j * lds + i] = z;
      Z[i * lds + j] = z;
      dot += lcol[i] * z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; w++) t += red[w];
      Z[j * lds + j] = (inv - t) * inv;
    }
    __syncthreads();
  }
  for (int c = tq; c < d; c += 4)
    for (int i = c + ti - (c & 63) + ((ti < (c & 63)) ? 64 : 0); i < d; i += 64) Zg[(int64_t)c * ldg + i] = Z[c * lds + i];
  for (int c = tid; c < s; c += 256) zdiag[D.col0 + c] = Z[c * lds + c];
}

// Simple element-wise task kernels: one CTA per 64x64 tile.
__global__ void __launch_bounds__(256) k_tile_op(const Task* __restrict__ tasks, int ntasks, Arenas ar, int op) {
  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int M = T.M, N = T.N;
  const int ntm = (M + 63) / 64;
  const int local = blockIdx.x - T.tile0;
  const int ti = local % ntm, tj = local / ntm;
  double* __restrict__ C = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int tid = threadIdx.x, li = tid & 63;
  const int i = ti * 64 + li;
  if (i >= M) return;
  for (int lj = tid >> 6; lj < 64; lj += 4) {
    const int j = tj * 64 + lj;
    if (j >= N) continue;
    double* p = C + i + (int64_t)j * T.ldc;
    if (op == LK_SET_IDENTITY)
      *p = (i == j) ? 1.0 : 0.0;
    else if (op == LK_SCALE)
      *p = T.alpha * (*p);
    else if (op == LK_SYMMETRIZE) {
      if (j > i) *p = C[j + (int64_t)i * T.ldc];
    }
  }
}

// In-place transpose of a square M x M matrix: one CTA per 32x32 tile pair of the lower triangle.
__global__ void __launch_bounds__(256) k_transpose(const Task* __restrict__ tasks, int ntasks, Arenas ar) {
  __shared__ double sa[32][33], sb[32][33];
  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int M = T.M;
  int ti, tj;
  tri_decode(blockIdx.x - T.tile0, ti, tj);
  double* __restrict__ C = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int tid = threadIdx.x, li = tid & 31;
  const int i = ti * 32 + li, i2 = tj * 32 + li;
  for (int lj = tid >> 5; lj < 32; lj += 8) {
    const int j = tj * 32 + lj, j2 = ti * 32 + lj;
    sa[lj][li] = (i < M && j < M) ? C[i + (int64_t)j * T.ldc] : 0.0;      // tile (ti,tj), element (li,lj)
    sb[lj][li] = (i2 < M && j2 < M) ? C[i2 + (int64_t)j2 * T.ldc] : 0.0;  // mirror tile (tj,ti), element (li,lj)
  }
  __syncthreads();
  for (int lj = tid >> 5; lj < 32; lj += 8) {
    const int j = tj * 32 + lj, j2 = ti * 32 + lj;
    if (i < M && j < M) C[i + (int64_t)j * T.ldc] = sb[li][lj];
    if (ti != tj && i2 < M && j2 < M) C[i2 + (int64_t)j2 * T.ldc] = sa[li][lj];
  }
}

// out[aux + i] = C[i,i]
__global__ void __launch_bounds__(256) k_diag_out(const Task* __restrict__ tasks, int ntasks, Arenas ar,
                                                  double* __restrict__ out) {
  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int i = (blockIdx.x - T.tile0) * 256 + threadIdx.x;
  if (i >= T.M) return;
  const double* __restrict__ C = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  out[(((int64_t)T.aux1 << 32) | (uint32_t)T.aux0) + i] = C[i + (int64_t)i * T.ldc];
}

// arena[amap[k]] = nzval[k] for every stored entry that belongs to the analysed triangle.
__global__ void k_scatter_values(const double* __restrict__ nzval, const int64_t* __restrict__ amap, int64_t nnz,
                                 double* __restrict__ arena) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  int64_t d = amap[k];
  if (d >= 0) arena[d] = nzval[k];
}

// ---------------------------------------------------------------------------------- host launchers ----
static size_t trsm_smem() { return (size_t)(64 * 64 + 64 + TRSM_ROWS * T_LDX) * sizeof(double); }
static size_t gemm_smem(bool ta, bool tb) {
  size_t a = ta ? G_A_STAGE_T : G_A_STAGE_N, b = tb ? G_B_STAGE_T : G_B_STAGE_N;
  return (a + b) * G_STAGES * sizeof(double);
}

cudaError_t kernels_init() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_gemm<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)gemm_smem(false, false));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_gemm<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)gemm_smem(false, true));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_gemm<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)gemm_smem(true, true));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_gemm<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)gemm_smem(true, false));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_trsm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_trsm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem());
  if (e != cudaSuccess) return e;
  const int small_smem = SMALL_FRONT_MAX * (SMALL_FRONT_MAX | 1) * (int)sizeof(double);
  e = cudaFuncSetAttribute(k_front_factor_small, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_front_selinv_small, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem);
  return e;
}

cudaError_t run_launch(const Launch& L, const Task* d_tasks, const Arenas& ar, const LaunchAux& aux,
                       cudaStream_t st) {
  if (L.grid <= 0 || L.ntasks <= 0) return cudaSuccess;
  const Task* t = d_tasks + L.task0;
  switch (L.kind) {
    case LK_GEMM_NT:
      k_gemm<false, false><<<L.grid, 256, gemm_smem(false, false), st>>>(t, L.ntasks, ar);
      break;
    case LK_GEMM_NN:
      k_gemm<false, true><<<L.grid, 256, gemm_smem(false, true), st>>>(t, L.ntasks, ar);
      break;
    case LK_GEMM_TN:
      k_gemm<true, true><<<L.grid, 256, gemm_smem(true, true), st>>>(t, L.ntasks, ar);
      break;
    case LK_GEMM_TT:
      k_gemm<true, false><<<L.grid, 256, gemm_smem(true, false), st>>>(t, L.ntasks, ar);
      break;
    case LK_POTRF:
      k_potrf64<<<L.grid, 128, 0, st>>>(t, L.ntasks, ar, aux.d_info);
      break;
    case LK_TRSM_RLT:
      k_trsm<true><<<L.grid, TRSM_ROWS, trsm_smem(), st>>>(t, L.ntasks, ar);
      break;
    case LK_TRSM_RLN:
      k_trsm<false><<<L.grid, TRSM_ROWS, trsm_smem(), st>>>(t, L.ntasks, ar);
      break;
    case LK_EXTEND_ADD:
      k_extend_add<<<L.grid, 256, 0, st>>>(t, L.ntasks, ar, aux.d_relmap);
      break;
    case LK_GATHER_SYM:
      k_gather_sym<<<L.grid, 256, 0, st>>>(t, L.ntasks, ar, aux.d_relmap);
      break;
    case LK_TRANSPOSE:
      k_transpose<<<L.grid, 256, 0, st>>>(t, L.ntasks, ar);
      break;
    case LK_SET_IDENTITY:
    case LK_SCALE:
    case LK_SYMMETRIZE:
      k_tile_op<<<L.grid, 256, 0, st>>>(t, L.ntasks, ar, L.kind);
      break;
    case LK_DIAG_OUT:
      k_diag_out<<<L.grid, 256, 0, st>>>(t, L.ntasks, ar, aux.d_out);
      break;
    case LK_FRONT_FACTOR_SMALL:
      k_front_factor_small<<<L.grid, 256, L.smem, st>>>(t, ar, aux.d_snodes, aux.d_child_idx, aux.d_relmap, aux.d_info);
      break;
    case LK_FRONT_SELINV_SMALL:
      k_front_selinv_small<<<L.grid, 256, L.smem, st>>>(t, ar, aux.d_snodes, aux.d_relmap, aux.d_sparent, aux.d_out);
      break;
    default:
      return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_scatter_values(const double* d_nzval, const int64_t* d_amap, int64_t nnz, double* d_arena,
                                  cudaStream_t st) {
  if (nnz <= 0) return cudaSuccess;
  int64_t grid = (nnz + 255) / 256;
  k_scatter_values<<<(unsigned)grid, 256, 0, st>>>(d_nzval, d_amap, nnz, d_arena);
  return cudaGetLastError();
}

}  // namespace gmrfb
