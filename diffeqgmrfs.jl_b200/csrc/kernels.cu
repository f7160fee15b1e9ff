// Dense tile engine for sm_100a: grouped (batched, variable-size) FP64 kernels driven by Task lists.
//
//  k_gemm2<TA,TB,CFG>  (gemm_engine.cuh) 128x64x16 / 64x64x16 tiles, cp.async pipeline, FP64 tensor cores (DMMA m8n8k4
//                  via mma.sync.aligned.m8n8k4.f64 — tcgen05 has no f64 kind; measured 37.1 TFLOP/s peak on B200,
//                  profiles/r01_fp64_probe.txt).  Used for every SYRK/GEMM of the multifrontal factorisation,
//                  selected inversion and the block-tridiagonal factor.
//  k_potrf64       Cholesky + inverse of one <=64x64 diagonal block per CTA in shared memory (DMMA panel updates).
//  k_apply_inv     X <- X L^{-T} / X L^{-1} as a DMMA product with the inverted <=64x64 triangle.
//  k_extend_add, k_gather_sym, ...  index-mapped assembly kernels of the multifrontal method.
//
// No CPU fallback exists for any of these; the host only builds task lists (plan.cpp) and launches.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "gemm_engine.cuh"
#include "kernels.hpp"
#include "skinny_kernels.cuh"
#include "sparse_kernels.hpp"
#include "tasks.hpp"

namespace gmrfb {

// ------------------------------------------------------------------------------------------ helpers ----
// Task of CTA `cta`: launches with more than one task carry a CTA -> task map in the Task-sized slots that follow
// their tasks (PlanBuilder::end), so the lookup is one load instead of a binary search over the tile0 prefix sums.
__device__ __forceinline__ int find_task(const Task* __restrict__ tasks, int ntasks, int cta) {
  if (ntasks == 1) return 0;
  return reinterpret_cast<const int32_t*>(tasks + ntasks)[cta];
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

// advance a lower-triangular tile index (ti, tj) by `step` tiles in row-major order of the triangle (0,0), (1,0), (1,1), ...
// (the per-tile sqrt-based decode cost more than the two DMMAs of a rank-8 tile update: ncu, profiles/r02_small_front_lines.md)
__device__ __forceinline__ void tri_advance(int& ti, int& tj, int step) {
  tj += step;
  while (tj > ti) {
    tj -= ti + 1;
    ti++;
  }
}
// x / n for 0 <= x < 65536 and 1 <= n < 65536 through a multiplication by ceil(2^32 / n): the staging loops of the fused
// small-front kernels index (column, row) pairs by a flat counter, and an integer division per element was a tenth of
// their issue slots
__device__ __forceinline__ uint32_t div_magic(uint32_t n) { return n <= 1 ? 0u : (uint32_t)((0x100000000ull + n - 1) / n); }
__device__ __forceinline__ int fast_div(int x, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)x, magic) : x; }

// right-looking Cholesky of the 8x8 block in registers (lower part); invd[c] = 1 / L_cc.  Returns the first failing
// column or 8.
__device__ __forceinline__ int chol8(double (&D)[8][8], double (&invd)[8]) {
  int bad = 8;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const double d = D[c][c];
    double r;
    if (d > 0.0) {
      r = rsqrt(d);
    } else {
      r = nan("");
      if (bad == 8) bad = c;
    }
    invd[c] = r;
    D[c][c] = d * r;
#pragma unroll
    for (int i = c + 1; i < 8; i++) D[i][c] *= r;
#pragma unroll
    for (int j = c + 1; j < 8; j++)
#pragma unroll
      for (int i = j; i < 8; i++) D[i][j] -= D[i][c] * D[j][c];
  }
  return bad;
}

// --------------------------------------------------------------------------------------------- GEMM ----
// The grouped DMMA GEMM engine lives in gemm_engine.cuh (k_gemm2<TA, TB, CFG>); two tile configurations are
// instantiated and chosen per launch by the plan builder (Launch::cfg):
//   GCFG_BIG   128x64 tiles, 8 warps, 3 stages, 2 CTAs/SM  - large, regular problems (33.5 TFLOP/s at n=4736, K=4096)
//   GCFG_SMALL  64x64 tiles, 2 stages, in two warp layouts chosen at launch time by the grid size:
//                 4 warps (32x32 warp tiles) at 4 CTAs/SM for launches of at least GEMM_W4_MIN_GRID CTAs,
//                 8 warps (16x32 warp tiles) at 3 CTAs/SM below (few CTAs: more warps per tile hide latency better)
// Measured with tools/probe/gemm_batched_probe.cu (profiles/r01_gemm_batched_probe.md): on batched launches over ragged
// fronts (K = 30 ... 600) two stages are 6-18 % faster than four (35 KB instead of 70 KB of shared memory per CTA leaves
// the L1 to the read-modify-write of the C tiles); four warps with 32x32 warp tiles (one LDS.128 pair feeds 16 DMMAs
// instead of 8) at 4 CTAs/SM add another 2-4 % on large grids and reach 35.1 TFLOP/s on one 4736^2 x 4096 product
// (8 warps: 33.5; cuBLAS DGEMM: 35.5).  The 128x64 configuration is only used on request (GMRFB_GEMM_BIG_MIN).
using GemmBig = GemmCfg<GEMM_TILE_M[GCFG_BIG], GEMM_TILE_N[GCFG_BIG], 4, 2, 16, 3, 2>;
using GemmSmall = GemmCfg<GEMM_TILE_M[GCFG_SMALL], GEMM_TILE_N[GCFG_SMALL], 4, 2, 16, 2, 3>;
using GemmSmall4 = GemmCfg<GEMM_TILE_M[GCFG_SMALL], GEMM_TILE_N[GCFG_SMALL], 2, 2, 16, 2, 4>;
constexpr int GEMM_W4_MIN_GRID = 1184;  // two full waves of the 4-warp layout (148 SMs x 4 CTAs x 2)

// -------------------------------------------------------------------------------------------- POTRF ----
// Cholesky AND inverse of an n x n (n <= 64) diagonal block, one CTA (8 warps) per block, everything in shared
// memory (identity padding up to 64).  The block sits on the dependent chain of every blocked factorisation, so it is
// organised for latency:
//   factor : 8-column panels.  Threads 0..63 own one matrix row each; every one of them factors the 8x8 diagonal
//            block of the panel redundantly in registers (no synchronisation inside a panel, one rsqrt per column)
//            and solves its own row against it; the rank-8 trailing update then runs as 8x8 DMMA tiles on all warps.
//            16 barriers per block instead of two per column.
//   invert : W = L^{-1} by recursive doubling — the eight 8x8 diagonal blocks in registers, then three merge levels
//            inv([A 0; B C]) = [A^{-1} 0; -C^{-1} B A^{-1}  C^{-1}], each two batched DMMA products.
// The inverse turns every later "X L^{-T}" / "X L^{-1}" with this block into a tensor-core product (k_apply_inv).
// Only <= 64x64 diagonal blocks are ever inverted; the error this adds is O(cond(L_jj) eps), as in blocked TRSMs of
// dense GPU libraries.  Task: a/lda/M = block, b/ldb = destination of W (arena B bits, or ar.dinv when TF_B_DINV),
// aux0 = global column for failure reports, TF_NOFACTOR = the block already holds L (invert only).
constexpr int PLD = 68;  // leading dimension of the 64x64 smem matrices: = 4 (mod 16) => conflict-free fragment reads

__device__ __forceinline__ double* task_b_ptr(const Task& T, const Arenas& ar) {
  return ((T.flags & TF_B_DINV) ? ar.dinv : ar.p[(T.flags >> TF_B_SHIFT) & 3]) + T.b;
}

__global__ void __launch_bounds__(256) k_potrf64(const Task* __restrict__ tasks, int ntasks, Arenas ar,
                                                 int* __restrict__ info) {
  extern __shared__ __align__(16) double psm[];
  double* S = psm;             // the block / its factor L, column-major S[c * PLD + r]
  double* W = S + 64 * PLD;    // L^{-1}
  double* Tm = W + 64 * PLD;   // scratch of the merge levels
  const Task T = tasks[blockIdx.x];
  const int n = T.M, lda = T.lda;
  double* __restrict__ A = ar.p[(T.flags >> TF_A_SHIFT) & 3] + T.a;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int e = tid + u * 256, i = e & 63, k = e >> 6;
      v[u] = (i == k) ? 1.0 : 0.0;
      if (i < n && k <= i) v[u] = A[i + (int64_t)k * lda];
    }
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int e = tid + u * 256, i = e & 63, k = e >> 6;
      S[k * PLD + i] = v[u];
      W[k * PLD + i] = 0.0;
    }
  }
  __syncthreads();
  const int nb8 = (n + 7) >> 3;
  if (!(T.flags & TF_NOFACTOR)) {
    bool bad = false;
    for (int p = 0; p < nb8; p++) {
      const int j0 = p * 8;
      const bool rowt = tid < 64 && tid >= j0;
      double x[8];
      if (rowt) {
        const int i = tid;
        double D[8][8], invd[8];
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
          for (int k = 0; k <= c; k++) D[c][k] = S[(j0 + k) * PLD + j0 + c];
#pragma unroll
        for (int c = 0; c < 8; c++) x[c] = S[(j0 + c) * PLD + i];
        const int badc = chol8(D, invd);
        if (badc < 8 && i == j0 + badc && i < n) bad = true;
        const int ii = i - j0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          x[c] *= invd[c];
#pragma unroll
          for (int j = c + 1; j < 8; j++) x[j] -= x[c] * D[j][c];
        }
#pragma unroll
        for (int c = 0; c < 8; c++)
          if (c > ii) x[c] = 0.0;  // rows of the diagonal block: L's row up to the diagonal, zero above
      }
      __syncthreads();  // every row thread has read the 8x8 diagonal block before its rows are overwritten
      if (rowt) {
#pragma unroll
        for (int c = 0; c < 8; c++) S[(j0 + c) * PLD + tid] = x[c];
      }
      __syncthreads();
      const int first = p + 1, m = nb8 - first;
      int ti0 = 0, tj0 = 0;
      tri_advance(ti0, tj0, warp);
      for (; ti0 < m; tri_advance(ti0, tj0, 8)) {
        const int ti = ti0 + first, tj = tj0 + first;
        double* cp = S + (tj * 8 + 2 * lc) * PLD + ti * 8 + lr;
        double c0 = cp[0], c1 = cp[PLD];
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
          const double a = -S[(j0 + 4 * kk + lc) * PLD + ti * 8 + lr];
          const double b = S[(j0 + 4 * kk + lc) * PLD + tj * 8 + lr];
          dmma884(c0, c1, a, b);
        }
        cp[0] = c0;
        cp[PLD] = c1;
      }
      __syncthreads();
    }
    if (bad) atomicMin(info, T.aux0 + tid);
    for (int e = tid; e < 64 * 64; e += 256) {
      const int i = e & 63, k = e >> 6;
      if (i < n && k <= i) A[i + (int64_t)k * lda] = S[k * PLD + i];
    }
  }
  // ---- W = L^{-1}: 8x8 diagonal blocks, one column per thread ----
  if (tid < 64) {
    const int bi = tid >> 3, cj = tid & 7, o = bi * 8;
    double l[8][8], x[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k <= r; k++) l[r][k] = S[(o + k) * PLD + o + r];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      double v = (r == cj) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < r; k++) v -= l[r][k] * x[k];
      x[r] = v / l[r][r];
    }
#pragma unroll
    for (int r = 0; r < 8; r++) W[(o + cj) * PLD + o + r] = x[r];
  }
  __syncthreads();
#pragma unroll 1
  for (int h = 8; h < 64; h *= 2) {
    const int tp = h >> 3, tpp = tp * tp, ntile = (64 / (2 * h)) * tpp;
    // T = L_BA W_AA
    for (int t = warp; t < ntile; t += 8) {
      const int q = t / tpp, rem = t - q * tpp, ti = rem / tp, tj = rem - ti * tp;
      const int ao = q * 2 * h, co = ao + h;
      double c0 = 0.0, c1 = 0.0;
      for (int k4 = 0; k4 < h; k4 += 4) {
        const double a = S[(ao + k4 + lc) * PLD + co + ti * 8 + lr];
        const double b = W[(ao + tj * 8 + lr) * PLD + ao + k4 + lc];
        dmma884(c0, c1, a, b);
      }
      double* tp_ = Tm + (ao + tj * 8 + 2 * lc) * PLD + co + ti * 8 + lr;
      tp_[0] = c0;
      tp_[PLD] = c1;
    }
    __syncthreads();
    // W_BA = -W_CC T
    for (int t = warp; t < ntile; t += 8) {
      const int q = t / tpp, rem = t - q * tpp, ti = rem / tp, tj = rem - ti * tp;
      const int ao = q * 2 * h, co = ao + h;
      double c0 = 0.0, c1 = 0.0;
      for (int k4 = 0; k4 < h; k4 += 4) {
        const double a = -W[(co + k4 + lc) * PLD + co + ti * 8 + lr];
        const double b = Tm[(ao + tj * 8 + lr) * PLD + co + k4 + lc];
        dmma884(c0, c1, a, b);
      }
      double* wp = W + (ao + tj * 8 + 2 * lc) * PLD + co + ti * 8 + lr;
      wp[0] = c0;
      wp[PLD] = c1;
    }
    __syncthreads();
  }
  {
    double* __restrict__ Wg = task_b_ptr(T, ar);
    const int ldb = T.ldb;
    const int nw = (T.flags & TF_B_DINV) ? 64 : n;  // scratch slots are full 64x64 (identity padded)
    for (int e = tid; e < 64 * 64; e += 256) {
      const int i = e & 63, k = e >> 6;
      if (i < nw && k < nw) Wg[i + (int64_t)k * ldb] = W[k * PLD + i];
    }
  }
}

// ------------------------------------------------------------------------------------- apply inverse ----
// X (M x N, N <= 64) <- +-X W' (TRANS: "X L^{-T}") or +-X W (!TRANS: "X L^{-1}") with W = L^{-1} from k_potrf64.
// The launch sits on the dependent chain of every blocked factorisation (POTRF -> apply inverse -> panel update), so it
// is cut for latency, not for throughput: one CTA = TRSM_ROWS = 32 rows, one warp = 8 rows = one DMMA row tile on its
// own SM sub-core (128 dependent-chain DMMAs per warp instead of 512 with 128-row CTAs), which also spreads a
// 4096-row panel over 128 SMs instead of 32.  The A fragments are loaded straight from global memory (issued before W
// is staged, so their latency overlaps), the result is written in place.
constexpr int AI_NT = 128;
template <bool TRANS>
__global__ void __launch_bounds__(AI_NT) k_apply_inv(const Task* __restrict__ tasks, int ntasks, Arenas ar) {
  __shared__ double Ws[64 * PLD];
  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int M = T.M, N = T.N;
  const double* __restrict__ Wg = task_b_ptr(T, ar);
  double* __restrict__ X = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int ldw = T.ldb, ldx = T.ldc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int row0 = (blockIdx.x - T.tile0) * TRSM_ROWS + warp * 8;
  const int r = row0 + lr;
  double a[16];
#pragma unroll
  for (int kk = 0; kk < 16; kk++) {
    const int c = 4 * kk + lc;
    a[kk] = (r < M && c < N) ? X[r + (int64_t)c * ldx] : 0.0;
  }
  const int nw = (T.flags & TF_B_DINV) ? 64 : N;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int e = tid + (half * 16 + u) * AI_NT, i = e & 63, k = e >> 6;
      // W is lower triangular: the strict upper part is never fetched
      v[u] = (i < nw && k < nw && i >= k) ? Wg[i + (int64_t)k * ldw] : ((i == k) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int e = tid + (half * 16 + u) * AI_NT, i = e & 63, k = e >> 6;
      Ws[k * PLD + i] = v[u];
    }
  }
  __syncthreads();
  if (row0 >= M) return;
  double acc[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; nt++) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
  for (int kk = 0; kk < 16; kk++)
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      // W is lower triangular: (X W')[:, n] needs k <= n, (X W)[:, n] needs k >= n
      if (TRANS ? (kk > 2 * nt + 1) : (kk < 2 * nt)) continue;
      const double b = TRANS ? Ws[(4 * kk + lc) * PLD + nt * 8 + lr] : Ws[(nt * 8 + lr) * PLD + 4 * kk + lc];
      dmma884(acc[nt][0], acc[nt][1], a[kk], b);
    }
  const double sgn = (T.flags & TF_NEG) ? -1.0 : 1.0;
  if (r < M) {
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int c = nt * 8 + 2 * lc + h;
        if (c < N) X[r + (int64_t)c * ldx] = sgn * acc[nt][h];
      }
  }
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
  // 16 columns per thread: all loads of a batch of 8 are issued before the first store (the kernel is bound by the
  // latency of the indexed read-modify-write, not by arithmetic)
#ifndef EA_BATCH
#define EA_BATCH 8
#endif
#pragma unroll
  for (int h = 0; h < 16 / EA_BATCH; h++) {
    double u[EA_BATCH], pv[EA_BATCH];
    double* pp[EA_BATCH];
    bool ok[EA_BATCH];
#pragma unroll
    for (int q = 0; q < EA_BATCH; q++) {
      const int lj = (tid >> 6) + 4 * (EA_BATCH * h + q), j = j0 + lj;
      ok[q] = j <= i && j < M;
      pp[q] = P + pr + (int64_t)rj[lj] * T.ldc;
      u[q] = ok[q] ? U[i + (int64_t)j * T.lda] : 0.0;
      pv[q] = ok[q] ? *pp[q] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < EA_BATCH; q++)
      if (ok[q]) *pp[q] = pv[q] + u[q];
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
#pragma unroll
  for (int h = 0; h < 16 / EA_BATCH; h++) {  // EA_BATCH gathered loads in flight per thread before the first store
    double v[EA_BATCH];
#pragma unroll
    for (int q = 0; q < EA_BATCH; q++) {
      const int lj = (tid >> 6) + 4 * (EA_BATCH * h + q);
      const int64_t b = rj[lj];
      v[q] = (j0 + lj < M) ? ((a >= b) ? Zp[a + b * T.lda] : Zp[b + a * T.lda]) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < EA_BATCH; q++) {
      const int j = j0 + (tid >> 6) + 4 * (EA_BATCH * h + q);
      if (j < M) Zc[i + (int64_t)j * T.ldc] = v[q];
    }
  }
}

// ------------------------------------------------------------------------ fused small-front kernels ----
// Fronts of order d <= SMALL_FRONT_MAX are processed by ONE CTA entirely in shared memory (the front is staged
// once, every operation of the multifrontal step runs on-chip, results are written once).  Both kernels work in
// 8-column panels: the 8x8 diagonal block is handled redundantly in registers by one thread per front row (no
// barrier inside a panel), everything of rank 8 or higher runs as 8x8x4 DMMA tiles on all warps.
//   factor : stage the assembled panel, extend-add the children's update matrices (fixed order), partial Cholesky
//            of the first s columns with the full trailing update, write L and the update matrix.
//   selinv : gather Z_RR from the parent's inverse front, then per panel J (right to left), with B = rows below J:
//            Y = L_BJ L_JJ^{-1},  Z_BJ = -Z_BB Y,  Z_JJ = L_JJ^{-T} L_JJ^{-1} - Y' Z_BJ;  write the inverse front.
// Shared layout: column-major with leading dimension ld = roundup(d, 8) + 4 (= 4 mod 8: conflict-free fragments).
// Task encoding: aux0 = supernode index.
__host__ __device__ __forceinline__ int sf_ld(int d) { return ((d + 7) & ~7) + 4; }

template <int NT>
__global__ void __launch_bounds__(NT) k_front_factor_small(const Task* __restrict__ tasks, Arenas ar,
                                                            const SnodeDesc* __restrict__ sd,
                                                            const int32_t* __restrict__ child_idx,
                                                            const int32_t* __restrict__ relmap,
                                                            int* __restrict__ info) {
  extern __shared__ __align__(16) double S[];
  __shared__ int32_t rel[SMALL_FRONT_MAX];
  const SnodeDesc D = sd[tasks[blockIdx.x].aux0];
  const int d = D.d, s = D.s, ldg = D.ld;
  const int dp = (d + 7) & ~7, lds = dp + 4;
  double* Pb = S + (size_t)dp * lds;  // zero-padded copy of the current panel, 8 columns x lds
  double* __restrict__ F = ar.p[0] + D.foff;
  const int tid = threadIdx.x, ti_ = tid & 63, tq = tid >> 6, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  // stage: the update-matrix part and the padding start from zero, the panel columns come from the arena.  All
  // global-memory loops of this kernel keep 8 independent loads per thread in flight (the kernel is latency bound).
  for (int e = tid; e < dp * lds; e += NT) S[e] = 0.0;
  const uint32_t mdp = div_magic((uint32_t)dp);
  __syncthreads();
  for (int base = tid; base < s * dp; base += 8 * NT) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int e = base + u * NT, c = fast_div(e, mdp), i = e - c * dp;
      v[u] = (e < s * dp && i >= c && i < d) ? F[(int64_t)c * ldg + i] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int e = base + u * NT, c = fast_div(e, mdp), i = e - c * dp;
      if (e < s * dp && i >= c && i < d) S[c * lds + i] = v[u];
    }
  }
  __syncthreads();
  for (int ci = 0; ci < D.nchild; ci++) {
    const SnodeDesc C = sd[child_idx[D.child0 + ci]];
    const int rc = C.d - C.s;
    const double* __restrict__ U = ar.p[0] + C.foff + (int64_t)C.s * C.ld + C.s;
    const int32_t* __restrict__ rl = relmap + C.rows_off + C.s;
    const bool cached = rc <= SMALL_FRONT_MAX;  // else: a child with a long boundary (its rows still map into this front)
    if (cached) {
      for (int i = tid; i < rc; i += NT) rel[i] = rl[i];
      __syncthreads();
    }
    const int32_t* __restrict__ rmap = cached ? rel : rl;
    // lower triangle of the child's update matrix, visited as (row chunk of 64) x column
    const int nrc64 = (rc + 63) >> 6;
    const int total = nrc64 * rc;  // items: (j, chunk)
    const bool fastd = total < 65536;
    const uint32_t mrc = div_magic((uint32_t)nrc64);
    for (int base = tq; base < total; base += 8 * (NT / 64)) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int it = base + u * (NT / 64), j = fastd ? fast_div(it, mrc) : it / nrc64, i = (it - j * nrc64) * 64 + ti_;
        v[u] = (it < total && i >= j && i < rc) ? U[i + (int64_t)j * C.ld] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int it = base + u * (NT / 64), j = fastd ? fast_div(it, mrc) : it / nrc64, i = (it - j * nrc64) * 64 + ti_;
        if (it < total && i >= j && i < rc) S[rmap[j] * lds + rmap[i]] += v[u];
      }
    }
    __syncthreads();
  }
  for (int j0 = 0; j0 < s; j0 += 8) {
    const int pw = min(8, s - j0);
    const int i = j0 + tid;
    const bool rowt = i < dp;
    double x[8];
    if (rowt) {
      double Dg[8][8], invd[8];
#pragma unroll
      for (int c = 0; c < 8; c++)
#pragma unroll
        for (int k = 0; k <= c; k++) Dg[c][k] = (c < pw) ? S[(j0 + k) * lds + j0 + c] : ((c == k) ? 1.0 : 0.0);
      const int bad = chol8(Dg, invd);
      if (bad < pw && tid == 0) atomicMin(info, D.col0 + j0 + bad);
#pragma unroll
      for (int c = 0; c < 8; c++) x[c] = (c < pw) ? S[(j0 + c) * lds + i] : 0.0;
      const int ii = i - j0;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        x[c] *= invd[c];
#pragma unroll
        for (int j = c + 1; j < 8; j++) x[j] -= x[c] * Dg[j][c];
      }
#pragma unroll
      for (int c = 0; c < 8; c++)
        if (c > ii) x[c] = 0.0;  // rows of the diagonal block: L's row up to the diagonal, zero above
    }
    __syncthreads();  // all row threads have read the diagonal block
    if (rowt) {
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (c < pw) S[(j0 + c) * lds + i] = x[c];
        Pb[c * lds + i] = x[c];
      }
    }
    __syncthreads();
    // trailing update S[i][k] -= sum_c P[i][c] P[k][c] for k >= j0 + pw, lower tiles
    {
      const int m0 = j0 + pw;  // first column that is updated
      const int t0 = m0 >> 3, m = (dp >> 3) - t0;
      int ti0 = 0, tj0 = 0;
      tri_advance(ti0, tj0, warp);
      for (; ti0 < m; tri_advance(ti0, tj0, NT / 32)) {
        const int ti = ti0 + t0, tj = tj0 + t0;
        double* cp = S + (tj * 8 + 2 * lc) * lds + ti * 8 + lr;
        double c0 = cp[0], c1 = cp[lds];
        const bool kvalid = (tj * 8 + lr) >= m0;  // columns of the panel itself are not updated
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
          const double a = -Pb[(4 * kk + lc) * lds + ti * 8 + lr];
          const double b = kvalid ? Pb[(4 * kk + lc) * lds + tj * 8 + lr] : 0.0;
          dmma884(c0, c1, a, b);
        }
        cp[0] = c0;
        cp[lds] = c1;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < d * dp; e += NT) {
    const int c = fast_div(e, mdp), i = e - c * dp;
    if (i >= c && i < d) F[(int64_t)c * ldg + i] = S[c * lds + i];
  }
}

template <int NT>
__global__ void __launch_bounds__(NT) k_front_selinv_small(const Task* __restrict__ tasks, Arenas ar,
                                                            const SnodeDesc* __restrict__ sd,
                                                            const int32_t* __restrict__ relmap,
                                                            const int32_t* __restrict__ sparent,
                                                            double* __restrict__ zdiag) {
  extern __shared__ __align__(16) double Z[];
  __shared__ int32_t rel[SMALL_FRONT_MAX];
  __shared__ double Wb[8][8];      // L_JJ^{-1} of the current panel
  __shared__ double part[NT / 32][64];   // per-warp partial sums of Y' Z_BJ
  const int sidx = tasks[blockIdx.x].aux0;
  const SnodeDesc D = sd[sidx];
  const int d = D.d, s = D.s, r = d - s, ldg = D.ld;
  const int dp = (d + 7) & ~7, lds = dp + 4;
  double* Lb = Z + (size_t)dp * lds;  // current panel of L (8 columns x lds, zero padded)
  double* Yb = Lb + 8 * lds;          // Y = L_BJ L_JJ^{-1}   (rows outside B are zero)
  const double* __restrict__ L = ar.p[0] + D.foff;
  double* __restrict__ Zg = ar.p[1] + D.foff;
  const int tid = threadIdx.x, ti_ = tid & 63, tq = tid >> 6, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  for (int e = tid; e < dp * lds; e += NT) Z[e] = 0.0;
  __syncthreads();
  if (r > 0) {
    const SnodeDesc P = sd[sparent[sidx]];
    const double* __restrict__ Zp = ar.p[1] + P.foff;
    const int32_t* __restrict__ rl = relmap + D.rows_off + s;
    for (int i = tid; i < r; i += NT) rel[i] = rl[i];
    __syncthreads();
    // Z_RR: lower triangle gathered from the parent's inverse front (8 independent loads in flight), mirrored
    const int nr64 = (r + 63) >> 6, total = nr64 * r;
    const uint32_t mr64 = div_magic((uint32_t)nr64);
    for (int base = tq; base < total; base += 8 * (NT / 64)) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int it = base + u * (NT / 64), j = fast_div(it, mr64), i = (it - j * nr64) * 64 + ti_;
        v[u] = (it < total && i >= j && i < r) ? Zp[(int64_t)rel[i] + (int64_t)rel[j] * P.ld] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int it = base + u * (NT / 64), j = fast_div(it, mr64), i = (it - j * nr64) * 64 + ti_;
        if (it < total && i >= j && i < r) {
          Z[(s + j) * lds + s + i] = v[u];
          Z[(s + i) * lds + s + j] = v[u];
        }
      }
    }
  }
  __syncthreads();
  // panel of L: Lb[c][i] = L[i][j0 + c], i >= j0 + c (8 * dp <= 8 * NT entries); the next panel is prefetched into
  // registers while the current one is processed
  double lnext[8];
  const uint32_t mdp = div_magic((uint32_t)dp);
  auto load_panel = [&](int jp) {
    const int pwp = min(8, s - jp);
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int e = tid + u * NT, c = fast_div(e, mdp), i = e - c * dp;
      lnext[u] = (jp >= 0 && e < 8 * dp && c < pwp && i >= jp + c && i < d) ? L[(int64_t)(jp + c) * ldg + i] : 0.0;
    }
  };
  load_panel(((s - 1) >> 3) << 3);
  for (int j0 = ((s - 1) >> 3) << 3; j0 >= 0; j0 -= 8) {
    const int pw = min(8, s - j0);
    const int m0 = j0 + pw;  // B = [m0, d)
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int e = tid + u * NT, c = fast_div(e, mdp), i = e - c * dp;
      if (e < 8 * dp) Lb[c * lds + i] = lnext[u];
    }
    __syncthreads();
    load_panel(j0 - 8);
    {
      // every row thread inverts the (identity padded) 8x8 triangle redundantly, then forms its row of Y
      const int i = m0 + tid;
      if (i < dp || tid == 0) {
        double l[8][8], w[8][8];
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
          for (int k = 0; k <= c; k++) l[c][k] = (c < pw) ? Lb[k * lds + j0 + c] : ((c == k) ? 1.0 : 0.0);
        double invd[8];
#pragma unroll
        for (int c = 0; c < 8; c++) invd[c] = 1.0 / l[c][c];
        // w = l^{-1}, column by column (forward substitution on the identity)
#pragma unroll
        for (int cj = 0; cj < 8; cj++) {
#pragma unroll
          for (int rr = 0; rr < 8; rr++) {
            if (rr < cj) {
              w[rr][cj] = 0.0;
            } else {
              double v = (rr == cj) ? 1.0 : 0.0;
#pragma unroll
              for (int k = cj; k < rr; k++) v -= l[rr][k] * w[k][cj];
              w[rr][cj] = v * invd[rr];
            }
          }
        }
        if (tid == 0) {
#pragma unroll
          for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 8; b++) Wb[a][b] = w[a][b];
        }
        if (i < dp) {
          double lrow[8], y[8];
#pragma unroll
          for (int k = 0; k < 8; k++) lrow[k] = Lb[k * lds + i];
#pragma unroll
          for (int c = 0; c < 8; c++) {
            double v = 0.0;
#pragma unroll
            for (int k = c; k < 8; k++) v += lrow[k] * w[k][c];
            y[c] = v;
          }
#pragma unroll
          for (int c = 0; c < 8; c++) Yb[c * lds + i] = y[c];
        }
      }
      // rows of the tile-aligned range below m0 do not belong to B
      for (int e = tid; e < 8 * 8; e += NT) {
        const int c = e >> 3, i = (m0 & ~7) + (e & 7);
        if (i < m0) Yb[c * lds + i] = 0.0;
      }
    }
    __syncthreads();
    // Z_BJ = -Z_BB Y (tiles of 8 rows), and the per-warp partial of Y' Z_BJ over the same rows
    {
      const int t0 = m0 >> 3, nt = (dp >> 3) - t0;
      double p0 = 0.0, p1 = 0.0;
      for (int t = warp; t < nt; t += NT / 32) {
        const int row0 = (t0 + t) * 8;
        double c0 = 0.0, c1 = 0.0;
        for (int k4 = t0 * 8; k4 < dp; k4 += 4) {
          const double a = -Z[(k4 + lc) * lds + row0 + lr];
          const double b = Yb[lr * lds + k4 + lc];
          dmma884(c0, c1, a, b);
        }
        const int gi = row0 + lr;
        if (gi >= m0) {
          // store both triangles: column j0 + c (rows in B) and row j0 + c
          if (2 * lc < pw) {
            Z[(j0 + 2 * lc) * lds + gi] = c0;
            Z[gi * lds + j0 + 2 * lc] = c0;
          }
          if (2 * lc + 1 < pw) {
            Z[(j0 + 2 * lc + 1) * lds + gi] = c1;
            Z[gi * lds + j0 + 2 * lc + 1] = c1;
          }
        }
        __syncwarp();
        // partial (8x8) += Y_tile' Z_BJ_tile  (k = the 8 rows of this tile)
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
          const int kr = row0 + 4 * kk + lc;
          const double a = Yb[lr * lds + kr];
          const double b = (kr >= m0) ? Z[(j0 + lr) * lds + kr] : 0.0;
          dmma884(p0, p1, a, b);
        }
      }
      part[warp][lr * 8 + 2 * lc] = p0;
      part[warp][lr * 8 + 2 * lc + 1] = p1;
    }
    __syncthreads();
    if (tid < 64) {
      const int a = tid >> 3, b = tid & 7;  // Z_JJ[a][b] = (W'W)[a][b] - sum_w part[w][a][b]
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) v += Wb[k][a] * Wb[k][b];
#pragma unroll
      for (int w = 0; w < NT / 32; w++) v -= part[w][a * 8 + b];
      if (a < pw && b < pw) Z[(j0 + b) * lds + j0 + a] = v;
    }
    __syncthreads();
  }
  for (int e = tid; e < d * dp; e += NT) {
    const int c = fast_div(e, mdp), i = e - c * dp;
    if (i >= c && i < d) Zg[(int64_t)c * ldg + i] = Z[c * lds + i];
  }
  for (int c = tid; c < s; c += NT) zdiag[D.col0 + c] = Z[c * lds + c];
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

// c (M x N) <- 0, one CTA per 64 x 64 tile; with TF_TRI only the tiles that meet the lower triangle.  The numeric
// factorisation clears exactly what it accumulates into (the lower triangles of the large fronts, the panels of the
// small ones) instead of the whole frontal arena: 3.5 GB instead of 8.3 GB on the 1M-node mesh.
__global__ void __launch_bounds__(256) k_zero_front(const Task* __restrict__ tasks, int ntasks, Arenas ar) {
  const int tix = find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int M = T.M, N = T.N;
  const int local = blockIdx.x - T.tile0;
  int ti, tj;
  if (T.flags & TF_TRI) {
    tri_decode(local, ti, tj);
  } else {
    const int ntm = (M + 63) / 64;
    ti = local % ntm;
    tj = local / ntm;
  }
  double* __restrict__ C = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  // fronts start on 128-byte boundaries and have even leading dimensions: row pairs are 16-byte aligned
  const int i = ti * 64 + 2 * (threadIdx.x & 31);
  if (i >= M) return;
  const bool pair = i + 1 < M;
#pragma unroll 8
  for (int lj = threadIdx.x >> 5; lj < 64; lj += 8) {
    const int j = tj * 64 + lj;
    if (j >= N) break;
    double* p = C + i + (int64_t)j * T.ldc;
    if (pair)
      *reinterpret_cast<double2*>(p) = make_double2(0.0, 0.0);
    else
      *p = 0.0;
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
static size_t potrf_smem() { return (size_t)(3 * 64 * PLD) * sizeof(double); }
template <bool TA, bool TB, class CFG>
static cudaError_t gemm_attr() {
  return cudaFuncSetAttribute(k_gemm2<TA, TB, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM);
}
template <bool TA, bool TB>
static void gemm_launch(const Launch& L, const Task* t, const Arenas& ar, cudaStream_t st) {
  if (L.cfg == GCFG_SMALL && L.grid >= GEMM_W4_MIN_GRID)
    k_gemm2<TA, TB, GemmSmall4><<<L.grid, GemmSmall4::NT, GemmSmall4::SMEM, st>>>(t, L.ntasks, ar);
  else if (L.cfg == GCFG_SMALL)
    k_gemm2<TA, TB, GemmSmall><<<L.grid, GemmSmall::NT, GemmSmall::SMEM, st>>>(t, L.ntasks, ar);
  else
    k_gemm2<TA, TB, GemmBig><<<L.grid, GemmBig::NT, GemmBig::SMEM, st>>>(t, L.ntasks, ar);
}

cudaError_t kernels_init() {
  cudaError_t e;
  if ((e = gemm_attr<false, false, GemmBig>()) != cudaSuccess) return e;
  if ((e = gemm_attr<false, true, GemmBig>()) != cudaSuccess) return e;
  if ((e = gemm_attr<true, true, GemmBig>()) != cudaSuccess) return e;
  if ((e = gemm_attr<true, false, GemmBig>()) != cudaSuccess) return e;
  if ((e = gemm_attr<false, false, GemmSmall>()) != cudaSuccess) return e;
  if ((e = gemm_attr<false, true, GemmSmall>()) != cudaSuccess) return e;
  if ((e = gemm_attr<true, true, GemmSmall>()) != cudaSuccess) return e;
  if ((e = gemm_attr<true, false, GemmSmall>()) != cudaSuccess) return e;
  if ((e = gemm_attr<false, false, GemmSmall4>()) != cudaSuccess) return e;
  if ((e = gemm_attr<false, true, GemmSmall4>()) != cudaSuccess) return e;
  if ((e = gemm_attr<true, true, GemmSmall4>()) != cudaSuccess) return e;
  if ((e = gemm_attr<true, false, GemmSmall4>()) != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_potrf64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_smem());
  if (e != cudaSuccess) return e;
  const int small_smem = small_front_smem(SMALL_FRONT_MAX);
  e = cudaFuncSetAttribute(k_front_factor_small<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_front_selinv_small<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_front_factor_small<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_front_selinv_small<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_front_factor_small<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_front_selinv_small<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem);
  if (e != cudaSuccess) return e;
  return mr_kernels_init();
}

// threads per CTA of the fused small-front kernels for a launch with `smem` bytes of dynamic shared memory
// (GMRFB_SF_THREADS="t72,t104,tbig" overrides the defaults; tuning aid)
static int small_front_threads(int smem) {
  static int cfg[3] = {0, 0, 0};
  if (cfg[0] == 0) {
    cfg[0] = 128, cfg[1] = 256, cfg[2] = 512;
    if (const char* e = std::getenv("GMRFB_SF_THREADS")) std::sscanf(e, "%d,%d,%d", &cfg[0], &cfg[1], &cfg[2]);
  }
  return smem <= small_front_smem(72) ? cfg[0] : smem <= small_front_smem(104) ? cfg[1] : cfg[2];
}

cudaError_t run_launch(const Launch& L, const Task* d_tasks, const Arenas& ar, const LaunchAux& aux,
                       cudaStream_t st) {
  if (L.grid <= 0 || L.ntasks <= 0) return cudaSuccess;
  const Task* t = d_tasks + L.task0;
  switch (L.kind) {
    case LK_GEMM_NT:
      gemm_launch<false, false>(L, t, ar, st);
      break;
    case LK_GEMM_NN:
      gemm_launch<false, true>(L, t, ar, st);
      break;
    case LK_GEMM_TN:
      gemm_launch<true, true>(L, t, ar, st);
      break;
    case LK_GEMM_TT:
      gemm_launch<true, false>(L, t, ar, st);
      break;
    case LK_SKINNY_NT:
      switch (L.cfg) {
        case 1: k_skinny_nt<1><<<L.grid, SK_ROWS, 0, st>>>(t, ar); break;
        case 2: k_skinny_nt<2><<<L.grid, SK_ROWS, 0, st>>>(t, ar); break;
        case 4: k_skinny_nt<4><<<L.grid, SK_ROWS, 0, st>>>(t, ar); break;
        default: k_skinny_nt<8><<<L.grid, SK_ROWS, 0, st>>>(t, ar); break;
      }
      break;
    case LK_SKINNY_NN:
      switch (L.cfg) {
        case 1: k_skinny_nn<1><<<L.grid, 256, 0, st>>>(t, ar); break;
        case 2: k_skinny_nn<2><<<L.grid, 256, 0, st>>>(t, ar); break;
        case 4: k_skinny_nn<4><<<L.grid, 256, 0, st>>>(t, ar); break;
        default: k_skinny_nn<8><<<L.grid, 256, 0, st>>>(t, ar); break;
      }
      break;
    case LK_POTRF:
      k_potrf64<<<L.grid, 256, potrf_smem(), st>>>(t, L.ntasks, ar, aux.d_info);
      break;
    case LK_TRSM_RLT:
      k_apply_inv<true><<<L.grid, AI_NT, 0, st>>>(t, L.ntasks, ar);
      break;
    case LK_TRSM_RLN:
      k_apply_inv<false><<<L.grid, AI_NT, 0, st>>>(t, L.ntasks, ar);
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
    case LK_ZERO_FRONT:
      k_zero_front<<<L.grid, 256, 0, st>>>(t, L.ntasks, ar);
      break;
    case LK_FRONT_FACTOR_SMALL: {
      // CTA width by size class: the kernels are latency bound (global loads of the children's update matrices), so
      // the classes that fit only one or two CTAs per SM get more threads = more loads in flight per SM
      const int nt = small_front_threads(L.smem);
      if (nt == 128)
        k_front_factor_small<128><<<L.grid, 128, L.smem, st>>>(t, ar, aux.d_snodes, aux.d_child_idx, aux.d_relmap, aux.d_info);
      else if (nt == 256)
        k_front_factor_small<256><<<L.grid, 256, L.smem, st>>>(t, ar, aux.d_snodes, aux.d_child_idx, aux.d_relmap, aux.d_info);
      else
        k_front_factor_small<512><<<L.grid, 512, L.smem, st>>>(t, ar, aux.d_snodes, aux.d_child_idx, aux.d_relmap, aux.d_info);
      break;
    }
    case LK_FRONT_SELINV_SMALL: {
      const int nt = small_front_threads(L.smem);
      if (nt == 128)
        k_front_selinv_small<128><<<L.grid, 128, L.smem, st>>>(t, ar, aux.d_snodes, aux.d_relmap, aux.d_sparent, aux.d_out);
      else if (nt == 256)
        k_front_selinv_small<256><<<L.grid, 256, L.smem, st>>>(t, ar, aux.d_snodes, aux.d_relmap, aux.d_sparent, aux.d_out);
      else
        k_front_selinv_small<512><<<L.grid, 512, L.smem, st>>>(t, ar, aux.d_snodes, aux.d_relmap, aux.d_sparent, aux.d_out);
      break;
    }
    case LK_MR_FWD_SMALL:
    case LK_MR_BWD_SMALL:
    case LK_MR_ASSEMBLE:
    case LK_MR_GATHER:
      return run_mr_launch(L, t, ar, aux, st);
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
