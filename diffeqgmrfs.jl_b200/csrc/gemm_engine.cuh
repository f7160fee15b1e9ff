// Grouped FP64 GEMM engine for sm_100a (FP64 tensor cores: DMMA m8n8k4; tcgen05 has no f64 kind).
//
//   C = beta*C + alpha*op(A)*op(B), all column-major, one CTA per BM x BN result tile of one Task.
//     TA=false: A is M x K (m contiguous)     TA=true: A is K x M (k contiguous), used as A'
//     TB=false: B is N x K (n contiguous), used as B'   TB=true: B is K x N (k contiguous)
//
// Shared-memory layout is [k][m] / [k][n] for every variant (transposed operands are transposed by the cp.async
// scatter: a warp copies 4 consecutive k of 8 columns per instruction, which is sector-exact on the global side and
// conflict-free on the shared side).  Sub-tile ownership is chosen so that the operands of two adjacent 8x8 DMMA
// sub-tiles are adjacent in shared memory: sub-tile pair p of warp wm covers rows p*16*WARPS_M + wm*16 + [0,16), and
// lane (lr, lc) supplies rows 2*lr, 2*lr+1 of that range => one conflict-free 128-bit load feeds two DMMA row
// sub-tiles (leading dimensions = 4 mod 16).  Fragments are double buffered in registers across the k4 steps.
// Partially filled tiles (ragged fronts of a batched launch, triangular results) skip inactive sub-tiles with
// warp-uniform predicates; the 16-row interleaving across warps keeps the remaining work balanced.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "kernels.hpp"
#include "tasks.hpp"

namespace gmrfb {

__device__ __forceinline__ void ge_cp_async8(double* smem_dst, const double* gsrc, bool valid) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 8 : 0;  // src-size 0 => the 8 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void ge_cp_async8_full(double* smem_dst, const double* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void ge_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void ge_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void ge_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
// Task of CTA `cta`: launches with more than one task carry a CTA -> task map right after their tasks
// (PlanBuilder::end in plan.hpp).
__device__ __forceinline__ int ge_find_task(const Task* __restrict__ tasks, int ntasks, int cta) {
  if (ntasks == 1) return 0;
  return reinterpret_cast<const int32_t*>(tasks + ntasks)[cta];
}

template <int BM_, int BN_, int WARPS_M_, int WARPS_N_, int BK_, int STAGES_, int MINB_>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, BK = BK_, STAGES = STAGES_,
                       MINB = MINB_;
  static constexpr int NT = 32 * WARPS_M * WARPS_N;
  static constexpr int SM = BM / (8 * WARPS_M), SN = BN / (8 * WARPS_N);  // 8x8 sub-tiles per warp
  static constexpr int HM = 16 * WARPS_M, HN = 16 * WARPS_N;              // slab heights (one sub-tile pair per warp)
  static constexpr int LDA = BM + 4, LDB = BN + 4;
  static constexpr int A_STAGE = BK * LDA, B_STAGE = BK * LDB;
  static constexpr int SMEM = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);
  static_assert(SM % 2 == 0 && SN % 2 == 0, "sub-tiles are owned in pairs");
  static_assert(BM % 16 == 0 && BN % 16 == 0 && BK % 4 == 0, "tile geometry");
  static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "copy loops");
  static_assert(NT % BM == 0 && NT % BN == 0 && (NT / 32) % (BK / 4) == 0, "copy geometry");
  static_assert(SM * SN <= 32, "activity mask");
};

// number of BM x BN tiles of an M x N result; lower-trapezoidal results skip tiles entirely above the diagonal
// (tile (tm, tn) is needed iff its first column tn*BN <= its last row tm*BM + BM - 1)
template <class CFG>
inline int gemm_tiles_cfg(int M, int N, bool tri) {
  const int tm = (M + CFG::BM - 1) / CFG::BM, tn = (N + CFG::BN - 1) / CFG::BN;
  if (!tri) return tm * tn;
  int t = 0;
  for (int i = 0; i < tm; i++) {
    const int w = (i * CFG::BM + CFG::BM - 1) / CFG::BN + 1;
    t += w < tn ? w : tn;
  }
  return t;
}

template <bool TA, bool TB, class CFG>
__global__ void __launch_bounds__(CFG::NT, CFG::MINB) k_gemm2(const Task* __restrict__ tasks, int ntasks, Arenas ar) {
  extern __shared__ __align__(16) double ge_smem[];
  constexpr int BM = CFG::BM, BN = CFG::BN, BK = CFG::BK, NT = CFG::NT, ST = CFG::STAGES;
  constexpr int SM = CFG::SM, SN = CFG::SN, HM = CFG::HM, HN = CFG::HN, LDA = CFG::LDA, LDB = CFG::LDB;
  double* As = ge_smem;
  double* Bs = ge_smem + ST * CFG::A_STAGE;

  const int tix = ge_find_task(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int local = blockIdx.x - T.tile0;
  const int M = T.M, N = T.N, K = T.K;
  const int ntn = (N + BN - 1) / BN;
  const bool tri = (T.flags & TF_TRI) != 0;
  int tm, tn;
  if (tri) {
    int rem = local;
    tm = 0;
    for (;;) {
      const int w = min((tm * BM + BM - 1) / BN + 1, ntn);
      if (rem < w) break;
      rem -= w;
      tm++;
    }
    tn = rem;
  } else {
    tm = local / ntn;
    tn = local - tm * ntn;
  }
  const int m0 = tm * BM, n0 = tn * BN;
  const double* __restrict__ A = ar.p[(T.flags >> TF_A_SHIFT) & 3] + T.a;
  const double* __restrict__ B = ar.p[(T.flags >> TF_B_SHIFT) & 3] + T.b;
  double* __restrict__ C = ar.p[(T.flags >> TF_C_SHIFT) & 3] + T.c;
  const int lda = T.lda, ldb = T.ldb, ldc = T.ldc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % CFG::WARPS_M, wn = warp / CFG::WARPS_M;
  const int lr = lane >> 2, lc = lane & 3;

  // activity of the warp's SM x SN sub-tiles (bit im*SN + in)
  unsigned active = 0;
#pragma unroll
  for (int im = 0; im < SM; im++)
#pragma unroll
    for (int in = 0; in < SN; in++) {
      const int r0 = m0 + (im >> 1) * HM + wm * 16 + (im & 1);  // first row; rows r0, r0+2, ..., r0+14
      const int c0 = n0 + (in >> 1) * HN + wn * 16 + (in & 1);
      bool on = (r0 < M) && (c0 < N);
      if (tri && c0 > r0 + 14) on = false;
      if (on) active |= 1u << (im * SN + in);
    }
  constexpr unsigned ALL = (SM * SN == 32) ? 0xffffffffu : ((1u << (SM * SN)) - 1u);
  // fast path: the tile lies inside the matrix and every sub-tile is active => unpredicated copies and DMMA stream
  const bool fast = (m0 + BM <= M) && (n0 + BN <= N) && (active == ALL);

  // Per-thread copy geometry.  Copy i of a thread moves element (m_b + i*MI, k_b + i*KI) of the operand tile:
  //   non-transposed operand (m contiguous in global): consecutive threads take consecutive m; (MI, KI) = (0, NT/BM)
  //   transposed operand (k contiguous in global): a warp takes 4 consecutive k of 8 columns; (MI, KI) = (2*NT/BK, 0)
  // so the global source advances by a fixed stride per copy and by a fixed increment per k-iteration, and the
  // shared destination by a compile-time stride: no per-copy index arithmetic is left in the main loop.
  constexpr int NA = (BM * BK) / NT, NB_ = (BN * BK) / NT;
  constexpr int A_MI = TA ? (NT / 32) / (BK / 4) * 8 : 0, A_KI = TA ? 0 : NT / BM;
  constexpr int B_MI = TB ? (NT / 32) / (BK / 4) * 8 : 0, B_KI = TB ? 0 : NT / BN;
  constexpr int A_DS = TA ? A_MI : A_KI * LDA, B_DS = TB ? B_MI : B_KI * LDB;  // shared stride per copy (doubles)
  const int a_m = TA ? (warp / (BK / 4)) * 8 + (lane >> 2) : tid % BM;
  const int a_k = TA ? (warp % (BK / 4)) * 4 + (lane & 3) : tid / BM;
  const int b_n = TB ? (warp / (BK / 4)) * 8 + (lane >> 2) : tid % BN;
  const int b_k = TB ? (warp % (BK / 4)) * 4 + (lane & 3) : tid / BN;
  const int64_t a_step = TA ? (int64_t)A_MI * lda : (int64_t)A_KI * lda;
  const int64_t b_step = TB ? (int64_t)B_MI * ldb : (int64_t)B_KI * ldb;
  const int64_t a_kinc = TA ? BK : (int64_t)BK * lda, b_kinc = TB ? BK : (int64_t)BK * ldb;
  const unsigned a_dst = (unsigned)__cvta_generic_to_shared(As + a_k * LDA + a_m);
  const unsigned b_dst = (unsigned)__cvta_generic_to_shared(Bs + b_k * LDB + b_n);

  // reduction range in BK steps; triangular operands (flags) shorten it per tile
  int nkt = (K + BK - 1) / BK;
  int kt0 = (TA && (T.flags & TF_KLOW)) ? min(m0 / BK, nkt) : 0;  // A' lower triangular: rows k < m0 are zero
  if (TB && (T.flags & TF_BLOW)) kt0 = max(kt0, min(n0 / BK, nkt));           // B (K x N) lower: k < n0 is zero
  if (!TB && (T.flags & TF_BUPP)) nkt = min(nkt, (n0 + BN + BK - 1) / BK);    // B (N x K) lower: k >= n0 + BN is zero
  // sources of the thread's first copy at k-iteration kt0
  const double* a_src = TA ? A + (kt0 * BK + a_k) + (int64_t)(m0 + a_m) * lda : A + (m0 + a_m) + (int64_t)(kt0 * BK + a_k) * lda;
  const double* b_src = TB ? B + (kt0 * BK + b_k) + (int64_t)(n0 + b_n) * ldb : B + (n0 + b_n) + (int64_t)(kt0 * BK + b_k) * ldb;

  // stages must be loaded in increasing k order (the source pointers advance with every call)
  auto load_stage = [&](auto pred_tag, int stage, int k0) {
    constexpr bool PRED = decltype(pred_tag)::value;
    const unsigned as = a_dst + stage * (CFG::A_STAGE * 8), bs = b_dst + stage * (CFG::B_STAGE * 8);
#pragma unroll
    for (int i = 0; i < NA; i++) {
      if (PRED) {
        const bool ok = (m0 + a_m + i * A_MI < M) && (k0 + a_k + i * A_KI < K);
        const double* src = ok ? a_src + i * a_step : A;
        const int sz = ok ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(as + i * (A_DS * 8)), "l"(src), "r"(sz));
      } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(as + i * (A_DS * 8)), "l"(a_src + i * a_step));
      }
    }
#pragma unroll
    for (int i = 0; i < NB_; i++) {
      if (PRED) {
        const bool ok = (n0 + b_n + i * B_MI < N) && (k0 + b_k + i * B_KI < K);
        const double* src = ok ? b_src + i * b_step : B;
        const int sz = ok ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(bs + i * (B_DS * 8)), "l"(src), "r"(sz));
      } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(bs + i * (B_DS * 8)), "l"(b_src + i * b_step));
      }
    }
    a_src += a_kinc;
    b_src += b_kinc;
  };

  double acc[SM][SN][2];
#pragma unroll
  for (int i = 0; i < SM; i++)
#pragma unroll
    for (int j = 0; j < SN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < ST - 1; s++) {
    if (kt0 + s < nkt) load_stage(std::true_type{}, (kt0 + s) % ST, (kt0 + s) * BK);
    ge_commit();
  }
  const int a_off = lc * LDA + wm * 16 + 2 * lr;
  const int b_off = lc * LDB + wn * 16 + 2 * lr;

  auto main_loop = [&](auto fast_tag) {
    constexpr bool FAST = decltype(fast_tag)::value;
    int stage = kt0 % ST, lstage = (kt0 + ST - 1) % ST;
    for (int kt = kt0; kt < nkt; kt++) {
#if !defined(GE_PROBE_MODE) || GE_PROBE_MODE < 2
      ge_wait<ST - 2>();
      __syncthreads();
#endif
#if !defined(GE_PROBE_MODE) || GE_PROBE_MODE < 1
      {
        const int nk = kt + ST - 1;
        if (nk < nkt) {
          if (FAST && (nk + 1) * BK <= K)
            load_stage(std::false_type{}, lstage, nk * BK);
          else
            load_stage(std::true_type{}, lstage, nk * BK);
        }
        ge_commit();
        lstage = (lstage + 1 == ST) ? 0 : lstage + 1;
      }
#endif
      const double* as = As + stage * CFG::A_STAGE + a_off;
      const double* bs = Bs + stage * CFG::B_STAGE + b_off;
      stage = (stage + 1 == ST) ? 0 : stage + 1;
      double2 fa[2][SM / 2], fb[2][SN / 2];
#pragma unroll
      for (int p = 0; p < SM / 2; p++) fa[0][p] = *reinterpret_cast<const double2*>(as + p * HM);
#pragma unroll
      for (int q = 0; q < SN / 2; q++) fb[0][q] = *reinterpret_cast<const double2*>(bs + q * HN);
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; k4++) {
        const int cur = k4 & 1, nxt = cur ^ 1;
        if (k4 + 1 < BK / 4) {
#pragma unroll
          for (int p = 0; p < SM / 2; p++)
            fa[nxt][p] = *reinterpret_cast<const double2*>(as + (k4 + 1) * 4 * LDA + p * HM);
#pragma unroll
          for (int q = 0; q < SN / 2; q++)
            fb[nxt][q] = *reinterpret_cast<const double2*>(bs + (k4 + 1) * 4 * LDB + q * HN);
        }
#pragma unroll
        for (int im = 0; im < SM; im++)
#pragma unroll
          for (int in = 0; in < SN; in++)
            if (FAST || (active & (1u << (im * SN + in))))
              ge_dmma(acc[im][in][0], acc[im][in][1], (im & 1) ? fa[cur][im >> 1].y : fa[cur][im >> 1].x,
                      (in & 1) ? fb[cur][in >> 1].y : fb[cur][in >> 1].x);
      }
    }
  };
  if (fast)
    main_loop(std::true_type{});
  else
    main_loop(std::false_type{});
  ge_wait<0>();

  const double alpha = T.alpha, beta = T.beta;
  // Epilogue: C = beta*C + alpha*acc.  acc[2p+a][2q+b][h] is element (row + a, col + 2h + b) of the thread's 2 x 4
  // block of sub-tile pair (p, q).  The read-modify-write is done one row pair p at a time with all loads issued
  // before the first store, so the global-load latency is paid once per group.
#pragma unroll
  for (int p = 0; p < SM / 2; p++) {
    const int row = m0 + p * HM + wm * 16 + 2 * lr;
    double cv[SN / 2][4][2];
    bool ok[SN / 2][4][2];
#pragma unroll
    for (int q = 0; q < SN / 2; q++)
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int a = 0; a < 2; a++) {
          const int col = n0 + q * HN + wn * 16 + 4 * lc + j, r = row + a;
          ok[q][j][a] = (active & (1u << ((2 * p + a) * SN + 2 * q + (j & 1)))) && r < M && col < N && (!tri || r >= col);
          cv[q][j][a] = 0.0;
          if (ok[q][j][a] && beta != 0.0) cv[q][j][a] = C[r + (int64_t)col * ldc];
        }
#pragma unroll
    for (int q = 0; q < SN / 2; q++)
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int a = 0; a < 2; a++) {
          const int col = n0 + q * HN + wn * 16 + 4 * lc + j, r = row + a;
          if (ok[q][j][a]) C[r + (int64_t)col * ldc] = beta * cv[q][j][a] + alpha * acc[2 * p + a][2 * q + (j & 1)][j >> 1];
        }
  }
}

}  // namespace gmrfb
