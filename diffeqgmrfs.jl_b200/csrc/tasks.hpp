// Task descriptors shared by the host plan builders and the CUDA kernels.
//
// Every numeric phase of the library (multifrontal factorisation, triangular solves, selected inversion,
// block-tridiagonal factorisation) is compiled on the host into a static list of *launches*; each launch is
// one kernel over a contiguous range of Task records (a "grouped"/batched launch: many dense problems of
// different sizes in one grid).  The lists depend only on the sparsity pattern, are uploaded once, and are
// replayed for every new set of values (Gauss-Newton refactorisation) without further host work.
#pragma once
#include <cstdint>

namespace gmrfb {

// Which arena an operand lives in (2 bits each in Task::flags).
enum : int32_t {
  TF_A_SHIFT = 0,   // bits 0-1: arena of operand a
  TF_B_SHIFT = 2,   // bits 2-3: arena of operand b
  TF_C_SHIFT = 4,   // bits 4-5: arena of operand c
  TF_TRI = 1 << 8,  // C is lower-trapezoidal: tiles strictly above the diagonal are skipped, diagonal tiles masked
  TF_NEG = 1 << 9,  // kind-specific: negate result
  TF_KLOW = 1 << 10,  // GEMM with transposed A (K x M) that is lower triangular (zero for k < m): skip k < m0
  TF_B_DINV = 1 << 11,    // operand b lives in the inverse-block scratch (Arenas::dinv), 64x64 slots with ld 64
  TF_NOFACTOR = 1 << 12,  // POTRF task: the block already holds L, only its inverse is produced
  TF_BLOW = 1 << 13,  // GEMM, B stored K x N (TB) and lower triangular (zero for k < n): skip k < n0
  TF_BUPP = 1 << 14,  // GEMM, B stored N x K (!TB) and lower triangular (zero for k > n): skip k >= n0 + BN
};

struct alignas(16) Task {
  int64_t a, b, c;  // element offsets into the arenas selected by flags
  int32_t M, N, K;
  int32_t lda, ldb, ldc;
  int32_t tile0;  // index of this task's first CTA inside its launch
  int32_t flags;
  int32_t aux0, aux1;  // kind-specific
  double alpha, beta;
};

enum LaunchKind : int32_t {
  LK_GEMM_NT = 0,     // C = beta C + alpha A B'        A: MxK, B: NxK
  LK_GEMM_NN = 1,     // C = beta C + alpha A B         A: MxK, B: KxN
  LK_GEMM_TN = 2,     // C = beta C + alpha A' B        A: KxM, B: KxN
  LK_POTRF = 3,       // in-place Cholesky of an n<=64 diagonal block (a, lda, M=n) + its inverse W (b, ldb); aux0 = global column
  LK_TRSM_RLT = 4,    // X <- X W' = X L^{-T}   W = L^{-1}: b (NxN, ldb), X: c (MxN, ldc), N<=64
  LK_TRSM_RLN = 5,    // X <- X W  = X L^{-1}
  LK_EXTEND_ADD = 6,  // P[rel[i], rel[j]] += U[i,j] (i>=j): U: a (MxM, lda), P: c (ldc), rel at aux0
  LK_GATHER_SYM = 7,  // Zc[i,j] = Zp[rel[i], rel[j]] symmetric read: Zp: a (lda), Zc: c (MxM, ldc), rel at aux0
  LK_SET_IDENTITY = 8,  // c (MxM, ldc) <- I
  LK_TRANSPOSE = 9,     // c (MxM, ldc) <- c' in place
  LK_SCALE = 10,        // c (MxN, ldc) <- alpha c
  LK_DIAG_OUT = 11,     // out[aux0 + i] = c[i,i], i < M   (out = int-indexed double buffer)
  LK_SYMMETRIZE = 12,   // c (MxM): copy lower triangle to the upper
  LK_FRONT_FACTOR_SMALL = 13,  // fused shared-memory factorisation of one small front per CTA (aux0 = supernode)
  LK_FRONT_SELINV_SMALL = 14,  // fused shared-memory selected inversion of one small front per CTA
  LK_GEMM_TT = 15,             // C = beta C + alpha A' B'       A: KxM, B: NxK
  LK_SKINNY_NT = 16,           // M <= 8 rows: C = beta C + alpha A B', bandwidth-bound streaming of B (skinny_kernels.cuh)
  LK_SKINNY_NN = 17,           // M <= 8 rows: C = beta C + alpha A B
  // panel (multi-right-hand-side) triangular solves, solve_mr.cu: the right-hand sides are a node-major panel
  // X (nr x n, leading dimension ldk) in arena 1, the supernodes' update panels U_J (nr x r_J) in arena 2
  LK_MR_FWD_SMALL = 32,  // forward substitution of one small supernode per CTA (aux0 = supernode)
  LK_MR_BWD_SMALL = 33,  // backward substitution of one small supernode per CTA
  LK_MR_ASSEMBLE = 34,   // X[:, C_J] += / U_J = children's update panels (aux0 = supernode; CTA = 8 right-hand sides)
  LK_MR_GATHER = 35,     // U_J <- X[:, below rows of J] (aux0 = supernode; CTA = 64 rows)
  LK_ZERO_FRONT = 36,    // c (M x N, ldc) <- 0 by 64 x 64 tiles; TF_TRI: only the tiles of the lower triangle (N = M)
};

struct Launch {
  int32_t kind;
  int32_t task0, ntasks;
  int32_t grid;  // number of CTAs
  double flops;  // useful floating-point operations of this launch (0 for data-movement kernels)
  double bytes;  // algorithmic bytes of this launch (0 where not accounted)
  int32_t smem;  // dynamic shared memory of the launch (fused small-front kernels)
  int32_t cfg;   // GEMM launches: tile configuration (GCFG_*), chosen by PlanBuilder::end()
  // look-ahead schedules (block-tridiagonal factor): before the launch its stream waits for event `wait_ev`, after it
  // the stream records event `rec_ev` (indices into the caller's event array; -1 = none)
  int16_t wait_ev = -1, rec_ev = -1;
};

// Profile slots for kernels that are not plan launches.
enum : int32_t {
  PK_FWD_LEVEL = 20,
  PK_BWD_LEVEL = 21,
  PK_SCATTER = 22,
  PK_MEMSET = 23,
  PK_PERM = 24,
  PK_FWD_ASM = 25,
  PK_BWD_RPART = 26,
  PK_FWD_SMALL = 27,
  PK_BWD_SMALL = 28,
  PK_PERM_MR = 29,
  PK_FEM = 30,
  PK_WIDE_FWD = 37,   // wide supernodes: y = W x and u -= L21 y
  PK_WIDE_BWD = 38,   // wide supernodes: x = W' t
  PK_WIDE_NORM = 39,  // |L_JJ|_inf, |W_J|_inf of the wide supernodes
  PK_MAX = 40
};

// GEMM tile configurations shared by the plan builder (tile counts) and the kernels (gemm_engine.cuh).
enum : int32_t { GCFG_BIG = 0, GCFG_SMALL = 1, GCFG_COUNT = 2 };
constexpr int GEMM_TILE_M[GCFG_COUNT] = {128, 64};
constexpr int GEMM_TILE_N[GCFG_COUNT] = {64, 64};
constexpr int NB = 64;           // panel width of the blocked POTRF/TRSM
constexpr int TRSM_ROWS = 32;    // rows per CTA in the apply-inverse (TRSM) kernels (4 warps of 8 rows)
constexpr int DINV_SLOT = 64 * 64;  // doubles per inverse-block scratch slot
constexpr int EA_TILE = 64;      // extend-add / gather tile edge
constexpr int SMALL_FRONT_MAX = 152;  // fronts up to this order are processed by one CTA in shared memory
// size classes of the fused small-front launches (one launch per class and level; smaller classes => more CTAs/SM)
constexpr int SMALL_FRONT_NCLASS = 5;
constexpr int SMALL_FRONT_CLASSES[SMALL_FRONT_NCLASS] = {48, 72, 104, 128, SMALL_FRONT_MAX};

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
// skinny products (skinny_kernels.cuh): rows of B per CTA, K slices of the nt variant, workspace doubles
constexpr int SKINNY_ROWS = 256, SKINNY_MAX_ROWS = 8, SKINNY_WS_HEAD = 512;
inline int skinny_nr(int m) { return m <= 1 ? 1 : m <= 2 ? 2 : m <= 4 ? 4 : 8; }
inline int skinny_slices(int N, int K) {
  const int chunks = cdiv(N, SKINNY_ROWS);
  int ks = 592 / chunks;  // about four CTAs per SM in flight
  ks = ks < K / 64 ? ks : K / 64;
  return ks < 1 ? 1 : ks > 32 ? 32 : ks;
}
inline int64_t skinny_workspace(int m, int N, int K) {
  return SKINNY_WS_HEAD + (int64_t)skinny_slices(N, K) * skinny_nr(m) * N;
}
// dynamic shared memory of the fused small-front kernels for a front of order d: the front (roundup(d,8) columns,
// leading dimension roundup(d,8)+4) plus two 8-column panel buffers
inline int small_front_smem(int d) {
  const int dp = (d + 7) & ~7, ld = dp + 4;
  return (dp + 16) * ld * (int)sizeof(double);
}
// number of BM x BN tiles of an M x N result under tile configuration cfg; lower-trapezoidal results skip tiles
// entirely above the diagonal (tile (tm, tn) is needed iff its first column tn*BN <= its last row tm*BM + BM - 1)
inline int gemm_tiles(int M, int N, bool tri, int cfg) {
  const int BM = GEMM_TILE_M[cfg], BN = GEMM_TILE_N[cfg];
  const int tm = cdiv(M, BM), tn = cdiv(N, BN);
  if (!tri) return tm * tn;
  int t = 0;
  for (int i = 0; i < tm; i++) {
    const int w = (i * BM + BM - 1) / BN + 1;
    t += w < tn ? w : tn;
  }
  return t;
}
// launch kinds whose kernels serve several CTAs per task and look their task up through the CTA -> task map
inline bool uses_cta_map(int kind) {
  switch (kind) {
    case LK_GEMM_NT: case LK_GEMM_NN: case LK_GEMM_TN: case LK_GEMM_TT: case LK_TRSM_RLT: case LK_TRSM_RLN:
    case LK_EXTEND_ADD: case LK_GATHER_SYM: case LK_SET_IDENTITY: case LK_TRANSPOSE: case LK_SCALE:
    case LK_DIAG_OUT: case LK_SYMMETRIZE: case LK_MR_GATHER: case LK_ZERO_FRONT:
      return true;
    default:
      return false;
  }
}
inline bool is_gemm_kind(int kind) {
  return kind == LK_GEMM_NT || kind == LK_GEMM_NN || kind == LK_GEMM_TN || kind == LK_GEMM_TT;
}
// Tile configuration of a GEMM launch.  With the 2-stage 64x64 configuration the small tile is at least as fast as the
// 128x64 one on every launch of the bench workload (batched ragged fronts: 7-12 % faster; one 4736^2 x 4096 product:
// 33.4 against 33.5 TFLOP/s; profiles/r01_gemm_batched_probe.md), so the 128x64 configuration is only chosen when
// GMRFB_GEMM_BIG_MIN asks for it (launches averaging at least that many 128x64 tiles per task); default: never.
int gemm_big_min();
bool gemm_lpt_order();  // GMRFB_GEMM_LPT=0 keeps the tasks of a GEMM launch in supernode order (A/B aid)
inline int choose_gemm_cfg(int big_tiles, int ntasks) {
  return (ntasks > 0 && big_tiles / ntasks >= gemm_big_min()) ? GCFG_BIG : GCFG_SMALL;
}

}  // namespace gmrfb
