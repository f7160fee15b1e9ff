// Plan builders: static launch lists for the multifrontal factorisation, the Takahashi selected inversion and
// the dense building blocks of the block-tridiagonal factor.  All dense work is expressed as batched
// ("grouped") launches of the tile engine in kernels.cu; fronts of one elimination-tree level advance in
// lock-step through the blocked POTRF/TRSM so one launch serves every front of the level.
#include "plan.hpp"

#include <algorithm>
#include <cstdlib>

namespace gmrfb {

int gemm_big_min() {
  static const int v = [] {
    const char* e = std::getenv("GMRFB_GEMM_BIG_MIN");  // tuning aid: 0 = always 128x64 tiles, huge = always 64x64
    return e ? std::atoi(e) : (1 << 30);
  }();
  return v;
}

bool gemm_lpt_order() {
  static const bool v = [] {
    const char* e = std::getenv("GMRFB_GEMM_LPT");
    return !(e && e[0] == '0');
  }();
  return v;
}

namespace {

inline int32_t arena_flags(int a, int b, int c) { return (a << TF_A_SHIFT) | (b << TF_B_SHIFT) | (c << TF_C_SHIFT); }

double gemm_flops(int M, int N, int K, bool tri) {
  if (!tri) return 2.0 * M * N * K;
  double n = std::min(M, N);
  return 2.0 * K * (n * (n + 1) / 2 + (double)(M - n) * n);
}

void add_gemm(PlanBuilder& B, Plan& P, int aa, int64_t a, int lda, int ab, int64_t b, int ldb, int ac, int64_t c,
              int ldc, int M, int N, int K, bool tri, double alpha, double beta) {
  if (M <= 0 || N <= 0) return;
  Task t = make_task();
  t.a = a;
  t.b = b;
  t.c = c;
  t.lda = lda;
  t.ldb = ldb;
  t.ldc = ldc;
  t.M = M;
  t.N = N;
  t.K = K;
  t.alpha = alpha;
  t.beta = beta;
  t.flags = arena_flags(aa, ab, ac) | (tri ? TF_TRI : 0);
  B.add(t, gemm_tiles(M, N, tri, GCFG_BIG));
  P.flops += gemm_flops(M, N, K, tri);
}

// X (M x n) <- X W' / X W with W = L^{-1} held in inverse-block scratch slot `slot` (written by a POTRF task)
void add_trsm(PlanBuilder& B, Plan& P, int64_t slot, int ax, int64_t x, int ldx, int M, int n) {
  if (M <= 0 || n <= 0) return;
  Task t = make_task();
  t.b = slot * DINV_SLOT;
  t.ldb = 64;
  t.c = x;
  t.ldc = ldx;
  t.M = M;
  t.N = n;
  t.flags = arena_flags(0, 0, ax) | TF_B_DINV;
  B.add(t, cdiv(M, TRSM_ROWS));
  P.flops += (double)M * n * n;
  P.dinv = std::max(P.dinv, (slot + 1) * (int64_t)DINV_SLOT);
}

// Cholesky (unless nofactor) + inverse of the nb x nb block at (arena, off, ld); inverse into scratch slot `slot`
void add_potrf(PlanBuilder& B, Plan& P, int arena, int64_t off, int ld, int nb, int col0, int64_t slot, bool nofactor) {
  Task t = make_task();
  t.a = off;
  t.lda = ld;
  t.M = nb;
  t.aux0 = col0;
  t.b = slot * DINV_SLOT;
  t.ldb = 64;
  t.flags = arena_flags(arena, 0, 0) | TF_B_DINV | (nofactor ? TF_NOFACTOR : 0);
  B.add(t, 1);
  if (!nofactor) P.flops += (double)nb * nb * nb / 3.0;
  P.dinv = std::max(P.dinv, (slot + 1) * (int64_t)DINV_SLOT);
}

}  // namespace

// ------------------------------------------------------------------------ partial factorisation batch ----
// Each problem is a d x d front (column-major, ld) whose first s columns are eliminated:
//   [F11 .; F21 F22] -> L11 = chol(F11), L21 = F21 L11^{-T}, F22 <- F22 - L21 L21'.
// Recursive blocking down to NB = 64 wide blocks (POTRF + inverse in shared memory, TRSM as a DMMA product with the
// inverse).  The F22 update is a single SYRK with K = s.
struct FactorProb {
  int arena;
  int64_t off;
  int ld, d, s, col0;
  int64_t slot0 = -1;  // >= 0: the inverses of this front's 64x64 diagonal blocks are kept in slots slot0 + j/64
                       // (the solves apply them); -1: transient scratch slots, reused by the next panel step
};

// Recursive blocking: columns [j0, j1) of every front (all earlier updates applied) are factorised as
//   factor [j0, jm);  trailing update of [jm, j1) with K = jm - j0;  factor [jm, j1)
// with jm a multiple of NB near the middle, down to single NB-wide blocks (POTRF + apply-inverse).  Most of the
// panel-update flops therefore run in GEMMs with K = width/2, width/4, ... instead of K = NB.
static void plan_factor_cols(PlanBuilder& B, Plan& P, const std::vector<FactorProb>& probs, int j0, int j1,
                             bool events = false) {
  if (j1 - j0 <= NB) {
    // with events: column block j0 / NB of L is final after the apply-inverse launch of this leaf (or after the POTRF
    // launch when no rows lie below the block)
    bool rows_below = false;
    for (auto& p : probs)
      if (j0 < p.s && p.d - (j0 + std::min(NB, p.s - j0)) > 0) rows_below = true;
    B.begin(LK_POTRF);
    {
      int64_t slot = 0;
      for (auto& p : probs) {
        if (j0 >= p.s) continue;
        int nb = std::min(NB, p.s - j0);
        add_potrf(B, P, p.arena, p.off + (int64_t)j0 * p.ld + j0, p.ld, nb, p.col0 + j0,
                  p.slot0 >= 0 ? p.slot0 + j0 / NB : slot++, false);
      }
    }
    if (events && !rows_below) B.set_record(j0 / NB);
    B.end();
    B.begin(LK_TRSM_RLT);
    if (events && rows_below) B.set_record(j0 / NB);
    {
      int64_t slot = 0;
      for (auto& p : probs) {
        if (j0 >= p.s) continue;
        int nb = std::min(NB, p.s - j0);
        int row0 = j0 + nb;
        add_trsm(B, P, p.slot0 >= 0 ? p.slot0 + j0 / NB : slot++, p.arena, p.off + (int64_t)j0 * p.ld + row0, p.ld,
                 p.d - row0, nb);
      }
    }
    B.end();
    return;
  }
  const int nblk = (j1 - j0 + NB - 1) / NB;
  const int jm = j0 + (nblk / 2) * NB;
  plan_factor_cols(B, P, probs, j0, jm, events);
  B.begin(LK_GEMM_NT);
  for (auto& p : probs) {
    if (jm >= p.s) continue;
    int c1 = std::min(j1, p.s);
    int64_t a = p.off + (int64_t)j0 * p.ld + jm;
    add_gemm(B, P, p.arena, a, p.ld, p.arena, a, p.ld, p.arena, p.off + (int64_t)jm * p.ld + jm, p.ld, p.d - jm,
             c1 - jm, jm - j0, true, -1.0, 1.0);
  }
  B.end();
  plan_factor_cols(B, P, probs, jm, j1, events);
}

static void plan_partial_factor_batch(PlanBuilder& B, Plan& P, const std::vector<FactorProb>& probs, int nbo) {
  (void)nbo;
  int max_s = 0;
  for (auto& p : probs) max_s = std::max(max_s, p.s);
  if (max_s > 0) plan_factor_cols(B, P, probs, 0, ((max_s + NB - 1) / NB) * NB);
  B.begin(LK_GEMM_NT);
  for (auto& p : probs) {
    int r = p.d - p.s;
    if (r <= 0 || p.s <= 0) continue;
    int64_t a = p.off + p.s;
    add_gemm(B, P, p.arena, a, p.ld, p.arena, a, p.ld, p.arena, p.off + (int64_t)p.s * p.ld + p.s, p.ld, r, r, p.s,
             true, -1.0, 1.0);
  }
  B.end();
}

// ------------------------------------------------------------------------------- blocked TRSM batch ----
// X (M x n) <- X L^{-T} (trans) or X L^{-1} (!trans), L n x n lower.  Left-looking over outer blocks of
// `nbo` columns (GEMM with the already-solved columns, full-width tiles), inner blocks of NB.
struct TrsmProb {
  int arenaL;
  int64_t loff;
  int ldl;
  int arenaX;
  int64_t xoff;
  int ldx;
  int M, n;
};

// Recursive over column ranges [j0, j1) (multiples of NB from 0, the same for every problem):
//   trans : solve [j0, jm);  X[:, jm:j1] -= X[:, j0:jm] L[jm:j1, j0:jm]';  solve [jm, j1)
//   !trans: solve [jm, j1);  X[:, j0:jm] -= X[:, jm:j1] L[jm:j1, j0:jm];   solve [j0, jm)
// so the updates are GEMMs with K = width/2, width/4, ...; the leaves multiply by the inverted diagonal blocks.
static void plan_trsm_cols(PlanBuilder& B, Plan& P, const std::vector<TrsmProb>& probs,
                           const std::vector<int64_t>& slot0, bool trans, int j0, int j1, bool events = false) {
  if (j1 - j0 <= NB) {
    B.begin(trans ? LK_TRSM_RLT : LK_TRSM_RLN);
    if (events) {  // wait for panel j of L; afterwards column block j of X is final
      B.set_wait(j0 / NB);
      B.set_record(j0 / NB);
    }
    for (size_t i = 0; i < probs.size(); i++) {
      const auto& p = probs[i];
      if (j0 >= p.n) continue;
      int nb = std::min(NB, p.n - j0);
      add_trsm(B, P, slot0[i] + j0 / NB, p.arenaX, p.xoff + (int64_t)j0 * p.ldx, p.ldx, p.M, nb);
    }
    B.end();
    return;
  }
  const int nblk = (j1 - j0 + NB - 1) / NB;
  const int jm = j0 + (nblk / 2) * NB;
  if (trans) {
    plan_trsm_cols(B, P, probs, slot0, trans, j0, jm, events);
    B.begin(LK_GEMM_NT);
    for (auto& p : probs) {
      if (jm >= p.n) continue;
      int c1 = std::min(j1, p.n);
      add_gemm(B, P, p.arenaX, p.xoff + (int64_t)j0 * p.ldx, p.ldx, p.arenaL, p.loff + (int64_t)j0 * p.ldl + jm, p.ldl,
               p.arenaX, p.xoff + (int64_t)jm * p.ldx, p.ldx, p.M, c1 - jm, jm - j0, false, -1.0, 1.0);
    }
    B.end();
    plan_trsm_cols(B, P, probs, slot0, trans, jm, j1, events);
  } else {
    plan_trsm_cols(B, P, probs, slot0, trans, jm, j1);
    B.begin(LK_GEMM_NN);
    for (auto& p : probs) {
      if (jm >= p.n) continue;
      int c1 = std::min(j1, p.n);
      add_gemm(B, P, p.arenaX, p.xoff + (int64_t)jm * p.ldx, p.ldx, p.arenaL, p.loff + (int64_t)j0 * p.ldl + jm, p.ldl,
               p.arenaX, p.xoff + (int64_t)j0 * p.ldx, p.ldx, p.M, jm - j0, c1 - jm, false, -1.0, 1.0);
    }
    B.end();
    plan_trsm_cols(B, P, probs, slot0, trans, j0, jm);
  }
}

static void plan_trsm_batch(PlanBuilder& B, Plan& P, const std::vector<TrsmProb>& probs, bool trans, int nbo) {
  (void)nbo;
  int max_n = 0;
  for (auto& p : probs) max_n = std::max(max_n, p.n);
  if (max_n == 0) return;
  // invert every <=64x64 diagonal block of every L once (one launch, all blocks in parallel)
  std::vector<int64_t> slot0(probs.size());
  {
    int64_t slot = 0;
    B.begin(LK_POTRF);
    for (size_t i = 0; i < probs.size(); i++) {
      const auto& p = probs[i];
      slot0[i] = slot;
      for (int jj = 0; jj < p.n; jj += NB)
        add_potrf(B, P, p.arenaL, p.loff + (int64_t)jj * p.ldl + jj, p.ldl, std::min(NB, p.n - jj), 0, slot++, true);
    }
    B.end();
  }
  plan_trsm_cols(B, P, probs, slot0, trans, 0, ((max_n + NB - 1) / NB) * NB);
}

// ---------------------------------------------------------------------- triangular inverse batch ----
// W = L^{-1} (n x n lower) by recursive doubling: invert the 64x64 diagonal blocks, then merge neighbouring
// inverted blocks  inv([A 0; B C]) = [A^{-1} 0; -C^{-1} B A^{-1}  C^{-1}]  level by level.  Every level is two
// batched GEMM launches over all merges of all problems, so the dependent chain is 1 + 2 log2(n/64) launches
// instead of ~3 n/64 for a substitution sweep.  W must be pre-set to the identity (zero strict upper triangle).
struct TrtriProb {
  int arenaL;
  int64_t loff;
  int ldl;
  int arenaW;   // W and the scratch T live in the same arena unless arenaT >= 0
  int64_t woff, toff;
  int ldw;
  int n;
  int arenaT = -1;
};

static void plan_trtri_batch(PlanBuilder& B, Plan& P, const std::vector<TrtriProb>& probs) {
  int max_n = 0;
  for (auto& p : probs) max_n = std::max(max_n, p.n);
  B.begin(LK_POTRF);  // invert-only: W_jj = L_jj^{-1} written straight into W's diagonal blocks
  for (auto& p : probs)
    for (int jj = 0; jj < p.n; jj += NB) {
      int nb = std::min(NB, p.n - jj);
      Task t = make_task();
      t.a = p.loff + (int64_t)jj * p.ldl + jj;
      t.lda = p.ldl;
      t.M = nb;
      t.b = p.woff + (int64_t)jj * p.ldw + jj;
      t.ldb = p.ldw;
      t.flags = arena_flags(p.arenaL, p.arenaW, 0) | TF_NOFACTOR;
      B.add(t, 1);
      P.flops += (double)nb * nb * nb / 3.0;
    }
  B.end();
  for (int h = NB; h < max_n; h *= 2) {
    B.begin(LK_GEMM_NN);  // T = B A^{-1}
    for (auto& p : probs)
      for (int j0 = 0; j0 + h < p.n; j0 += 2 * h) {
        int j1 = j0 + h, hc = std::min(h, p.n - j1);
        add_gemm(B, P, p.arenaL, p.loff + (int64_t)j0 * p.ldl + j1, p.ldl, p.arenaW, p.woff + (int64_t)j0 * p.ldw + j0,
                 p.ldw, p.arenaT >= 0 ? p.arenaT : p.arenaW, p.toff + (int64_t)j0 * p.ldw + j1, p.ldw, hc, h, h, false,
                 1.0, 0.0);
      }
    B.end();
    B.begin(LK_GEMM_NN);  // W_BA = -C^{-1} T
    for (auto& p : probs)
      for (int j0 = 0; j0 + h < p.n; j0 += 2 * h) {
        int j1 = j0 + h, hc = std::min(h, p.n - j1);
        add_gemm(B, P, p.arenaW, p.woff + (int64_t)j1 * p.ldw + j1, p.ldw, p.arenaT >= 0 ? p.arenaT : p.arenaW,
                 p.toff + (int64_t)j0 * p.ldw + j1, p.ldw, p.arenaW, p.woff + (int64_t)j0 * p.ldw + j1, p.ldw, hc, h, hc,
                 false, -1.0, 0.0);
      }
    B.end();
  }
}

// Full inverses W_J = L_JJ^{-1} of the wide supernodes (all levels in one batch: the dependent chain is
// 2 + 2 log2(max s / 64) launches), written to arena AR_WINV at woff[J] with leading dimension ldw[J]; the scratch of
// the merge levels lies `toff_base` doubles further in the same arena.
void build_wide_inverse_plan(const Symbolic& S, const std::vector<int32_t>& wide, const std::vector<int64_t>& woff,
                             const std::vector<int32_t>& ldw, int64_t toff_base, Plan& P) {
  P = Plan();
  if (wide.empty()) return;
  PlanBuilder B(P);
  B.begin(LK_SET_IDENTITY);
  for (size_t i = 0; i < wide.size(); i++) {
    const int sc = S.ncols(wide[i]);
    Task t = make_task();
    t.c = woff[i];
    t.ldc = ldw[i];
    t.M = sc;
    t.N = sc;
    t.flags = arena_flags(0, 0, AR_WINV);
    B.add(t, cdiv(sc, 64) * cdiv(sc, 64));
  }
  B.end();
  std::vector<TrtriProb> tp;
  for (size_t i = 0; i < wide.size(); i++) {
    const int32_t s = wide[i];
    TrtriProb p{AR_FRONT, S.foff[s], S.ld[s], AR_WINV, woff[i], toff_base + woff[i], ldw[i], S.ncols(s)};
    p.arenaT = AR_WINV;
    tp.push_back(p);
  }
  plan_trtri_batch(B, P, tp);
}

void plan_trtri(PlanBuilder& B, Plan& P, int arenaL, int64_t loff, int ldl, int arenaW, int64_t woff, int ldw, int arenaT,
                int64_t toff, int n) {
  TrtriProb p{arenaL, loff, ldl, arenaW, woff, toff, ldw, n};
  p.arenaT = arenaT;
  plan_trtri_batch(B, P, std::vector<TrtriProb>{p});
}
void plan_potrf_events(PlanBuilder& B, Plan& P, int arena, int64_t off, int n, int ld, int col0) {
  FactorProb fp{arena, off, ld, n, n, col0};
  fp.slot0 = 0;  // kept: slot j holds the inverse of diagonal block j until the next factorisation with this slot set
  std::vector<FactorProb> v{fp};
  if (n > 0) plan_factor_cols(B, P, v, 0, ((n + NB - 1) / NB) * NB, true);
}
void plan_trsm_rlt_events(PlanBuilder& B, Plan& P, int arenaL, int64_t loff, int ldl, int arenaX, int64_t xoff, int M,
                          int n, int ldx) {
  std::vector<TrsmProb> v{{arenaL, loff, ldl, arenaX, xoff, ldx, M, n}};
  std::vector<int64_t> slot0{0};
  if (n > 0) plan_trsm_cols(B, P, v, slot0, true, 0, ((n + NB - 1) / NB) * NB, true);
}
void plan_potrf(PlanBuilder& B, Plan& P, int arena, int64_t off, int n, int ld, int col0) {
  std::vector<FactorProb> v{{arena, off, ld, n, n, col0}};
  plan_partial_factor_batch(B, P, v, 0);
}
void plan_trsm_rlt(PlanBuilder& B, Plan& P, int arenaL, int64_t loff, int ldl, int arenaX, int64_t xoff, int M,
                   int n, int ldx) {
  std::vector<TrsmProb> v{{arenaL, loff, ldl, arenaX, xoff, ldx, M, n}};
  plan_trsm_batch(B, P, v, true, 0);
}
void plan_trsm_rln(PlanBuilder& B, Plan& P, int arenaL, int64_t loff, int ldl, int arenaX, int64_t xoff, int M,
                   int n, int ldx, bool negate) {
  std::vector<TrsmProb> v{{arenaL, loff, ldl, arenaX, xoff, ldx, M, n}};
  plan_trsm_batch(B, P, v, false, 0);
  if (negate) {
    B.begin(LK_SCALE);
    Task t = make_task();
    t.c = xoff;
    t.ldc = ldx;
    t.M = M;
    t.N = n;
    t.alpha = -1.0;
    t.flags = arena_flags(0, 0, arenaX);
    B.add(t, cdiv(M, 64) * cdiv(n, 64));
    B.end();
  }
}

// ------------------------------------------------------------------------------------ sparse factor ----
// Clears what the factorisation accumulates into, before the matrix values are scattered into the fronts: the panel of
// every small front (the fused kernel keeps the rest of the front in shared memory and writes it in full) and the
// lower triangle of every large front.
void build_zero_plan(const Symbolic& S, Plan& P) {
  PlanBuilder B(P);
  B.begin(LK_ZERO_FRONT);
  for (int32_t s = 0; s < S.nsuper; s++) {
    const int d = S.front_order(s), sc = S.ncols(s);
    Task t = make_task();
    t.c = S.foff[s];
    t.ldc = S.ld[s];
    t.M = d;
    t.flags = arena_flags(0, 0, AR_FRONT);
    if (d <= SMALL_FRONT_MAX) {
      t.N = sc;
      B.add(t, cdiv(d, 64) * cdiv(sc, 64));
      B.add_bytes(8.0 * d * sc);
    } else {
      t.N = d;
      t.flags |= TF_TRI;
      const int nt = cdiv(d, 64);
      B.add(t, nt * (nt + 1) / 2);
      B.add_bytes(8.0 * 64 * 64 * (nt * (nt + 1) / 2));
    }
  }
  B.end();
}

void build_factor_plan(const Symbolic& S, Plan& P) {
  PlanBuilder B(P);
  P.winv_slot.assign(S.nsuper, -1);
  P.kept_slots = 0;
  for (size_t lev = 0; lev < S.levels.size(); lev++) {
    const auto& all = S.levels[lev].snodes;
    // small fronts: one fused shared-memory kernel launch per size class
    std::vector<int32_t> sn;
    {
      std::vector<int32_t> cls[SMALL_FRONT_NCLASS];
      for (int32_t s : all) {
        int d = S.front_order(s);
        if (d > SMALL_FRONT_MAX) {
          sn.push_back(s);
          continue;
        }
        int c = 0;
        while (d > SMALL_FRONT_CLASSES[c]) c++;
        cls[c].push_back(s);
      }
      for (int c = 0; c < SMALL_FRONT_NCLASS; c++) {
        if (cls[c].empty()) continue;
        B.begin(LK_FRONT_FACTOR_SMALL);
        int maxd = 0;
        for (int32_t s : cls[c]) {
          Task t = make_task();
          t.aux0 = s;
          B.add(t, 1);
          double d = S.front_order(s), sc = S.ncols(s), r = d - sc;
          maxd = std::max(maxd, (int)d);
          P.flops += sc * sc * sc / 3.0 + sc * sc * r + sc * r * (r + 1);
          B.add_bytes(8.0 * (d * sc + r * r));
        }
        B.set_smem(small_front_smem(maxd));
        B.end();
      }
    }
    // 1. assemble children update matrices, one child rank per launch (deterministic, no atomics)
    int maxc = 0;
    for (int32_t s : sn) maxc = std::max(maxc, S.child_ptr[s + 1] - S.child_ptr[s]);
    for (int k = 0; k < maxc; k++) {
      B.begin(LK_EXTEND_ADD);
      for (int32_t s : sn) {
        if (S.child_ptr[s + 1] - S.child_ptr[s] <= k) continue;
        int32_t c = S.child_idx[S.child_ptr[s] + k];
        int sc = S.ncols(c), dc = S.front_order(c), rc = dc - sc;
        if (rc <= 0) continue;
        Task t = make_task();
        t.a = S.foff[c] + (int64_t)sc * S.ld[c] + sc;
        t.lda = S.ld[c];
        t.M = rc;
        t.c = S.foff[s];
        t.ldc = S.ld[s];
        int64_t ro = S.rptr[c] + sc;
        t.aux0 = (int32_t)(ro & 0xffffffff);
        t.aux1 = (int32_t)(ro >> 32);
        t.flags = arena_flags(AR_FRONT, 0, AR_FRONT);
        int nt = cdiv(rc, EA_TILE);
        B.add(t, nt * (nt + 1) / 2);
      }
      B.end();
    }
    // 2. partial factorisation of every front of the level
    // the inverses of the 64x64 diagonal blocks are kept (one slot each) for the triangular solves
    std::vector<FactorProb> probs;
    probs.reserve(sn.size());
    for (int32_t s : sn) {
      FactorProb fp{AR_FRONT, S.foff[s], S.ld[s], S.front_order(s), S.ncols(s), S.sptr[s]};
      fp.slot0 = P.kept_slots;
      P.winv_slot[s] = P.kept_slots;
      P.kept_slots += cdiv(S.ncols(s), NB);
      probs.push_back(fp);
    }
    plan_partial_factor_batch(B, P, probs, 0);
  }
}

// -------------------------------------------------------------------------------- selected inversion ----
// For a front with columns C (s) and below-rows R (r), given Z_RR (from the parent):
//   T    = -Z_RR L21                     (r x s)
//   H    = I - L21' T = I + L21' Z_RR L21 (s x s)
//   Z_RC = T L11^{-1},   Z_CC = L11^{-T} H L11^{-1}
void build_selinv_plan(const Symbolic& S, Plan& P, const std::vector<int32_t>* wide_idx,
                       const std::vector<int64_t>* wide_off, const std::vector<int32_t>* wide_ld) {
  PlanBuilder B(P);
  for (int lev = (int)S.levels.size() - 1; lev >= 0; lev--) {
    const auto& all = S.levels[lev].snodes;
    std::vector<int32_t> sn;
    {
      std::vector<int32_t> cls[SMALL_FRONT_NCLASS];
      for (int32_t s : all) {
        int d = S.front_order(s);
        if (d > SMALL_FRONT_MAX) {
          sn.push_back(s);
          continue;
        }
        int c = 0;
        while (d > SMALL_FRONT_CLASSES[c]) c++;
        cls[c].push_back(s);
      }
      for (int c = 0; c < SMALL_FRONT_NCLASS; c++) {
        if (cls[c].empty()) continue;
        B.begin(LK_FRONT_SELINV_SMALL);
        int maxd = 0;
        for (int32_t s : cls[c]) {
          Task t = make_task();
          t.aux0 = s;
          B.add(t, 1);
          double d = S.front_order(s), sc = S.ncols(s), r = d - sc;
          maxd = std::max(maxd, (int)d);
          P.flops += 2.0 * (sc * r * r + sc * sc * r + sc * sc * sc / 3.0);
          B.add_bytes(8.0 * (d * sc + d * d / 2 + r * r));
        }
        B.set_smem(small_front_smem(maxd));
        B.end();
      }
    }
    if (sn.empty()) continue;
    // ---- large fronts of this level: GEMM-rich formulation with the explicit inverse W = L11^{-1} ----
    //   Y' = W' L21'                 (s x r, parked in the unused upper-right block of the inverse front)
    //   Z_RC = -Z_RR Y               (r x s)
    //   Z_CC = W'W - Y' Z_RC         (s x s, lower triangle)
    // W is used for the variances only (never for the factor or the solves), where its conditioning-dependent
    // error (~cond(L11) eps) is far inside the 1e-8 tolerance and cannot propagate to posterior means.
    // (supernodes whose full inverse the factorisation already keeps for the solves - `wide` - read it from AR_WINV)
    std::vector<int64_t> woff(sn.size()), toff(sn.size());
    std::vector<int> ldw(sn.size()), warena(sn.size(), AR_WORK);
    {
      int64_t off = 0;
      for (size_t i = 0; i < sn.size(); i++) {
        int sc = S.ncols(sn[i]);
        const int32_t wi = wide_idx ? (*wide_idx)[sn[i]] : -1;
        if (wi >= 0) {
          warena[i] = AR_WINV;
          woff[i] = (*wide_off)[wi];
          ldw[i] = (*wide_ld)[wi];
          toff[i] = -1;
          continue;
        }
        ldw[i] = (sc + 1) & ~1;
        woff[i] = off;
        off += (int64_t)ldw[i] * sc;
        off = (off + 15) & ~(int64_t)15;
        toff[i] = off;
        off += (int64_t)ldw[i] * sc;
        off = (off + 15) & ~(int64_t)15;
      }
      P.scratch = std::max(P.scratch, off);
    }
    B.begin(LK_GATHER_SYM);
    for (int32_t s : sn) {
      int sc = S.ncols(s), d = S.front_order(s), r = d - sc;
      int32_t p = S.sparent[s];
      if (r <= 0 || p < 0) continue;
      Task t = make_task();
      t.a = S.foff[p];
      t.lda = S.ld[p];
      t.c = S.foff[s] + (int64_t)sc * S.ld[s] + sc;
      t.ldc = S.ld[s];
      t.M = r;
      int64_t ro = S.rptr[s] + sc;
      t.aux0 = (int32_t)(ro & 0xffffffff);
      t.aux1 = (int32_t)(ro >> 32);
      t.flags = arena_flags(AR_ZINV, 0, AR_ZINV);
      int nt = cdiv(r, EA_TILE);
      B.add(t, nt * nt);
    }
    B.end();
    // W = I L11^{-1}
    B.begin(LK_SET_IDENTITY);
    for (size_t i = 0; i < sn.size(); i++) {
      if (warena[i] != AR_WORK) continue;
      int sc = S.ncols(sn[i]);
      Task t = make_task();
      t.c = woff[i];
      t.ldc = ldw[i];
      t.M = sc;
      t.N = sc;
      t.flags = arena_flags(0, 0, AR_WORK);
      B.add(t, cdiv(sc, 64) * cdiv(sc, 64));
    }
    B.end();
    {
      std::vector<TrtriProb> tp;
      for (size_t i = 0; i < sn.size(); i++) {
        if (warena[i] != AR_WORK) continue;
        int32_t s = sn[i];
        tp.push_back({AR_FRONT, S.foff[s], S.ld[s], AR_WORK, woff[i], toff[i], ldw[i], S.ncols(s)});
      }
      if (!tp.empty()) plan_trtri_batch(B, P, tp);
    }
    B.begin(LK_GEMM_TT);  // Y' = W' L21'
    for (size_t i = 0; i < sn.size(); i++) {
      int32_t s = sn[i];
      int sc = S.ncols(s), d = S.front_order(s), r = d - sc, ld = S.ld[s];
      if (r <= 0) continue;
      Task t = make_task();
      t.a = woff[i];
      t.lda = ldw[i];
      t.b = S.foff[s] + sc;
      t.ldb = ld;
      t.c = S.foff[s] + (int64_t)sc * ld;
      t.ldc = ld;
      t.M = sc;
      t.N = r;
      t.K = sc;
      t.alpha = 1.0;
      t.beta = 0.0;
      t.flags = arena_flags(warena[i], AR_FRONT, AR_ZINV) | TF_KLOW;
      B.add(t, gemm_tiles(sc, r, false, GCFG_BIG));
      P.flops += (double)sc * sc * r;
    }
    B.end();
    B.begin(LK_GEMM_NT);  // Z_RC = -Z_RR Y
    for (int32_t s : sn) {
      int sc = S.ncols(s), d = S.front_order(s), r = d - sc, ld = S.ld[s];
      add_gemm(B, P, AR_ZINV, S.foff[s] + (int64_t)sc * ld + sc, ld, AR_ZINV, S.foff[s] + (int64_t)sc * ld, ld, AR_ZINV,
               S.foff[s] + sc, ld, r, sc, r, false, -1.0, 0.0);
    }
    B.end();
    B.begin(LK_GEMM_TN);  // Z_CC = W'W (lower)
    for (size_t i = 0; i < sn.size(); i++) {
      int32_t s = sn[i];
      int sc = S.ncols(s);
      Task t = make_task();
      t.a = woff[i];
      t.lda = ldw[i];
      t.b = woff[i];
      t.ldb = ldw[i];
      t.c = S.foff[s];
      t.ldc = S.ld[s];
      t.M = sc;
      t.N = sc;
      t.K = sc;
      t.alpha = 1.0;
      t.beta = 0.0;
      t.flags = arena_flags(warena[i], warena[i], AR_ZINV) | TF_TRI | TF_KLOW;
      B.add(t, gemm_tiles(sc, sc, true, GCFG_BIG));
      P.flops += (double)sc * sc * sc / 3.0;
    }
    B.end();
    B.begin(LK_GEMM_NN);  // Z_CC -= Y' Z_RC (lower)
    for (int32_t s : sn) {
      int sc = S.ncols(s), d = S.front_order(s), r = d - sc, ld = S.ld[s];
      if (r <= 0) continue;
      add_gemm(B, P, AR_ZINV, S.foff[s] + (int64_t)sc * ld, ld, AR_ZINV, S.foff[s] + sc, ld, AR_ZINV, S.foff[s], ld, sc,
               sc, r, true, -1.0, 1.0);
    }
    B.end();
    B.begin(LK_DIAG_OUT);
    for (int32_t s : sn) {
      int sc = S.ncols(s);
      Task t = make_task();
      t.c = S.foff[s];
      t.ldc = S.ld[s];
      t.M = sc;
      t.aux0 = S.sptr[s];
      t.aux1 = 0;
      t.flags = arena_flags(0, 0, AR_ZINV);
      B.add(t, cdiv(sc, 256));
    }
    B.end();
  }
}

// ------------------------------------------------------------------------------- panel solves ----
// Sweeps over L for a node-major panel of nr right-hand sides (solve_mr.cu): arena 0 = fronts, arena 1 = the panel X
// (nr x n, leading dimension ldk), arena 2 = the update panels U_J (nr x r_J at uoff_J * ldk), inverse-block operands =
// the inverses the factorisation kept (winv_slot).  One plan per panel width.
void build_solve_mr_plans(const Symbolic& S, const std::vector<int64_t>& winv_slot, int nr, int ldk, Plan& fwd, Plan& bwd) {
  std::vector<int64_t> uoff(S.nsuper, 0);
  {
    int64_t uo = 0;
    for (int32_t s = 0; s < S.nsuper; s++) {
      uoff[s] = uo;
      uo += S.front_order(s) - S.ncols(s);
    }
  }
  auto trap_bytes = [&](int32_t s) {
    const double d = S.front_order(s), sc = S.ncols(s);
    return 8.0 * (sc * d - sc * (sc - 1) / 2);
  };
  auto split = [&](const std::vector<int32_t>& all, std::vector<int32_t> (&cls)[SMALL_FRONT_NCLASS], std::vector<int32_t>& big) {
    for (int32_t s : all) {
      const int d = S.front_order(s);
      if (d > SMALL_FRONT_MAX) {
        big.push_back(s);
        continue;
      }
      int c = 0;
      while (d > SMALL_FRONT_CLASSES[c]) c++;
      cls[c].push_back(s);
    }
  };
  auto small_launches = [&](PlanBuilder& B, Plan& P, int kind, std::vector<int32_t> (&cls)[SMALL_FRONT_NCLASS]) {
    for (int c = 0; c < SMALL_FRONT_NCLASS; c++) {
      if (cls[c].empty()) continue;
      B.begin(kind);
      int maxd = 0;
      for (int32_t s : cls[c]) {
        Task t = make_task();
        t.aux0 = s;
        B.add(t, 1);
        maxd = std::max(maxd, S.front_order(s));
        B.add_bytes(trap_bytes(s));
        const double d = S.front_order(s), sc = S.ncols(s);
        P.flops += 2.0 * nr * (sc * d - sc * (sc - 1) / 2);
      }
      B.set_smem(maxd);  // largest front order of the launch (the kernels size their shared memory from it)
      B.end();
    }
  };
  auto trsm_probs = [&](const std::vector<int32_t>& big, std::vector<int64_t>& slot0) {
    std::vector<TrsmProb> tp;
    slot0.clear();
    for (int32_t s : big) {
      tp.push_back({AR_FRONT, S.foff[s], S.ld[s], 1, (int64_t)S.sptr[s] * ldk, ldk, nr, S.ncols(s)});
      slot0.push_back(winv_slot[s]);
    }
    return tp;
  };
  {  // ---- forward: leaves to root ----
    PlanBuilder B(fwd);
    for (size_t lev = 0; lev < S.levels.size(); lev++) {
      std::vector<int32_t> cls[SMALL_FRONT_NCLASS], big;
      split(S.levels[lev].snodes, cls, big);
      small_launches(B, fwd, LK_MR_FWD_SMALL, cls);
      if (big.empty()) continue;
      B.begin(LK_MR_ASSEMBLE);
      for (int32_t s : big) {
        Task t = make_task();
        t.aux0 = s;
        B.add(t, 1);
      }
      B.end();
      std::vector<int64_t> slot0;
      std::vector<TrsmProb> tp = trsm_probs(big, slot0);
      int max_n = 0;
      for (auto& p : tp) max_n = std::max(max_n, p.n);
      plan_trsm_cols(B, fwd, tp, slot0, true, 0, ((max_n + NB - 1) / NB) * NB);
      B.begin(LK_GEMM_NT);  // U_J -= X_J L21'
      for (int32_t s : big) {
        const int sc = S.ncols(s), r = S.front_order(s) - sc;
        if (r <= 0) continue;
        add_gemm(B, fwd, 1, (int64_t)S.sptr[s] * ldk, ldk, AR_FRONT, S.foff[s] + sc, S.ld[s], 2, uoff[s] * ldk, ldk, nr, r, sc,
                 false, -1.0, 1.0);
        B.add_bytes(8.0 * (double)r * sc);
      }
      B.end();
    }
  }
  {  // ---- backward: root to leaves ----
    PlanBuilder B(bwd);
    for (int lev = (int)S.levels.size() - 1; lev >= 0; lev--) {
      std::vector<int32_t> cls[SMALL_FRONT_NCLASS], big;
      split(S.levels[lev].snodes, cls, big);
      if (!big.empty()) {
        B.begin(LK_MR_GATHER);
        for (int32_t s : big) {
          const int r = S.front_order(s) - S.ncols(s);
          if (r <= 0) continue;
          Task t = make_task();
          t.aux0 = s;
          B.add(t, cdiv(r, 64));
        }
        B.end();
        B.begin(LK_GEMM_NN);  // X_J -= U_J L21
        for (int32_t s : big) {
          const int sc = S.ncols(s), r = S.front_order(s) - sc;
          if (r <= 0) continue;
          add_gemm(B, bwd, 2, uoff[s] * ldk, ldk, AR_FRONT, S.foff[s] + sc, S.ld[s], 1, (int64_t)S.sptr[s] * ldk, ldk, nr, sc, r,
                   false, -1.0, 1.0);
          B.add_bytes(8.0 * (double)r * sc);
        }
        B.end();
        std::vector<int64_t> slot0;
        std::vector<TrsmProb> tp = trsm_probs(big, slot0);
        int max_n = 0;
        for (auto& p : tp) max_n = std::max(max_n, p.n);
        plan_trsm_cols(B, bwd, tp, slot0, false, 0, ((max_n + NB - 1) / NB) * NB);
      }
      small_launches(B, bwd, LK_MR_BWD_SMALL, cls);
    }
  }
}

}  // namespace gmrfb
