// Host-side plan builders: compile a symbolic structure into static launch/task lists for the tile engine.
#pragma once
#include <algorithm>
#include <vector>

#include "symbolic.hpp"
#include "tasks.hpp"

namespace gmrfb {

struct Plan {
  std::vector<Task> tasks;
  std::vector<Launch> launches;
  int64_t scratch = 0;  // doubles of scratch arena (AR_WORK) the plan needs
  int64_t dinv = 0;     // doubles of inverse-block scratch (Arenas::dinv) the plan needs
  double flops = 0;  // floating-point operations of the GEMM/SYRK/TRSM/POTRF tasks (useful work, not tile padding)
  // sparse factor plan: first kept inverse-block slot of every supernode factored by the blocked path (-1: none; the
  // fused small-front kernels keep no inverses), and the number of kept slots
  std::vector<int64_t> winv_slot;
  int64_t kept_slots = 0;
};

// Incremental builder: open a launch, append tasks with their CTA counts, close it (empty launches vanish).
class PlanBuilder {
 public:
  explicit PlanBuilder(Plan& p) : P(p) {}
  void begin(int kind) {
    cur.kind = kind;
    cur.task0 = (int32_t)P.tasks.size();
    cur.ntasks = 0;
    cur.grid = 0;
    cur.bytes = 0;
    cur.smem = 0;
    cur.cfg = 0;
    cur.wait_ev = cur.rec_ev = -1;
    flops0 = P.flops;
  }
  void add(Task t, int ctas) {
    if (ctas <= 0) return;
    t.tile0 = cur.grid;
    P.tasks.push_back(t);
    cur.ntasks++;
    cur.grid += ctas;
  }
  void set_smem(int bytes) { cur.smem = bytes; }
  void set_wait(int ev) { cur.wait_ev = (int16_t)ev; }
  void set_record(int ev) { cur.rec_ev = (int16_t)ev; }
  int open_tasks() const { return cur.ntasks; }
  void set_cfg(int cfg) { cur.cfg = cfg; }
  void add_bytes(double b) { cur.bytes += b; }
  void end() {
    cur.flops = P.flops - flops0;
    if (is_gemm_kind(cur.kind) && cur.ntasks > 0) {
      // longest tiles first (K descending): the CTAs of a launch are dispatched in index order, so the last, partially
      // filled wave is made of the shortest tiles (the tasks of one launch are independent: any order is valid)
      if (cur.ntasks > 1 && gemm_lpt_order())
        std::stable_sort(P.tasks.begin() + cur.task0, P.tasks.begin() + cur.task0 + cur.ntasks,
                         [](const Task& a, const Task& b) { return a.K > b.K; });
      // GEMM tasks were added with their GCFG_BIG tile counts; pick the launch's tile configuration and re-tile
      cur.cfg = choose_gemm_cfg(cur.grid, cur.ntasks);
      cur.grid = 0;
      for (int32_t i = cur.task0; i < cur.task0 + cur.ntasks; i++) {
        Task& t = P.tasks[i];
        t.tile0 = cur.grid;
        cur.grid += gemm_tiles(t.M, t.N, (t.flags & TF_TRI) != 0, cur.cfg);
      }
    }
    if (cur.ntasks > 1 && uses_cta_map(cur.kind)) {
      // CTA -> task map of the launch, stored in Task-sized slots right after its tasks (kernels read
      // reinterpret_cast<const int32_t*>(tasks + ntasks)[blockIdx.x] instead of searching the tile0 prefix sums)
      const size_t nslots = ((size_t)cur.grid * sizeof(int32_t) + sizeof(Task) - 1) / sizeof(Task);
      const size_t first = P.tasks.size();
      P.tasks.resize(first + nslots, Task{});
      int32_t* map = reinterpret_cast<int32_t*>(P.tasks.data() + first);
      for (int32_t i = 0; i < cur.ntasks; i++) {
        const int32_t lo = P.tasks[cur.task0 + i].tile0;
        const int32_t hi = (i + 1 < cur.ntasks) ? P.tasks[cur.task0 + i + 1].tile0 : cur.grid;
        for (int32_t c = lo; c < hi; c++) map[c] = i;
      }
    }
    if (cur.ntasks > 0) P.launches.push_back(cur);
  }

 private:
  Plan& P;
  Launch cur{};
  double flops0 = 0;
};

inline Task make_task() {
  Task t{};
  t.alpha = 1.0;
  t.beta = 0.0;
  return t;
}

// Arena indices used by the sparse plans.
enum { AR_FRONT = 0, AR_ZINV = 1, AR_WORK = 2, AR_WINV = 3 };

// One launch that clears the parts of the frontal arena the factorisation accumulates into (runs before the scatter).
void build_zero_plan(const Symbolic& S, Plan& P);
// Multifrontal numeric factorisation of every front, level by level (arena 0 = frontal arena).
void build_factor_plan(const Symbolic& S, Plan& P);
// Takahashi selected inversion, top-down (arena 0 = factor fronts, arena 1 = inverse fronts);
// DIAG_OUT tasks write diag(Z) by internal column index.
// wide_idx (per supernode: index into wide_off / wide_ld or -1), when given: those supernodes' inverses W_J are read
// from arena AR_WINV (kept by the factorisation) instead of being recomputed
void build_selinv_plan(const Symbolic& S, Plan& P, const std::vector<int32_t>* wide_idx = nullptr,
                       const std::vector<int64_t>* wide_off = nullptr, const std::vector<int32_t>* wide_ld = nullptr);
void build_wide_inverse_plan(const Symbolic& S, const std::vector<int32_t>& wide, const std::vector<int64_t>& woff,
                             const std::vector<int32_t>& ldw, int64_t toff_base, Plan& P);

// Panel (multi-right-hand-side) forward / backward sweeps for nr right-hand sides held node-major with leading dimension
// ldk (solve_mr.cu); winv_slot = Plan::winv_slot of the factor plan (kept inverses of the 64 x 64 diagonal blocks).
void build_solve_mr_plans(const Symbolic& S, const std::vector<int64_t>& winv_slot, int nr, int ldk, Plan& fwd, Plan& bwd);

// Dense building blocks shared with the block-tridiagonal path (offsets relative to a moving base).
//  blocked in-place Cholesky of the n x n matrix at `off` (ld), reporting failures at column col0 + j
void plan_potrf(PlanBuilder& B, Plan& P, int arena, int64_t off, int n, int ld, int col0);
//  W (n x n at woff, ldw; its strict upper triangle must already be zero) <- L^{-1} by recursive doubling; T is an
//  n x n scratch (ldw) in arena arenaT
void plan_trtri(PlanBuilder& B, Plan& P, int arenaL, int64_t loff, int ldl, int arenaW, int64_t woff, int ldw, int arenaT,
                int64_t toff, int n);
//  the same with per-panel events for look-ahead (column block j = columns [64 j, 64 j + 64)): the inverses of the
//  diagonal blocks are kept in inverse-block slots 0, 1, ... (slot j = block j); every leaf of plan_potrf_events records
//  event j once column block j of L is final, every leaf of plan_trsm_rlt_events waits for event j before it reads
//  block row j of L and slot j (it inverts nothing itself) and records event j (of a second event array) once column
//  block j of X is final
void plan_potrf_events(PlanBuilder& B, Plan& P, int arena, int64_t off, int n, int ld, int col0);
void plan_trsm_rlt_events(PlanBuilder& B, Plan& P, int arenaL, int64_t loff, int ldl, int arenaX, int64_t xoff, int M,
                          int n, int ldx);
//  X (M x n at xoff, ldx) <- X L^{-T}, L n x n lower at loff (ldl); left-looking blocked
void plan_trsm_rlt(PlanBuilder& B, Plan& P, int arenaL, int64_t loff, int ldl, int arenaX, int64_t xoff, int M,
                   int n, int ldx);
//  X <- sign * X L^{-1}
void plan_trsm_rln(PlanBuilder& B, Plan& P, int arenaL, int64_t loff, int ldl, int arenaX, int64_t xoff, int M,
                   int n, int ldx, bool negate);

}  // namespace gmrfb
