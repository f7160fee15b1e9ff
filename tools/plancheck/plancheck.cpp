// Host-only access to the static launch plans (csrc/plan.cpp) for tools/plancheck/plancheck.py: analyse a pattern with
// the library's own symbolic code, build the zero / factor / selected-inversion / wide-inverse plans and hand the
// Symbolic arrays, the Task records and the Launch records to Python, where the plans are (1) checked for hazards
// between the tasks of one launch (the tasks of a launch run concurrently on the GPU: no two may write the same entry,
// none may read what another one writes) and (2) interpreted with NumPy against dense linear algebra.  No CUDA.
//   g++ -O2 -shared -fPIC -std=c++17 -I diffeqgmrfs.jl_b200/csrc -I /usr/local/cuda/include tools/plancheck/plancheck.cpp \
//       diffeqgmrfs.jl_b200/csrc/symbolic.cpp diffeqgmrfs.jl_b200/csrc/plan.cpp -o /tmp/libplancheck.so -lpthread
#include <cstring>
#include <string>
#include <vector>

#include "plan.hpp"
#include "symbolic.hpp"

namespace gmrfb {
// the CUDA half of the analysis is not linked here: the host path of analyze_pattern is used (dev == nullptr)
std::string SymDevice::adjacency(int64_t, const int64_t*, const int64_t*, int, std::vector<int64_t>&, std::vector<int32_t>&,
                                 bool&) { return "no device"; }
std::string SymDevice::permuted_adjacency(const std::vector<int32_t>&, const std::vector<int32_t>&, std::vector<int64_t>&,
                                          std::vector<int32_t>&) { return "no device"; }
std::string SymDevice::internal_adjacency(const std::vector<int32_t>&, const std::vector<int32_t>&, std::vector<int64_t>&,
                                          std::vector<int32_t>&) { return "no device"; }
std::string SymDevice::maps(Symbolic&) { return "no device"; }
}  // namespace gmrfb

using namespace gmrfb;

struct PC {
  Symbolic S;
  Plan plans[6];  // 0 zero, 1 factor, 2 selinv, 3 wide inverse, 4 / 5 panel sweeps forward / backward
  std::vector<int32_t> wide, wide_ld, wide_idx;
  std::vector<int64_t> wide_off;
  int64_t wide_doubles = 0;
  std::string err;
};

extern "C" {

int pc_sizeof_task() { return (int)sizeof(Task); }
int pc_sizeof_launch() { return (int)sizeof(Launch); }

PC* pc_create(int64_t n, const int64_t* colptr, const int64_t* rowval, const int64_t* perm, int ordering_kind,
              int coord_dim, const double* coords) {
  PC* h = new PC();
  AnalyzeOptions o;
  o.ordering_kind = ordering_kind;
  o.base = 0;
  o.coord_dim = coords ? coord_dim : 0;
  o.coords = coords;
  h->err = analyze_pattern(n, colptr, rowval, perm, o, h->S);
  return h;
}
const char* pc_error(PC* h) { return h->err.c_str(); }
void pc_destroy(PC* h) { delete h; }

// sizes: [n, nsuper, arena, nrows (= rptr[nsuper]), nnzA, nchild_idx]
void pc_sizes(PC* h, int64_t* out) {
  const Symbolic& S = h->S;
  out[0] = S.n;
  out[1] = S.nsuper;
  out[2] = S.arena;
  out[3] = S.rptr[S.nsuper];
  out[4] = S.nnzA;
  out[5] = (int64_t)S.child_idx.size();
}
void pc_arrays(PC* h, int32_t* sptr, int64_t* rptr, int32_t* rows, int32_t* relmap, int32_t* ld, int64_t* foff,
               int32_t* sparent, int32_t* child_ptr, int32_t* child_idx, int64_t* amap, int32_t* perm) {
  const Symbolic& S = h->S;
  std::memcpy(sptr, S.sptr.data(), S.sptr.size() * 4);
  std::memcpy(rptr, S.rptr.data(), S.rptr.size() * 8);
  std::memcpy(rows, S.rows.data(), S.rows.size() * 4);
  std::memcpy(relmap, S.relmap.data(), S.relmap.size() * 4);
  std::memcpy(ld, S.ld.data(), S.ld.size() * 4);
  std::memcpy(foff, S.foff.data(), S.foff.size() * 8);
  std::memcpy(sparent, S.sparent.data(), S.sparent.size() * 4);
  std::memcpy(child_ptr, S.child_ptr.data(), S.child_ptr.size() * 4);
  std::memcpy(child_idx, S.child_idx.data(), S.child_idx.size() * 4);
  std::memcpy(amap, S.amap.data(), S.amap.size() * 8);
  std::memcpy(perm, S.perm.data(), S.perm.size() * 4);
}

// wide supernodes exactly as api.cu lays them out (columns in [wide_min, wide_max], fronts above the small-front limit);
// returns their number; out: [doubles of the inverses (the TRTRI scratch lies that much further)]
int64_t pc_set_wide(PC* h, int wide_min, int wide_max, int64_t* out) {
  const Symbolic& S = h->S;
  h->wide.clear(), h->wide_off.clear(), h->wide_ld.clear();
  int64_t off = 0;
  for (int32_t s = 0; s < S.nsuper; s++) {
    if (S.front_order(s) <= SMALL_FRONT_MAX || S.ncols(s) < wide_min || S.ncols(s) > wide_max) continue;
    const int sc = S.ncols(s), ldw = (sc + 1) & ~1;
    h->wide.push_back(s);
    h->wide_off.push_back(off);
    h->wide_ld.push_back(ldw);
    off += (int64_t)ldw * sc;
    off = (off + 15) & ~(int64_t)15;
  }
  h->wide_doubles = off;
  h->wide_idx.assign(S.nsuper, -1);
  for (size_t i = 0; i < h->wide.size(); i++) h->wide_idx[h->wide[i]] = (int32_t)i;
  out[0] = off;
  return (int64_t)h->wide.size();
}
void pc_get_wide(PC* h, int32_t* wide, int64_t* off, int32_t* ldw) {
  std::memcpy(wide, h->wide.data(), h->wide.size() * 4);
  std::memcpy(off, h->wide_off.data(), h->wide_off.size() * 8);
  std::memcpy(ldw, h->wide_ld.data(), h->wide_ld.size() * 4);
}

// build plan `which`; out: [task records (incl. the CTA -> task map slots), launches, scratch doubles, dinv doubles,
// kept inverse slots]
void pc_build(PC* h, int which, int use_wide, int64_t* out) {
  Plan& P = h->plans[which];
  P = Plan();
  if (which == 0) build_zero_plan(h->S, P);
  if (which == 1) build_factor_plan(h->S, P);
  if (which == 2) {
    if (use_wide && !h->wide.empty())
      build_selinv_plan(h->S, P, &h->wide_idx, &h->wide_off, &h->wide_ld);
    else
      build_selinv_plan(h->S, P);
  }
  if (which == 3) build_wide_inverse_plan(h->S, h->wide, h->wide_off, h->wide_ld, h->wide_doubles, P);
  out[0] = (int64_t)P.tasks.size();
  out[1] = (int64_t)P.launches.size();
  out[2] = P.scratch;
  out[3] = P.dinv;
  out[4] = P.kept_slots;
}
// panel (multi-right-hand-side) sweeps for nr right-hand sides with leading dimension ldk: plans 4 (forward) and 5
// (backward); they apply the inverse blocks the factor plan keeps, so that plan is (re)built first.
// out: [tasks fwd, launches fwd, tasks bwd, launches bwd, dinv doubles, update-panel rows (sum of r_J)]
void pc_build_mr(PC* h, int nr, int ldk, int64_t* out) {
  h->plans[1] = Plan();
  build_factor_plan(h->S, h->plans[1]);
  h->plans[4] = Plan();
  h->plans[5] = Plan();
  build_solve_mr_plans(h->S, h->plans[1].winv_slot, nr, ldk, h->plans[4], h->plans[5]);
  out[0] = (int64_t)h->plans[4].tasks.size();
  out[1] = (int64_t)h->plans[4].launches.size();
  out[2] = (int64_t)h->plans[5].tasks.size();
  out[3] = (int64_t)h->plans[5].launches.size();
  out[4] = std::max(h->plans[4].dinv, h->plans[5].dinv);
  int64_t ur = 0;
  for (int32_t s = 0; s < h->S.nsuper; s++) ur += h->S.front_order(s) - h->S.ncols(s);
  out[5] = ur;
}
void pc_get(PC* h, int which, void* tasks, void* launches, int64_t* winv_slot) {
  const Plan& P = h->plans[which];
  std::memcpy(tasks, P.tasks.data(), P.tasks.size() * sizeof(Task));
  std::memcpy(launches, P.launches.data(), P.launches.size() * sizeof(Launch));
  if (winv_slot && !P.winv_slot.empty()) std::memcpy(winv_slot, P.winv_slot.data(), P.winv_slot.size() * 8);
}

// ---- block-tridiagonal look-ahead schedule (btd.cu: btd_alloc builds these plans, btd_run_factor runs them) ----
// The POTRF and TRSM plans come from plan.cpp (plan_potrf_events / plan_trsm_rlt_events: the event numbering under test);
// the rank-512 SYRK launches restate the loop of btd_alloc.  Offsets are relative to the current block's slot
// [L_i | C_i] (the previous slot lies `slot` doubles before).  which: 0 POTRF_i, 1 TRSM_{i+1}, 2 SYRK_{i+1};
// time-sharded factor (gmrfb_btd_dist_factor: a further lane behind POTRF_i): 3 W_i = L_i^-1 by recursive doubling
// (btd_prepare_winv: arenas 0 = slot i, 1 = W_i, 2 = scratch), 4 / 5 first / later step of the spike recurrence
// (arenas 0 = slot i, 1 = spike block i with block i-1 one block before, 2 = [T | E_l | Q], 3 = W_i).
struct PCB {
  int b, ld;
  int64_t slot;
  Plan plans[6];
};
PCB* pcb_create(int b) {
  PCB* h = new PCB();
  h->b = b;
  h->ld = (b + 1) & ~1;
  h->slot = 2 * (int64_t)h->ld * b;
  const int64_t coff = (int64_t)h->ld * b;
  {
    PlanBuilder B(h->plans[0]);
    plan_potrf_events(B, h->plans[0], 0, 0, b, h->ld, 0);
  }
  {
    PlanBuilder B(h->plans[1]);
    plan_trsm_rlt_events(B, h->plans[1], 0, -h->slot, h->ld, 0, coff, b, b, h->ld);
  }
  {
    Plan& P = h->plans[2];
    PlanBuilder B(P);
    const int LA_SYRK_K = 512;
    for (int c0 = 0; c0 < b; c0 += LA_SYRK_K) {
      const int kc = std::min(LA_SYRK_K, b - c0);
      B.begin(LK_GEMM_NT);
      B.set_wait((c0 + kc - 1) / NB);
      Task t = make_task();
      t.a = coff + (int64_t)c0 * h->ld;
      t.b = t.a;
      t.c = 0;
      t.lda = t.ldb = t.ldc = h->ld;
      t.M = t.N = b;
      t.K = kc;
      t.alpha = -1.0;
      t.beta = 1.0;
      t.flags = TF_TRI;
      B.add(t, gemm_tiles(b, b, true, GCFG_BIG));
      P.flops += (double)kc * b * (b + 1);
      B.end();
    }
  }
  {
    PlanBuilder B(h->plans[3]);
    plan_trtri(B, h->plans[3], 0, 0, h->ld, 1, 0, h->ld, 2, 0, b);
  }
  {
    const int64_t bs = coff;
    const int ld = h->ld;
    auto gemm = [&](PlanBuilder& B, int kind, int aa, int64_t a, int ab, int64_t bo, int ac, int64_t c, bool tri, double alpha,
                    double beta, int32_t extra) {
      B.begin(kind);
      Task t = make_task();
      t.a = a;
      t.b = bo;
      t.c = c;
      t.lda = t.ldb = t.ldc = ld;
      t.M = t.N = t.K = b;
      t.alpha = alpha;
      t.beta = beta;
      t.flags = (aa << TF_A_SHIFT) | (ab << TF_B_SHIFT) | (ac << TF_C_SHIFT) | (tri ? TF_TRI : 0) | extra;
      B.add(t, gemm_tiles(b, b, tri, GCFG_BIG));
      B.end();
    };
    {  // S_1 = E_l' W_1';  Q = S_1 S_1'
      PlanBuilder B(h->plans[4]);
      gemm(B, LK_GEMM_TT, 2, bs, 3, 0, 1, 0, false, 1.0, 0.0, TF_BUPP);
      gemm(B, LK_GEMM_NT, 1, 0, 1, 0, 2, 2 * bs, true, 1.0, 1.0, 0);
    }
    {  // T = S_{i-1} C_i';  S_i = -T W_i';  Q += S_i S_i'
      PlanBuilder B(h->plans[5]);
      gemm(B, LK_GEMM_NT, 1, -bs, 0, bs, 2, 0, false, 1.0, 0.0, 0);
      gemm(B, LK_GEMM_NT, 2, 0, 3, 0, 1, 0, false, -1.0, 0.0, TF_BUPP);
      gemm(B, LK_GEMM_NT, 1, 0, 1, 0, 2, 2 * bs, true, 1.0, 1.0, 0);
    }
  }
  return h;
}
void pcb_destroy(PCB* h) { delete h; }
// out: [ld, slot, then per plan: tasks, launches, dinv doubles]
void pcb_sizes(PCB* h, int64_t* out) {
  out[0] = h->ld;
  out[1] = h->slot;
  for (int w = 0; w < 6; w++) {
    out[2 + 3 * w] = (int64_t)h->plans[w].tasks.size();
    out[3 + 3 * w] = (int64_t)h->plans[w].launches.size();
    out[4 + 3 * w] = h->plans[w].dinv;
  }
}
void pcb_get(PCB* h, int which, void* tasks, void* launches) {
  const Plan& P = h->plans[which];
  std::memcpy(tasks, P.tasks.data(), P.tasks.size() * sizeof(Task));
  std::memcpy(launches, P.launches.data(), P.launches.size() * sizeof(Launch));
}
}
