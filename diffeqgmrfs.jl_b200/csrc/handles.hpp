// Opaque handle definitions behind include/gmrfb.h.
#pragma once
#include <map>
#include <memory>
#include <vector>

#include "common.hpp"

struct gmrfb_sym {
  gmrfb_ctx* ctx = nullptr;
  gmrfb::Symbolic S;
  bool dev_ready = false;
  gmrfb::DevBuf<int64_t> d_amap;
  gmrfb::DevBuf<int32_t> d_relmap, d_rows, d_perm, d_post, d_child_idx, d_level_lists, d_sparent;
  gmrfb::DevBuf<gmrfb::SnodeDesc> d_snodes;
  std::vector<int32_t> level_off, level_maxd;
  // solves: per level, the supernodes solved by the fused one-warp kernels (fronts of order <= SOLVE_SMALL_MAX) and the
  // others (assemble / block-step / R-part launches); offsets into d_solve_lists
  gmrfb::DevBuf<int32_t> d_solve_lists;
  std::vector<int32_t> small_off, small_cnt, big_off, big_cnt;
  std::vector<double> small_bytes;
  std::vector<double> level_bytes, level_vec_bytes, level_flops;  // algorithmic work of one solve sweep per level
  int64_t uvec_rows = 0;
  gmrfb::DevPlan factor_plan, selinv_plan, zero_plan;
  // solve schedule: per level, one launch per 64-column block step
  struct SolveLevel {
    std::vector<gmrfb::Launch> fwd_steps, bwd_steps;  // block steps of every large supernode of the level
    gmrfb::Launch rpart;
    // the same with the wide supernodes taken out of the block steps and solved through their full inverses
    std::vector<gmrfb::Launch> fwd_steps_nw, bwd_steps_nw;
    gmrfb::Launch wide_trmv, wide_below, wide_bwd;
  };
  std::vector<SolveLevel> solve_levels;
  gmrfb::DevBuf<gmrfb::Task> d_solve_tasks;
  int64_t partial_doubles = 0;
  // wide supernodes (>= SOLVE_WIDE_MIN columns): the factorisation keeps W_J = L_JJ^{-1} at wide_off[i] (leading
  // dimension wide_ld[i]) of the factor's wide-inverse buffer; wide_doubles = size of the W region (the TRTRI scratch
  // of the same size follows it)
  std::vector<int32_t> wide;
  std::vector<int64_t> wide_off;
  std::vector<int32_t> wide_ld;
  int64_t wide_doubles = 0;
  gmrfb::DevPlan wide_plan;
  gmrfb::Launch wide_norms;
  bool wide_enabled = false;
  // panel (multi-right-hand-side) sweeps, one pair of plans per panel width (built on first use)
  struct MrPlans {
    gmrfb::DevPlan fwd, bwd;
    int nr = 0, ldk = 0;
  };
  std::map<int, std::unique_ptr<MrPlans>> mr_plans;
  // captured numeric phases (factorisation, sweeps, selected inversion) of the factors of this pattern, keyed by every
  // device buffer the kernels were captured with
  gmrfb::GraphCache graphs;
};

struct gmrfb_fac {
  gmrfb_ctx* ctx = nullptr;
  gmrfb_sym* sym = nullptr;
  // hash of the addresses of every device buffer of this factor (part of every graph key)
  uint64_t buffers_key() const {
    return gmrfb::graph_key({(uint64_t)(uintptr_t)arena.p, (uint64_t)(uintptr_t)zarena.p, (uint64_t)(uintptr_t)zwork.p,
                             (uint64_t)(uintptr_t)zdiag.p, (uint64_t)(uintptr_t)nzval.p, (uint64_t)(uintptr_t)xwork.p,
                             (uint64_t)(uintptr_t)ywork.p, (uint64_t)(uintptr_t)bwork.p, (uint64_t)(uintptr_t)owork.p,
                             (uint64_t)(uintptr_t)uvec.p, (uint64_t)(uintptr_t)partial.p, (uint64_t)(uintptr_t)dinv.p,
                             (uint64_t)(uintptr_t)dinv_sel.p, (uint64_t)(uintptr_t)mr_x.p, (uint64_t)(uintptr_t)mr_u.p,
                             (uint64_t)(uintptr_t)winv_full.p, (uint64_t)(uintptr_t)wide_norms.p});
  }
  gmrfb::DevBuf<double> arena, zarena, zwork, zdiag, nzval, xwork, ywork, bwork, owork, uvec, partial, dinv, dinv_sel;
  // panel solves: node-major panel X (MR_MAX x n), update panels U (MR_MAX x sum of r_J), column-major staging (n x MR_MAX)
  gmrfb::DevBuf<double> mr_x, mr_u, mr_io;
  // full inverses of the wide supernodes + TRTRI scratch; |L_JJ|_1, |W_J|_1 per wide supernode (device / host);
  // wide_ok: the last factorisation found every cond_inf(L_JJ) below the threshold => sweeps use the inverses
  gmrfb::DevBuf<double> winv_full, wide_norms;
  std::vector<double> wide_norms_host;
  bool wide_ok = false;
  double wide_cond = 0.0;
  gmrfb::DevBuf<double> meanbuf, rbmc_x, refine_ws;  // persistent workspaces of gmrfb_sample / gmrfb_var_rbmc
  bool factored = false, z_valid = false, logdet_valid = false;
  int32_t status = GMRFB_ERR_STATE;
  int64_t fail_column = -1;
  double logdet = 0;
};

struct gmrfb_spm {
  gmrfb_ctx* ctx = nullptr;
  int64_t m = 0, n = 0, nnz = 0;
  // host copy of the pattern (0-based) for symbolic work
  std::vector<int64_t> colptr;
  std::vector<int32_t> rowidx;
  // device CSC and the row-wise (transposed) copy used by y = A x
  gmrfb::DevBuf<int64_t> d_colptr, d_rowptr, d_tmap;
  gmrfb::DevBuf<int32_t> d_rowidx, d_colidx;
  gmrfb::DevBuf<double> d_val, d_tval;
  bool owned_by_plan = false;
};

namespace gmrfb {
// fill `M` (host pattern copy + device CSC / row-wise arrays) from a base-`base` CSC pattern; nzval may be NULL (spm.cu)
gmrfb_status spm_build(gmrfb_ctx* ctx, gmrfb_spm* M, int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval,
                       const double* nzval, int base);
}  // namespace gmrfb
