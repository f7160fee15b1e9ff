// C ABI: contexts, symbolic analysis handles, numeric factorisation, solves, samples, marginal variances.
// (sparse-matrix helpers live in spm.cu, the block-tridiagonal path in btd.cu)
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "common.hpp"
#include "handles.hpp"

namespace gmrfb {

std::string& global_error() {
  static thread_local std::string e;
  return e;
}

gmrfb_status run_plan(gmrfb_ctx* ctx, const DevPlan& P, const Arenas& ar, const LaunchAux& aux) {
  for (const Launch& L : P.host.launches) {
    ProfScope ps(ctx, L.kind, L.flops, L.bytes, L.grid, L.ntasks);
    cudaError_t e = run_launch(L, P.tasks.p, ar, aux, ctx->stream);
    if (e != cudaSuccess)
      return fail(ctx, GMRFB_ERR_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(e));
    ctx->launches++;
  }
  return GMRFB_OK;
}

}  // namespace gmrfb

using namespace gmrfb;

// ------------------------------------------------------------------------------------------ context ----
extern "C" int32_t gmrfb_version(void) { return 100; }

extern "C" gmrfb_status gmrfb_ctx_create(int32_t device, gmrfb_ctx** out) try {
  if (!out) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(nullptr, GMRFB_ERR_CUDA,
                std::string("gmrfb_ctx_create: no CUDA device available (") +
                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                    "); libgmrfb has no CPU fallback");
  }
  if (device < 0 || device >= count) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_ctx_create: bad device index");
  std::unique_ptr<gmrfb_ctx> c(new gmrfb_ctx());
  c->device = device;
  GMRFB_CU(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  GMRFB_CU(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(nullptr, GMRFB_ERR_CUDA, "gmrfb_ctx_create: device is not sm_100-class; libgmrfb is built for sm_100a only");
  c->sm_count = prop.multiProcessorCount;
  // the context's stream gets the highest priority so that a second, default-priority lane (ctx->stream2: GPU-filling
  // GEMMs of the time-sharded factor) cannot starve the small dependent kernels queued here: the CTA dispatcher serves
  // pending blocks of higher-priority streams first whenever an SM frees resources
  int prio_least = 0, prio_greatest = 0;
  GMRFB_CU(nullptr, cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  GMRFB_CU(nullptr, cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
  GMRFB_CU(nullptr, cudaMalloc((void**)&c->d_info, sizeof(int)));
  GMRFB_CU(nullptr, cudaMalloc((void**)&c->d_info_init, sizeof(int)));
  {
    const int big = INT_MAX;
    GMRFB_CU(nullptr, cudaMemcpy(c->d_info_init, &big, sizeof(int), cudaMemcpyHostToDevice));
    const char* g = getenv("GMRFB_GRAPHS");
    c->use_graphs = !(g && g[0] == '0');
    const char* po = getenv("GMRFB_POISON");
    c->poison = po && po[0] == '1';
  }
  GMRFB_CU(nullptr, cudaMalloc((void**)&c->d_scalar, 16 * sizeof(double)));
  GMRFB_CU(nullptr, kernels_init());
  GMRFB_CU(nullptr, sparse_kernels_init());
  DevPool::get().ctx_created();
  *out = c.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_pool_trim(int64_t keep_bytes, int64_t* cached_bytes_out) try {
  DevPool::get().trim(keep_bytes > 0 ? (size_t)keep_bytes : 0);
  if (cached_bytes_out) *cached_bytes_out = (int64_t)DevPool::get().cached();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_ctx_destroy(gmrfb_ctx* ctx) try {
  if (!ctx) return GMRFB_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->d_info) cudaFree(ctx->d_info);
  if (ctx->d_info_init) cudaFree(ctx->d_info_init);
  if (ctx->d_scalar) cudaFree(ctx->d_scalar);
  if (ctx->stream2) {
    cudaStreamSynchronize(ctx->stream2);
    cudaStreamDestroy(ctx->stream2);
  }
  if (ctx->ev_lane) cudaEventDestroy(ctx->ev_lane);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  DevPool::get().ctx_destroyed();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" const char* gmrfb_last_error(gmrfb_ctx* ctx) {
  return (ctx && !ctx->err.empty()) ? ctx->err.c_str() : global_error().c_str();
}

extern "C" gmrfb_status gmrfb_ctx_sync(gmrfb_ctx* ctx) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "ctx is NULL");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH
extern "C" gmrfb_status gmrfb_ctx_profile_begin(gmrfb_ctx* ctx) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "ctx is NULL");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->prof.clear();
  ctx->profiling = true;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

static const char* prof_name(int kind) {
  switch (kind) {
    case LK_GEMM_NT: return "k_gemm<NT> (DMMA)";
    case LK_GEMM_NN: return "k_gemm<NN> (DMMA)";
    case LK_GEMM_TN: return "k_gemm<TN> (DMMA)";
    case LK_GEMM_TT: return "k_gemm<TT> (DMMA)";
    case LK_SKINNY_NT: return "k_skinny_nt";
    case LK_SKINNY_NN: return "k_skinny_nn";
    case LK_POTRF: return "k_potrf64";
    case LK_TRSM_RLT: return "k_apply_inv<T>";
    case LK_TRSM_RLN: return "k_apply_inv<N>";
    case LK_EXTEND_ADD: return "k_extend_add";
    case LK_GATHER_SYM: return "k_gather_sym";
    case LK_SET_IDENTITY: return "k_tile_op<identity>";
    case LK_TRANSPOSE: return "k_transpose";
    case LK_SCALE: return "k_tile_op<scale>";
    case LK_DIAG_OUT: return "k_diag_out";
    case LK_SYMMETRIZE: return "k_tile_op<symmetrize>";
    case LK_FRONT_FACTOR_SMALL: return "k_front_factor_small";
    case LK_FRONT_SELINV_SMALL: return "k_front_selinv_small";
    case PK_FWD_LEVEL: return "k_fwd_step";
    case PK_BWD_LEVEL: return "k_bwd_step";
    case PK_FWD_ASM: return "k_fwd_assemble";
    case PK_BWD_RPART: return "k_bwd_rpart";
    case PK_WIDE_FWD: return "k_wide_gemv";
    case PK_WIDE_BWD: return "k_wide_trmv_t";
    case PK_WIDE_NORM: return "k_wide_norms";
    case PK_FWD_SMALL: return "k_fwd_small";
    case PK_BWD_SMALL: return "k_bwd_small";
    case PK_SCATTER: return "k_scatter_values";
    case PK_MEMSET: return "memset(front arena)";
    case PK_PERM: return "k_perm_gather/scatter";
    case PK_PERM_MR: return "k_mr_perm_in/out";
    case PK_FEM: return "k_fem_assemble";
    case LK_MR_FWD_SMALL: return "k_mr_fwd_small";
    case LK_MR_BWD_SMALL: return "k_mr_bwd_small";
    case LK_MR_ASSEMBLE: return "k_mr_assemble";
    case LK_MR_GATHER: return "k_mr_gather";
    case LK_ZERO_FRONT: return "k_zero_front";
  }
  return "other";
}

extern "C" gmrfb_status gmrfb_ctx_profile_end(gmrfb_ctx* ctx, gmrfb_profile_entry* entries, int32_t cap,
                                              int32_t* count) try {
  if (!ctx || !count) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_ctx_profile_end: NULL argument");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->profiling = false;
  gmrfb_profile_entry acc[PK_MAX];
  memset(acc, 0, sizeof(acc));
  // GMRFB_PROFILE_DUMP=<file>: one CSV line per launch (kind, name, grid, ms, flops, bytes) for kernel tuning
  FILE* dump = nullptr;
  if (const char* path = getenv("GMRFB_PROFILE_DUMP")) {
    dump = fopen(path, "a");
    if (dump) fprintf(dump, "seq,kind,name,grid,ntasks,ms,flops,bytes\n");
  }
  int seq = 0;
  for (auto& r : ctx->prof) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (dump)
      fprintf(dump, "%d,%d,%s,%d,%d,%.5f,%.6e,%.6e\n", seq++, r.kind, prof_name(r.kind), r.grid, r.ntasks, ms, r.flops,
              r.bytes);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    int k = (r.kind >= 0 && r.kind < PK_MAX) ? r.kind : PK_MAX - 1;
    acc[k].kind = k;
    acc[k].launches++;
    acc[k].ms += ms;
    acc[k].flops += r.flops;
    acc[k].bytes += r.bytes;
  }
  ctx->prof.clear();
  if (dump) fclose(dump);
  int32_t n = 0;
  for (int k = 0; k < PK_MAX; k++) {
    if (acc[k].launches == 0) continue;
    if (entries && n < cap) {
      entries[n] = acc[k];
      strncpy(entries[n].name, prof_name(k), sizeof(entries[n].name) - 1);
    }
    n++;
  }
  *count = n;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" uint64_t gmrfb_ctx_stream(gmrfb_ctx* ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }
extern "C" int64_t gmrfb_ctx_launch_count(gmrfb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------ symbolic ----
extern "C" gmrfb_status gmrfb_analyze(gmrfb_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval,
                                      const int64_t* perm, const gmrfb_analyze_opts* opts, gmrfb_sym** out) try {
  if (!out) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_analyze: out is NULL");
  *out = nullptr;
  AnalyzeOptions o;
  if (opts) {
    o.ordering_kind = opts->ordering_kind;
    o.storage = opts->storage;
    o.base = opts->base;
    o.coord_dim = opts->coord_dim;
    o.coords = opts->coords;
    o.nd_leaf = opts->nd_leaf;
    o.relax_small = opts->relax_small;
    o.relax_zeros = opts->relax_zeros;
  }
  std::unique_ptr<gmrfb_sym> s(new gmrfb_sym());
  s->ctx = ctx;
  // with a context the data-parallel phases of the analysis run on its device (symbolic_gpu.cu); GMRFB_SYM_HOST=1 forces
  // the host loops (the two paths are bit-identical, tests/test_symbolic.py)
  std::unique_ptr<SymDevice> dev;
  const char* force_host = getenv("GMRFB_SYM_HOST");
  if (ctx && !(force_host && force_host[0] == '1')) {
    GMRFB_CU(ctx, cudaSetDevice(ctx->device));
    dev.reset(new SymDevice((void*)ctx->stream));
  }
  std::string err = analyze_pattern(n, colptr, rowval, perm, o, s->S, dev.get());
  if (!err.empty()) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_analyze: " + err);
  *out = s.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_sym_destroy(gmrfb_sym* sym) try {
  if (!sym) return GMRFB_OK;
  if (sym->ctx) cudaSetDevice(sym->ctx->device);
  delete sym;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_sym_get_info(const gmrfb_sym* sym, gmrfb_sym_info* info) try {
  if (!sym || !info) return fail(sym ? sym->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_sym_get_info: NULL argument");
  const Symbolic& S = sym->S;
  info->n = S.n;
  info->nnz_lower_A = S.nnz_lower_A;
  info->nnz_L = S.nnzL;
  info->nnz_L_stored = S.nnzL_stored;
  info->flops = S.flops;
  info->nsuper = S.nsuper;
  info->nlevels = (int64_t)S.levels.size();
  info->max_front = S.max_front;
  info->front_bytes = S.arena * (int64_t)sizeof(double);
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_sym_get(const gmrfb_sym* sym, int64_t* perm, int64_t* parent, int64_t* colcount,
                                      int64_t* super_ptr, int64_t* ipost) try {
  if (!sym) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_sym_get: sym is NULL");
  const Symbolic& S = sym->S;
  const int b = S.base;
  for (int64_t k = 0; k < S.n; k++) {
    if (perm) perm[k] = S.perm_user[k] + b;
    if (parent) parent[k] = S.parent_user[k] < 0 ? b - 1 : S.parent_user[k] + b;
    if (colcount) colcount[k] = S.colcount_user[k];
    if (ipost) ipost[k] = S.ipost[k] + b;
  }
  if (super_ptr)
    for (int32_t s = 0; s <= S.nsuper; s++) super_ptr[s] = S.sptr[s] + b;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_sym_get_super_rows(const gmrfb_sym* sym, int64_t s, int64_t* rows, int64_t cap,
                                                 int64_t* nrows) try {
  if (!sym || s < 0 || s >= sym->S.nsuper)
    return fail(sym ? sym->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_sym_get_super_rows: bad argument");
  const Symbolic& S = sym->S;
  int64_t cnt = S.rptr[s + 1] - S.rptr[s];
  if (nrows) *nrows = cnt;
  if (rows)
    for (int64_t k = 0; k < std::min(cnt, cap); k++) rows[k] = S.rows[S.rptr[s] + k] + S.base;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_sym_get_maps(const gmrfb_sym* sym, int64_t* amap, int64_t* relmap, int64_t* nnz_out,
                                           int64_t* total_rows_out) try {
  if (!sym) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_sym_get_maps: sym is NULL");
  const Symbolic& S = sym->S;
  if (nnz_out) *nnz_out = (int64_t)S.amap.size();
  if (total_rows_out) *total_rows_out = (int64_t)S.relmap.size();
  if (amap) std::copy(S.amap.begin(), S.amap.end(), amap);
  if (relmap)
    for (size_t k = 0; k < S.relmap.size(); k++) relmap[k] = S.relmap[k];
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// Upload everything the numeric phases need (once per symbolic handle).
static gmrfb_status sym_ensure_device(gmrfb_sym* sym) {
  if (sym->dev_ready) return GMRFB_OK;
  gmrfb_ctx* ctx = sym->ctx;
  if (!ctx) return fail(nullptr, GMRFB_ERR_STATE, "symbolic handle was created without a context (host-only analysis)");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const Symbolic& S = sym->S;
  cudaStream_t st = ctx->stream;
  GMRFB_CU(ctx, sym->d_amap.upload(S.amap, st));
  GMRFB_CU(ctx, sym->d_relmap.upload(S.relmap, st));
  GMRFB_CU(ctx, sym->d_rows.upload(S.rows, st));
  GMRFB_CU(ctx, sym->d_perm.upload(S.perm, st));
  GMRFB_CU(ctx, sym->d_post.upload(S.post, st));
  GMRFB_CU(ctx, sym->d_child_idx.upload(S.child_idx, st));
  GMRFB_CU(ctx, sym->d_sparent.upload(S.sparent, st));
  std::vector<SnodeDesc> sd(S.nsuper);
  int64_t uo = 0;
  for (int32_t s = 0; s < S.nsuper; s++) {
    SnodeDesc& D = sd[s];
    D.foff = S.foff[s];
    D.rows_off = S.rptr[s];
    D.uoff = uo;
    D.ld = S.ld[s];
    D.d = S.front_order(s);
    D.s = S.ncols(s);
    D.col0 = S.sptr[s];
    D.child0 = S.child_ptr[s];
    D.nchild = S.child_ptr[s + 1] - S.child_ptr[s];
    uo += D.d - D.s;
  }
  sym->uvec_rows = uo;
  GMRFB_CU(ctx, sym->d_snodes.upload(sd, st));
  std::vector<int32_t> lists;
  sym->level_off.assign(1, 0);
  sym->level_maxd.clear();
  sym->level_bytes.clear();
  sym->level_vec_bytes.clear();
  sym->level_flops.clear();
  for (auto& L : S.levels) {
    int md = 0;
    double lb = 0, vb = 0, lf = 0;
    for (int32_t s : L.snodes) {
      lists.push_back(s);
      md = std::max(md, S.front_order(s));
      double d = S.front_order(s), sc = S.ncols(s);
      double trap = sc * d - sc * (sc - 1) / 2;
      lb += 8.0 * trap;   // factor values read once per sweep
      vb += 16.0 * d;     // solution/update vector entries read + written per right-hand side
      lf += 2.0 * trap;   // one multiply-add per stored factor entry per right-hand side
    }
    sym->level_bytes.push_back(lb);
    sym->level_vec_bytes.push_back(vb);
    sym->level_flops.push_back(lf);
    sym->level_off.push_back((int32_t)lists.size());
    sym->level_maxd.push_back(md);
  }
  GMRFB_CU(ctx, sym->d_level_lists.upload(lists, st));
  // the factor plan comes first: the solve tasks refer to the inverse-block slots it keeps
  build_factor_plan(S, sym->factor_plan.host);
  // ---- solve schedule ----
  auto solve_small = [&](int32_t s) { return S.front_order(s) <= SOLVE_SMALL_MAX; };
  {
    std::vector<int32_t> sl;
    sym->small_off.clear(), sym->small_cnt.clear(), sym->big_off.clear(), sym->big_cnt.clear(), sym->small_bytes.clear();
    for (auto& L : S.levels) {
      double sb = 0;
      sym->small_off.push_back((int32_t)sl.size());
      for (int32_t s : L.snodes)
        if (solve_small(s)) {
          sl.push_back(s);
          const double d = S.front_order(s), sc = S.ncols(s);
          sb += 8.0 * (sc * d - sc * (sc - 1) / 2);
        }
      sym->small_cnt.push_back((int32_t)sl.size() - sym->small_off.back());
      sym->small_bytes.push_back(sb);
      sym->big_off.push_back((int32_t)sl.size());
      for (int32_t s : L.snodes)
        if (!solve_small(s)) sl.push_back(s);
      sym->big_cnt.push_back((int32_t)sl.size() - sym->big_off.back());
    }
    GMRFB_CU(ctx, sym->d_solve_lists.upload(sl, st));
  }
  {
    std::vector<Task> tasks;
    sym->solve_levels.assign(S.levels.size(), gmrfb_sym::SolveLevel());
    int64_t poff = 0;
    std::vector<int64_t> part_off(S.nsuper, 0);
    std::vector<int32_t> nchunk(S.nsuper, 0);
    for (int32_t s = 0; s < S.nsuper; s++) {
      int r = S.front_order(s) - S.ncols(s);
      nchunk[s] = cdiv(r, SOLVE_BR_ROWS);
      part_off[s] = poff;
      poff += (int64_t)nchunk[s] * S.ncols(s) * SOLVE_NRC;
    }
    sym->partial_doubles = poff;
    auto base_task = [&](int32_t s, int k) {
      Task t = make_task();
      t.a = S.foff[s];
      t.lda = S.ld[s];
      t.M = S.front_order(s);
      t.N = S.ncols(s);
      t.K = k;
      t.ldb = S.sptr[s];
      t.b = sd[s].uoff;
      t.c = part_off[s];
      t.ldc = nchunk[s];
      t.aux0 = (int32_t)(S.rptr[s] & 0xffffffff);
      t.aux1 = (int32_t)(S.rptr[s] >> 32);
      // offset (in doubles) of the supernode's kept inverse blocks, carried in the bits of `alpha`; -1 = none
      const int64_t ws = sym->factor_plan.host.winv_slot[s];
      const int64_t woff = ws >= 0 ? ws * (int64_t)DINV_SLOT : -1;
      static_assert(sizeof(double) == sizeof(int64_t), "bit cast");
      std::memcpy(&t.alpha, &woff, sizeof(double));
      return t;
    };
    // wide supernodes keep their full inverse (see sparse_kernels.cu); GMRFB_WIDE_INV=0 turns the path off
    {
      const char* we = std::getenv("GMRFB_WIDE_INV");
      sym->wide_enabled = !(we && we[0] == '0');
      sym->wide.clear(), sym->wide_off.clear(), sym->wide_ld.clear();
      int64_t off = 0;
      const char* wm = std::getenv("GMRFB_WIDE_MIN");
      const int wide_min = wm ? std::max(65, std::atoi(wm)) : SOLVE_WIDE_MIN;
      if (sym->wide_enabled)
        for (int32_t s = 0; s < S.nsuper; s++) {
          // (beyond SOLVE_WIDE_MAX columns - the time-slice separators of a space-time precision - the s^2 doubles
          //  of an inverse and the s^3/3 flops of its TRTRI are no longer small against the front itself)
          if (solve_small(s) || S.ncols(s) < wide_min || S.ncols(s) > SOLVE_WIDE_MAX) continue;
          const int sc = S.ncols(s), ldw = (sc + 1) & ~1;
          sym->wide.push_back(s);
          sym->wide_off.push_back(off);
          sym->wide_ld.push_back(ldw);
          off += (int64_t)ldw * sc;
          off = (off + 15) & ~(int64_t)15;
        }
      sym->wide_doubles = off;
      // inverses + TRTRI scratch must stay a small addition to the frontal arena
      if (sym->wide.empty() || 2 * off > S.arena / 8) {
        sym->wide.clear(), sym->wide_off.clear(), sym->wide_ld.clear();
        sym->wide_doubles = 0;
        sym->wide_enabled = false;
      }
    }
    std::vector<int32_t> wide_idx(S.nsuper, -1);
    for (size_t i = 0; i < sym->wide.size(); i++) wide_idx[sym->wide[i]] = (int32_t)i;
    auto wide_task = [&](int32_t s) {  // solve-task fields, K = leading dimension of W_J, alpha = bits of its offset
      Task t = base_task(s, 0);
      const int32_t wi = wide_idx[s];
      t.K = sym->wide_ld[wi];
      t.aux0 = wi;  // (the row-list offset is not needed by the wide kernels)
      const int64_t wo = sym->wide_off[wi];
      std::memcpy(&t.alpha, &wo, sizeof(double));
      return t;
    };
    auto build_steps = [&](const std::vector<int32_t>& sn, bool skip_wide, std::vector<Launch>& fwd,
                           std::vector<Launch>& bwd) {
      auto skip = [&](int32_t s) { return solve_small(s) || (skip_wide && wide_idx[s] >= 0); };
      int maxs = 0;
      for (int32_t s : sn)
        if (!skip(s)) maxs = std::max(maxs, S.ncols(s));
      int nsteps = cdiv(maxs, 64);
      for (int k = 0; k < nsteps; k++) {
        Launch Lf{};
        Lf.kind = PK_FWD_LEVEL;
        Lf.task0 = (int32_t)tasks.size();
        Launch Lb = Lf;
        // forward step k: CTAs over the rows below block k (at least one CTA to solve and publish the block)
        for (int32_t s : sn) {
          int sc = S.ncols(s), d = S.front_order(s);
          if (k * 64 >= sc || skip(s)) continue;
          int nb = std::min(64, sc - k * 64);
          Task t = base_task(s, k);
          t.tile0 = Lf.grid;
          int ctas = std::max(1, cdiv(d - (k * 64 + nb), SOLVE_FS_ROWS));
          tasks.push_back(t);
          Lf.ntasks++;
          Lf.grid += ctas;
          double trap = (double)nb * (d - k * 64) - (double)nb * (nb - 1) / 2;
          Lf.bytes += 8.0 * trap;
          Lf.flops += 2.0 * trap;
        }
        fwd.push_back(Lf);
        Lb.kind = PK_BWD_LEVEL;
        Lb.task0 = (int32_t)tasks.size();
        for (int32_t s : sn) {
          int sc = S.ncols(s);
          if (k * 64 >= sc || skip(s)) continue;
          int nb = std::min(64, sc - k * 64);
          Task t = base_task(s, k);
          t.tile0 = Lb.grid;
          tasks.push_back(t);
          Lb.ntasks++;
          Lb.grid += k + 1;
          double tri = (double)nb * (k * 64) + (double)nb * (nb + 1) / 2;  // strip of L11 left of and incl. the block
          Lb.bytes += 8.0 * tri;
          Lb.flops += 2.0 * tri;
        }
        bwd.push_back(Lb);
      }
    };
    for (size_t l = 0; l < S.levels.size(); l++) {
      const auto& sn = S.levels[l].snodes;
      auto& SL = sym->solve_levels[l];
      build_steps(sn, false, SL.fwd_steps, SL.bwd_steps);
      if (sym->wide_enabled) {
        build_steps(sn, true, SL.fwd_steps_nw, SL.bwd_steps_nw);
        auto wide_launch = [&](int kind, auto ctas_of, auto bytes_of) {
          Launch L{};
          L.kind = kind;
          L.task0 = (int32_t)tasks.size();
          for (int32_t s : sn) {
            if (wide_idx[s] < 0) continue;
            const int c = ctas_of(s);
            if (c <= 0) continue;
            Task t = wide_task(s);
            t.tile0 = L.grid;
            tasks.push_back(t);
            L.ntasks++;
            L.grid += c;
            L.bytes += bytes_of(s);
            L.flops += bytes_of(s) / 4.0;
          }
          return L;
        };
        auto tri_bytes = [&](int32_t s) { double sc = S.ncols(s); return 8.0 * sc * (sc + 1) / 2; };
        SL.wide_trmv = wide_launch(PK_WIDE_FWD, [&](int32_t s) { return cdiv(S.ncols(s), SOLVE_WG_ROWS); }, tri_bytes);
        SL.wide_below = wide_launch(PK_WIDE_FWD, [&](int32_t s) { return cdiv(S.front_order(s) - S.ncols(s), SOLVE_WG_ROWS); },
                                    [&](int32_t s) { return 8.0 * (double)(S.front_order(s) - S.ncols(s)) * S.ncols(s); });
        SL.wide_bwd = wide_launch(PK_WIDE_BWD, [&](int32_t s) { return cdiv(S.ncols(s), 4); }, tri_bytes);
      }
      Launch Lr{};
      Lr.kind = PK_BWD_RPART;
      Lr.task0 = (int32_t)tasks.size();
      for (int32_t s : sn) {
        int sc = S.ncols(s), d = S.front_order(s), r = d - sc;
        if (r <= 0 || solve_small(s)) continue;
        Task t = base_task(s, 0);
        t.tile0 = Lr.grid;
        tasks.push_back(t);
        Lr.ntasks++;
        Lr.grid += cdiv(sc, 64) * nchunk[s];
        Lr.bytes += 8.0 * (double)r * sc;
        Lr.flops += 2.0 * (double)r * sc;
      }
      SL.rpart = Lr;
    }
    if (sym->wide_enabled) {
      Launch Ln{};
      Ln.kind = PK_WIDE_NORM;
      Ln.task0 = (int32_t)tasks.size();
      for (int32_t s : sym->wide) {
        Task t = wide_task(s);
        t.tile0 = Ln.grid;
        tasks.push_back(t);
        Ln.ntasks++;
        Ln.grid += cdiv(S.ncols(s), 8);
        Ln.bytes += 8.0 * (double)S.ncols(s) * (S.ncols(s) + 1);
      }
      sym->wide_norms = Ln;
      build_wide_inverse_plan(S, sym->wide, sym->wide_off, sym->wide_ld, sym->wide_doubles, sym->wide_plan.host);
      GMRFB_CU(ctx, sym->wide_plan.tasks.upload(sym->wide_plan.host.tasks, st));
      sym->wide_plan.ready = true;
    }
    GMRFB_CU(ctx, sym->d_solve_tasks.upload(tasks, st));
  }
  GMRFB_CU(ctx, sym->factor_plan.tasks.upload(sym->factor_plan.host.tasks, st));
  sym->factor_plan.ready = true;
  build_zero_plan(S, sym->zero_plan.host);
  GMRFB_CU(ctx, sym->zero_plan.tasks.upload(sym->zero_plan.host.tasks, st));
  sym->zero_plan.ready = true;
  sym->dev_ready = true;
  return GMRFB_OK;
}

static gmrfb_status sym_ensure_selinv(gmrfb_sym* sym) {
  if (sym->selinv_plan.ready) return GMRFB_OK;
  gmrfb_ctx* ctx = sym->ctx;
  if (sym->wide_enabled) {
    std::vector<int32_t> wide_idx(sym->S.nsuper, -1);
    for (size_t i = 0; i < sym->wide.size(); i++) wide_idx[sym->wide[i]] = (int32_t)i;
    build_selinv_plan(sym->S, sym->selinv_plan.host, &wide_idx, &sym->wide_off, &sym->wide_ld);
  } else {
    build_selinv_plan(sym->S, sym->selinv_plan.host);
  }
  GMRFB_CU(ctx, sym->selinv_plan.tasks.upload(sym->selinv_plan.host.tasks, ctx->stream));
  sym->selinv_plan.ready = true;
  return GMRFB_OK;
}

// ------------------------------------------------------------------------------------------- numeric ----
extern "C" gmrfb_status gmrfb_fac_create(gmrfb_sym* sym, gmrfb_fac** out) try {
  if (!sym || !out) return fail(sym ? sym->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fac_create: NULL argument");
  *out = nullptr;
  gmrfb_status rc = sym_ensure_device(sym);
  if (rc != GMRFB_OK) return rc;
  gmrfb_ctx* ctx = sym->ctx;
  std::unique_ptr<gmrfb_fac> f(new gmrfb_fac());
  f->sym = sym;
  f->ctx = ctx;
  const Symbolic& S = sym->S;
  GMRFB_CU(ctx, f->arena.alloc((size_t)std::max<int64_t>(S.arena, 1)));
  GMRFB_CU(ctx, f->nzval.alloc((size_t)std::max<int64_t>(S.nnzA, 1)));
  GMRFB_CU(ctx, f->xwork.alloc((size_t)std::max<int64_t>(S.n, 1) * SOLVE_NRC));
  GMRFB_CU(ctx, f->ywork.alloc((size_t)std::max<int64_t>(S.n, 1) * SOLVE_NRC));
  GMRFB_CU(ctx, f->owork.alloc((size_t)std::max<int64_t>(S.n, 1) * SOLVE_NRC));
  GMRFB_CU(ctx, f->partial.alloc((size_t)std::max<int64_t>(sym->partial_doubles, 1)));
  GMRFB_CU(ctx, f->bwork.alloc((size_t)std::max<int64_t>(S.n, 1) * SOLVE_NRC));
  GMRFB_CU(ctx, f->uvec.alloc((size_t)std::max<int64_t>(sym->uvec_rows, 1) * SOLVE_NRC));
  GMRFB_CU(ctx, f->dinv.alloc((size_t)std::max<int64_t>(sym->factor_plan.host.dinv, 1)));
  if (sym->wide_enabled) {
    GMRFB_CU(ctx, f->winv_full.alloc((size_t)(2 * sym->wide_doubles)));
    GMRFB_CU(ctx, f->wide_norms.alloc(2 * sym->wide.size()));
    f->wide_norms_host.assign(2 * sym->wide.size(), 0.0);
  }
  *out = f.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fac_destroy(gmrfb_fac* fac) try {
  if (!fac) return GMRFB_OK;
  cudaSetDevice(fac->ctx->device);
  cudaStreamSynchronize(fac->ctx->stream);
  delete fac;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_factorize_dev(gmrfb_fac* fac, const double* d_nzval) try {
  if (!fac || !d_nzval) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_factorize_dev: NULL argument");
  gmrfb_ctx* ctx = fac->ctx;
  gmrfb_sym* sym = fac->sym;
  const Symbolic& S = sym->S;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  fac->factored = false;
  fac->z_valid = false;
  fac->logdet_valid = false;
  auto body = [&]() -> gmrfb_status {
    // (the plan's first launch clears the parts of the fronts that are accumulated into; GMRFB_POISON=1 fills the whole
    //  arena with NaNs first — a debugging aid that proves nothing outside the cleared parts is ever read)
    if (ctx->poison) GMRFB_CU(ctx, cudaMemsetAsync(fac->arena.p, 0xff, fac->arena.n * sizeof(double), st));
    GMRFB_CU(ctx, cudaMemcpyAsync(ctx->d_info, ctx->d_info_init, sizeof(int), cudaMemcpyDeviceToDevice, st));
    {
      Arenas zar{{fac->arena.p, nullptr, nullptr, nullptr}};
      gmrfb_status zrc = run_plan(ctx, sym->zero_plan, zar, LaunchAux());
      if (zrc != GMRFB_OK) return zrc;
    }
    {
      ProfScope ps(ctx, PK_SCATTER, 0, (double)S.nnzA * 16.0 + (double)S.nnz_lower_A * 8.0);
      GMRFB_CU(ctx, launch_scatter_values(d_nzval, sym->d_amap.p, S.nnzA, fac->arena.p, st));
    }
    ctx->launches++;
    Arenas ar{{fac->arena.p, nullptr, nullptr, nullptr}};
    ar.dinv = fac->dinv.p;
    LaunchAux aux;
    aux.d_info = ctx->d_info;
    aux.d_relmap = sym->d_relmap.p;
    aux.d_snodes = sym->d_snodes.p;
    aux.d_child_idx = sym->d_child_idx.p;
    aux.d_sparent = sym->d_sparent.p;
    gmrfb_status frc = run_plan(ctx, sym->factor_plan, ar, aux);
    if (frc != GMRFB_OK || !sym->wide_enabled) return frc;
    // full inverses of the wide supernodes (the solves use them when they are well enough conditioned) and the two
    // infinity norms per supernode that decide it
    GMRFB_CU(ctx, cudaMemsetAsync(fac->wide_norms.p, 0, fac->wide_norms.n * sizeof(double), st));
    ar.p[AR_WINV] = fac->winv_full.p;
    frc = run_plan(ctx, sym->wide_plan, ar, aux);
    if (frc != GMRFB_OK) return frc;
    {
      const Launch& Ln = sym->wide_norms;
      ProfScope ps(ctx, PK_WIDE_NORM, 0, Ln.bytes, Ln.grid, Ln.ntasks);
      GMRFB_CU(ctx, launch_wide_norms(sym->d_solve_tasks.p + Ln.task0, Ln.ntasks, Ln.grid, fac->arena.p, fac->winv_full.p,
                                      fac->wide_norms.p, st));
    }
    ctx->launches++;
    return GMRFB_OK;
  };
  fac->wide_ok = false;
  gmrfb_status rc = run_graphed(ctx, sym->graphs, graph_key({1, (uint64_t)(uintptr_t)d_nzval, fac->buffers_key()}), body);
  if (rc != GMRFB_OK) return rc;
  int info = 0;
  GMRFB_CU(ctx, cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (sym->wide_enabled)
    GMRFB_CU(ctx, cudaMemcpyAsync(fac->wide_norms_host.data(), fac->wide_norms.p, fac->wide_norms.n * sizeof(double),
                                  cudaMemcpyDeviceToHost, st));
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (sym->wide_enabled && info == INT_MAX) {
    // cond_1(L_JJ) bounds the error the explicit inverse adds to a solve (~ cond * eps): beyond the threshold the
    // sweeps keep to the block-step substitution (GMRFB_WIDE_COND_MAX, default 1e5)
    const char* ce = std::getenv("GMRFB_WIDE_COND_MAX");
    const double cond_max = ce ? std::atof(ce) : 1e5;
    double c = 0.0;
    for (size_t i = 0; i < sym->wide.size(); i++)
      c = std::max(c, fac->wide_norms_host[2 * i] * fac->wide_norms_host[2 * i + 1]);
    fac->wide_cond = c;
    fac->wide_ok = std::isfinite(c) && c <= cond_max;
    if (std::getenv("GMRFB_WIDE_DEBUG"))
      fprintf(stderr, "[gmrfb] wide supernodes: %zu, max cond_1(L_JJ) = %.3e, inverse path %s\n", sym->wide.size(), c,
              fac->wide_ok ? "on" : "off");
  }
  if (info != INT_MAX) {
    fac->status = GMRFB_ERR_NOT_SPD;
    fac->fail_column = (info >= 0 && info < S.n) ? S.post[info] : -1;
    return fail(ctx, GMRFB_ERR_NOT_SPD,
                "matrix is not positive definite (pivot failed at permuted column " + std::to_string(fac->fail_column) + ")");
  }
  fac->status = GMRFB_OK;
  fac->fail_column = -1;
  fac->factored = true;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_factorize(gmrfb_fac* fac, const double* nzval) try {
  if (!fac || !nzval) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_factorize: NULL argument");
  gmrfb_ctx* ctx = fac->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  GMRFB_CU(ctx, cudaMemcpyAsync(fac->nzval.p, nzval, (size_t)fac->sym->S.nnzA * sizeof(double), cudaMemcpyHostToDevice,
                                ctx->stream));
  return gmrfb_factorize_dev(fac, fac->nzval.p);
}
GMRFB_ABI_CATCH

static gmrfb_status fac_diag_host(gmrfb_fac* fac, std::vector<double>& dl) {
  // diag(L) in the internal ordering
  gmrfb_ctx* ctx = fac->ctx;
  const Symbolic& S = fac->sym->S;
  DevBuf<double> d;
  GMRFB_CU(ctx, d.alloc((size_t)std::max<int64_t>(S.n, 1)));
  GMRFB_CU(ctx, launch_diag_L(fac->sym->d_snodes.p, S.nsuper, fac->arena.p, d.p, ctx->stream));
  ctx->launches++;
  dl.resize(S.n);
  GMRFB_CU(ctx, cudaMemcpyAsync(dl.data(), d.p, (size_t)S.n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}

extern "C" gmrfb_status gmrfb_fac_get_info(gmrfb_fac* fac, gmrfb_fac_info* info) try {
  if (!fac || !info) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fac_get_info: NULL argument");
  GMRFB_CU(fac->ctx, cudaSetDevice(fac->ctx->device));
  info->status = fac->status;
  info->fail_column = fac->fail_column;
  info->nnz_L = fac->sym->S.nnzL_stored;
  info->logdet = NAN;
  if (fac->factored) {
    if (!fac->logdet_valid) {
      std::vector<double> dl;
      gmrfb_status rc = fac_diag_host(fac, dl);
      if (rc != GMRFB_OK) return rc;
      double s = 0;
      for (double v : dl) s += std::log(v);
      fac->logdet = 2.0 * s;
      fac->logdet_valid = true;
    }
    info->logdet = fac->logdet;
  }
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fac_diag(gmrfb_fac* fac, double* diagL) try {
  if (!fac || !diagL) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fac_diag: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_fac_diag: no successful factorisation");
  GMRFB_CU(fac->ctx, cudaSetDevice(fac->ctx->device));
  std::vector<double> dl;
  gmrfb_status rc = fac_diag_host(fac, dl);
  if (rc != GMRFB_OK) return rc;
  const Symbolic& S = fac->sym->S;
  for (int64_t k = 0; k < S.n; k++) diagL[S.post[k]] = dl[k];
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fac_get_L(gmrfb_fac* fac, int32_t base, int32_t drop_zeros, int64_t* colptr,
                                        int64_t* rowval, double* nzval) try {
  if (!fac || !colptr) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fac_get_L: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_fac_get_L: no successful factorisation");
  gmrfb_ctx* ctx = fac->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const Symbolic& S = fac->sym->S;
  // L in the perm_user ordering: user column ku = post[k]; rows mapped the same way and sorted.
  const bool want_vals = rowval && nzval;
  std::vector<double> front;
  std::vector<std::pair<int64_t, double>> col;
  // first pass sizes (and second pass fills) column by column in user order: build per-internal-column data
  std::vector<std::vector<std::pair<int64_t, double>>> cols;
  if (want_vals || drop_zeros) cols.resize(S.n);
  std::vector<int64_t> cnt(S.n, 0);
  for (int32_t s = 0; s < S.nsuper; s++) {
    int d = S.front_order(s), sc = S.ncols(s), ld = S.ld[s];
    if (want_vals || drop_zeros) {
      front.resize((size_t)ld * sc);
      GMRFB_CU(ctx, cudaMemcpyAsync(front.data(), fac->arena.p + S.foff[s], front.size() * sizeof(double),
                                    cudaMemcpyDeviceToHost, ctx->stream));
      GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int j = 0; j < sc; j++) {
      int32_t k = S.sptr[s] + j;
      int64_t ku = S.post[k];
      for (int i = j; i < d; i++) {
        double v = (want_vals || drop_zeros) ? front[(size_t)j * ld + i] : 1.0;
        if (drop_zeros && v == 0.0 && i != j) continue;
        cnt[ku]++;
        if (want_vals || drop_zeros) cols[ku].push_back({(int64_t)S.post[S.rows[S.rptr[s] + i]], v});
      }
    }
  }
  colptr[0] = base;
  for (int64_t k = 0; k < S.n; k++) colptr[k + 1] = colptr[k] + cnt[k];
  if (want_vals) {
    for (int64_t k = 0; k < S.n; k++) {
      auto& c = cols[k];
      std::sort(c.begin(), c.end());
      int64_t o = colptr[k] - base;
      for (auto& e : c) {
        rowval[o] = e.first + base;
        nzval[o] = e.second;
        o++;
      }
    }
  }
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// ---------------------------------------------------------------------------------------------- solves ----
namespace {

// Level-scheduled sweeps.  Forward: working vector in `w` (destroyed), solution into `y`.  Backward: right-hand
// side in `t` (destroyed), solution into `xs`.  Vectors are n x nr (internal ordering, leading dimension n).
gmrfb_status sweep_fwd(gmrfb_fac* fac, double* w, double* y, int nr) {
  gmrfb_ctx* ctx = fac->ctx;
  gmrfb_sym* sym = fac->sym;
  const int64_t n = sym->S.n;
  const int nlev = (int)sym->S.levels.size();
  const bool wide = sym->wide_enabled && fac->wide_ok;
  for (int l = 0; l < nlev; l++) {
    if (sym->small_cnt[l] > 0) {  // fused: children's contributions, substitution and update vector by one warp each
      ProfScope ps(ctx, PK_FWD_SMALL, 0, sym->small_bytes[l], sym->small_cnt[l], sym->small_cnt[l]);
      GMRFB_CU(ctx, launch_fwd_small(sym->d_snodes.p, sym->d_solve_lists.p + sym->small_off[l], sym->small_cnt[l],
                                     sym->d_child_idx.p, sym->d_relmap.p, fac->arena.p, w, y, n, fac->uvec.p, nr, ctx->stream));
      ctx->launches++;
    }
    if (sym->big_cnt[l] > 0) {
      // leaves have no children: the same kernel just zeroes their update vectors
      ProfScope ps(ctx, PK_FWD_ASM, 0, 0, sym->big_cnt[l], sym->big_cnt[l]);
      GMRFB_CU(ctx, launch_fwd_assemble(sym->d_snodes.p, sym->d_solve_lists.p + sym->big_off[l], sym->big_cnt[l],
                                        sym->d_child_idx.p, sym->d_relmap.p, w, n, fac->uvec.p, ctx->stream));
      ctx->launches++;
    }
    const auto& SL = sym->solve_levels[l];
    for (const Launch& L : (wide ? SL.fwd_steps_nw : SL.fwd_steps)) {
      ProfScope ps(ctx, PK_FWD_LEVEL, L.flops * nr, L.bytes, L.grid, L.ntasks);
      GMRFB_CU(ctx, launch_fwd_step(sym->d_solve_tasks.p + L.task0, L.ntasks, L.grid, fac->arena.p, w, y, n,
                                    fac->uvec.p, nr, fac->dinv.p, ctx->stream));
      ctx->launches++;
    }
    if (wide && SL.wide_trmv.grid > 0) {  // y_J = W_J x_J, then u_J -= L21 y_J
      {
        const Launch& L = SL.wide_trmv;
        ProfScope ps(ctx, PK_WIDE_FWD, L.flops * nr, L.bytes, L.grid, L.ntasks);
        GMRFB_CU(ctx, launch_wide_fwd(sym->d_solve_tasks.p + L.task0, L.ntasks, L.grid, 0, fac->arena.p, fac->winv_full.p, w,
                                      y, n, fac->uvec.p, nr, ctx->stream));
        ctx->launches++;
      }
      if (SL.wide_below.grid > 0) {
        const Launch& L = SL.wide_below;
        ProfScope ps(ctx, PK_WIDE_FWD, L.flops * nr, L.bytes, L.grid, L.ntasks);
        GMRFB_CU(ctx, launch_wide_fwd_below(sym->d_solve_tasks.p + L.task0, L.ntasks, L.grid, fac->arena.p, y, n, fac->uvec.p,
                                            nr, ctx->stream));
        ctx->launches++;
      }
    }
  }
  return GMRFB_OK;
}

gmrfb_status sweep_bwd(gmrfb_fac* fac, double* t, double* xs, int nr) {
  gmrfb_ctx* ctx = fac->ctx;
  gmrfb_sym* sym = fac->sym;
  const int64_t n = sym->S.n;
  const int nlev = (int)sym->S.levels.size();
  const bool wide = sym->wide_enabled && fac->wide_ok;
  for (int l = nlev - 1; l >= 0; l--) {
    const auto& SL = sym->solve_levels[l];
    if (SL.rpart.grid > 0) {
      ProfScope ps(ctx, PK_BWD_RPART, SL.rpart.flops * nr, SL.rpart.bytes, SL.rpart.grid, SL.rpart.ntasks);
      GMRFB_CU(ctx, launch_bwd_rpart(sym->d_solve_tasks.p + SL.rpart.task0, SL.rpart.ntasks, SL.rpart.grid,
                                     fac->arena.p, sym->d_rows.p, xs, n, fac->partial.p, nr, ctx->stream));
      ctx->launches++;
    }
    const auto& bsteps = wide ? SL.bwd_steps_nw : SL.bwd_steps;
    for (int k = (int)bsteps.size() - 1; k >= 0; k--) {
      const Launch& L = bsteps[k];
      ProfScope ps(ctx, PK_BWD_LEVEL, L.flops * nr, L.bytes, L.grid, L.ntasks);
      GMRFB_CU(ctx, launch_bwd_step(sym->d_solve_tasks.p + L.task0, L.ntasks, L.grid, fac->arena.p, t, xs, n,
                                    fac->partial.p, nr, fac->dinv.p, ctx->stream));
      ctx->launches++;
    }
    if (wide && SL.wide_bwd.grid > 0) {  // x_J = W_J' (t_J - R-part partial sums)
      const Launch& L = SL.wide_bwd;
      ProfScope ps(ctx, PK_WIDE_BWD, L.flops * nr, L.bytes, L.grid, L.ntasks);
      GMRFB_CU(ctx, launch_wide_bwd(sym->d_solve_tasks.p + L.task0, L.ntasks, L.grid, fac->winv_full.p, t, xs, n,
                                    fac->partial.p, nr, ctx->stream));
      ctx->launches++;
    }
    if (sym->small_cnt[l] > 0) {
      ProfScope ps(ctx, PK_BWD_SMALL, 0, sym->small_bytes[l], sym->small_cnt[l], sym->small_cnt[l]);
      GMRFB_CU(ctx, launch_bwd_small(sym->d_snodes.p, sym->d_solve_lists.p + sym->small_off[l], sym->small_cnt[l],
                                     sym->d_rows.p, fac->arena.p, t, xs, n, nr, ctx->stream));
      ctx->launches++;
    }
  }
  return GMRFB_OK;
}


// ---- panel (multi-right-hand-side) sweeps: solve_mr.cu + build_solve_mr_plans ----
gmrfb_status sym_get_mr(gmrfb_sym* sym, int nr, gmrfb_sym::MrPlans** out) {
  gmrfb_ctx* ctx = sym->ctx;
  auto it = sym->mr_plans.find(nr);
  if (it == sym->mr_plans.end()) {
    if (sym->mr_plans.size() >= 8) {  // bounded cache of panel widths; queued launches may still read the old task lists
      GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
      sym->mr_plans.clear();
    }
    std::unique_ptr<gmrfb_sym::MrPlans> mp(new gmrfb_sym::MrPlans());
    mp->nr = nr;
    mp->ldk = (nr + 7) & ~7;
    build_solve_mr_plans(sym->S, sym->factor_plan.host.winv_slot, nr, mp->ldk, mp->fwd.host, mp->bwd.host);
    GMRFB_CU(ctx, mp->fwd.tasks.upload(mp->fwd.host.tasks, ctx->stream));
    GMRFB_CU(ctx, mp->bwd.tasks.upload(mp->bwd.host.tasks, ctx->stream));
    mp->fwd.ready = mp->bwd.ready = true;
    it = sym->mr_plans.emplace(nr, std::move(mp)).first;
  }
  *out = it->second.get();
  return GMRFB_OK;
}

gmrfb_status fac_ensure_mr(gmrfb_fac* fac, bool staging) {
  gmrfb_ctx* ctx = fac->ctx;
  const int64_t n = std::max<int64_t>(fac->sym->S.n, 1);
  if (!fac->mr_x.p) GMRFB_CU(ctx, fac->mr_x.alloc((size_t)n * MR_MAX));
  if (!fac->mr_u.p) GMRFB_CU(ctx, fac->mr_u.alloc((size_t)std::max<int64_t>(fac->sym->uvec_rows, 1) * MR_MAX));
  if (staging && !fac->mr_io.p) GMRFB_CU(ctx, fac->mr_io.alloc((size_t)n * MR_MAX));
  return GMRFB_OK;
}

// The panel fac->mr_x (nr x n, node-major, internal ordering) through L^{-1} and / or L^{-T}, in place.
gmrfb_status sweep_panel(gmrfb_fac* fac, bool fwd, bool bwd, int nr) {
  gmrfb_sym* sym = fac->sym;
  gmrfb_sym::MrPlans* mp = nullptr;
  gmrfb_status rc = sym_get_mr(sym, nr, &mp);
  if (rc != GMRFB_OK) return rc;
  Arenas ar{{fac->arena.p, fac->mr_x.p, fac->mr_u.p, nullptr}};
  ar.dinv = fac->dinv.p;
  LaunchAux aux;
  aux.d_relmap = sym->d_relmap.p;
  aux.d_snodes = sym->d_snodes.p;
  aux.d_child_idx = sym->d_child_idx.p;
  aux.d_rows = sym->d_rows.p;
  aux.nr = nr;
  aux.ldk = mp->ldk;
  auto body = [&]() -> gmrfb_status {
    gmrfb_status r = GMRFB_OK;
    if (fwd) r = run_plan(fac->ctx, mp->fwd, ar, aux);
    if (r == GMRFB_OK && bwd) r = run_plan(fac->ctx, mp->bwd, ar, aux);
    return r;
  };
  return run_graphed(fac->ctx, sym->graphs,
                     graph_key({3, (uint64_t)nr, (uint64_t)fwd, (uint64_t)bwd, fac->buffers_key(),
                                (uint64_t)(uintptr_t)mp->fwd.tasks.p}), body);
}

// split nrhs right-hand sides into the fewest, equally wide panels of at most MR_MAX columns
inline int64_t panel_width(int64_t nrhs) {
  const int64_t chunks = (nrhs + MR_MAX - 1) / MR_MAX;
  return (nrhs + chunks - 1) / chunks;
}

struct ModeSpec {
  bool fwd, bwd;
  bool in_perm;   // gather input through perm (original ordering) instead of post (perm ordering)
  bool out_perm;  // scatter output through perm instead of post
};
bool mode_spec(int mode, ModeSpec& m) {
  switch (mode) {
    case GMRFB_SOLVE_A: m = {true, true, true, true}; return true;
    case GMRFB_SOLVE_PTL: m = {true, false, true, false}; return true;
    case GMRFB_SOLVE_UP: m = {false, true, false, true}; return true;
    case GMRFB_SOLVE_L: m = {true, false, false, false}; return true;
    case GMRFB_SOLVE_LT: m = {false, true, false, false}; return true;
  }
  return false;
}

// d_X (device, n x nrhs, ldx) solved in place, NRC columns at a time; optional mean added on output.
gmrfb_status solve_device(gmrfb_fac* fac, int mode, const double* d_in, int64_t ldin, double* d_out, int64_t ldout,
                          int64_t nrhs, const double* d_mean) {
  gmrfb_ctx* ctx = fac->ctx;
  gmrfb_sym* sym = fac->sym;
  const int64_t n = sym->S.n;
  ModeSpec m;
  if (!mode_spec(mode, m)) return fail(ctx, GMRFB_ERR_INVALID, "unknown solve mode");
  if (nrhs > SOLVE_NRC) {
    // batches: one sweep over L per panel of up to MR_MAX right-hand sides
    gmrfb_status rc = fac_ensure_mr(fac, false);
    if (rc != GMRFB_OK) return rc;
    const int64_t pw = panel_width(nrhs);
    for (int64_t c0 = 0; c0 < nrhs; c0 += pw) {
      const int nr = (int)std::min<int64_t>(pw, nrhs - c0), ldk = (nr + 7) & ~7;
      {
        ProfScope ps(ctx, PK_PERM_MR, 0, 16.0 * n * nr);
        GMRFB_CU(ctx, launch_mr_perm_in(d_in + c0 * ldin, ldin, fac->mr_x.p, ldk, m.in_perm ? sym->d_perm.p : sym->d_post.p, n,
                                        nr, ctx->stream));
      }
      ctx->launches++;
      rc = sweep_panel(fac, m.fwd, m.bwd, nr);
      if (rc != GMRFB_OK) return rc;
      {
        ProfScope ps(ctx, PK_PERM_MR, 0, 16.0 * n * nr);
        GMRFB_CU(ctx, launch_mr_perm_out(fac->mr_x.p, ldk, d_out + c0 * ldout, ldout,
                                         m.out_perm ? sym->d_perm.p : sym->d_post.p, n, nr, d_mean, ctx->stream));
      }
      ctx->launches++;
    }
    return GMRFB_OK;
  }
  for (int64_t c0 = 0; c0 < nrhs; c0 += SOLVE_NRC) {
    int nr = (int)std::min<int64_t>(SOLVE_NRC, nrhs - c0);
    auto body = [&]() -> gmrfb_status {
      // fwd: xwork -> ywork;  bwd: ywork -> xwork
      double* first = m.fwd ? fac->xwork.p : fac->ywork.p;
      GMRFB_CU(ctx, launch_perm_gather(d_in + c0 * ldin, ldin, first, n, m.in_perm ? sym->d_perm.p : sym->d_post.p, n, nr,
                                       ctx->stream));
      ctx->launches++;
      gmrfb_status rc = GMRFB_OK;
      if (m.fwd) rc = sweep_fwd(fac, fac->xwork.p, fac->ywork.p, nr);
      if (rc != GMRFB_OK) return rc;
      if (m.bwd) rc = sweep_bwd(fac, fac->ywork.p, fac->xwork.p, nr);
      if (rc != GMRFB_OK) return rc;
      const double* result = m.bwd ? fac->xwork.p : fac->ywork.p;
      GMRFB_CU(ctx, launch_perm_scatter(result, n, d_out + c0 * ldout, ldout, m.out_perm ? sym->d_perm.p : sym->d_post.p,
                                        n, nr, d_mean, ctx->stream));
      ctx->launches++;
      return GMRFB_OK;
    };
    gmrfb_status rc = run_graphed(ctx, sym->graphs,
                                  graph_key({2, (uint64_t)mode, (uint64_t)(uintptr_t)(d_in + c0 * ldin), (uint64_t)ldin,
                                             (uint64_t)(uintptr_t)(d_out + c0 * ldout), (uint64_t)ldout, (uint64_t)nr,
                                             (uint64_t)(uintptr_t)d_mean, fac->buffers_key(), (uint64_t)fac->wide_ok}), body);
    if (rc != GMRFB_OK) return rc;
  }
  return GMRFB_OK;
}

}  // namespace

extern "C" gmrfb_status gmrfb_solve_dev(gmrfb_fac* fac, int32_t mode, double* d_X, int64_t ldx, int64_t nrhs) try {
  if (!fac || !d_X) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_solve_dev: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_solve: no successful factorisation");
  if (ldx < fac->sym->S.n || nrhs < 0) return fail(fac->ctx, GMRFB_ERR_INVALID, "gmrfb_solve: bad ldx/nrhs");
  GMRFB_CU(fac->ctx, cudaSetDevice(fac->ctx->device));
  // in-place: stage each chunk through bwork so gather/scatter never alias
  gmrfb_ctx* ctx = fac->ctx;
  const int64_t n = fac->sym->S.n;
  if (nrhs > SOLVE_NRC)  // panel path: the gather into the panel finishes before the scatter back starts
    return solve_device(fac, mode, d_X, ldx, d_X, ldx, nrhs, nullptr);
  for (int64_t c0 = 0; c0 < nrhs; c0 += SOLVE_NRC) {
    int nr = (int)std::min<int64_t>(SOLVE_NRC, nrhs - c0);
    GMRFB_CU(ctx, cudaMemcpy2DAsync(fac->bwork.p, n * sizeof(double), d_X + c0 * ldx, ldx * sizeof(double),
                                    n * sizeof(double), nr, cudaMemcpyDeviceToDevice, ctx->stream));
    gmrfb_status rc = solve_device(fac, mode, fac->bwork.p, n, d_X + c0 * ldx, ldx, nr, nullptr);
    if (rc != GMRFB_OK) return rc;
  }
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_solve(gmrfb_fac* fac, int32_t mode, double* X, int64_t ldx, int64_t nrhs) try {
  if (!fac || !X) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_solve: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_solve: no successful factorisation");
  gmrfb_ctx* ctx = fac->ctx;
  const int64_t n = fac->sym->S.n;
  if (ldx < n || nrhs < 0) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_solve: bad ldx/nrhs");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  DevBuf<double>& out = fac->owork;  // persistent staging: no cudaMalloc/cudaFree on the solve path
  if (nrhs > SOLVE_NRC) {
    gmrfb_status rc = fac_ensure_mr(fac, true);
    if (rc != GMRFB_OK) return rc;
    const int64_t pw = panel_width(nrhs);
    for (int64_t c0 = 0; c0 < nrhs; c0 += pw) {
      const int64_t nr = std::min<int64_t>(pw, nrhs - c0);
      GMRFB_CU(ctx, cudaMemcpy2DAsync(fac->mr_io.p, n * sizeof(double), X + c0 * ldx, ldx * sizeof(double),
                                      n * sizeof(double), nr, cudaMemcpyHostToDevice, ctx->stream));
      rc = solve_device(fac, mode, fac->mr_io.p, n, fac->mr_io.p, n, nr, nullptr);
      if (rc != GMRFB_OK) return rc;
      GMRFB_CU(ctx, cudaMemcpy2DAsync(X + c0 * ldx, ldx * sizeof(double), fac->mr_io.p, n * sizeof(double),
                                      n * sizeof(double), nr, cudaMemcpyDeviceToHost, ctx->stream));
      GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return GMRFB_OK;
  }
  for (int64_t c0 = 0; c0 < nrhs; c0 += SOLVE_NRC) {
    int nr = (int)std::min<int64_t>(SOLVE_NRC, nrhs - c0);
    GMRFB_CU(ctx, cudaMemcpy2DAsync(fac->bwork.p, n * sizeof(double), X + c0 * ldx, ldx * sizeof(double),
                                    n * sizeof(double), nr, cudaMemcpyHostToDevice, ctx->stream));
    gmrfb_status rc = solve_device(fac, mode, fac->bwork.p, n, out.p, n, nr, nullptr);
    if (rc != GMRFB_OK) return rc;
    GMRFB_CU(ctx, cudaMemcpy2DAsync(X + c0 * ldx, ldx * sizeof(double), out.p, n * sizeof(double), n * sizeof(double), nr,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_solve_refined(gmrfb_fac* fac, const gmrfb_spm* Q, double* X, int64_t ldx, int64_t nrhs,
                                            int32_t max_iter, double* resid_out) try {
  if (!fac || !Q || !X) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_solve_refined: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_solve_refined: no successful factorisation");
  gmrfb_ctx* ctx = fac->ctx;
  gmrfb_sym* sym = fac->sym;
  const int64_t n = sym->S.n;
  if (Q->m != n || Q->n != n) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_solve_refined: Q has the wrong shape");
  if (ldx < n || nrhs < 0 || max_iter < 0) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_solve_refined: bad ldx/nrhs/max_iter");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // workspaces (n doubles each, original ordering): b, x, r / correction, best iterate
  if (fac->refine_ws.n < (size_t)(4 * std::max<int64_t>(n, 1))) GMRFB_CU(ctx, fac->refine_ws.alloc((size_t)(4 * std::max<int64_t>(n, 1))));
  double *db = fac->refine_ws.p, *dx = db + n, *dr = dx + n, *dbest = dr + n;
  auto dot_host = [&](const double* a, double* out) -> gmrfb_status {
    GMRFB_CU(ctx, launch_dot(a, a, n, ctx->d_scalar, st));
    ctx->launches++;
    GMRFB_CU(ctx, cudaMemcpyAsync(out, ctx->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, st));
    GMRFB_CU(ctx, cudaStreamSynchronize(st));
    return GMRFB_OK;
  };
  for (int64_t c = 0; c < nrhs; c++) {
    GMRFB_CU(ctx, cudaMemcpyAsync(db, X + c * ldx, n * sizeof(double), cudaMemcpyHostToDevice, st));
    double bb = 0;
    gmrfb_status rc = dot_host(db, &bb);
    if (rc != GMRFB_OK) return rc;
    rc = solve_device(fac, GMRFB_SOLVE_A, db, n, dx, n, 1, nullptr);
    if (rc != GMRFB_OK) return rc;
    double best = -1.0;
    for (int it = 0;; it++) {
      // r = b - Q x   (Q symmetric: its CSC columns double as rows)
      GMRFB_CU(ctx, cudaMemcpyAsync(dr, db, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      GMRFB_CU(ctx, launch_spmv_rows(n, Q->d_colptr.p, Q->d_rowidx.p, Q->d_val.p, dx, dr, -1.0, 1.0, st));
      ctx->launches++;
      double rr = 0;
      rc = dot_host(dr, &rr);
      if (rc != GMRFB_OK) return rc;
      const double res = bb > 0 ? std::sqrt(rr / bb) : std::sqrt(rr);
      if (best >= 0 && !(res < best)) break;  // no further progress: keep the previous iterate
      best = res;
      GMRFB_CU(ctx, cudaMemcpyAsync(dbest, dx, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      if (it >= max_iter || res == 0.0) break;
      rc = solve_device(fac, GMRFB_SOLVE_A, dr, n, dr, n, 1, nullptr);  // correction, in place
      if (rc != GMRFB_OK) return rc;
      GMRFB_CU(ctx, launch_axpby(n, 1.0, dx, 1.0, dr, dx, st));
      ctx->launches++;
    }
    if (resid_out) resid_out[c] = best;
    GMRFB_CU(ctx, cudaMemcpyAsync(X + c * ldx, dbest, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    GMRFB_CU(ctx, cudaStreamSynchronize(st));
  }
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_sample(gmrfb_fac* fac, const double* mean, const double* Z, int64_t ldz, double* X,
                                     int64_t ldx, int64_t nrhs) try {
  if (!fac || !Z || !X) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_sample: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_sample: no successful factorisation");
  gmrfb_ctx* ctx = fac->ctx;
  const int64_t n = fac->sym->S.n;
  if (ldz < n || ldx < n || nrhs < 0) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_sample: bad leading dimension");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  DevBuf<double>& out = fac->owork;
  if (mean) {  // persistent in the handle: no allocation (and no device-wide synchronisation of a release) per call
    if (!fac->meanbuf.p) GMRFB_CU(ctx, fac->meanbuf.alloc((size_t)std::max<int64_t>(n, 1)));
    GMRFB_CU(ctx, cudaMemcpyAsync(fac->meanbuf.p, mean, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  const double* dmean = mean ? fac->meanbuf.p : nullptr;
  if (nrhs > SOLVE_NRC) {
    gmrfb_status rc = fac_ensure_mr(fac, true);
    if (rc != GMRFB_OK) return rc;
    const int64_t pw = panel_width(nrhs);
    for (int64_t c0 = 0; c0 < nrhs; c0 += pw) {
      const int64_t nr = std::min<int64_t>(pw, nrhs - c0);
      GMRFB_CU(ctx, cudaMemcpy2DAsync(fac->mr_io.p, n * sizeof(double), Z + c0 * ldz, ldz * sizeof(double),
                                      n * sizeof(double), nr, cudaMemcpyHostToDevice, ctx->stream));
      rc = solve_device(fac, GMRFB_SOLVE_UP, fac->mr_io.p, n, fac->mr_io.p, n, nr, dmean);
      if (rc != GMRFB_OK) return rc;
      GMRFB_CU(ctx, cudaMemcpy2DAsync(X + c0 * ldx, ldx * sizeof(double), fac->mr_io.p, n * sizeof(double),
                                      n * sizeof(double), nr, cudaMemcpyDeviceToHost, ctx->stream));
      GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return GMRFB_OK;
  }
  for (int64_t c0 = 0; c0 < nrhs; c0 += SOLVE_NRC) {
    int nr = (int)std::min<int64_t>(SOLVE_NRC, nrhs - c0);
    GMRFB_CU(ctx, cudaMemcpy2DAsync(fac->bwork.p, n * sizeof(double), Z + c0 * ldz, ldz * sizeof(double),
                                    n * sizeof(double), nr, cudaMemcpyHostToDevice, ctx->stream));
    gmrfb_status rc = solve_device(fac, GMRFB_SOLVE_UP, fac->bwork.p, n, out.p, n, nr, dmean);
    if (rc != GMRFB_OK) return rc;
    GMRFB_CU(ctx, cudaMemcpy2DAsync(X + c0 * ldx, ldx * sizeof(double), out.p, n * sizeof(double), n * sizeof(double), nr,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// ------------------------------------------------------------------------------------ marginal variances ----
static gmrfb_status selinv_run(gmrfb_fac* fac) {
  gmrfb_ctx* ctx = fac->ctx;
  gmrfb_sym* sym = fac->sym;
  gmrfb_status rc = sym_ensure_selinv(sym);
  if (rc != GMRFB_OK) return rc;
  if (!fac->zarena.p) GMRFB_CU(ctx, fac->zarena.alloc(fac->arena.n));
  if (!fac->zdiag.p) GMRFB_CU(ctx, fac->zdiag.alloc((size_t)std::max<int64_t>(sym->S.n, 1)));
  if (!fac->zwork.p) GMRFB_CU(ctx, fac->zwork.alloc((size_t)std::max<int64_t>(sym->selinv_plan.host.scratch, 1)));
  // the selected inversion has its own inverse-block scratch: fac->dinv keeps the factor's inverses for the solves
  if ((int64_t)fac->dinv_sel.n < sym->selinv_plan.host.dinv)
    GMRFB_CU(ctx, fac->dinv_sel.alloc((size_t)std::max<int64_t>(sym->selinv_plan.host.dinv, 1)));
  Arenas ar{{fac->arena.p, fac->zarena.p, fac->zwork.p, fac->winv_full.p}};
  ar.dinv = fac->dinv_sel.p;
  LaunchAux aux;
  aux.d_info = ctx->d_info;
  aux.d_relmap = sym->d_relmap.p;
  aux.d_out = fac->zdiag.p;
  aux.d_snodes = sym->d_snodes.p;
  aux.d_child_idx = sym->d_child_idx.p;
  aux.d_sparent = sym->d_sparent.p;
  rc = run_graphed(ctx, sym->graphs, graph_key({4, fac->buffers_key()}),
                   [&]() -> gmrfb_status { return run_plan(ctx, sym->selinv_plan, ar, aux); });
  if (rc != GMRFB_OK) return rc;
  fac->z_valid = true;
  return GMRFB_OK;
}

extern "C" gmrfb_status gmrfb_var_selinv_dev(gmrfb_fac* fac, double* d_var_out) try {
  if (!fac || !d_var_out) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_var_selinv: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_var_selinv: no successful factorisation");
  gmrfb_ctx* ctx = fac->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  gmrfb_status rc = selinv_run(fac);
  if (rc != GMRFB_OK) return rc;
  const int64_t n = fac->sym->S.n;
  GMRFB_CU(ctx, launch_perm_scatter(fac->zdiag.p, n, d_var_out, n, fac->sym->d_perm.p, n, 1, nullptr, ctx->stream));
  ctx->launches++;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_var_selinv(gmrfb_fac* fac, double* var_out) try {
  if (!fac || !var_out) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_var_selinv: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_var_selinv: no successful factorisation");
  gmrfb_ctx* ctx = fac->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t n = fac->sym->S.n;
  DevBuf<double>& out = fac->owork;
  gmrfb_status rc = gmrfb_var_selinv_dev(fac, out.p);
  if (rc != GMRFB_OK) return rc;
  GMRFB_CU(ctx, cudaMemcpyAsync(var_out, out.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_selinv_entries(gmrfb_fac* fac, int32_t base, int64_t count, const int64_t* rows,
                                             const int64_t* cols, double* out) try {
  if (!fac || (count > 0 && (!rows || !cols || !out)))
    return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_selinv_entries: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_selinv_entries: no successful factorisation");
  gmrfb_ctx* ctx = fac->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const Symbolic& S = fac->sym->S;
  std::vector<int64_t> map(count);
  for (int64_t k = 0; k < count; k++) {
    int64_t r = rows[k] - base, c = cols[k] - base;
    if (r < 0 || r >= S.n || c < 0 || c >= S.n) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_selinv_entries: index out of range");
    int32_t a = S.iperm[r], b = S.iperm[c];
    if (a < b) std::swap(a, b);
    int32_t s = S.snode[b];
    int32_t f = S.sptr[s], l = S.sptr[s + 1] - 1;
    int64_t lr;
    if (a <= l) {
      lr = a - f;
    } else {
      auto bg = S.rows.begin() + S.rptr[s] + (l - f + 1), en = S.rows.begin() + S.rptr[s + 1];
      auto it = std::lower_bound(bg, en, a);
      if (it == en || *it != a)
        return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_selinv_entries: entry lies outside the filled pattern of the factor");
      lr = it - (S.rows.begin() + S.rptr[s]);
    }
    map[k] = S.foff[s] + (int64_t)(b - f) * S.ld[s] + lr;
  }
  if (!fac->z_valid) {
    gmrfb_status rc = selinv_run(fac);
    if (rc != GMRFB_OK) return rc;
  }
  if (count == 0) return GMRFB_OK;
  DevBuf<int64_t> dmap;
  DevBuf<double> dout;
  GMRFB_CU(ctx, dmap.upload(map, ctx->stream));
  GMRFB_CU(ctx, dout.alloc((size_t)count));
  GMRFB_CU(ctx, launch_gather_values(fac->zarena.p, dmap.p, count, dout.p, ctx->stream));
  ctx->launches++;
  GMRFB_CU(ctx, cudaMemcpyAsync(out, dout.p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

static gmrfb_status var_rbmc_core(gmrfb_fac* fac, const gmrfb_spm* Q, const double* Z, int64_t ldz, int64_t nsamp,
                                  double* var_out, bool out_on_device);

extern "C" gmrfb_status gmrfb_var_rbmc(gmrfb_fac* fac, const gmrfb_spm* Q, const double* Z, int64_t ldz,
                                       int64_t nsamp, double* var_out) try {
  return var_rbmc_core(fac, Q, Z, ldz, nsamp, var_out, false);
}
GMRFB_ABI_CATCH
extern "C" gmrfb_status gmrfb_var_rbmc_dev(gmrfb_fac* fac, const gmrfb_spm* Q, const double* Z, int64_t ldz,
                                           int64_t nsamp, double* d_var_out) try {
  return var_rbmc_core(fac, Q, Z, ldz, nsamp, d_var_out, true);
}
GMRFB_ABI_CATCH

static gmrfb_status var_rbmc_core(gmrfb_fac* fac, const gmrfb_spm* Q, const double* Z, int64_t ldz, int64_t nsamp,
                                  double* var_out, bool out_on_device) {
  if (!fac || !Q || !Z || !var_out) return fail(fac ? fac->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_var_rbmc: NULL argument");
  if (!fac->factored) return fail(fac->ctx, GMRFB_ERR_STATE, "gmrfb_var_rbmc: no successful factorisation");
  gmrfb_ctx* ctx = fac->ctx;
  gmrfb_sym* sym = fac->sym;
  const int64_t n = sym->S.n;
  if (Q->m != n || Q->n != n) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_var_rbmc: Q has the wrong shape");
  if (ldz < n || nsamp <= 0) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_var_rbmc: bad ldz/nsamp");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t ldk = (nsamp + 3) & ~(int64_t)3;
  // sample panel (node-major, original ordering) and result: persistent in the handle, grown on demand
  if ((int64_t)fac->rbmc_x.n < ldk * std::max<int64_t>(n, 1)) GMRFB_CU(ctx, fac->rbmc_x.alloc((size_t)ldk * std::max<int64_t>(n, 1)));
  DevBuf<double>& Xs = fac->rbmc_x;
  DevBuf<double>& dvar = fac->owork;
  if (nsamp > SOLVE_NRC) {
    // all samples of a panel in one backward sweep over L (RBMCStrategy(50): one sweep instead of 13)
    gmrfb_status rc = fac_ensure_mr(fac, true);
    if (rc != GMRFB_OK) return rc;
    const int64_t pw = panel_width(nsamp);
    for (int64_t c0 = 0; c0 < nsamp; c0 += pw) {
      const int nr = (int)std::min<int64_t>(pw, nsamp - c0), lk = (nr + 7) & ~7;
      // Z may be a host or a device pointer (unified addressing): normals drawn on the device need no PCIe round trip
      GMRFB_CU(ctx, cudaMemcpy2DAsync(fac->mr_io.p, n * sizeof(double), Z + c0 * ldz, ldz * sizeof(double),
                                      n * sizeof(double), nr, cudaMemcpyDefault, ctx->stream));
      {
        ProfScope ps(ctx, PK_PERM_MR, 0, 16.0 * n * nr);
        GMRFB_CU(ctx, launch_mr_perm_in(fac->mr_io.p, n, fac->mr_x.p, lk, sym->d_post.p, n, nr, ctx->stream));
      }
      ctx->launches++;
      rc = sweep_panel(fac, false, true, nr);
      if (rc != GMRFB_OK) return rc;
      {
        ProfScope ps(ctx, PK_PERM_MR, 0, 16.0 * n * nr);
        GMRFB_CU(ctx, launch_mr_perm_nodemajor(fac->mr_x.p, lk, Xs.p, ldk, sym->d_perm.p, n, (int)c0, nr, ctx->stream));
      }
      ctx->launches++;
    }
  } else
  for (int64_t c0 = 0; c0 < nsamp; c0 += SOLVE_NRC) {
    int nr = (int)std::min<int64_t>(SOLVE_NRC, nsamp - c0);
    GMRFB_CU(ctx, cudaMemcpy2DAsync(fac->bwork.p, n * sizeof(double), Z + c0 * ldz, ldz * sizeof(double),
                                    n * sizeof(double), nr, cudaMemcpyDefault, ctx->stream));
    GMRFB_CU(ctx, launch_perm_gather(fac->bwork.p, n, fac->ywork.p, n, sym->d_post.p, n, nr, ctx->stream));
    ctx->launches++;
    gmrfb_status rc = sweep_bwd(fac, fac->ywork.p, fac->xwork.p, nr);
    if (rc != GMRFB_OK) return rc;
    GMRFB_CU(ctx, launch_perm_scatter_nodemajor(fac->xwork.p, n, Xs.p, ldk, sym->d_perm.p, n, (int)c0, nr, ctx->stream));
    ctx->launches++;
  }
  // Q is symmetric: its CSC columns double as rows
  GMRFB_CU(ctx, launch_rbmc(n, Q->d_colptr.p, Q->d_rowidx.p, Q->d_val.p, Xs.p, ldk, (int)nsamp,
                            out_on_device ? var_out : dvar.p, ctx->stream));
  ctx->launches++;
  if (!out_on_device) {
    GMRFB_CU(ctx, cudaMemcpyAsync(var_out, dvar.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return GMRFB_OK;
}
