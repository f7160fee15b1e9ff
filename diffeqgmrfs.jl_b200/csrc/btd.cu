// C ABI: dense block-tridiagonal Cholesky (space-time GMRFs with implicit time stepping).
// Restates src/tridiagonal_cholesky.jl:65-82 (factor) and the intended semantics of :24-63 (solves) on the
// tile engine:  L_1 = chol(D_1);  C_i = B_i L_{i-1}^{-T};  L_i = chol(D_i - C_i C_i').
//
// Device layout: one slot per time block, slot i = [ L_i (b x b, ld) | C_i (b x b, ld) ], contiguous, so the
// task list of one block step is position independent and is replayed with a moving base pointer.
// Right-hand sides are kept "node-major" (nrhs x n, column-major) on the device so that every solve step is a
// right-sided GEMM/TRSM of the same kernels that factorise.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <functional>
#include <cmath>
#include <memory>

#include "common.hpp"
#include "handles.hpp"

using namespace gmrfb;

struct gmrfb_btd {
  gmrfb_ctx* ctx = nullptr;
  int64_t b = 0, N = 0;
  int ld = 0;
  int64_t slot = 0;  // doubles per block slot
  DevBuf<double> arena, dinv;  // dinv: scratch of inverted <=64x64 diagonal blocks
  DevPlan plan_first, plan_step;
  // look-ahead schedule (default for b >= 512): C_{i+1} = B_{i+1} L_i^{-T} is queued on a second stream and follows the
  // POTRF of block i panel by panel (one event per 64-column panel), so the two latency-bound chains overlap; the
  // inverses of L_i's diagonal blocks are kept in one of two alternating slot sets instead of being recomputed
  DevPlan la_potrf, la_trsm, la_syrk;
  DevBuf<double> la_dinv;  // 2 sets of ceil(b / 64) inverse-block slots
  std::vector<cudaEvent_t> la_ev[2];   // [set][j]: panel j of L_i is final (set = i & 1)
  std::vector<cudaEvent_t> la_cev[2];  // [set][j]: column block j of C_i is final
  cudaEvent_t la_done = nullptr;
  cudaStream_t la_stream = nullptr, la_stream2 = nullptr;  // TRSM lane, SYRK lane
  bool lookahead = false;
  int32_t status = GMRFB_ERR_STATE;
  int64_t fail_block = -1;
  bool factored = false;
  // Solves apply W_i = L_i^{-1} (computed once per factorisation, on the first solve, by recursive doubling from the
  // inverted 64x64 diagonal blocks): every block step of a sweep is then two GEMM launches whose cost is reading C_i
  // and W_i once, instead of a chain of b/64 dependent substitution steps.
  DevBuf<double> winv;  // N blocks of b x b (ld), strict upper triangles zero
  DevBuf<double> skws;  // workspace of the skinny kernels (arrival counters + K-slice partial sums), passed as Arenas::dinv
  bool winv_ready = false;
  DevPlan plan_winv;
  // solve plans are built per nrhs on demand
  int plan_nrhs = -1;
  DevPlan fwd_first, fwd_step, bwd_last, bwd_step;
  DevPlan sel_last, sel_step;
  bool sel_ready = false;
  // called by btd_run_factor after the launches of block i have been queued (time-sharded factor: queues W_i and the
  // spike step of block i on a second stream while the dependent chain of block i + 1 runs on the first)
  std::function<gmrfb_status(int64_t)> after_block;
};

namespace {

__global__ void k_btd_scatter_csc(int64_t ncols, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                  const double* __restrict__ nzval, int base, int64_t b, int64_t N, int ld, int64_t slot,
                                  double* __restrict__ arena) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  const int64_t bj = c / b, cj = c % b;
  if (bj >= N) return;
  for (int64_t p = colptr[c] - base; p < colptr[c + 1] - base; p++) {
    const int64_t r = rowval[p] - base;
    const int64_t bi = r / b, ri = r % b;
    if (bi >= N) continue;
    if (bi == bj)
      arena[bi * slot + cj * ld + ri] = nzval[p];
    else if (bi == bj + 1)
      arena[bi * slot + (int64_t)ld * b + cj * ld + ri] = nzval[p];
  }
}

// out (cols x rows, ldo) = in' (rows x cols, ldi)
__global__ void k_transpose_rect(const double* __restrict__ in, int64_t ldi, double* __restrict__ out, int64_t ldo,
                                 int64_t rows, int64_t cols) {
  __shared__ double t[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int64_t r = r0 + threadIdx.x, c = c0 + j;
    t[j][threadIdx.x] = (r < rows && c < cols) ? in[r + c * ldi] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int64_t c = c0 + threadIdx.x, r = r0 + j;
    if (r < rows && c < cols) out[c + r * ldo] = t[threadIdx.x][j];
  }
}

__global__ void k_btd_diag(const double* __restrict__ arena, int64_t slot, int ld, int64_t b, int64_t N,
                           double* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= b * N) return;
  int64_t i = k / b, j = k % b;
  out[k] = arena[i * slot + j * ld + j];
}

__global__ void k_diag_strided(const double* __restrict__ M, int ld, int64_t b, double* __restrict__ out) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < b) out[j] = M[j * ld + j];
}

// zero the strict upper triangle of a b x b block (the factorisation only ever touches the lower triangle of L_i; the
// solves stream whole blocks and rely on structural zeros above the diagonal)
__global__ void k_zero_upper(double* __restrict__ M, int ld, int64_t b) {
  const int64_t j = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < j) M[i + j * ld] = 0.0;
}

gmrfb_status upload_plan(gmrfb_ctx* ctx, DevPlan& P) {
  GMRFB_CU(ctx, P.tasks.upload(P.host.tasks, ctx->stream));
  P.ready = true;
  return GMRFB_OK;
}

// grow the inverse-block scratch to what plan `P` needs (stream-ordered: earlier launches have been queued before)
gmrfb_status ensure_dinv(gmrfb_btd* f, const Plan& P) {
  if ((int64_t)f->dinv.n >= P.dinv) return GMRFB_OK;
  gmrfb_ctx* ctx = f->ctx;
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  GMRFB_CU(ctx, f->dinv.alloc((size_t)P.dinv));
  return GMRFB_OK;
}

gmrfb_status btd_alloc(gmrfb_ctx* ctx, int64_t b, int64_t N, std::unique_ptr<gmrfb_btd>& f) {
  if (b <= 0 || N <= 0) return fail(ctx, GMRFB_ERR_INVALID, "btd: block size and block count must be positive");
  if (b > 60000) return fail(ctx, GMRFB_ERR_INVALID, "btd: block size too large");
  f.reset(new gmrfb_btd());
  f->ctx = ctx;
  f->b = b;
  f->N = N;
  f->ld = (int)((b + 1) & ~(int64_t)1);
  f->slot = 2 * (int64_t)f->ld * b;
  GMRFB_CU(ctx, f->arena.alloc((size_t)(f->slot * N)));
  // factor plans (offsets relative to the current slot; the previous slot is at -slot)
  {
    PlanBuilder B(f->plan_first.host);
    plan_potrf(B, f->plan_first.host, 0, 0, (int)b, f->ld, 0);
  }
  {
    Plan& P = f->plan_step.host;
    PlanBuilder B(P);
    const int64_t coff = (int64_t)f->ld * b;
    plan_trsm_rlt(B, P, 0, -f->slot, f->ld, 0, coff, (int)b, (int)b, f->ld);
    B.begin(LK_GEMM_NT);
    {
      Task t = make_task();
      t.a = coff;
      t.b = coff;
      t.c = 0;
      t.lda = t.ldb = t.ldc = f->ld;
      t.M = t.N = t.K = (int)b;
      t.alpha = -1.0;
      t.beta = 1.0;
      t.flags = TF_TRI;
      B.add(t, gemm_tiles((int)b, (int)b, true, GCFG_BIG));
      P.flops += (double)b * b * (b + 1);
    }
    B.end();
    plan_potrf(B, P, 0, 0, (int)b, f->ld, 0);
  }
  gmrfb_status rc = upload_plan(ctx, f->plan_first);
  if (rc != GMRFB_OK) return rc;
  if ((rc = ensure_dinv(f.get(), f->plan_first.host)) != GMRFB_OK) return rc;
  if ((rc = ensure_dinv(f.get(), f->plan_step.host)) != GMRFB_OK) return rc;
  if ((rc = upload_plan(ctx, f->plan_step)) != GMRFB_OK) return rc;
  {
    const char* e = getenv("GMRFB_BTD_LOOKAHEAD");
    f->lookahead = e ? (e[0] != '0') : (b >= 512 && N > 1);
  }
  if (f->lookahead) {
    const int64_t coff = (int64_t)f->ld * b;
    {
      PlanBuilder B(f->la_potrf.host);
      plan_potrf_events(B, f->la_potrf.host, 0, 0, (int)b, f->ld, 0);
    }
    {
      PlanBuilder B(f->la_trsm.host);
      plan_trsm_rlt_events(B, f->la_trsm.host, 0, -f->slot, f->ld, 0, coff, (int)b, (int)b, f->ld);
    }
    {
      // D_{i+1} -= C C' accumulated over K chunks of LA_SYRK_K columns of C, each launch waiting for the last column
      // block of its chunk: the rank-k updates follow the TRSM as its columns become final
      Plan& P = f->la_syrk.host;
      PlanBuilder B(P);
      const int LA_SYRK_K = 512;
      for (int c0 = 0; c0 < (int)b; c0 += LA_SYRK_K) {
        const int kc = std::min(LA_SYRK_K, (int)b - c0);
        B.begin(LK_GEMM_NT);
        B.set_wait((c0 + kc - 1) / NB);
        Task t = make_task();
        t.a = coff + (int64_t)c0 * f->ld;
        t.b = t.a;
        t.c = 0;
        t.lda = t.ldb = t.ldc = f->ld;
        t.M = t.N = (int)b;
        t.K = kc;
        t.alpha = -1.0;
        t.beta = 1.0;
        t.flags = TF_TRI;
        B.add(t, gemm_tiles((int)b, (int)b, true, GCFG_BIG));
        P.flops += (double)kc * b * (b + 1);
        B.end();
      }
    }
    if ((rc = upload_plan(ctx, f->la_potrf)) != GMRFB_OK) return rc;
    if ((rc = upload_plan(ctx, f->la_trsm)) != GMRFB_OK) return rc;
    if ((rc = upload_plan(ctx, f->la_syrk)) != GMRFB_OK) return rc;
    const int nblk = cdiv((int)b, NB);
    GMRFB_CU(ctx, f->la_dinv.alloc((size_t)2 * nblk * DINV_SLOT));
    for (int set = 0; set < 2; set++) {
      f->la_ev[set].resize(nblk);
      f->la_cev[set].resize(nblk);
      for (auto& ev : f->la_ev[set]) GMRFB_CU(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      for (auto& ev : f->la_cev[set]) GMRFB_CU(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    GMRFB_CU(ctx, cudaEventCreateWithFlags(&f->la_done, cudaEventDisableTiming));
    int prio_least = 0, prio_greatest = 0;
    GMRFB_CU(ctx, cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    GMRFB_CU(ctx, cudaStreamCreateWithPriority(&f->la_stream, cudaStreamNonBlocking, prio_least));
    GMRFB_CU(ctx, cudaStreamCreateWithPriority(&f->la_stream2, cudaStreamNonBlocking, prio_least));
  }
  return GMRFB_OK;
}

// one plan on `st`, honouring the per-launch events of a look-ahead schedule
gmrfb_status run_plan_events(gmrfb_ctx* ctx, const DevPlan& P, const Arenas& ar, const LaunchAux& aux, cudaStream_t st,
                             const std::vector<cudaEvent_t>* wait_evs, const std::vector<cudaEvent_t>* rec_evs) {
  for (const Launch& L : P.host.launches) {
    if (L.wait_ev >= 0 && wait_evs) GMRFB_CU(ctx, cudaStreamWaitEvent(st, (*wait_evs)[L.wait_ev], 0));
    cudaError_t e = run_launch(L, P.tasks.p, ar, aux, st);
    if (e != cudaSuccess) return fail(ctx, GMRFB_ERR_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(e));
    ctx->launches++;
    if (L.rec_ev >= 0 && rec_evs) GMRFB_CU(ctx, cudaEventRecord((*rec_evs)[L.rec_ev], st));
  }
  return GMRFB_OK;
}

gmrfb_status btd_run_factor(gmrfb_btd* f) {
  gmrfb_ctx* ctx = f->ctx;
  const int big = INT_MAX;
  GMRFB_CU(ctx, cudaMemcpyAsync(ctx->d_info, &big, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  LaunchAux aux;
  aux.d_info = ctx->d_info;
  if (f->lookahead && !ctx->profiling) {
    // stream 1 (ctx->stream, highest priority): POTRF_i, recording one event per finished 64-column panel of L_i
    // la_stream :  TRSM_{i+1} (C_{i+1} = B_{i+1} L_i^{-T}), every leaf waiting for the panel of L_i it reads and recording
    //              one event per finished column block of C_{i+1}
    // la_stream2:  SYRK_{i+1} (D_{i+1} -= C_{i+1} C_{i+1}') as rank-512 updates, each waiting for its columns of C_{i+1}
    // => the two latency-bound chains and the GPU-filling rank-k updates overlap; POTRF_{i+1} waits for the last update
    const int64_t setsz = (int64_t)cdiv((int)f->b, NB) * DINV_SLOT;
    GMRFB_CU(ctx, cudaEventRecord(f->la_done, ctx->stream));
    GMRFB_CU(ctx, cudaStreamWaitEvent(f->la_stream, f->la_done, 0));  // the input blocks are in place
    GMRFB_CU(ctx, cudaStreamWaitEvent(f->la_stream2, f->la_done, 0));
    for (int64_t i = 0; i < f->N; i++) {
      const int set = (int)(i & 1);
      Arenas ar{{f->arena.p + i * f->slot, nullptr, nullptr, nullptr}};
      gmrfb_status rc;
      if (i > 0) {
        ar.dinv = f->la_dinv.p + (int64_t)(1 - set) * setsz;  // inverses kept by POTRF_{i-1}
        rc = run_plan_events(ctx, f->la_trsm, ar, aux, f->la_stream, &f->la_ev[1 - set], &f->la_cev[set]);
        if (rc != GMRFB_OK) return rc;
        rc = run_plan_events(ctx, f->la_syrk, ar, aux, f->la_stream2, &f->la_cev[set], nullptr);
        if (rc != GMRFB_OK) return rc;
        GMRFB_CU(ctx, cudaEventRecord(f->la_done, f->la_stream2));
        GMRFB_CU(ctx, cudaStreamWaitEvent(ctx->stream, f->la_done, 0));
      }
      ar.dinv = f->la_dinv.p + (int64_t)set * setsz;
      rc = run_plan_events(ctx, f->la_potrf, ar, aux, ctx->stream, nullptr, &f->la_ev[set]);
      if (rc != GMRFB_OK) return rc;
      // time-sharded factor: W_i = L_i^-1 and the spike step of block i are queued on a further lane behind POTRF_i
      if (f->after_block && (rc = f->after_block(i)) != GMRFB_OK) return rc;
    }
  } else
  for (int64_t i = 0; i < f->N; i++) {
    Arenas ar{{f->arena.p + i * f->slot, nullptr, nullptr, nullptr}};
    ar.dinv = f->dinv.p;
    gmrfb_status rc = run_plan(ctx, i == 0 ? f->plan_first : f->plan_step, ar, aux);
    if (rc != GMRFB_OK) return rc;
    if (f->after_block && (rc = f->after_block(i)) != GMRFB_OK) return rc;
  }
  int info = 0;
  GMRFB_CU(ctx, cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  f->status = GMRFB_OK;
  f->fail_block = -1;
  f->factored = true;
  f->winv_ready = false;
  if (info != INT_MAX) {
    // locate the first block whose diagonal is poisoned
    std::vector<double> dg((size_t)(f->b * f->N));
    DevBuf<double> dd;
    GMRFB_CU(ctx, dd.alloc(dg.size()));
    k_btd_diag<<<(unsigned)((dg.size() + 255) / 256), 256, 0, ctx->stream>>>(f->arena.p, f->slot, f->ld, f->b, f->N, dd.p);
    GMRFB_CU(ctx, cudaGetLastError());
    GMRFB_CU(ctx, cudaMemcpyAsync(dg.data(), dd.p, dg.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < dg.size(); k++)
      if (!(dg[k] > 0.0)) {
        f->fail_block = (int64_t)(k / f->b);
        break;
      }
    f->status = GMRFB_ERR_NOT_SPD;
    f->factored = false;
    return fail(ctx, GMRFB_ERR_NOT_SPD, "btd: block " + std::to_string(f->fail_block) + " is not positive definite");
  }
  return GMRFB_OK;
}

}  // namespace

// copy the dense input blocks (host or device memory, column-major b x b) into the factor's slots
static gmrfb_status btd_load_dense(gmrfb_btd* f, const double* D, const double* Bsub) {
  gmrfb_ctx* ctx = f->ctx;
  const int64_t b = f->b;
  for (int64_t i = 0; i < f->N; i++) {
    GMRFB_CU(ctx, cudaMemcpy2DAsync(f->arena.p + i * f->slot, f->ld * sizeof(double), D + i * b * b, b * sizeof(double),
                                    b * sizeof(double), b, cudaMemcpyDefault, ctx->stream));
    if (i > 0)
      GMRFB_CU(ctx, cudaMemcpy2DAsync(f->arena.p + i * f->slot + (int64_t)f->ld * b, f->ld * sizeof(double),
                                      Bsub + (i - 1) * b * b, b * sizeof(double), b * sizeof(double), b,
                                      cudaMemcpyDefault, ctx->stream));
  }
  return GMRFB_OK;
}

extern "C" gmrfb_status gmrfb_btd_factor_dense(gmrfb_ctx* ctx, int64_t b, int64_t nblocks, const double* D,
                                               const double* Bsub, gmrfb_btd** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_factor_dense: ctx is NULL");
  if (!out || !D || (nblocks > 1 && !Bsub)) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_factor_dense: NULL argument");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_btd> f;
  gmrfb_status rc = btd_alloc(ctx, b, nblocks, f);
  if (rc != GMRFB_OK) return rc;
  if ((rc = btd_load_dense(f.get(), D, Bsub)) != GMRFB_OK) return rc;
  rc = btd_run_factor(f.get());
  *out = f.release();  // the handle is returned even when not SPD so that get_info can report the block
  return rc;
}
GMRFB_ABI_CATCH

// Block-tridiagonal precision of a constant-mesh implicit-Euler state-space model built directly in block form
// (ingredients of src/spdes/shallow_water.jl:198-228): with G = M + dt K and noise precision q,
//   D_1 = Q_0 + q M'M,   D_t = q (G'G + M'M) (1 < t < N),   D_N = q G'G,   B_t = -q G'M  (every sub-diagonal block),
// so four b-by-b blocks describe the whole N-block matrix: the arena is filled on the device from them, without the
// N-block host array of gmrfb_btd_factor_dense or the sparse -> dense gather of src/tridiagonal_cholesky.jl:73,76.
extern "C" gmrfb_status gmrfb_btd_factor_ssm(gmrfb_ctx* ctx, int64_t b, int64_t nblocks, const double* D_first,
                                             const double* D_mid, const double* D_last, const double* B_sub,
                                             gmrfb_btd** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_factor_ssm: ctx is NULL");
  if (!out || !D_first || (nblocks > 1 && (!D_last || !B_sub)) || (nblocks > 2 && !D_mid))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_factor_ssm: NULL argument");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_btd> f;
  gmrfb_status rc = btd_alloc(ctx, b, nblocks, f);
  if (rc != GMRFB_OK) return rc;
  // stage the distinct blocks once in device memory, then replicate device-to-device
  DevBuf<double> stage;
  GMRFB_CU(ctx, stage.alloc((size_t)(4 * b * b)));
  const double* src[4] = {D_first, D_mid, D_last, B_sub};
  for (int k = 0; k < 4; k++)
    if (src[k]) GMRFB_CU(ctx, cudaMemcpyAsync(stage.p + k * b * b, src[k], b * b * sizeof(double), cudaMemcpyDefault, ctx->stream));
  for (int64_t i = 0; i < nblocks; i++) {
    const int which = (i == 0) ? 0 : (i == nblocks - 1 ? 2 : 1);
    GMRFB_CU(ctx, cudaMemcpy2DAsync(f->arena.p + i * f->slot, f->ld * sizeof(double), stage.p + which * b * b,
                                    b * sizeof(double), b * sizeof(double), b, cudaMemcpyDeviceToDevice, ctx->stream));
    if (i > 0)
      GMRFB_CU(ctx, cudaMemcpy2DAsync(f->arena.p + i * f->slot + (int64_t)f->ld * b, f->ld * sizeof(double),
                                      stage.p + 3 * b * b, b * sizeof(double), b * sizeof(double), b,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
  }
  rc = btd_run_factor(f.get());  // synchronises the stream: `stage` may be released afterwards
  *out = f.release();
  return rc;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_btd_factor(gmrfb_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval,
                                         const double* nzval, int32_t base, int64_t nblocks, gmrfb_btd** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_factor: ctx is NULL");
  if (!out || !colptr || !rowval || !nzval || n <= 0 || nblocks <= 0 || (base != 0 && base != 1))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_factor: bad argument");
  *out = nullptr;
  const int64_t b = n / nblocks;  // src/tridiagonal_cholesky.jl:66 — the remainder rows are ignored
  if (b <= 0) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_factor: more blocks than rows");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_btd> f;
  gmrfb_status rc = btd_alloc(ctx, b, nblocks, f);
  if (rc != GMRFB_OK) return rc;
  const int64_t nnz = colptr[n] - base;
  DevBuf<int64_t> dcp, drv;
  DevBuf<double> dnz;
  GMRFB_CU(ctx, dcp.alloc((size_t)n + 1));
  GMRFB_CU(ctx, drv.alloc((size_t)std::max<int64_t>(nnz, 1)));
  GMRFB_CU(ctx, dnz.alloc((size_t)std::max<int64_t>(nnz, 1)));
  GMRFB_CU(ctx, cudaMemcpyAsync(dcp.p, colptr, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  GMRFB_CU(ctx, cudaMemcpyAsync(drv.p, rowval, nnz * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  GMRFB_CU(ctx, cudaMemcpyAsync(dnz.p, nzval, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GMRFB_CU(ctx, cudaMemsetAsync(f->arena.p, 0, f->arena.n * sizeof(double), ctx->stream));
  k_btd_scatter_csc<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, dcp.p, drv.p, dnz.p, base, b, nblocks,
                                                                          f->ld, f->slot, f->arena.p);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches++;
  rc = btd_run_factor(f.get());
  *out = f.release();
  return rc;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_btd_destroy(gmrfb_btd* f) try {
  if (!f) return GMRFB_OK;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  for (cudaStream_t st : {f->la_stream, f->la_stream2})
    if (st) {
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
    }
  for (int set = 0; set < 2; set++) {
    for (auto ev : f->la_ev[set]) cudaEventDestroy(ev);
    for (auto ev : f->la_cev[set]) cudaEventDestroy(ev);
  }
  if (f->la_done) cudaEventDestroy(f->la_done);
  delete f;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_btd_get_info(gmrfb_btd* f, gmrfb_btd_info* info) try {
  if (!f || !info) return fail(f ? f->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_get_info: NULL argument");
  info->b = f->b;
  info->nblocks = f->N;
  info->status = f->status;
  info->fail_block = f->fail_block;
  const double b3 = (double)f->b * f->b * f->b;
  info->flops = (double)(f->N - 1) * (7.0 / 3.0) * b3 + b3 / 3.0;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_btd_get_block(gmrfb_btd* f, int64_t i, int32_t which, double* out, int64_t ldo) try {
  if (!f || !out) return fail(f ? f->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_get_block: NULL argument");
  gmrfb_ctx* ctx = f->ctx;
  if (!f->factored) return fail(ctx, GMRFB_ERR_STATE, "gmrfb_btd_get_block: no successful factorisation");
  const int64_t b = f->b;
  if (ldo < b) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_get_block: ldo too small");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const double* src;
  if (which == GMRFB_BTD_BLOCK_L) {
    if (i < 0 || i >= f->N) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_get_block: block index out of range");
    src = f->arena.p + i * f->slot;
  } else if (which == GMRFB_BTD_BLOCK_C) {
    if (i < 0 || i >= f->N - 1) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_get_block: block index out of range");
    src = f->arena.p + (i + 1) * f->slot + (int64_t)f->ld * b;
  } else {
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_get_block: bad selector");
  }
  GMRFB_CU(ctx, cudaMemcpy2DAsync(out, ldo * sizeof(double), src, f->ld * sizeof(double), b * sizeof(double), b,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (which == GMRFB_BTD_BLOCK_L)
    for (int64_t c = 1; c < b; c++)
      for (int64_t r = 0; r < c; r++) out[r + c * ldo] = 0.0;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_btd_logdet(gmrfb_btd* f, double* logdet) try {
  if (!f || !logdet) return fail(f ? f->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_logdet: NULL argument");
  gmrfb_ctx* ctx = f->ctx;
  if (!f->factored) return fail(ctx, GMRFB_ERR_STATE, "gmrfb_btd_logdet: no successful factorisation");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::vector<double> dg((size_t)(f->b * f->N));
  DevBuf<double> dd;
  GMRFB_CU(ctx, dd.alloc(dg.size()));
  k_btd_diag<<<(unsigned)((dg.size() + 255) / 256), 256, 0, ctx->stream>>>(f->arena.p, f->slot, f->ld, f->b, f->N, dd.p);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches++;
  GMRFB_CU(ctx, cudaMemcpyAsync(dg.data(), dd.p, dg.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  double s = 0;
  for (double v : dg) s += std::log(v);
  *logdet = 2.0 * s;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// ---------------------------------------------------------------------------------------------- solve ----
// Device RHS layout: node-major, nrhs x (b*N) with leading dimension ldr; block i occupies columns [i*b, (i+1)*b).
// Two such buffers are used alternately (a GEMM cannot run in place):
//   forward   U_i = X_i - Y_{i-1} C_i'          (in X),   Y_i = U_i L_i^{-T}   (into Y; X_i is consumed as scratch)
//   backward  V_i = Y_i - X_{i+1} C_{i+1}       (in Y),   X_i = V_i L_i^{-1}   (into X; Y_i is consumed as scratch)
// Arenas (all moving with the block): 0 = factor slot i, 1 = X block i, 2 = Y block i, 3 = W_i.
// allocate and clear W, build the recursive-doubling plan (stream-ordered on ctx->stream)
static gmrfb_status btd_prepare_winv(gmrfb_btd* f) {
  gmrfb_ctx* ctx = f->ctx;
  const int b = (int)f->b, ld = f->ld;
  const int64_t bs = (int64_t)ld * b;
  if (f->winv.n < (size_t)(bs * f->N)) GMRFB_CU(ctx, f->winv.alloc((size_t)(bs * f->N)));
  GMRFB_CU(ctx, cudaMemsetAsync(f->winv.p, 0, (size_t)(bs * f->N) * sizeof(double), ctx->stream));
  if (!f->plan_winv.ready) {
    PlanBuilder B(f->plan_winv.host);
    plan_trtri(B, f->plan_winv.host, 0, 0, ld, 1, 0, ld, 2, 0, b);
    gmrfb_status rc = upload_plan(ctx, f->plan_winv);
    if (rc != GMRFB_OK) return rc;
    if ((rc = ensure_dinv(f, f->plan_winv.host)) != GMRFB_OK) return rc;
  }
  return GMRFB_OK;
}
// W_i = L_i^{-1} of block i on ctx->stream (scratch: b x b, ld)
static gmrfb_status btd_winv_block(gmrfb_btd* f, int64_t i, double* scratch) {
  gmrfb_ctx* ctx = f->ctx;
  const int b = (int)f->b, ld = f->ld;
  const int64_t bs = (int64_t)ld * b;
  if (b > 1) {
    k_zero_upper<<<dim3((unsigned)((b + 255) / 256), (unsigned)b), 256, 0, ctx->stream>>>(f->arena.p + i * f->slot, ld, b);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches++;
  }
  LaunchAux aux;
  aux.d_info = ctx->d_info;
  Arenas ar{{f->arena.p + i * f->slot, f->winv.p + i * bs, scratch, nullptr}};
  ar.dinv = f->dinv.p;
  return run_plan(ctx, f->plan_winv, ar, aux);
}
static gmrfb_status btd_ensure_winv(gmrfb_btd* f) {
  if (f->winv_ready) return GMRFB_OK;
  gmrfb_ctx* ctx = f->ctx;
  gmrfb_status rc = btd_prepare_winv(f);
  if (rc != GMRFB_OK) return rc;
  DevBuf<double> scratch;
  GMRFB_CU(ctx, scratch.alloc((size_t)((int64_t)f->ld * f->b)));
  for (int64_t i = 0; i < f->N; i++)
    if ((rc = btd_winv_block(f, i, scratch.p)) != GMRFB_OK) return rc;
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));  // scratch is released on return
  f->winv_ready = true;
  return GMRFB_OK;
}

static void btd_add_solve_gemm(PlanBuilder& B, Plan& P, int kind, int aa, int64_t a, int ab, int64_t b, int ldb, int ac,
                               int nrhs, int n, int ldr, double alpha, double beta, int32_t extra) {
  // few right-hand sides: bandwidth-bound streaming kernels; otherwise the tile engine
  const bool skinny = nrhs <= SKINNY_MAX_ROWS;
  B.begin(skinny ? (kind == LK_GEMM_NT ? LK_SKINNY_NT : LK_SKINNY_NN) : kind);
  Task t = make_task();
  t.a = a;
  t.lda = ldr;
  t.b = b;
  t.ldb = ldb;
  t.c = 0;
  t.ldc = ldr;
  t.M = nrhs;
  t.N = n;
  t.K = n;
  t.alpha = alpha;
  t.beta = beta;
  t.flags = (aa << TF_A_SHIFT) | (ab << TF_B_SHIFT) | (ac << TF_C_SHIFT) | extra;
  if (!skinny) {
    B.add(t, gemm_tiles(nrhs, n, false, GCFG_BIG));
  } else if (kind == LK_GEMM_NT) {
    t.aux0 = skinny_slices(n, n);
    B.add(t, cdiv(n, SKINNY_ROWS) * t.aux0);
    B.set_cfg(skinny_nr(nrhs));
    P.dinv = std::max(P.dinv, skinny_workspace(nrhs, n, n));
  } else {
    B.add(t, cdiv(n, 8));
    B.set_cfg(skinny_nr(nrhs));
  }
  // algorithmic bytes: the matrix operand once (a triangular W: its lower half)
  B.add_bytes(8.0 * n * (double)n * ((extra & (TF_BUPP | TF_BLOW)) ? 0.5 : 1.0));
  B.end();
}

static gmrfb_status btd_build_solve_plans(gmrfb_btd* f, int nrhs, int ldr) {
  if (f->plan_nrhs == nrhs) return GMRFB_OK;
  gmrfb_ctx* ctx = f->ctx;
  const int b = (int)f->b, ld = f->ld;
  const int64_t coff = (int64_t)ld * b;
  const int64_t xstep = (int64_t)ldr * b;  // doubles between consecutive blocks of a right-hand-side buffer
  for (DevPlan* dp : {&f->fwd_first, &f->fwd_step, &f->bwd_last, &f->bwd_step}) {
    dp->host = Plan();
    dp->tasks.release();
    dp->ready = false;
  }
  // Applying W_i = L_i^{-1} alone loses accuracy on ill-conditioned blocks (measured: residual 7e-6 against 2e-7 for
  // substitution on the Burgers posterior of config 2).  One step of iterative refinement against L_i itself,
  //     y = W t;   r = t - L y;   y += W r,
  // restores the accuracy of a substitution at the price of two more streaming products per block.
  auto fwd_apply = [&](PlanBuilder& B, Plan& BP) {
    btd_add_solve_gemm(B, BP, LK_GEMM_NT, 1, 0, 3, 0, ld, 2, nrhs, b, ldr, 1.0, 0.0, TF_BUPP);   // Y_i = X_i W_i'
    btd_add_solve_gemm(B, BP, LK_GEMM_NT, 2, 0, 0, 0, ld, 1, nrhs, b, ldr, -1.0, 1.0, TF_BUPP);  // X_i -= Y_i L_i'
    btd_add_solve_gemm(B, BP, LK_GEMM_NT, 1, 0, 3, 0, ld, 2, nrhs, b, ldr, 1.0, 1.0, TF_BUPP);   // Y_i += X_i W_i'
  };
  auto bwd_apply = [&](PlanBuilder& B, Plan& BP) {
    btd_add_solve_gemm(B, BP, LK_GEMM_NN, 2, 0, 3, 0, ld, 1, nrhs, b, ldr, 1.0, 0.0, TF_BLOW);   // X_i = Y_i W_i
    btd_add_solve_gemm(B, BP, LK_GEMM_NN, 1, 0, 0, 0, ld, 2, nrhs, b, ldr, -1.0, 1.0, TF_BLOW);  // Y_i -= X_i L_i
    btd_add_solve_gemm(B, BP, LK_GEMM_NN, 2, 0, 3, 0, ld, 1, nrhs, b, ldr, 1.0, 1.0, TF_BLOW);   // X_i += Y_i W_i
  };
  {  // Y_1 = X_1 L_1^{-T}
    Plan& BP = f->fwd_first.host;
    PlanBuilder B(BP);
    fwd_apply(B, BP);
  }
  {  // X_i -= Y_{i-1} C_i';  Y_i = X_i L_i^{-T}
    Plan& BP = f->fwd_step.host;
    PlanBuilder B(BP);
    btd_add_solve_gemm(B, BP, LK_GEMM_NT, 2, -xstep, 0, coff, ld, 1, nrhs, b, ldr, -1.0, 1.0, 0);
    fwd_apply(B, BP);
  }
  {  // X_N = Y_N L_N^{-1}
    Plan& BP = f->bwd_last.host;
    PlanBuilder B(BP);
    bwd_apply(B, BP);
  }
  {  // Y_i -= X_{i+1} C_{i+1} (C_{i+1} sits in the next slot);  X_i = Y_i L_i^{-1}
    Plan& BP = f->bwd_step.host;
    PlanBuilder B(BP);
    btd_add_solve_gemm(B, BP, LK_GEMM_NN, 1, xstep, 0, f->slot + coff, ld, 2, nrhs, b, ldr, -1.0, 1.0, 0);
    bwd_apply(B, BP);
  }
  gmrfb_status rc;
  for (DevPlan* dp : {&f->fwd_first, &f->fwd_step, &f->bwd_last, &f->bwd_step}) {
    if ((rc = upload_plan(ctx, *dp)) != GMRFB_OK) return rc;
    if ((int64_t)f->skws.n < dp->host.dinv) {
      GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
      GMRFB_CU(ctx, f->skws.alloc((size_t)dp->host.dinv));
      // the arrival counters at the head of the workspace must start at zero (the kernels leave them at zero)
      GMRFB_CU(ctx, cudaMemsetAsync(f->skws.p, 0, (size_t)dp->host.dinv * sizeof(double), ctx->stream));
    }
  }
  f->plan_nrhs = nrhs;
  return GMRFB_OK;
}

// Run the forward and/or backward block sweeps of factor `f` on a device node-major buffer dXt (nrhs x b*N), in place.
static gmrfb_status btd_sweep(gmrfb_btd* f, bool fwd, bool bwd, double* dXt, int ldr, int nrhs) {
  gmrfb_ctx* ctx = f->ctx;
  gmrfb_status rc = btd_ensure_winv(f);
  if (rc != GMRFB_OK) return rc;
  if ((rc = btd_build_solve_plans(f, nrhs, ldr)) != GMRFB_OK) return rc;
  LaunchAux aux;
  aux.d_info = ctx->d_info;
  const int64_t xstep = (int64_t)ldr * f->b, bs = (int64_t)f->ld * f->b;
  const size_t bytes = (size_t)(xstep * f->N) * sizeof(double);
  DevBuf<double> Y;
  GMRFB_CU(ctx, Y.alloc((size_t)(xstep * f->N), ctx->stream));
  if (!fwd) GMRFB_CU(ctx, cudaMemcpyAsync(Y.p, dXt, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  auto arenas = [&](int64_t i) {
    Arenas ar{{f->arena.p + i * f->slot, dXt + i * xstep, Y.p + i * xstep, f->winv.p + i * bs}};
    ar.dinv = f->skws.p;
    return ar;
  };
  if (fwd) {
    for (int64_t i = 0; i < f->N; i++) {
      rc = run_plan(ctx, i == 0 ? f->fwd_first : f->fwd_step, arenas(i), aux);
      if (rc != GMRFB_OK) return rc;
    }
    if (!bwd) GMRFB_CU(ctx, cudaMemcpyAsync(dXt, Y.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (bwd) {
    for (int64_t i = f->N - 1; i >= 0; i--) {
      rc = run_plan(ctx, i == f->N - 1 ? f->bwd_last : f->bwd_step, arenas(i), aux);
      if (rc != GMRFB_OK) return rc;
    }
  }
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));  // Y is released on return
  return GMRFB_OK;
}

static gmrfb_status transpose_dev(gmrfb_ctx* ctx, const double* in, int64_t ldi, double* out, int64_t ldo, int64_t rows,
                                  int64_t cols) {
  dim3 tb(32, 8);
  k_transpose_rect<<<dim3((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32)), tb, 0, ctx->stream>>>(in, ldi, out, ldo,
                                                                                                          rows, cols);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches++;
  return GMRFB_OK;
}

extern "C" gmrfb_status gmrfb_btd_solve(gmrfb_btd* f, int32_t mode, double* X, int64_t ldx, int64_t nrhs) try {
  if (!f || !X) return fail(f ? f->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_solve: NULL argument");
  gmrfb_ctx* ctx = f->ctx;
  if (!f->factored) return fail(ctx, GMRFB_ERR_STATE, "gmrfb_btd_solve: no successful factorisation");
  const int64_t n = f->b * f->N;
  if (ldx < n || nrhs <= 0 || nrhs > 4096) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_solve: bad ldx/nrhs");
  if (mode < GMRFB_BTD_SOLVE_A || mode > GMRFB_BTD_SOLVE_BWD) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_solve: bad mode");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int ldr = (int)((nrhs + 1) & ~(int64_t)1);
  DevBuf<double> dX, dXt;
  GMRFB_CU(ctx, dX.alloc((size_t)(n * nrhs), ctx->stream));
  GMRFB_CU(ctx, dXt.alloc((size_t)((int64_t)ldr * n), ctx->stream));
  GMRFB_CU(ctx, cudaMemcpy2DAsync(dX.p, n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double), nrhs,
                                  cudaMemcpyHostToDevice, ctx->stream));
  gmrfb_status rc = transpose_dev(ctx, dX.p, n, dXt.p, ldr, n, nrhs);
  if (rc != GMRFB_OK) return rc;
  rc = btd_sweep(f, mode != GMRFB_BTD_SOLVE_BWD, mode != GMRFB_BTD_SOLVE_FWD, dXt.p, ldr, (int)nrhs);
  if (rc != GMRFB_OK) return rc;
  rc = transpose_dev(ctx, dXt.p, ldr, dX.p, n, nrhs, n);
  if (rc != GMRFB_OK) return rc;
  GMRFB_CU(ctx, cudaMemcpy2DAsync(X, ldx * sizeof(double), dX.p, n * sizeof(double), n * sizeof(double), nrhs,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// ------------------------------------------------------------------------------- selected inversion ----
// S_N = L_N^{-T} L_N^{-1};  S_i = L_i^{-T} (I + C_{i+1}' S_{i+1} C_{i+1}) L_i^{-1}.
// Arena 0 = factor slot i (C_{i+1} in the next slot), 1 = S_{i+1} (b x b, ld), 2 = T scratch, 3 = S_i (output).
extern "C" gmrfb_status gmrfb_btd_selinv_diag(gmrfb_btd* f, double* var_out) try {
  if (!f || !var_out) return fail(f ? f->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_selinv_diag: NULL argument");
  gmrfb_ctx* ctx = f->ctx;
  if (!f->factored) return fail(ctx, GMRFB_ERR_STATE, "gmrfb_btd_selinv_diag: no successful factorisation");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int b = (int)f->b, ld = f->ld;
  const int64_t coff = (int64_t)ld * b;
  if (!f->sel_ready) {
    auto tail = [&](Plan& P, PlanBuilder& B) {
      // arena 3 holds H; S = (H L^{-1})' L^{-1}
      plan_trsm_rln(B, P, 0, 0, ld, 3, 0, b, b, ld, false);
      B.begin(LK_TRANSPOSE);
      Task t = make_task();
      t.c = 0;
      t.ldc = ld;
      t.M = b;
      t.flags = 3 << TF_C_SHIFT;
      int nt = cdiv(b, 32);
      B.add(t, nt * (nt + 1) / 2);
      B.end();
      plan_trsm_rln(B, P, 0, 0, ld, 3, 0, b, b, ld, false);
    };
    auto ident = [&](PlanBuilder& B) {
      B.begin(LK_SET_IDENTITY);
      Task t = make_task();
      t.c = 0;
      t.ldc = ld;
      t.M = t.N = b;
      t.flags = 3 << TF_C_SHIFT;
      B.add(t, cdiv(b, 64) * cdiv(b, 64));
      B.end();
    };
    {
      Plan& P = f->sel_last.host;
      PlanBuilder B(P);
      ident(B);
      tail(P, B);
    }
    {
      Plan& P = f->sel_step.host;
      PlanBuilder B(P);
      // T = -S_{i+1} C_{i+1}  (S symmetric, both triangles valid)
      B.begin(LK_GEMM_NN);
      Task t = make_task();
      t.a = 0;
      t.lda = ld;
      t.b = f->slot + coff;
      t.ldb = ld;
      t.c = 0;
      t.ldc = ld;
      t.M = t.N = t.K = b;
      t.alpha = -1.0;
      t.beta = 0.0;
      t.flags = (1 << TF_A_SHIFT) | (0 << TF_B_SHIFT) | (2 << TF_C_SHIFT);
      B.add(t, gemm_tiles(b, b, false, GCFG_BIG));
      B.end();
      ident(B);
      // H = I - C_{i+1}' T
      B.begin(LK_GEMM_TN);
      Task u = make_task();
      u.a = f->slot + coff;
      u.lda = ld;
      u.b = 0;
      u.ldb = ld;
      u.c = 0;
      u.ldc = ld;
      u.M = u.N = u.K = b;
      u.alpha = -1.0;
      u.beta = 1.0;
      u.flags = (0 << TF_A_SHIFT) | (2 << TF_B_SHIFT) | (3 << TF_C_SHIFT);
      B.add(u, gemm_tiles(b, b, false, GCFG_BIG));
      B.end();
      tail(P, B);
    }
    gmrfb_status rc;
    if ((rc = upload_plan(ctx, f->sel_last)) != GMRFB_OK) return rc;
    if ((rc = upload_plan(ctx, f->sel_step)) != GMRFB_OK) return rc;
    if ((rc = ensure_dinv(f, f->sel_last.host)) != GMRFB_OK) return rc;
    if ((rc = ensure_dinv(f, f->sel_step.host)) != GMRFB_OK) return rc;
    f->sel_ready = true;
  }
  DevBuf<double> S0, S1, T, dvar;
  GMRFB_CU(ctx, S0.alloc((size_t)coff));
  GMRFB_CU(ctx, S1.alloc((size_t)coff));
  GMRFB_CU(ctx, T.alloc((size_t)coff));
  GMRFB_CU(ctx, dvar.alloc((size_t)(f->b * f->N)));
  LaunchAux aux;
  aux.d_info = ctx->d_info;
  double* prev = S0.p;
  double* cur = S1.p;
  for (int64_t i = f->N - 1; i >= 0; i--) {
    Arenas ar{{f->arena.p + i * f->slot, prev, T.p, cur}};
    ar.dinv = f->dinv.p;
    gmrfb_status rc = run_plan(ctx, i == f->N - 1 ? f->sel_last : f->sel_step, ar, aux);
    if (rc != GMRFB_OK) return rc;
    k_diag_strided<<<(unsigned)((b + 255) / 256), 256, 0, ctx->stream>>>(cur, ld, b, dvar.p + i * f->b);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches++;
    std::swap(prev, cur);
  }
  GMRFB_CU(ctx, cudaMemcpyAsync(var_out, dvar.p, f->b * f->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// ======================================================================= time-sharded block tridiagonal ====
// Rank r owns a contiguous slab of blocks.  Ranks 0..P-2 treat their last block as a *separator*; all other blocks
// are interior.  With separators ordered last, the Cholesky factor is
//      [ L_II        ]      L_II  = block-diagonal of the per-rank interior block-tridiagonal factors,
//      [ L_SI   L_SS ]      L_SI  = A_SI L_II^{-T}: per rank one dense block V (own separator x last interior block)
// and a *spike* W (previous rank's separator x every interior block), and L_SS L_SS' = S_hat, the (P-1)-block
// tridiagonal Schur complement.  The only inter-rank exchange is an all-gather of three b x b blocks per rank
// (factor) and two b x nrhs panels per rank (solve); it is performed by the host (torch.distributed / NCCL).
struct gmrfb_btd_dist {
  gmrfb_ctx* ctx = nullptr;
  int rank = 0, P = 1;
  int64_t b = 0, nloc = 0, ni = 0;  // ni = interior blocks
  bool has_sep = false, has_spike = false;
  int ld = 0;
  gmrfb_btd* interior = nullptr;
  gmrfb_btd* reduced = nullptr;
  DevBuf<double> W;      // ni blocks of b x ld (spike), block i at W + i*ld*b
  DevBuf<double> V;      // b x ld
  DevBuf<double> iface;  // 3 blocks of b x b (ld = b): [Dsep - VV', Q, R]
  DevBuf<double> Xt, Sx; // solve state: local node-major rhs; separator solutions
  int solve_nrhs = 0, solve_ldr = 0;
  bool reduced_ready = false;
  ~gmrfb_btd_dist() {  // also runs on the error paths of gmrfb_btd_dist_create
    if (interior) gmrfb_btd_destroy(interior);
    if (reduced) gmrfb_btd_destroy(reduced);
  }
};

namespace {

// dst (b x b, ldd) = alpha * x (ldx) + beta * y (ldy)  (y may be NULL)
__global__ void k_block_axpby(int64_t b, double alpha, const double* __restrict__ x, int64_t ldx, double beta,
                              const double* __restrict__ y, int64_t ldy, double* __restrict__ dst, int64_t ldd) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= b * b) return;
  int64_t i = e % b, j = e / b;
  double v = alpha * x[i + j * ldx];
  if (y) v += beta * y[i + j * ldy];
  dst[i + j * ldd] = v;
}

gmrfb_status block_axpby(gmrfb_ctx* ctx, int64_t b, double alpha, const double* x, int64_t ldx, double beta,
                         const double* y, int64_t ldy, double* dst, int64_t ldd) {
  k_block_axpby<<<(unsigned)((b * b + 255) / 256), 256, 0, ctx->stream>>>(b, alpha, x, ldx, beta, y, ldy, dst, ldd);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches++;
  return GMRFB_OK;
}

// one-off GEMM through the tile engine: C = beta C + alpha op(A) op(B), pointers given directly
gmrfb_status gemm_once(gmrfb_ctx* ctx, int kind, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
                       int M, int N, int K, bool tri, double alpha, double beta) {
  Task t = make_task();
  t.a = t.b = t.c = 0;
  t.lda = lda;
  t.ldb = ldb;
  t.ldc = ldc;
  t.M = M;
  t.N = N;
  t.K = K;
  t.alpha = alpha;
  t.beta = beta;
  t.tile0 = 0;
  t.flags = (0 << TF_A_SHIFT) | (1 << TF_B_SHIFT) | (2 << TF_C_SHIFT) | (tri ? TF_TRI : 0);
  DevBuf<Task> dt;
  std::vector<Task> ht{t};
  GMRFB_CU(ctx, dt.upload(ht, ctx->stream));
  Launch L{};
  L.kind = kind;
  L.task0 = 0;
  L.ntasks = 1;
  L.cfg = choose_gemm_cfg(gemm_tiles(M, N, tri, GCFG_BIG), 1);
  L.grid = gemm_tiles(M, N, tri, L.cfg);
  Arenas ar{{const_cast<double*>(A), const_cast<double*>(B), C, nullptr}};
  LaunchAux aux;
  GMRFB_CU(ctx, run_launch(L, dt.p, ar, aux, ctx->stream));
  ctx->launches++;
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));  // dt is released on return
  return GMRFB_OK;
}

}  // namespace

extern "C" gmrfb_status gmrfb_btd_dist_create(gmrfb_ctx* ctx, int32_t rank, int32_t nranks, int64_t b, int64_t nloc,
                                              const double* D_local, const double* B_local, gmrfb_btd_dist** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_dist_create: ctx is NULL");
  if (!out || !D_local || !B_local || rank < 0 || nranks < 1 || rank >= nranks || b <= 0)
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_dist_create: bad argument");
  const bool has_sep = rank < nranks - 1;
  if (nloc < (has_sep ? 2 : 1))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_dist_create: every rank but the last needs at least 2 blocks");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_btd_dist> h(new gmrfb_btd_dist());
  h->ctx = ctx;
  h->rank = rank;
  h->P = nranks;
  h->b = b;
  h->nloc = nloc;
  h->has_sep = has_sep;
  h->has_spike = rank > 0;
  h->ni = has_sep ? nloc - 1 : nloc;
  // interior factor (blocks 0..ni-1); couplings inside the interior are B_local[:,:,1..ni-1]
  const bool lanes = (h->has_spike || h->has_sep) && !getenv("GMRFB_BTD_DIST_SERIAL");
  gmrfb_status rc;
  if (!lanes) {
    rc = gmrfb_btd_factor_dense(ctx, b, h->ni, D_local, h->ni > 1 ? B_local + b * b : nullptr, &h->interior);
    if (rc != GMRFB_OK) return rc;
  } else {
    std::unique_ptr<gmrfb_btd> f;
    if ((rc = btd_alloc(ctx, b, h->ni, f)) != GMRFB_OK) return rc;
    h->interior = f.release();
    if ((rc = btd_load_dense(h->interior, D_local, h->ni > 1 ? B_local + b * b : nullptr)) != GMRFB_OK) return rc;
  }
  gmrfb_btd* F = h->interior;
  const int ld = F->ld;
  h->ld = ld;
  const int64_t bs = (int64_t)ld * b;
  const int ib = (int)b;
  GMRFB_CU(ctx, h->iface.alloc((size_t)(3 * b * b)));
  GMRFB_CU(ctx, cudaMemsetAsync(h->iface.p, 0, 3 * b * b * sizeof(double), ctx->stream));
  DevBuf<double> tmp;
  GMRFB_CU(ctx, tmp.alloc((size_t)bs));
  // The couplings with the neighbouring separators are eliminated with W_i = L_i^{-1} of the interior blocks (the same
  // inverses the solves use), so the whole spike recurrence is a replayed static plan of GEMMs without host round trips.
  // Two lanes: the factor of block i + 1 (a dependent chain of 64-column steps that leaves most SMs idle) runs on
  // ctx->stream while W_i and the spike step of block i (large GEMMs) run on ctx->stream2; the only ordering between
  // the lanes is "block i factored" (one event).  GMRFB_BTD_DIST_SERIAL=1 restores the one-stream order (A/B runs).
  DevBuf<double> work, wscratch;
  DevPlan first, step;
  if (lanes) {
    if ((rc = btd_prepare_winv(F)) != GMRFB_OK) return rc;
    GMRFB_CU(ctx, wscratch.alloc((size_t)bs));
    if (!ctx->stream2) {
      int prio_least = 0, prio_greatest = 0;
      GMRFB_CU(ctx, cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
      GMRFB_CU(ctx, cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, prio_least));
    }
    if (!ctx->ev_lane) GMRFB_CU(ctx, cudaEventCreateWithFlags(&ctx->ev_lane, cudaEventDisableTiming));
  } else if (h->has_spike || h->has_sep) {
    rc = btd_ensure_winv(F);
    if (rc != GMRFB_OK) return rc;
  }
  if (h->has_spike) {
    GMRFB_CU(ctx, h->W.alloc((size_t)(bs * h->ni)));
    // scratch arena: [ T (b x b, ld) | E_l (b x b, ld) | Q (b x b, ld) ]
    GMRFB_CU(ctx, work.alloc((size_t)(3 * bs)));
    GMRFB_CU(ctx, cudaMemsetAsync(work.p + 2 * bs, 0, (size_t)bs * sizeof(double), ctx->stream));
    // E_l = B_local[:,:,0] (rows: first interior block, cols: previous separator)
    GMRFB_CU(ctx, cudaMemcpy2DAsync(work.p + bs, ld * sizeof(double), B_local, b * sizeof(double), b * sizeof(double), b,
                                    cudaMemcpyDefault, ctx->stream));
    // arenas: 0 = factor slot i, 1 = spike block i (block i-1 at -bs), 2 = work, 3 = W_i = L_i^{-1}
    auto gemm = [&](PlanBuilder& B, int kind, int aa, int64_t a, int ab, int64_t bo, int ac, int64_t c, bool tri, double alpha,
                    double beta, int32_t extra) {
      B.begin(kind);
      Task t = make_task();
      t.a = a;
      t.b = bo;
      t.c = c;
      t.lda = t.ldb = t.ldc = ld;
      t.M = t.N = t.K = ib;
      t.alpha = alpha;
      t.beta = beta;
      t.flags = (aa << TF_A_SHIFT) | (ab << TF_B_SHIFT) | (ac << TF_C_SHIFT) | (tri ? TF_TRI : 0) | extra;
      B.add(t, gemm_tiles(ib, ib, tri, GCFG_BIG));
      B.end();
    };
    {  // S_1 = E_l' W_1';  Q = S_1 S_1'
      PlanBuilder B(first.host);
      gemm(B, LK_GEMM_TT, 2, bs, 3, 0, 1, 0, false, 1.0, 0.0, TF_BUPP);
      gemm(B, LK_GEMM_NT, 1, 0, 1, 0, 2, 2 * bs, true, 1.0, 1.0, 0);
    }
    {  // T = S_{i-1} C_i';  S_i = -T W_i';  Q += S_i S_i'
      PlanBuilder B(step.host);
      gemm(B, LK_GEMM_NT, 1, -bs, 0, bs, 2, 0, false, 1.0, 0.0, 0);
      gemm(B, LK_GEMM_NT, 2, 0, 3, 0, 1, 0, false, -1.0, 0.0, TF_BUPP);
      gemm(B, LK_GEMM_NT, 1, 0, 1, 0, 2, 2 * bs, true, 1.0, 1.0, 0);
    }
    if ((rc = upload_plan(ctx, first)) != GMRFB_OK) return rc;
    if ((rc = upload_plan(ctx, step)) != GMRFB_OK) return rc;
  }
  LaunchAux aux;
  aux.d_info = ctx->d_info;
  auto spike_step = [&](int64_t i) -> gmrfb_status {
    Arenas ar{{F->arena.p + i * F->slot, h->W.p + i * bs, work.p, F->winv.p + i * bs}};
    return run_plan(ctx, i == 0 ? first : step, ar, aux);
  };
  if (lanes) {
    // everything queued so far on ctx->stream (input copies, memsets, plan uploads) precedes the first event
    F->after_block = [&](int64_t i) -> gmrfb_status {
      GMRFB_CU(ctx, cudaEventRecord(ctx->ev_lane, ctx->stream));
      std::swap(ctx->stream, ctx->stream2);  // queue on the second lane
      gmrfb_status r = GMRFB_OK;
      cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->ev_lane, 0);
      if (e != cudaSuccess) r = fail(ctx, GMRFB_ERR_CUDA, std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(e));
      if (r == GMRFB_OK) r = btd_winv_block(F, i, wscratch.p);
      if (r == GMRFB_OK && h->has_spike) r = spike_step(i);
      std::swap(ctx->stream, ctx->stream2);
      return r;
    };
    rc = btd_run_factor(F);  // synchronises ctx->stream and reports a non-SPD block
    F->after_block = nullptr;
    GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream2));
    if (rc != GMRFB_OK) return rc;
    F->winv_ready = true;
  } else if (h->has_spike) {
    for (int64_t i = 0; i < h->ni; i++)
      if ((rc = spike_step(i)) != GMRFB_OK) return rc;
  }
  if (h->has_spike) {
    // Q (lower triangle) -> iface block 1 (ld = b)
    GMRFB_CU(ctx, cudaMemcpy2DAsync(h->iface.p + b * b, b * sizeof(double), work.p + 2 * bs, ld * sizeof(double),
                                    b * sizeof(double), b, cudaMemcpyDeviceToDevice, ctx->stream));
    GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));  // work and the plans are released at the end of this function
  }
  if (h->has_sep) {
    GMRFB_CU(ctx, h->V.alloc((size_t)bs));
    // V = E_r L_ni^{-T},  E_r = B_local[:,:,nloc-1] (rows: separator, cols: last interior block)
    GMRFB_CU(ctx, cudaMemcpy2DAsync(tmp.p, ld * sizeof(double), B_local + (nloc - 1) * b * b, b * sizeof(double),
                                    b * sizeof(double), b, cudaMemcpyDefault, ctx->stream));
    rc = gemm_once(ctx, LK_GEMM_NT, tmp.p, ld, F->winv.p + (h->ni - 1) * bs, ld, h->V.p, ld, ib, ib, ib, false, 1.0, 0.0);
    if (rc != GMRFB_OK) return rc;
    // iface0 = D_sep - V V'
    GMRFB_CU(ctx, cudaMemcpyAsync(h->iface.p, D_local + (nloc - 1) * b * b, b * b * sizeof(double), cudaMemcpyDefault,
                                  ctx->stream));
    rc = gemm_once(ctx, LK_GEMM_NT, h->V.p, ld, h->V.p, ld, h->iface.p, ib, ib, ib, ib, true, -1.0, 1.0);
    if (rc != GMRFB_OK) return rc;
    if (h->has_spike) {
      // R = V W_ni'  (block (S_r, S_{r-1}) of L_SI L_SI')
      rc = gemm_once(ctx, LK_GEMM_NT, h->V.p, ld, h->W.p + (h->ni - 1) * bs, ld, h->iface.p + 2 * b * b, ib, ib, ib, ib, false,
                     1.0, 0.0);
      if (rc != GMRFB_OK) return rc;
    }
  }
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  *out = h.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" int64_t gmrfb_btd_dist_iface_count(const gmrfb_btd_dist* h) { return h ? 3 * h->b * h->b : 0; }

extern "C" gmrfb_status gmrfb_btd_dist_get_iface(gmrfb_btd_dist* h, double* d_out) try {
  if (!h || !d_out) return fail(h ? h->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_dist_get_iface: NULL argument");
  GMRFB_CU(h->ctx, cudaSetDevice(h->ctx->device));
  GMRFB_CU(h->ctx, cudaMemcpyAsync(d_out, h->iface.p, 3 * h->b * h->b * sizeof(double), cudaMemcpyDeviceToDevice,
                                   h->ctx->stream));
  GMRFB_CU(h->ctx, cudaStreamSynchronize(h->ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_btd_dist_reduce(gmrfb_btd_dist* h, const double* d_all) try {
  if (!h || (h->P > 1 && !d_all)) return fail(h ? h->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_dist_reduce: NULL argument");
  gmrfb_ctx* ctx = h->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  if (h->P == 1) {
    h->reduced_ready = true;
    return GMRFB_OK;
  }
  const int64_t b = h->b, nb = h->P - 1, cnt = 3 * b * b;
  std::unique_ptr<gmrfb_btd> R;
  gmrfb_status rc = btd_alloc(ctx, b, nb, R);
  if (rc != GMRFB_OK) return rc;
  GMRFB_CU(ctx, cudaMemsetAsync(R->arena.p, 0, R->arena.n * sizeof(double), ctx->stream));
  for (int64_t r = 0; r < nb; r++) {
    // D_hat_r = iface0_r - Q_{r+1};   B_hat_r = -R_r (coupling with separator r-1)
    rc = block_axpby(ctx, b, 1.0, d_all + r * cnt, b, -1.0, d_all + (r + 1) * cnt + b * b, b, R->arena.p + r * R->slot, R->ld);
    if (rc != GMRFB_OK) return rc;
    if (r > 0) {
      rc = block_axpby(ctx, b, -1.0, d_all + r * cnt + 2 * b * b, b, 0.0, nullptr, 0,
                       R->arena.p + r * R->slot + (int64_t)R->ld * b, R->ld);
      if (rc != GMRFB_OK) return rc;
    }
  }
  rc = btd_run_factor(R.get());
  if (rc != GMRFB_OK) return rc;
  if (h->reduced) gmrfb_btd_destroy(h->reduced);
  h->reduced = R.release();
  h->reduced_ready = true;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" int64_t gmrfb_btd_dist_solve_count(const gmrfb_btd_dist* h, int64_t nrhs) {
  if (!h) return 0;
  const int64_t ldr = (nrhs + 1) & ~(int64_t)1;
  return 2 * ldr * h->b;
}

// Phase 1: local forward elimination; d_send (device) receives [ (b_S - V y_last)' | (sum_i W_i y_i)' ], each ldr x b.
extern "C" gmrfb_status gmrfb_btd_dist_solve_begin(gmrfb_btd_dist* h, const double* X_local, int64_t ldx, int64_t nrhs,
                                                   double* d_send) try {
  if (!h || !X_local || (h->P > 1 && !d_send))
    return fail(h ? h->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_dist_solve_begin: NULL argument");
  gmrfb_ctx* ctx = h->ctx;
  if (!h->reduced_ready) return fail(ctx, GMRFB_ERR_STATE, "gmrfb_btd_dist_solve_begin: call gmrfb_btd_dist_reduce first");
  const int64_t b = h->b, n = b * h->nloc;
  if (ldx < n || nrhs <= 0 || nrhs > 4096) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_btd_dist_solve_begin: bad ldx/nrhs");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int ldr = (int)((nrhs + 1) & ~(int64_t)1);
  const int inr = (int)nrhs, ib = (int)b;
  h->solve_nrhs = inr;
  h->solve_ldr = ldr;
  DevBuf<double> dX;
  GMRFB_CU(ctx, dX.alloc((size_t)(n * nrhs), ctx->stream));
  GMRFB_CU(ctx, h->Xt.alloc((size_t)((int64_t)ldr * n)));
  GMRFB_CU(ctx, cudaMemsetAsync(h->Xt.p, 0, h->Xt.n * sizeof(double), ctx->stream));
  GMRFB_CU(ctx, cudaMemcpy2DAsync(dX.p, n * sizeof(double), X_local, ldx * sizeof(double), n * sizeof(double), nrhs,
                                  cudaMemcpyHostToDevice, ctx->stream));
  gmrfb_status rc = transpose_dev(ctx, dX.p, n, h->Xt.p, ldr, n, nrhs);
  if (rc != GMRFB_OK) return rc;
  rc = btd_sweep(h->interior, true, false, h->Xt.p, ldr, inr);
  if (rc != GMRFB_OK) return rc;
  if (h->P == 1) return GMRFB_OK;
  const int64_t xstep = (int64_t)ldr * b, bs = (int64_t)h->ld * b;
  GMRFB_CU(ctx, cudaMemsetAsync(d_send, 0, 2 * xstep * sizeof(double), ctx->stream));
  if (h->has_sep) {
    // send0' = b_S' - y_last' V'
    GMRFB_CU(ctx, cudaMemcpyAsync(d_send, h->Xt.p + h->ni * xstep, xstep * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    rc = gemm_once(ctx, LK_GEMM_NT, h->Xt.p + (h->ni - 1) * xstep, ldr, h->V.p, h->ld, d_send, ldr, inr, ib, ib, false, -1.0, 1.0);
    if (rc != GMRFB_OK) return rc;
  }
  if (h->has_spike) {
    for (int64_t i = 0; i < h->ni; i++) {
      rc = gemm_once(ctx, LK_GEMM_NT, h->Xt.p + i * xstep, ldr, h->W.p + i * bs, h->ld, d_send + xstep, ldr, inr, ib, ib, false,
                     1.0, i == 0 ? 0.0 : 1.0);
      if (rc != GMRFB_OK) return rc;
    }
  }
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// Phase 2: reduced solve (redundant on every rank), local back-substitution, copy out this rank's rows.
extern "C" gmrfb_status gmrfb_btd_dist_solve_end(gmrfb_btd_dist* h, const double* d_all, double* X_local, int64_t ldx,
                                                 int64_t nrhs) try {
  if (!h || !X_local || (h->P > 1 && !d_all))
    return fail(h ? h->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_dist_solve_end: NULL argument");
  gmrfb_ctx* ctx = h->ctx;
  if (h->solve_nrhs != nrhs || !h->Xt.p) return fail(ctx, GMRFB_ERR_STATE, "gmrfb_btd_dist_solve_end: no matching solve_begin");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t b = h->b, n = b * h->nloc;
  const int ldr = h->solve_ldr, inr = (int)nrhs, ib = (int)b;
  const int64_t xstep = (int64_t)ldr * b, bs = (int64_t)h->ld * b;
  gmrfb_status rc;
  if (h->P > 1) {
    const int64_t nb = h->P - 1, cnt = 2 * xstep;
    GMRFB_CU(ctx, h->Sx.alloc((size_t)(xstep * nb)));
    for (int64_t r = 0; r < nb; r++) {
      // b_hat_r = send0_r - u_{r+1}   (treated as ldr x b blocks with leading dimension ldr)
      GMRFB_CU(ctx, launch_axpby(xstep, 1.0, d_all + r * cnt, -1.0, d_all + (r + 1) * cnt + xstep, h->Sx.p + r * xstep, ctx->stream));
      ctx->launches++;
    }
    rc = btd_sweep(h->reduced, true, true, h->Sx.p, ldr, inr);
    if (rc != GMRFB_OK) return rc;
    // y_i -= x_{S_{r-1}}' W_i  (all interior blocks);  y_last -= x_{S_r}' V
    if (h->has_spike) {
      for (int64_t i = 0; i < h->ni; i++) {
        rc = gemm_once(ctx, LK_GEMM_NN, h->Sx.p + (h->rank - 1) * xstep, ldr, h->W.p + i * bs, h->ld, h->Xt.p + i * xstep, ldr,
                       inr, ib, ib, false, -1.0, 1.0);
        if (rc != GMRFB_OK) return rc;
      }
    }
    if (h->has_sep) {
      rc = gemm_once(ctx, LK_GEMM_NN, h->Sx.p + h->rank * xstep, ldr, h->V.p, h->ld, h->Xt.p + (h->ni - 1) * xstep, ldr, inr, ib,
                     ib, false, -1.0, 1.0);
      if (rc != GMRFB_OK) return rc;
      GMRFB_CU(ctx, cudaMemcpyAsync(h->Xt.p + h->ni * xstep, h->Sx.p + h->rank * xstep, xstep * sizeof(double),
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  rc = btd_sweep(h->interior, false, true, h->Xt.p, ldr, inr);
  if (rc != GMRFB_OK) return rc;
  DevBuf<double> dX;
  GMRFB_CU(ctx, dX.alloc((size_t)(n * nrhs), ctx->stream));
  rc = transpose_dev(ctx, h->Xt.p, ldr, dX.p, n, nrhs, n);
  if (rc != GMRFB_OK) return rc;
  GMRFB_CU(ctx, cudaMemcpy2DAsync(X_local, ldx * sizeof(double), dX.p, n * sizeof(double), n * sizeof(double), nrhs,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_btd_dist_logdet(gmrfb_btd_dist* h, double* local_part, double* reduced_part) try {
  if (!h || !local_part || !reduced_part)
    return fail(h ? h->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_btd_dist_logdet: NULL argument");
  if (!h->reduced_ready) return fail(h->ctx, GMRFB_ERR_STATE, "gmrfb_btd_dist_logdet: call gmrfb_btd_dist_reduce first");
  gmrfb_status rc = gmrfb_btd_logdet(h->interior, local_part);
  if (rc != GMRFB_OK) return rc;
  *reduced_part = 0.0;
  if (h->reduced) rc = gmrfb_btd_logdet(h->reduced, reduced_part);
  return rc;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_btd_dist_destroy(gmrfb_btd_dist* h) try {
  if (!h) return GMRFB_OK;
  cudaSetDevice(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  delete h;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH
