// C ABI: Gauss-Newton on the device for a bilinear collocation residual (SURVEY.md §8f N1 + N2).
//
// The reference's explicit loop (scripts/solve_burger.jl:143-180) is, per iteration,
//     f, J = residual and tangent at x_k                     (:127-134; a host SciPy/Ferrite loop in the reference)
//     A    = Q + noise * J'J                                 (:145)
//     x+   = A^{-1} (Q mu + noise * J'(J x_k + y - f))       (:146-148, cholesky(A; perm = p) \ rhs)
//     stop when the relative change of  (mu-x)'Q(mu-x) + noise*|y-f|^2  is below tol, or after max_steps (:171-180)
// with the Burgers collocation residual  f(w) = A1 w - A0 w + dt (A1 w).*(D w) - dt nu D2 w.  That residual is an
// instance of
//     f(w) = L w + c (A w) .* (D w),      J(w) = L + c (diag(D w) A + diag(A w) D),
// with sparse L, A, D of one shape; an optional elementwise cubic term e .* w.^3 (square systems: the elliptic problem
// -lap u + u^3 = g of _research/elliptic_chen24.jl:231-285 with lumped mass, L = stiffness, e = mass) adds
// diag(3 e w^2) to the tangent.  The caller passes the union pattern of L, A, D in CSC form with three aligned value
// arrays; everything else happens here without leaving the device: residual and tangent (one row-wise and one
// entry-wise kernel), the fixed-pattern assembly Q + noise J'J, the numeric refactorisation on the analysed pattern,
// the solve, the objective (one scalar per iteration crosses PCIe, for the stopping rule).
#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "common.hpp"
#include "handles.hpp"

using namespace gmrfb;

struct gmrfb_gn {
  gmrfb_ctx* ctx = nullptr;
  const gmrfb_spm* Q = nullptr;   // prior precision (borrowed)
  gmrfb_spm* J = nullptr;         // tangent on the union pattern (values rewritten every iteration)
  gmrfb_postprec* plan = nullptr; // Q + noise J'J on a fixed pattern
  gmrfb_sym* sym = nullptr;
  gmrfb_fac* fac = nullptr;
  int64_t m = 0, n = 0, nnz = 0;
  double c = 0, noise = 0;
  DevBuf<double> lval, aval, dval;        // aligned to J's CSC positions
  DevBuf<double> cubic;                   // optional e (length n = m): f += e .* w.^3
  DevBuf<int64_t> diagpos;                // CSC position of J[i,i] (cubic term)
  DevBuf<double> y, mu, qmu, x, aw, dw, r, t, rhs, d, qd;
  int32_t steps = 0;
};

namespace {

// rows of the union pattern (row-wise copy of J): aw = A w, dw = D w, r = y - (L w + c aw .* dw)
__global__ void k_gn_residual(int64_t m, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                              const int64_t* __restrict__ tmap, const double* __restrict__ lval,
                              const double* __restrict__ aval, const double* __restrict__ dval, double c,
                              const double* __restrict__ w, const double* __restrict__ y,
                              const double* __restrict__ cubic, double* __restrict__ aw, double* __restrict__ dw,
                              double* __restrict__ r) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double sl = 0.0, sa = 0.0, sd = 0.0;
  for (int64_t q = rowptr[i]; q < rowptr[i + 1]; q++) {
    const int64_t p = tmap[q];
    const double wj = w[colidx[q]];
    sl += lval[p] * wj;
    sa += aval[p] * wj;
    sd += dval[p] * wj;
  }
  aw[i] = sa;
  dw[i] = sd;
  double f = sl + c * sa * sd;
  if (cubic) f += cubic[i] * w[i] * w[i] * w[i];
  r[i] = y[i] - f;
}

// J[i,i] += 3 e_i w_i^2
__global__ void k_gn_cubic_diag(int64_t n, const int64_t* __restrict__ diagpos, const double* __restrict__ cubic,
                                const double* __restrict__ w, double* __restrict__ jval) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) jval[diagpos[i]] += 3.0 * cubic[i] * w[i] * w[i];
}

// entries of the tangent in CSC order: J = L + c (diag(D w) A + diag(A w) D)
__global__ void k_gn_tangent(int64_t nnz, const int32_t* __restrict__ rowidx, const double* __restrict__ lval,
                             const double* __restrict__ aval, const double* __restrict__ dval, double c,
                             const double* __restrict__ aw, const double* __restrict__ dw, double* __restrict__ jval) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const int32_t i = rowidx[p];
  jval[p] = lval[p] + c * (dw[i] * aval[p] + aw[i] * dval[p]);
}

// out = a + s * b
__global__ void k_gn_axpy(int64_t n, const double* a, double s, const double* __restrict__ b, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + s * b[i];
}

inline unsigned blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

// residual, auxiliary products and the objective at the current iterate g->x
gmrfb_status gn_eval(gmrfb_gn* g, double* obj) {
  gmrfb_ctx* ctx = g->ctx;
  cudaStream_t st = ctx->stream;
  const gmrfb_spm* J = g->J;
  k_gn_residual<<<blocks(g->m), 256, 0, st>>>(g->m, J->d_rowptr.p, J->d_colidx.p, J->d_tmap.p, g->lval.p, g->aval.p,
                                              g->dval.p, g->c, g->x.p, g->y.p, g->cubic.p, g->aw.p, g->dw.p, g->r.p);
  GMRFB_CU(ctx, cudaGetLastError());
  // objective = (mu - x)' Q (mu - x) + noise * r'r
  GMRFB_CU(ctx, launch_axpby(g->n, 1.0, g->mu.p, -1.0, g->x.p, g->d.p, st));
  GMRFB_CU(ctx, launch_spmv_rows(g->n, g->Q->d_rowptr.p, g->Q->d_colidx.p, g->Q->d_tval.p, g->d.p, g->qd.p, 1.0, 0.0, st));
  GMRFB_CU(ctx, launch_dot(g->d.p, g->qd.p, g->n, ctx->d_scalar, st));
  GMRFB_CU(ctx, launch_dot(g->r.p, g->r.p, g->m, ctx->d_scalar + 1, st));
  ctx->launches += 5;
  double h[2];
  GMRFB_CU(ctx, cudaMemcpyAsync(h, ctx->d_scalar, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  *obj = h[0] + g->noise * h[1];
  return GMRFB_OK;
}

// one Gauss-Newton step from g->x (gn_eval must have run at g->x)
gmrfb_status gn_step(gmrfb_gn* g) {
  gmrfb_ctx* ctx = g->ctx;
  cudaStream_t st = ctx->stream;
  gmrfb_spm* J = g->J;
  k_gn_tangent<<<blocks(g->nnz), 256, 0, st>>>(g->nnz, J->d_rowidx.p, g->lval.p, g->aval.p, g->dval.p, g->c, g->aw.p,
                                               g->dw.p, J->d_val.p);
  GMRFB_CU(ctx, cudaGetLastError());
  if (g->cubic.p) {
    k_gn_cubic_diag<<<blocks(g->n), 256, 0, st>>>(g->n, g->diagpos.p, g->cubic.p, g->x.p, J->d_val.p);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches++;
  }
  GMRFB_CU(ctx, launch_gather_values(J->d_val.p, J->d_tmap.p, J->nnz, J->d_tval.p, st));
  // t = J x + r ;  rhs = Q mu + noise * J' t
  GMRFB_CU(ctx, launch_spmv_rows(g->m, J->d_rowptr.p, J->d_colidx.p, J->d_tval.p, g->x.p, g->t.p, 1.0, 0.0, st));
  k_gn_axpy<<<blocks(g->m), 256, 0, st>>>(g->m, g->t.p, 1.0, g->r.p, g->t.p);
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_spmv_rows(g->n, J->d_colptr.p, J->d_rowidx.p, J->d_val.p, g->t.p, g->rhs.p, 1.0, 0.0, st));
  k_gn_axpy<<<blocks(g->n), 256, 0, st>>>(g->n, g->qmu.p, g->noise, g->rhs.p, g->x.p);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches += 6;
  const gmrfb_spm* Apost = nullptr;
  gmrfb_status rc = gmrfb_postprec_compute(g->plan, g->noise, nullptr, &Apost);
  if (rc != GMRFB_OK) return rc;
  if ((rc = gmrfb_factorize_dev(g->fac, Apost->d_val.p)) != GMRFB_OK) return rc;
  return gmrfb_solve_dev(g->fac, GMRFB_SOLVE_A, g->x.p, g->n, 1);
}

}  // namespace

extern "C" gmrfb_status gmrfb_gn_create(gmrfb_ctx* ctx, const gmrfb_spm* Q, int64_t m, const int64_t* colptr,
                                        const int64_t* rowval, const double* lval, const double* aval,
                                        const double* dval, const double* cubic, int32_t base, double c, double noise,
                                        const double* y, const double* mu, const int64_t* perm,
                                        const gmrfb_analyze_opts* opts, gmrfb_gn** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_gn_create: ctx is NULL");
  if (!Q || !colptr || !rowval || !lval || !aval || !dval || !y || !mu || !out || m <= 0 || Q->m != Q->n)
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_gn_create: bad argument");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_gn> g(new gmrfb_gn());
  g->ctx = ctx;
  g->Q = Q;
  g->m = m;
  g->n = Q->n;
  g->c = c;
  g->noise = noise;
  gmrfb_status rc = gmrfb_spm_create(ctx, m, g->n, colptr, rowval, lval, base, &g->J);
  if (rc != GMRFB_OK) return rc;
  auto cleanup = [&](gmrfb_status code) {
    gmrfb_gn_destroy(g.release());
    return code;
  };
  g->nnz = g->J->nnz;
  cudaStream_t st = ctx->stream;
  auto up = [&](DevBuf<double>& b, const double* h, int64_t cnt) -> cudaError_t {
    cudaError_t e = b.alloc((size_t)std::max<int64_t>(cnt, 1));
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(b.p, h, cnt * sizeof(double), cudaMemcpyHostToDevice, st);
  };
  if (up(g->lval, lval, g->nnz) != cudaSuccess || up(g->aval, aval, g->nnz) != cudaSuccess ||
      up(g->dval, dval, g->nnz) != cudaSuccess || up(g->y, y, m) != cudaSuccess || up(g->mu, mu, g->n) != cudaSuccess)
    return cleanup(fail(ctx, GMRFB_ERR_ALLOC, "gmrfb_gn_create: device allocation failed"));
  if (cubic) {
    if (m != g->n) return cleanup(fail(ctx, GMRFB_ERR_INVALID, "gmrfb_gn_create: the cubic term needs a square system"));
    std::vector<int64_t> dp((size_t)g->n, -1);
    for (int64_t j = 0; j < g->n; j++)
      for (int64_t p = g->J->colptr[j]; p < g->J->colptr[j + 1]; p++)
        if (g->J->rowidx[p] == j) dp[j] = p;
    for (int64_t j = 0; j < g->n; j++)
      if (dp[j] < 0) return cleanup(fail(ctx, GMRFB_ERR_INVALID, "gmrfb_gn_create: the pattern must hold the diagonal for the cubic term"));
    if (up(g->cubic, cubic, g->n) != cudaSuccess || g->diagpos.upload(dp, st) != cudaSuccess)
      return cleanup(fail(ctx, GMRFB_ERR_ALLOC, "gmrfb_gn_create: device allocation failed"));
  }
  for (DevBuf<double>* b : {&g->qmu, &g->x, &g->rhs, &g->d, &g->qd})
    if (b->alloc((size_t)g->n) != cudaSuccess) return cleanup(fail(ctx, GMRFB_ERR_ALLOC, "gmrfb_gn_create: device allocation failed"));
  for (DevBuf<double>* b : {&g->aw, &g->dw, &g->r, &g->t})
    if (b->alloc((size_t)m) != cudaSuccess) return cleanup(fail(ctx, GMRFB_ERR_ALLOC, "gmrfb_gn_create: device allocation failed"));
  // Q mu once
  if (launch_spmv_rows(g->n, Q->d_rowptr.p, Q->d_colidx.p, Q->d_tval.p, g->mu.p, g->qmu.p, 1.0, 0.0, st) != cudaSuccess)
    return cleanup(fail(ctx, GMRFB_ERR_CUDA, "gmrfb_gn_create: SpMV launch failed"));
  ctx->launches++;
  if ((rc = gmrfb_postprec_create(ctx, Q, g->J, &g->plan)) != GMRFB_OK) return cleanup(rc);
  // symbolic analysis of the pattern of Q + J'J (once; every iteration refactorises numerically on it)
  const gmrfb_spm* Apost = nullptr;
  if ((rc = gmrfb_postprec_compute(g->plan, noise, nullptr, &Apost)) != GMRFB_OK) return cleanup(rc);
  {
    std::vector<int64_t> cp(Apost->colptr.begin(), Apost->colptr.end()), ri(Apost->rowidx.begin(), Apost->rowidx.end());
    gmrfb_analyze_opts o{};
    if (opts) o = *opts;
    if (perm) o.ordering_kind = GMRFB_ORDER_GIVEN;
    const int32_t pbase = o.base;
    o.base = 0;  // the pattern handed over here is 0-based; a `base`-based perm is shifted below
    std::vector<int64_t> p0;
    if (perm) {
      p0.assign(perm, perm + g->n);
      for (auto& v : p0) v -= pbase;
    }
    if ((rc = gmrfb_analyze(ctx, g->n, cp.data(), ri.data(), perm ? p0.data() : nullptr, &o, &g->sym)) != GMRFB_OK)
      return cleanup(rc);
  }
  if ((rc = gmrfb_fac_create(g->sym, &g->fac)) != GMRFB_OK) return cleanup(rc);
  if (cudaStreamSynchronize(st) != cudaSuccess) return cleanup(fail(ctx, GMRFB_ERR_CUDA, "gmrfb_gn_create: stream error"));
  *out = g.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_gn_optimize(gmrfb_gn* g, double* x, int32_t max_steps, double rel_tol, int32_t* steps,
                                          double* obj_hist) try {
  if (!g || !x || max_steps < 0) return fail(g ? g->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_gn_optimize: bad argument");
  gmrfb_ctx* ctx = g->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  GMRFB_CU(ctx, cudaMemcpyAsync(g->x.p, x, g->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  double obj = 0.0, last = INFINITY;
  gmrfb_status rc = gn_eval(g, &obj);
  if (rc != GMRFB_OK) return rc;
  if (obj_hist) obj_hist[0] = obj;
  int32_t k = 0;
  // stopping rule of scripts/solve_burger.jl:171-180
  while (std::fabs(last - obj) / std::fabs(obj) > rel_tol && k < max_steps) {
    if ((rc = gn_step(g)) != GMRFB_OK) return rc;
    last = obj;
    if ((rc = gn_eval(g, &obj)) != GMRFB_OK) return rc;
    k++;
    if (obj_hist) obj_hist[k] = obj;
  }
  g->steps = k;
  if (steps) *steps = k;
  GMRFB_CU(ctx, cudaMemcpyAsync(x, g->x.p, g->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_gn_get(gmrfb_gn* g, gmrfb_fac** fac, gmrfb_sym** sym, const gmrfb_spm** J,
                                     const gmrfb_spm** Qpost) try {
  if (!g) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_gn_get: NULL handle");
  if (fac) *fac = g->fac;
  if (sym) *sym = g->sym;
  if (J) *J = g->J;
  if (Qpost) {
    const gmrfb_spm* A = nullptr;
    // the plan's result matrix (values of the last assembly)
    gmrfb_status rc = gmrfb_postprec_result(g->plan, &A);
    if (rc != GMRFB_OK) return rc;
    *Qpost = A;
  }
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_gn_destroy(gmrfb_gn* g) try {
  if (!g) return GMRFB_OK;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  if (g->fac) gmrfb_fac_destroy(g->fac);
  if (g->sym) gmrfb_sym_destroy(g->sym);
  if (g->plan) gmrfb_postprec_destroy(g->plan);
  if (g->J) gmrfb_spm_destroy(g->J);
  delete g;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH
