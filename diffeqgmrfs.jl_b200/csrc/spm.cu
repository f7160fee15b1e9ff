// C ABI: device-resident sparse matrices, SpMV, sqmahal and the fixed-pattern posterior-precision assembly
//   Qpost = Q + A' diag(w) A    (condition_on_observations, scripts/darcy/solve_darcy_gmrf-fem.jl:165-167;
//                                Q + noise*J'*J, scripts/solve_burger.jl:145)
#include <algorithm>
#include <memory>
#include <thread>
#include <climits>
#include <cstdlib>

#include "common.hpp"
#include "handles.hpp"
#include "spgemm_kernels.cuh"

using namespace gmrfb;

struct gmrfb_postprec {
  gmrfb_ctx* ctx = nullptr;
  const gmrfb_spm* Q = nullptr;
  const gmrfb_spm* A = nullptr;
  gmrfb_spm out;  // owned result matrix
  DevBuf<int64_t> d_qsrc, d_pptr, d_pa, d_pb;
  DevBuf<int32_t> d_prow;
  DevBuf<double> d_w;
  int64_t nprod = 0;
};

gmrfb_status gmrfb::spm_build(gmrfb_ctx* ctx, gmrfb_spm* M, int64_t m, int64_t n, const int64_t* colptr,
                              const int64_t* rowval, const double* nzval, int base) {
  M->ctx = ctx;
  M->m = m;
  M->n = n;
  if (colptr[0] != base) return fail(ctx, GMRFB_ERR_INVALID, "spm: colptr[0] must equal base");
  M->nnz = colptr[n] - base;
  M->colptr.resize(n + 1);
  M->rowidx.resize(M->nnz);
  for (int64_t j = 0; j <= n; j++) M->colptr[j] = colptr[j] - base;
  for (int64_t j = 0; j < n; j++) {
    if (M->colptr[j + 1] < M->colptr[j]) return fail(ctx, GMRFB_ERR_INVALID, "spm: colptr is not monotone");
    for (int64_t p = M->colptr[j]; p < M->colptr[j + 1]; p++) {
      int64_t r = rowval[p] - base;
      if (r < 0 || r >= m) return fail(ctx, GMRFB_ERR_INVALID, "spm: row index out of range");
      if (p > M->colptr[j] && rowval[p] <= rowval[p - 1])
        return fail(ctx, GMRFB_ERR_INVALID, "spm: row indices must be strictly increasing within a column");
      M->rowidx[p] = (int32_t)r;
    }
  }
  // row-wise copy: rowptr/colidx and the map from row-wise position to CSC position
  std::vector<int64_t> rowptr(m + 1, 0), tmap(M->nnz);
  std::vector<int32_t> colidx(M->nnz);
  for (int64_t p = 0; p < M->nnz; p++) rowptr[M->rowidx[p] + 1]++;
  for (int64_t i = 0; i < m; i++) rowptr[i + 1] += rowptr[i];
  {
    std::vector<int64_t> fillp(rowptr.begin(), rowptr.end() - 1);
    for (int64_t j = 0; j < n; j++)
      for (int64_t p = M->colptr[j]; p < M->colptr[j + 1]; p++) {
        int64_t q = fillp[M->rowidx[p]]++;
        colidx[q] = (int32_t)j;
        tmap[q] = p;
      }
  }
  cudaStream_t st = ctx->stream;
  GMRFB_CU(ctx, M->d_colptr.upload(M->colptr, st));
  GMRFB_CU(ctx, M->d_rowidx.upload(M->rowidx, st));
  GMRFB_CU(ctx, M->d_rowptr.upload(rowptr, st));
  GMRFB_CU(ctx, M->d_colidx.upload(colidx, st));
  GMRFB_CU(ctx, M->d_tmap.upload(tmap, st));
  GMRFB_CU(ctx, M->d_val.alloc((size_t)std::max<int64_t>(M->nnz, 1)));
  GMRFB_CU(ctx, M->d_tval.alloc((size_t)std::max<int64_t>(M->nnz, 1)));
  if (nzval) {
    GMRFB_CU(ctx, cudaMemcpyAsync(M->d_val.p, nzval, M->nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    GMRFB_CU(ctx, launch_gather_values(M->d_val.p, M->d_tmap.p, M->nnz, M->d_tval.p, st));
    ctx->launches++;
    GMRFB_CU(ctx, cudaStreamSynchronize(st));
  }
  return GMRFB_OK;
}

extern "C" gmrfb_status gmrfb_spm_create(gmrfb_ctx* ctx, int64_t m, int64_t n, const int64_t* colptr,
                                         const int64_t* rowval, const double* nzval, int32_t base, gmrfb_spm** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_spm_create: ctx is NULL");
  if (!out || !colptr || m < 0 || n < 0 || (base != 0 && base != 1))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_spm_create: bad argument");
  if (m > 2000000000 || n > 2000000000) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_spm_create: dimension too large");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_spm> M(new gmrfb_spm());
  gmrfb_status rc = spm_build(ctx, M.get(), m, n, colptr, rowval, nzval, base);
  if (rc != GMRFB_OK) return rc;
  *out = M.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_spm_set_values(gmrfb_spm* A, const double* nzval) try {
  if (!A || !nzval) return fail(A ? A->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_spm_set_values: NULL argument");
  gmrfb_ctx* ctx = A->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  GMRFB_CU(ctx, cudaMemcpyAsync(A->d_val.p, nzval, A->nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GMRFB_CU(ctx, launch_gather_values(A->d_val.p, A->d_tmap.p, A->nnz, A->d_tval.p, ctx->stream));
  ctx->launches++;
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_spm_destroy(gmrfb_spm* A) try {
  if (!A) return GMRFB_OK;
  if (A->owned_by_plan) return fail(A->ctx, GMRFB_ERR_INVALID, "matrix is owned by a posterior-precision plan");
  cudaSetDevice(A->ctx->device);
  cudaStreamSynchronize(A->ctx->stream);
  delete A;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_spm_dims(const gmrfb_spm* A, int64_t* m, int64_t* n, int64_t* nnz) try {
  if (!A) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_spm_dims: NULL argument");
  if (m) *m = A->m;
  if (n) *n = A->n;
  if (nnz) *nnz = A->nnz;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_spm_get(const gmrfb_spm* A, int32_t base, int64_t* colptr, int64_t* rowval,
                                      double* nzval) try {
  if (!A) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_spm_get: NULL argument");
  gmrfb_ctx* ctx = A->ctx;
  if (colptr)
    for (int64_t j = 0; j <= A->n; j++) colptr[j] = A->colptr[j] + base;
  if (rowval)
    for (int64_t p = 0; p < A->nnz; p++) rowval[p] = A->rowidx[p] + base;
  if (nzval) {
    GMRFB_CU(ctx, cudaSetDevice(ctx->device));
    GMRFB_CU(ctx, cudaMemcpyAsync(nzval, A->d_val.p, A->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" const double* gmrfb_spm_values_dev(const gmrfb_spm* A) { return A ? A->d_val.p : nullptr; }

extern "C" gmrfb_status gmrfb_spmv(const gmrfb_spm* A, int32_t trans, double alpha, const double* x, double beta,
                                   double* y) try {
  if (!A || !x || !y) return fail(A ? A->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_spmv: NULL argument");
  gmrfb_ctx* ctx = A->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t nx = trans ? A->m : A->n, ny = trans ? A->n : A->m;
  DevBuf<double> dx, dy;
  GMRFB_CU(ctx, dx.alloc((size_t)std::max<int64_t>(nx, 1), ctx->stream));
  GMRFB_CU(ctx, dy.alloc((size_t)std::max<int64_t>(ny, 1), ctx->stream));
  GMRFB_CU(ctx, cudaMemcpyAsync(dx.p, x, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (beta != 0.0) GMRFB_CU(ctx, cudaMemcpyAsync(dy.p, y, ny * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (trans)  // y_j = sum over column j of A: the CSC arrays are the rows of A'
    GMRFB_CU(ctx, launch_spmv_rows(A->n, A->d_colptr.p, A->d_rowidx.p, A->d_val.p, dx.p, dy.p, alpha, beta, ctx->stream));
  else
    GMRFB_CU(ctx, launch_spmv_rows(A->m, A->d_rowptr.p, A->d_colidx.p, A->d_tval.p, dx.p, dy.p, alpha, beta, ctx->stream));
  ctx->launches++;
  GMRFB_CU(ctx, cudaMemcpyAsync(y, dy.p, ny * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_sqmahal(const gmrfb_spm* Q, const double* mu, const double* v, double* out) try {
  if (!Q || !v || !out) return fail(Q ? Q->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_sqmahal: NULL argument");
  if (Q->m != Q->n) return fail(Q->ctx, GMRFB_ERR_INVALID, "gmrfb_sqmahal: Q must be square");
  gmrfb_ctx* ctx = Q->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t n = Q->n;
  DevBuf<double> dv, dmu, dd, dt;
  GMRFB_CU(ctx, dv.alloc((size_t)std::max<int64_t>(n, 1), ctx->stream));
  GMRFB_CU(ctx, dd.alloc((size_t)std::max<int64_t>(n, 1), ctx->stream));
  GMRFB_CU(ctx, dt.alloc((size_t)std::max<int64_t>(n, 1), ctx->stream));
  GMRFB_CU(ctx, cudaMemcpyAsync(dv.p, v, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (mu) {
    GMRFB_CU(ctx, dmu.alloc((size_t)std::max<int64_t>(n, 1), ctx->stream));
    GMRFB_CU(ctx, cudaMemcpyAsync(dmu.p, mu, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  GMRFB_CU(ctx, launch_axpby(n, 1.0, dv.p, -1.0, mu ? dmu.p : nullptr, dd.p, ctx->stream));
  GMRFB_CU(ctx, launch_spmv_rows(n, Q->d_rowptr.p, Q->d_colidx.p, Q->d_tval.p, dd.p, dt.p, 1.0, 0.0, ctx->stream));
  GMRFB_CU(ctx, launch_dot(dd.p, dt.p, n, ctx->d_scalar, ctx->stream));
  ctx->launches += 3;
  GMRFB_CU(ctx, cudaMemcpyAsync(out, ctx->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// ------------------------------------------------------------------------------ posterior precision ----
extern "C" gmrfb_status gmrfb_postprec_create(gmrfb_ctx* ctx, const gmrfb_spm* Q, const gmrfb_spm* A,
                                              gmrfb_postprec** out) try {
  if (!ctx || !Q || !A || !out) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_postprec_create: NULL argument");
  if (Q->m != Q->n || A->n != Q->n) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_postprec_create: shape mismatch");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t n = Q->n, m = A->m;
  std::vector<int64_t> ocolptr, qsrc, pptr, pa, pb;
  std::vector<int32_t> orow, prow;
  spgemm::postprec_pattern(n, m, Q->colptr.data(), Q->rowidx.data(), A->colptr.data(), A->rowidx.data(), A->nnz, 0, ocolptr,
                           orow, qsrc, pptr, prow, pa, pb);
  std::unique_ptr<gmrfb_postprec> P(new gmrfb_postprec());
  P->ctx = ctx;
  P->Q = Q;
  P->A = A;
  P->nprod = (int64_t)prow.size();
  std::vector<int64_t> orow64(orow.begin(), orow.end());
  gmrfb_status rc = spm_build(ctx, &P->out, n, n, ocolptr.data(), orow64.data(), nullptr, 0);
  if (rc != GMRFB_OK) return rc;
  P->out.owned_by_plan = true;
  cudaStream_t st = ctx->stream;
  GMRFB_CU(ctx, P->d_qsrc.upload(qsrc, st));
  GMRFB_CU(ctx, P->d_pptr.upload(pptr, st));
  GMRFB_CU(ctx, P->d_pa.upload(pa, st));
  GMRFB_CU(ctx, P->d_pb.upload(pb, st));
  GMRFB_CU(ctx, P->d_prow.upload(prow, st));
  GMRFB_CU(ctx, P->d_w.alloc((size_t)std::max<int64_t>(m, 1)));
  *out = P.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_postprec_destroy(gmrfb_postprec* plan) try {
  if (!plan) return GMRFB_OK;
  cudaSetDevice(plan->ctx->device);
  cudaStreamSynchronize(plan->ctx->stream);
  delete plan;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_postprec_compute(gmrfb_postprec* plan, double qeps_scalar, const double* qeps_diag,
                                               const gmrfb_spm** Qpost) try {
  if (!plan) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_postprec_compute: NULL plan");
  gmrfb_ctx* ctx = plan->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  if (qeps_diag)  // host or device pointer (unified addressing)
    GMRFB_CU(ctx, cudaMemcpyAsync(plan->d_w.p, qeps_diag, plan->A->m * sizeof(double), cudaMemcpyDefault, ctx->stream));
  GMRFB_CU(ctx, launch_postprec(plan->out.nnz, plan->d_qsrc.p, plan->Q->d_val.p, plan->d_pptr.p, plan->d_prow.p,
                                plan->d_pa.p, plan->d_pb.p, plan->A->d_val.p, qeps_diag ? plan->d_w.p : nullptr,
                                qeps_scalar, plan->out.d_val.p, ctx->stream));
  GMRFB_CU(ctx, launch_gather_values(plan->out.d_val.p, plan->out.d_tmap.p, plan->out.nnz, plan->out.d_tval.p, ctx->stream));
  ctx->launches += 2;
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (Qpost) *Qpost = &plan->out;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_postprec_result(gmrfb_postprec* plan, const gmrfb_spm** Qpost) try {
  if (!plan || !Qpost) return fail(plan ? plan->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_postprec_result: NULL argument");
  *Qpost = &plan->out;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// ---------------------------------------------------------------------------- fixed-pattern product ----
// C = alpha A diag(w) B on patterns fixed at creation: the higher Matern powers K Mt^-1 K Mt^-1 K of
// src/spdes/shallow_water.jl:186 (commented there, needed for smoothness 2 in two dimensions,
// scripts/darcy/solve_darcy_gmrf-fem.jl:97) and any other repeated sparse product of a prior construction.
// Symbolic (host, once): the pattern of C, column by column, as the sorted union of the columns of A selected by the
// rows of B(:, j).  Numeric (device): one thread per nonzero C[i, j] walks B(:, j) in ascending row order and finds
// A[i, k] by binary search in the sorted column k of A - fixed summation order, no atomics, no product lists.
struct gmrfb_spgemm {
  gmrfb_ctx* ctx = nullptr;
  const gmrfb_spm* A = nullptr;
  const gmrfb_spm* B = nullptr;
  gmrfb_spm out;              // owned result matrix
  DevBuf<int32_t> d_colnz;    // column of every (CSC) nonzero of C
  DevBuf<double> d_w;
};

extern "C" gmrfb_status gmrfb_spgemm_create(gmrfb_ctx* ctx, const gmrfb_spm* A, const gmrfb_spm* B, gmrfb_spgemm** out) try {
  if (!ctx || !A || !B || !out) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_spgemm_create: NULL argument");
  if (A->n != B->m) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_spgemm_create: shape mismatch");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t m = A->m, n = B->n;
  std::vector<int64_t> ccolptr, crow;
  std::vector<int32_t> colnz;
  spgemm::product_pattern(m, n, A->colptr.data(), A->rowidx.data(), B->colptr.data(), B->rowidx.data(), ccolptr, crow, colnz);
  std::unique_ptr<gmrfb_spgemm> P(new gmrfb_spgemm());
  P->ctx = ctx;
  P->A = A;
  P->B = B;
  gmrfb_status rc = spm_build(ctx, &P->out, m, n, ccolptr.data(), crow.data(), nullptr, 0);
  if (rc != GMRFB_OK) return rc;
  P->out.owned_by_plan = true;
  GMRFB_CU(ctx, P->d_colnz.upload(colnz, ctx->stream));
  GMRFB_CU(ctx, P->d_w.alloc((size_t)std::max<int64_t>(A->n, 1)));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));  // `colnz` is a pageable temporary
  *out = P.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_spgemm_destroy(gmrfb_spgemm* plan) try {
  if (!plan) return GMRFB_OK;
  cudaSetDevice(plan->ctx->device);
  cudaStreamSynchronize(plan->ctx->stream);
  delete plan;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_spgemm_compute(gmrfb_spgemm* plan, double alpha, const double* w_diag,
                                             const gmrfb_spm** C_out) try {
  if (!plan) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_spgemm_compute: NULL plan");
  gmrfb_ctx* ctx = plan->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (w_diag)  // host or device pointer (unified addressing)
    GMRFB_CU(ctx, cudaMemcpyAsync(plan->d_w.p, w_diag, plan->A->n * sizeof(double), cudaMemcpyDefault, st));
  const int64_t nnz = plan->out.nnz;
  if (nnz > 0) {
    spgemm::k_spgemm<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, plan->out.d_rowidx.p, plan->d_colnz.p, plan->A->d_colptr.p,
                                                           plan->A->d_rowidx.p, plan->A->d_val.p, plan->B->d_colptr.p,
                                                           plan->B->d_rowidx.p, plan->B->d_val.p,
                                                           w_diag ? plan->d_w.p : nullptr, alpha, plan->out.d_val.p);
    GMRFB_CU(ctx, cudaGetLastError());
  }
  GMRFB_CU(ctx, launch_gather_values(plan->out.d_val.p, plan->out.d_tmap.p, nnz, plan->out.d_tval.p, st));
  ctx->launches += 2;
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (C_out) *C_out = &plan->out;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

// ------------------------------------------------------------------------------- evaluation metrics ----
// src/metrics.jl:3-13 on the device: pred = E x (or x), then
//   rmse = sqrt(mean((pred - truth).^2)),  max_err = max |pred - truth|,  rel_err = |pred - truth| / |truth|.
namespace {

constexpr int MET_BLOCKS = 512;

// per-block partials in a fixed order (bit-reproducible): [sum (p-t)^2, max |p-t|, sum t^2]
__global__ void __launch_bounds__(256) k_metrics_partial(int64_t n, const double* __restrict__ pred,
                                                         const double* __restrict__ truth, double* __restrict__ part) {
  __shared__ double sh[3][8];
  double s2 = 0.0, mx = 0.0, t2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double d = pred[i] - truth[i], t = truth[i];
    s2 += d * d;
    mx = fmax(mx, fabs(d));
    t2 += t * t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh[0][warp] = s2;
    sh[1][warp] = mx;
    sh[2][warp] = t2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int w = 0; w < 8; w++) {
      a += sh[0][w];
      b = fmax(b, sh[1][w]);
      c += sh[2][w];
    }
    part[3 * blockIdx.x] = a;
    part[3 * blockIdx.x + 1] = b;
    part[3 * blockIdx.x + 2] = c;
  }
}

__global__ void k_metrics_final(int nb, int64_t n, const double* __restrict__ part, double* __restrict__ out) {
  double a = 0.0, b = 0.0, c = 0.0;
  for (int k = 0; k < nb; k++) {
    a += part[3 * k];
    b = fmax(b, part[3 * k + 1]);
    c += part[3 * k + 2];
  }
  out[0] = sqrt(a / (double)n);
  out[1] = b;
  out[2] = sqrt(a) / sqrt(c);
}

}  // namespace

extern "C" gmrfb_status gmrfb_metrics(gmrfb_ctx* ctx, const gmrfb_spm* E, const double* x, const double* truth,
                                      int64_t ntruth, double* out3) try {
  if (!ctx || !x || !truth || !out3 || ntruth <= 0) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_metrics: bad argument");
  if (E && E->m != ntruth) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_metrics: E has the wrong number of rows");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t nx = E ? E->n : ntruth;
  DevBuf<double> dx, dt, dp, part;
  GMRFB_CU(ctx, dx.alloc((size_t)nx, ctx->stream));
  GMRFB_CU(ctx, dt.alloc((size_t)ntruth, ctx->stream));
  GMRFB_CU(ctx, part.alloc(3 * MET_BLOCKS + 3, ctx->stream));
  GMRFB_CU(ctx, cudaMemcpyAsync(dx.p, x, nx * sizeof(double), cudaMemcpyDefault, st));
  GMRFB_CU(ctx, cudaMemcpyAsync(dt.p, truth, ntruth * sizeof(double), cudaMemcpyDefault, st));
  const double* pred = dx.p;
  if (E) {
    GMRFB_CU(ctx, dp.alloc((size_t)ntruth, ctx->stream));
    GMRFB_CU(ctx, launch_spmv_rows(E->m, E->d_rowptr.p, E->d_colidx.p, E->d_tval.p, dx.p, dp.p, 1.0, 0.0, st));
    ctx->launches++;
    pred = dp.p;
  }
  const int nb = (int)std::min<int64_t>(MET_BLOCKS, (ntruth + 255) / 256);
  k_metrics_partial<<<nb, 256, 0, st>>>(ntruth, pred, dt.p, part.p);
  GMRFB_CU(ctx, cudaGetLastError());
  k_metrics_final<<<1, 1, 0, st>>>(nb, ntruth, part.p, part.p + 3 * MET_BLOCKS);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches += 2;
  GMRFB_CU(ctx, cudaMemcpyAsync(out3, part.p + 3 * MET_BLOCKS, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH
