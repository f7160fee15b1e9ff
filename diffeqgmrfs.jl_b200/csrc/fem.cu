// C ABI: P1 finite-element assembly on the device for the two matrices the reference rebuilds on its hot loops:
//   * the Darcy stiffness G(a) = sum_T a_T K_T of a new coefficient field for every problem of the dataset loop
//     (src/problems/darcy.jl:5-63 assemble_darcy_diff_matrix; coefficient looked up by nearest grid index,
//     src/datasets/darcy.jl:30-34; called from form_observations in scripts/darcy/solve_darcy_gmrf-fem.jl:104-137,178);
//   * the Matern prior precision Q = ratio * K' Mt^-1 K, K = kappa^2 Mt + G (src/spdes/shallow_water.jl:177-194);
//   * the Gauss-Newton tangent of the cubic reaction term, J = s G + 3 int u^2 phi_i phi_j and
//     f = s G u + int u^3 phi_i, by quadrature of the current iterate (_research/elliptic_chen24.jl:180-285
//     assemble_J_diff_and_f / assemble_J_cube / f_and_J), reassembled at every Gauss-Newton step.
// The mesh is analysed once on the host (pattern of the stiffness matrix, the list of element entries that fall on
// every nonzero, the grid cell of every element); every assembly is then one gather kernel over the nonzeros —
// deterministic (fixed summation order), no atomics — writing straight into a device matrix that the posterior-
// precision plan (gmrfb_postprec) and the factorisation consume without a host round trip.
#include <algorithm>
#include <cmath>
#include <memory>
#include <numeric>

#include "common.hpp"
#include "fem_pattern.hpp"
#include "handles.hpp"

using namespace gmrfb;

struct gmrfb_postprec;
extern "C" gmrfb_status gmrfb_postprec_create(gmrfb_ctx*, const gmrfb_spm*, const gmrfb_spm*, gmrfb_postprec**);
extern "C" gmrfb_status gmrfb_postprec_destroy(gmrfb_postprec*);
extern "C" gmrfb_status gmrfb_postprec_compute(gmrfb_postprec*, double, const double*, const gmrfb_spm**);

struct gmrfb_fem {
  gmrfb_ctx* ctx = nullptr;
  int64_t nn = 0, nt = 0;
  gmrfb_spm G;  // stiffness pattern (n x n, full symmetric storage), values of the last assembly
  gmrfb_spm K;  // same pattern: kappa^2 Mt + G of the last Matern call
  gmrfb_spm Z;  // empty n x n matrix (the "Q" of the posterior-precision plan that forms K' W K)
  gmrfb_spm J;  // same pattern: tangent of the last gmrfb_fem_assemble_cubic call
  bool G_built = false, K_built = false;
  DevBuf<double> d_kloc;   // nt x 9: geometric element matrices (unit coefficient), entry t * 9 + 3 i + j
  DevBuf<double> d_area;   // nt
  DevBuf<double> d_mass;   // nn: lumped mass
  DevBuf<int64_t> d_cptr;  // nnz + 1: contributions of every nonzero
  DevBuf<int32_t> d_cidx;  // element entry t * 9 + 3 i + j of every contribution (ascending per nonzero)
  DevBuf<int64_t> d_diag;  // nn: position of the diagonal entry of every column
  DevBuf<int32_t> d_cell;  // nt: coefficient-grid cell of every element (after gmrfb_fem_set_coeff_grid)
  DevBuf<double> d_coeff;  // staged coefficient grid
  DevBuf<uint8_t> d_presc; // nn: prescribed (Dirichlet) dofs of the last call
  DevBuf<double> d_w;      // nn: ratio / Mt
  DevBuf<int32_t> d_tris;  // 3 nt: vertices of every element
  DevBuf<double> d_u, d_f; // nn: staged iterate / residual of the cubic tangent
  DevBuf<double> d_quad;   // 4 nq: barycentric quadrature points and weights of the last requested degree
  int quad_degree = 0, nq = 0;
  std::vector<double> centroid;  // 2 nt (host): element centroids = the quadrature points of the P1 rule
  std::vector<double> mass;      // nn (host copy)
  int64_t ncell = 0;
  gmrfb_postprec* matern_plan = nullptr;
};

namespace {

__global__ void k_fem_geom(int64_t nt, const double* __restrict__ nodes, const int32_t* __restrict__ tris,
                           double* __restrict__ kloc, double* __restrict__ area) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const int32_t v0 = tris[3 * t], v1 = tris[3 * t + 1], v2 = tris[3 * t + 2];
  const double x0 = nodes[2 * (int64_t)v0], y0 = nodes[2 * (int64_t)v0 + 1];
  const double x1 = nodes[2 * (int64_t)v1], y1 = nodes[2 * (int64_t)v1 + 1];
  const double x2 = nodes[2 * (int64_t)v2], y2 = nodes[2 * (int64_t)v2 + 1];
  const double a2 = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0);  // twice the signed area
  const double ar = 0.5 * fabs(a2);
  // gradients of the hat functions: g_k = rot90(edge opposite to vertex k) / (2 area)
  const double ex[3] = {x2 - x1, x0 - x2, x1 - x0}, ey[3] = {y2 - y1, y0 - y2, y1 - y0};
  double gx[3], gy[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    gx[k] = -ey[k] / a2;
    gy[k] = ex[k] / a2;
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) kloc[9 * t + 3 * i + j] = (gx[i] * gx[j] + gy[i] * gy[j]) * ar;
  area[t] = ar;
}

// G[k] = sum over the element entries on nonzero k of coeff(element) * K_element[i, j]; rows of prescribed dofs
// become identity rows (the reference applies its Dirichlet constraints to the assembled matrix the same way)
__global__ void k_fem_assemble(int64_t nnz, const int64_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                               const double* __restrict__ kloc, const int32_t* __restrict__ cell,
                               const double* __restrict__ coeff, const int32_t* __restrict__ rowidx,
                               const int64_t* __restrict__ diag, const uint8_t* __restrict__ presc,
                               double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int32_t r = rowidx[k];
  if (presc && presc[r]) {
    out[k] = (k == diag[r]) ? 1.0 : 0.0;  // diag[r] is the position of entry (r, r)
    return;
  }
  double v = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
    const int32_t e = cidx[p];
    const double a = coeff ? coeff[cell[e / 9]] : 1.0;
    v += a * kloc[e];
  }
  out[k] = v;
}

// mass[i] = sum over elements at node i of area / 3 (gathered through the diagonal's contribution list: the diagonal
// entry of node i receives exactly one contribution per element at i)
__global__ void k_fem_mass(int64_t nn, const int64_t* __restrict__ diag, const int64_t* __restrict__ cptr,
                           const int32_t* __restrict__ cidx, const double* __restrict__ area, double* __restrict__ mass) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int64_t k = diag[i];
  double m = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) m += area[cidx[p] / 9] / 3.0;
  mass[i] = m;
}

// K = kappa^2 Mt + G with Mt = lumped mass (prescribed dofs: Mt_ii = presc_mass, G_ii = 1, as
// src/spdes/shallow_water.jl:178-181), and w_i = ratio / Mt_i
__global__ void k_fem_matern_k(int64_t nn, const int64_t* __restrict__ diag, const double* __restrict__ mass,
                               const uint8_t* __restrict__ presc, double presc_mass, double kappa2, double ratio,
                               double* __restrict__ kval, double* __restrict__ w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const bool p = presc && presc[i];
  const double mt = p ? presc_mass : mass[i];
  if (p) kval[diag[i]] = 1.0;
  kval[diag[i]] += kappa2 * mt;
  w[i] = ratio / mt;
}

// Tangent of the cubic term (_research/elliptic_chen24.jl:231-278) fused with s * stiffness (:180-228), one thread per
// nonzero: J[k] = sum over the element entries (t, i, j) on k of  s K_t[i, j] + 3 |t| sum_q w_q l_i(q) l_j(q) u_q^2,
// u_q = sum_k u[v_k] l_k(q); rows of prescribed dofs are skipped (stay zero) as the reference's `continue` does.
__global__ void k_fem_cubic_J(int64_t nnz, const int64_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                              const double* __restrict__ kloc, const double* __restrict__ area,
                              const int32_t* __restrict__ tris, const double* __restrict__ u,
                              const int32_t* __restrict__ rowidx, const uint8_t* __restrict__ presc, double s,
                              const double* __restrict__ quad, int nq, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  if (presc && presc[rowidx[k]]) {
    out[k] = 0.0;
    return;
  }
  double v = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
    const int32_t e = cidx[p];
    const int32_t t = e / 9, i = (e % 9) / 3, j = e % 3;
    const double w0 = u[tris[3 * t]], w1 = u[tris[3 * t + 1]], w2 = u[tris[3 * t + 2]];
    double acc = 0.0;
    for (int q = 0; q < nq; q++) {
      const double* l = quad + 4 * q;
      const double uq = l[0] * w0 + l[1] * w1 + l[2] * w2;
      acc += l[3] * l[i] * l[j] * uq * uq;
    }
    v += s * kloc[e] + 3.0 * area[t] * acc;
  }
  out[k] = v;
}

// f[i] = sum over the elements t at node i (local index a) of  s sum_j K_t[a, j] u_j + |t| sum_q w_q l_a(q) u_q^3
// (gathered through the diagonal's contribution list: one entry per element at i); prescribed rows stay zero
__global__ void k_fem_cubic_f(int64_t nn, const int64_t* __restrict__ diag, const int64_t* __restrict__ cptr,
                              const int32_t* __restrict__ cidx, const double* __restrict__ kloc,
                              const double* __restrict__ area, const int32_t* __restrict__ tris,
                              const double* __restrict__ u, const uint8_t* __restrict__ presc, double s,
                              const double* __restrict__ quad, int nq, double* __restrict__ f) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  if (presc && presc[i]) {
    f[i] = 0.0;
    return;
  }
  const int64_t k = diag[i];
  double v = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
    const int32_t e = cidx[p];
    const int32_t t = e / 9, a = e % 3;
    const double w0 = u[tris[3 * t]], w1 = u[tris[3 * t + 1]], w2 = u[tris[3 * t + 2]];
    double acc = 0.0;
    for (int q = 0; q < nq; q++) {
      const double* l = quad + 4 * q;
      const double uq = l[0] * w0 + l[1] * w1 + l[2] * w2;
      acc += l[3] * l[a] * uq * uq * uq;
    }
    const double* kr = kloc + 9 * (int64_t)t + 3 * a;
    v += s * (kr[0] * w0 + kr[1] * w1 + kr[2] * w2) + area[t] * acc;
  }
  f[i] = v;
}

// symmetric triangle rules exact to degree 1, 2 (3 interior points: QuadratureRule{RefTriangle}(2)) and 4 (6 points);
// rows: barycentric coordinates l0, l1, l2 and the weight (weights sum to 1: d Omega = weight * area)
std::vector<double> tri_rule(int degree) {
  if (degree == 1) return {1.0 / 3, 1.0 / 3, 1.0 / 3, 1.0};
  if (degree == 2) {
    const double a = 1.0 / 6, b = 2.0 / 3, w = 1.0 / 3;
    return {b, a, a, w, a, b, a, w, a, a, b, w};
  }
  std::vector<double> r;
  const double as[2] = {0.445948490915965, 0.091576213509771}, ws[2] = {0.223381589678011, 0.109951743655322};
  for (int g = 0; g < 2; g++) {
    const double a = as[g], b = 1.0 - 2.0 * a, w = ws[g];
    const double rows[12] = {b, a, a, w, a, b, a, w, a, a, b, w};
    r.insert(r.end(), rows, rows + 12);
  }
  return r;
}

}  // namespace

extern "C" gmrfb_status gmrfb_fem_create(gmrfb_ctx* ctx, int64_t nnodes, const double* nodes, int64_t ntri,
                                         const int64_t* tris, int32_t base, gmrfb_fem** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_fem_create: ctx is NULL");
  if (!out || !nodes || !tris || nnodes <= 0 || ntri <= 0 || (base != 0 && base != 1))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem_create: bad argument");
  if (nnodes > 2000000000 || ntri > 200000000) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem_create: mesh too large");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_fem> F(new gmrfb_fem());
  F->ctx = ctx;
  F->nn = nnodes;
  F->nt = ntri;
  std::vector<int32_t> tr((size_t)3 * ntri);
  for (int64_t k = 0; k < 3 * ntri; k++) {
    const int64_t v = tris[k] - base;
    if (v < 0 || v >= nnodes) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem_create: vertex index out of range");
    tr[k] = (int32_t)v;
  }
  F->centroid.resize((size_t)2 * ntri);
  for (int64_t t = 0; t < ntri; t++)
    for (int c = 0; c < 2; c++)
      F->centroid[2 * t + c] = (nodes[2 * (int64_t)tr[3 * t] + c] + nodes[2 * (int64_t)tr[3 * t + 1] + c] +
                                nodes[2 * (int64_t)tr[3 * t + 2] + c]) / 3.0;
  // pattern of the stiffness matrix and the element entries that fall on every nonzero
  const int64_t ne = 9 * ntri;
  ElementPattern P;
  if (!build_element_pattern(nnodes, ntri, 3, tr.data(), P))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem_create: a node belongs to no element");
  const std::vector<int64_t>&colptr = P.colptr, &rowval = P.rowval, &cptr = P.cptr, &diag = P.diag;
  const std::vector<int32_t>& cidx = P.cidx;
  gmrfb_status rc = spm_build(ctx, &F->G, nnodes, nnodes, colptr.data(), rowval.data(), nullptr, 0);
  if (rc != GMRFB_OK) return rc;
  if ((rc = spm_build(ctx, &F->K, nnodes, nnodes, colptr.data(), rowval.data(), nullptr, 0)) != GMRFB_OK) return rc;
  if ((rc = spm_build(ctx, &F->J, nnodes, nnodes, colptr.data(), rowval.data(), nullptr, 0)) != GMRFB_OK) return rc;
  {
    std::vector<int64_t> zp((size_t)nnodes + 1, 0);
    if ((rc = spm_build(ctx, &F->Z, nnodes, nnodes, zp.data(), nullptr, nullptr, 0)) != GMRFB_OK) return rc;
  }
  F->G.owned_by_plan = F->K.owned_by_plan = F->J.owned_by_plan = F->Z.owned_by_plan = true;  // borrowed views: not destroyed by the caller
  cudaStream_t st = ctx->stream;
  DevBuf<double> d_nodes;
  DevBuf<int32_t>& d_tris = F->d_tris;
  {
    std::vector<double> nd(nodes, nodes + 2 * nnodes);
    GMRFB_CU(ctx, d_nodes.upload(nd, st));
  }
  GMRFB_CU(ctx, d_tris.upload(tr, st));
  GMRFB_CU(ctx, F->d_kloc.alloc((size_t)ne));
  GMRFB_CU(ctx, F->d_area.alloc((size_t)ntri));
  GMRFB_CU(ctx, F->d_mass.alloc((size_t)nnodes));
  GMRFB_CU(ctx, F->d_cptr.upload(cptr, st));
  GMRFB_CU(ctx, F->d_cidx.upload(cidx, st));
  GMRFB_CU(ctx, F->d_diag.upload(diag, st));
  k_fem_geom<<<(unsigned)((ntri + 255) / 256), 256, 0, st>>>(ntri, d_nodes.p, d_tris.p, F->d_kloc.p, F->d_area.p);
  GMRFB_CU(ctx, cudaGetLastError());
  k_fem_mass<<<(unsigned)((nnodes + 255) / 256), 256, 0, st>>>(nnodes, F->d_diag.p, F->d_cptr.p, F->d_cidx.p, F->d_area.p,
                                                              F->d_mass.p);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches += 2;
  F->mass.resize((size_t)nnodes);
  GMRFB_CU(ctx, cudaMemcpyAsync(F->mass.data(), F->d_mass.p, nnodes * sizeof(double), cudaMemcpyDeviceToHost, st));
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  *out = F.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem_destroy(gmrfb_fem* F) try {
  if (!F) return GMRFB_OK;
  cudaSetDevice(F->ctx->device);
  cudaStreamSynchronize(F->ctx->stream);
  if (F->matern_plan) gmrfb_postprec_destroy(F->matern_plan);
  delete F;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem_get_mass(gmrfb_fem* F, double* mass_out) try {
  if (!F || !mass_out) return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem_get_mass: NULL argument");
  std::copy(F->mass.begin(), F->mass.end(), mass_out);
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem_set_coeff_grid(gmrfb_fem* F, int64_t gx, const double* x_coords, int64_t gy,
                                                 const double* y_coords) try {
  if (!F || !x_coords || !y_coords || gx <= 0 || gy <= 0)
    return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem_set_coeff_grid: bad argument");
  gmrfb_ctx* ctx = F->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  // nearest grid index per axis (argmin |coords - x|, first minimum on ties: src/datasets/darcy.jl:30-34)
  auto nearest = [](const double* c, int64_t g, double x) {
    int64_t best = 0;
    double bd = std::fabs(c[0] - x);
    for (int64_t k = 1; k < g; k++) {
      const double d = std::fabs(c[k] - x);
      if (d < bd) {
        bd = d;
        best = k;
      }
    }
    return best;
  };
  const bool xs = std::is_sorted(x_coords, x_coords + gx), ys = std::is_sorted(y_coords, y_coords + gy);
  auto nearest_sorted = [](const double* c, int64_t g, double x) {
    const int64_t hi = std::lower_bound(c, c + g, x) - c;
    if (hi <= 0) return (int64_t)0;
    if (hi >= g) return g - 1;
    return (std::fabs(c[hi - 1] - x) <= std::fabs(c[hi] - x)) ? hi - 1 : hi;  // ties: the first (lower) index
  };
  std::vector<int32_t> cell((size_t)F->nt);
  for (int64_t t = 0; t < F->nt; t++) {
    const double x = F->centroid[2 * t], y = F->centroid[2 * t + 1];
    const int64_t ix = xs ? nearest_sorted(x_coords, gx, x) : nearest(x_coords, gx, x);
    const int64_t iy = ys ? nearest_sorted(y_coords, gy, y) : nearest(y_coords, gy, y);
    cell[t] = (int32_t)(ix + iy * gx);  // coeff_mat[x_idx, y_idx] of a column-major gx x gy array
  }
  GMRFB_CU(ctx, F->d_cell.upload(cell, ctx->stream));
  F->ncell = gx * gy;
  GMRFB_CU(ctx, F->d_coeff.alloc((size_t)F->ncell));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

static gmrfb_status fem_upload_presc(gmrfb_fem* F, const uint8_t* prescribed) {
  gmrfb_ctx* ctx = F->ctx;
  if (!prescribed) return GMRFB_OK;
  if (!F->d_presc.p) GMRFB_CU(ctx, F->d_presc.alloc((size_t)F->nn));
  GMRFB_CU(ctx, cudaMemcpyAsync(F->d_presc.p, prescribed, F->nn, cudaMemcpyDefault, ctx->stream));
  return GMRFB_OK;
}

extern "C" gmrfb_status gmrfb_fem_assemble(gmrfb_fem* F, const double* coeff_grid, const uint8_t* prescribed,
                                           const gmrfb_spm** G_out) try {
  if (!F) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_fem_assemble: NULL handle");
  gmrfb_ctx* ctx = F->ctx;
  if (coeff_grid && !F->d_cell.p)
    return fail(ctx, GMRFB_ERR_STATE, "gmrfb_fem_assemble: call gmrfb_fem_set_coeff_grid before passing a coefficient grid");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (coeff_grid)  // host or device pointer (unified addressing)
    GMRFB_CU(ctx, cudaMemcpyAsync(F->d_coeff.p, coeff_grid, F->ncell * sizeof(double), cudaMemcpyDefault, st));
  gmrfb_status rc = fem_upload_presc(F, prescribed);
  if (rc != GMRFB_OK) return rc;
  {
    ProfScope ps(ctx, PK_FEM, 0, 12.0 * 9 * F->nt + 8.0 * F->G.nnz, 0, 0);
    k_fem_assemble<<<(unsigned)((F->G.nnz + 255) / 256), 256, 0, st>>>(
        F->G.nnz, F->d_cptr.p, F->d_cidx.p, F->d_kloc.p, F->d_cell.p, coeff_grid ? F->d_coeff.p : nullptr, F->G.d_rowidx.p,
        F->d_diag.p, prescribed ? F->d_presc.p : nullptr, F->G.d_val.p);
  }
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->G.d_val.p, F->G.d_tmap.p, F->G.nnz, F->G.d_tval.p, st));
  ctx->launches += 2;
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  F->G_built = true;
  if (G_out) *G_out = &F->G;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem_matern_precision(gmrfb_fem* F, double kappa, double ratio, const uint8_t* prescribed,
                                                   double prescribed_mass, const gmrfb_spm** Q_out) try {
  if (!F || !Q_out) return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem_matern_precision: NULL argument");
  gmrfb_ctx* ctx = F->ctx;
  if (!(kappa > 0) || !(ratio > 0) || (prescribed && !(prescribed_mass > 0)))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem_matern_precision: kappa, ratio and prescribed_mass must be positive");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  gmrfb_status rc = fem_upload_presc(F, prescribed);
  if (rc != GMRFB_OK) return rc;
  if (!F->d_w.p) GMRFB_CU(ctx, F->d_w.alloc((size_t)F->nn));
  // K = kappa^2 Mt + G (unit coefficient; no identity rows: the reference only touches the diagonal of prescribed dofs)
  k_fem_assemble<<<(unsigned)((F->K.nnz + 255) / 256), 256, 0, st>>>(F->K.nnz, F->d_cptr.p, F->d_cidx.p, F->d_kloc.p, nullptr,
                                                                   nullptr, F->K.d_rowidx.p, F->d_diag.p, nullptr,
                                                                   F->K.d_val.p);
  GMRFB_CU(ctx, cudaGetLastError());
  k_fem_matern_k<<<(unsigned)((F->nn + 255) / 256), 256, 0, st>>>(F->nn, F->d_diag.p, F->d_mass.p,
                                                                prescribed ? F->d_presc.p : nullptr, prescribed_mass,
                                                                kappa * kappa, ratio, F->K.d_val.p, F->d_w.p);
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->K.d_val.p, F->K.d_tmap.p, F->K.nnz, F->K.d_tval.p, st));
  ctx->launches += 3;
  // Q = K' diag(ratio / Mt) K on a fixed pattern: the plan is symbolic work done once per mesh
  if (!F->matern_plan) {
    rc = gmrfb_postprec_create(ctx, &F->Z, &F->K, &F->matern_plan);
    if (rc != GMRFB_OK) return rc;
  }
  return gmrfb_postprec_compute(F->matern_plan, 0.0, F->d_w.p, Q_out);
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem_assemble_cubic(gmrfb_fem* F, const double* u, int32_t quad_degree,
                                                 double stiffness_scale, const uint8_t* prescribed,
                                                 const gmrfb_spm** J_out, double* f_out) try {
  if (!F || !u) return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem_assemble_cubic: NULL argument");
  gmrfb_ctx* ctx = F->ctx;
  if (quad_degree != 1 && quad_degree != 2 && quad_degree != 4)
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem_assemble_cubic: quad_degree must be 1, 2 or 4");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (F->quad_degree != quad_degree) {
    const std::vector<double> r = tri_rule(quad_degree);
    GMRFB_CU(ctx, cudaStreamSynchronize(st));  // a previous call may still read the old table
    GMRFB_CU(ctx, F->d_quad.upload(r, st));
    GMRFB_CU(ctx, cudaStreamSynchronize(st));  // `r` is a pageable temporary
    F->quad_degree = quad_degree;
    F->nq = (int)(r.size() / 4);
  }
  if (!F->d_u.p) {
    GMRFB_CU(ctx, F->d_u.alloc((size_t)F->nn));
    GMRFB_CU(ctx, F->d_f.alloc((size_t)F->nn));
  }
  GMRFB_CU(ctx, cudaMemcpyAsync(F->d_u.p, u, F->nn * sizeof(double), cudaMemcpyDefault, st));  // host or device
  gmrfb_status rc = fem_upload_presc(F, prescribed);
  if (rc != GMRFB_OK) return rc;
  const uint8_t* pr = prescribed ? F->d_presc.p : nullptr;
  {
    ProfScope ps(ctx, PK_FEM, 0, 12.0 * 9 * F->nt + 8.0 * F->J.nnz, 0, 0);
    k_fem_cubic_J<<<(unsigned)((F->J.nnz + 255) / 256), 256, 0, st>>>(F->J.nnz, F->d_cptr.p, F->d_cidx.p, F->d_kloc.p,
                                                                     F->d_area.p, F->d_tris.p, F->d_u.p, F->J.d_rowidx.p,
                                                                     pr, stiffness_scale, F->d_quad.p, F->nq, F->J.d_val.p);
  }
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->J.d_val.p, F->J.d_tmap.p, F->J.nnz, F->J.d_tval.p, st));
  ctx->launches += 2;
  if (f_out) {
    k_fem_cubic_f<<<(unsigned)((F->nn + 255) / 256), 256, 0, st>>>(F->nn, F->d_diag.p, F->d_cptr.p, F->d_cidx.p, F->d_kloc.p,
                                                                  F->d_area.p, F->d_tris.p, F->d_u.p, pr, stiffness_scale,
                                                                  F->d_quad.p, F->nq, F->d_f.p);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches += 1;
    GMRFB_CU(ctx, cudaMemcpyAsync(f_out, F->d_f.p, F->nn * sizeof(double), cudaMemcpyDefault, st));
  }
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (J_out) *J_out = &F->J;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH
