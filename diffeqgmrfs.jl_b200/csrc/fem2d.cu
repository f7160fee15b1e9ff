// C ABI: finite elements on Lagrange triangles of order 1 or 2 with the cell values Ferrite computes for them
// (isoparametric geometry, `QuadratureRule{RefTriangle}(order + 1)`): the discretisations the reference's scripts
// actually run - `uniform_unit_square_discretization(...; element_order = 2)` (src/utils.jl:20-38) for the Darcy
// dataset loop and `generate_grid(QuadraticTriangle, ...)` (_research/elliptic_chen24.jl:118-122) for the elliptic
// Gauss-Newton solve.  fem.cu is the P1 special case with one coefficient per element; here
//   * assemble_darcy_diff_matrix   src/problems/darcy.jl:5-63: Ge[i, j] += coeff(x_q) grad phi_i . grad phi_j dOmega with
//     the coefficient looked up AT EVERY QUADRATURE POINT by nearest grid index (:39, src/datasets/darcy.jl:30-34), and
//     the load fe[i] += beta phi_i dOmega (:46);
//   * the element-lumped mass of src/spdes/shallow_water.jl:115 and the Matern prior ratio K' Mt^-1 K (:172-190), plus
//     the third power ratio K Mt^-1 K Mt^-1 K of the commented line :186 (smoothness 2 in two dimensions,
//     scripts/darcy/solve_darcy_gmrf-fem.jl:97) through a fixed-pattern sparse product;
//   * assemble_J_cube / assemble_J_diff_and_f / f_and_J   _research/elliptic_chen24.jl:180-285.
// Same gather form as fem.cu / fem1d.cu: the mesh is analysed once on the host (pattern, the element entries on every
// nonzero), the geometry (physical gradients and dOmega at every quadrature point of every element = what
// `reinit!(cellvalues, cell)` computes) once on the device, and every assembly is one thread per nonzero with a fixed
// summation order and no atomics.  Quadratic triangles are numbered as Ferrite's QuadraticTriangle: vertices 0, 1, 2,
// then the nodes on the edges (0,1), (1,2), (2,0).
#include <algorithm>
#include <cmath>
#include <memory>
#include <numeric>

#include "common.hpp"
#include "fem_pattern.hpp"
#include "handles.hpp"
#include "sparse_kernels.hpp"
#include "fem2d_kernels.cuh"

using namespace gmrfb;

struct gmrfb_postprec;
extern "C" gmrfb_status gmrfb_postprec_create(gmrfb_ctx*, const gmrfb_spm*, const gmrfb_spm*, gmrfb_postprec**);
extern "C" gmrfb_status gmrfb_postprec_destroy(gmrfb_postprec*);
extern "C" gmrfb_status gmrfb_postprec_compute(gmrfb_postprec*, double, const double*, const gmrfb_spm**);
struct gmrfb_spgemm;
extern "C" gmrfb_status gmrfb_spgemm_create(gmrfb_ctx*, const gmrfb_spm*, const gmrfb_spm*, gmrfb_spgemm**);
extern "C" gmrfb_status gmrfb_spgemm_destroy(gmrfb_spgemm*);
extern "C" gmrfb_status gmrfb_spgemm_compute(gmrfb_spgemm*, double, const double*, const gmrfb_spm**);

struct gmrfb_fem2d {
  gmrfb_ctx* ctx = nullptr;
  int64_t nn = 0, ne = 0;
  int order = 1, npe = 3, nq = 0, quad_degree = 0;
  gmrfb_spm G;  // stiffness of the last gmrfb_fem2d_stiffness call
  gmrfb_spm K;  // kappa^2 Mt + G of the last Matern call
  gmrfb_spm M;  // mass of the last gmrfb_fem2d_mass call
  gmrfb_spm J;  // tangent of the last gmrfb_fem2d_assemble_cubic call
  gmrfb_spm Z;  // empty n x n matrix (the "Q" of the plan that forms K' W K)
  std::vector<double> nodes;   // 2 nn (host): quadrature-point coordinates for the coefficient lookup
  std::vector<int32_t> conn;   // ne x npe (host)
  std::vector<double> shape;   // nq x npe (host copy of d_shape)
  DevBuf<int32_t> d_conn;
  DevBuf<double> d_shape;      // nq x npe: N_a(q)
  DevBuf<double> d_grad;       // ne x nq x npe x 2: physical gradient of N_a at quadrature point q of element t
  DevBuf<double> d_jxw;        // ne x nq: dOmega = weight |det J| / 2
  DevBuf<int64_t> d_cptr, d_diag;
  DevBuf<int32_t> d_cidx;      // element entry t npe^2 + i npe + j of every contribution (ascending per nonzero)
  DevBuf<int32_t> d_cellq;     // ne x nq: coefficient-grid cell of every quadrature point
  DevBuf<double> d_coeff;
  int64_t ncell = 0;
  DevBuf<uint8_t> d_presc;
  DevBuf<double> d_mlump, d_w, d_u, d_f;
  std::vector<double> mlump;   // host copy of the last lumped mass
  int mlump_kind = -1;
  gmrfb_postprec* matern_plan = nullptr;  // K' W K
  gmrfb_spgemm* matern_plan3 = nullptr;   // (K' W K) W K
};

namespace {

using namespace gmrfb::fem2d;

inline unsigned blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

gmrfb_status upload_presc(gmrfb_fem2d* F, const uint8_t* prescribed) {
  gmrfb_ctx* ctx = F->ctx;
  if (!prescribed) return GMRFB_OK;
  if (!F->d_presc.p) GMRFB_CU(ctx, F->d_presc.alloc((size_t)F->nn));
  GMRFB_CU(ctx, cudaMemcpyAsync(F->d_presc.p, prescribed, F->nn, cudaMemcpyDefault, ctx->stream));
  return GMRFB_OK;
}

// lumped mass of `kind` (1 row sums, 2 scaled diagonal; 0 = the default of the order) into d_mlump / mlump
gmrfb_status ensure_lumped(gmrfb_fem2d* F, int kind) {
  gmrfb_ctx* ctx = F->ctx;
  if (kind == 0) kind = F->order == 1 ? 1 : 2;
  if (F->mlump_kind == kind) return GMRFB_OK;
  cudaStream_t st = ctx->stream;
  if (!F->d_mlump.p) GMRFB_CU(ctx, F->d_mlump.alloc((size_t)F->nn));
  k_fem2d_lump<<<blocks(F->nn), 256, 0, st>>>(F->nn, F->npe, F->nq, kind, F->d_diag.p, F->d_cptr.p, F->d_cidx.p,
                                             F->d_shape.p, F->d_jxw.p, F->d_mlump.p);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches += 1;
  F->mlump.resize((size_t)F->nn);
  GMRFB_CU(ctx, cudaMemcpyAsync(F->mlump.data(), F->d_mlump.p, F->nn * sizeof(double), cudaMemcpyDeviceToHost, st));
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  F->mlump_kind = kind;
  return GMRFB_OK;
}

}  // namespace

extern "C" gmrfb_status gmrfb_fem2d_create(gmrfb_ctx* ctx, int32_t order, int64_t nnodes, const double* nodes,
                                           int64_t nelem, const int64_t* elems, int32_t base, int32_t quad_degree,
                                           gmrfb_fem2d** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_fem2d_create: ctx is NULL");
  if (!out || !nodes || !elems || nnodes <= 0 || nelem <= 0 || (base != 0 && base != 1) || (order != 1 && order != 2))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem2d_create: bad argument (order must be 1 or 2)");
  if (quad_degree == 0) quad_degree = order + 1;  // QuadratureRule{RefTriangle}(element_order + 1), src/utils.jl:30
  if (quad_degree < 1 || quad_degree > 4)
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem2d_create: quad_degree must be 0 (order + 1), 1, 2, 3 or 4");
  const int npe = order == 1 ? 3 : 6;
  if (nnodes > 2000000000 || nelem > 2000000000 / (npe * npe))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem2d_create: mesh too large");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_fem2d> F(new gmrfb_fem2d());
  F->ctx = ctx;
  F->nn = nnodes;
  F->ne = nelem;
  F->order = order;
  F->npe = npe;
  F->quad_degree = quad_degree;
  F->conn.resize((size_t)npe * nelem);
  for (int64_t k = 0; k < (int64_t)npe * nelem; k++) {
    const int64_t v = elems[k] - base;
    if (v < 0 || v >= nnodes) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem2d_create: node index out of range");
    F->conn[k] = (int32_t)v;
  }
  F->nodes.assign(nodes, nodes + 2 * nnodes);
  std::vector<double> dref, wq;
  const int nq = build_tables(order, quad_degree, F->shape, dref, wq);
  F->nq = nq;
  ElementPattern P;
  if (!build_element_pattern(nnodes, nelem, npe, F->conn.data(), P))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem2d_create: a node belongs to no element");
  gmrfb_status rc;
  for (gmrfb_spm* S : {&F->G, &F->K, &F->M, &F->J})
    if ((rc = spm_build(ctx, S, nnodes, nnodes, P.colptr.data(), P.rowval.data(), nullptr, 0)) != GMRFB_OK) return rc;
  {
    std::vector<int64_t> zp((size_t)nnodes + 1, 0);
    if ((rc = spm_build(ctx, &F->Z, nnodes, nnodes, zp.data(), nullptr, nullptr, 0)) != GMRFB_OK) return rc;
  }
  F->G.owned_by_plan = F->K.owned_by_plan = F->M.owned_by_plan = F->J.owned_by_plan = F->Z.owned_by_plan = true;
  cudaStream_t st = ctx->stream;
  DevBuf<double> d_nodes, d_dref, d_wq;
  DevBuf<int> d_bad;
  GMRFB_CU(ctx, d_nodes.upload(F->nodes, st));
  GMRFB_CU(ctx, d_dref.upload(dref, st));
  GMRFB_CU(ctx, d_wq.upload(wq, st));
  GMRFB_CU(ctx, F->d_conn.upload(F->conn, st));
  GMRFB_CU(ctx, F->d_shape.upload(F->shape, st));
  GMRFB_CU(ctx, F->d_cptr.upload(P.cptr, st));
  GMRFB_CU(ctx, F->d_cidx.upload(P.cidx, st));
  GMRFB_CU(ctx, F->d_diag.upload(P.diag, st));
  GMRFB_CU(ctx, F->d_grad.alloc((size_t)nelem * nq * npe * 2));
  GMRFB_CU(ctx, F->d_jxw.alloc((size_t)nelem * nq));
  GMRFB_CU(ctx, d_bad.alloc(1));
  GMRFB_CU(ctx, cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
  k_fem2d_geom<<<blocks(nelem * nq), 256, 0, st>>>(nelem, npe, nq, d_nodes.p, F->d_conn.p, d_dref.p, d_wq.p, F->d_grad.p,
                                                  F->d_jxw.p, d_bad.p);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches += 1;
  int bad = 0;
  GMRFB_CU(ctx, cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (bad) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem2d_create: degenerate element (det J = 0 at a quadrature point)");
  *out = F.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem2d_destroy(gmrfb_fem2d* F) try {
  if (!F) return GMRFB_OK;
  cudaSetDevice(F->ctx->device);
  cudaStreamSynchronize(F->ctx->stream);
  if (F->matern_plan3) gmrfb_spgemm_destroy(F->matern_plan3);
  if (F->matern_plan) gmrfb_postprec_destroy(F->matern_plan);
  delete F;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem2d_info(gmrfb_fem2d* F, int32_t* order, int32_t* nodes_per_element, int32_t* nquad,
                                         int64_t* nnz) try {
  if (!F) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_fem2d_info: NULL handle");
  if (order) *order = F->order;
  if (nodes_per_element) *nodes_per_element = F->npe;
  if (nquad) *nquad = F->nq;
  if (nnz) *nnz = F->G.nnz;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem2d_set_coeff_grid(gmrfb_fem2d* F, int64_t gx, const double* x_coords, int64_t gy,
                                                   const double* y_coords) try {
  if (!F || !x_coords || !y_coords || gx <= 0 || gy <= 0)
    return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem2d_set_coeff_grid: bad argument");
  gmrfb_ctx* ctx = F->ctx;
  if (gx * gy > 2000000000) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem2d_set_coeff_grid: grid too large");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::vector<int32_t> cell((size_t)F->ne * F->nq);
  quad_point_cells(F->ne, F->npe, F->nq, F->conn.data(), F->nodes.data(), F->shape.data(), gx, x_coords, gy, y_coords,
                   cell.data());
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));  // a previous assembly may still read the old cells
  GMRFB_CU(ctx, F->d_cellq.upload(cell, ctx->stream));
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));  // `cell` is a pageable temporary
  F->ncell = gx * gy;
  GMRFB_CU(ctx, F->d_coeff.alloc((size_t)F->ncell));
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem2d_stiffness(gmrfb_fem2d* F, const double* coeff_grid, const uint8_t* prescribed,
                                              double beta, const gmrfb_spm** G_out, double* load_out) try {
  if (!F) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_fem2d_stiffness: NULL handle");
  gmrfb_ctx* ctx = F->ctx;
  if (coeff_grid && !F->d_cellq.p)
    return fail(ctx, GMRFB_ERR_STATE, "gmrfb_fem2d_stiffness: call gmrfb_fem2d_set_coeff_grid before passing a coefficient grid");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (coeff_grid)  // host or device pointer (unified addressing)
    GMRFB_CU(ctx, cudaMemcpyAsync(F->d_coeff.p, coeff_grid, F->ncell * sizeof(double), cudaMemcpyDefault, st));
  gmrfb_status rc = upload_presc(F, prescribed);
  if (rc != GMRFB_OK) return rc;
  const uint8_t* pr = prescribed ? F->d_presc.p : nullptr;
  {
    ProfScope ps(ctx, PK_FEM, 0, (8.0 * 2 * F->npe + 12.0) * F->nq * F->ne + 8.0 * F->G.nnz, 0, 0);
    k_fem2d_stiffness<<<blocks(F->G.nnz), 256, 0, st>>>(F->G.nnz, F->npe, F->nq, F->d_cptr.p, F->d_cidx.p, F->d_grad.p,
                                                       F->d_jxw.p, F->d_cellq.p, coeff_grid ? F->d_coeff.p : nullptr,
                                                       F->G.d_rowidx.p, F->d_diag.p, pr, F->G.d_val.p);
  }
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->G.d_val.p, F->G.d_tmap.p, F->G.nnz, F->G.d_tval.p, st));
  ctx->launches += 2;
  if (load_out) {
    if (!F->d_f.p) {
      GMRFB_CU(ctx, F->d_u.alloc((size_t)F->nn));
      GMRFB_CU(ctx, F->d_f.alloc((size_t)F->nn));
    }
    k_fem2d_load<<<blocks(F->nn), 256, 0, st>>>(F->nn, F->npe, F->nq, F->d_diag.p, F->d_cptr.p, F->d_cidx.p, F->d_shape.p,
                                               F->d_jxw.p, pr, beta, F->d_f.p);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches += 1;
    GMRFB_CU(ctx, cudaMemcpyAsync(load_out, F->d_f.p, F->nn * sizeof(double), cudaMemcpyDefault, st));
  }
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (G_out) *G_out = &F->G;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem2d_mass(gmrfb_fem2d* F, int32_t lumping, const gmrfb_spm** M_out, double* lumped_out) try {
  if (!F) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_fem2d_mass: NULL handle");
  gmrfb_ctx* ctx = F->ctx;
  if (lumping < 0 || lumping > 3) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem2d_mass: lumping must be 0 ... 3");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (lumping == 0) {
    k_fem2d_mass<<<blocks(F->M.nnz), 256, 0, st>>>(F->M.nnz, F->npe, F->nq, F->d_cptr.p, F->d_cidx.p, F->d_shape.p,
                                                  F->d_jxw.p, F->M.d_val.p);
  } else {
    gmrfb_status rc = ensure_lumped(F, lumping == 3 ? 0 : lumping);
    if (rc != GMRFB_OK) return rc;
    k_fem2d_set_lumped<<<blocks(F->M.nnz), 256, 0, st>>>(F->M.nnz, F->M.d_rowidx.p, F->d_diag.p, F->d_mlump.p, F->M.d_val.p);
    if (lumped_out) std::copy(F->mlump.begin(), F->mlump.end(), lumped_out);
  }
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->M.d_val.p, F->M.d_tmap.p, F->M.nnz, F->M.d_tval.p, st));
  ctx->launches += 2;
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (M_out) *M_out = &F->M;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem2d_matern_precision(gmrfb_fem2d* F, double kappa, double ratio, int32_t alpha,
                                                     const uint8_t* prescribed, double prescribed_mass,
                                                     const gmrfb_spm** Q_out) try {
  if (!F || !Q_out) return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem2d_matern_precision: NULL argument");
  gmrfb_ctx* ctx = F->ctx;
  if (!(kappa > 0) || !(ratio > 0) || (prescribed && !(prescribed_mass > 0)) || (alpha != 2 && alpha != 3))
    return fail(ctx, GMRFB_ERR_INVALID,
                "gmrfb_fem2d_matern_precision: kappa, ratio and prescribed_mass must be positive, alpha 2 or 3");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  gmrfb_status rc = upload_presc(F, prescribed);
  if (rc != GMRFB_OK) return rc;
  if ((rc = ensure_lumped(F, 0)) != GMRFB_OK) return rc;
  if (!F->d_w.p) GMRFB_CU(ctx, F->d_w.alloc((size_t)F->nn));
  // K = kappa^2 Mt + G (unit coefficient; the reference only touches the diagonal entries of prescribed dofs)
  k_fem2d_stiffness<<<blocks(F->K.nnz), 256, 0, st>>>(F->K.nnz, F->npe, F->nq, F->d_cptr.p, F->d_cidx.p, F->d_grad.p,
                                                     F->d_jxw.p, nullptr, nullptr, F->K.d_rowidx.p, F->d_diag.p, nullptr,
                                                     F->K.d_val.p);
  GMRFB_CU(ctx, cudaGetLastError());
  // alpha 2: w = ratio / Mt;  alpha 3: w = 1 / Mt for both products, the ratio applied by the last one
  k_fem2d_matern_k<<<blocks(F->nn), 256, 0, st>>>(F->nn, F->d_diag.p, F->d_mlump.p, prescribed ? F->d_presc.p : nullptr,
                                                 prescribed_mass, kappa * kappa, alpha == 2 ? ratio : 1.0, F->K.d_val.p,
                                                 F->d_w.p);
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->K.d_val.p, F->K.d_tmap.p, F->K.nnz, F->K.d_tval.p, st));
  ctx->launches += 3;
  if (!F->matern_plan) {  // symbolic work, once per mesh
    rc = gmrfb_postprec_create(ctx, &F->Z, &F->K, &F->matern_plan);
    if (rc != GMRFB_OK) return rc;
  }
  const gmrfb_spm* Q2 = nullptr;
  rc = gmrfb_postprec_compute(F->matern_plan, 0.0, F->d_w.p, &Q2);  // K' W K
  if (rc != GMRFB_OK || alpha == 2) {
    *Q_out = Q2;
    return rc;
  }
  if (!F->matern_plan3) {
    rc = gmrfb_spgemm_create(ctx, Q2, &F->K, &F->matern_plan3);
    if (rc != GMRFB_OK) return rc;
  }
  return gmrfb_spgemm_compute(F->matern_plan3, ratio, F->d_w.p, Q_out);  // ratio (K W K) W K
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem2d_assemble_cubic(gmrfb_fem2d* F, const double* u, double stiffness_scale,
                                                   const uint8_t* prescribed, const gmrfb_spm** J_out, double* f_out) try {
  if (!F || !u) return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem2d_assemble_cubic: NULL argument");
  gmrfb_ctx* ctx = F->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (!F->d_u.p) {
    GMRFB_CU(ctx, F->d_u.alloc((size_t)F->nn));
    GMRFB_CU(ctx, F->d_f.alloc((size_t)F->nn));
  }
  GMRFB_CU(ctx, cudaMemcpyAsync(F->d_u.p, u, F->nn * sizeof(double), cudaMemcpyDefault, st));  // host or device
  gmrfb_status rc = upload_presc(F, prescribed);
  if (rc != GMRFB_OK) return rc;
  const uint8_t* pr = prescribed ? F->d_presc.p : nullptr;
  {
    ProfScope ps(ctx, PK_FEM, 0, (8.0 * 2 * F->npe + 8.0) * F->nq * F->ne + 8.0 * F->J.nnz, 0, 0);
    k_fem2d_cubic_J<<<blocks(F->J.nnz), 256, 0, st>>>(F->J.nnz, F->npe, F->nq, F->d_cptr.p, F->d_cidx.p, F->d_conn.p,
                                                     F->d_shape.p, F->d_grad.p, F->d_jxw.p, F->d_u.p, F->J.d_rowidx.p, pr,
                                                     stiffness_scale, F->J.d_val.p);
  }
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->J.d_val.p, F->J.d_tmap.p, F->J.nnz, F->J.d_tval.p, st));
  ctx->launches += 2;
  if (f_out) {
    k_fem2d_cubic_f<<<blocks(F->nn), 256, 0, st>>>(F->nn, F->npe, F->nq, F->d_diag.p, F->d_cptr.p, F->d_cidx.p, F->d_conn.p,
                                                  F->d_shape.p, F->d_grad.p, F->d_jxw.p, F->d_u.p, pr, stiffness_scale,
                                                  F->d_f.p);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches += 1;
    GMRFB_CU(ctx, cudaMemcpyAsync(f_out, F->d_f.p, F->nn * sizeof(double), cudaMemcpyDefault, st));
  }
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (J_out) *J_out = &F->J;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH
