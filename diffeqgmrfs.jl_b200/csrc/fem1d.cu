// C ABI: finite elements on a 1-D mesh (Lagrange lines of order 1 or 2) for the Burgers Gauss-Newton loop:
//   * assemble_burgers_mass_diffusion_matrices(disc; lumping)      src/problems/burgers.jl:61-98
//   * assemble_burgers_advection_matrix(disc, cur_weights)         src/problems/burgers.jl:5-59   (u u_x and its tangent)
//   * the space-time tangent the script builds from them at every Gauss-Newton step,
//       J = J_static + dt J_adv,  f = J_static w + dt f_adv,  J_static = M_{t+1} - M_t + dt nu G_{t+1}
//                                                                  scripts/burgers/solve_burgers_gmrf-fem.jl:115-142
//     (the reference loops over the time steps, assembles one spatial matrix per step with a scalar cell loop, lifts
//     each to space-time and concatenates; here one kernel writes all values of the fixed space-time pattern).
// Same gather form as fem.cu: the mesh is analysed once on the host, every assembly is one thread per nonzero with a
// fixed summation order.  Quadratic lines are numbered (left, right, middle) as Ferrite's QuadraticLine.
#include <cmath>
#include <memory>

#include "common.hpp"
#include "fem_pattern.hpp"
#include "handles.hpp"

using namespace gmrfb;

struct gmrfb_fem1d {
  gmrfb_ctx* ctx = nullptr;
  int64_t nn = 0, ne = 0;
  int npe = 2, nq = 2;
  gmrfb_spm M, G, A;  // spatial pattern: mass, stiffness, advection tangent of the last call
  std::unique_ptr<gmrfb_spm> J;  // space-time tangent ((nt - 1) nn x nt nn) of the last gmrfb_fem1d_spacetime_tangent call
  int64_t nt = 0;     // number of time steps J was built for (0: not built)
  bool static_built = false, static_lumped = false;
  std::vector<uint8_t> static_presc;  // mask the static matrices were assembled with
  DevBuf<int32_t> d_elems;  // ne x npe
  DevBuf<double> d_shape;   // nq x (2 npe + 1): N_k(xi_q), dN_k/dxi(xi_q), weight
  DevBuf<double> d_jac;     // ne x nq: dx/dxi at the quadrature points
  DevBuf<int64_t> d_cptr, d_diag;
  DevBuf<int32_t> d_cidx;
  DevBuf<int32_t> d_colnz;  // column of every (CSC) nonzero of the spatial pattern
  DevBuf<uint8_t> d_presc;
  DevBuf<double> d_u, d_v, d_mlump;
  // space-time pattern: per J nonzero the spatial nonzero it comes from and which part (0: -M of block (t, t);
  // 1: M + dt nu G + dt A(u_{t+1}) of block (t, t+1)); column block s owns the nonzeros [jblk[s], jblk[s + 1])
  DevBuf<int32_t> d_jmap;
  DevBuf<int64_t> d_jblk;
  std::vector<int64_t> jblk;
  DevBuf<double> d_w, d_f;
};

namespace {

constexpr int MAXPE = 3, MAXQ = 4;

__global__ void k_fem1d_jac(int64_t ne, int npe, int nq, const double* __restrict__ xe, const double* __restrict__ shape,
                            double* __restrict__ jac) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ne * nq) return;
  const int64_t e = t / nq;
  const int q = (int)(t % nq);
  const double* sh = shape + (size_t)q * (2 * npe + 1);
  double j = 0.0;
  for (int k = 0; k < npe; k++) j += xe[e * npe + k] * sh[npe + k];
  jac[t] = j;
}

// M[k] = sum phi_i phi_j dOmega, G[k] = sum phi_i' phi_j' dOmega over the element entries on nonzero k; rows and
// columns of prescribed dofs are zero (src/problems/burgers.jl:88-93)
__global__ void k_fem1d_static(int64_t nnz, int npe, int nq, const int64_t* __restrict__ cptr,
                               const int32_t* __restrict__ cidx, const double* __restrict__ shape,
                               const double* __restrict__ jac, const int32_t* __restrict__ rowidx,
                               const int32_t* __restrict__ colidx_of_nz, const uint8_t* __restrict__ presc,
                               double* __restrict__ Mv, double* __restrict__ Gv) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  if (presc && (presc[rowidx[k]] || presc[colidx_of_nz[k]])) {
    Mv[k] = 0.0;
    Gv[k] = 0.0;
    return;
  }
  const int npe2 = npe * npe, ld = 2 * npe + 1;
  double m = 0.0, g = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
    const int32_t c = cidx[p];
    const int32_t e = c / npe2, i = (c % npe2) / npe, j = c % npe;
    for (int q = 0; q < nq; q++) {
      const double* sh = shape + (size_t)q * ld;
      const double jq = jac[(int64_t)e * nq + q], dO = sh[2 * npe] * jq;
      m += sh[i] * sh[j] * dO;
      g += (sh[npe + i] / jq) * (sh[npe + j] / jq) * dO;
    }
  }
  Mv[k] = m;
  Gv[k] = g;
}

// row sums of the consistent mass (through the diagonal's contribution list: one entry per element at the node)
__global__ void k_fem1d_lump(int64_t nn, int npe, int nq, const int64_t* __restrict__ diag, const int64_t* __restrict__ cptr,
                             const int32_t* __restrict__ cidx, const double* __restrict__ shape,
                             const double* __restrict__ jac, const int32_t* __restrict__ elems,
                             const uint8_t* __restrict__ presc, double* __restrict__ ml) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  if (presc && presc[i]) {
    ml[i] = 0.0;
    return;
  }
  const int npe2 = npe * npe, ld = 2 * npe + 1;
  double m = 0.0;
  for (int64_t p = cptr[diag[i]]; p < cptr[diag[i] + 1]; p++) {
    const int32_t c = cidx[p];
    const int32_t e = c / npe2, a = c % npe;
    for (int q = 0; q < nq; q++) {
      const double* sh = shape + (size_t)q * ld;
      const double dO = sh[2 * npe] * jac[(int64_t)e * nq + q];
      for (int j = 0; j < npe; j++)
        if (!(presc && presc[elems[e * npe + j]])) m += sh[a] * sh[j] * dO;
    }
  }
  ml[i] = m;
}

__global__ void k_fem1d_set_lumped(int64_t nnz, const int32_t* __restrict__ rowidx, const int64_t* __restrict__ diag,
                                   const double* __restrict__ ml, double* __restrict__ Mv) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int32_t r = rowidx[k];
  Mv[k] = (k == diag[r]) ? ml[r] : 0.0;
}

// advection tangent entry of element entry (e, i, j) for the iterate u (length nn):
//   sum_q phi_i (phi_j u_x + u phi_j') dOmega      (src/problems/burgers.jl:40-48)
__device__ __forceinline__ double adv_entry(int npe, int nq, const double* __restrict__ shape, const double* __restrict__ jac,
                                            const int32_t* __restrict__ elems, const double* __restrict__ u, int32_t e, int i,
                                            int j) {
  const int ld = 2 * npe + 1;
  double w[MAXPE];
  for (int k = 0; k < npe; k++) w[k] = u[elems[e * npe + k]];
  double acc = 0.0;
  for (int q = 0; q < nq; q++) {
    const double* sh = shape + (size_t)q * ld;
    const double jq = jac[(int64_t)e * nq + q], dO = sh[2 * npe] * jq;
    double uq = 0.0, ux = 0.0;
    for (int k = 0; k < npe; k++) {
      uq += sh[k] * w[k];
      ux += sh[npe + k] * w[k];
    }
    ux /= jq;
    acc += sh[i] * (sh[j] * ux + uq * (sh[npe + j] / jq)) * dO;
  }
  return acc;
}

// sum_q phi_a u u_x dOmega of element e                                     (src/problems/burgers.jl:49)
__device__ __forceinline__ double adv_load(int npe, int nq, const double* __restrict__ shape, const double* __restrict__ jac,
                                           const int32_t* __restrict__ elems, const double* __restrict__ u, int32_t e, int a) {
  const int ld = 2 * npe + 1;
  double w[MAXPE];
  for (int k = 0; k < npe; k++) w[k] = u[elems[e * npe + k]];
  double acc = 0.0;
  for (int q = 0; q < nq; q++) {
    const double* sh = shape + (size_t)q * ld;
    const double jq = jac[(int64_t)e * nq + q], dO = sh[2 * npe] * jq;
    double uq = 0.0, ux = 0.0;
    for (int k = 0; k < npe; k++) {
      uq += sh[k] * w[k];
      ux += sh[npe + k] * w[k];
    }
    acc += sh[a] * uq * (ux / jq) * dO;
  }
  return acc;
}

__global__ void k_fem1d_adv(int64_t nnz, int npe, int nq, const int64_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                            const double* __restrict__ shape, const double* __restrict__ jac,
                            const int32_t* __restrict__ elems, const double* __restrict__ u,
                            const int32_t* __restrict__ rowidx, const int32_t* __restrict__ colidx_of_nz,
                            const uint8_t* __restrict__ presc, double* __restrict__ Av) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  if (presc && (presc[rowidx[k]] || presc[colidx_of_nz[k]])) {
    Av[k] = 0.0;
    return;
  }
  const int npe2 = npe * npe;
  double a = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
    const int32_t c = cidx[p];
    a += adv_entry(npe, nq, shape, jac, elems, u, c / npe2, (c % npe2) / npe, c % npe);
  }
  Av[k] = a;
}

__global__ void k_fem1d_adv_v(int64_t nn, int npe, int nq, const int64_t* __restrict__ diag, const int64_t* __restrict__ cptr,
                              const int32_t* __restrict__ cidx, const double* __restrict__ shape,
                              const double* __restrict__ jac, const int32_t* __restrict__ elems,
                              const double* __restrict__ u, const uint8_t* __restrict__ presc, double* __restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  if (presc && presc[i]) {
    v[i] = 0.0;
    return;
  }
  const int npe2 = npe * npe;
  double a = 0.0;
  for (int64_t p = cptr[diag[i]]; p < cptr[diag[i] + 1]; p++) {
    const int32_t c = cidx[p];
    a += adv_load(npe, nq, shape, jac, elems, u, c / npe2, c % npe);
  }
  v[i] = a;
}

// all values of the space-time tangent: blockIdx.y = column block s (time step s), x over its nonzeros
__global__ void k_fem1d_st_J(int64_t nn, int npe, int nq, const int64_t* __restrict__ jblk, const int32_t* __restrict__ jmap,
                             const int64_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                             const double* __restrict__ shape, const double* __restrict__ jac,
                             const int32_t* __restrict__ elems, const double* __restrict__ w,
                             const int32_t* __restrict__ rowidx, const int32_t* __restrict__ colidx_of_nz,
                             const uint8_t* __restrict__ presc, const double* __restrict__ Mv,
                             const double* __restrict__ Gv, double dt, double nu, double* __restrict__ Jv) {
  const int64_t s = blockIdx.y;
  const int64_t kj = jblk[s] + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kj >= jblk[s + 1]) return;
  const int32_t code = jmap[kj];
  const int64_t k = code >> 1;
  if ((code & 1) == 0) {
    Jv[kj] = -Mv[k];
    return;
  }
  double a = 0.0;
  if (!(presc && (presc[rowidx[k]] || presc[colidx_of_nz[k]]))) {
    const int npe2 = npe * npe;
    const double* u = w + s * nn;
    for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
      const int32_t c = cidx[p];
      a += adv_entry(npe, nq, shape, jac, elems, u, c / npe2, (c % npe2) / npe, c % npe);
    }
  }
  Jv[kj] = Mv[k] + dt * nu * Gv[k] + dt * a;
}

// f[(t - 1) nn + i] = sum_j (M + dt nu G)[i, j] w_t[j] - sum_j M[i, j] w_{t-1}[j] + dt v_i(w_t),  t = 1 .. nt - 1
// (M and G are symmetric: row i is read as column i)
__global__ void k_fem1d_st_f(int64_t nn, int npe, int nq, const int64_t* __restrict__ colptr,
                             const int32_t* __restrict__ rowidx, const int64_t* __restrict__ diag,
                             const int64_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                             const double* __restrict__ shape, const double* __restrict__ jac,
                             const int32_t* __restrict__ elems, const double* __restrict__ w,
                             const uint8_t* __restrict__ presc, const double* __restrict__ Mv,
                             const double* __restrict__ Gv, double dt, double nu, double* __restrict__ f) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t t = (int64_t)blockIdx.y + 1;
  if (i >= nn) return;
  const double* wt = w + t * nn;
  const double* wp = w + (t - 1) * nn;
  double acc = 0.0;
  for (int64_t k = colptr[i]; k < colptr[i + 1]; k++) {
    const int32_t r = rowidx[k];
    acc += (Mv[k] + dt * nu * Gv[k]) * wt[r] - Mv[k] * wp[r];
  }
  if (!(presc && presc[i])) {
    const int npe2 = npe * npe;
    double a = 0.0;
    for (int64_t p = cptr[diag[i]]; p < cptr[diag[i] + 1]; p++) {
      const int32_t c = cidx[p];
      a += adv_load(npe, nq, shape, jac, elems, wt, c / npe2, c % npe);
    }
    acc += dt * a;
  }
  f[(t - 1) * nn + i] = acc;
}

// Gauss-Legendre rule on [-1, 1] with n points (Newton iteration on P_n; n <= MAXQ)
void gauss_legendre(int n, double* xi, double* wq) {
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < n; k++) {
    double x = std::cos(pi * (k + 0.75) / (n + 0.5));
    double dp = 1.0;
    for (int it = 0; it < 100; it++) {
      double p0 = 1.0, p1 = x;
      for (int j = 2; j <= n; j++) {
        const double p2 = ((2.0 * j - 1.0) * x * p1 - (j - 1.0) * p0) / j;
        p0 = p1;
        p1 = p2;
      }
      dp = n * (x * p1 - p0) / (x * x - 1.0);
      const double dx = p1 / dp;
      x -= dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    xi[n - 1 - k] = x;  // ascending
    wq[n - 1 - k] = 2.0 / ((1.0 - x * x) * dp * dp);
  }
}

gmrfb_status upload_presc(gmrfb_fem1d* F, const uint8_t* prescribed) {
  if (!prescribed) return GMRFB_OK;
  gmrfb_ctx* ctx = F->ctx;
  if (!F->d_presc.p) GMRFB_CU(ctx, F->d_presc.alloc((size_t)F->nn));
  GMRFB_CU(ctx, cudaMemcpyAsync(F->d_presc.p, prescribed, F->nn, cudaMemcpyDefault, ctx->stream));
  return GMRFB_OK;
}

// consistent (or lumped) mass and stiffness into F->M / F->G; cached for an unchanged (mask, lumping)
gmrfb_status build_static(gmrfb_fem1d* F, bool lumping, const uint8_t* prescribed) {
  gmrfb_ctx* ctx = F->ctx;
  cudaStream_t st = ctx->stream;
  std::vector<uint8_t> mask;
  if (prescribed) {
    mask.resize((size_t)F->nn);
    GMRFB_CU(ctx, cudaMemcpyAsync(mask.data(), prescribed, F->nn, cudaMemcpyDefault, st));
    GMRFB_CU(ctx, cudaStreamSynchronize(st));
    for (auto& b : mask) b = b ? 1 : 0;
  }
  if (F->static_built && F->static_lumped == lumping && F->static_presc == mask) return GMRFB_OK;
  gmrfb_status rc = upload_presc(F, prescribed);
  if (rc != GMRFB_OK) return rc;
  const uint8_t* pr = prescribed ? F->d_presc.p : nullptr;
  const unsigned gb = (unsigned)((F->M.nnz + 255) / 256);
  k_fem1d_static<<<gb, 256, 0, st>>>(F->M.nnz, F->npe, F->nq, F->d_cptr.p, F->d_cidx.p, F->d_shape.p, F->d_jac.p,
                                     F->M.d_rowidx.p, F->d_colnz.p, pr, F->M.d_val.p, F->G.d_val.p);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches += 1;
  if (lumping) {
    if (!F->d_mlump.p) GMRFB_CU(ctx, F->d_mlump.alloc((size_t)F->nn));
    k_fem1d_lump<<<(unsigned)((F->nn + 255) / 256), 256, 0, st>>>(F->nn, F->npe, F->nq, F->d_diag.p, F->d_cptr.p, F->d_cidx.p,
                                                                 F->d_shape.p, F->d_jac.p, F->d_elems.p, pr, F->d_mlump.p);
    GMRFB_CU(ctx, cudaGetLastError());
    k_fem1d_set_lumped<<<gb, 256, 0, st>>>(F->M.nnz, F->M.d_rowidx.p, F->d_diag.p, F->d_mlump.p, F->M.d_val.p);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches += 2;
  }
  GMRFB_CU(ctx, launch_gather_values(F->M.d_val.p, F->M.d_tmap.p, F->M.nnz, F->M.d_tval.p, st));
  GMRFB_CU(ctx, launch_gather_values(F->G.d_val.p, F->G.d_tmap.p, F->G.nnz, F->G.d_tval.p, st));
  ctx->launches += 2;
  F->static_built = true;
  F->static_lumped = lumping;
  F->static_presc = mask;
  return GMRFB_OK;
}

}  // namespace

extern "C" gmrfb_status gmrfb_fem1d_create(gmrfb_ctx* ctx, int64_t nnodes, int64_t nelem, const int64_t* elems,
                                           const double* elem_x, int32_t order, int32_t base, int32_t nquad,
                                           gmrfb_fem1d** out) try {
  if (!ctx) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_fem1d_create: ctx is NULL");
  if (!out || !elem_x || !elems || nnodes <= 0 || nelem <= 0 || (base != 0 && base != 1) || (order != 1 && order != 2))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem1d_create: bad argument (order must be 1 or 2)");
  if (nquad == 0) nquad = order + 1;  // QuadratureRule{RefLine}(element_order + 1), src/utils.jl:45
  if (nquad < 1 || nquad > MAXQ) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem1d_create: nquad must be 1..4");
  if (nnodes > 2000000000 || nelem > 200000000) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem1d_create: mesh too large");
  *out = nullptr;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<gmrfb_fem1d> F(new gmrfb_fem1d());
  F->ctx = ctx;
  F->nn = nnodes;
  F->ne = nelem;
  const int npe = F->npe = order + 1;
  F->nq = nquad;
  std::vector<int32_t> el((size_t)npe * nelem);
  for (int64_t k = 0; k < npe * nelem; k++) {
    const int64_t v = elems[k] - base;
    if (v < 0 || v >= nnodes) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem1d_create: node index out of range");
    el[k] = (int32_t)v;
  }
  ElementPattern P;
  if (!build_element_pattern(nnodes, nelem, npe, el.data(), P))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem1d_create: a node belongs to no element");
  gmrfb_status rc;
  for (gmrfb_spm* S : {&F->M, &F->G, &F->A}) {
    if ((rc = spm_build(ctx, S, nnodes, nnodes, P.colptr.data(), P.rowval.data(), nullptr, 0)) != GMRFB_OK) return rc;
    S->owned_by_plan = true;  // borrowed views: not destroyed by the caller
  }
  // shape table: N_k, dN_k/dxi at the Gauss points and the weights
  double xi[MAXQ], wq[MAXQ];
  gauss_legendre(nquad, xi, wq);
  std::vector<double> shape((size_t)nquad * (2 * npe + 1));
  for (int q = 0; q < nquad; q++) {
    double* sh = shape.data() + (size_t)q * (2 * npe + 1);
    const double s = xi[q];
    if (npe == 2) {
      sh[0] = 0.5 * (1 - s), sh[1] = 0.5 * (1 + s);
      sh[2] = -0.5, sh[3] = 0.5;
    } else {
      sh[0] = 0.5 * s * (s - 1), sh[1] = 0.5 * s * (s + 1), sh[2] = 1 - s * s;
      sh[3] = s - 0.5, sh[4] = s + 0.5, sh[5] = -2 * s;
    }
    sh[2 * npe] = wq[q];
  }
  cudaStream_t st = ctx->stream;
  DevBuf<double> d_x;
  {
    std::vector<double> xv(elem_x, elem_x + (size_t)npe * nelem);
    GMRFB_CU(ctx, d_x.upload(xv, st));
    GMRFB_CU(ctx, cudaStreamSynchronize(st));
  }
  GMRFB_CU(ctx, F->d_elems.upload(el, st));
  GMRFB_CU(ctx, F->d_shape.upload(shape, st));
  GMRFB_CU(ctx, F->d_cptr.upload(P.cptr, st));
  GMRFB_CU(ctx, F->d_cidx.upload(P.cidx, st));
  GMRFB_CU(ctx, F->d_diag.upload(P.diag, st));
  {
    std::vector<int32_t> colnz(P.rowval.size());
    for (int64_t c = 0; c < nnodes; c++)
      for (int64_t k = P.colptr[c]; k < P.colptr[c + 1]; k++) colnz[k] = (int32_t)c;
    GMRFB_CU(ctx, F->d_colnz.upload(colnz, st));
  }
  GMRFB_CU(ctx, F->d_jac.alloc((size_t)nelem * nquad));
  k_fem1d_jac<<<(unsigned)((nelem * nquad + 255) / 256), 256, 0, st>>>(nelem, npe, nquad, d_x.p, F->d_shape.p, F->d_jac.p);
  GMRFB_CU(ctx, cudaGetLastError());
  ctx->launches += 1;
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  {  // an inverted or degenerate element would silently give a wrong matrix
    std::vector<double> jac((size_t)nelem * nquad);
    GMRFB_CU(ctx, cudaMemcpy(jac.data(), F->d_jac.p, jac.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (double j : jac)
      if (!(j > 0)) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem1d_create: element with non-positive Jacobian");
  }
  *out = F.release();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem1d_destroy(gmrfb_fem1d* F) try {
  if (!F) return GMRFB_OK;
  cudaSetDevice(F->ctx->device);
  cudaStreamSynchronize(F->ctx->stream);
  delete F;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem1d_mass_stiffness(gmrfb_fem1d* F, int32_t lumping, const uint8_t* prescribed,
                                                   const gmrfb_spm** M_out, const gmrfb_spm** G_out) try {
  if (!F) return fail(nullptr, GMRFB_ERR_INVALID, "gmrfb_fem1d_mass_stiffness: NULL handle");
  gmrfb_ctx* ctx = F->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  gmrfb_status rc = build_static(F, lumping != 0, prescribed);
  if (rc != GMRFB_OK) return rc;
  GMRFB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (M_out) *M_out = &F->M;
  if (G_out) *G_out = &F->G;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem1d_advection(gmrfb_fem1d* F, const double* u, const uint8_t* prescribed,
                                              const gmrfb_spm** A_out, double* v_out) try {
  if (!F || !u) return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem1d_advection: NULL argument");
  gmrfb_ctx* ctx = F->ctx;
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (!F->d_u.p) {
    GMRFB_CU(ctx, F->d_u.alloc((size_t)F->nn));
    GMRFB_CU(ctx, F->d_v.alloc((size_t)F->nn));
  }
  GMRFB_CU(ctx, cudaMemcpyAsync(F->d_u.p, u, F->nn * sizeof(double), cudaMemcpyDefault, st));
  gmrfb_status rc = upload_presc(F, prescribed);
  if (rc != GMRFB_OK) return rc;
  const uint8_t* pr = prescribed ? F->d_presc.p : nullptr;
  k_fem1d_adv<<<(unsigned)((F->A.nnz + 255) / 256), 256, 0, st>>>(F->A.nnz, F->npe, F->nq, F->d_cptr.p, F->d_cidx.p,
                                                                 F->d_shape.p, F->d_jac.p, F->d_elems.p, F->d_u.p,
                                                                 F->A.d_rowidx.p, F->d_colnz.p, pr, F->A.d_val.p);
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->A.d_val.p, F->A.d_tmap.p, F->A.nnz, F->A.d_tval.p, st));
  ctx->launches += 2;
  if (v_out) {
    k_fem1d_adv_v<<<(unsigned)((F->nn + 255) / 256), 256, 0, st>>>(F->nn, F->npe, F->nq, F->d_diag.p, F->d_cptr.p, F->d_cidx.p,
                                                                  F->d_shape.p, F->d_jac.p, F->d_elems.p, F->d_u.p, pr,
                                                                  F->d_v.p);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches += 1;
    GMRFB_CU(ctx, cudaMemcpyAsync(v_out, F->d_v.p, F->nn * sizeof(double), cudaMemcpyDefault, st));
  }
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (A_out) *A_out = &F->A;
  return GMRFB_OK;
}
GMRFB_ABI_CATCH

extern "C" gmrfb_status gmrfb_fem1d_spacetime_tangent(gmrfb_fem1d* F, int64_t nt, double dt, double nu, const double* w,
                                                      const uint8_t* prescribed, const gmrfb_spm** J_out,
                                                      double* f_out) try {
  if (!F || !w) return fail(F ? F->ctx : nullptr, GMRFB_ERR_INVALID, "gmrfb_fem1d_spacetime_tangent: NULL argument");
  gmrfb_ctx* ctx = F->ctx;
  if (nt < 2 || nt > 65535) return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem1d_spacetime_tangent: nt must be in 2..65535");
  GMRFB_CU(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t nn = F->nn, nnzP = F->M.nnz;
  if (2 * nnzP * (nt - 1) >= ((int64_t)1 << 30))
    return fail(ctx, GMRFB_ERR_INVALID, "gmrfb_fem1d_spacetime_tangent: space-time pattern too large");
  gmrfb_status rc = build_static(F, false, prescribed);
  if (rc != GMRFB_OK) return rc;
  if (F->nt != nt) {  // pattern of J: column block s holds block row s - 1 (part 1, s >= 1) above block row s (part 0, s <= nt - 2)
    GMRFB_CU(ctx, cudaStreamSynchronize(st));
    const std::vector<int64_t>& cp = F->M.colptr;
    const std::vector<int32_t>& ri = F->M.rowidx;
    std::vector<int64_t> jcp((size_t)(nt * nn) + 1, 0), jri;
    std::vector<int32_t> jmap;
    jri.reserve((size_t)(2 * nnzP * (nt - 1)));
    jmap.reserve((size_t)(2 * nnzP * (nt - 1)));
    F->jblk.assign((size_t)nt + 1, 0);
    for (int64_t s = 0; s < nt; s++) {
      for (int64_t c = 0; c < nn; c++) {
        if (s >= 1)
          for (int64_t k = cp[c]; k < cp[c + 1]; k++) {
            jri.push_back((s - 1) * nn + ri[k]);
            jmap.push_back((int32_t)(2 * k + 1));
          }
        if (s <= nt - 2)
          for (int64_t k = cp[c]; k < cp[c + 1]; k++) {
            jri.push_back(s * nn + ri[k]);
            jmap.push_back((int32_t)(2 * k));
          }
        jcp[s * nn + c + 1] = (int64_t)jri.size();
      }
      F->jblk[s + 1] = (int64_t)jri.size();
    }
    F->nt = 0;
    F->J.reset(new gmrfb_spm());
    if ((rc = spm_build(ctx, F->J.get(), (nt - 1) * nn, nt * nn, jcp.data(), jri.data(), nullptr, 0)) != GMRFB_OK) return rc;
    F->J->owned_by_plan = true;
    GMRFB_CU(ctx, F->d_jmap.upload(jmap, st));
    GMRFB_CU(ctx, F->d_jblk.upload(F->jblk, st));
    GMRFB_CU(ctx, F->d_w.alloc((size_t)(nt * nn)));
    GMRFB_CU(ctx, F->d_f.alloc((size_t)((nt - 1) * nn)));
    GMRFB_CU(ctx, cudaStreamSynchronize(st));  // jmap is a pageable temporary
    F->nt = nt;
  }
  GMRFB_CU(ctx, cudaMemcpyAsync(F->d_w.p, w, nt * nn * sizeof(double), cudaMemcpyDefault, st));  // host or device
  const uint8_t* pr = prescribed ? F->d_presc.p : nullptr;
  int64_t maxblk = 0;
  for (int64_t s = 0; s < nt; s++) maxblk = std::max(maxblk, F->jblk[s + 1] - F->jblk[s]);
  {
    ProfScope ps(ctx, PK_FEM, 0, 8.0 * F->J->nnz + 8.0 * nt * nn, 0, 0);
    k_fem1d_st_J<<<dim3((unsigned)((maxblk + 255) / 256), (unsigned)nt), 256, 0, st>>>(
        nn, F->npe, F->nq, F->d_jblk.p, F->d_jmap.p, F->d_cptr.p, F->d_cidx.p, F->d_shape.p, F->d_jac.p, F->d_elems.p,
        F->d_w.p, F->M.d_rowidx.p, F->d_colnz.p, pr, F->M.d_val.p, F->G.d_val.p, dt, nu, F->J->d_val.p);
  }
  GMRFB_CU(ctx, cudaGetLastError());
  GMRFB_CU(ctx, launch_gather_values(F->J->d_val.p, F->J->d_tmap.p, F->J->nnz, F->J->d_tval.p, st));
  ctx->launches += 2;
  if (f_out) {
    k_fem1d_st_f<<<dim3((unsigned)((nn + 255) / 256), (unsigned)(nt - 1)), 256, 0, st>>>(
        nn, F->npe, F->nq, F->M.d_colptr.p, F->M.d_rowidx.p, F->d_diag.p, F->d_cptr.p, F->d_cidx.p, F->d_shape.p,
        F->d_jac.p, F->d_elems.p, F->d_w.p, pr, F->M.d_val.p, F->G.d_val.p, dt, nu, F->d_f.p);
    GMRFB_CU(ctx, cudaGetLastError());
    ctx->launches += 1;
    GMRFB_CU(ctx, cudaMemcpyAsync(f_out, F->d_f.p, (nt - 1) * nn * sizeof(double), cudaMemcpyDefault, st));
  }
  GMRFB_CU(ctx, cudaStreamSynchronize(st));
  if (J_out) *J_out = F->J.get();
  return GMRFB_OK;
}
GMRFB_ABI_CATCH
