// Kernels and host-side tables of the Lagrange-triangle assembler (fem2d.cu).  No CUDA runtime calls in here: the
// file is also compiled as plain C++ by tools/probe/fem2d_emul.cpp, which runs every kernel thread by thread on the
// host and compares with the oracle (the container that builds this library has no GPU).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace gmrfb {
namespace fem2d {

constexpr int MAXPE = 6;

// barycentric points l0, l1, l2 and weights (sum 1: dOmega = weight * area) of the symmetric rules of degree 1 (centroid),
// 2 (3 interior points), 3 (Dunavant's 4 points with the negative centroid weight) and 4 (6 points)
inline std::vector<double> tri_rule2d(int degree) {
  if (degree == 1) return {1.0 / 3, 1.0 / 3, 1.0 / 3, 1.0};
  if (degree == 2) {
    const double a = 1.0 / 6, b = 2.0 / 3, w = 1.0 / 3;
    return {b, a, a, w, a, b, a, w, a, a, b, w};
  }
  if (degree == 3) {
    const double t = 1.0 / 3, w0 = -27.0 / 48, w1 = 25.0 / 48;
    return {t, t, t, w0, 0.6, 0.2, 0.2, w1, 0.2, 0.6, 0.2, w1, 0.2, 0.2, 0.6, w1};
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

// N[a] and dN[a][b] = dN_a / dl_b (barycentric coordinates as independent variables) at the barycentric point l
inline void tri_shape(int order, const double* l, double* N, double (*dN)[3]) {
  for (int a = 0; a < MAXPE; a++) {
    N[a] = 0.0;
    for (int b = 0; b < 3; b++) dN[a][b] = 0.0;
  }
  if (order == 1) {
    for (int a = 0; a < 3; a++) {
      N[a] = l[a];
      dN[a][a] = 1.0;
    }
    return;
  }
  for (int v = 0; v < 3; v++) {
    N[v] = l[v] * (2.0 * l[v] - 1.0);
    dN[v][v] = 4.0 * l[v] - 1.0;
  }
  const int ea[3] = {0, 1, 2}, eb[3] = {1, 2, 0};
  for (int e = 0; e < 3; e++) {
    N[3 + e] = 4.0 * l[ea[e]] * l[eb[e]];
    dN[3 + e][ea[e]] = 4.0 * l[eb[e]];
    dN[3 + e][eb[e]] = 4.0 * l[ea[e]];
  }
}

// nearest grid index: argmin |coords - x|, the first minimum on ties (src/datasets/darcy.jl:30-34)
inline int64_t nearest_index(const double* c, int64_t g, double x, bool sorted) {
  if (sorted) {
    const int64_t hi = std::lower_bound(c, c + g, x) - c;
    if (hi <= 0) return 0;
    int64_t best;
    if (hi >= g) best = g - 1;
    else best = (std::fabs(c[hi - 1] - x) <= std::fabs(c[hi] - x)) ? hi - 1 : hi;
    while (best > 0 && std::fabs(c[best - 1] - x) <= std::fabs(c[best] - x)) best--;  // repeated coordinates
    return best;
  }
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
}

// reference tables of one (order, rule): shape[q npe + a] = N_a(q); dref[(q npe + a) 2 + {0, 1}] = dN_a/dxi, dN_a/deta
// with xi = l1, eta = l2; wq[q].  Returns the number of quadrature points.
inline int build_tables(int order, int quad_degree, std::vector<double>& shape, std::vector<double>& dref,
                        std::vector<double>& wq) {
  const int npe = order == 1 ? 3 : 6;
  const std::vector<double> rule = tri_rule2d(quad_degree);
  const int nq = (int)(rule.size() / 4);
  shape.assign((size_t)nq * npe, 0.0);
  dref.assign((size_t)nq * npe * 2, 0.0);
  wq.assign((size_t)nq, 0.0);
  for (int q = 0; q < nq; q++) {
    double N[MAXPE], dN[MAXPE][3];
    tri_shape(order, &rule[4 * q], N, dN);
    for (int a = 0; a < npe; a++) {
      shape[(size_t)q * npe + a] = N[a];
      dref[((size_t)q * npe + a) * 2] = dN[a][1] - dN[a][0];
      dref[((size_t)q * npe + a) * 2 + 1] = dN[a][2] - dN[a][0];
    }
    wq[q] = rule[4 * q + 3];
  }
  return nq;
}

// coefficient-grid cell of every quadrature point: x_q = spatial_coordinate(cellvalues, q_point, cell_coords), then
// coeff_mat[x_idx, y_idx] of a column-major gx x gy array = entry x_idx + y_idx gx
inline void quad_point_cells(int64_t ne, int npe, int nq, const int32_t* conn, const double* nodes, const double* shape,
                             int64_t gx, const double* x_coords, int64_t gy, const double* y_coords, int32_t* cell) {
  const bool xs = std::is_sorted(x_coords, x_coords + gx), ys = std::is_sorted(y_coords, y_coords + gy);
  for (int64_t t = 0; t < ne; t++)
    for (int q = 0; q < nq; q++) {
      double x = 0.0, y = 0.0;
      for (int a = 0; a < npe; a++) {
        const int64_t v = conn[t * npe + a];
        x += shape[(size_t)q * npe + a] * nodes[2 * v];
        y += shape[(size_t)q * npe + a] * nodes[2 * v + 1];
      }
      const int64_t ix = nearest_index(x_coords, gx, x, xs), iy = nearest_index(y_coords, gy, y, ys);
      cell[(size_t)t * nq + q] = (int32_t)(ix + iy * gx);
    }
}

// physical gradients and dOmega at quadrature point q of element t; reference coordinates xi = l1, eta = l2, so that
// dref[(q npe + a) 2 + {0, 1}] = dN_a/dxi, dN_a/deta and [dN/dx, dN/dy] = [dN/dxi, dN/deta] J^-1, J = d(x, y)/d(xi, eta)
__global__ void k_fem2d_geom(int64_t ne, int npe, int nq, const double* __restrict__ nodes, const int32_t* __restrict__ conn,
                             const double* __restrict__ dref, const double* __restrict__ wq, double* __restrict__ grad,
                             double* __restrict__ jxw, int* __restrict__ bad) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ne * nq) return;
  const int64_t t = idx / nq;
  const int q = (int)(idx - t * nq);
  const int32_t* c = conn + t * npe;
  const double* dr = dref + (size_t)q * npe * 2;
  // J = sum_a x_a (x) dN_a/dref.  The reference derivatives sum to zero (partition of unity), so the coordinates are
  // taken relative to the first vertex: same Jacobian in exact arithmetic, without the cancellation of absolute
  // coordinates of size 1 in an element of size h (relative rounding error eps instead of eps / h)
  const double x0 = nodes[2 * (int64_t)c[0]], y0 = nodes[2 * (int64_t)c[0] + 1];
  double j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;
  for (int a = 1; a < npe; a++) {
    const double x = nodes[2 * (int64_t)c[a]] - x0, y = nodes[2 * (int64_t)c[a] + 1] - y0;
    j00 += x * dr[2 * a];
    j01 += x * dr[2 * a + 1];
    j10 += y * dr[2 * a];
    j11 += y * dr[2 * a + 1];
  }
  const double det = j00 * j11 - j01 * j10;
  if (!(fabs(det) > 0.0)) atomicExch(bad, 1);
  const double inv = 1.0 / det;
  double* g = grad + (size_t)idx * npe * 2;
  for (int a = 0; a < npe; a++) {
    g[2 * a] = (dr[2 * a] * j11 - dr[2 * a + 1] * j10) * inv;
    g[2 * a + 1] = (-dr[2 * a] * j01 + dr[2 * a + 1] * j00) * inv;
  }
  jxw[idx] = 0.5 * fabs(det) * wq[q];
}

__device__ __forceinline__ void entry_decode(int32_t e, int npe, int64_t& t, int& i, int& j) {
  const int npe2 = npe * npe;
  t = e / npe2;
  const int r = e - (int32_t)t * npe2;
  i = r / npe;
  j = r - i * npe;
}

// G[k] = sum over the element entries (t, i, j) on nonzero k of  sum_q coeff(cell(t, q)) grad phi_i . grad phi_j dOmega;
// rows of prescribed dofs become identity rows
__global__ void k_fem2d_stiffness(int64_t nnz, int npe, int nq, const int64_t* __restrict__ cptr,
                                  const int32_t* __restrict__ cidx, const double* __restrict__ grad,
                                  const double* __restrict__ jxw, const int32_t* __restrict__ cellq,
                                  const double* __restrict__ coeff, const int32_t* __restrict__ rowidx,
                                  const int64_t* __restrict__ diag, const uint8_t* __restrict__ presc,
                                  double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int32_t r = rowidx[k];
  if (presc && presc[r]) {
    out[k] = (k == diag[r]) ? 1.0 : 0.0;
    return;
  }
  double v = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
    int64_t t;
    int i, j;
    entry_decode(cidx[p], npe, t, i, j);
    double ve = 0.0;
    for (int q = 0; q < nq; q++) {
      const int64_t tq = t * nq + q;
      const double* g = grad + (size_t)tq * npe * 2;
      const double a = coeff ? coeff[cellq[tq]] : 1.0;
      ve += a * (g[2 * i] * g[2 * j] + g[2 * i + 1] * g[2 * j + 1]) * jxw[tq];
    }
    v += ve;
  }
  out[k] = v;
}

// f[i] = beta sum over the elements (t, a) at node i of sum_q phi_a dOmega (through the diagonal's contribution list:
// one entry per element at the node); prescribed rows are zero
__global__ void k_fem2d_load(int64_t nn, int npe, int nq, const int64_t* __restrict__ diag, const int64_t* __restrict__ cptr,
                             const int32_t* __restrict__ cidx, const double* __restrict__ shape,
                             const double* __restrict__ jxw, const uint8_t* __restrict__ presc, double beta,
                             double* __restrict__ f) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  if (presc && presc[n]) {
    f[n] = 0.0;
    return;
  }
  double v = 0.0;
  for (int64_t p = cptr[diag[n]]; p < cptr[diag[n] + 1]; p++) {
    int64_t t;
    int i, j;
    entry_decode(cidx[p], npe, t, i, j);
    double ve = 0.0;
    for (int q = 0; q < nq; q++) ve += shape[q * npe + i] * jxw[t * nq + q];
    v += ve;
  }
  f[n] = beta * v;
}

// consistent mass M[k] = sum phi_i phi_j dOmega
__global__ void k_fem2d_mass(int64_t nnz, int npe, int nq, const int64_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                             const double* __restrict__ shape, const double* __restrict__ jxw, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  double v = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
    int64_t t;
    int i, j;
    entry_decode(cidx[p], npe, t, i, j);
    double ve = 0.0;
    for (int q = 0; q < nq; q++) ve += shape[q * npe + i] * shape[q * npe + j] * jxw[t * nq + q];
    v += ve;
  }
  out[k] = v;
}

// element-lumped mass (`lump_matrix(me, ip)`, src/spdes/shallow_water.jl:115): kind 1 = row sums of the element mass,
// kind 2 = its diagonal scaled to the element's total mass, diag(me) sum(me) / sum(diag(me))
__global__ void k_fem2d_lump(int64_t nn, int npe, int nq, int kind, const int64_t* __restrict__ diag,
                             const int64_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                             const double* __restrict__ shape, const double* __restrict__ jxw, double* __restrict__ ml) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  double m = 0.0;
  for (int64_t p = cptr[diag[n]]; p < cptr[diag[n] + 1]; p++) {
    int64_t t;
    int a, j;
    entry_decode(cidx[p], npe, t, a, j);
    if (kind == 1) {
      // sum_j me[a, j] = sum_q phi_a (sum_j phi_j) dOmega, written out as the reference sums it
      double row = 0.0;
      for (int b = 0; b < npe; b++) {
        double mab = 0.0;
        for (int q = 0; q < nq; q++) mab += shape[q * npe + a] * shape[q * npe + b] * jxw[t * nq + q];
        row += mab;
      }
      m += row;
    } else {
      double total = 0.0, dsum = 0.0, maa = 0.0;
      for (int b = 0; b < npe; b++)
        for (int c = 0; c < npe; c++) {
          double mbc = 0.0;
          for (int q = 0; q < nq; q++) mbc += shape[q * npe + b] * shape[q * npe + c] * jxw[t * nq + q];
          total += mbc;
          if (b == c) {
            dsum += mbc;
            if (b == a) maa = mbc;
          }
        }
      m += maa * (total / dsum);
    }
  }
  ml[n] = m;
}

__global__ void k_fem2d_set_lumped(int64_t nnz, const int32_t* __restrict__ rowidx, const int64_t* __restrict__ diag,
                                   const double* __restrict__ ml, double* __restrict__ Mv) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int32_t r = rowidx[k];
  Mv[k] = (k == diag[r]) ? ml[r] : 0.0;
}

// K = kappa^2 Mt + G (prescribed dofs: Mt_ii = presc_mass, G_ii = 1; src/spdes/shallow_water.jl:172-181), w_i = 1 / Mt_i
// scaled by `wscale`
__global__ void k_fem2d_matern_k(int64_t nn, const int64_t* __restrict__ diag, const double* __restrict__ mass,
                                 const uint8_t* __restrict__ presc, double presc_mass, double kappa2, double wscale,
                                 double* __restrict__ kval, double* __restrict__ w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const bool p = presc && presc[i];
  const double mt = p ? presc_mass : mass[i];
  if (p) kval[diag[i]] = 1.0;
  kval[diag[i]] += kappa2 * mt;
  w[i] = wscale / mt;
}

// J[k] = sum over the element entries (t, i, j) on k of  sum_q (s grad phi_i . grad phi_j + 3 phi_i u_q^2 phi_j) dOmega,
// u_q = sum_a u[conn(t, a)] phi_a(q); rows of prescribed dofs are skipped (_research/elliptic_chen24.jl:207-209, :259-261)
__global__ void k_fem2d_cubic_J(int64_t nnz, int npe, int nq, const int64_t* __restrict__ cptr,
                                const int32_t* __restrict__ cidx, const int32_t* __restrict__ conn,
                                const double* __restrict__ shape, const double* __restrict__ grad,
                                const double* __restrict__ jxw, const double* __restrict__ u,
                                const int32_t* __restrict__ rowidx, const uint8_t* __restrict__ presc, double s,
                                double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  if (presc && presc[rowidx[k]]) {
    out[k] = 0.0;
    return;
  }
  double v = 0.0;
  for (int64_t p = cptr[k]; p < cptr[k + 1]; p++) {
    int64_t t;
    int i, j;
    entry_decode(cidx[p], npe, t, i, j);
    double w[MAXPE];
    for (int a = 0; a < npe; a++) w[a] = u[conn[t * npe + a]];
    double ve = 0.0;
    for (int q = 0; q < nq; q++) {
      const int64_t tq = t * nq + q;
      const double* sh = shape + q * npe;
      const double* g = grad + (size_t)tq * npe * 2;
      double uq = 0.0;
      for (int a = 0; a < npe; a++) uq += w[a] * sh[a];
      ve += (s * (g[2 * i] * g[2 * j] + g[2 * i + 1] * g[2 * j + 1]) + 3.0 * sh[i] * uq * uq * sh[j]) * jxw[tq];
    }
    v += ve;
  }
  out[k] = v;
}

// f[n] = sum over the elements (t, a) at node n of  sum_q (s grad phi_a . grad u_q + phi_a u_q^3) dOmega
__global__ void k_fem2d_cubic_f(int64_t nn, int npe, int nq, const int64_t* __restrict__ diag,
                                const int64_t* __restrict__ cptr, const int32_t* __restrict__ cidx,
                                const int32_t* __restrict__ conn, const double* __restrict__ shape,
                                const double* __restrict__ grad, const double* __restrict__ jxw,
                                const double* __restrict__ u, const uint8_t* __restrict__ presc, double s,
                                double* __restrict__ f) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  if (presc && presc[n]) {
    f[n] = 0.0;
    return;
  }
  double v = 0.0;
  for (int64_t p = cptr[diag[n]]; p < cptr[diag[n] + 1]; p++) {
    int64_t t;
    int a, j;
    entry_decode(cidx[p], npe, t, a, j);
    double w[MAXPE];
    for (int b = 0; b < npe; b++) w[b] = u[conn[t * npe + b]];
    double ve = 0.0;
    for (int q = 0; q < nq; q++) {
      const int64_t tq = t * nq + q;
      const double* sh = shape + q * npe;
      const double* g = grad + (size_t)tq * npe * 2;
      double uq = 0.0, ux = 0.0, uy = 0.0;
      for (int b = 0; b < npe; b++) {
        uq += w[b] * sh[b];
        ux += w[b] * g[2 * b];
        uy += w[b] * g[2 * b + 1];
      }
      ve += (s * (g[2 * a] * ux + g[2 * a + 1] * uy) + sh[a] * uq * uq * uq) * jxw[tq];
    }
    v += ve;
  }
  f[n] = v;
}

}  // namespace fem2d
}  // namespace gmrfb
