// Host emulation of the fem2d / spgemm kernels: the kernel headers are compiled as plain C++ (the CUDA qualifiers are
// defined away, blockIdx / threadIdx are globals) and every kernel is run thread by thread.  The container that
// builds libgmrfb.so has no GPU; this checks the kernel arithmetic, the reference tables, the pattern / contribution
// lists and the coefficient lookup against the oracle before the GPU run (tools/probe/fem2d_emul.py drives it).
//   g++ -O2 -shared -fPIC -std=c++17 -I diffeqgmrfs.jl_b200/csrc tools/probe/fem2d_emul.cpp -o /tmp/libfem2d_emul.so
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
struct Dim3 {
  unsigned x = 0, y = 0, z = 0;
};
static Dim3 blockIdx, blockDim, threadIdx;
static int atomicExch(int* p, int v) {
  int o = *p;
  *p = v;
  return o;
}

#include "fem2d_kernels.cuh"
#include "fem_pattern.hpp"
#include "spgemm_kernels.cuh"

using namespace gmrfb;
using namespace gmrfb::fem2d;

// run `body` for every thread of a 1-D launch covering n items with 256-thread blocks
template <class F>
static void launch(int64_t n, F body) {
  blockDim.x = 256;
  for (unsigned b = 0; b < (unsigned)((n + 255) / 256); b++)
    for (unsigned t = 0; t < 256; t++) {
      blockIdx.x = b;
      threadIdx.x = t;
      body();
    }
}

struct Emul {
  int64_t nn, ne;
  int order, npe, nq;
  std::vector<double> nodes, shape, dref, wq, grad, jxw;
  std::vector<int32_t> conn, cellq, rowidx;
  ElementPattern P;
};

extern "C" {

Emul* emul_create(int order, int quad_degree, int64_t nn, const double* nodes, int64_t ne, const int64_t* elems) {
  Emul* E = new Emul();
  E->nn = nn;
  E->ne = ne;
  E->order = order;
  E->npe = order == 1 ? 3 : 6;
  E->nodes.assign(nodes, nodes + 2 * nn);
  E->conn.resize((size_t)E->npe * ne);
  for (int64_t k = 0; k < (int64_t)E->npe * ne; k++) E->conn[k] = (int32_t)elems[k];
  E->nq = build_tables(order, quad_degree ? quad_degree : order + 1, E->shape, E->dref, E->wq);
  if (!build_element_pattern(nn, ne, E->npe, E->conn.data(), E->P)) return nullptr;
  E->rowidx.assign(E->P.rowval.begin(), E->P.rowval.end());
  E->grad.resize((size_t)ne * E->nq * E->npe * 2);
  E->jxw.resize((size_t)ne * E->nq);
  int bad = 0;
  launch(ne * E->nq, [&] {
    k_fem2d_geom(ne, E->npe, E->nq, E->nodes.data(), E->conn.data(), E->dref.data(), E->wq.data(), E->grad.data(),
                 E->jxw.data(), &bad);
  });
  if (bad) return nullptr;
  return E;
}
void emul_destroy(Emul* E) { delete E; }
int64_t emul_nnz(Emul* E) { return (int64_t)E->P.rowval.size(); }
int emul_nq(Emul* E) { return E->nq; }
void emul_pattern(Emul* E, int64_t* colptr, int64_t* rowval) {
  std::memcpy(colptr, E->P.colptr.data(), (E->nn + 1) * sizeof(int64_t));
  std::memcpy(rowval, E->P.rowval.data(), E->P.rowval.size() * sizeof(int64_t));
}
void emul_set_grid(Emul* E, int64_t gx, const double* xc, int64_t gy, const double* yc) {
  E->cellq.resize((size_t)E->ne * E->nq);
  quad_point_cells(E->ne, E->npe, E->nq, E->conn.data(), E->nodes.data(), E->shape.data(), gx, xc, gy, yc, E->cellq.data());
}
void emul_stiffness(Emul* E, const double* coeff, const uint8_t* presc, double beta, double* out, double* load) {
  const int64_t nnz = emul_nnz(E);
  launch(nnz, [&] {
    k_fem2d_stiffness(nnz, E->npe, E->nq, E->P.cptr.data(), E->P.cidx.data(), E->grad.data(), E->jxw.data(),
                      E->cellq.data(), coeff, E->rowidx.data(), E->P.diag.data(), presc, out);
  });
  if (load)
    launch(E->nn, [&] {
      k_fem2d_load(E->nn, E->npe, E->nq, E->P.diag.data(), E->P.cptr.data(), E->P.cidx.data(), E->shape.data(),
                   E->jxw.data(), presc, beta, load);
    });
}
void emul_mass(Emul* E, int lumping, double* out, double* ml) {
  const int64_t nnz = emul_nnz(E);
  if (lumping == 0) {
    launch(nnz, [&] {
      k_fem2d_mass(nnz, E->npe, E->nq, E->P.cptr.data(), E->P.cidx.data(), E->shape.data(), E->jxw.data(), out);
    });
    return;
  }
  launch(E->nn, [&] {
    k_fem2d_lump(E->nn, E->npe, E->nq, lumping, E->P.diag.data(), E->P.cptr.data(), E->P.cidx.data(), E->shape.data(),
                 E->jxw.data(), ml);
  });
  launch(nnz, [&] { k_fem2d_set_lumped(nnz, E->rowidx.data(), E->P.diag.data(), ml, out); });
}
// K values (unit stiffness, then the Matern diagonal) and w
void emul_matern_k(Emul* E, const double* ml, const uint8_t* presc, double presc_mass, double kappa2, double wscale,
                   double* kval, double* w) {
  const int64_t nnz = emul_nnz(E);
  launch(nnz, [&] {
    k_fem2d_stiffness(nnz, E->npe, E->nq, E->P.cptr.data(), E->P.cidx.data(), E->grad.data(), E->jxw.data(), nullptr,
                      nullptr, E->rowidx.data(), E->P.diag.data(), nullptr, kval);
  });
  launch(E->nn, [&] { k_fem2d_matern_k(E->nn, E->P.diag.data(), ml, presc, presc_mass, kappa2, wscale, kval, w); });
}
void emul_cubic(Emul* E, const double* u, double s, const uint8_t* presc, double* Jout, double* fout) {
  const int64_t nnz = emul_nnz(E);
  launch(nnz, [&] {
    k_fem2d_cubic_J(nnz, E->npe, E->nq, E->P.cptr.data(), E->P.cidx.data(), E->conn.data(), E->shape.data(),
                    E->grad.data(), E->jxw.data(), u, E->rowidx.data(), presc, s, Jout);
  });
  launch(E->nn, [&] {
    k_fem2d_cubic_f(E->nn, E->npe, E->nq, E->P.diag.data(), E->P.cptr.data(), E->P.cidx.data(), E->conn.data(),
                    E->shape.data(), E->grad.data(), E->jxw.data(), u, presc, s, fout);
  });
}

// C = alpha A diag(w) B: pattern into caller arrays sized by a first call with crow == NULL (returns nnz)
int64_t emul_spgemm(int64_t m, int64_t k, int64_t n, const int64_t* acolptr, const int32_t* arow, const double* aval,
                    const int64_t* bcolptr, const int32_t* brow, const double* bval, const double* w, double alpha,
                    int64_t* ccolptr, int32_t* crow_out, double* cval) {
  std::vector<int64_t> cp, cr;
  std::vector<int32_t> colnz;
  spgemm::product_pattern(m, n, acolptr, arow, bcolptr, brow, cp, cr, colnz);
  const int64_t nnz = (int64_t)cr.size();
  if (!crow_out) return nnz;
  std::memcpy(ccolptr, cp.data(), (n + 1) * sizeof(int64_t));
  std::vector<int32_t> cr32(cr.begin(), cr.end());
  std::memcpy(crow_out, cr32.data(), nnz * sizeof(int32_t));
  launch(nnz, [&] {
    spgemm::k_spgemm(nnz, cr32.data(), colnz.data(), acolptr, arow, aval, bcolptr, brow, bval, w, alpha, cval);
  });
  return nnz;
}

// Qpost = Q + A' diag(w) A through the plan of spgemm::postprec_pattern with `nthreads` workers; values by the formula of
// k_postprec (sparse_kernels.cu).  First call with orow_out == NULL returns nnz(Qpost).
int64_t emul_postprec(int64_t n, int64_t m, const int64_t* qcolptr, const int32_t* qrow, const double* qval,
                      const int64_t* acolptr, const int32_t* arow, const double* aval, int64_t nnzA, const double* w,
                      int nthreads, int64_t* ocolptr_out, int32_t* orow_out, double* oval) {
  std::vector<int64_t> ocolptr, qsrc, pptr, pa, pb;
  std::vector<int32_t> orow, prow;
  spgemm::postprec_pattern(n, m, qcolptr, qrow, acolptr, arow, nnzA, nthreads, ocolptr, orow, qsrc, pptr, prow, pa, pb);
  const int64_t nnz = (int64_t)orow.size();
  if (!orow_out) return nnz;
  std::memcpy(ocolptr_out, ocolptr.data(), (n + 1) * sizeof(int64_t));
  std::memcpy(orow_out, orow.data(), nnz * sizeof(int32_t));
  for (int64_t k = 0; k < nnz; k++) {
    double v = qsrc[k] >= 0 ? qval[qsrc[k]] : 0.0;
    for (int64_t t = pptr[k]; t < pptr[k + 1]; t++) v += w[prow[t]] * aval[pa[t]] * aval[pb[t]];
    oval[k] = v;
  }
  return nnz;
}
}
