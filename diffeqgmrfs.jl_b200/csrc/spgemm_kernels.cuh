// Kernel and symbolic phase of the fixed-pattern sparse product (gmrfb_spgemm, spm.cu).  No CUDA runtime calls in here:
// also compiled as plain C++ by tools/probe/fem2d_emul.cpp (host emulation against SciPy).
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <system_error>
#include <thread>
#include <vector>

namespace gmrfb {
namespace spgemm {

// run body(0), ..., body(nblocks - 1), one worker thread per block; a block whose thread cannot be created runs on
// the calling thread (no exception leaves a joinable thread behind)
template <class F>
inline void run_blocks(int nblocks, F body) {
  std::vector<std::thread> pool;
  std::vector<int> here(1, 0);
  for (int t = 1; t < nblocks; t++) {
    try {
      pool.emplace_back(body, t);
    } catch (const std::system_error&) {
      here.push_back(t);
    }
  }
  for (int t : here) body(t);
  for (std::thread& th : pool) th.join();
}

// Symbolic phase of Qpost = Q + A' diag(w) A (gmrfb_postprec, spm.cu): the pattern of Qpost (ocolptr / orow), for every
// output entry the position of the Q entry it starts from (qsrc, -1: none) and its list of products
// w[prow] A[pa] A[pb] (pptr).  nthreads = 0: GMRFB_HOST_THREADS or the hardware threads (at most 16).
inline void postprec_pattern(int64_t n, int64_t m, const int64_t* Qcolptr, const int32_t* Qrow, const int64_t* Acolptr,
                             const int32_t* Arow, int64_t nnzA, int nthreads, std::vector<int64_t>& ocolptr,
                             std::vector<int32_t>& orow, std::vector<int64_t>& qsrc, std::vector<int64_t>& pptr,
                             std::vector<int32_t>& prow, std::vector<int64_t>& pa, std::vector<int64_t>& pb) {
  // rows of A (host): rowptr / (col, csc position)
  std::vector<int64_t> rptr(m + 1, 0);
  for (int64_t p = 0; p < nnzA; p++) rptr[Arow[p] + 1]++;
  for (int64_t i = 0; i < m; i++) rptr[i + 1] += rptr[i];
  std::vector<int32_t> rcol(nnzA);
  std::vector<int64_t> rpos(nnzA);
  {
    std::vector<int64_t> fillp(rptr.begin(), rptr.end() - 1);
    for (int64_t j = 0; j < n; j++)
      for (int64_t p = Acolptr[j]; p < Acolptr[j + 1]; p++) {
        int64_t q = fillp[Arow[p]]++;
        rcol[q] = (int32_t)j;
        rpos[q] = p;
      }
  }
  // column by column: merge pattern(Q[:,j]) with the columns i reached through shared rows k.  Columns are independent,
  // so contiguous column blocks are built by worker threads (GMRFB_HOST_THREADS, default: the hardware threads, at
  // most 16) and concatenated in column order: the plan is the same for any number of workers.
  struct Prod {
    int32_t i;   // output row
    int32_t k;   // observation row
    int64_t pa;  // position of A[k,i]
    int64_t pb;  // position of A[k,j]
  };
  struct Part {
    std::vector<int64_t> colcnt, qsrc, pend, pa, pb;  // pend: end of every output entry's product list, block-local
    std::vector<int32_t> orow, prow;
  };
  if (nthreads <= 0) {
    nthreads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = std::getenv("GMRFB_HOST_THREADS")) nthreads = std::max(1, std::atoi(e));
    if (nnzA < 200000) nthreads = 1;
  }
  nthreads = (int)std::min<int64_t>(nthreads, std::max<int64_t>(n, 1));
  std::vector<Part> parts(nthreads);
  auto build_block = [&](int t) {
    Part& R = parts[t];
    const int64_t j0 = n * t / nthreads, j1 = n * (t + 1) / nthreads;
    std::vector<Prod> prods;
    R.colcnt.reserve((size_t)(j1 - j0));
    for (int64_t j = j0; j < j1; j++) {
      prods.clear();
      for (int64_t p = Acolptr[j]; p < Acolptr[j + 1]; p++) {
        int32_t k = Arow[p];
        for (int64_t q = rptr[k]; q < rptr[k + 1]; q++) prods.push_back({rcol[q], k, rpos[q], p});
      }
      std::sort(prods.begin(), prods.end(), [](const Prod& a, const Prod& b) { return a.i < b.i || (a.i == b.i && a.k < b.k); });
      size_t ip = 0;
      int64_t qp = Qcolptr[j], qe = Qcolptr[j + 1];
      const size_t before = R.orow.size();
      while (ip < prods.size() || qp < qe) {
        int32_t ri = ip < prods.size() ? prods[ip].i : INT32_MAX;
        int32_t rq = qp < qe ? Qrow[qp] : INT32_MAX;
        int32_t r = std::min(ri, rq);
        R.orow.push_back(r);
        R.qsrc.push_back(rq == r ? qp++ : -1);
        while (ip < prods.size() && prods[ip].i == r) {
          R.prow.push_back(prods[ip].k);
          R.pa.push_back(prods[ip].pa);
          R.pb.push_back(prods[ip].pb);
          ip++;
        }
        R.pend.push_back((int64_t)R.prow.size());
      }
      R.colcnt.push_back((int64_t)(R.orow.size() - before));
    }
  };
  run_blocks(nthreads, build_block);
  ocolptr.assign((size_t)n + 1, 0);
  qsrc.clear(), pptr.clear(), pa.clear(), pb.clear(), orow.clear(), prow.clear();
  {
    size_t nout = 0, nprod = 0;
    for (const Part& R : parts) nout += R.orow.size(), nprod += R.prow.size();
    orow.reserve(nout), qsrc.reserve(nout), pptr.reserve(nout + 1);
    prow.reserve(nprod), pa.reserve(nprod), pb.reserve(nprod);
    pptr.push_back(0);
    int64_t j = 0;
    for (const Part& R : parts) {
      const int64_t base = (int64_t)prow.size();
      for (int64_t c : R.colcnt) {
        ocolptr[j + 1] = ocolptr[j] + c;
        j++;
      }
      orow.insert(orow.end(), R.orow.begin(), R.orow.end());
      qsrc.insert(qsrc.end(), R.qsrc.begin(), R.qsrc.end());
      for (int64_t e : R.pend) pptr.push_back(base + e);
      prow.insert(prow.end(), R.prow.begin(), R.prow.end());
      pa.insert(pa.end(), R.pa.begin(), R.pa.end());
      pb.insert(pb.end(), R.pb.begin(), R.pb.end());
    }
    parts.clear();
    parts.shrink_to_fit();
  }
}

// pattern of C = A B (A: m x k, B: k x n, CSC with sorted columns): column j is the sorted union of the columns of A
// selected by the rows of B(:, j).  colnz[p] = column of the CSC position p of C.
inline void product_pattern(int64_t m, int64_t n, const int64_t* acolptr, const int32_t* arow, const int64_t* bcolptr,
                            const int32_t* brow, std::vector<int64_t>& ccolptr, std::vector<int64_t>& crow,
                            std::vector<int32_t>& colnz, int nthreads = 0) {
  if (nthreads <= 0) {
    nthreads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = std::getenv("GMRFB_HOST_THREADS")) nthreads = std::max(1, std::atoi(e));
    if (n < 20000) nthreads = 1;
  }
  nthreads = (int)std::min<int64_t>(nthreads, std::max<int64_t>(n, 1));
  struct Part {
    std::vector<int64_t> colcnt, crow;
  };
  std::vector<Part> parts(nthreads);
  auto build_block = [&](int t) {  // contiguous column blocks, concatenated in column order below
    Part& R = parts[t];
    const int64_t j0 = n * t / nthreads, j1 = n * (t + 1) / nthreads;
    std::vector<int64_t> mark((size_t)m, -1);
    std::vector<int32_t> rows;
    for (int64_t j = j0; j < j1; j++) {
      rows.clear();
      for (int64_t p = bcolptr[j]; p < bcolptr[j + 1]; p++) {
        const int32_t kk = brow[p];
        for (int64_t q = acolptr[kk]; q < acolptr[kk + 1]; q++) {
          const int32_t i = arow[q];
          if (mark[i] != j) {
            mark[i] = j;
            rows.push_back(i);
          }
        }
      }
      std::sort(rows.begin(), rows.end());
      R.crow.insert(R.crow.end(), rows.begin(), rows.end());
      R.colcnt.push_back((int64_t)rows.size());
    }
  };
  run_blocks(nthreads, build_block);
  ccolptr.assign((size_t)n + 1, 0);
  crow.clear();
  colnz.clear();
  int64_t j = 0;
  for (const Part& R : parts) {
    crow.insert(crow.end(), R.crow.begin(), R.crow.end());
    for (int64_t c : R.colcnt) {
      ccolptr[j + 1] = ccolptr[j] + c;
      colnz.insert(colnz.end(), (size_t)c, (int32_t)j);
      j++;
    }
  }
}

__global__ void k_spgemm(int64_t nnzC, const int32_t* __restrict__ crow, const int32_t* __restrict__ ccol,
                         const int64_t* __restrict__ acolptr, const int32_t* __restrict__ arow,
                         const double* __restrict__ aval, const int64_t* __restrict__ bcolptr,
                         const int32_t* __restrict__ brow, const double* __restrict__ bval,
                         const double* __restrict__ w, double alpha, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnzC) return;
  const int32_t i = crow[k], j = ccol[k];
  double v = 0.0;
  for (int64_t p = bcolptr[j]; p < bcolptr[j + 1]; p++) {
    const int32_t kk = brow[p];
    int64_t lo = acolptr[kk], hi = acolptr[kk + 1];  // first position in [lo, hi) with arow >= i
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (arow[mid] < i) lo = mid + 1;
      else hi = mid;
    }
    if (lo < acolptr[kk + 1] && arow[lo] == i) v += aval[lo] * (w ? w[kk] : 1.0) * bval[p];
  }
  out[k] = alpha * v;
}
}  // namespace spgemm
}  // namespace gmrfb
