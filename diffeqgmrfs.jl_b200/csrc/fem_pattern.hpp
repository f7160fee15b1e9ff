// Sparsity pattern of an element-assembled matrix together with, for every nonzero, the list of element-matrix
// entries that are summed into it (the "gather" form of finite-element assembly: one thread per nonzero, fixed
// summation order, no atomics).  Shared by the triangle (fem.cu) and line (fem1d.cu) assemblers.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

namespace gmrfb {

struct ElementPattern {
  std::vector<int64_t> colptr;  // nnodes + 1
  std::vector<int64_t> rowval;  // nnz, ascending within a column
  std::vector<int64_t> cptr;    // nnz + 1: contribution list of every nonzero
  std::vector<int32_t> cidx;    // element entry e * npe^2 + i * npe + j (i: local row, j: local column), ascending
  std::vector<int64_t> diag;    // nnodes: position of entry (c, c)
};

// conn: nelem x npe node ids (0-based).  Returns false when some node belongs to no element.
inline bool build_element_pattern(int64_t nnodes, int64_t nelem, int npe, const int32_t* conn, ElementPattern& P) {
  const int npe2 = npe * npe;
  const int64_t ne = (int64_t)npe2 * nelem;
  std::vector<int64_t> key((size_t)ne);
  std::vector<int32_t> ord((size_t)ne);
  std::iota(ord.begin(), ord.end(), 0);
  for (int64_t t = 0; t < nelem; t++)
    for (int i = 0; i < npe; i++)
      for (int j = 0; j < npe; j++)
        key[npe2 * t + npe * i + j] = (int64_t)conn[npe * t + j] * nnodes + conn[npe * t + i];
  std::sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return key[a] < key[b] || (key[a] == key[b] && a < b); });
  P.colptr.assign((size_t)nnodes + 1, 0);
  P.diag.assign((size_t)nnodes, -1);
  P.rowval.clear();
  P.cptr.clear();
  P.cidx.assign((size_t)ne, 0);
  P.rowval.reserve((size_t)ne / 2);
  P.cptr.reserve((size_t)ne / 2);
  int64_t prev = -1;
  for (int64_t q = 0; q < ne; q++) {
    const int32_t e = ord[q];
    if (key[e] != prev) {
      prev = key[e];
      const int64_t c = prev / nnodes, r = prev % nnodes;
      if (r == c) P.diag[c] = (int64_t)P.rowval.size();
      P.rowval.push_back(r);
      P.cptr.push_back(q);
      P.colptr[c + 1]++;
    }
    P.cidx[q] = e;
  }
  P.cptr.push_back(ne);
  for (int64_t c = 0; c < nnodes; c++) {
    if (P.diag[c] < 0) return false;
    P.colptr[c + 1] += P.colptr[c];
  }
  return true;
}

}  // namespace gmrfb
