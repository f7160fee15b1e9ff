// Data-parallel phases of the symbolic analysis as CUDA kernels (BASELINE north star: "symbolic analysis ... on the
// GPU").  Given the fill-reducing permutation, the elimination tree and the supernode partition (sequential graph
// algorithms, symbolic.cpp), everything that is an independent per-row / per-entry computation runs here:
//   * validation of the CSC pattern + symmetric adjacency without the diagonal (structurally symmetric input, the only
//     kind the reference produces: `Symmetric(sparse(...))` of an assembled matrix; anything else takes the host path)
//   * the permuted adjacency (rows gathered through the permutation, neighbours relabelled, every row sorted)
//   * the internal (postordered) adjacency above the diagonal
//   * relmap: position of every below-row of a front inside its parent's front
//   * amap:   arena slot of every stored matrix entry
// Results are bit-identical to the host loops they replace (tests/test_symbolic.py compares the two paths).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "symbolic.hpp"

namespace gmrfb {

namespace {

template <typename T>
struct DBuf {  // plain owning device buffer (the analysis runs before any handle exists: no pooling needed)
  T* p = nullptr;
  size_t n = 0;
  ~DBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    n = count;
    return cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
  }
  cudaError_t upload(const T* h, size_t count, cudaStream_t st) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess || count == 0) return e;
    return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st);
  }
};

#define SG_CU(call)                                                                   \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      cudaGetLastError();                                                             \
      return std::string("symbolic (GPU): " #call ": ") + cudaGetErrorString(e_);     \
    }                                                                                 \
  } while (0)

constexpr int SORT_MAX = 256;  // rows up to this length are sorted by their thread; longer rows are flagged for the host

// flags: bit 0 = malformed pattern, bit 1 = not structurally symmetric
__global__ void k_pattern_check(int64_t n, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int base,
                                int64_t nnz, int32_t* __restrict__ deg, int* __restrict__ flags) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int64_t p0 = colptr[c] - base, p1 = colptr[c + 1] - base;
  if (p0 > p1 || p0 < 0 || p1 > nnz) {
    atomicOr(flags, 1);
    return;
  }
  int32_t d = 0;
  for (int64_t p = p0; p < p1; p++) {
    const int64_t r = rowval[p] - base;
    if (r < 0 || r >= n || (p > p0 && rowval[p] <= rowval[p - 1])) {
      atomicOr(flags, 1);
      return;
    }
    if (r == c) continue;
    d++;
    // the mirrored entry (c, r) must be stored in column r
    int64_t lo = colptr[r] - base, hi = colptr[r + 1] - base;
    if (lo > hi || lo < 0 || hi > nnz) {
      atomicOr(flags, 1);
      return;
    }
    bool found = false;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int64_t v = rowval[mid] - base;
      if (v == c) {
        found = true;
        break;
      }
      if (v < c)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (!found) atomicOr(flags, 2);
  }
  deg[c] = d;
}

__global__ void k_adj_fill(int64_t n, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int base,
                           const int64_t* __restrict__ xadj, int32_t* __restrict__ adj) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  int64_t o = xadj[c];
  for (int64_t p = colptr[c] - base; p < colptr[c + 1] - base; p++) {
    const int64_t r = rowval[p] - base;
    if (r != c) adj[o++] = (int32_t)r;
  }
}

__device__ __forceinline__ void insertion_sort(int32_t* a, int len) {
  for (int i = 1; i < len; i++) {
    const int32_t v = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > v) {
      a[j + 1] = a[j];
      j--;
    }
    a[j + 1] = v;
  }
}

// out row k = sorted { label[adj[p]] : p in row src[k] of (xadj, adj), label > lower[k] }.  With count_only the row
// lengths are written to cnt instead.  Rows longer than SORT_MAX are left unsorted and counted in *nlong.
__global__ void k_rows_map(int64_t n, const int64_t* __restrict__ xadj, const int32_t* __restrict__ adj,
                           const int32_t* __restrict__ src, const int32_t* __restrict__ label, bool filter_above,
                           const int64_t* __restrict__ oxadj, int32_t* __restrict__ oadj, int64_t* __restrict__ cnt,
                           int* __restrict__ nlong) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t v = src[k];
  if (cnt) {
    int64_t c = 0;
    for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) c += (!filter_above || label[adj[p]] > k);
    cnt[k] = c;
    return;
  }
  int64_t o = oxadj[k];
  for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) {
    const int32_t l = label[adj[p]];
    if (!filter_above || l > k) oadj[o++] = l;
  }
  const int64_t len = o - oxadj[k];
  if (len <= SORT_MAX)
    insertion_sort(oadj + oxadj[k], (int)len);
  else
    atomicAdd(nlong, 1);
}

// relmap[q] for q in the below part of front s: position of row rows[q] in the parent's (sorted) row list
__global__ void k_relmap(int32_t nsuper, const int32_t* __restrict__ sptr, const int64_t* __restrict__ rptr,
                         const int32_t* __restrict__ rows, const int32_t* __restrict__ sparent,
                         int32_t* __restrict__ relmap, int* __restrict__ bad) {
  const int32_t s = blockIdx.x;
  if (s >= nsuper) return;
  const int32_t p = sparent[s];
  const int64_t b0 = rptr[s] + (sptr[s + 1] - sptr[s]), b1 = rptr[s + 1];
  if (p < 0) return;
  const int64_t q0 = rptr[p], q1 = rptr[p + 1];
  const int32_t pf = sptr[p], pl = sptr[p + 1] - 1;  // the parent's own columns come first: rows[q0 + j] = pf + j
  for (int64_t k = b0 + threadIdx.x; k < b1; k += blockDim.x) {
    const int32_t i = rows[k];
    int64_t pos;
    if (i <= pl) {
      pos = i - pf;
      if (i < pf) atomicOr(bad, 1);
    } else {
      int64_t lo = q0 + (pl - pf + 1), hi = q1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rows[mid] < i)
          lo = mid + 1;
        else
          hi = mid;
      }
      if (lo >= q1 || rows[lo] != i) atomicOr(bad, 1);
      pos = lo - q0;
    }
    relmap[k] = (int32_t)pos;
  }
}

// amap[p] = arena slot of stored entry p (or -1 for the mirrored triangle); bad: bit 0 = entry outside the structure,
// bit 1 = STORAGE_LOWER entry above the diagonal, bit 2 = STORAGE_UPPER entry below the diagonal
__global__ void k_amap(int64_t n, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, int base,
                       int storage, const int32_t* __restrict__ iperm, const int32_t* __restrict__ snode,
                       const int32_t* __restrict__ sptr, const int64_t* __restrict__ rptr,
                       const int32_t* __restrict__ rows, const int64_t* __restrict__ foff, const int32_t* __restrict__ ld,
                       int64_t* __restrict__ amap, int* __restrict__ bad) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int32_t jc = iperm[c];
  for (int64_t p = colptr[c] - base; p < colptr[c + 1] - base; p++) {
    const int64_t r = rowval[p] - base;
    int32_t i = iperm[r], j = jc;
    if (storage == 0) {
      if (i < j) {
        amap[p] = -1;
        continue;
      }
    } else {
      if (storage == 1 && r < c) atomicOr(bad, 2);
      if (storage == 2 && r > c) atomicOr(bad, 4);
      if (i < j) {
        const int32_t t = i;
        i = j;
        j = t;
      }
    }
    const int32_t s = snode[j], f = sptr[s], l = sptr[s + 1] - 1;
    int64_t lr;
    if (i <= l) {
      lr = i - f;
    } else {
      int64_t lo = rptr[s] + (l - f + 1), hi = rptr[s + 1];
      const int64_t end = hi;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rows[mid] < i)
          lo = mid + 1;
        else
          hi = mid;
      }
      if (lo >= end || rows[lo] != i) {
        atomicOr(bad, 1);
        amap[p] = -1;
        continue;
      }
      lr = lo - rptr[s];
    }
    amap[p] = foff[s] + (int64_t)(j - f) * ld[s] + lr;
  }
}

inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

struct SymDevice::Impl {
  cudaStream_t st = nullptr;
  int64_t n = 0, nnz = 0;
  int base = 0;
  DBuf<int64_t> colptr, rowval, xadj;
  DBuf<int32_t> adj;
  DBuf<int> flags;
};

SymDevice::SymDevice(void* stream) : im(new Impl()) { im->st = (cudaStream_t)stream; }
SymDevice::~SymDevice() { delete im; }

// host-side fix-up of the (rare) rows that were too long for the per-thread sort
static void sort_long_rows(int64_t n, const std::vector<int64_t>& x, std::vector<int32_t>& a) {
  for (int64_t k = 0; k < n; k++)
    if (x[k + 1] - x[k] > SORT_MAX) std::sort(a.begin() + x[k], a.begin() + x[k + 1]);
}

std::string SymDevice::adjacency(int64_t n, const int64_t* colptr, const int64_t* rowval, int base,
                                 std::vector<int64_t>& xadj, std::vector<int32_t>& adj, bool& handled) {
  handled = false;
  Impl& I = *im;
  I.n = n;
  I.base = base;
  I.nnz = colptr[n] - base;
  if (I.nnz < 0) return "colptr is not monotone";
  SG_CU(I.colptr.upload(colptr, (size_t)n + 1, I.st));
  SG_CU(I.rowval.upload(rowval, (size_t)I.nnz, I.st));
  DBuf<int32_t> deg;
  SG_CU(deg.alloc((size_t)n));
  SG_CU(I.flags.alloc(4));
  SG_CU(cudaMemsetAsync(I.flags.p, 0, 4 * sizeof(int), I.st));
  SG_CU(cudaMemsetAsync(deg.p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(int32_t), I.st));
  if (n > 0) k_pattern_check<<<nblk(n), 256, 0, I.st>>>(n, I.colptr.p, I.rowval.p, base, I.nnz, deg.p, I.flags.p);
  SG_CU(cudaGetLastError());
  int flags = 0;
  std::vector<int32_t> hdeg((size_t)n);
  SG_CU(cudaMemcpyAsync(&flags, I.flags.p, sizeof(int), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaMemcpyAsync(hdeg.data(), deg.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaStreamSynchronize(I.st));
  if (flags) return "";  // malformed or not structurally symmetric: the host path validates / symmetrises and reports
  xadj.assign((size_t)n + 1, 0);
  for (int64_t i = 0; i < n; i++) xadj[i + 1] = xadj[i] + hdeg[i];
  adj.resize((size_t)xadj[n]);
  SG_CU(I.xadj.upload(xadj.data(), xadj.size(), I.st));
  SG_CU(I.adj.alloc(adj.size()));
  if (n > 0) k_adj_fill<<<nblk(n), 256, 0, I.st>>>(n, I.colptr.p, I.rowval.p, base, I.xadj.p, I.adj.p);
  SG_CU(cudaGetLastError());
  SG_CU(cudaMemcpyAsync(adj.data(), I.adj.p, adj.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaStreamSynchronize(I.st));
  handled = true;
  return "";
}

// (oxadj, oadj) = rows of the device-resident adjacency gathered through `src`, relabelled by `label`, optionally only
// labels above the row index, each row sorted
static std::string rows_map(SymDevice::Impl& I, const DBuf<int64_t>& dx, const DBuf<int32_t>& da, const std::vector<int32_t>& src,
                            const std::vector<int32_t>& label, bool filter_above, std::vector<int64_t>& oxadj,
                            std::vector<int32_t>& oadj, DBuf<int64_t>* keep_x, DBuf<int32_t>* keep_a) {
  const int64_t n = I.n;
  DBuf<int32_t> dsrc, dlabel, doadj_local;
  DBuf<int64_t> dcnt, doxadj_local;
  DBuf<int64_t>& doxadj = keep_x ? *keep_x : doxadj_local;
  DBuf<int32_t>& doadj = keep_a ? *keep_a : doadj_local;
  SG_CU(dsrc.upload(src.data(), src.size(), I.st));
  SG_CU(dlabel.upload(label.data(), label.size(), I.st));
  SG_CU(dcnt.alloc((size_t)n));
  if (n > 0) k_rows_map<<<nblk(n), 256, 0, I.st>>>(n, dx.p, da.p, dsrc.p, dlabel.p, filter_above, nullptr, nullptr, dcnt.p, nullptr);
  SG_CU(cudaGetLastError());
  std::vector<int64_t> cnt((size_t)n);
  SG_CU(cudaMemcpyAsync(cnt.data(), dcnt.p, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaStreamSynchronize(I.st));
  oxadj.assign((size_t)n + 1, 0);
  for (int64_t k = 0; k < n; k++) oxadj[k + 1] = oxadj[k] + cnt[k];
  oadj.resize((size_t)oxadj[n]);
  SG_CU(doxadj.upload(oxadj.data(), oxadj.size(), I.st));
  SG_CU(doadj.alloc(oadj.size()));
  SG_CU(cudaMemsetAsync(I.flags.p, 0, 4 * sizeof(int), I.st));
  if (n > 0) k_rows_map<<<nblk(n), 256, 0, I.st>>>(n, dx.p, da.p, dsrc.p, dlabel.p, filter_above, doxadj.p, doadj.p, nullptr, I.flags.p);
  SG_CU(cudaGetLastError());
  int nlong = 0;
  SG_CU(cudaMemcpyAsync(&nlong, I.flags.p, sizeof(int), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaMemcpyAsync(oadj.data(), doadj.p, oadj.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaStreamSynchronize(I.st));
  if (nlong > 0) {
    sort_long_rows(n, oxadj, oadj);
    if (keep_a) SG_CU(cudaMemcpyAsync(doadj.p, oadj.data(), oadj.size() * sizeof(int32_t), cudaMemcpyHostToDevice, I.st));
  }
  return "";
}

struct SymDevice::Perm {
  DBuf<int64_t> pxadj;
  DBuf<int32_t> padj;
};

std::string SymDevice::permuted_adjacency(const std::vector<int32_t>& perm_user, const std::vector<int32_t>& ipu,
                                          std::vector<int64_t>& pxadj, std::vector<int32_t>& padj) {
  if (!pm) pm = new Perm();
  return rows_map(*im, im->xadj, im->adj, perm_user, ipu, false, pxadj, padj, &pm->pxadj, &pm->padj);
}

std::string SymDevice::internal_adjacency(const std::vector<int32_t>& post, const std::vector<int32_t>& ipost,
                                          std::vector<int64_t>& ixadj, std::vector<int32_t>& iadj) {
  if (!pm) return "symbolic (GPU): internal_adjacency before permuted_adjacency";
  return rows_map(*im, pm->pxadj, pm->padj, post, ipost, true, ixadj, iadj, nullptr, nullptr);
}

std::string SymDevice::maps(Symbolic& S) {
  Impl& I = *im;
  delete pm;  // the adjacency copies are no longer needed
  pm = nullptr;
  const int64_t n = S.n;
  DBuf<int32_t> sptr, rows, sparent, relmap, iperm, snode, ld;
  DBuf<int64_t> rptr, foff, amap;
  SG_CU(sptr.upload(S.sptr.data(), S.sptr.size(), I.st));
  SG_CU(rptr.upload(S.rptr.data(), S.rptr.size(), I.st));
  SG_CU(rows.upload(S.rows.data(), S.rows.size(), I.st));
  SG_CU(sparent.upload(S.sparent.data(), S.sparent.size(), I.st));
  SG_CU(iperm.upload(S.iperm.data(), S.iperm.size(), I.st));
  SG_CU(snode.upload(S.snode.data(), S.snode.size(), I.st));
  SG_CU(ld.upload(S.ld.data(), S.ld.size(), I.st));
  SG_CU(foff.upload(S.foff.data(), S.foff.size(), I.st));
  SG_CU(relmap.alloc(S.rows.size()));
  SG_CU(amap.alloc((size_t)S.nnzA));
  SG_CU(cudaMemsetAsync(relmap.p, 0xff, std::max<size_t>(S.rows.size(), 1) * sizeof(int32_t), I.st));  // -1
  SG_CU(cudaMemsetAsync(I.flags.p, 0, 4 * sizeof(int), I.st));
  if (S.nsuper > 0) k_relmap<<<(unsigned)S.nsuper, 64, 0, I.st>>>(S.nsuper, sptr.p, rptr.p, rows.p, sparent.p, relmap.p, I.flags.p);
  SG_CU(cudaGetLastError());
  if (n > 0)
    k_amap<<<nblk(n), 256, 0, I.st>>>(n, I.colptr.p, I.rowval.p, I.base, S.storage, iperm.p, snode.p, sptr.p, rptr.p, rows.p,
                                      foff.p, ld.p, amap.p, I.flags.p + 1);
  SG_CU(cudaGetLastError());
  int flags[2] = {0, 0};
  S.relmap.resize(S.rows.size());
  S.amap.resize((size_t)S.nnzA);
  SG_CU(cudaMemcpyAsync(flags, I.flags.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaMemcpyAsync(S.relmap.data(), relmap.p, S.relmap.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaMemcpyAsync(S.amap.data(), amap.p, S.amap.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, I.st));
  SG_CU(cudaStreamSynchronize(I.st));
  if (flags[0]) return "internal error: child row missing from parent front";
  if (flags[1] & 2) return "STORAGE_LOWER matrix has an entry above the diagonal";
  if (flags[1] & 4) return "STORAGE_UPPER matrix has an entry below the diagonal";
  if (flags[1] & 1) return "internal error: matrix entry outside the symbolic structure";
  return "";
}

}  // namespace gmrfb
