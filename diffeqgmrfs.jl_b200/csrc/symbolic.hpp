// Host-side symbolic analysis for the supernodal multifrontal Cholesky.
// Everything here is integer graph work on the sparsity pattern; it produces the permutation,
// elimination tree, column counts, supernode partition, frontal layout and the index maps that the
// CUDA numeric phases (numeric.cu, solve.cu, selinv.cu) consume without further host computation.
//
// Replaces CHOLMOD analyze/analyze_p as reached from `cholesky(Symmetric(A); perm=p)`
// (reference: scripts/solve_burger.jl:147, scripts/darcy/solve_darcy_fem.jl:93).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gmrfb {

struct AnalyzeOptions {
  int ordering_kind = 2;  // GMRFB_ORDER_*
  int storage = 0;        // GMRFB_STORAGE_*
  int base = 0;
  int coord_dim = 0;
  const double* coords = nullptr;
  int nd_leaf = 0;
  int relax_small = 0;
  double relax_zeros = 0.0;
};

// One level of the supernodal elimination tree (height from the leaves).
struct Level {
  std::vector<int32_t> snodes;  // supernodes of this level, ascending
};

struct Symbolic {
  int64_t n = 0;
  int64_t nnzA = 0;  // stored entries of the analysed matrix (user's nzval length)
  int base = 0;
  int storage = 0;

  // Orderings (all 0-based).  perm_user: new->old, the CHOLMOD-style `.p`.  post: internal k -> position in
  // the perm_user ordering (an etree postorder).  perm: internal new->old.  iperm: old -> internal.
  std::vector<int32_t> perm_user, post, ipost, perm, iperm;
  // Elimination tree and column counts in the perm_user ordering (what CHOLMOD would report).
  std::vector<int32_t> parent_user, colcount_user;
  // Same in the internal (postordered) numbering.
  std::vector<int32_t> parent, colcount;

  int64_t nnzL = 0;         // Σ colcount
  double flops = 0;         // Σ colcount^2
  int64_t nnzL_stored = 0;  // supernodal trapezoids incl. relaxation zeros
  int64_t nnz_lower_A = 0;

  // Supernodes (internal numbering).
  int32_t nsuper = 0;
  std::vector<int32_t> sptr;     // [nsuper+1] first column
  std::vector<int32_t> snode;    // [n] supernode of a column
  std::vector<int32_t> sparent;  // [nsuper] parent supernode or -1
  std::vector<int64_t> rptr;     // [nsuper+1] offsets into rows / relmap
  std::vector<int32_t> rows;     // row structure of each front (its columns first, then sorted below-rows)
  std::vector<int32_t> relmap;   // same shape: for below-row i of J, its position in the parent's front (-1 for cols)
  std::vector<int32_t> ld;       // [nsuper] leading dimension of the front
  std::vector<int64_t> foff;     // [nsuper] offset (in doubles) of the front in the arena
  int64_t arena = 0;             // doubles in the frontal arena
  std::vector<int32_t> level;    // [nsuper]
  std::vector<Level> levels;
  std::vector<int32_t> child_ptr, child_idx;  // children lists of the supernodal tree (ascending)
  int32_t max_front = 0;

  // Scatter map: user's nz k -> arena offset of its frontal entry, or -1 if the entry is the mirrored
  // triangle (FULL storage) / outside the analysed triangle.
  std::vector<int64_t> amap;

  int front_order(int32_t s) const { return (int)(rptr[s + 1] - rptr[s]); }
  int ncols(int32_t s) const { return sptr[s + 1] - sptr[s]; }
};

// CUDA implementations of the data-parallel phases of the analysis (symbolic_gpu.cu): pattern validation + symmetric
// adjacency, permuted / internal adjacency, relmap and amap.  Every method returns "" on success; results are
// bit-identical to the host loops of analyze_pattern.
class SymDevice {
 public:
  explicit SymDevice(void* cuda_stream);
  ~SymDevice();
  SymDevice(const SymDevice&) = delete;
  SymDevice& operator=(const SymDevice&) = delete;
  // handled = false: the pattern is malformed or not structurally symmetric (the host path deals with it)
  std::string adjacency(int64_t n, const int64_t* colptr, const int64_t* rowval, int base, std::vector<int64_t>& xadj,
                        std::vector<int32_t>& adj, bool& handled);
  std::string permuted_adjacency(const std::vector<int32_t>& perm_user, const std::vector<int32_t>& ipu,
                                 std::vector<int64_t>& pxadj, std::vector<int32_t>& padj);
  std::string internal_adjacency(const std::vector<int32_t>& post, const std::vector<int32_t>& ipost,
                                 std::vector<int64_t>& ixadj, std::vector<int32_t>& iadj);
  std::string maps(Symbolic& S);  // S.relmap and S.amap from S.rows / S.sptr / S.foff / ...
  struct Impl;
  struct Perm;

 private:
  Impl* im = nullptr;
  Perm* pm = nullptr;
};

// Returns "" on success, else an error message.  colptr/rowval are `base`-based 64-bit CSC.  With `dev` the
// data-parallel phases run as CUDA kernels (the ordering, the elimination tree, the column counts, the supernode
// partition and the bottom-up row structures are sequential graph algorithms and stay on the host).
std::string analyze_pattern(int64_t n, const int64_t* colptr, const int64_t* rowval, const int64_t* perm,
                            const AnalyzeOptions& opt, Symbolic& S, SymDevice* dev = nullptr);

// Pieces exposed for tests.
void etree_lower(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                 std::vector<int32_t>& parent);
void postorder_tree(int32_t n, const std::vector<int32_t>& parent, std::vector<int32_t>& post);
void column_counts(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                   const std::vector<int32_t>& parent, const std::vector<int32_t>& post,
                   std::vector<int32_t>& colcount);
void amd_order(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj, std::vector<int32_t>& perm,
               int32_t nfree = -1);  // vertices >= nfree: halo (never eliminated); -1: none
void nested_dissection(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                       int leaf, int coord_dim, const double* coords, std::vector<int32_t>& perm,
                       bool amd_leaves = false);  // amd_leaves: halo-AMD inside the leaf subdomains (default leaf 200)

}  // namespace gmrfb
