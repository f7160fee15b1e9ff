// Panel (multi-right-hand-side) supernodal triangular solves: every right-hand side of a batch in ONE sweep over L.
//
// The reference's real multi-RHS workload is RBMCStrategy(50) (scripts/darcy/solve_darcy_gmrf-fem.jl:100,174,192: 50
// samples x = P' L^{-T} z per posterior) and batches of posterior samples (`rand`, :191).  The level-scheduled kernels of
// sparse_kernels.cu process at most SOLVE_NRC = 4 columns per pass and re-read the whole factor for every pass; here
// the right-hand sides are a node-major panel
//     X  (nr x n, leading dimension ldk: the nr values of one node are contiguous)          -> Arenas::p[1]
//     U_J (nr x r_J per supernode, same leading dimension; multifrontal update panels)      -> Arenas::p[2]
// so that, for the larger supernodes, a sweep is made of the same grouped tensor-core launches that factorise
// (plan.cpp, build_solve_mr_plans):
//     forward :  X_J <- X_J L11^{-T}  (blocked TRSM with the kept inverses of the 64 x 64 diagonal blocks)
//                U_J <- U_J - X_J L21'                                   (one grouped DMMA GEMM per level)
//     backward:  U_J <- X[:, below rows]   (gather);   X_J <- (X_J - U_J L21) L11^{-1}
// and the small supernodes (front order <= SMALL_FRONT_MAX: 85 % of the supernodes, a quarter of the factor) are
// solved by the kernels below: one CTA per supernode, the working panel v (d x nr) in shared memory, THREAD = one
// right-hand side (x a row group), so the substitution needs no reduction across threads, the factor entries are
// warp-uniform (broadcast) loads, and every factor entry is read once per sweep for all nr right-hand sides.
// All sums have a fixed order: results are bit-reproducible and independent of the batch composition.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"
#include "sparse_kernels.hpp"
#include "tasks.hpp"

namespace gmrfb {

namespace {

constexpr int MR_Q = 32;  // right-hand sides per CTA of the small-supernode kernels (one per lane)
constexpr int MR_G = 8;   // row groups = warps per CTA
constexpr int MR_CH = 8;  // factor columns staged in shared memory at a time (two elimination steps of four)
constexpr int MR_NT = MR_Q * MR_G;
constexpr int MR_PF = (MR_CH * SMALL_FRONT_MAX + MR_NT - 1) / MR_NT;  // prefetch registers per thread for one chunk

// ------------------------------------------------------------------------------ small supernodes ----
// One CTA = one supernode x 32 right-hand sides; lane = right-hand side, warp = row group.  The working panel v
// (d x 32) lives in shared memory.  The factor panel is streamed through shared memory in chunks of MR_CH columns
// (coalesced loads by the whole CTA, the next chunk prefetched into registers while the current one is applied), so
// the substitution itself never waits for global memory; 1 / L_cc is computed once per column while staging.
// Columns are eliminated four at a time: every thread solves the 4 x 4 triangle of its right-hand side in registers.

struct MrChunk {
  double pf[MR_PF];
};

// chunk k = columns [8k, 8k + 8) of the panel, rows [8k, d): element e = c * dlen + (i - 8k)
__device__ __forceinline__ void mr_chunk_load(MrChunk& ck, const double* __restrict__ F, int ld, int d, int s, int k) {
  const int r0 = k * MR_CH, dlen = d - r0, ncol = min(MR_CH, s - r0);
#pragma unroll
  for (int u = 0; u < MR_PF; u++) {
    const int e = threadIdx.x + u * MR_NT;
    const int c = e / dlen, i = e - c * dlen + r0;
    ck.pf[u] = (c < ncol && i >= r0 + c) ? F[i + (int64_t)(r0 + c) * ld] : 0.0;
  }
}
// Lc[c * dp + i] (absolute row i), rd[c] = 1 / L_cc (1 for padding columns)
__device__ __forceinline__ void mr_chunk_store(const MrChunk& ck, double* __restrict__ Lc, double* __restrict__ rd, int dp,
                                               int d, int s, int k) {
  const int r0 = k * MR_CH, dlen = d - r0, ncol = min(MR_CH, s - r0);
#pragma unroll
  for (int u = 0; u < MR_PF; u++) {
    const int e = threadIdx.x + u * MR_NT;
    const int c = e / dlen, i = e - c * dlen + r0;
    if (c < MR_CH) {
      Lc[c * dp + i] = ck.pf[u];
      if (i == r0 + c) rd[c] = (c < ncol) ? 1.0 / ck.pf[u] : 1.0;
    }
  }
}

// forward:  v = [X_J; 0] + children's update panels;  for c: y_c = v_c / L_cc, v_i -= L_ic y_c;  X_J = y, U_J = v[s:]
__global__ void __launch_bounds__(MR_NT) k_mr_fwd_small(const Task* __restrict__ tasks, Arenas ar,
                                                        const SnodeDesc* __restrict__ sd,
                                                        const int32_t* __restrict__ child_idx,
                                                        const int32_t* __restrict__ relmap, int nr, int ldk, int dp) {
  extern __shared__ __align__(16) double sm[];
  double* v = sm;                               // v[i * MR_Q + lane]
  double* Lc = v + (size_t)dp * MR_Q;           // two chunk buffers of MR_CH * dp
  double* rd = Lc + 2 * (size_t)MR_CH * dp;     // two sets of MR_CH reciprocal diagonals
  const SnodeDesc D = sd[tasks[blockIdx.x].aux0];
  const int d = D.d, s = D.s, r = d - s, ld = D.ld;
  const double* __restrict__ F = ar.p[0] + D.foff;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int q = blockIdx.y * MR_Q + lane;
  const bool qa = q < nr;
  double* __restrict__ X = ar.p[1] + (int64_t)D.col0 * ldk + q;
  double* __restrict__ U = ar.p[2] + D.uoff * ldk + q;
  MrChunk ck;
  mr_chunk_load(ck, F, ld, d, s, 0);
  for (int i = g; i < d; i += MR_G) v[i * MR_Q + lane] = (i < s && qa) ? X[(int64_t)i * ldk] : 0.0;
  mr_chunk_store(ck, Lc, rd, dp, d, s, 0);
  __syncthreads();
  for (int ci = 0; ci < D.nchild; ci++) {
    const SnodeDesc C = sd[child_idx[D.child0 + ci]];
    const int rc = C.d - C.s;
    const double* __restrict__ Uc = ar.p[2] + C.uoff * ldk + q;
    const int32_t* __restrict__ rel = relmap + C.rows_off + C.s;
    for (int i0 = g; i0 < rc; i0 += 4 * MR_G) {  // four independent loads in flight per thread
      int p[4];
      double u[4];
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int i = i0 + t * MR_G;
        p[t] = i < rc ? rel[i] : -1;
        u[t] = (i < rc && qa) ? Uc[(int64_t)i * ldk] : 0.0;
      }
#pragma unroll
      for (int t = 0; t < 4; t++)
        if (p[t] >= 0) v[p[t] * MR_Q + lane] += u[t];  // distinct rows within one child
    }
    __syncthreads();
  }
  const int nchunk = (s + MR_CH - 1) / MR_CH;
  for (int k = 0; k < nchunk; k++) {
    const double* __restrict__ L = Lc + (size_t)(k & 1) * MR_CH * dp;
    const double* __restrict__ rdk = rd + (k & 1) * MR_CH;
    const bool more = k + 1 < nchunk;
    if (more) mr_chunk_load(ck, F, ld, d, s, k + 1);
    const int nhalf = (k * MR_CH + 4 < s) ? 2 : 1;
    for (int h = 0; h < nhalf; h++) {
      const int c0 = k * MR_CH + 4 * h, nc = min(4, s - c0), cl = 4 * h;  // cl: column inside the chunk
      double y[4];
#pragma unroll
      for (int a = 0; a < 4; a++) {
        double t = (a < nc) ? v[(c0 + a) * MR_Q + lane] : 0.0;
#pragma unroll
        for (int b = 0; b < a; b++) t -= L[(cl + b) * dp + c0 + a] * y[b];
        y[a] = t * rdk[cl + a];
      }
      if (g == 0 && qa)
#pragma unroll
        for (int a = 0; a < 4; a++)
          if (a < nc) X[(int64_t)(c0 + a) * ldk] = y[a];
      const double* __restrict__ l0 = L + (size_t)cl * dp;
#pragma unroll 2
      for (int i = c0 + nc + g; i < d; i += MR_G) {
        double t = v[i * MR_Q + lane];
        t -= l0[i] * y[0];
        t -= l0[dp + i] * y[1];
        t -= l0[2 * dp + i] * y[2];
        t -= l0[3 * dp + i] * y[3];
        v[i * MR_Q + lane] = t;
      }
      if (h == nhalf - 1 && more) mr_chunk_store(ck, Lc + (size_t)((k + 1) & 1) * MR_CH * dp, rd + ((k + 1) & 1) * MR_CH, dp, d, s, k + 1);
      __syncthreads();
    }
  }
  if (qa)
    for (int i = g; i < r; i += MR_G) U[(int64_t)i * ldk] = v[(s + i) * MR_Q + lane];
}

// backward:  v = [X_J (= t_J); X[below rows]];  for c descending: x_c = (v_c - sum_{i>c} L_ic v_i) / L_cc;  X_J = x
// Four columns at a time: the row groups accumulate the four dot products over their rows, the partial sums are
// combined in group order by warp 0, which solves the transposed 4 x 4 triangle.
__global__ void __launch_bounds__(MR_NT) k_mr_bwd_small(const Task* __restrict__ tasks, Arenas ar,
                                                        const SnodeDesc* __restrict__ sd,
                                                        const int32_t* __restrict__ rows, int nr, int ldk, int dp) {
  extern __shared__ __align__(16) double sm[];
  double* v = sm;
  double* Lc = v + (size_t)dp * MR_Q;
  double* rd = Lc + 2 * (size_t)MR_CH * dp;
  double* red = rd + 2 * MR_CH;  // red[(g * 4 + a) * MR_Q + lane]
  const SnodeDesc D = sd[tasks[blockIdx.x].aux0];
  const int d = D.d, s = D.s, ld = D.ld;
  const double* __restrict__ F = ar.p[0] + D.foff;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int q = blockIdx.y * MR_Q + lane;
  const bool qa = q < nr;
  const double* __restrict__ Xall = ar.p[1] + q;
  double* __restrict__ X = ar.p[1] + (int64_t)D.col0 * ldk + q;
  const int32_t* __restrict__ rw = rows + D.rows_off;
  const int nchunk = (s + MR_CH - 1) / MR_CH;
  MrChunk ck;
  mr_chunk_load(ck, F, ld, d, s, nchunk - 1);
  for (int i0 = g; i0 < d; i0 += 4 * MR_G) {  // four independent (index, value) load pairs in flight per thread
    int64_t node[4];
    double x[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int i = i0 + t * MR_G;
      node[t] = i < s ? (int64_t)D.col0 + i : (i < d ? (int64_t)rw[i] : 0);
    }
#pragma unroll
    for (int t = 0; t < 4; t++) x[t] = (i0 + t * MR_G < d && qa) ? Xall[node[t] * ldk] : 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++)
      if (i0 + t * MR_G < d) v[(i0 + t * MR_G) * MR_Q + lane] = x[t];
  }
  mr_chunk_store(ck, Lc + (size_t)((nchunk - 1) & 1) * MR_CH * dp, rd + ((nchunk - 1) & 1) * MR_CH, dp, d, s, nchunk - 1);
  __syncthreads();
  for (int k = nchunk - 1; k >= 0; k--) {
    const double* __restrict__ L = Lc + (size_t)(k & 1) * MR_CH * dp;
    const double* __restrict__ rdk = rd + (k & 1) * MR_CH;
    const bool more = k > 0;
    if (more) mr_chunk_load(ck, F, ld, d, s, k - 1);
    const int nhalf = (k * MR_CH + 4 < s) ? 2 : 1;
    for (int h = nhalf - 1; h >= 0; h--) {
      const int c0 = k * MR_CH + 4 * h, nc = min(4, s - c0), cl = 4 * h;
      const double* __restrict__ l0 = L + (size_t)cl * dp;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
      for (int i = c0 + nc + g; i < d; i += MR_G) {
        const double vi = v[i * MR_Q + lane];
        a0 += l0[i] * vi;
        a1 += l0[dp + i] * vi;
        a2 += l0[2 * dp + i] * vi;
        a3 += l0[3 * dp + i] * vi;
      }
      red[(g * 4 + 0) * MR_Q + lane] = a0;
      red[(g * 4 + 1) * MR_Q + lane] = a1;
      red[(g * 4 + 2) * MR_Q + lane] = a2;
      red[(g * 4 + 3) * MR_Q + lane] = a3;
      __syncthreads();
      if (g == 0) {
        double x[4];
#pragma unroll
        for (int a = 3; a >= 0; a--) {
          double t = 0.0;
          if (a < nc) {
            t = v[(c0 + a) * MR_Q + lane];
#pragma unroll
            for (int gg = 0; gg < MR_G; gg++) t -= red[(gg * 4 + a) * MR_Q + lane];
          }
#pragma unroll
          for (int b = a + 1; b < 4; b++) t -= L[(cl + a) * dp + c0 + b] * x[b];  // L_ba = entry (c0+b, c0+a)
          x[a] = t * rdk[cl + a];
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
          if (a < nc) v[(c0 + a) * MR_Q + lane] = x[a];
      }
      if (h == 0 && more) mr_chunk_store(ck, Lc + (size_t)((k - 1) & 1) * MR_CH * dp, rd + ((k - 1) & 1) * MR_CH, dp, d, s, k - 1);
      __syncthreads();
    }
  }
  if (qa)
    for (int i = g; i < s; i += MR_G) X[(int64_t)i * ldk] = v[i * MR_Q + lane];
}

// ------------------------------------------------------------------------ larger supernodes: assembly ----
// Forward assembly of supernode J: U_J <- 0, then for every child c (fixed order) the rows of U_c are added into
// X[:, C_J] (rows mapping into J's columns) or U_J (rows mapping below).  CTA = (supernode, slice of 8 right-hand
// sides): a slice is touched by one CTA only, children are processed one after the other => deterministic.
__global__ void __launch_bounds__(256) k_mr_assemble(const Task* __restrict__ tasks, Arenas ar,
                                                     const SnodeDesc* __restrict__ sd,
                                                     const int32_t* __restrict__ child_idx,
                                                     const int32_t* __restrict__ relmap, int nr, int ldk, int nslice) {
  const int t = blockIdx.x / nslice, sl = blockIdx.x - t * nslice;
  const SnodeDesc D = sd[tasks[t].aux0];
  const int s = D.s, r = D.d - D.s;
  double* __restrict__ X = ar.p[1] + (int64_t)D.col0 * ldk;
  double* __restrict__ U = ar.p[2] + D.uoff * ldk;
  const int q = sl * 8 + (threadIdx.x & 7), i0 = threadIdx.x >> 3;
  const bool qa = q < nr;
  if (qa)
    for (int i = i0; i < r; i += 32) U[q + (int64_t)i * ldk] = 0.0;
  for (int ci = 0; ci < D.nchild; ci++) {
    __syncthreads();
    const SnodeDesc C = sd[child_idx[D.child0 + ci]];
    const int rc = C.d - C.s;
    const double* __restrict__ Uc = ar.p[2] + C.uoff * ldk;
    const int32_t* __restrict__ rel = relmap + C.rows_off + C.s;
    if (qa)
      for (int i = i0; i < rc; i += 32) {
        const int p = rel[i];
        double* dst = (p < s) ? X + q + (int64_t)p * ldk : U + q + (int64_t)(p - s) * ldk;
        *dst += Uc[q + (int64_t)i * ldk];
      }
  }
}

// Backward gather: U_J[:, i] = X[:, rows_J[s + i]].  CTA = 64 below-rows of one supernode.
__global__ void __launch_bounds__(256) k_mr_gather(const Task* __restrict__ tasks, int ntasks, Arenas ar,
                                                   const SnodeDesc* __restrict__ sd, const int32_t* __restrict__ rows,
                                                   int nr, int ldk) {
  const int tix = (ntasks == 1) ? 0 : reinterpret_cast<const int32_t*>(tasks + ntasks)[blockIdx.x];
  const Task T = tasks[tix];
  const SnodeDesc D = sd[T.aux0];
  const int s = D.s, r = D.d - D.s;
  const double* __restrict__ X = ar.p[1];
  double* __restrict__ U = ar.p[2] + D.uoff * ldk;
  const int32_t* __restrict__ rw = rows + D.rows_off + s;
  const int i_lo = (blockIdx.x - T.tile0) * 64, i_hi = min(r, i_lo + 64);
  const int q = threadIdx.x & 63, i0 = threadIdx.x >> 6;
  if (q >= nr) return;
  for (int i = i_lo + i0; i < i_hi; i += 4) U[q + (int64_t)i * ldk] = X[q + (int64_t)rw[i] * ldk];
}

// ----------------------------------------------------------------------------- panel in / out ----
// X[q + k*ldk] = src[perm[k] + q*lds]   (column-major n x nr  ->  node-major panel in the internal ordering)
__global__ void k_mr_perm_in(const double* __restrict__ src, int64_t lds, double* __restrict__ X, int ldk,
                             const int32_t* __restrict__ perm, int64_t n, int nr) {
  __shared__ double t[32][33];
  // tile of 32 nodes x 32 right-hand sides: reads run along the nodes, writes along the right-hand sides
  const int64_t k0 = (int64_t)blockIdx.x * 32;
  const int q0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t k = k0 + tx;
  const int64_t p = k < n ? perm[k] : 0;
  for (int j = ty; j < 32; j += 8) t[j][tx] = (k < n && q0 + j < nr) ? src[p + (int64_t)(q0 + j) * lds] : 0.0;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t kk = k0 + j;
    if (kk < n && q0 + tx < nr) X[(q0 + tx) + kk * ldk] = t[tx][j];
  }
}
// dst[perm[k] + q*ldd] = X[q + k*ldk] (+ add[perm[k]])
__global__ void k_mr_perm_out(const double* __restrict__ X, int ldk, double* __restrict__ dst, int64_t ldd,
                              const int32_t* __restrict__ perm, int64_t n, int nr, const double* __restrict__ add) {
  __shared__ double t[32][33];
  const int64_t k0 = (int64_t)blockIdx.x * 32;
  const int q0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t kk = k0 + j;
    t[j][tx] = (kk < n && q0 + tx < nr) ? X[(q0 + tx) + kk * ldk] : 0.0;
  }
  __syncthreads();
  const int64_t k = k0 + tx;
  if (k >= n) return;
  const int64_t p = perm[k];
  const double a = add ? add[p] : 0.0;
  for (int j = ty; j < 32; j += 8)
    if (q0 + j < nr) dst[p + (int64_t)(q0 + j) * ldd] = t[tx][j] + a;
}
// dst[(q0 + q) + perm[k]*ldd] = X[q + k*ldk]   (node-major -> node-major through a permutation, RBMC sample panel)
__global__ void k_mr_perm_nodemajor(const double* __restrict__ X, int ldk, double* __restrict__ dst, int64_t ldd,
                                    const int32_t* __restrict__ perm, int64_t n, int q0, int nr) {
  const int q = threadIdx.x & 63;
  const int64_t k = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 6);
  if (k >= n || q >= nr) return;
  dst[(q0 + q) + (int64_t)perm[k] * ldd] = X[q + k * ldk];
}

}  // namespace

// dynamic shared memory of the small-supernode panel kernels for fronts of order <= d (dp = d rounded up to even)
static int mr_small_smem(int dp) { return (dp * MR_Q + 2 * MR_CH * dp + 2 * MR_CH + MR_G * 4 * MR_Q) * (int)sizeof(double); }

cudaError_t mr_kernels_init() {
  const int smem = mr_small_smem((SMALL_FRONT_MAX + 1) & ~1);
  cudaError_t e = cudaFuncSetAttribute(k_mr_fwd_small, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_mr_bwd_small, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// Launch::smem of the small-supernode launches holds the largest front order of the launch (the shared-memory size
// depends on it, not on a byte count computed by the plan).
cudaError_t run_mr_launch(const Launch& L, const Task* t, const Arenas& ar, const LaunchAux& aux, cudaStream_t st) {
  const int nr = aux.nr, ldk = aux.ldk;
  switch (L.kind) {
    case LK_MR_FWD_SMALL: {
      const int dp = (L.smem + 1) & ~1;
      k_mr_fwd_small<<<dim3(L.grid, (nr + MR_Q - 1) / MR_Q), MR_NT, mr_small_smem(dp), st>>>(
          t, ar, aux.d_snodes, aux.d_child_idx, aux.d_relmap, nr, ldk, dp);
      break;
    }
    case LK_MR_BWD_SMALL: {
      const int dp = (L.smem + 1) & ~1;
      k_mr_bwd_small<<<dim3(L.grid, (nr + MR_Q - 1) / MR_Q), MR_NT, mr_small_smem(dp), st>>>(t, ar, aux.d_snodes, aux.d_rows,
                                                                                           nr, ldk, dp);
      break;
    }
    case LK_MR_ASSEMBLE: {
      const int nslice = (nr + 7) / 8;
      k_mr_assemble<<<L.grid * nslice, 256, 0, st>>>(t, ar, aux.d_snodes, aux.d_child_idx, aux.d_relmap, nr, ldk, nslice);
      break;
    }
    case LK_MR_GATHER:
      k_mr_gather<<<L.grid, 256, 0, st>>>(t, L.ntasks, ar, aux.d_snodes, aux.d_rows, nr, ldk);
      break;
    default:
      return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_mr_perm_in(const double* src, int64_t lds, double* X, int ldk, const int32_t* perm, int64_t n,
                              int nr, cudaStream_t st) {
  if (n <= 0 || nr <= 0) return cudaSuccess;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((nr + 31) / 32));
  k_mr_perm_in<<<grid, 256, 0, st>>>(src, lds, X, ldk, perm, n, nr);
  return cudaGetLastError();
}
cudaError_t launch_mr_perm_out(const double* X, int ldk, double* dst, int64_t ldd, const int32_t* perm, int64_t n,
                               int nr, const double* add, cudaStream_t st) {
  if (n <= 0 || nr <= 0) return cudaSuccess;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((nr + 31) / 32));
  k_mr_perm_out<<<grid, 256, 0, st>>>(X, ldk, dst, ldd, perm, n, nr, add);
  return cudaGetLastError();
}
cudaError_t launch_mr_perm_nodemajor(const double* X, int ldk, double* dst, int64_t ldd, const int32_t* perm,
                                     int64_t n, int q0, int nr, cudaStream_t st) {
  if (n <= 0 || nr <= 0) return cudaSuccess;
  k_mr_perm_nodemajor<<<(unsigned)((n + 3) / 4), 256, 0, st>>>(X, ldk, dst, ldd, perm, n, q0, nr);
  return cudaGetLastError();
}

}  // namespace gmrfb
