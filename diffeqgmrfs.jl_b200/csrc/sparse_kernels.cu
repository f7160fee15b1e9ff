// Bandwidth-bound kernels of the sparse path: supernodal triangular solves (level-scheduled, multifrontal
// update vectors => deterministic, atomic-free), permutations, SpMV / SpMM, RBMC variance accumulation and the
// fixed-pattern posterior-precision assembly.
#include <cuda_runtime.h>

#include <cstdint>

#include "sparse_kernels.hpp"
#include "tasks.hpp"

namespace gmrfb {

constexpr int DLD = 65;  // padded leading dimension of the 64x64 diagonal block in shared memory

// ------------------------------------------------------------------------------------- permutations ----
// dst[k + q*ldd] = src[perm[k] + q*lds]   (gather rows through perm)
__global__ void k_perm_gather(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd,
                              const int32_t* __restrict__ perm, int64_t n, int nrhs) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t p = perm[k];
  for (int q = 0; q < nrhs; q++) dst[k + q * ldd] = src[p + q * lds];
}
// dst[perm[k] + q*ldd] = src[k + q*lds] (+ add[perm[k]] if add)
__global__ void k_perm_scatter(const double* __restrict__ src, int64_t lds, double* __restrict__ dst, int64_t ldd,
                               const int32_t* __restrict__ perm, int64_t n, int nrhs,
                               const double* __restrict__ add) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t p = perm[k];
  double a = add ? add[p] : 0.0;
  for (int q = 0; q < nrhs; q++) dst[p + q * ldd] = src[k + q * lds] + a;
}

// ------------------------------------------------------------------------- supernodal triangular solves ----
// Level-scheduled, multifrontal-style solves.  The right-hand side lives in global memory (x: n x nr, internal
// ordering); every supernode J owns an update vector u_J (r_J x NRC) holding minus the accumulated
// contributions to its below-rows.  Large supernodes are spread over many CTAs: a level is processed as one
// assembly launch plus one launch per 64-column block step, and every factor entry is read exactly once with
// rows contiguous across threads.  No atomics: all reductions have a fixed order (bit-reproducible results).
//
// Solve task encoding (Task): a = front offset, lda = ld, M = d, N = s, K = block step, ldb = col0,
//   b = u offset (rows), c = partial-sum offset (doubles), ldc = number of row chunks of the R part,
//   aux0/aux1 = offset of the front's row list.

constexpr int FS_ROWS = 128;   // rows per CTA in the forward step
constexpr int BR_ROWS = 512;   // rows per CTA in the backward R-part reduction

__device__ __forceinline__ int find_task_s(const Task* __restrict__ tasks, int ntasks, int cta) {
  int lo = 0, hi = ntasks - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (tasks[mid].tile0 <= cta)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// x[C_J] += sum_children u_c (entries mapping into the columns), u_J = sum_children u_c (entries mapping below)
template <int NRC>
__global__ void __launch_bounds__(256) k_fwd_assemble(const SnodeDesc* __restrict__ sd,
                                                      const int32_t* __restrict__ list,
                                                      const int32_t* __restrict__ child_idx,
                                                      const int32_t* __restrict__ relmap, double* __restrict__ x,
                                                      int64_t ldx, double* __restrict__ uvec) {
  const SnodeDesc D = sd[list[blockIdx.x]];
  const int s = D.s, r = D.d - D.s;
  const int tid = threadIdx.x;
  double* __restrict__ uj = uvec + D.uoff * NRC;
  for (int i = tid; i < r * NRC; i += 256) uj[i] = 0.0;
  __syncthreads();
  for (int ci = 0; ci < D.nchild; ci++) {
    const SnodeDesc C = sd[child_idx[D.child0 + ci]];
    const int rc = C.d - C.s;
    const double* __restrict__ uc = uvec + C.uoff * NRC;
    const int32_t* __restrict__ rel = relmap + C.rows_off + C.s;
    for (int i = tid; i < rc; i += 256) {
      const int p = rel[i];
#pragma unroll
      for (int q = 0; q < NRC; q++) {
        const double v = uc[i + q * rc];
        if (p < s)
          x[D.col0 + p + q * ldx] += v;
        else
          uj[(p - s) + q * r] += v;
      }
    }
    __syncthreads();
  }
}

// Block step k of the forward solve of supernode J: every CTA solves the 64x64 diagonal block for y_k
// (redundantly, identical arithmetic), CTA 0 publishes y_k, and each CTA updates its 128 rows below the block.
template <int NRC>
__global__ void __launch_bounds__(FS_ROWS) k_fwd_step(const Task* __restrict__ tasks, int ntasks,
                                                      const double* __restrict__ F, double* __restrict__ x,
                                                      double* __restrict__ ysol, int64_t ldx,
                                                      double* __restrict__ uvec, int nr,
                                                      const double* __restrict__ dinv) {
  __shared__ double Ld[64 * DLD];
  __shared__ double invd[64];
  __shared__ double yk[NRC][64];
  const int tix = find_task_s(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int chunk = blockIdx.x - T.tile0;
  const int d = T.M, s = T.N, ld = T.lda, col0 = T.ldb;
  const int k0 = T.K * 64;
  const int nb = min(64, s - k0);
  const int r = d - s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ Fj = F + T.a;
  // this thread's row of the block below the diagonal: all loads are issued first, so that their latency overlaps
  // the (sequential) solve of the diagonal block
  const int row = k0 + nb + chunk * FS_ROWS + tid;
  double f[64];
  if (row < d) {
    const double* __restrict__ fr = Fj + row + (int64_t)k0 * ld;
#pragma unroll
    for (int c = 0; c < 64; c++) f[c] = (c < nb) ? fr[(int64_t)c * ld] : 0.0;
  }
  // the factorisation kept W_kk = L_kk^{-1} of this diagonal block (task field alpha): y_k = W_kk x_k is a parallel
  // product instead of a 64-step substitution chain
  const int64_t woff = __double_as_longlong(T.alpha);
  if (woff >= 0) {
    const double* __restrict__ Wk = dinv + woff + (int64_t)T.K * DINV_SLOT;  // nb x nb, leading dimension 64
    for (int e = tid; e < 64 * 64; e += FS_ROWS) {
      const int i = e & 63, j = e >> 6;
      Ld[j * DLD + i] = (i < nb && j <= i) ? Wk[i + j * 64] : 0.0;
    }
    for (int e = tid; e < NRC * 64; e += FS_ROWS) {
      const int q = e >> 6, j = e & 63;
      yk[q][j] = (q < nr && j < nb) ? x[col0 + k0 + j + q * ldx] : 0.0;  // x_k staged; overwritten by y_k below
    }
    __syncthreads();
    double part[NRC];
#pragma unroll
    for (int q = 0; q < NRC; q++) part[q] = 0.0;
    const int i = tid & 63, j0 = (tid >> 6) * 32;
    for (int j = j0; j < j0 + 32; j++) {
      const double wv = Ld[j * DLD + i];
#pragma unroll
      for (int q = 0; q < NRC; q++) part[q] += wv * yk[q][j];
    }
    __syncthreads();  // all reads of x_k done
    if (tid >= 64)
#pragma unroll
      for (int q = 0; q < NRC; q++) Ld[q * DLD + i] = part[q];  // upper halves parked in Ld (free now)
    __syncthreads();
    if (tid < 64) {
#pragma unroll
      for (int q = 0; q < NRC; q++) {
        const double yv = part[q] + Ld[q * DLD + i];
        yk[q][i] = yv;
        if (chunk == 0 && q < nr && i < nb) ysol[col0 + k0 + i + q * ldx] = yv;
      }
    }
    __syncthreads();
  } else {
  for (int e = tid; e < nb * nb; e += FS_ROWS) {
    int i = e % nb, j = e / nb;
    if (i >= j) Ld[j * DLD + i] = Fj[(k0 + i) + (int64_t)(k0 + j) * ld];
  }
  __syncthreads();
  if (tid < nb) invd[tid] = 1.0 / Ld[tid * DLD + tid];
  __syncthreads();
  if (warp < NRC) {
    double v0 = 0.0, v1 = 0.0;
    if (warp < nr) {
      const double* xq = x + col0 + k0 + warp * ldx;
      if (lane < nb) v0 = xq[lane];
      if (lane + 32 < nb) v1 = xq[lane + 32];
      for (int c = 0; c < nb; c++) {
        double yc = __shfl_sync(0xffffffffu, (c < 32) ? v0 : v1, c & 31) * invd[c];
        if (c < 32) {
          if (lane == c) v0 = yc;
          if (lane > c && lane < nb) v0 -= Ld[c * DLD + lane] * yc;
        } else if (lane == c - 32) {
          v1 = yc;
        }
        if (lane + 32 > c && lane + 32 < nb) v1 -= Ld[c * DLD + lane + 32] * yc;
      }
    }
    yk[warp][lane] = v0;
    yk[warp][lane + 32] = v1;
    if (chunk == 0 && warp < nr) {
      // published into a separate buffer: the other CTAs of this launch still read w_k from x
      double* yq = ysol + col0 + k0 + warp * ldx;
      if (lane < nb) yq[lane] = v0;
      if (lane + 32 < nb) yq[lane + 32] = v1;
    }
  }
  __syncthreads();
  }
  if (row >= d) return;
  double acc[NRC];
#pragma unroll
  for (int q = 0; q < NRC; q++) acc[q] = 0.0;
#pragma unroll
  for (int c = 0; c < 64; c++)
#pragma unroll
    for (int q = 0; q < NRC; q++) acc[q] += f[c] * yk[q][c];
  if (row < s) {
#pragma unroll
    for (int q = 0; q < NRC; q++)
      if (q < nr) x[col0 + row + q * ldx] -= acc[q];
  } else {
    double* __restrict__ uj = uvec + T.b * NRC;
#pragma unroll
    for (int q = 0; q < NRC; q++) uj[(row - s) + q * r] -= acc[q];
  }
}

// ------------------------------------------------------------------ wide supernodes: full inverse ----
// Supernodes with many columns (the top separators: up to ~3000 columns on the 1M-node mesh) put s/64 dependent block
// steps on the chain of every sweep - 128 launches of 10-20 us with a handful of CTAs each.  For those the
// factorisation keeps the full inverse W_J = L_JJ^{-1} (recursive-doubling TRTRI, the same product the selected
// inversion needs), and a sweep over J becomes two bandwidth-bound products with no dependency inside the supernode:
//   forward : y_J = W_J x_J,            u_J -= L21 y_J
//   backward: x_J = W_J' (t_J - L21' x_R)        (the R-part reduction stays k_bwd_rpart)
// The explicit inverse is only used when the factorisation measured cond_1(L_JJ) = |L_JJ|_1 |W_J|_1 below a
// threshold (api.cu); otherwise the block-step path above runs.
// Wide task encoding: the solve-task fields, with K = leading dimension of W_J and alpha = bits of W_J's offset.
constexpr int WG_ROWS = 32;   // rows per CTA of the row-oriented products
constexpr int WG_CH = 128;    // vector entries staged per chunk

// out[i] = sum_j A[i, j] v[j] for WG_ROWS rows per CTA: lanes = rows (contiguous in memory), the 8 warps take the
// columns j = w (mod 8) of every staged chunk (8 independent loads in flight per thread: the product is latency bound
// on grids of a few dozen CTAs); fixed-order reduction over the warps.
//   TRI : A = W_J (lower triangular), v = x[C_J],  y[C_J] = result                (forward, first product)
//   !TRI: A = L21 of the front,       v = y[C_J],  u_J -= result                  (forward, second product)
constexpr int WG_WARPS = 8;
template <int NRC, bool TRI>
__global__ void __launch_bounds__(32 * WG_WARPS) k_wide_gemv(const Task* __restrict__ tasks, int ntasks, const double* __restrict__ F,
                                                   const double* __restrict__ Wf, const double* __restrict__ v,
                                                   double* __restrict__ ysol, int64_t ldx, double* __restrict__ uvec,
                                                   int nr) {
  __shared__ double vs[NRC][WG_CH];
  __shared__ double red[WG_WARPS - 1][WG_ROWS][NRC];
  const int tix = find_task_s(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int d = T.M, s = T.N, col0 = T.ldb;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = (blockIdx.x - T.tile0) * WG_ROWS, i = i0 + lane;
  const int nrows = TRI ? s : d - s;
  const int lda = TRI ? T.K : T.lda;
  const double* __restrict__ A = TRI ? Wf + __double_as_longlong(T.alpha) : F + T.a + s;
  const int ncols = TRI ? min(s, i0 + WG_ROWS) : s;
  double acc[NRC];
#pragma unroll
  for (int q = 0; q < NRC; q++) acc[q] = 0.0;
  for (int c0 = 0; c0 < ncols; c0 += WG_CH) {
    __syncthreads();  // the previous chunk has been consumed
    for (int e = tid; e < NRC * WG_CH; e += 32 * WG_WARPS) {
      const int q = e / WG_CH, j = e % WG_CH;
      vs[q][j] = (q < nr && c0 + j < ncols) ? v[col0 + c0 + j + q * ldx] : 0.0;
    }
    __syncthreads();
    if (i < nrows) {
      const int jn = min(WG_CH, ncols - c0);
      const double* __restrict__ ap = A + i + (int64_t)c0 * lda;
      // all loads of the chunk are issued before the first multiply-add (WG_CH / WG_WARPS independent loads in flight)
      constexpr int U = WG_CH / WG_WARPS;
      double a[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int j = warp + u * WG_WARPS;
        a[u] = (j < jn && !(TRI && c0 + j > i)) ? ap[(int64_t)j * lda] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int j = warp + u * WG_WARPS;
#pragma unroll
        for (int q = 0; q < NRC; q++) acc[q] += a[u] * vs[q][j];
      }
    }
  }
  if (warp > 0)
#pragma unroll
    for (int q = 0; q < NRC; q++) red[warp - 1][lane][q] = acc[q];
  __syncthreads();
  if (warp == 0 && i < nrows) {
    const int r = d - s;
    double* __restrict__ uj = uvec + T.b * NRC;
#pragma unroll
    for (int q = 0; q < NRC; q++) {
      double val = acc[q];
#pragma unroll
      for (int w = 0; w < WG_WARPS - 1; w++) val += red[w][lane][q];
      if (TRI) {
        if (q < nr) ysol[col0 + i + q * ldx] = val;
      } else {
        uj[i + q * r] -= val;
      }
    }
  }
}

// x[C_J][j] = sum_{i >= j} W_J[i, j] t'[i],  t' = t[C_J] - sum over the R-part chunks (fixed order).  One warp per
// column (rows contiguous across lanes), 4 columns per CTA; t' is staged chunk by chunk.
template <int NRC>
__global__ void __launch_bounds__(128) k_wide_trmv_t(const Task* __restrict__ tasks, int ntasks,
                                                     const double* __restrict__ Wf, const double* __restrict__ t,
                                                     double* __restrict__ xsol, int64_t ldx,
                                                     const double* __restrict__ partial, int nr) {
  __shared__ double ts[NRC][WG_CH];
  const int tix = find_task_s(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int s = T.N, col0 = T.ldb, ldw = T.K, nchunk = T.ldc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = (blockIdx.x - T.tile0) * 4, j = j0 + warp;
  const double* __restrict__ W = Wf + __double_as_longlong(T.alpha);
  const double* __restrict__ pj = partial + T.c;
  double acc[NRC];
#pragma unroll
  for (int q = 0; q < NRC; q++) acc[q] = 0.0;
  for (int c0 = (j0 / WG_CH) * WG_CH; c0 < s; c0 += WG_CH) {
    __syncthreads();
    for (int e = tid; e < NRC * WG_CH; e += 128) {
      const int q = e / WG_CH, ii = e % WG_CH, i = c0 + ii;
      double val = 0.0;
      if (q < nr && i < s) {
        val = t[col0 + i + q * ldx];
        for (int ch = 0; ch < nchunk; ch++) val -= pj[(int64_t)ch * s * NRC + (int64_t)i * NRC + q];
      }
      ts[q][ii] = val;
    }
    __syncthreads();
    if (j < s) {
      const double* __restrict__ wc = W + (int64_t)j * ldw + c0;
      double a[WG_CH / 32];
#pragma unroll
      for (int u = 0; u < WG_CH / 32; u++) {
        const int ii = lane + 32 * u, i = c0 + ii;
        a[u] = (i >= j && i < s) ? wc[ii] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < WG_CH / 32; u++) {
        const int ii = lane + 32 * u;
#pragma unroll
        for (int q = 0; q < NRC; q++) acc[q] += a[u] * ts[q][ii];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NRC; q++) {
    double val = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if (lane == 0 && j < s && q < nr) xsol[col0 + j + q * ldx] = val;
  }
}

// norms[2 w] = max(norms[2 w], |L_JJ|_1 over this CTA's columns), norms[2 w + 1] likewise for W_J (w = T.aux0, the
// index of the wide supernode): one warp per column (rows contiguous across lanes), 8 columns per CTA.  Non-negative
// doubles order like their bit patterns, so the maximum is an integer atomicMax (order independent => deterministic).
__global__ void __launch_bounds__(256) k_wide_norms(const Task* __restrict__ tasks, int ntasks, const double* __restrict__ F,
                                                    const double* __restrict__ Wf, double* __restrict__ norms) {
  const int tix = find_task_s(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int s = T.N, ld = T.lda, ldw = T.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = (blockIdx.x - T.tile0) * 8 + warp;
  if (j >= s) return;
  const double* __restrict__ lp = F + T.a + (int64_t)j * ld;
  const double* __restrict__ wp = Wf + __double_as_longlong(T.alpha) + (int64_t)j * ldw;
  double sl = 0.0, sw = 0.0;
#pragma unroll 4
  for (int i = j + lane; i < s; i += 32) {
    sl += fabs(lp[i]);
    sw += fabs(wp[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  if (lane < 2) {
    double m = lane == 0 ? sl : sw;
    if (!(m == m)) m = __longlong_as_double(0x7ff0000000000000LL);  // NaN counts as infinite
    atomicMax(reinterpret_cast<unsigned long long*>(norms) + 2 * T.aux0 + lane, (unsigned long long)__double_as_longlong(m));
  }
}

// ------------------------------------------------------------------ fused solves of small supernodes ----
// Supernodes whose front has order d <= SOLVE_SMALL_MAX (the large majority: 31 of 38 thousand on the 1M-node mesh,
// a quarter of the factor's bytes) are solved by ONE WARP each, the whole panel [L11; L21] streamed once, column
// by column, with the working vector (x_J on top of u_J) in shared memory:
//   forward : v = [x_J; 0] + contributions of the children's update vectors (fixed order);
//             for c = 0..s-1:  y_c = v_c / L_cc,  v_i -= L_ic y_c (i > c);      u_J = v[s..d)
//   backward: v = [t_J; x[rows of R]];
//             for c = s-1..0:  x_c = (v_c - sum_{i>c} L_ic v_i) / L_cc
// This replaces, for these supernodes, the assemble / block-step / R-part launches, whose CTAs spent their time on a
// 64-step substitution executed by one of their four warps.  Column c+1 is loaded while column c is applied.
constexpr int SS_ROWS = (SOLVE_SMALL_MAX + 31) / 32;  // front rows per lane
constexpr int SS_WARPS = 4;

template <int NRC>
__global__ void __launch_bounds__(32 * SS_WARPS) k_fwd_small(const SnodeDesc* __restrict__ sd,
                                                             const int32_t* __restrict__ list, int count,
                                                             const int32_t* __restrict__ child_idx,
                                                             const int32_t* __restrict__ relmap,
                                                             const double* __restrict__ F, double* __restrict__ w,
                                                             double* __restrict__ ysol, int64_t ldx,
                                                             double* __restrict__ uvec, int nr) {
  __shared__ double sv[SS_WARPS][NRC][SOLVE_SMALL_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * SS_WARPS + warp;
  if (idx >= count) return;
  const SnodeDesc D = sd[list[idx]];
  const int d = D.d, s = D.s, r = d - s, ld = D.ld;
  const double* __restrict__ Fj = F + D.foff;
  double(*v)[SOLVE_SMALL_MAX] = sv[warp];
  // first column's values are requested before anything else
  double lnext[SS_ROWS];
#pragma unroll
  for (int t = 0; t < SS_ROWS; t++) {
    const int i = lane + 32 * t;
    lnext[t] = (i < d) ? Fj[i] : 0.0;
  }
  for (int i = lane; i < d; i += 32)
#pragma unroll
    for (int q = 0; q < NRC; q++) v[q][i] = (i < s && q < nr) ? w[D.col0 + i + q * ldx] : 0.0;
  __syncwarp();
  for (int ci = 0; ci < D.nchild; ci++) {
    const SnodeDesc C = sd[child_idx[D.child0 + ci]];
    const int rc = C.d - C.s;
    const double* __restrict__ uc = uvec + C.uoff * NRC;
    const int32_t* __restrict__ rel = relmap + C.rows_off + C.s;
    for (int i = lane; i < rc; i += 32) {
      const int p = rel[i];  // distinct rows of this front: no two lanes hit the same entry
#pragma unroll
      for (int q = 0; q < NRC; q++) v[q][p] += uc[i + q * rc];
    }
    __syncwarp();
  }
  for (int c = 0; c < s; c++) {
    double lcur[SS_ROWS];
#pragma unroll
    for (int t = 0; t < SS_ROWS; t++) lcur[t] = lnext[t];
    if (c + 1 < s) {
      const double* __restrict__ fc = Fj + (int64_t)(c + 1) * ld;
#pragma unroll
      for (int t = 0; t < SS_ROWS; t++) {
        const int i = lane + 32 * t;
        lnext[t] = (i > c && i < d) ? fc[i] : 0.0;
      }
    }
    double diag = 0.0;  // L_cc sits in slot c / 32 of lane c % 32 (select chain: no dynamic register indexing)
#pragma unroll
    for (int t = 0; t < SS_ROWS; t++)
      if (t == (c >> 5)) diag = lcur[t];
    const double inv = 1.0 / __shfl_sync(0xffffffffu, diag, c & 31);
    double yc[NRC];
#pragma unroll
    for (int q = 0; q < NRC; q++) yc[q] = v[q][c] * inv;
#pragma unroll
    for (int t = 0; t < SS_ROWS; t++) {
      const int i = lane + 32 * t;
      if (i > c && i < d)
#pragma unroll
        for (int q = 0; q < NRC; q++) v[q][i] -= lcur[t] * yc[q];
    }
    if (lane == 0)
#pragma unroll
      for (int q = 0; q < NRC; q++)
        if (q < nr) ysol[D.col0 + c + q * ldx] = yc[q];
    __syncwarp();
  }
  double* __restrict__ uj = uvec + D.uoff * NRC;
  for (int i = lane; i < r; i += 32)
#pragma unroll
    for (int q = 0; q < NRC; q++) uj[i + q * r] = v[q][s + i];
}

template <int NRC>
__global__ void __launch_bounds__(32 * SS_WARPS) k_bwd_small(const SnodeDesc* __restrict__ sd,
                                                             const int32_t* __restrict__ list, int count,
                                                             const int32_t* __restrict__ rows,
                                                             const double* __restrict__ F, const double* __restrict__ tv,
                                                             double* __restrict__ xsol, int64_t ldx, int nr) {
  __shared__ double sv[SS_WARPS][NRC][SOLVE_SMALL_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * SS_WARPS + warp;
  if (idx >= count) return;
  const SnodeDesc D = sd[list[idx]];
  const int d = D.d, s = D.s, ld = D.ld;
  const double* __restrict__ Fj = F + D.foff;
  const int32_t* __restrict__ rw = rows + D.rows_off;
  double(*v)[SOLVE_SMALL_MAX] = sv[warp];
  double lnext[SS_ROWS];
  {
    const double* __restrict__ fc = Fj + (int64_t)(s - 1) * ld;
#pragma unroll
    for (int t = 0; t < SS_ROWS; t++) {
      const int i = lane + 32 * t;
      lnext[t] = (i >= s - 1 && i < d) ? fc[i] : 0.0;
    }
  }
  for (int i = lane; i < d; i += 32) {
    const int64_t g = (i < s) ? (int64_t)D.col0 + i : (int64_t)rw[i];
#pragma unroll
    for (int q = 0; q < NRC; q++) v[q][i] = (q < nr) ? (i < s ? tv[g + q * ldx] : xsol[g + q * ldx]) : 0.0;
  }
  __syncwarp();
  for (int c = s - 1; c >= 0; c--) {
    double lcur[SS_ROWS];
#pragma unroll
    for (int t = 0; t < SS_ROWS; t++) lcur[t] = lnext[t];
    if (c > 0) {
      const double* __restrict__ fc = Fj + (int64_t)(c - 1) * ld;
#pragma unroll
      for (int t = 0; t < SS_ROWS; t++) {
        const int i = lane + 32 * t;
        lnext[t] = (i >= c - 1 && i < d) ? fc[i] : 0.0;
      }
    }
    double dot[NRC];
#pragma unroll
    for (int q = 0; q < NRC; q++) dot[q] = 0.0;
#pragma unroll
    for (int t = 0; t < SS_ROWS; t++) {
      const int i = lane + 32 * t;
      if (i > c && i < d)
#pragma unroll
        for (int q = 0; q < NRC; q++) dot[q] += lcur[t] * v[q][i];
    }
#pragma unroll
    for (int q = 0; q < NRC; q++)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
    double diag = 0.0;
#pragma unroll
    for (int t = 0; t < SS_ROWS; t++)
      if (t == (c >> 5)) diag = lcur[t];
    const double inv = 1.0 / __shfl_sync(0xffffffffu, diag, c & 31);
    if (lane == 0)
#pragma unroll
      for (int q = 0; q < NRC; q++) v[q][c] = (v[q][c] - dot[q]) * inv;
    __syncwarp();
  }
  for (int i = lane; i < s; i += 32)
#pragma unroll
    for (int q = 0; q < NRC; q++)
      if (q < nr) xsol[D.col0 + i + q * ldx] = v[q][i];
}

// Backward, R part: partial[chunk][c][q] = sum_{rows of the chunk} L21[row, c] * x[rows[row]][q] for the 64 columns
// of block kb.  One CTA = 64 columns x BR_ROWS rows; fixed-order reductions (shuffle tree, then warps in order).
template <int NRC>
__global__ void __launch_bounds__(128) k_bwd_rpart(const Task* __restrict__ tasks, int ntasks,
                                                   const double* __restrict__ F, const int32_t* __restrict__ rows,
                                                   const double* __restrict__ x, int64_t ldx,
                                                   double* __restrict__ partial, int nr) {
  __shared__ double red[4][64][NRC];
  const int tix = find_task_s(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int local = blockIdx.x - T.tile0;
  const int d = T.M, s = T.N, ld = T.lda;
  const int r = d - s;
  const int nchunk = T.ldc;
  const int kb = local / nchunk, chunk = local % nchunk;
  const int k0 = kb * 64;
  const int nb = min(64, s - k0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ Fj = F + T.a;
  const int32_t* __restrict__ rw = rows + (((int64_t)T.aux1 << 32) | (uint32_t)T.aux0) + s;
  // each thread owns up to BR_ROWS/128 rows of the chunk
  constexpr int RPT = BR_ROWS / 128;
  double xr[RPT][NRC];
  int rloc[RPT];
#pragma unroll
  for (int t = 0; t < RPT; t++) {
    const int rr = chunk * BR_ROWS + t * 128 + tid;
    rloc[t] = rr < r ? rr : -1;
    const int64_t g = rr < r ? rw[rr] : 0;
#pragma unroll
    for (int q = 0; q < NRC; q++) xr[t][q] = (rr < r && q < nr) ? x[g + q * ldx] : 0.0;
  }
  for (int c0 = 0; c0 < nb; c0 += 8) {
    double acc[8][NRC];
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int q = 0; q < NRC; q++) acc[u][q] = 0.0;
#pragma unroll
    for (int t = 0; t < RPT; t++) {
      if (rloc[t] < 0) continue;
      const double* __restrict__ fr = Fj + (s + rloc[t]) + (int64_t)(k0 + c0) * ld;
      double f[8];
#pragma unroll
      for (int u = 0; u < 8; u++) f[u] = (c0 + u < nb) ? fr[(int64_t)u * ld] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
        for (int q = 0; q < NRC; q++) acc[u][q] += f[u] * xr[t][q];
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int q = 0; q < NRC; q++) {
        double v = acc[u][q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][c0 + u][q] = v;
      }
  }
  __syncthreads();
  double* __restrict__ pj = partial + T.c + (int64_t)chunk * s * NRC;
  for (int e = tid; e < nb * NRC; e += 128) {
    const int c = e / NRC, q = e % NRC;
    pj[(int64_t)(k0 + c) * NRC + q] = ((red[0][c][q] + red[1][c][q]) + red[2][c][q]) + red[3][c][q];
  }
}

// Block step k (descending) of the backward solve: t_k = x_k - sum_chunks partial - (updates already applied by
// later blocks); solve L_kk' x_k = t_k; CTA j < k applies x[C_j] -= L[k-rows, j-cols]' x_k, CTA j == k publishes x_k.
template <int NRC>
__global__ void __launch_bounds__(128) k_bwd_step(const Task* __restrict__ tasks, int ntasks,
                                                  const double* __restrict__ F, double* __restrict__ x,
                                                  double* __restrict__ xsol, int64_t ldx,
                                                  const double* __restrict__ partial, int nr,
                                                  const double* __restrict__ dinv) {
  __shared__ double Ld[64 * DLD];
  __shared__ double invd[64];
  __shared__ double xk[NRC][64];
  __shared__ double half[64][NRC];
  const int tix = find_task_s(tasks, ntasks, blockIdx.x);
  const Task T = tasks[tix];
  const int j = blockIdx.x - T.tile0;
  const int s = T.N, ld = T.lda, col0 = T.ldb;
  const int k = T.K, k0 = k * 64;
  const int nb = min(64, s - k0);
  const int nchunk = T.ldc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* __restrict__ Fj = F + T.a;
  // update role (j < k): thread = (column cc of block j, row half hf) of L[k-rows, j-cols]; loads issued up front
  const int cc = tid & 63, hf = tid >> 6;
  const int r0 = hf * 32;
  double f[32];
  if (j != k) {
    const double* __restrict__ fc = Fj + k0 + (int64_t)(j * 64 + cc) * ld;
#pragma unroll
    for (int u = 0; u < 32; u++) f[u] = (r0 + u < nb) ? fc[r0 + u] : 0.0;
  }
  const int64_t woff = __double_as_longlong(T.alpha);
  if (woff >= 0) {
    // x_k = W_kk' t_k with the stored inverse of the diagonal block (see k_fwd_step)
    const double* __restrict__ Wk = dinv + woff + (int64_t)k * DINV_SLOT;
    for (int e = tid; e < 64 * 64; e += 128) {
      const int i = e & 63, jj = e >> 6;
      Ld[jj * DLD + i] = (i < nb && jj <= i) ? Wk[i + jj * 64] : 0.0;
    }
    // t_k = x_k - sum over the R-part chunks (fixed order)
    for (int e = tid; e < NRC * 64; e += 128) {
      const int q = e >> 6, i = e & 63;
      double v = 0.0;
      if (q < nr && i < nb) {
        v = x[col0 + k0 + i + q * ldx];
        const double* __restrict__ pj = partial + T.c;
        for (int ch = 0; ch < nchunk; ch++) v -= pj[(int64_t)ch * s * NRC + (int64_t)(k0 + i) * NRC + q];
      }
      xk[q][i] = v;
    }
    __syncthreads();
    double part[NRC];
#pragma unroll
    for (int q = 0; q < NRC; q++) part[q] = 0.0;
    const int jc = tid & 63, i0 = (tid >> 6) * 32;
    for (int i = i0; i < i0 + 32; i++) {
      const double wv = Ld[jc * DLD + i];  // W[i, jc]: column jc of W, conflict-free through the padding
#pragma unroll
      for (int q = 0; q < NRC; q++) part[q] += wv * xk[q][i];
    }
    __syncthreads();
    if (tid >= 64)
#pragma unroll
      for (int q = 0; q < NRC; q++) half[jc][q] = part[q];
    __syncthreads();
    if (tid < 64) {
#pragma unroll
      for (int q = 0; q < NRC; q++) {
        const double xv = part[q] + half[jc][q];
        xk[q][jc] = xv;
        if (j == k && q < nr && jc < nb) xsol[col0 + k0 + jc + q * ldx] = xv;
      }
    }
    __syncthreads();
  } else {
  for (int e = tid; e < nb * nb; e += 128) {
    int i = e % nb, jj = e / nb;
    if (i >= jj) Ld[jj * DLD + i] = Fj[(k0 + i) + (int64_t)(k0 + jj) * ld];
  }
  __syncthreads();
  if (tid < nb) invd[tid] = 1.0 / Ld[tid * DLD + tid];
  __syncthreads();
  if (warp < NRC) {
    double v0 = 0.0, v1 = 0.0;
    if (warp < nr) {
      const double* xq = x + col0 + k0 + warp * ldx;
      if (lane < nb) v0 = xq[lane];
      if (lane + 32 < nb) v1 = xq[lane + 32];
      const double* __restrict__ pj = partial + T.c;
      for (int ch = 0; ch < nchunk; ch++) {
        const double* pc = pj + (int64_t)ch * s * NRC;
        if (lane < nb) v0 -= pc[(int64_t)(k0 + lane) * NRC + warp];
        if (lane + 32 < nb) v1 -= pc[(int64_t)(k0 + lane + 32) * NRC + warp];
      }
      for (int c = nb - 1; c >= 0; c--) {
        double xc = __shfl_sync(0xffffffffu, (c < 32) ? v0 : v1, c & 31) * invd[c];
        if (c < 32) {
          if (lane == c) v0 = xc;
        } else {
          if (lane == c - 32) v1 = xc;
          if (lane + 32 < c) v1 -= Ld[(lane + 32) * DLD + c] * xc;
        }
        if (lane < c && lane < nb) v0 -= Ld[lane * DLD + c] * xc;
      }
    }
    xk[warp][lane] = v0;
    xk[warp][lane + 32] = v1;
    if (j == k && warp < nr) {
      double* xq = xsol + col0 + k0 + warp * ldx;
      if (lane < nb) xq[lane] = v0;
      if (lane + 32 < nb) xq[lane + 32] = v1;
    }
  }
  __syncthreads();
  }
  if (j == k) return;
  double acc[NRC];
#pragma unroll
  for (int q = 0; q < NRC; q++) acc[q] = 0.0;
#pragma unroll
  for (int u = 0; u < 32; u++)
#pragma unroll
    for (int q = 0; q < NRC; q++) acc[q] += f[u] * xk[q][r0 + u];
  if (hf == 1) {
#pragma unroll
    for (int q = 0; q < NRC; q++) half[cc][q] = acc[q];
  }
  __syncthreads();
  if (hf == 0) {
#pragma unroll
    for (int q = 0; q < NRC; q++)
      if (q < nr) x[col0 + j * 64 + cc + q * ldx] -= acc[q] + half[cc][q];
  }
}

// --------------------------------------------------------------------------------------- SpMV/SpMM ----
// y = alpha * G x + beta * y for a matrix stored by rows of G (ptr/idx/val): 8 lanes per row.
__global__ void k_spmv_rows(int64_t nrows, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                            const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                            double alpha, double beta) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  double acc = 0.0;
  if (row < nrows) {
    for (int64_t p = ptr[row] + sub; p < ptr[row + 1]; p += 8) acc += val[p] * x[idx[p]];
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (row < nrows && sub == 0) y[row] = alpha * acc + (beta != 0.0 ? beta * y[row] : 0.0);
}

// RBMC accumulation (Siden et al. 2018): for row i, with samples X node-major (X[k + j*ldk], k < nsamp):
//   var_i = 1/Q_ii + (1/nsamp) sum_k ( sum_{j != i} Q_ij x_j^(k) )^2 / Q_ii^2.   One warp per row.
__global__ void k_rbmc(int64_t n, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                       const double* __restrict__ val, const double* __restrict__ X, int64_t ldk, int nsamp,
                       double* __restrict__ var) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double qii = 0.0, sumsq = 0.0;
  for (int k0 = 0; k0 < nsamp; k0 += 32) {
    const int k = k0 + lane;
    double t = 0.0;
    for (int64_t p = ptr[row]; p < ptr[row + 1]; p++) {
      const int64_t j = idx[p];
      const double q = val[p];
      if (j == row) {
        qii = q;
      } else if (k < nsamp) {
        t += q * X[k + j * ldk];
      }
    }
    sumsq += t * t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
  if (lane == 0) var[row] = 1.0 / qii + sumsq / ((double)nsamp * qii * qii);
}

// dst[k + p*ldk] = src[p + k*lds]  (RHS-major -> node-major through a permutation: dst row = perm[p])
__global__ void k_perm_scatter_nodemajor(const double* __restrict__ src, int64_t lds, double* __restrict__ dst,
                                         int64_t ldk, const int32_t* __restrict__ perm, int64_t n, int k0, int nr) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int64_t o = perm[p];
  for (int q = 0; q < nr; q++) dst[(k0 + q) + o * ldk] = src[p + q * lds];
}

// values of a matrix re-laid out through an index map: dst[k] = src[map[k]]
__global__ void k_gather_values(const double* __restrict__ src, const int64_t* __restrict__ map, int64_t nnz,
                                double* __restrict__ dst) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) dst[k] = src[map[k]];
}

// Posterior precision values on a fixed pattern:
//   out[k] = Q-part + sum over precomputed products: out[k] = (qsrc[k] >= 0 ? Qval[qsrc[k]] : 0)
//            + sum_{t in [pptr[k], pptr[k+1])} w[prow[t]] * Aval[pa[t]] * Aval[pb[t]]
__global__ void k_postprec(int64_t nnz_out, const int64_t* __restrict__ qsrc, const double* __restrict__ Qval,
                           const int64_t* __restrict__ pptr, const int32_t* __restrict__ prow,
                           const int64_t* __restrict__ pa, const int64_t* __restrict__ pb,
                           const double* __restrict__ Aval, const double* __restrict__ wdiag, double wscalar,
                           double* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz_out) return;
  double v = qsrc[k] >= 0 ? Qval[qsrc[k]] : 0.0;
  for (int64_t t = pptr[k]; t < pptr[k + 1]; t++) {
    double w = wdiag ? wdiag[prow[t]] : wscalar;
    v += w * Aval[pa[t]] * Aval[pb[t]];
  }
  out[k] = v;
}

// sum_i a_i * b_i  ->  out (single double, atomically accumulated; out must be zeroed first)
__global__ void k_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* out) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += a[i] * b[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[i];
    atomicAdd(out, s);
  }
}

__global__ void k_axpby(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                        double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}

// log-determinant pieces: out[k] = F[diag of internal column k]
__global__ void k_diag_L(const SnodeDesc* __restrict__ sd, int nsuper, const double* __restrict__ F,
                         double* __restrict__ out) {
  int s = blockIdx.x;
  if (s >= nsuper) return;
  const SnodeDesc D = sd[s];
  for (int i = threadIdx.x; i < D.s; i += blockDim.x) out[D.col0 + i] = F[D.foff + (int64_t)i * D.ld + i];
}

// ------------------------------------------------------------------------------------ host wrappers ----
static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

cudaError_t launch_perm_gather(const double* src, int64_t lds, double* dst, int64_t ldd, const int32_t* perm,
                               int64_t n, int nrhs, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_perm_gather<<<blocks_for(n, 256), 256, 0, st>>>(src, lds, dst, ldd, perm, n, nrhs);
  return cudaGetLastError();
}
cudaError_t launch_perm_scatter(const double* src, int64_t lds, double* dst, int64_t ldd, const int32_t* perm,
                                int64_t n, int nrhs, const double* add, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_perm_scatter<<<blocks_for(n, 256), 256, 0, st>>>(src, lds, dst, ldd, perm, n, nrhs, add);
  return cudaGetLastError();
}
cudaError_t launch_perm_scatter_nodemajor(const double* src, int64_t lds, double* dst, int64_t ldk,
                                          const int32_t* perm, int64_t n, int k0, int nr, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_perm_scatter_nodemajor<<<blocks_for(n, 256), 256, 0, st>>>(src, lds, dst, ldk, perm, n, k0, nr);
  return cudaGetLastError();
}

cudaError_t sparse_kernels_init() { return cudaSuccess; }

cudaError_t launch_fwd_assemble(const SnodeDesc* sd, const int32_t* list, int count, const int32_t* child_idx,
                                const int32_t* relmap, double* x, int64_t ldx, double* uvec, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  k_fwd_assemble<SOLVE_NRC><<<count, 256, 0, st>>>(sd, list, child_idx, relmap, x, ldx, uvec);
  return cudaGetLastError();
}
cudaError_t launch_fwd_small(const SnodeDesc* sd, const int32_t* list, int count, const int32_t* child_idx,
                             const int32_t* relmap, const double* F, double* w, double* ysol, int64_t ldx, double* uvec,
                             int nr, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  k_fwd_small<SOLVE_NRC><<<(count + SS_WARPS - 1) / SS_WARPS, 32 * SS_WARPS, 0, st>>>(sd, list, count, child_idx, relmap, F, w,
                                                                                        ysol, ldx, uvec, nr);
  return cudaGetLastError();
}
cudaError_t launch_bwd_small(const SnodeDesc* sd, const int32_t* list, int count, const int32_t* rows, const double* F,
                             const double* t, double* xsol, int64_t ldx, int nr, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  k_bwd_small<SOLVE_NRC><<<(count + SS_WARPS - 1) / SS_WARPS, 32 * SS_WARPS, 0, st>>>(sd, list, count, rows, F, t, xsol, ldx,
                                                                                        nr);
  return cudaGetLastError();
}
cudaError_t launch_fwd_step(const Task* tasks, int ntasks, int grid, const double* F, double* w, double* ysol,
                            int64_t ldx, double* uvec, int nr, const double* dinv, cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  k_fwd_step<SOLVE_NRC><<<grid, FS_ROWS, 0, st>>>(tasks, ntasks, F, w, ysol, ldx, uvec, nr, dinv);
  return cudaGetLastError();
}
cudaError_t launch_bwd_rpart(const Task* tasks, int ntasks, int grid, const double* F, const int32_t* rows,
                             const double* x, int64_t ldx, double* partial, int nr, cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  k_bwd_rpart<SOLVE_NRC><<<grid, 128, 0, st>>>(tasks, ntasks, F, rows, x, ldx, partial, nr);
  return cudaGetLastError();
}
cudaError_t launch_bwd_step(const Task* tasks, int ntasks, int grid, const double* F, double* t, double* xsol,
                            int64_t ldx, const double* partial, int nr, const double* dinv, cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  k_bwd_step<SOLVE_NRC><<<grid, 128, 0, st>>>(tasks, ntasks, F, t, xsol, ldx, partial, nr, dinv);
  return cudaGetLastError();
}

cudaError_t launch_wide_fwd(const Task* tasks, int ntasks, int grid_trmv, int grid_gemv, const double* F, const double* Wf,
                            const double* w, double* ysol, int64_t ldx, double* uvec, int nr, cudaStream_t st) {
  if (grid_trmv > 0) k_wide_gemv<SOLVE_NRC, true><<<grid_trmv, 32 * WG_WARPS, 0, st>>>(tasks, ntasks, F, Wf, w, ysol, ldx, uvec, nr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return cudaSuccess;
}
cudaError_t launch_wide_fwd_below(const Task* tasks, int ntasks, int grid, const double* F, const double* ysol, int64_t ldx,
                                  double* uvec, int nr, cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  k_wide_gemv<SOLVE_NRC, false><<<grid, 32 * WG_WARPS, 0, st>>>(tasks, ntasks, F, nullptr, ysol, nullptr, ldx, uvec, nr);
  return cudaGetLastError();
}
cudaError_t launch_wide_bwd(const Task* tasks, int ntasks, int grid, const double* Wf, const double* t, double* xsol,
                            int64_t ldx, const double* partial, int nr, cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  k_wide_trmv_t<SOLVE_NRC><<<grid, 128, 0, st>>>(tasks, ntasks, Wf, t, xsol, ldx, partial, nr);
  return cudaGetLastError();
}
cudaError_t launch_wide_norms(const Task* tasks, int ntasks, int grid, const double* F, const double* Wf, double* norms,
                              cudaStream_t st) {
  if (grid <= 0) return cudaSuccess;
  k_wide_norms<<<grid, 256, 0, st>>>(tasks, ntasks, F, Wf, norms);
  return cudaGetLastError();
}

cudaError_t launch_spmv_rows(int64_t nrows, const int64_t* ptr, const int32_t* idx, const double* val,
                             const double* x, double* y, double alpha, double beta, cudaStream_t st) {
  if (nrows <= 0) return cudaSuccess;
  k_spmv_rows<<<blocks_for(nrows * 8, 256), 256, 0, st>>>(nrows, ptr, idx, val, x, y, alpha, beta);
  return cudaGetLastError();
}
cudaError_t launch_rbmc(int64_t n, const int64_t* ptr, const int32_t* idx, const double* val, const double* X,
                        int64_t ldk, int nsamp, double* var, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_rbmc<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, ptr, idx, val, X, ldk, nsamp, var);
  return cudaGetLastError();
}
cudaError_t launch_gather_values(const double* src, const int64_t* map, int64_t nnz, double* dst, cudaStream_t st) {
  if (nnz <= 0) return cudaSuccess;
  k_gather_values<<<blocks_for(nnz, 256), 256, 0, st>>>(src, map, nnz, dst);
  return cudaGetLastError();
}
cudaError_t launch_postprec(int64_t nnz_out, const int64_t* qsrc, const double* Qval, const int64_t* pptr,
                            const int32_t* prow, const int64_t* pa, const int64_t* pb, const double* Aval,
                            const double* wdiag, double wscalar, double* out, cudaStream_t st) {
  if (nnz_out <= 0) return cudaSuccess;
  k_postprec<<<blocks_for(nnz_out, 256), 256, 0, st>>>(nnz_out, qsrc, Qval, pptr, prow, pa, pb, Aval, wdiag, wscalar,
                                                       out);
  return cudaGetLastError();
}
cudaError_t launch_dot(const double* a, const double* b, int64_t n, double* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), st);
  if (e != cudaSuccess) return e;
  if (n <= 0) return cudaSuccess;
  unsigned g = blocks_for(n, 256);
  if (g > 1024) g = 1024;
  k_dot<<<g, 256, 0, st>>>(a, b, n, out);
  return cudaGetLastError();
}
cudaError_t launch_axpby(int64_t n, double a, const double* x, double b, const double* y, double* out,
                         cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_axpby<<<blocks_for(n, 256), 256, 0, st>>>(n, a, x, b, y, out);
  return cudaGetLastError();
}
cudaError_t launch_diag_L(const SnodeDesc* sd, int nsuper, const double* F, double* out, cudaStream_t st) {
  if (nsuper <= 0) return cudaSuccess;
  k_diag_L<<<nsuper, 64, 0, st>>>(sd, nsuper, F, out);
  return cudaGetLastError();
}

}  // namespace gmrfb
