// Probe of the grouped DMMA GEMM engine on BATCHED launches shaped like the mid levels of the 1M-node bench problem
// (hundreds to thousands of ragged fronts per launch, short K), with the engine's per-task switches toggled:
// TF_KTAIL (skip zero-filled k4 steps past K) and TF_CPRE (accumulators start from C).  Every variant is checked
// against a scalar reference kernel on a small batch first.  Build: nvcc -arch=sm_100a -O3 -std=c++17
//   [-I <dir with another gemm_engine.cuh> -DENGINE_TAG=\"base\"] gemm_batched_probe.cu -o gemm_batched_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifdef ENGINE_DIR_BASE
#include "gemm_engine.cuh"  // resolved through -I (an older engine for A/B runs)
#else
#include "../../diffeqgmrfs.jl_b200/csrc/gemm_engine.cuh"
#endif
#ifndef ENGINE_TAG
#define ENGINE_TAG "tree"
#endif
using namespace gmrfb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_fill(double* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) {
    unsigned x = (unsigned)i * 2654435761u + seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = ((double)(x & 0xffff) / 65536.0) - 0.5;
  }
}
// one thread per result element of one task (blockIdx.z = task)
__global__ void k_ref(const Task* tasks, int TA, int TB, const double* A, const double* B, double* C) {
  const Task T = tasks[blockIdx.z];
  int i = blockIdx.x * 16 + threadIdx.x, j = blockIdx.y * 16 + threadIdx.y;
  const bool tri = (T.flags & TF_TRI) != 0;
  if (i >= T.M || j >= T.N || (tri && j > i)) return;
  double s = 0;
  for (int k = 0; k < T.K; k++) {
    double a = TA ? A[T.a + k + (size_t)i * T.lda] : A[T.a + i + (size_t)k * T.lda];
    double b = TB ? B[T.b + k + (size_t)j * T.ldb] : B[T.b + j + (size_t)k * T.ldb];
    s += a * b;
  }
  double* c = C + T.c + i + (size_t)j * T.ldc;
  *c = T.beta * (*c) + T.alpha * s;
}

template <class CFG>
const char* cfg_name() {
  static char buf[64];
  snprintf(buf, sizeof buf, "%dx%dx%d/w%dx%d/s%d/b%d", CFG::BM, CFG::BN, CFG::BK, CFG::WARPS_M, CFG::WARPS_N, CFG::STAGES, CFG::MINB);
  return buf;
}

struct Batch {
  std::vector<Task> tasks;   // tasks followed by the CTA map slots
  int ntasks = 0, grid = 0;
  double flops = 0;
  size_t a_elems = 0, c_elems = 0;
};

// A batch of `nt` problems C (M x N inside a front of leading dimension ldc) -= A B' (NT) or its variants; operands of
// successive problems are laid out one after another like fronts in the frontal arena.  `odd` shifts every offset by
// one element (misaligned 128-bit pairs => scalar epilogue).
template <class CFG>
Batch make_batch(int nt, int M, int N, int K, bool tri, bool TA, bool TB, int32_t extra_flags, bool odd, int jitter) {
  Batch b;
  b.ntasks = nt;
  size_t aoff = odd ? 1 : 0, coff = odd ? 1 : 0;
  unsigned rng = 12345u;
  for (int t = 0; t < nt; t++) {
    rng = rng * 1664525u + 1013904223u;
    const int dm = jitter ? (int)((rng >> 8) % (2 * jitter + 1)) - jitter : 0;
    rng = rng * 1664525u + 1013904223u;
    const int dk = jitter ? (int)((rng >> 8) % (jitter + 1)) - jitter / 2 : 0;
    const int m = std::max(8, M + dm), n = tri ? m : std::max(8, N + dm / 2), k = std::max(4, K + dk);
    Task T{};
    T.M = m; T.N = n; T.K = k;
    T.alpha = -1.0; T.beta = 1.0;
    T.lda = (TA ? k : m) + 2; T.ldb = (TB ? k : n) + 2; T.ldc = m + 2;
    T.lda += T.lda & 1; T.ldb += T.ldb & 1; T.ldc += T.ldc & 1;   // even leading dimensions, as in the arenas
    T.a = (int64_t)aoff; aoff += (size_t)T.lda * (TA ? m : k);
    if (tri && !TA && !TB) { T.b = T.a; T.ldb = T.lda; }  // SYRK: both operands are the same panel L21
    else { T.b = (int64_t)aoff; aoff += (size_t)T.ldb * (TB ? n : k); }
    T.c = (int64_t)coff; coff += (size_t)T.ldc * n;
    aoff += aoff & 1; coff += coff & 1;
    if (odd) { aoff |= 1; coff |= 1; }
    T.flags = (0 << TF_A_SHIFT) | (0 << TF_B_SHIFT) | (1 << TF_C_SHIFT) | (tri ? TF_TRI : 0) | extra_flags;
#ifdef GMRFB_GEMM_SWITCHES
    if (!((k % 16) != 0 && (k % 16) <= 12)) T.flags &= ~TF_KTAIL;
#endif
    T.tile0 = b.grid;
    b.grid += gemm_tiles_cfg<CFG>(m, n, tri);
    b.flops += tri ? (double)k * ((double)n * (n + 1) + 2.0 * (m - n) * n) : 2.0 * m * n * k;
    b.tasks.push_back(T);
  }
  b.a_elems = aoff; b.c_elems = coff;
  const size_t nslots = ((size_t)b.grid * sizeof(int32_t) + sizeof(Task) - 1) / sizeof(Task);
  b.tasks.resize(nt + nslots, Task{});
  int32_t* map = reinterpret_cast<int32_t*>(b.tasks.data() + nt);
  for (int t = 0; t < nt; t++) {
    const int lo = b.tasks[t].tile0, hi = t + 1 < nt ? b.tasks[t + 1].tile0 : b.grid;
    for (int c = lo; c < hi; c++) map[c] = t;
  }
  return b;
}

struct Bufs { double *A, *C, *Cr; size_t na, nc; Task* dt; size_t ntask_cap; };

template <bool TA, bool TB, class CFG>
void run(const char* wname, const char* vname, Bufs& buf, const Batch& b, bool check) {
  static bool init = false;
  if (!init) { CK(cudaFuncSetAttribute(k_gemm2<TA, TB, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM)); init = true; }
  if (b.a_elems > buf.na || b.c_elems > buf.nc || b.tasks.size() > buf.ntask_cap) { printf("%s: batch too large\n", wname); return; }
  CK(cudaMemcpy(buf.dt, b.tasks.data(), b.tasks.size() * sizeof(Task), cudaMemcpyHostToDevice));
  Arenas ar{{buf.A, buf.C, nullptr, nullptr}};
  if (check) {
    k_fill<<<(unsigned)((b.c_elems + 255) / 256), 256>>>(buf.C, b.c_elems, 7u);
    CK(cudaMemcpy(buf.Cr, buf.C, b.c_elems * 8, cudaMemcpyDeviceToDevice));
    k_gemm2<TA, TB, CFG><<<b.grid, CFG::NT, CFG::SMEM>>>(buf.dt, b.ntasks, ar);
    CK(cudaGetLastError());
    int mm = 0, mn = 0;
    for (int t = 0; t < b.ntasks; t++) { mm = std::max(mm, b.tasks[t].M); mn = std::max(mn, b.tasks[t].N); }
    dim3 g((mm + 15) / 16, (mn + 15) / 16, b.ntasks), blk(16, 16);
    k_ref<<<g, blk>>>(buf.dt, TA, TB, buf.A, buf.A, buf.Cr);
    CK(cudaDeviceSynchronize());
    std::vector<double> h(b.c_elems), hr(b.c_elems);
    CK(cudaMemcpy(h.data(), buf.C, b.c_elems * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), buf.Cr, b.c_elems * 8, cudaMemcpyDeviceToHost));
    double err = 0;
    for (size_t i = 0; i < b.c_elems; i++) err = std::max(err, std::fabs(h[i] - hr[i]));
    printf("  check [%s] %-10s %-12s %c%c %s: max abs err %.3e %s\n", ENGINE_TAG, wname, vname, TA ? 'T' : 'N', TB ? 'N' : 'T',
           cfg_name<CFG>(), err, err < 1e-11 ? "ok" : "FAIL");
    return;
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; i++) k_gemm2<TA, TB, CFG><<<b.grid, CFG::NT, CFG::SMEM>>>(buf.dt, b.ntasks, ar);
  CK(cudaDeviceSynchronize());
  const int reps = 10;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) k_gemm2<TA, TB, CFG><<<b.grid, CFG::NT, CFG::SMEM>>>(buf.dt, b.ntasks, ar);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  printf("[%s] %-10s %-12s %c%c %s grid=%6d: %8.3f ms %6.2f TFLOP/s\n", ENGINE_TAG, wname, vname, TA ? 'T' : 'N', TB ? 'N' : 'T',
         cfg_name<CFG>(), b.grid, ms, b.flops / ms * 1e-9);
}

#ifdef GMRFB_GEMM_SWITCHES
static const int32_t VARIANT_FLAGS[] = {0, TF_KTAIL, TF_CPRE, TF_KTAIL | TF_CPRE};
static const char* VARIANT_NAMES[] = {"plain", "ktail", "cpre", "ktail+cpre"};
static const int NVAR = 4;
#else
static const int32_t VARIANT_FLAGS[] = {0};
static const char* VARIANT_NAMES[] = {"plain"};
static const int NVAR = 1;
#endif

template <bool TA, bool TB, class CFG>
void workload(const char* wname, Bufs& buf, int nt, int M, int N, int K, bool tri, int jitter) {
  for (int v = 0; v < NVAR; v++) {
    Batch b = make_batch<CFG>(nt, M, N, K, tri, TA, TB, VARIANT_FLAGS[v], false, jitter);
    run<TA, TB, CFG>(wname, VARIANT_NAMES[v], buf, b, false);
  }
}
template <bool TA, bool TB, class CFG>
void checks(Bufs& buf) {
  for (int v = 0; v < NVAR; v++)
    for (int odd = 0; odd < 2; odd++) {
      Batch b1 = make_batch<CFG>(7, 203, 203, 29, true, TA, TB, VARIANT_FLAGS[v], odd, 40);
      run<TA, TB, CFG>(odd ? "tri/odd" : "tri", VARIANT_NAMES[v], buf, b1, true);
      Batch b2 = make_batch<CFG>(5, 333, 150, 77, false, TA, TB, VARIANT_FLAGS[v], odd, 60);
      run<TA, TB, CFG>(odd ? "rect/odd" : "rect", VARIANT_NAMES[v], buf, b2, true);
    }
}

int main() {
  Bufs buf;
  buf.na = (size_t)400 << 20; buf.nc = (size_t)400 << 20;  // 3.2 GB each
  CK(cudaMalloc(&buf.A, buf.na * 8)); CK(cudaMalloc(&buf.C, buf.nc * 8)); CK(cudaMalloc(&buf.Cr, ((size_t)8 << 20) * 8));
  buf.ntask_cap = 1 << 16;
  CK(cudaMalloc(&buf.dt, buf.ntask_cap * sizeof(Task)));
  k_fill<<<(unsigned)((buf.na + 255) / 256), 256>>>(buf.A, buf.na, 1u);
  CK(cudaMemset(buf.C, 0, buf.nc * 8));
  CK(cudaDeviceSynchronize());
  using Small = GemmCfg<64, 64, 2, 2, 16, 2, 4>;  // the library's configuration
  using Big = GemmCfg<128, 64, 4, 2, 16, 3, 2>;
  {
    Bufs small = buf; small.nc = (size_t)8 << 20;
    checks<false, false, Small>(small);
    checks<false, true, Small>(small);
    checks<true, false, Small>(small);
    checks<true, true, Small>(small);
    checks<false, false, Big>(small);
    checks<true, true, Big>(small);
  }
  // factor: trailing SYRK of a level (NT, lower-trapezoidal), shapes of the bench levels (fronts, r, K = s)
#define SHORTK(CFG)                                                              \
  workload<false, false, CFG>("syrk1714", buf, 1714, 200, 200, 29, true, 30);    \
  workload<false, false, CFG>("syrk1020", buf, 1020, 250, 250, 56, true, 40);    \
  workload<false, false, CFG>("syrk288", buf, 288, 450, 450, 90, true, 60);      \
  workload<false, false, CFG>("syrk64", buf, 64, 1000, 1000, 256, true, 100);    \
  workload<false, false, CFG>("zrr1020", buf, 1020, 250, 56, 250, false, 30);    \
  workload<false, false, CFG>("zrr108", buf, 108, 600, 214, 600, false, 60);     \
  workload<false, true, CFG>("zrr108", buf, 108, 600, 214, 600, false, 60);
  SHORTK(Small)
#ifdef PROBE_EXTRA_CFGS
  using S9 = GemmCfg<64, 64, 4, 2, 16, 2, 4>;   // 64-register cap: 4 CTAs/SM
  using S10 = GemmCfg<64, 64, 2, 2, 16, 2, 4>;  // 4 warps, 32x32 warp tiles, 4 CTAs/SM
  using S11 = GemmCfg<64, 64, 2, 2, 16, 2, 5>;
  using S12 = GemmCfg<64, 64, 4, 2, 8, 3, 3>;   // BK = 8
  SHORTK(S9)
  SHORTK(S10)
  SHORTK(S11)
  SHORTK(S12)
  workload<false, false, S10>("syrk16", buf, 16, 1800, 1800, 450, true, 100);
  workload<false, false, S10>("big1", buf, 1, 4736, 4736, 4096, false, 0);
  workload<false, false, Small>("big1", buf, 1, 4736, 4736, 4096, false, 0);
#endif
  workload<false, false, Big>("syrk64", buf, 64, 1000, 1000, 256, true, 100);
  workload<false, false, Small>("syrk16", buf, 16, 1800, 1800, 450, true, 100);
  workload<false, false, Big>("syrk16", buf, 16, 1800, 1800, 450, true, 100);
  workload<false, false, Big>("zrr108", buf, 108, 600, 214, 600, false, 60);
  workload<true, true, Small>("zrr108", buf, 108, 600, 214, 600, false, 60);
  return 0;
}
