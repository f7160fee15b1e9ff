// Stand-alone correctness + throughput probe of the grouped DMMA GEMM engine (gemm_engine.cuh) over tile configurations.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../diffeqgmrfs.jl_b200/csrc/gemm_engine.cuh"
using namespace gmrfb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_ref(int TA, int TB, int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C,
                      int ldc, double alpha, double beta, int tri) {
  int i = blockIdx.x * 16 + threadIdx.x, j = blockIdx.y * 16 + threadIdx.y;
  if (i >= M || j >= N || (tri && j > i)) return;
  double s = 0;
  for (int k = 0; k < K; k++) {
    double a = TA ? A[k + (size_t)i * lda] : A[i + (size_t)k * lda];
    double b = TB ? B[k + (size_t)j * ldb] : B[j + (size_t)k * ldb];
    s += a * b;
  }
  C[i + (size_t)j * ldc] = beta * C[i + (size_t)j * ldc] + alpha * s;
}
__global__ void k_fill(double* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) {
    unsigned x = (unsigned)i * 2654435761u + seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = ((double)(x & 0xffff) / 65536.0) - 0.5;
  }
}

template <bool TA, bool TB, class CFG>
void run(const char* name, int M, int N, int K, int tri, bool check, double* A, double* B, double* C, double* Cr) {
  static bool init = false;
  if (!init) { CK(cudaFuncSetAttribute(k_gemm2<TA, TB, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM)); init = true; }
  Task t{}; t.a = 0; t.b = 0; t.c = 0; t.M = M; t.N = N; t.K = K; t.tile0 = 0; t.alpha = -1; t.beta = 1;
  t.lda = TA ? K + 2 : M + 2; t.ldb = TB ? K + 2 : N + 2; t.ldc = M + 2;
  t.flags = (0 << TF_A_SHIFT) | (1 << TF_B_SHIFT) | (2 << TF_C_SHIFT) | (tri ? TF_TRI : 0);
  static Task* dt = nullptr; if (!dt) CK(cudaMalloc(&dt, sizeof(Task)));
  CK(cudaMemcpy(dt, &t, sizeof(Task), cudaMemcpyHostToDevice));
  Arenas ar{{A, B, C, nullptr}};
  const int grid = gemm_tiles_cfg<CFG>(M, N, tri);
  if (check) {
    size_t nc = (size_t)t.ldc * N;
    k_fill<<<(unsigned)((nc + 255) / 256), 256>>>(C, nc, 7u);
    CK(cudaMemcpy(Cr, C, nc * 8, cudaMemcpyDeviceToDevice));
    k_gemm2<TA, TB, CFG><<<grid, CFG::NT, CFG::SMEM>>>(dt, 1, ar);
    CK(cudaGetLastError());
    dim3 g((M + 15) / 16, (N + 15) / 16), b(16, 16);
    k_ref<<<g, b>>>(TA, TB, M, N, K, A, t.lda, B, t.ldb, Cr, t.ldc, t.alpha, t.beta, tri);
    CK(cudaDeviceSynchronize());
    std::vector<double> h(nc), hr(nc);
    CK(cudaMemcpy(h.data(), C, nc * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hr.data(), Cr, nc * 8, cudaMemcpyDeviceToHost));
    double err = 0; for (size_t i = 0; i < nc; i++) err = fmax(err, fabs(h[i] - hr[i]));
    printf("  check %s %c%c M=%d N=%d K=%d tri=%d: max abs err %.3e %s\n", name, TA ? 'T' : 'N', TB ? 'N' : 'T', M, N, K, tri, err, err < 1e-9 * K ? "ok" : "FAIL");
    return;
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; i++) k_gemm2<TA, TB, CFG><<<grid, CFG::NT, CFG::SMEM>>>(dt, 1, ar);
  CK(cudaDeviceSynchronize());
  const int reps = 3;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) k_gemm2<TA, TB, CFG><<<grid, CFG::NT, CFG::SMEM>>>(dt, 1, ar);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  double flops = tri ? (double)M * N * K : 2.0 * M * N * K;
  printf("%-22s %c%c M=%5d N=%5d K=%5d tri=%d grid=%5d: %8.3f ms %6.2f TFLOP/s\n", name, TA ? 'T' : 'N', TB ? 'N' : 'T', M, N, K, tri, grid, ms, flops / ms * 1e-9);
}

template <class CFG>
void suite(const char* name, double* A, double* B, double* C, double* Cr) {
  // correctness on ragged shapes
  run<false, false, CFG>(name, 333, 217, 77, 0, true, A, B, C, Cr);
  run<false, false, CFG>(name, 401, 401, 50, 1, true, A, B, C, Cr);
  run<false, true, CFG>(name, 333, 217, 77, 0, true, A, B, C, Cr);
  run<true, true, CFG>(name, 333, 217, 77, 0, true, A, B, C, Cr);
  run<true, false, CFG>(name, 333, 217, 77, 0, true, A, B, C, Cr);
  // throughput
  run<false, false, CFG>(name, 4736, 4736, 4096, 0, false, A, B, C, Cr);
  run<false, true, CFG>(name, 4736, 4736, 4096, 0, false, A, B, C, Cr);
  run<true, true, CFG>(name, 4736, 4736, 4096, 0, false, A, B, C, Cr);
  run<true, false, CFG>(name, 4736, 4736, 4096, 0, false, A, B, C, Cr);
  run<false, false, CFG>(name, 4736, 4736, 256, 0, false, A, B, C, Cr);
  run<false, false, CFG>(name, 3000, 3000, 256, 1, false, A, B, C, Cr);
  run<false, false, CFG>(name, 3000, 3000, 64, 1, false, A, B, C, Cr);
  run<false, false, CFG>(name, 1000, 1000, 128, 1, false, A, B, C, Cr);
}

int main() {
  const size_t nmax = (size_t)4800 * 4800;
  double *A, *B, *C, *Cr;
  CK(cudaMalloc(&A, nmax * 8)); CK(cudaMalloc(&B, nmax * 8)); CK(cudaMalloc(&C, nmax * 8)); CK(cudaMalloc(&Cr, nmax * 8));
  k_fill<<<(unsigned)((nmax + 255) / 256), 256>>>(A, nmax, 1u);
  k_fill<<<(unsigned)((nmax + 255) / 256), 256>>>(B, nmax, 2u);
  CK(cudaDeviceSynchronize());
  suite<GemmCfg<128, 64, 4, 2, 16, 3, 2>>("128x64 w4x2 k16 s3 b2", A, B, C, Cr);
  suite<GemmCfg<128, 128, 4, 2, 16, 4, 1>>("128x128 w4x2 k16 s4 b1", A, B, C, Cr);
  suite<GemmCfg<64, 64, 4, 2, 16, 4, 3>>("64x64 w4x2 k16 s4 b3", A, B, C, Cr);
  return 0;
}
