// Bottleneck isolation for the DMMA GEMM engine: GE_PROBE_MODE 0 = full kernel, 1 = no cp.async in the main loop,
// 2 = additionally no barrier (pure LDS + DMMA stream).  Results are numerically meaningless for modes > 0.
#include <cstdio>
#include <cstdlib>
#include "../../diffeqgmrfs.jl_b200/csrc/gemm_engine.cuh"
using namespace gmrfb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
template <class CFG>
void run(const char* name, int n, int K, double* A, double* B, double* C, int reps) {
  CK(cudaFuncSetAttribute(k_gemm2<false, false, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM));
  Task t{}; t.M = n; t.N = n; t.K = K; t.alpha = -1; t.beta = 1; t.lda = n; t.ldb = n; t.ldc = n;
  t.flags = (0 << TF_A_SHIFT) | (1 << TF_B_SHIFT) | (2 << TF_C_SHIFT);
  Task* dt; CK(cudaMalloc(&dt, sizeof(Task))); CK(cudaMemcpy(dt, &t, sizeof(Task), cudaMemcpyHostToDevice));
  Arenas ar{{A, B, C, nullptr}};
  const int grid = gemm_tiles_cfg<CFG>(n, n, false);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; i++) k_gemm2<false, false, CFG><<<grid, CFG::NT, CFG::SMEM>>>(dt, 1, ar);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) k_gemm2<false, false, CFG><<<grid, CFG::NT, CFG::SMEM>>>(dt, 1, ar);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  printf("mode %d %-22s n=%d K=%d: %8.3f ms %6.2f TFLOP/s\n", GE_PROBE_MODE, name, n, K, ms, 2.0 * n * n * K / ms * 1e-9);
}
int main(int argc, char** argv) {
  int reps = argc > 1 ? atoi(argv[1]) : 3;
  const size_t nmax = (size_t)4800 * 4800;
  double *A, *B, *C;
  CK(cudaMalloc(&A, nmax * 8)); CK(cudaMalloc(&B, nmax * 8)); CK(cudaMalloc(&C, nmax * 8));
  CK(cudaMemset(A, 0, nmax * 8)); CK(cudaMemset(B, 0, nmax * 8)); CK(cudaMemset(C, 0, nmax * 8));
  run<GemmCfg<128, 64, 4, 2, 16, 3, 2>>("128x64 w4x2 k16 s3 b2", 4736, 4096, A, B, C, reps);
  run<GemmCfg<128, 128, 4, 2, 16, 4, 1>>("128x128 w4x2 k16 s4 b1", 4736, 4096, A, B, C, reps);
  return 0;
}
