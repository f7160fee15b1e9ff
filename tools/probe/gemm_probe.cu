// Stand-alone throughput probe of the library's DMMA GEMM kernel on one large problem (tuning aid).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../diffeqgmrfs.jl_b200/csrc/kernels.cu"
using namespace gmrfb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 4096;
  int K = argc > 2 ? atoi(argv[2]) : n;
  int tri = argc > 3 ? atoi(argv[3]) : 0;
  CK(kernels_init());
  double *A, *B, *C;
  CK(cudaMalloc(&A, sizeof(double) * (size_t)n * K)); CK(cudaMalloc(&B, sizeof(double) * (size_t)n * K)); CK(cudaMalloc(&C, sizeof(double) * (size_t)n * n));
  CK(cudaMemset(A, 0, sizeof(double) * (size_t)n * K)); CK(cudaMemset(B, 0, sizeof(double) * (size_t)n * K)); CK(cudaMemset(C, 0, sizeof(double) * (size_t)n * n));
  Task t{}; t.a = 0; t.b = 0; t.c = 0; t.M = n; t.N = n; t.K = K; t.lda = n; t.ldb = n; t.ldc = n; t.tile0 = 0; t.alpha = -1; t.beta = 1;
  t.flags = (0 << TF_A_SHIFT) | (1 << TF_B_SHIFT) | (2 << TF_C_SHIFT) | (tri ? TF_TRI : 0);
  Task* dt; CK(cudaMalloc(&dt, sizeof(Task))); CK(cudaMemcpy(dt, &t, sizeof(Task), cudaMemcpyHostToDevice));
  Arenas ar{{A, B, C, nullptr}};
  LaunchAux aux;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int kinds[3] = {LK_GEMM_NT, LK_GEMM_NN, LK_GEMM_TN};
  const char* names[3] = {"NT", "NN", "TN"};
  for (int v = 0; v < 3; v++) {
    Launch L{}; L.kind = kinds[v]; L.task0 = 0; L.ntasks = 1; L.grid = gemm_tiles(n, n, tri);
    if (v > 0) { /* operand shapes: K x n storage needs ld >= K */ t.lda = (v == 2) ? K : n; t.ldb = K; CK(cudaMemcpy(dt, &t, sizeof(Task), cudaMemcpyHostToDevice)); }
    for (int i = 0; i < 2; i++) CK(run_launch(L, dt, ar, aux, 0));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 3; i++) CK(run_launch(L, dt, ar, aux, 0));
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    double flops = tri ? (double)n * n * K : 2.0 * n * n * K;
    printf("BK=%d STAGES=%d %s n=%d K=%d tri=%d grid=%d: %.3f ms  %.2f TFLOP/s\n", GEMM_BK, G_STAGES, names[v], n, K, tri, L.grid, ms, flops / ms * 1e-9);
  }
  return 0;
}
