// Probe: raw FP64 throughput of the DMMA (mma.sync m8n8k4 f64) and DFMA pipes on B200.
// Register-only loops; reports TFLOP/s for several warps-per-SM settings.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void dmma_loop(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_loop(double* out, int iters) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void atomic_loop(double* tgt, int n, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; it++) atomicAdd(&tgt[(t * 17 + it * 9973) % n], 1.0);
}

int main() {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nsm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 32 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  int wps[] = {4, 8, 16, 32};
  for (int w : wps) {
    for (int rep = 0; rep < 2; rep++) {
      dmma_loop<16><<<nsm, w * 32>>>(out, iters);
      CK(cudaDeviceSynchronize());
    }
    cudaEventRecord(e0);
    dmma_loop<16><<<nsm, w * 32>>>(out, iters);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * 8 * 4 * 16 * (double)iters * w * nsm;
    printf("DMMA m8n8k4 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", w, ms, flops / ms * 1e-9);
  }
  for (int w : wps) {
    dfma_loop<16><<<nsm, w * 32>>>(out, iters); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    dfma_loop<16><<<nsm, w * 32>>>(out, iters);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 32 * 16 * (double)iters * w * nsm;
    printf("DFMA        warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", w, ms, flops / ms * 1e-9);
  }
  {
    int n = 1 << 24; double* tgt; CK(cudaMalloc(&tgt, sizeof(double) * n)); CK(cudaMemset(tgt, 0, sizeof(double) * n));
    atomic_loop<<<nsm * 8, 256>>>(tgt, n, 100); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    atomic_loop<<<nsm * 8, 256>>>(tgt, n, 100);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("atomicAdd(double) scattered: %.2f Gatom/s\n", (double)nsm * 8 * 256 * 100 / ms * 1e-6);
  }
  // kernel launch latency (empty kernel back-to-back)
  {
    cudaEventRecord(e0);
    for (int i = 0; i < 1000; i++) dfma_loop<1><<<1, 32>>>(out, 0);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("empty-kernel launch back-to-back: %.2f us each\n", ms);
  }
  return 0;
}
