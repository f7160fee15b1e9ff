// Internal definitions shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/gmrfb.h"
#include "kernels.hpp"
#include "plan.hpp"
#include "sparse_kernels.hpp"
#include "symbolic.hpp"

struct gmrfb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int* d_info = nullptr;       // POTRF failure column
  double* d_scalar = nullptr;  // small device scratch (reductions)
  int sm_count = 0;
  // optional per-kernel profiling (CUDA events around every launch)
  bool profiling = false;
  struct ProfRec {
    int32_t kind;
    cudaEvent_t e0, e1;
    double flops, bytes;
    int32_t grid, ntasks;
  };
  std::vector<ProfRec> prof;
};

namespace gmrfb {

std::string& global_error();

inline gmrfb_status fail(gmrfb_ctx* ctx, gmrfb_status code, const std::string& msg) {
  if (ctx)
    ctx->err = msg;
  else
    global_error() = msg;
  return code;
}

#define GMRFB_CU(ctx, call)                                                                          \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      cudaGetLastError();                                                                            \
      return ::gmrfb::fail((ctx), e_ == cudaErrorMemoryAllocation ? GMRFB_ERR_ALLOC : GMRFB_ERR_CUDA, \
                           std::string(#call) + ": " + cudaGetErrorString(e_));                      \
    }                                                                                                \
  } while (0)

// Owning device buffer.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const std::vector<T>& h, cudaStream_t st) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    e = cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);  // the host vector may be a temporary
  }
};

// A plan uploaded to the device.
struct DevPlan {
  Plan host;
  DevBuf<Task> tasks;
  bool ready = false;
};

// Profiling hooks: bracket one launch of `kind` with events when ctx->profiling is on.
struct ProfScope {
  gmrfb_ctx* ctx;
  bool on;
  gmrfb_ctx::ProfRec rec;
  ProfScope(gmrfb_ctx* c, int32_t kind, double flops, double bytes, int32_t grid = 0, int32_t ntasks = 0)
      : ctx(c), on(c && c->profiling) {
    if (!on) return;
    rec.kind = kind;
    rec.grid = grid;
    rec.ntasks = ntasks;
    rec.flops = flops;
    rec.bytes = bytes;
    cudaEventCreate(&rec.e0);
    cudaEventCreate(&rec.e1);
    cudaEventRecord(rec.e0, ctx->stream);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(rec.e1, ctx->stream);
    ctx->prof.push_back(rec);
  }
};

// Execute every launch of a plan on the context's stream.
gmrfb_status run_plan(gmrfb_ctx* ctx, const DevPlan& P, const Arenas& ar, const LaunchAux& aux);

}  // namespace gmrfb
