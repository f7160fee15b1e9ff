// Internal definitions shared by the C-ABI translation units.
#pragma once
#include <exception>
#include <new>
#include <cuda_runtime.h>

#include <cstdint>
#include <initializer_list>
#include <cstdio>
#include <cstdlib>
#include <mutex>
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
  // second stream + event for two-lane schedules (created on first use; the time-sharded block-tridiagonal factor
  // runs its GPU-filling spike GEMMs there while the latency-bound chain of the next block runs on `stream`)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_lane = nullptr;
  std::string err;
  int64_t launches = 0;
  int* d_info = nullptr;       // POTRF failure column
  int* d_info_init = nullptr;  // device constant INT_MAX (d_info is reset by a device-to-device copy: graph-capturable)
  double* d_scalar = nullptr;  // small device scratch (reductions)
  bool poison = false;         // GMRFB_POISON=1: NaN-fill the frontal arena before every factorisation (debug)
  bool use_graphs = true;      // replay the static launch lists as CUDA graphs (GMRFB_GRAPHS=0 disables)
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

// Every gmrfb_status entry point is a function-try-block closed by this macro: no C++ exception (std::bad_alloc of a
// host vector, std::system_error of a worker thread, ...) crosses the C ABI.  The message goes to the process-wide slot
// (gmrfb_last_error(NULL)); gmrfb_last_error(ctx) returns it too while the context has no message of its own.
#define GMRFB_ABI_CATCH                                                                                        \
  catch (const std::bad_alloc&) {                                                                              \
    return gmrfb::fail(nullptr, GMRFB_ERR_ALLOC, "host allocation failed (std::bad_alloc)");                   \
  }                                                                                                            \
  catch (const std::exception& e_) {                                                                           \
    return gmrfb::fail(nullptr, GMRFB_ERR_INTERNAL, std::string("unexpected C++ exception: ") + e_.what());    \
  }                                                                                                            \
  catch (...) {                                                                                                \
    return gmrfb::fail(nullptr, GMRFB_ERR_INTERNAL, "unexpected C++ exception");                               \
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

// Cache of released device allocations.  Handles of this library own several GB of fronts / factor blocks each, and a
// loop that creates and destroys them (one posterior per dataset problem, one time-sharded factor per rank) otherwise
// pays cudaMalloc + cudaFree of those GB every iteration - and concurrent processes on one node serialise in the
// driver on them (profiles/r01_multi_gpu.md).  Released buffers of >= 1 MB are kept (after the same device
// synchronisation cudaFree implies) and handed to the next allocation of (nearly) the same size on the same device.
// GMRFB_POOL=0 disables the cache, GMRFB_POOL_MAX_GB bounds it (default 24); gmrfb_pool_trim and the
// destruction of the last context give the cached memory back to the driver.
class DevPool {
 public:
  static DevPool& get() {
    static DevPool P;
    return P;
  }
  // `st`: the stream the buffer will be used on first (nullptr: unknown).  A cached buffer that was released with work
  // still queued on its stream carries an event; the new user waits for it (stream-ordered if `st` is known).
  cudaError_t alloc(void** out, size_t bytes, size_t* cap, cudaStream_t st = nullptr) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (enabled_ && bytes >= kMin) {
      std::unique_lock<std::mutex> lk(mu_);
      int best = -1;
      for (int i = 0; i < (int)free_.size(); i++)
        if (free_[i].dev == dev && free_[i].bytes >= bytes && free_[i].bytes <= bytes + bytes / 8 &&
            (best < 0 || free_[i].bytes < free_[best].bytes))
          best = i;
      if (best >= 0) {
        const Ent e = free_[best];
        cached_ -= e.bytes;
        free_.erase(free_.begin() + best);
        lk.unlock();
        if (e.ev) {
          if (st)
            cudaStreamWaitEvent(st, e.ev, 0);
          else
            cudaEventSynchronize(e.ev);
          cudaEventDestroy(e.ev);
        }
        *out = e.p;
        *cap = e.bytes;
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give the cached memory back and retry once
      cudaGetLastError();
      trim(0);
      e = cudaMalloc(out, bytes);
    }
    *cap = bytes;
    return e;
  }
  // `st`: the only stream that may still have work queued on the buffer (temporaries of one API call); the release is
  // then an event record instead of a device-wide synchronisation.  nullptr: unknown users => synchronise the device
  // (what cudaFree would have done).
  void release(void* p, size_t cap, cudaStream_t st = nullptr) {
    if (!enabled_ || cap < kMin) {
      cudaFree(p);
      return;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaEvent_t ev = nullptr;
    if (st && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
      if (cudaEventRecord(ev, st) != cudaSuccess) {
        cudaEventDestroy(ev);
        ev = nullptr;
      }
    }
    if (!ev) cudaDeviceSynchronize();
    {
      std::lock_guard<std::mutex> lk(mu_);
      free_.push_back({p, cap, dev, ev});
      cached_ += cap;
    }
    trim(max_bytes_);
  }
  size_t cached() {
    std::lock_guard<std::mutex> lk(mu_);
    return cached_;
  }
  // live contexts of the process: the last one to go trims the cache
  void ctx_created() {
    std::lock_guard<std::mutex> lk(mu_);
    live_ctx_++;
  }
  void ctx_destroyed() {
    bool last;
    {
      std::lock_guard<std::mutex> lk(mu_);
      last = --live_ctx_ <= 0;
    }
    if (last) trim(0);
  }
  // free cached buffers (oldest first) until at most `keep` bytes remain cached
  void trim(size_t keep) {
    std::lock_guard<std::mutex> lk(mu_);
    while (cached_ > keep && !free_.empty()) {
      if (free_.front().ev) cudaEventDestroy(free_.front().ev);  // cudaFree waits for the device anyway
      cudaFree(free_.front().p);
      cached_ -= free_.front().bytes;
      free_.erase(free_.begin());
    }
  }

 private:
  DevPool() {
    const char* e = std::getenv("GMRFB_POOL");
    enabled_ = !(e && e[0] == '0');
    const char* m = std::getenv("GMRFB_POOL_MAX_GB");
    max_bytes_ = (size_t)((m ? std::atof(m) : 24.0) * 1e9);
  }
  static constexpr size_t kMin = (size_t)1 << 20;
  struct Ent {
    void* p;
    size_t bytes;
    int dev;
    cudaEvent_t ev;  // recorded on the releasing stream, or nullptr (device was synchronised)
  };
  std::mutex mu_;
  std::vector<Ent> free_;
  size_t cached_ = 0, max_bytes_ = 0;
  int live_ctx_ = 0;
  bool enabled_ = true;
};

// Owning device buffer.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  size_t cap = 0;  // bytes of the underlying allocation (>= n * sizeof(T) when it came from the cache)
  cudaStream_t st = nullptr;  // set for temporaries of one API call: the only stream that uses the buffer
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) DevPool::get().release(p, cap, st);
    p = nullptr;
    n = 0;
    cap = 0;
  }
  // stream != nullptr marks a temporary that is only ever used on that stream (stream-ordered reuse, no device-wide
  // synchronisation when it is released)
  cudaError_t alloc(size_t count, cudaStream_t stream = nullptr) {
    release();
    st = stream;
    if (count == 0) return cudaSuccess;
    cudaError_t e = DevPool::get().alloc((void**)&p, count * sizeof(T), &cap, stream);
    if (e == cudaSuccess)
      n = count;
    else
      p = nullptr;
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

// The numeric phases are static launch lists (hundreds of small dependent kernels for the top of the elimination tree):
// the first execution of a phase with a given set of buffers is captured into a CUDA graph, later executions replay the
// graph with one launch call (no per-kernel host work, back-to-back scheduling on the device).  `key` identifies
// everything baked into the captured kernel arguments (buffer addresses, sizes, modes); a different key re-captures.
// The cache belongs to the symbolic handle and is shared by its factors: a new factor that the buffer pool hands the
// same device buffers (the dataset loop) replays the graphs of its predecessor.  Profiling (per-launch events) bypasses
// the graphs.
struct GraphCache {
  struct Ent {
    uint64_t key = 0;
    cudaGraphExec_t exec = nullptr;
    int64_t nodes = 0;  // kernels of the graph (for the launch counter)
    uint64_t stamp = 0;
  };
  std::vector<Ent> ents;
  std::vector<uint64_t> seen;  // keys executed once without a graph: a phase is captured on its SECOND execution, so that
                               // one-shot factors (a dataset loop that creates a factor per problem) never pay capture +
                               // instantiation for nothing
  uint64_t clock = 0;
  ~GraphCache() { clear(); }
  void clear() {
    for (auto& e : ents)
      if (e.exec) cudaGraphExecDestroy(e.exec);
    ents.clear();
  }
};
inline uint64_t graph_key(std::initializer_list<uint64_t> parts) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint64_t v : parts) {
    h ^= v;
    h *= 0x100000001b3ull;
    h ^= h >> 29;
  }
  return h ? h : 1;
}
template <class Body>
gmrfb_status run_graphed(gmrfb_ctx* ctx, GraphCache& gc, uint64_t key, Body&& body) {
  if (!ctx->use_graphs || ctx->profiling) return body();
  for (auto& e : gc.ents)
    if (e.key == key && e.exec) {
      e.stamp = ++gc.clock;
      GMRFB_CU(ctx, cudaGraphLaunch(e.exec, ctx->stream));
      ctx->launches += e.nodes;
      return GMRFB_OK;
    }
  {
    bool second = false;
    for (uint64_t k : gc.seen) second = second || k == key;
    if (!second) {
      if (gc.seen.size() >= 64) gc.seen.erase(gc.seen.begin());
      gc.seen.push_back(key);
      return body();
    }
  }
  const int64_t l0 = ctx->launches;
  GMRFB_CU(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  gmrfb_status rc = body();
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
  if (rc != GMRFB_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (ce != cudaSuccess || !graph) {
    cudaGetLastError();
    return fail(ctx, GMRFB_ERR_CUDA, std::string("CUDA graph capture failed: ") + cudaGetErrorString(ce));
  }
  GraphCache::Ent e;
  ce = cudaGraphInstantiate(&e.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, GMRFB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
  }
  e.key = key;
  e.nodes = ctx->launches - l0;
  e.stamp = ++gc.clock;
  if (gc.ents.size() >= 12) {  // bounded: drop the least recently used graph
    size_t lru = 0;
    for (size_t i = 1; i < gc.ents.size(); i++)
      if (gc.ents[i].stamp < gc.ents[lru].stamp) lru = i;
    cudaGraphExecDestroy(gc.ents[lru].exec);
    gc.ents.erase(gc.ents.begin() + lru);
  }
  gc.ents.push_back(e);
  GMRFB_CU(ctx, cudaGraphLaunch(e.exec, ctx->stream));
  return GMRFB_OK;
}

}  // namespace gmrfb
