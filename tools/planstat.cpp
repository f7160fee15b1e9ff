// Host-only comparison of the static launch plans (factor and selected inversion) that two orderings of the same
// matrix produce - launches, dependent POTRF steps, GEMM tiles and GEMM flops by K - without a GPU.
// Input: binary file [n, nnz (int64)] [colptr (n+1 int64)] [rowval (nnz int64)] [coords (2n double)], 0-based CSC.
// Build: g++ -O2 -std=c++17 -Idiffeqgmrfs.jl_b200/csrc -I/usr/local/cuda/include tools/planstat.cpp \
//          diffeqgmrfs.jl_b200/csrc/symbolic.cpp diffeqgmrfs.jl_b200/csrc/plan.cpp -o /tmp/planstat
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include "plan.hpp"
#include "symbolic.hpp"
using namespace gmrfb;
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  int64_t n, nnz; fread(&n, 8, 1, f); fread(&nnz, 8, 1, f);
  std::vector<int64_t> cp(n + 1), ri(nnz); fread(cp.data(), 8, n + 1, f); fread(ri.data(), 8, nnz, f);
  std::vector<double> co(2 * n); fread(co.data(), 8, 2 * n, f); fclose(f);
  for (int mode = 0; mode < 2; mode++) {
    AnalyzeOptions o; o.ordering_kind = 2; o.base = 0;
    if (mode == 0) { o.coord_dim = 2; o.coords = co.data(); }
    Symbolic S; std::string e = analyze_pattern(n, cp.data(), ri.data(), nullptr, o, S);
    if (!e.empty()) { printf("error %s\n", e.c_str()); return 1; }
    printf("== %s ND: nnzL %.1fM flops %.4g nsuper %d levels %zu max_front %d arena %.2f GB\n", mode == 0 ? "geometric" : "graph", S.nnzL / 1e6, S.flops, S.nsuper, S.levels.size(), S.max_front, S.arena * 8e-9);
    for (int ph = 0; ph < 2; ph++) {
      Plan P; if (ph == 0) build_factor_plan(S, P); else build_selinv_plan(S, P);
      std::map<int, std::pair<int, double>> by;  // kind -> (launches, flops)
      double kb[6] = {0, 0, 0, 0, 0, 0}, tiles = 0, model_ns = 0; int chain = 0, small_grid = 0;
      for (auto& L : P.launches) {
        by[L.kind].first++; by[L.kind].second += L.flops;
        if (L.grid < 148) small_grid++;
        if (L.kind == LK_POTRF) chain++;
        if (is_gemm_kind(L.kind)) {
          tiles += L.grid;
          for (int t = L.task0; t < L.task0 + L.ntasks; t++) {
            const Task& T = P.tasks[t];
            const bool tri = T.flags & TF_TRI;
            double fl = tri ? (double)T.K * ((double)std::min(T.M, T.N) * (std::min(T.M, T.N) + 1) + 2.0 * (T.M - std::min(T.M, T.N)) * std::min(T.M, T.N)) : 2.0 * T.M * T.N * T.K;
            int b = T.K <= 32 ? 0 : T.K <= 64 ? 1 : T.K <= 128 ? 2 : T.K <= 256 ? 3 : T.K <= 512 ? 4 : 5;
            kb[b] += fl;
            // per-tile cost of the 2-stage 64x64 engine fitted to tools/probe/gemm_batched_probe.cu on B200
            // (K = 29 ... 4096: 19, 25, 34, 74, 124, 1003 ns per tile, GPU-wide): 12 ns + 0.243 ns * K
            model_ns += gemm_tiles(T.M, T.N, tri, GCFG_SMALL) * (12.0 + 0.243 * T.K);
          }
        }
      }
      printf("  %s plan: %zu launches (%d with grid < 148, %d POTRF steps), flops %.4g, GEMM tiles %.0f\n", ph == 0 ? "factor" : "selinv", P.launches.size(), small_grid, chain, P.flops, tiles);
      printf("    GEMM time predicted by the per-tile model (saturated GPU, no launch latency): %.1f ms\n", model_ns * 1e-6);
      printf("    GEMM GFLOP by K: <=32 %.1f | <=64 %.1f | <=128 %.1f | <=256 %.1f | <=512 %.1f | >512 %.1f\n", kb[0] / 1e9, kb[1] / 1e9, kb[2] / 1e9, kb[3] / 1e9, kb[4] / 1e9, kb[5] / 1e9);
      for (auto& kv : by) printf("    kind %2d: %4d launches %.4g flops\n", kv.first, kv.second.first, kv.second.second);
    }
  }
  return 0;
}
