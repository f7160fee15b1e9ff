// Symbolic analysis (host, integer): ordering, elimination tree, column counts, supernodes, frontal layout.
// See symbolic.hpp.  Algorithms are the published ones, restated here from their descriptions:
//  - elimination tree: Liu (1990), ancestor path compression;
//  - column counts: Gilbert, Ng & Peyton (1994) skeleton-leaf / least-common-ancestor scheme;
//  - nested dissection: George & Liu automatic ND (BFS level-structure bisection from a pseudo-peripheral
//    vertex) with a boundary-layer vertex separator; coordinate bisection when node coordinates are supplied;
//    separators refined to a minimum vertex cover of the cut edges (Liu 1989 / Pothen & Fan 1990: Hopcroft-Karp
//    maximum matching + Koenig's theorem); optional halo-AMD ordering of the leaf subdomains;
//  - approximate minimum degree: Amestoy, Davis & Duff (1996) on the quotient graph, with an optional halo;
//  - relaxed supernode amalgamation in the spirit of CHOLMOD's nrelax/zrelax rule, tuned for wide GPU fronts.
#include "symbolic.hpp"

#include <algorithm>
#include <atomic>
#include <system_error>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <numeric>

namespace gmrfb {

// ---------------------------------------------------------------------------------------------------------
// Elimination tree of a symmetric pattern given as sorted adjacency lists (no diagonal), both triangles.
void etree_lower(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                 std::vector<int32_t>& parent) {
  parent.assign(n, -1);
  std::vector<int32_t> anc(n, -1);
  for (int32_t i = 0; i < n; i++) {
    for (int64_t p = xadj[i]; p < xadj[i + 1]; p++) {
      int32_t k = adj[p];
      if (k >= i) break;  // sorted: only neighbours below i
      // climb from k to the current root of its subtree, redirecting every visited node to i
      while (k != -1 && k < i) {
        int32_t next = anc[k];
        anc[k] = i;
        if (next == -1) parent[k] = i;
        k = next;
      }
    }
  }
}

// Postorder of a forest: children visited in ascending order, roots in ascending order.
void postorder_tree(int32_t n, const std::vector<int32_t>& parent, std::vector<int32_t>& post) {
  std::vector<int32_t> head(n, -1), next(n, -1);
  for (int32_t j = n - 1; j >= 0; j--) {
    int32_t p = parent[j];
    if (p >= 0) {
      next[j] = head[p];
      head[p] = j;
    }
  }
  post.clear();
  post.reserve(n);
  std::vector<int32_t> stack;
  for (int32_t r = 0; r < n; r++) {
    if (parent[r] != -1) continue;
    stack.push_back(r);
    while (!stack.empty()) {
      int32_t v = stack.back();
      int32_t c = head[v];
      if (c == -1) {
        post.push_back(v);
        stack.pop_back();
      } else {
        head[v] = next[c];
        stack.push_back(c);
      }
    }
  }
}

// Column counts of L (incl. diagonal) without forming L.  post[k] = k-th vertex of a postorder.
void column_counts(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                   const std::vector<int32_t>& parent, const std::vector<int32_t>& post,
                   std::vector<int32_t>& colcount) {
  std::vector<int32_t> first(n, -1), maxfirst(n, -1), prevleaf(n, -1), setroot(n);
  std::vector<int64_t> delta(n, 0);
  // first[j]: postorder rank of the first descendant of j; a vertex is an etree leaf iff it is its own first.
  for (int32_t k = 0; k < n; k++) {
    int32_t j = post[k];
    delta[j] = (first[j] == -1) ? 1 : 0;
    for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
  }
  std::iota(setroot.begin(), setroot.end(), 0);
  for (int32_t k = 0; k < n; k++) {
    int32_t j = post[k];
    if (parent[j] != -1) delta[parent[j]]--;
    // rows i > j with A(i,j) != 0: is j a leaf of the row subtree of i?
    for (int64_t p = xadj[j]; p < xadj[j + 1]; p++) {
      int32_t i = adj[p];
      if (i <= j) continue;
      if (first[j] <= maxfirst[i]) continue;  // j lies under a previously seen leaf's subtree: not a leaf
      maxfirst[i] = first[j];
      int32_t jprev = prevleaf[i];
      prevleaf[i] = j;
      delta[j]++;
      if (jprev != -1) {
        // least common ancestor of jprev and j = representative of jprev's merged set
        int32_t q = jprev;
        while (q != setroot[q]) q = setroot[q];
        for (int32_t s = jprev; s != q;) {
          int32_t sn = setroot[s];
          setroot[s] = q;
          s = sn;
        }
        delta[q]--;
      }
    }
    if (parent[j] != -1) setroot[j] = parent[j];
  }
  // accumulate along the tree: children are always numbered below their parent
  for (int32_t j = 0; j < n; j++)
    if (parent[j] != -1) delta[parent[j]] += delta[j];
  colcount.resize(n);
  for (int32_t j = 0; j < n; j++) colcount[j] = (int32_t)delta[j];
}

// ---------------------------------------------------------------------------------------------------------
// Nested dissection.
namespace {

// int32 with relaxed atomic loads and stores: `label` is the one per-vertex array a worker reads for vertices OUTSIDE its
// own range (a neighbour across a separator may be relabelled by another worker at that moment; either value it can
// see differs from the reader's own, unique label, so the comparison has one outcome)
struct RelaxedI32 {
  std::atomic<int32_t> v;
  RelaxedI32(int32_t x = 0) : v(x) {}
  RelaxedI32(const RelaxedI32& o) : v(o.v.load(std::memory_order_relaxed)) {}
  operator int32_t() const { return v.load(std::memory_order_relaxed); }
  RelaxedI32& operator=(int32_t x) {
    v.store(x, std::memory_order_relaxed);
    return *this;
  }
};

// per-vertex state of a dissection, shared by all workers: every open range owns a disjoint set of vertices, and a
// worker touches `verts` positions, `lvl`, `stamp` and `posv` entries of its own range only
struct NDShared {
  std::vector<int32_t> verts;   // current elimination order, rearranged in place
  std::vector<RelaxedI32> label;  // subset id of each vertex
  std::vector<int32_t> lvl;     // BFS level scratch (valid where stamp matches)
  std::vector<int32_t> stamp;   // visit stamp
  std::vector<int32_t> posv;    // vertex -> position in its current range (separator refinement)
  std::atomic<int32_t> next_label{1}, next_stamp{1};  // ids are unique across workers
  explicit NDShared(int32_t n) : verts(n), label(n, RelaxedI32(0)), lvl(n, 0), stamp(n, 0), posv(n, 0) {
    std::iota(verts.begin(), verts.end(), 0);
  }
};

struct NDWork {
  int32_t n;
  const std::vector<int64_t>& xadj;
  const std::vector<int32_t>& adj;
  std::vector<int32_t>& verts;
  std::vector<RelaxedI32>& label;
  std::vector<int32_t>& lvl;
  std::vector<int32_t>& stamp;
  std::vector<int32_t> queue;   // this worker's BFS queue
  std::atomic<int32_t>& next_label;
  std::atomic<int32_t>& next_stamp;
  int leaf;
  int coord_dim;
  const double* coords;
  NDWork(int32_t n_, const std::vector<int64_t>& xa, const std::vector<int32_t>& a, NDShared& sh)
      : n(n_), xadj(xa), adj(a), verts(sh.verts), label(sh.label), lvl(sh.lvl), stamp(sh.stamp),
        next_label(sh.next_label), next_stamp(sh.next_stamp) {}

  // BFS inside subset `lab` from `root`; fills queue (visit order) and lvl; returns number of levels.
  int bfs(int32_t root, int32_t lab, int32_t st, bool append) {
    if (!append) queue.clear();
    size_t head = queue.size();
    queue.push_back(root);
    stamp[root] = st;
    lvl[root] = 0;
    int maxl = 0;
    while (head < queue.size()) {
      int32_t v = queue[head++];
      int32_t lv = lvl[v];
      for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) {
        int32_t u = adj[p];
        if (label[u] != lab || stamp[u] == st) continue;
        stamp[u] = st;
        lvl[u] = lv + 1;
        if (lv + 1 > maxl) maxl = lv + 1;
        queue.push_back(u);
      }
    }
    return maxl + 1;
  }

  int32_t degree_in(int32_t v, int32_t lab) const {
    int32_t d = 0;
    for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) d += (label[adj[p]] == lab);
    return d;
  }
};

}  // namespace

void nested_dissection(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                       int leaf, int coord_dim, const double* coords, std::vector<int32_t>& perm, bool amd_leaves) {
  NDShared SH(n);
  const int leaf_size = leaf > 0 ? leaf : (amd_leaves ? 200 : 24);
  auto make_work = [&]() {
    NDWork w(n, xadj, adj, SH);
    w.leaf = leaf_size;
    w.coord_dim = coord_dim;
    w.coords = coords;
    return w;
  };
  struct Range {
    int32_t lo, hi;
  };
  struct Scratch {
    std::vector<int32_t> side;  // 0 = A, 1 = B, 2 = S for vertices of the current range (by position)
    std::vector<int32_t> tmp;
  };
  std::vector<int32_t>& posv = SH.posv;
  // One bisection step: splits range r into A | B | S in place, pushes the open parts onto `todo`, ranges that are not
  // split further onto `leaves`.  Everything it reads and writes belongs to the vertices of r (see NDShared), so open
  // ranges can be processed in any order, and concurrently, with the same result.
  auto process_range = [&](NDWork& W, Scratch& SC, Range r, std::vector<Range>& todo, std::vector<Range>& leaves) {
    std::vector<int32_t>& side = SC.side;
    std::vector<int32_t>& tmp = SC.tmp;
    int32_t m = r.hi - r.lo;
    if (m <= W.leaf) {  // leaf: keeps the order it inherited (BFS order of the parent bisection) unless amd_leaves
      leaves.push_back(r);
      return;
    }
    int32_t lab = W.next_label++;
    for (int32_t k = r.lo; k < r.hi; k++) W.label[W.verts[k]] = lab;
    // --- choose a bipartition A0 | B of the range; mark with stamp: in A0 <=> inA[v] ---
    // inA encoded in lvl sign via separate stamp array `side` indexed by position.
    side.assign(m, 1);
    bool split_done = false;
    std::vector<int32_t>& q = W.queue;
    if (W.coords && W.coord_dim > 0) {
      // coordinate bisection along the widest axis, at the median
      int cd = W.coord_dim, best = 0;
      double bestw = -1;
      for (int d = 0; d < cd; d++) {
        double lo = 1e300, hi = -1e300;
        for (int32_t k = r.lo; k < r.hi; k++) {
          double c = W.coords[(int64_t)W.verts[k] * cd + d];
          lo = std::min(lo, c);
          hi = std::max(hi, c);
        }
        if (hi - lo > bestw) {
          bestw = hi - lo;
          best = d;
        }
      }
      tmp.assign(W.verts.begin() + r.lo, W.verts.begin() + r.hi);
      int32_t half = m / 2;
      std::nth_element(tmp.begin(), tmp.begin() + half, tmp.end(), [&](int32_t a, int32_t b) {
        double ca = W.coords[(int64_t)a * cd + best], cb = W.coords[(int64_t)b * cd + best];
        return ca < cb || (ca == cb && a < b);
      });
      // Equal coordinates (structured meshes: a whole grid line shares the median value) must not be split between the
      // two sides - a cut that jogs through a grid line makes the separator a line longer and thicker.  Snap the cut
      // to the nearest gap between distinct coordinate values (GMRFB_ND_SNAP=0 restores the plain median split).
      static const bool snap = !(std::getenv("GMRFB_ND_SNAP") && std::getenv("GMRFB_ND_SNAP")[0] == '0');
      if (snap && half > 0 && half < m) {
        const double cmed = W.coords[(int64_t)tmp[half] * cd + best];
        int32_t nless = 0, nleq = 0;
        for (int32_t k = 0; k < m; k++) {
          const double c = W.coords[(int64_t)tmp[k] * cd + best];
          nless += (c < cmed);
          nleq += (c <= cmed);
        }
        int32_t cut = -1;
        bool strict = true;
        if (nless > 0 && (nleq >= m || std::abs(nless - m / 2) <= std::abs(nleq - m / 2))) {
          cut = nless;
          strict = true;
        } else if (nleq < m) {
          cut = nleq;
          strict = false;
        }
        if (cut > 0 && cut < m && cut != half) {
          std::stable_partition(tmp.begin(), tmp.end(), [&](int32_t a) {
            const double c = W.coords[(int64_t)a * cd + best];
            return strict ? c < cmed : c <= cmed;
          });
          half = cut;
        }
      }
      // stamp A0 members
      int32_t st = W.next_stamp++;
      for (int32_t k = 0; k < half; k++) W.stamp[tmp[k]] = st;
      // reorder range as tmp (A0 first), set side
      for (int32_t k = 0; k < m; k++) {
        W.verts[r.lo + k] = tmp[k];
        side[k] = (k < half) ? 0 : 1;
      }
      // lvl used as "in A0" flag via stamp
      split_done = true;
      // boundary of A0 -> S
      for (int32_t k = 0; k < half; k++) {
        int32_t v = W.verts[r.lo + k];
        for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) {
          int32_t u = adj[p];
          if (W.label[u] == lab && W.stamp[u] != st) {
            side[k] = 2;
            break;
          }
        }
      }
    } else {
      // pseudo-peripheral root by repeated BFS
      int32_t root = W.verts[r.lo];
      int32_t st = W.next_stamp++;
      int nl = W.bfs(root, lab, st, false);
      if ((int32_t)q.size() == m) {
        for (int it = 0; it < 4; it++) {
          // candidate: minimum in-subset degree among the last level
          int32_t cand = q.back(), cd = INT32_MAX;
          for (int32_t k = (int32_t)q.size() - 1; k >= 0 && W.lvl[q[k]] == nl - 1; k--) {
            int32_t d = W.degree_in(q[k], lab);
            if (d < cd || (d == cd && q[k] < cand)) {
              cd = d;
              cand = q[k];
            }
          }
          int32_t st2 = W.next_stamp++;
          int nl2 = W.bfs(cand, lab, st2, false);
          st = st2;
          root = cand;
          if (nl2 <= nl) {
            nl = nl2;
            break;
          }
          nl = nl2;
        }
      }
      if ((int32_t)q.size() < m) {
        // disconnected: gather whole components into A until half the range is covered; S is empty.
        int32_t half = m / 2;
        int32_t last_start = 0;
        if ((int32_t)q.size() < half) {
          for (int32_t k = r.lo; k < r.hi && (int32_t)q.size() < half; k++) {
            int32_t v = W.verts[k];
            if (W.stamp[v] != st) {
              last_start = (int32_t)q.size();
              W.bfs(v, lab, st, true);
            }
          }
        }
        int32_t na = (int32_t)q.size();
        if (na == m) na = last_start;  // everything gathered: the last component becomes B
        tmp.clear();
        tmp.insert(tmp.end(), q.begin(), q.end());
        for (int32_t k = r.lo; k < r.hi; k++)
          if (W.stamp[W.verts[k]] != st) tmp.push_back(W.verts[k]);
        for (int32_t k = 0; k < m; k++) {
          W.verts[r.lo + k] = tmp[k];
          side[k] = (k < na) ? 0 : 1;
        }
        split_done = true;
      }
      if (!split_done) {
        // connected: q holds the BFS order from `root`, levels in W.lvl, nl levels
        if (nl < 3) {  // (near-)clique: no useful separator, treat as leaf
          leaves.push_back(r);
          return;
        }
        // smallest level index mcut with |levels <= mcut| >= m/2, but keep at least one level on each side
        std::vector<int32_t> cnt(nl, 0);
        for (int32_t v : q) cnt[W.lvl[v]]++;
        int32_t acc = 0, mcut = 0;
        for (int l = 0; l < nl; l++) {
          acc += cnt[l];
          if (acc * 2 >= m) {
            mcut = l;
            break;
          }
        }
        if (mcut >= nl - 1) mcut = nl - 2;
        if (mcut < 1) mcut = 1;
        // tuning aid GMRFB_ND_BALANCE=f (0 < f < 0.5, default off): among the levels whose cut leaves at least a
        // fraction f of the vertices on either side, take the narrowest one instead of the median level
        static const double balance = std::getenv("GMRFB_ND_BALANCE") ? std::atof(std::getenv("GMRFB_ND_BALANCE")) : 0.0;
        if (balance > 0.0 && balance < 0.5) {
          int64_t below = 0;
          int32_t best = mcut;
          for (int l = 0; l < nl - 1; l++) {
            // cutting at level l: A = levels < l, S subset of level l, B = levels > l
            const int64_t above = (int64_t)m - below - cnt[l];
            if (l >= 1 && (double)below >= balance * m && (double)above >= balance * m && cnt[l] < cnt[best]) best = l;
            below += cnt[l];
          }
          mcut = best;
        }
        tmp.assign(q.begin(), q.end());
        for (int32_t k = 0; k < m; k++) {
          int32_t v = tmp[k];
          W.verts[r.lo + k] = v;
          int32_t l = W.lvl[v];
          if (l < mcut)
            side[k] = 0;
          else if (l > mcut)
            side[k] = 1;
          else {
            // level mcut: separator only if adjacent to level mcut+1
            bool b = false;
            for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) {
              int32_t u = adj[p];
              if (W.label[u] == lab && W.lvl[u] == mcut + 1) {
                b = true;
                break;
              }
            }
            side[k] = b ? 2 : 0;
          }
        }
        split_done = true;
      }
    }
    // Separator refinement: the one-sided boundary S of A0 is replaced by a MINIMUM VERTEX COVER of the bipartite
    // graph of cut edges between A0 = A + S and B (Koenig's theorem on a Hopcroft-Karp maximum matching; Liu 1989,
    // Pothen & Fan 1990).  Any cover is a valid vertex separator; the minimum one is never larger than either boundary
    // and straightens the jagged separators that coordinate cuts produce on jittered or unstructured meshes (1M-node
    // bench mesh: nnz(L) 215 M -> 181 M, factor flops 2.1e11 -> 1.5e11, profiles/r01_ordering_study.md).
    // GMRFB_ND_COVER=0 restores the plain boundary separators.
    static const bool cover = !(std::getenv("GMRFB_ND_COVER") && std::getenv("GMRFB_ND_COVER")[0] == '0');
    if (cover && split_done) {
      for (int32_t k = 0; k < m; k++) posv[W.verts[r.lo + k]] = k;
      std::vector<int32_t> sl, bl, bid(m, -1);  // separator positions, boundary-of-B positions, position -> B id
      for (int32_t k = 0; k < m; k++)
        if (side[k] == 2) sl.push_back(k);
      std::vector<int64_t> ex(sl.size() + 1, 0);
      std::vector<int32_t> ea;
      for (size_t i = 0; i < sl.size(); i++) {
        const int32_t v = W.verts[r.lo + sl[i]];
        for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) {
          const int32_t u = adj[p];
          if (W.label[u] != lab) continue;
          const int32_t ku = posv[u];
          if (side[ku] != 1) continue;
          if (bid[ku] < 0) {
            bid[ku] = (int32_t)bl.size();
            bl.push_back(ku);
          }
          ea.push_back(bid[ku]);
        }
        ex[i + 1] = (int64_t)ea.size();
      }
      const int32_t nS = (int32_t)sl.size(), nB = (int32_t)bl.size();
      if (nS > 0 && nB > 0) {
        std::vector<int32_t> ms(nS, -1), mb(nB, -1), dist(nS), itp(nS), bq;
        // Hopcroft-Karp
        for (;;) {
          bq.clear();
          for (int32_t i = 0; i < nS; i++) {
            dist[i] = ms[i] < 0 ? 0 : -1;
            if (ms[i] < 0) bq.push_back(i);
          }
          bool found = false;
          for (size_t h = 0; h < bq.size(); h++) {
            const int32_t i = bq[h];
            for (int64_t p = ex[i]; p < ex[i + 1]; p++) {
              const int32_t j = mb[ea[p]];
              if (j < 0) found = true;
              else if (dist[j] < 0) {
                dist[j] = dist[i] + 1;
                bq.push_back(j);
              }
            }
          }
          if (!found) break;
          for (int32_t i = 0; i < nS; i++) itp[i] = 0;
          // iterative DFS along the layered graph
          std::vector<int32_t> stack;
          for (int32_t root = 0; root < nS; root++) {
            if (ms[root] >= 0) continue;
            stack.assign(1, root);
            while (!stack.empty()) {
              const int32_t i = stack.back();
              if (itp[i] >= (int32_t)(ex[i + 1] - ex[i])) {
                dist[i] = -2;  // dead end
                stack.pop_back();
                continue;
              }
              const int32_t b = ea[ex[i] + itp[i]++];
              const int32_t j = mb[b];
              if (j < 0) {
                // augment along the stack: the edge chosen at each stack vertex is its (itp - 1)-th
                for (int32_t t = (int32_t)stack.size() - 1; t >= 0; t--) {
                  const int32_t si = stack[t];
                  const int32_t bb = ea[ex[si] + itp[si] - 1];
                  mb[bb] = si;
                  ms[si] = bb;
                }
                stack.clear();
              } else if (dist[j] == dist[i] + 1) {
                stack.push_back(j);
              }
            }
          }
        }
        // Koenig: Z = vertices reachable from unmatched S vertices by alternating paths; cover = (S \ Z) + (B & Z)
        std::vector<char> zs(nS, 0), zb(nB, 0);
        bq.clear();
        for (int32_t i = 0; i < nS; i++)
          if (ms[i] < 0) {
            zs[i] = 1;
            bq.push_back(i);
          }
        for (size_t h = 0; h < bq.size(); h++) {
          const int32_t i = bq[h];
          for (int64_t p = ex[i]; p < ex[i + 1]; p++) {
            const int32_t b = ea[p];
            if (zb[b] || ms[i] == b) continue;
            zb[b] = 1;
            const int32_t j = mb[b];
            if (j >= 0 && !zs[j]) {
              zs[j] = 1;
              bq.push_back(j);
            }
          }
        }
        for (int32_t i = 0; i < nS; i++)
          if (zs[i]) side[sl[i]] = 0;  // not needed in the separator: back to A
        for (int32_t b = 0; b < nB; b++)
          if (zb[b]) side[bl[b]] = 2;  // B vertex that covers its cut edges
      }
    }
    // stable partition of the range into A | B | S
    tmp.resize(m);
    int32_t na = 0, nb = 0, ns = 0;
    for (int32_t k = 0; k < m; k++) na += (side[k] == 0), nb += (side[k] == 1), ns += (side[k] == 2);
    if (nb == 0 || na == 0) {
      // degenerate split (everything on one side): accept as leaf to guarantee progress
      if (ns == 0 || na + nb == 0) {
        leaves.push_back(r);
        return;
      }
    }
    int32_t pa = 0, pb = na, ps = na + nb;
    for (int32_t k = 0; k < m; k++) {
      int32_t v = W.verts[r.lo + k];
      if (side[k] == 0)
        tmp[pa++] = v;
      else if (side[k] == 1)
        tmp[pb++] = v;
      else
        tmp[ps++] = v;
    }
    std::copy(tmp.begin(), tmp.end(), W.verts.begin() + r.lo);
    // separator vertices leave the active graph
    for (int32_t k = na + nb; k < m; k++) W.label[W.verts[r.lo + k]] = -1;
    if (nb > 0) todo.push_back({r.lo + na, r.lo + na + nb});
    if (na > 0) todo.push_back({r.lo, r.lo + na});
  };

  // Workers take open ranges (largest first) from a shared counter.  Phase 1: rounds in which every open range is split
  // ONCE, until there are enough ranges to share out (the root alone, then its two parts side by side, then four, ...:
  // the sequential part is about twice the root's bisection instead of one bisection per level).  Phase 2: every range
  // is dissected depth first down to its leaves.  GMRFB_ND_THREADS (default: the hardware threads, at most 16;
  // 1 = no worker threads).
  int nthreads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char* e = std::getenv("GMRFB_ND_THREADS")) nthreads = std::max(1, std::atoi(e));
  if (n < 50000) nthreads = 1;
  std::vector<Range> leaves;  // ranges that are not split further
  std::vector<Range> open;
  open.push_back({0, n});
  auto run = [&](bool to_leaves) {
    std::sort(open.begin(), open.end(), [](const Range& a, const Range& b) { return a.hi - a.lo > b.hi - b.lo; });
    const int nw = (int)std::min<size_t>((size_t)nthreads, open.size());
    std::atomic<size_t> next{0};
    std::vector<std::vector<Range>> wleaves(nw), wopen(nw);
    auto worker = [&](int t) {
      NDWork W = make_work();
      Scratch SC;
      std::vector<Range> todo;
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= open.size()) break;
        if (!to_leaves) {
          process_range(W, SC, open[i], wopen[t], wleaves[t]);
          continue;
        }
        todo.assign(1, open[i]);
        while (!todo.empty()) {
          Range r = todo.back();
          todo.pop_back();
          process_range(W, SC, r, todo, wleaves[t]);
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nw; t++) {
      try {
        pool.emplace_back(worker, t);
      } catch (const std::system_error&) {
        break;  // the ranges come from a shared counter: fewer workers take more of them
      }
    }
    worker(0);
    for (std::thread& th : pool) th.join();
    open.clear();
    for (int t = 0; t < nw; t++) {
      leaves.insert(leaves.end(), wleaves[t].begin(), wleaves[t].end());
      open.insert(open.end(), wopen[t].begin(), wopen[t].end());
    }
  };
  if (nthreads > 1)
    while (!open.empty() && open.size() < (size_t)(2 * nthreads)) run(false);
  if (!open.empty()) run(true);
  NDWork W = make_work();  // (the leaf ordering below reads W.verts)
  if (amd_leaves) {
    // Halo-AMD inside every leaf subdomain: the leaf's vertices are ordered by approximate minimum degree on the leaf's
    // subgraph extended by its halo (the neighbours outside the leaf: separator vertices of the enclosing dissections,
    // all eliminated after the leaf), so fill against the separators is accounted for in the degrees.
    std::vector<int32_t> loc(n, -1), lperm, ladj, halo, order;
    std::vector<int64_t> lxadj;
    std::vector<std::vector<int32_t>> hadj;
    for (const Range& r : leaves) {
      const int32_t m = r.hi - r.lo;
      if (m < 3) continue;
      for (int32_t k = 0; k < m; k++) loc[W.verts[r.lo + k]] = k;
      halo.clear();
      for (int32_t k = 0; k < m; k++) {
        const int32_t v = W.verts[r.lo + k];
        for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) {
          const int32_t u = adj[p];
          if (loc[u] < 0) {
            loc[u] = m + (int32_t)halo.size();
            halo.push_back(u);
          }
        }
      }
      const int32_t nl = m + (int32_t)halo.size();
      hadj.assign(halo.size(), std::vector<int32_t>());
      lxadj.assign(nl + 1, 0);
      ladj.clear();
      for (int32_t k = 0; k < m; k++) {
        const int32_t v = W.verts[r.lo + k];
        for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) {
          const int32_t lu = loc[adj[p]];
          ladj.push_back(lu);
          if (lu >= m) hadj[lu - m].push_back(k);
        }
        lxadj[k + 1] = (int64_t)ladj.size();
      }
      for (size_t h = 0; h < halo.size(); h++) {  // halo rows: their neighbours inside the leaf (symmetric graph)
        ladj.insert(ladj.end(), hadj[h].begin(), hadj[h].end());
        lxadj[m + h + 1] = (int64_t)ladj.size();
      }
      amd_order(nl, lxadj, ladj, lperm, m);
      order.assign(W.verts.begin() + r.lo, W.verts.begin() + r.hi);
      for (int32_t k = 0; k < m; k++) W.verts[r.lo + k] = order[lperm[k]];
      for (int32_t v : order) loc[v] = -1;
      for (int32_t u : halo) loc[u] = -1;
    }
  }
  perm = W.verts;
}

// ---------------------------------------------------------------------------------------------------------
static std::string build_adjacency(int64_t n, const int64_t* colptr, const int64_t* rowval, int base,
                                   std::vector<int64_t>& xadj, std::vector<int32_t>& adj) {
  // symmetrised pattern without the diagonal, sorted, de-duplicated
  int64_t nnz = colptr[n] - base;
  std::vector<int64_t> deg(n + 1, 0);
  for (int64_t c = 0; c < n; c++) {
    int64_t p0 = colptr[c] - base, p1 = colptr[c + 1] - base;
    if (p0 > p1 || p0 < 0 || p1 > nnz) return "colptr is not monotone";
    for (int64_t p = p0; p < p1; p++) {
      int64_t r = rowval[p] - base;
      if (r < 0 || r >= n) return "row index out of range";
      if (p > p0 && rowval[p] <= rowval[p - 1]) return "row indices must be strictly increasing within a column";
      if (r != c) {
        deg[r + 1]++;
        deg[c + 1]++;
      }
    }
  }
  std::vector<int64_t> start(n + 1, 0);
  for (int64_t i = 0; i < n; i++) start[i + 1] = start[i] + deg[i + 1];
  std::vector<int32_t> raw(start[n]);
  std::vector<int64_t> fill(start.begin(), start.end() - 1);
  for (int64_t c = 0; c < n; c++)
    for (int64_t p = colptr[c] - base; p < colptr[c + 1] - base; p++) {
      int64_t r = rowval[p] - base;
      if (r != c) {
        raw[fill[r]++] = (int32_t)c;
        raw[fill[c]++] = (int32_t)r;
      }
    }
  xadj.assign(n + 1, 0);
  adj.clear();
  adj.reserve(raw.size() / 2 + n);
  for (int64_t i = 0; i < n; i++) {
    auto b = raw.begin() + start[i], e = raw.begin() + start[i + 1];
    std::sort(b, e);
    auto e2 = std::unique(b, e);
    adj.insert(adj.end(), b, e2);
    xadj[i + 1] = (int64_t)adj.size();
  }
  return "";
}

// ------------------------------------------------------------------------------- approximate minimum degree ----
// Approximate minimum degree ordering on the quotient graph (Amestoy, Davis & Duff, SIAM J. Matrix Anal. Appl. 17
// (1996) 886-905), restated from the paper: elements absorb the eliminated variables' adjacency, external degrees
// are bounded by  d_i <- min(n - k, d_i + |L_p \ i|, |A_i \ i| + |L_p \ i| + sum_{e in E_i \ p} |L_e \ L_p|)
// with the |L_e \ L_p| obtained in one pass by the timestamp trick, indistinguishable variables are merged into
// supervariables (hash buckets), variables whose adjacency is covered by the new element are mass-eliminated and
// elements that became subsets of it are absorbed (aggressive absorption).  Dense rows (degree above
// max(16, 10 sqrt(n)), SuiteSparse's rule) are detected ONCE, removed and ordered last (amd_order below).  This is the reordering the reference gets from CHOLMOD's default `cholesky(A)`
// call (no `perm`); ties are broken by this implementation's list order, not SuiteSparse's.
namespace {

struct AmdWork {
  int32_t n;
  std::vector<int32_t> iw;             // adjacency storage: for a variable its elements then its variables; for an element its variables
  std::vector<int64_t> pe;             // start of node i's list in iw, or -1 when the node is dead (absorbed)
  std::vector<int32_t> len, elen;      // list length; number of leading elements (variables only)
  std::vector<int32_t> nv;             // supervariable size (0: absorbed variable; negative: currently in L_p)
  std::vector<int32_t> degree;         // approximate external degree (variables) / |L_e| (elements)
  std::vector<int32_t> parent;         // absorbed node -> node that absorbed it
  std::vector<char> is_elem;
  std::vector<int64_t> w;              // timestamps
  std::vector<int32_t> head, next, last;  // degree lists
  int64_t pfree = 0;

  // compact the live lists to the front of iw (no list is being scanned when this runs)
  void collect() {
    for (int32_t i = 0; i < n; i++) {
      if (pe[i] >= 0 && len[i] > 0) {
        const int64_t p = pe[i];
        pe[i] = iw[p];       // save the first entry
        iw[p] = -(i + 1);    // mark the head of i's list
      } else if (pe[i] >= 0) {
        pe[i] = -2;  // live node with an empty list: re-pointed below
      }
    }
    int64_t dst = 0, src = 0;
    while (src < pfree) {
      const int32_t v = iw[src++];
      if (v < 0) {
        const int32_t i = -v - 1;
        iw[dst] = (int32_t)pe[i];
        pe[i] = dst++;
        for (int32_t k = 1; k < len[i]; k++) iw[dst++] = iw[src++];
      }
    }
    for (int32_t i = 0; i < n; i++)
      if (pe[i] == -2) pe[i] = dst;
    pfree = dst;
  }
  void list_remove(int32_t i) {
    const int32_t d = degree[i];
    if (last[i] >= 0) next[last[i]] = next[i]; else head[d] = next[i];
    if (next[i] >= 0) last[next[i]] = last[i];
  }
  void list_insert(int32_t i, int32_t d) {
    next[i] = head[d];
    last[i] = -1;
    if (head[d] >= 0) last[head[d]] = i;
    head[d] = i;
  }
};

}  // namespace

// Vertices nfree .. n-1 (if any) are a HALO: they take part in the quotient graph (so the degrees of the free vertices
// next to them are right) but are never eliminated, merged or reported - the "halo-AMD" used to order the leaf
// subdomains of a nested dissection, whose halo is the surrounding separators (eliminated later).
static void amd_order_impl(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                           std::vector<int32_t>& perm, int32_t nfree, bool detect_dense);

void amd_order(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
               std::vector<int32_t>& perm, int32_t nfree) {
  amd_order_impl(n, xadj, adj, perm, nfree, true);
}

static void amd_order_impl(int32_t n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                           std::vector<int32_t>& perm, int32_t nfree, bool detect_dense) {
  if (nfree < 0 || nfree > n) nfree = n;
  perm.assign(nfree, 0);
  if (nfree == 0) return;
  // Dense rows (degree above max(16, 10 sqrt(n)), the SuiteSparse AMD rule) are taken out and ordered last: a vertex
  // adjacent to nearly everything is rescanned at every one of its neighbours' eliminations (quadratic time) and ends
  // up in the last front whatever the order.
  if (detect_dense) {
    const double thresh = std::max(16.0, 10.0 * std::sqrt((double)n));
    std::vector<int32_t> dense;
    for (int32_t i = 0; i < nfree; i++)
      if ((double)(xadj[i + 1] - xadj[i]) > thresh) dense.push_back(i);
    if (!dense.empty() && (int32_t)dense.size() < nfree) {
      std::vector<int32_t> newid(n, -1), oldid;
      std::vector<char> isd(n, 0);
      for (int32_t d : dense) isd[d] = 1;
      for (int32_t i = 0; i < n; i++)
        if (!isd[i]) {
          newid[i] = (int32_t)oldid.size();
          oldid.push_back(i);
        }
      const int32_t n2 = (int32_t)oldid.size(), nfree2 = nfree - (int32_t)dense.size();
      std::vector<int64_t> x2(n2 + 1, 0);
      std::vector<int32_t> a2;
      a2.reserve(adj.size());
      for (int32_t k = 0; k < n2; k++) {
        const int32_t i = oldid[k];
        for (int64_t p = xadj[i]; p < xadj[i + 1]; p++)
          if (!isd[adj[p]]) a2.push_back(newid[adj[p]]);
        x2[k + 1] = (int64_t)a2.size();
      }
      std::vector<int32_t> p2;
      // dense detection runs once, against the threshold of the original n: the reduced graph is ordered as it is (a
      // recursive re-detection with the smaller n2 would peel further vertices level after level)
      amd_order_impl(n2, x2, a2, p2, nfree2, false);
      std::stable_sort(dense.begin(), dense.end(), [&](int32_t a, int32_t b) {
        return (xadj[a + 1] - xadj[a]) < (xadj[b + 1] - xadj[b]);
      });
      int32_t k = 0;
      for (int32_t v : p2) perm[k++] = oldid[v];
      for (int32_t d : dense) perm[k++] = d;
      return;
    }
  }
  AmdWork W;
  W.n = n;
  const int64_t nnz = xadj[n];
  W.iw.resize((size_t)(nnz + nnz / 4 + 2 * (int64_t)n + 16));
  W.pe.resize(n); W.len.resize(n); W.elen.assign(n, 0); W.nv.assign(n, 1); W.degree.resize(n);
  W.parent.assign(n, -1); W.is_elem.assign(n, 0); W.w.assign(n, 1);
  W.head.assign(n + 1, -1); W.next.assign(n, -1); W.last.assign(n, -1);
  std::copy(adj.begin(), adj.end(), W.iw.begin());
  W.pfree = nnz;
  std::vector<int32_t> pivots;  // elimination sequence of the pivot (super)variables
  pivots.reserve(n);
  int32_t nel = 0;
  for (int32_t i = 0; i < n; i++) {
    W.pe[i] = xadj[i];
    W.len[i] = (int32_t)(xadj[i + 1] - xadj[i]);
    W.degree[i] = W.len[i];
  }
  for (int32_t i = 0; i < n; i++) {
    if (i >= nfree) continue;  // halo vertices are never candidates
    if (W.degree[i] == 0) {  // isolated vertex: eliminate at once
      W.is_elem[i] = 1; W.w[i] = 0; W.pe[i] = -1; W.elen[i] = -1;
      pivots.push_back(i);
      nel++;
    } else {
      W.list_insert(i, W.degree[i]);
    }
  }
  int64_t wflg = 2;
  int32_t mindeg = 1;
  std::vector<int32_t> lme;         // the new element's variable list
  std::vector<int32_t> keep_e, keep_v;
  std::vector<int32_t> hhead(n, -1), hnext(n, -1), hval(n, 0), used;  // hash buckets of this pivot
  auto& iw = W.iw; auto& pe = W.pe; auto& len = W.len; auto& elen = W.elen; auto& nv = W.nv; auto& degree = W.degree;
  auto& w = W.w;

  while (nel < nfree) {
    while (mindeg <= n && W.head[mindeg] < 0) mindeg++;
    const int32_t me = W.head[mindeg];
    W.list_remove(me);
    const int32_t elenme = elen[me];
    int32_t nvpiv = nv[me];
    nel += nvpiv;
    nv[me] = -nvpiv;
    int32_t degme = 0;
    // ---- L_me = (A_me  U  union of L_e, e in E_me) \ me ----
    lme.clear();
    auto take = [&](int32_t i) {
      const int32_t nvi = nv[i];
      if (nvi > 0) {
        degme += nvi;
        nv[i] = -nvi;
        lme.push_back(i);
        if (i < nfree) W.list_remove(i);
      }
    };
    {
      const int64_t p0 = pe[me];
      for (int32_t k = 0; k < elenme; k++) {
        const int32_t e = iw[p0 + k];
        if (pe[e] < 0) continue;
        const int64_t pj = pe[e];
        for (int32_t t = 0; t < len[e]; t++) take(iw[pj + t]);
        pe[e] = -1;  // element e is absorbed into me
        W.parent[e] = me;
        w[e] = 0;
      }
      for (int32_t k = elenme; k < len[me]; k++) take(iw[p0 + k]);
    }
    // store L_me (me's old list and the absorbed elements' lists are garbage now)
    pe[me] = -1;
    if (W.pfree + (int64_t)lme.size() > (int64_t)iw.size()) W.collect();
    if (W.pfree + (int64_t)lme.size() > (int64_t)iw.size()) iw.resize((size_t)(W.pfree + lme.size() + n));
    const int64_t pme1 = W.pfree;
    std::copy(lme.begin(), lme.end(), iw.begin() + pme1);
    W.pfree += (int64_t)lme.size();
    pe[me] = pme1;
    len[me] = (int32_t)lme.size();
    W.is_elem[me] = 1;
    degree[me] = degme;
    elen[me] = -1;
    pivots.push_back(me);
    const int64_t pme2 = pme1 + len[me];
    // timestamps must stay below overflow: w values are at most wflg + n
    if (wflg > ((int64_t)1 << 60)) {
      for (int32_t i = 0; i < n; i++) if (w[i] != 0) w[i] = 1;
      wflg = 2;
    }
    // ---- scan 1: w[e] - wflg = |L_e \ L_me| for every element e adjacent to a variable of L_me ----
    int64_t wmax = wflg;
    for (int64_t p = pme1; p < pme2; p++) {
      const int32_t i = iw[p];
      const int32_t eln = elen[i];
      if (eln <= 0) continue;
      const int32_t nvi = -nv[i];
      const int64_t wnvi = wflg - nvi;
      const int64_t pi = pe[i];
      for (int32_t k = 0; k < eln; k++) {
        const int32_t e = iw[pi + k];
        const int64_t we = w[e];
        if (we >= wflg) w[e] = we - nvi;
        else if (we != 0) {
          w[e] = degree[e] + wnvi;
          if (w[e] > wmax) wmax = w[e];
        }
      }
    }
    // ---- scan 2: degree update, list pruning, hashing ----
    used.clear();
    for (int64_t p = pme1; p < pme2; p++) {
      const int32_t i = iw[p];
      const int64_t p1 = pe[i];
      const int64_t p2 = p1 + elen[i];
      int64_t pn;
      uint32_t hash = 0;
      int64_t deg = 0;
      // the pruned list gets me in front, so the survivors are gathered in a scratch first
      keep_e.clear(); keep_v.clear();
      for (int64_t q = p1; q < p2; q++) {
        const int32_t e = iw[q];
        const int64_t we = w[e];
        if (we == 0) continue;  // absorbed earlier
        const int64_t dext = we - wflg;
        if (dext > 0) {
          deg += dext;
          keep_e.push_back(e);
          hash += (uint32_t)e;
        } else {  // L_e is a subset of L_me: aggressive absorption
          pe[e] = -1;
          W.parent[e] = me;
          w[e] = 0;
        }
      }
      const int64_t pend = p1 + len[i];
      for (int64_t q = p2; q < pend; q++) {
        const int32_t j = iw[q];
        const int32_t nvj = nv[j];
        if (nvj > 0) {  // live and outside L_me
          deg += nvj;
          keep_v.push_back(j);
          hash += (uint32_t)j;
        }
      }
      if (keep_e.empty() && keep_v.empty() && i < nfree) {
        // mass elimination: i's adjacency is me alone
        pe[i] = -1;
        W.parent[i] = me;
        const int32_t nvi = -nv[i];
        degme -= nvi;
        nvpiv += nvi;
        nel += nvi;
        nv[i] = 0;
        elen[i] = -1;
        continue;
      }
      if (i < nfree) degree[i] = (int32_t)std::min<int64_t>(degree[i], deg);
      // new list: me, surviving elements, surviving variables (never longer than the old list: the variables of
      // L_me that i was adjacent to, or at least one absorbed element, made room - otherwise append at pfree)
      const int64_t newlen = 1 + (int64_t)keep_e.size() + (int64_t)keep_v.size();
      int64_t base = p1;
      if (newlen > len[i]) {  // cannot happen (me or an absorbed element always frees a slot); kept as a guard
        if (W.pfree + newlen > (int64_t)iw.size()) iw.resize((size_t)(W.pfree + newlen + n));
        base = W.pfree;
        W.pfree += newlen;
        pe[i] = base;
      }
      pn = base;
      iw[pn++] = me;
      for (int32_t e : keep_e) iw[pn++] = e;
      for (int32_t j : keep_v) iw[pn++] = j;
      elen[i] = 1 + (int32_t)keep_e.size();
      len[i] = (int32_t)newlen;
      if (i >= nfree) continue;  // halo: list kept consistent, no degree, no supervariable detection
      const int32_t hb = (int32_t)(hash % (uint32_t)n);
      hval[i] = (int32_t)hash;
      if (hhead[hb] < 0) used.push_back(hb);
      hnext[i] = hhead[hb];
      hhead[hb] = i;
    }
    degree[me] = degme;
    wflg = wmax + 1;  // every timestamp handed out in scan 1 is now in the past
    // ---- supervariable detection inside each hash bucket ----
    for (int32_t hb : used) {
      for (int32_t i = hhead[hb]; i >= 0; i = hnext[i]) {
        if (nv[i] == 0 || hnext[i] < 0) continue;
        // mark i's list
        wflg++;
        const int64_t pi = pe[i];
        for (int32_t k = 1; k < len[i]; k++) w[iw[pi + k]] = (w[iw[pi + k]] == 0 ? 0 : wflg);
        int32_t prev = i;
        for (int32_t j = hnext[i]; j >= 0; j = hnext[j]) {
          bool same = nv[j] != 0 && hval[j] == hval[i] && len[j] == len[i] && elen[j] == elen[i];
          if (same) {
            const int64_t pj = pe[j];
            for (int32_t k = 1; k < len[j] && same; k++) same = (w[iw[pj + k]] == wflg);
          }
          if (same) {  // j is indistinguishable from i
            pe[j] = -1;
            W.parent[j] = i;
            nv[i] += nv[j];  // both negative: sizes add
            nv[j] = 0;
            elen[j] = -1;
            hnext[prev] = hnext[j];
          } else {
            prev = j;
          }
        }
      }
      hhead[hb] = -1;
    }
    wflg += 2;
    // ---- finalise the new element and the degrees of its variables ----
    int64_t p = pme1;
    const int32_t nleft = (nfree - nel) + (n - nfree);
    for (int64_t q = pme1; q < pme2; q++) {
      const int32_t i = iw[q];
      const int32_t nvi = -nv[i];
      if (nvi <= 0) continue;  // absorbed above
      nv[i] = nvi;
      if (i >= nfree) {  // halo member of the new element
        iw[p++] = i;
        continue;
      }
      int64_t deg = (int64_t)degree[i] + degme - nvi;
      deg = std::min<int64_t>(deg, nleft - nvi);
      if (deg < 1) deg = 1;  // only possible for the last variables; keeps list 0 free for nothing special
      if (deg > n) deg = n;
      degree[i] = (int32_t)deg;
      W.list_insert(i, (int32_t)deg);
      if (deg < mindeg) mindeg = (int32_t)deg;
      iw[p++] = i;
    }
    nv[me] = nvpiv;
    len[me] = (int32_t)(p - pme1);
    if (len[me] == 0) {
      pe[me] = -1;
      w[me] = 0;
    }
    // make sure the timestamp of the new element is a live one
    if (len[me] > 0) w[me] = 1;
  }
  // ---- ordering: pivots in elimination order, each followed by the variables absorbed into it ----
  std::vector<int32_t> rep(n, -1);
  std::vector<int32_t> cnt_head(n, -1), cnt_next(n, -1);
  std::vector<char> is_pivot(n, 0);
  for (int32_t v : pivots) is_pivot[v] = 1;
  for (int32_t i = nfree - 1; i >= 0; i--) {
    if (is_pivot[i]) continue;
    int32_t r = i;
    while (!is_pivot[r]) r = W.parent[r];  // variables are absorbed by variables that end up as pivots or by pivots
    cnt_next[i] = cnt_head[r];
    cnt_head[r] = i;
  }
  int32_t k = 0;
  for (int32_t v : pivots) {
    perm[k++] = v;
    for (int32_t i = cnt_head[v]; i >= 0; i = cnt_next[i]) perm[k++] = i;
  }
}

// GMRFB_SYM_TIMING=1: wall time of every phase of the analysis on stderr (tuning aid)
struct PhaseTimer {
  bool on;
  std::chrono::steady_clock::time_point t0;
  PhaseTimer() : on(std::getenv("GMRFB_SYM_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[gmrfb analyze] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

std::string analyze_pattern(int64_t n64, const int64_t* colptr, const int64_t* rowval, const int64_t* perm_in,
                            const AnalyzeOptions& opt, Symbolic& S, SymDevice* dev) {
  if (n64 < 0 || n64 > (int64_t)2000000000) return "n out of range";
  if (!colptr || (n64 > 0 && !rowval)) return "null pattern";
  if (opt.base != 0 && opt.base != 1) return "base must be 0 or 1";
  if (opt.storage < 0 || opt.storage > 2) return "bad storage kind";
  const int32_t n = (int32_t)n64;
  const int base = opt.base;
  S = Symbolic();
  S.n = n;
  S.base = base;
  S.storage = opt.storage;
  S.nnzA = colptr[n] - base;
  if (colptr[0] != base) return "colptr[0] must equal base";

  std::vector<int64_t> xadj;
  std::vector<int32_t> adj;
  PhaseTimer pt;
  std::string err;
  bool on_dev = false;
  if (dev && opt.storage == 0) {
    // structurally symmetric input with both triangles stored: validated and turned into the adjacency on the GPU
    err = dev->adjacency(n, colptr, rowval, base, xadj, adj, on_dev);
    if (!err.empty()) return err;
  }
  if (!on_dev) {
    dev = nullptr;  // LOWER / UPPER storage or an unsymmetric pattern: the whole analysis takes the host path
    err = build_adjacency(n, colptr, rowval, base, xadj, adj);
    if (!err.empty()) return err;
  }
  pt.lap(on_dev ? "adjacency (GPU)" : "adjacency");

  // ---- ordering ----
  S.perm_user.resize(n);
  if (opt.ordering_kind == 0) {
    if (!perm_in) return "ORDER_GIVEN needs a permutation";
    std::vector<char> seen(n, 0);
    for (int32_t k = 0; k < n; k++) {
      int64_t v = perm_in[k] - base;
      if (v < 0 || v >= n || seen[v]) return "perm is not a permutation";
      seen[v] = 1;
      S.perm_user[k] = (int32_t)v;
    }
  } else if (opt.ordering_kind == 1) {
    std::iota(S.perm_user.begin(), S.perm_user.end(), 0);
  } else if (opt.ordering_kind == 2) {
    if (opt.coords && (opt.coord_dim < 1 || opt.coord_dim > 3)) return "coord_dim must be 1..3";
    nested_dissection(n, xadj, adj, opt.nd_leaf > 0 ? opt.nd_leaf : (std::getenv("GMRFB_ND_LEAF") ? std::atoi(std::getenv("GMRFB_ND_LEAF")) : 0), opt.coords ? opt.coord_dim : 0, opt.coords, S.perm_user, false);
  } else if (opt.ordering_kind == 3) {
    amd_order(n, xadj, adj, S.perm_user);
  } else if (opt.ordering_kind == 4) {
    if (opt.coords && (opt.coord_dim < 1 || opt.coord_dim > 3)) return "coord_dim must be 1..3";
    nested_dissection(n, xadj, adj, opt.nd_leaf > 0 ? opt.nd_leaf : (std::getenv("GMRFB_ND_LEAF") ? std::atoi(std::getenv("GMRFB_ND_LEAF")) : 0), opt.coords ? opt.coord_dim : 0, opt.coords, S.perm_user, true);
  } else {
    return "unknown ordering kind";
  }

  pt.lap("ordering");
  // ---- permuted adjacency (perm_user ordering) ----
  std::vector<int32_t> ipu(n);
  for (int32_t k = 0; k < n; k++) ipu[S.perm_user[k]] = k;
  std::vector<int64_t> pxadj(n + 1, 0);
  std::vector<int32_t> padj(adj.size());
  if (dev) {
    err = dev->permuted_adjacency(S.perm_user, ipu, pxadj, padj);
    if (!err.empty()) return err;
  } else {
    for (int32_t k = 0; k < n; k++) pxadj[k + 1] = pxadj[k] + (xadj[S.perm_user[k] + 1] - xadj[S.perm_user[k]]);
    for (int32_t k = 0; k < n; k++) {
      int32_t v = S.perm_user[k];
      int64_t o = pxadj[k];
      for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) padj[o++] = ipu[adj[p]];
      std::sort(padj.begin() + pxadj[k], padj.begin() + pxadj[k + 1]);
    }
  }
  S.nnz_lower_A = (int64_t)adj.size() / 2 + n;

  pt.lap("permuted adjacency");
  // ---- etree, postorder, column counts in the perm_user ordering ----
  etree_lower(n, pxadj, padj, S.parent_user);
  postorder_tree(n, S.parent_user, S.post);
  column_counts(n, pxadj, padj, S.parent_user, S.post, S.colcount_user);
  S.nnzL = 0;
  S.flops = 0;
  for (int32_t j = 0; j < n; j++) {
    S.nnzL += S.colcount_user[j];
    S.flops += (double)S.colcount_user[j] * (double)S.colcount_user[j];
  }

  pt.lap("etree + postorder + colcounts");
  // ---- internal (postordered) numbering ----
  S.ipost.resize(n);
  S.perm.resize(n);
  S.iperm.resize(n);
  for (int32_t k = 0; k < n; k++) S.ipost[S.post[k]] = k;
  for (int32_t k = 0; k < n; k++) {
    S.perm[k] = S.perm_user[S.post[k]];
    S.iperm[S.perm[k]] = k;
  }
  S.parent.resize(n);
  S.colcount.resize(n);
  for (int32_t k = 0; k < n; k++) {
    int32_t pu = S.parent_user[S.post[k]];
    S.parent[k] = pu < 0 ? -1 : S.ipost[pu];
    S.colcount[k] = S.colcount_user[S.post[k]];
  }
  // internal adjacency (needed for the supernode row structures): neighbours above each column
  std::vector<int64_t> ixadj(n + 1, 0);
  std::vector<int32_t> iadj;  // for internal column k: sorted internal neighbours i > k
  if (dev) {
    err = dev->internal_adjacency(S.post, S.ipost, ixadj, iadj);
    if (!err.empty()) return err;
  } else {
    std::vector<int64_t> cnt(n + 1, 0);
    for (int32_t k = 0; k < n; k++) {
      int32_t ku = S.post[k];
      int64_t c = 0;
      for (int64_t p = pxadj[ku]; p < pxadj[ku + 1]; p++) c += (S.ipost[padj[p]] > k);
      cnt[k + 1] = c;
    }
    for (int32_t k = 0; k < n; k++) ixadj[k + 1] = ixadj[k] + cnt[k + 1];
    iadj.resize(ixadj[n]);
    for (int32_t k = 0; k < n; k++) {
      int32_t ku = S.post[k];
      int64_t o = ixadj[k];
      for (int64_t p = pxadj[ku]; p < pxadj[ku + 1]; p++) {
        int32_t i = S.ipost[padj[p]];
        if (i > k) iadj[o++] = i;
      }
      std::sort(iadj.begin() + ixadj[k], iadj.begin() + ixadj[k + 1]);
    }
  }

  pt.lap("internal numbering/adjacency");
  // ---- fundamental supernodes ----
  std::vector<int32_t> nchild(n, 0);
  for (int32_t k = 0; k < n; k++)
    if (S.parent[k] >= 0) nchild[S.parent[k]]++;
  struct SN {
    int32_t first, last;  // column range
    int64_t true_nnz;     // Σ colcount over its columns
    int32_t order;        // front order = colcount of first column after merges (s + r)
  };
  std::vector<SN> sn;
  sn.reserve(n / 2 + 1);
  // tuning aids (only when the caller leaves the options at their defaults): GMRFB_RELAX_SMALL, GMRFB_RELAX_ZEROS,
  // GMRFB_RELAX_TIERS="z32,z64,z128,zbig" (zero fractions of the tiered default rule)
  auto env_num = [](const char* name, double dflt) {
    const char* e = std::getenv(name);
    return e ? std::atof(e) : dflt;
  };
  const int relax_small = opt.relax_small > 0 ? opt.relax_small : (int)env_num("GMRFB_RELAX_SMALL", 16);
  const double relax_zeros = opt.relax_zeros > 0 ? opt.relax_zeros : env_num("GMRFB_RELAX_ZEROS", 0.0);  // 0 => tiered rule
  double tier32 = 0.5, tier64 = 0.3, tier128 = 0.05, tierbig = 0.05;  // tier64: B200 sweep, profiles/r01_amalgamation_sweep.md
  if (const char* e = std::getenv("GMRFB_RELAX_TIERS")) std::sscanf(e, "%lf,%lf,%lf,%lf", &tier32, &tier64, &tier128, &tierbig);
  auto trapezoid = [](int64_t s, int64_t d) { return s * d - s * (s - 1) / 2; };
  for (int32_t k = 0; k < n;) {
    int32_t f = k;
    int64_t tn = S.colcount[k];
    while (k + 1 < n && S.parent[k] == k + 1 && S.colcount[k + 1] == S.colcount[k] - 1 && nchild[k + 1] == 1) {
      k++;
      tn += S.colcount[k];
    }
    SN cur{f, k, tn, S.colcount[f]};
    // relaxed amalgamation with the immediately preceding supernode when it is a child of `cur`
    while (!sn.empty()) {
      SN& c = sn.back();
      int32_t pc = S.parent[c.last];
      if (pc < cur.first || pc > cur.last) break;  // not a child
      int64_t s_new = (cur.last - cur.first + 1) + (c.last - c.first + 1);
      int64_t d_new = (c.last - c.first + 1) + cur.order;
      int64_t stored = trapezoid(s_new, d_new);
      int64_t truen = c.true_nnz + cur.true_nnz;
      double zfrac = (double)(stored - truen) / (double)stored;
      bool merge;
      if (s_new <= relax_small)
        merge = true;
      else if (relax_zeros > 0)
        merge = zfrac <= relax_zeros;
      else if (s_new <= 32)
        merge = zfrac <= tier32;
      else if (s_new <= 64)
        merge = zfrac <= tier64;
      else if (s_new <= 128)
        merge = zfrac <= tier128;
      else
        merge = zfrac <= tierbig;
      if (!merge) break;
      cur.first = c.first;
      cur.true_nnz = truen;
      cur.order = (int32_t)d_new;
      sn.pop_back();
    }
    sn.push_back(cur);
    k++;
  }
  S.nsuper = (int32_t)sn.size();
  S.sptr.resize(S.nsuper + 1);
  S.snode.resize(n);
  for (int32_t s = 0; s < S.nsuper; s++) {
    S.sptr[s] = sn[s].first;
    for (int32_t k = sn[s].first; k <= sn[s].last; k++) S.snode[k] = s;
  }
  S.sptr[S.nsuper] = n;
  S.sparent.assign(S.nsuper, -1);
  for (int32_t s = 0; s < S.nsuper; s++) {
    int32_t p = S.parent[sn[s].last];
    S.sparent[s] = p < 0 ? -1 : S.snode[p];
  }
  // children lists
  S.child_ptr.assign(S.nsuper + 1, 0);
  for (int32_t s = 0; s < S.nsuper; s++)
    if (S.sparent[s] >= 0) S.child_ptr[S.sparent[s] + 1]++;
  for (int32_t s = 0; s < S.nsuper; s++) S.child_ptr[s + 1] += S.child_ptr[s];
  S.child_idx.resize(S.child_ptr[S.nsuper]);
  {
    std::vector<int32_t> fillp(S.child_ptr.begin(), S.child_ptr.end() - 1);
    for (int32_t s = 0; s < S.nsuper; s++)
      if (S.sparent[s] >= 0) S.child_idx[fillp[S.sparent[s]]++] = s;
  }

  pt.lap("supernodes");
  // ---- row structures (bottom-up union) ----
  S.rptr.assign(S.nsuper + 1, 0);
  S.rows.clear();
  S.rows.reserve((size_t)n * 4);
  {
    std::vector<int32_t> mark(n, -1);
    std::vector<int32_t> below;
    for (int32_t s = 0; s < S.nsuper; s++) {
      int32_t f = S.sptr[s], l = S.sptr[s + 1] - 1;
      below.clear();
      for (int32_t k = f; k <= l; k++) {
        S.rows.push_back(k);
        mark[k] = s;
      }
      for (int32_t k = f; k <= l; k++)
        for (int64_t p = ixadj[k]; p < ixadj[k + 1]; p++) {
          int32_t i = iadj[p];
          if (i > l && mark[i] != s) {
            mark[i] = s;
            below.push_back(i);
          }
        }
      for (int32_t ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ci++) {
        int32_t c = S.child_idx[ci];
        int32_t sc = S.sptr[c + 1] - S.sptr[c];
        for (int64_t p = S.rptr[c] + sc; p < S.rptr[c + 1]; p++) {
          int32_t i = S.rows[p];
          if (i > l && mark[i] != s) {
            mark[i] = s;
            below.push_back(i);
          }
        }
      }
      std::sort(below.begin(), below.end());
      S.rows.insert(S.rows.end(), below.begin(), below.end());
      S.rptr[s + 1] = (int64_t)S.rows.size();
      if ((int32_t)(S.rptr[s + 1] - S.rptr[s]) != sn[s].order)
        return "internal error: column counts disagree with the supernodal row structure";
    }
  }
  pt.lap("row structures");
  // ---- relmap, layout, levels ----
  S.relmap.assign(S.rows.size(), -1);
  S.ld.resize(S.nsuper);
  S.foff.resize(S.nsuper);
  S.level.assign(S.nsuper, 0);
  S.arena = 0;
  S.nnzL_stored = 0;
  S.max_front = 0;
  for (int32_t s = 0; s < S.nsuper; s++) {
    int32_t d = S.front_order(s), sc = S.ncols(s);
    S.max_front = std::max(S.max_front, d);
    S.ld[s] = (d + 1) & ~1;  // even leading dimension: 16-byte aligned columns
    S.foff[s] = S.arena;
    S.arena += (int64_t)S.ld[s] * d;
    S.arena = (S.arena + 15) & ~(int64_t)15;  // 128-byte aligned fronts
    S.nnzL_stored += trapezoid(sc, d);
    int32_t p = S.sparent[s];
    if (p >= 0) {
      // positions of s's below-rows inside the parent's (sorted) row list (with `dev`: k_relmap, below)
      int64_t q = S.rptr[p];
      const int64_t qe = S.rptr[p + 1];
      if (!dev)
        for (int64_t k = S.rptr[s] + sc; k < S.rptr[s + 1]; k++) {
          int32_t i = S.rows[k];
          while (q < qe && S.rows[q] < i) q++;
          if (q >= qe || S.rows[q] != i) return "internal error: child row missing from parent front";
          S.relmap[k] = (int32_t)(q - S.rptr[p]);
        }
      S.level[p] = std::max(S.level[p], S.level[s] + 1);
    }
  }
  int32_t nlev = 0;
  for (int32_t s = 0; s < S.nsuper; s++) nlev = std::max(nlev, S.level[s] + 1);
  S.levels.assign(nlev, Level());
  for (int32_t s = 0; s < S.nsuper; s++) S.levels[S.level[s]].snodes.push_back(s);

  pt.lap("relmap + layout + levels");
  // ---- scatter map of the user's stored entries into the frontal arena ----
  if (dev) {
    err = dev->maps(S);  // relmap + amap on the GPU
    pt.lap("relmap + amap (GPU)");
    return err;
  }
  S.amap.assign(S.nnzA, -1);
  for (int64_t c = 0; c < n; c++) {
    int32_t jc = S.iperm[c];
    for (int64_t p = colptr[c] - base; p < colptr[c + 1] - base; p++) {
      int64_t r = rowval[p] - base;
      int32_t ir = S.iperm[r];
      int32_t i = ir, j = jc;
      if (opt.storage == 0) {
        // both triangles stored: take the copy that lands in the lower triangle of the permuted matrix;
        // the diagonal is taken once
        if (i < j) continue;
      } else {
        if (opt.storage == 1 && r < c) return "STORAGE_LOWER matrix has an entry above the diagonal";
        if (opt.storage == 2 && r > c) return "STORAGE_UPPER matrix has an entry below the diagonal";
        if (i < j) std::swap(i, j);
      }
      int32_t s = S.snode[j];
      int32_t f = S.sptr[s], l = S.sptr[s + 1] - 1;
      int64_t lr;
      if (i <= l) {
        lr = i - f;
      } else {
        auto b = S.rows.begin() + S.rptr[s] + (l - f + 1), e = S.rows.begin() + S.rptr[s + 1];
        auto it = std::lower_bound(b, e, i);
        if (it == e || *it != i) return "internal error: matrix entry outside the symbolic structure";
        lr = (it - (S.rows.begin() + S.rptr[s]));
      }
      S.amap[p] = S.foff[s] + (int64_t)(j - f) * S.ld[s] + lr;
    }
  }
  pt.lap("amap");
  return "";
}

}  // namespace gmrfb
