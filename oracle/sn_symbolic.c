/*
 * oracle/sn_symbolic.c — TEST / BASELINE INFRASTRUCTURE ONLY (never linked into libgmrfb).
 *
 * Supernodal symbolic analysis for the CPU baseline, independent of the product's host code (csrc/symbolic.cpp):
 * what CHOLMOD's `cholmod_analyze_p` + `cholmod_super_symbolic` do for the reference's
 * `cholesky(Symmetric(A); perm=p)` (scripts/solve_burger.jl:147, scripts/darcy/solve_darcy_gmrf-fem.jl:174), restated
 * from the published algorithms: Liu's elimination tree, an etree postorder, row-subtree column counts, fundamental
 * supernodes, relaxed amalgamation with CHOLMOD's default thresholds (nrelax = 4, 16, 48; zrelax = 0.8, 0.1, 0.05),
 * and the supernodal row structures by a bottom-up union.
 *
 * Input : symmetric pattern with both triangles (CSC, 0-based int64) and a fill-reducing permutation perm (new->old).
 * Output: perm_out (the permutation followed by the postorder: supernodes are contiguous column ranges, children
 *         precede parents), sptr, rptr and rows in the layout oracle/supernodal_chol.c consumes.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t idx;

static struct {
  idx ns;
  idx* sptr;
  idx* rptr;
  idx* rows;
  idx nrows;
  idx nnzL;
  double flops;
} R;

static int cmp_idx(const void* a, const void* b) {
  const idx x = *(const idx*)a, y = *(const idx*)b;
  return x < y ? -1 : x > y;
}

/* upper-triangular pattern (rows i <= j) of P A P' by columns; returns 0 on success */
static int upper_pattern(idx n, const idx* Ap, const idx* Ai, const idx* inv, idx** Cp_out, idx** Ci_out) {
  idx* Cp = (idx*)calloc((size_t)n + 1, sizeof(idx));
  if (!Cp) return -1;
  for (idx c = 0; c < n; c++)
    for (idx p = Ap[c]; p < Ap[c + 1]; p++)
      if (inv[Ai[p]] <= inv[c]) Cp[inv[c] + 1]++;
  for (idx k = 0; k < n; k++) Cp[k + 1] += Cp[k];
  idx* Ci = (idx*)malloc(sizeof(idx) * (size_t)(Cp[n] > 0 ? Cp[n] : 1));
  idx* w = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  if (!Ci || !w) return -1;
  memcpy(w, Cp, sizeof(idx) * (size_t)n);
  for (idx c = 0; c < n; c++)
    for (idx p = Ap[c]; p < Ap[c + 1]; p++)
      if (inv[Ai[p]] <= inv[c]) Ci[w[inv[c]]++] = inv[Ai[p]];
  free(w);
  *Cp_out = Cp;
  *Ci_out = Ci;
  return 0;
}

static void etree(idx n, const idx* Cp, const idx* Ci, idx* parent) {
  idx* anc = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  for (idx k = 0; k < n; k++) {
    parent[k] = -1;
    anc[k] = -1;
    for (idx p = Cp[k]; p < Cp[k + 1]; p++) {
      idx i = Ci[p];
      while (i != -1 && i < k) {
        idx next = anc[i];
        anc[i] = k;
        if (next == -1) parent[i] = k;
        i = next;
      }
    }
  }
  free(anc);
}

/* depth-first postorder of the forest `parent`, children visited in increasing order */
static void postorder(idx n, const idx* parent, idx* post) {
  idx* head = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* next = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* stack = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  for (idx k = 0; k < n; k++) head[k] = -1;
  for (idx k = n - 1; k >= 0; k--)
    if (parent[k] != -1) {
      next[k] = head[parent[k]];
      head[parent[k]] = k;
    }
  idx cnt = 0;
  for (idx r = 0; r < n; r++) {
    if (parent[r] != -1) continue;
    idx top = 0;
    stack[0] = r;
    while (top >= 0) {
      idx v = stack[top], c = head[v];
      if (c == -1) {
        post[cnt++] = v;
        top--;
      } else {
        head[v] = next[c];
        stack[++top] = c;
      }
    }
  }
  free(head);
  free(next);
  free(stack);
}

/* Phase 1: analysis.  perm_out[n] receives the postordered permutation; returns the number of supernodes (< 0: error). */
idx orc_sn_symbolic(idx n, const idx* Ap, const idx* Ai, const idx* perm, idx* perm_out) {
  idx *inv = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1)), *Cp = NULL, *Ci = NULL;
  idx* parent = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* post = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  if (!inv || !parent || !post) return -1;
  for (idx k = 0; k < n; k++) inv[perm[k]] = k;
  if (upper_pattern(n, Ap, Ai, inv, &Cp, &Ci)) return -1;
  etree(n, Cp, Ci, parent);
  postorder(n, parent, post);
  free(Cp);
  free(Ci);
  for (idx k = 0; k < n; k++) perm_out[k] = perm[post[k]];
  for (idx k = 0; k < n; k++) inv[perm_out[k]] = k;
  if (upper_pattern(n, Ap, Ai, inv, &Cp, &Ci)) return -1;
  etree(n, Cp, Ci, parent); /* now a postordered tree: parent[k] > k, subtrees contiguous */
  /* column counts by row-subtree traversal: row k of L = etree reach of the entries of column k of the upper triangle */
  idx* cc = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* mark = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* nchild = (idx*)calloc((size_t)(n > 0 ? n : 1), sizeof(idx));
  if (!cc || !mark || !nchild) return -1;
  for (idx k = 0; k < n; k++) {
    cc[k] = 1;
    mark[k] = -1;
    if (parent[k] != -1) nchild[parent[k]]++;
  }
  for (idx k = 0; k < n; k++) {
    mark[k] = k;
    for (idx p = Cp[k]; p < Cp[k + 1]; p++)
      for (idx i = Ci[p]; i < k && mark[i] != k; i = parent[i]) {
        cc[i]++;
        mark[i] = k;
      }
  }
  R.nnzL = 0;
  R.flops = 0;
  for (idx k = 0; k < n; k++) {
    R.nnzL += cc[k];
    R.flops += (double)cc[k] * (double)cc[k];
  }
  /* fundamental supernodes, then relaxed amalgamation of a supernode with the supernode that directly follows it
   * when that one holds its etree parent (the last child of a chain): w = columns, h = front order, z = explicit zeros */
  idx* first = (idx*)malloc(sizeof(idx) * (size_t)(n + 1));
  idx* wd = (idx*)malloc(sizeof(idx) * (size_t)(n + 1));
  idx* ht = (idx*)malloc(sizeof(idx) * (size_t)(n + 1));
  double* zz = (double*)malloc(sizeof(double) * (size_t)(n + 1));
  if (!first || !wd || !ht || !zz) return -1;
  idx ns = 0;
  for (idx j = 0; j < n; j++) {
    const int cont = j > 0 && parent[j - 1] == j && cc[j - 1] == cc[j] + 1 && nchild[j] == 1;
    if (cont) {
      wd[ns - 1]++;
      continue;
    }
    /* a new fundamental supernode starts at j; first try to merge the finished one into ... (done below, on close) */
    first[ns] = j;
    wd[ns] = 1;
    ht[ns] = cc[j];
    zz[ns] = 0;
    ns++;
  }
  /* merge pass over the fundamental supernodes (left to right, with a stack of open results) */
  idx m = 0; /* number of supernodes in the result so far; arrays reused in place (m <= current index) */
  for (idx s = 0; s < ns; s++) {
    idx f = first[s], w = wd[s], h = ht[s];
    double z = zz[s];
    while (m > 0) {
      const idx pf = first[m - 1], pw = wd[m - 1], ph = ht[m - 1];
      const idx plast = pf + pw - 1;
      if (plast + 1 != f || parent[plast] < f || parent[plast] > f + w - 1) break;
      const double nw = (double)(pw + w), nh = (double)pw + (double)h;
      const double tnew = nw * nh - nw * (nw - 1) / 2;
      const double told = ((double)pw * ph - (double)pw * (pw - 1) / 2) + ((double)w * h - (double)w * (w - 1) / 2);
      const double znew = tnew - told + zz[m - 1] + z;
      const double frac = znew / tnew;
      const int ok = nw <= 4 || (nw <= 16 && frac <= 0.8) || (nw <= 48 && frac <= 0.1) || frac <= 0.05;
      if (!ok) break;
      f = pf;
      w = pw + w;
      h = pw + h;
      z = znew;
      m--;
    }
    first[m] = f;
    wd[m] = w;
    ht[m] = h;
    zz[m] = z;
    m++;
  }
  ns = m;
  /* row structures, bottom-up: own columns, then sorted union of the below-diagonal pattern of A's columns and of the
   * children's below rows */
  idx* snode = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* sptr = (idx*)malloc(sizeof(idx) * (size_t)(ns + 1));
  idx* rptr = (idx*)malloc(sizeof(idx) * (size_t)(ns + 1));
  idx* chead = (idx*)malloc(sizeof(idx) * (size_t)(ns > 0 ? ns : 1));
  idx* cnext = (idx*)malloc(sizeof(idx) * (size_t)(ns > 0 ? ns : 1));
  if (!snode || !sptr || !rptr || !chead || !cnext) return -1;
  idx cap = 0;
  for (idx s = 0; s < ns; s++) {
    sptr[s] = first[s];
    cap += ht[s];
    chead[s] = -1;
    for (idx j = first[s]; j < first[s] + wd[s]; j++) snode[j] = s;
  }
  sptr[ns] = n;
  idx* rows = (idx*)malloc(sizeof(idx) * (size_t)(cap > 0 ? cap : 1));
  if (!rows) return -1;
  /* lower adjacency of the permuted matrix by columns = transpose of the upper pattern: build CSR of the upper = CSC
   * of the lower */
  idx* Lp = (idx*)calloc((size_t)n + 1, sizeof(idx));
  for (idx j = 0; j < n; j++)
    for (idx p = Cp[j]; p < Cp[j + 1]; p++)
      if (Ci[p] < j) Lp[Ci[p] + 1]++;
  for (idx k = 0; k < n; k++) Lp[k + 1] += Lp[k];
  idx* Li = (idx*)malloc(sizeof(idx) * (size_t)(Lp[n] > 0 ? Lp[n] : 1));
  idx* wq = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  if (!Lp || !Li || !wq) return -1;
  memcpy(wq, Lp, sizeof(idx) * (size_t)n);
  for (idx j = 0; j < n; j++)
    for (idx p = Cp[j]; p < Cp[j + 1]; p++)
      if (Ci[p] < j) Li[wq[Ci[p]]++] = j; /* column Ci[p] has the below-diagonal row j */
  for (idx k = 0; k < n; k++) mark[k] = -1;
  idx pos = 0;
  for (idx s = 0; s < ns; s++) {
    const idx f = sptr[s], l = sptr[s + 1] - 1;
    rptr[s] = pos;
    for (idx j = f; j <= l; j++) rows[pos++] = j;
    const idx below0 = pos;
    for (idx j = f; j <= l; j++)
      for (idx p = Lp[j]; p < Lp[j + 1]; p++) {
        const idx i = Li[p];
        if (i > l && mark[i] != s) {
          mark[i] = s;
          if (pos >= cap) return -2;
          rows[pos++] = i;
        }
      }
    for (idx c = chead[s]; c != -1; c = cnext[c])
      for (idx p = rptr[c] + (sptr[c + 1] - sptr[c]); p < rptr[c + 1]; p++) {
        const idx i = rows[p];
        if (i > l && mark[i] != s) {
          mark[i] = s;
          if (pos >= cap) return -2;
          rows[pos++] = i;
        }
      }
    qsort(rows + below0, (size_t)(pos - below0), sizeof(idx), cmp_idx);
    rptr[s + 1] = pos;
    if (pos > below0) { /* parent supernode = the one holding the first below row */
      const idx ps = snode[rows[below0]];
      cnext[s] = chead[ps];
      chead[ps] = s;
    }
  }
  free(inv);
  free(parent);
  free(post);
  free(Cp);
  free(Ci);
  free(cc);
  free(mark);
  free(nchild);
  free(first);
  free(wd);
  free(ht);
  free(zz);
  free(snode);
  free(chead);
  free(cnext);
  free(Lp);
  free(Li);
  free(wq);
  free(R.sptr);
  free(R.rptr);
  free(R.rows);
  R.ns = ns;
  R.sptr = sptr;
  R.rptr = rptr;
  R.rows = rows;
  R.nrows = pos;
  return ns;
}

idx orc_sn_symbolic_nrows(void) { return R.nrows; }
idx orc_sn_symbolic_nnzL(void) { return R.nnzL; }
double orc_sn_symbolic_flops(void) { return R.flops; }

/* Phase 2: copy out sptr[ns+1], rptr[ns+1], rows[nrows] and release the cached result. */
void orc_sn_symbolic_fetch(idx* sptr, idx* rptr, idx* rows) {
  memcpy(sptr, R.sptr, sizeof(idx) * (size_t)(R.ns + 1));
  memcpy(rptr, R.rptr, sizeof(idx) * (size_t)(R.ns + 1));
  memcpy(rows, R.rows, sizeof(idx) * (size_t)R.nrows);
  free(R.sptr);
  free(R.rptr);
  free(R.rows);
  R.sptr = R.rptr = R.rows = NULL;
  R.ns = R.nrows = 0;
}
