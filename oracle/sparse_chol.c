/*
 * oracle/sparse_chol.c — TEST INFRASTRUCTURE ONLY (never linked or imported by the product library).
 *
 * CPU restatement of the sparse-Cholesky path that the reference reaches through Julia's
 * `cholesky(Symmetric(A); perm=p)` / `F \ b` / `F.UP \ z` / `F.PtL \ b`
 * (reference call sites: scripts/solve_burger.jl:147-148, scripts/darcy/solve_darcy_fem.jl:93,
 * src/tridiagonal_cholesky.jl:20-22,39-41).  The arithmetic itself lives in SuiteSparse CHOLMOD, a
 * third-party dependency that is NOT vendored under /root/reference and is not installed here (Julia 1.10.5
 * bundles SuiteSparse 7.2.x — hpc/Singularity.def:17).  What is restated is therefore CHOLMOD's *published*
 * simplicial algorithm (Davis, "Direct Methods for Sparse Linear Systems", 2006; Liu 1990 for the
 * elimination tree): elimination tree by ancestor path compression, column counts by row-subtree
 * traversal, up-looking numeric factorisation, and the Takahashi recurrences (Erisman & Tinney 1975) for the
 * selected inverse.  PARITY UNPINNED: the reference ships no golden vectors for this path
 * (test/runtests.jl:5-10 is Aqua lint only); this oracle is pinned instead against dense LAPACK
 * (tests/test_oracle.py) — for SPD Q and a fixed permutation L is unique, so any correct FP64
 * implementation must agree to rounding.
 *
 * All indices are 0-based int64.  A is symmetric with BOTH triangles stored (CSC, sorted rows).
 * perm[k] = original index of the k-th pivot (CHOLMOD's new->old convention, `F.p`).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t idx;

/* Permuted upper-triangular pattern: for new column k the new rows i <= k, sorted. Returns malloc'ed CSC. */
static int permute_upper(idx n, const idx* Ap, const idx* Ai, const double* Ax, const idx* perm, idx** Cp_out,
                         idx** Ci_out, double** Cx_out) {
  idx* inv = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* Cp = (idx*)calloc((size_t)n + 1, sizeof(idx));
  if (!inv || !Cp) return -1;
  for (idx k = 0; k < n; k++) inv[perm[k]] = k;
  for (idx c = 0; c < n; c++)
    for (idx p = Ap[c]; p < Ap[c + 1]; p++) {
      idx i = inv[Ai[p]], j = inv[c];
      if (i <= j) Cp[j + 1]++;
    }
  for (idx k = 0; k < n; k++) Cp[k + 1] += Cp[k];
  idx nz = Cp[n];
  idx* Ci = (idx*)malloc(sizeof(idx) * (size_t)(nz > 0 ? nz : 1));
  double* Cx = (double*)malloc(sizeof(double) * (size_t)(nz > 0 ? nz : 1));
  idx* w = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  if (!Ci || !Cx || !w) return -1;
  memcpy(w, Cp, sizeof(idx) * (size_t)n);
  for (idx c = 0; c < n; c++)
    for (idx p = Ap[c]; p < Ap[c + 1]; p++) {
      idx i = inv[Ai[p]], j = inv[c];
      if (i <= j) {
        idx q = w[j]++;
        Ci[q] = i;
        Cx[q] = Ax ? Ax[p] : 0.0;
      }
    }
  /* sort rows within each column (insertion sort: columns are short) */
  for (idx j = 0; j < n; j++)
    for (idx p = Cp[j] + 1; p < Cp[j + 1]; p++) {
      idx r = Ci[p];
      double v = Cx[p];
      idx q = p - 1;
      while (q >= Cp[j] && Ci[q] > r) {
        Ci[q + 1] = Ci[q];
        Cx[q + 1] = Cx[q];
        q--;
      }
      Ci[q + 1] = r;
      Cx[q + 1] = v;
    }
  free(inv);
  free(w);
  *Cp_out = Cp;
  *Ci_out = Ci;
  *Cx_out = Cx;
  return 0;
}

/* Elimination tree of the permuted matrix from its upper-triangular columns (Liu 1990). */
static void etree_upper(idx n, const idx* Cp, const idx* Ci, idx* parent) {
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

/* Nonzero pattern of row k of L: the etree reach of the entries of column k of the upper triangle.
 * Writes the pattern into s[top..n-1] in topological order and returns top.  w is a mark array (w[i] == k marks). */
static idx ereach(idx n, const idx* Cp, const idx* Ci, idx k, const idx* parent, idx* s, idx* w) {
  idx top = n;
  w[k] = k;
  for (idx p = Cp[k]; p < Cp[k + 1]; p++) {
    idx i = Ci[p];
    if (i > k) continue;
    idx len = 0;
    for (; w[i] != k; i = parent[i]) {
      s[len++] = i;
      w[i] = k;
    }
    while (len > 0) s[--top] = s[--len];
  }
  return top;
}

/* Symbolic analysis with a given permutation: parent[n] (root = -1) and colcount[n] (incl. diagonal). */
int orc_symbolic(idx n, const idx* Ap, const idx* Ai, const idx* perm, idx* parent, idx* colcount) {
  idx *Cp, *Ci;
  double* Cx;
  if (permute_upper(n, Ap, Ai, NULL, perm, &Cp, &Ci, &Cx)) return -1;
  etree_upper(n, Cp, Ci, parent);
  idx* s = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* w = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  for (idx k = 0; k < n; k++) {
    w[k] = -1;
    colcount[k] = 1;
  }
  for (idx k = 0; k < n; k++) {
    idx top = ereach(n, Cp, Ci, k, parent, s, w);
    for (idx t = top; t < n; t++) colcount[s[t]]++; /* L(k, s[t]) is nonzero */
  }
  free(s);
  free(w);
  free(Cp);
  free(Ci);
  free(Cx);
  return 0;
}

/* Up-looking numeric Cholesky  P A P' = L L'.  Lp[n+1] must hold the column pointers (prefix sum of colcount);
 * Li/Lx sized Lp[n].  Rows come out sorted.  Returns 0, or k+1 if pivot k is not positive. */
int orc_cholesky(idx n, const idx* Ap, const idx* Ai, const double* Ax, const idx* perm, const idx* parent,
                 const idx* Lp, idx* Li, double* Lx) {
  idx *Cp, *Ci;
  double* Cx;
  if (permute_upper(n, Ap, Ai, Ax, perm, &Cp, &Ci, &Cx)) return -1;
  idx* s = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* w = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx* fillp = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  double* x = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  int rc = 0;
  for (idx k = 0; k < n; k++) {
    w[k] = -1;
    fillp[k] = Lp[k];
  }
  for (idx k = 0; k < n; k++) {
    idx top = ereach(n, Cp, Ci, k, parent, s, w);
    double d = 0.0;
    x[k] = 0.0;
    for (idx p = Cp[k]; p < Cp[k + 1]; p++) {
      if (Ci[p] < k)
        x[Ci[p]] = Cx[p];
      else if (Ci[p] == k)
        d = Cx[p];
    }
    /* solve L(0:k-1,0:k-1) * l = A(0:k-1,k) along the row pattern */
    for (idx t = top; t < n; t++) {
      idx i = s[t];
      double lki = x[i] / Lx[Lp[i]]; /* L(k,i) */
      x[i] = 0.0;
      for (idx p = Lp[i] + 1; p < fillp[i]; p++) x[Li[p]] -= Lx[p] * lki;
      d -= lki * lki;
      idx q = fillp[i]++;
      Li[q] = k;
      Lx[q] = lki;
    }
    if (!(d > 0.0)) {
      rc = (int)(k + 1);
      break;
    }
    idx q = fillp[k]++;
    Li[q] = k;
    Lx[q] = sqrt(d);
  }
  free(s);
  free(w);
  free(fillp);
  free(x);
  free(Cp);
  free(Ci);
  free(Cx);
  return rc;
}

/* x <- L^{-1} x (nrhs columns, ld n) */
void orc_lsolve(idx n, const idx* Lp, const idx* Li, const double* Lx, double* X, idx nrhs) {
  for (idx r = 0; r < nrhs; r++) {
    double* x = X + r * n;
    for (idx j = 0; j < n; j++) {
      x[j] /= Lx[Lp[j]];
      double xj = x[j];
      for (idx p = Lp[j] + 1; p < Lp[j + 1]; p++) x[Li[p]] -= Lx[p] * xj;
    }
  }
}

/* x <- L^{-T} x */
void orc_ltsolve(idx n, const idx* Lp, const idx* Li, const double* Lx, double* X, idx nrhs) {
  for (idx r = 0; r < nrhs; r++) {
    double* x = X + r * n;
    for (idx j = n - 1; j >= 0; j--) {
      double xj = x[j];
      for (idx p = Lp[j] + 1; p < Lp[j + 1]; p++) xj -= Lx[p] * x[Li[p]];
      x[j] = xj / Lx[Lp[j]];
    }
  }
}

static idx find_row(const idx* Li, idx lo, idx hi, idx r) {
  while (lo < hi) {
    idx mid = (lo + hi) >> 1;
    if (Li[mid] < r)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* Takahashi selected inverse on the pattern of L: Zx[p] = (A^{-1})_{perm}(Li[p], j) for p in column j.
 *   Z_ij = ( delta_ij / L_jj - sum_{k>j, L_kj != 0} L_kj Z_{ik} ) / L_jj,   columns j = n-1 .. 0. */
int orc_selinv(idx n, const idx* Lp, const idx* Li, const double* Lx, double* Zx) {
  for (idx j = n - 1; j >= 0; j--) {
    double ljj = Lx[Lp[j]];
    for (idx pi = Lp[j + 1] - 1; pi >= Lp[j]; pi--) {
      idx i = Li[pi];
      double acc = (i == j) ? 1.0 / ljj : 0.0;
      for (idx pk = Lp[j] + 1; pk < Lp[j + 1]; pk++) {
        idx k = Li[pk];
        idx a = i > k ? i : k, b = i > k ? k : i; /* Z(a,b), a >= b > j: stored in column b */
        idx q = find_row(Li, Lp[b], Lp[b + 1], a);
        if (q >= Lp[b + 1] || Li[q] != a) return -2; /* cannot happen for a Cholesky pattern */
        acc -= Lx[pk] * Zx[q];
      }
      Zx[pi] = acc / ljj;
    }
  }
  return 0;
}
