/*
 * oracle/supernodal_chol.c — TEST / BASELINE INFRASTRUCTURE ONLY (never linked into libgmrfb).
 *
 * CPU supernodal multifrontal Cholesky, triangular solves and Takahashi selected inversion with BLAS-3 on the
 * supernodes — the algorithm class CHOLMOD's supernodal path uses (dense potrf/trsm/syrk per supernode), which is
 * what the reference reaches through `cholesky(Symmetric(A); perm=p)` (scripts/solve_burger.jl:147,
 * scripts/darcy/solve_darcy_gmrf-fem.jl:188).  It is the *timed CPU baseline* of bench.py ("port") and a second,
 * independent numeric check of the scalar oracle in sparse_chol.c.  BLAS/LAPACK are reached through function
 * pointers handed in from Python (scipy.linalg.cython_blas / cython_lapack -> OpenBLAS, all host threads).
 *
 * Input structure (0-based, "internal" postordered numbering in which supernodes are contiguous column ranges and
 * children precede parents): sptr[ns+1], rptr[ns+1], rows[] (per supernode: its own columns first, then the
 * below-diagonal rows ascending), sparent[ns] (-1 for roots), and the permuted matrix's lower triangle as CSC.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t idx;
typedef int bint; /* LP64 BLAS integer */

typedef void (*dgemm_t)(char*, char*, bint*, bint*, bint*, double*, double*, bint*, double*, bint*, double*, double*, bint*);
typedef void (*dsyrk_t)(char*, char*, bint*, bint*, double*, double*, bint*, double*, double*, bint*);
typedef void (*dtrsm_t)(char*, char*, char*, char*, bint*, bint*, double*, double*, bint*, double*, bint*);
typedef void (*dtrsv_t)(char*, char*, char*, bint*, double*, bint*, double*, bint*);
typedef void (*dgemv_t)(char*, bint*, bint*, double*, double*, bint*, double*, bint*, double*, double*, bint*);
typedef void (*dsymm_t)(char*, char*, bint*, bint*, double*, double*, bint*, double*, bint*, double*, double*, bint*);
typedef void (*dpotrf_t)(char*, bint*, double*, bint*, bint*);
typedef void (*dtrtri_t)(char*, char*, bint*, double*, bint*, bint*);
typedef void (*dlauum_t)(char*, bint*, double*, bint*, bint*);

static struct {
  dgemm_t dgemm;
  dsyrk_t dsyrk;
  dtrsm_t dtrsm;
  dtrsv_t dtrsv;
  dgemv_t dgemv;
  dsymm_t dsymm;
  dpotrf_t dpotrf;
  dtrtri_t dtrtri;
  dlauum_t dlauum;
} B;

void sn_set_blas(void* dgemm, void* dsyrk, void* dtrsm, void* dtrsv, void* dgemv, void* dsymm, void* dpotrf, void* dtrtri,
                 void* dlauum) {
  B.dgemm = (dgemm_t)dgemm;
  B.dsyrk = (dsyrk_t)dsyrk;
  B.dtrsm = (dtrsm_t)dtrsm;
  B.dtrsv = (dtrsv_t)dtrsv;
  B.dgemv = (dgemv_t)dgemv;
  B.dsymm = (dsymm_t)dsymm;
  B.dpotrf = (dpotrf_t)dpotrf;
  B.dtrtri = (dtrtri_t)dtrtri;
  B.dlauum = (dlauum_t)dlauum;
}

/* Numeric factorisation.  Lx: factor panels, panel s at loff[s], d_s x s_s column-major (ld = d_s), where the
 * s_s x s_s top block holds L11 (lower) and the rest L21.  Returns 0, or 1 + failing internal column. */
idx sn_factor(idx n, idx ns, const idx* sptr, const idx* rptr, const idx* rows, const idx* sparent, const idx* Cp,
              const idx* Ci, const double* Cx, const idx* loff, double* Lx) {
  idx* map = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  double** upd = (double**)calloc((size_t)(ns > 0 ? ns : 1), sizeof(double*));
  idx* child_head = (idx*)malloc(sizeof(idx) * (size_t)(ns > 0 ? ns : 1));
  idx* child_next = (idx*)malloc(sizeof(idx) * (size_t)(ns > 0 ? ns : 1));
  idx maxd = 0, fail = 0;
  if (!map || !upd || !child_head || !child_next) return -1;
  for (idx s = 0; s < ns; s++) {
    child_head[s] = -1;
    idx d = rptr[s + 1] - rptr[s];
    if (d > maxd) maxd = d;
  }
  for (idx s = ns - 1; s >= 0; s--) /* descending insert => ascending child lists */
    if (sparent[s] >= 0) {
      child_next[s] = child_head[sparent[s]];
      child_head[sparent[s]] = s;
    }
  double* F = (double*)malloc(sizeof(double) * (size_t)(maxd * maxd > 0 ? maxd * maxd : 1));
  if (!F) return -1;
  char L_ = 'L', R_ = 'R', T_ = 'T', N_ = 'N';
  double one = 1.0, mone = -1.0;
  for (idx s = 0; s < ns && !fail; s++) {
    const idx c0 = sptr[s], sc = sptr[s + 1] - sptr[s];
    const idx d = rptr[s + 1] - rptr[s], r = d - sc;
    const idx* rw = rows + rptr[s];
    memset(F, 0, sizeof(double) * (size_t)(d * d));
    for (idx i = 0; i < d; i++) map[rw[i]] = i;
    for (idx j = 0; j < sc; j++)
      for (idx p = Cp[c0 + j]; p < Cp[c0 + j + 1]; p++) F[map[Ci[p]] + j * d] = Cx[p];
    for (idx c = child_head[s]; c >= 0; c = child_next[c]) {
      const idx scc = sptr[c + 1] - sptr[c], dc = rptr[c + 1] - rptr[c], rc = dc - scc;
      const idx* crw = rows + rptr[c] + scc;
      const double* U = upd[c];
      for (idx j = 0; j < rc; j++) {
        const idx pj = map[crw[j]];
        for (idx i = j; i < rc; i++) F[map[crw[i]] + pj * d] += U[i + j * rc];
      }
      free(upd[c]);
      upd[c] = NULL;
    }
    bint bd = (bint)d, bs = (bint)sc, br = (bint)r, info = 0;
    B.dpotrf(&L_, &bs, F, &bd, &info);
    if (info != 0) {
      fail = 1 + c0 + (info > 0 ? info - 1 : 0);
      break;
    }
    if (r > 0) {
      B.dtrsm(&R_, &L_, &T_, &N_, &br, &bs, &one, F, &bd, F + sc, &bd);
      B.dsyrk(&L_, &N_, &br, &bs, &mone, F + sc, &bd, &one, F + sc + sc * d, &bd);
      double* U = (double*)malloc(sizeof(double) * (size_t)(r * r));
      if (!U) {
        fail = -1;
        break;
      }
      for (idx j = 0; j < r; j++) memcpy(U + j * r + j, F + (sc + j) * d + sc + j, sizeof(double) * (size_t)(r - j));
      upd[s] = U;
    }
    memcpy(Lx + loff[s], F, sizeof(double) * (size_t)(d * sc));
  }
  for (idx s = 0; s < ns; s++) free(upd[s]);
  free(F);
  free(map);
  free(upd);
  free(child_head);
  free(child_next);
  return fail;
}

/* x (n x nrhs, column-major, internal ordering) <- L^{-1} x (fwd != 0) and/or L^{-T} x (bwd != 0). */
void sn_solve(idx n, idx ns, const idx* sptr, const idx* rptr, const idx* rows, const idx* loff, const double* Lx,
              double* X, idx nrhs, int fwd, int bwd) {
  idx maxr = 0;
  for (idx s = 0; s < ns; s++) {
    idx r = (rptr[s + 1] - rptr[s]) - (sptr[s + 1] - sptr[s]);
    if (r > maxr) maxr = r;
  }
  double* t = (double*)malloc(sizeof(double) * (size_t)(maxr > 0 ? maxr : 1));
  char L_ = 'L', T_ = 'T', N_ = 'N';
  double one = 1.0, mone = -1.0, zero = 0.0;
  bint i1 = 1;
  for (idx q = 0; q < nrhs; q++) {
    double* x = X + q * n;
    if (fwd)
      for (idx s = 0; s < ns; s++) {
        const idx c0 = sptr[s], sc = sptr[s + 1] - sptr[s], d = rptr[s + 1] - rptr[s], r = d - sc;
        const idx* rw = rows + rptr[s] + sc;
        const double* P = Lx + loff[s];
        bint bd = (bint)d, bs = (bint)sc, br = (bint)r;
        B.dtrsv(&L_, &N_, &N_, &bs, (double*)P, &bd, x + c0, &i1);
        if (r > 0) {
          B.dgemv(&N_, &br, &bs, &one, (double*)P + sc, &bd, x + c0, &i1, &zero, t, &i1);
          for (idx i = 0; i < r; i++) x[rw[i]] -= t[i];
        }
      }
    if (bwd)
      for (idx s = ns - 1; s >= 0; s--) {
        const idx c0 = sptr[s], sc = sptr[s + 1] - sptr[s], d = rptr[s + 1] - rptr[s], r = d - sc;
        const idx* rw = rows + rptr[s] + sc;
        const double* P = Lx + loff[s];
        bint bd = (bint)d, bs = (bint)sc, br = (bint)r;
        if (r > 0) {
          for (idx i = 0; i < r; i++) t[i] = x[rw[i]];
          B.dgemv(&T_, &br, &bs, &mone, (double*)P + sc, &bd, t, &i1, &one, x + c0, &i1);
        }
        B.dtrsv(&L_, &T_, &N_, &bs, (double*)P, &bd, x + c0, &i1);
      }
  }
  free(t);
}

/* Takahashi selected inversion, supernodal, top-down.  zdiag[k] = (A^{-1})_kk in the internal ordering.
 * Each supernode's dense inverse front Z (d x d, symmetric, full) is kept until its last child has gathered
 * Z_RR from it.   Z_RC = -Z_RR Y,  Z_CC = W'W - Y' Z_RC  with  W = L11^{-1},  Y = L21 W. */
idx sn_selinv(idx n, idx ns, const idx* sptr, const idx* rptr, const idx* rows, const idx* sparent, const idx* loff,
              const double* Lx, double* zdiag) {
  double** Zf = (double**)calloc((size_t)(ns > 0 ? ns : 1), sizeof(double*));
  idx* nchild = (idx*)calloc((size_t)(ns > 0 ? ns : 1), sizeof(idx));
  idx* map = (idx*)malloc(sizeof(idx) * (size_t)(n > 0 ? n : 1));
  idx maxd = 0;
  if (!Zf || !nchild || !map) return -1;
  for (idx s = 0; s < ns; s++) {
    if (sparent[s] >= 0) nchild[sparent[s]]++;
    idx d = rptr[s + 1] - rptr[s];
    if (d > maxd) maxd = d;
  }
  double* W = (double*)malloc(sizeof(double) * (size_t)(maxd * maxd > 0 ? maxd * maxd : 1));
  double* Y = (double*)malloc(sizeof(double) * (size_t)(maxd * maxd > 0 ? maxd * maxd : 1));
  if (!W || !Y) return -1;
  char L_ = 'L', R_ = 'R', T_ = 'T', N_ = 'N';
  double one = 1.0, mone = -1.0, zero = 0.0;
  idx rc = 0;
  for (idx s = ns - 1; s >= 0; s--) {
    const idx c0 = sptr[s], sc = sptr[s + 1] - sptr[s], d = rptr[s + 1] - rptr[s], r = d - sc;
    const idx* rw = rows + rptr[s];
    const double* P = Lx + loff[s];
    double* Z = (double*)malloc(sizeof(double) * (size_t)(d * d));
    if (!Z) {
      rc = -1;
      break;
    }
    bint bd = (bint)d, bs = (bint)sc, br = (bint)r, info = 0;
    /* W = L11^{-1} (lower) */
    for (idx j = 0; j < sc; j++) {
      for (idx i = 0; i < j; i++) W[i + j * sc] = 0.0;
      for (idx i = j; i < sc; i++) W[i + j * sc] = P[i + j * d];
    }
    B.dtrtri(&L_, &N_, &bs, W, &bs, &info);
    if (r > 0) {
      const idx p = sparent[s];
      const idx dp = rptr[p + 1] - rptr[p];
      const idx* prw = rows + rptr[p];
      const double* Zp = Zf[p];
      for (idx i = 0; i < dp; i++) map[prw[i]] = i;
      /* Z_RR gather (full symmetric) into Z[sc:, sc:] */
      for (idx j = 0; j < r; j++) {
        const idx pj = map[rw[sc + j]];
        for (idx i = 0; i < r; i++) Z[(sc + i) + (sc + j) * d] = Zp[map[rw[sc + i]] + pj * dp];
      }
      if (--nchild[p] == 0) {
        free(Zf[p]);
        Zf[p] = NULL;
      }
      /* Y = L21 W  (r x sc) */
      for (idx j = 0; j < sc; j++) memcpy(Y + j * r, P + sc + j * d, sizeof(double) * (size_t)r);
      B.dtrsm(&R_, &L_, &N_, &N_, &br, &bs, &one, (double*)P, &bd, Y, &br);
      /* Z_RC = -Z_RR Y */
      B.dsymm(&L_, &L_, &br, &bs, &mone, Z + sc + sc * d, &bd, Y, &br, &zero, Z + sc, &bd);
    }
    /* Z_CC = W'W (lower, in W) */
    B.dlauum(&L_, &bs, W, &bs, &info);
    for (idx j = 0; j < sc; j++) {
      for (idx i = 0; i < j; i++) Z[i + j * d] = 0.0;
      for (idx i = j; i < sc; i++) Z[i + j * d] = W[i + j * sc];
    }
    if (r > 0) {
      /* Z_CC -= Y' Z_RC : full s x s product, only the lower part is kept */
      B.dgemm(&T_, &N_, &bs, &bs, &br, &mone, Y, &br, Z + sc, &bd, &one, Z, &bd);
    }
    /* symmetrise: upper of Z_CC and Z_CR */
    for (idx j = 0; j < sc; j++) {
      for (idx i = j + 1; i < sc; i++) Z[j + i * d] = Z[i + j * d];
      for (idx i = sc; i < d; i++) Z[j + i * d] = Z[i + j * d];
      zdiag[c0 + j] = Z[j + j * d];
    }
    if (nchild[s] > 0)
      Zf[s] = Z;
    else
      free(Z);
  }
  for (idx s = 0; s < ns; s++) free(Zf[s]);
  free(Zf);
  free(nchild);
  free(map);
  free(W);
  free(Y);
  return rc;
}
