/* kalman_c.c - plain-C port of the reference StandardFilter log-likelihood and its reverse-mode
 * gradient (ORACLE / CPU BASELINE - test infrastructure, never linked into the product).
 *
 * Forward: reference pymc_statespace/filters/kalman_filter.py:231-284 (mask -> update -> predict),
 * one call per (draw); gradient: the adjoint recursion of SURVEY.md appendix B (what PyTensor's
 * Scan.L_op computes for the same graph).  OpenMP over draws = "all host cores".
 * Straightforward loops on small stack matrices; this is what a careful CPU implementation of the
 * reference path looks like, and it is what bench.py times as the CPU baseline.
 *
 *   gcc -O3 -march=native -fopenmp -shared -fPIC oracle/kalman_c.c -o oracle/_build/libkalman_c.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXM 32
#define MAXP 8
#define LOG_2PI 1.8378770664093454835606594728112

static inline __attribute__((always_inline)) void matmul(double *C, const double *A, const double *B, int r, int k, int c, int ta, int tb) {
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += (ta ? A[l * r + i] : A[i * k + l]) * (tb ? B[j * k + l] : B[l * c + j]);
      C[i * c + j] = s;
    }
}

/* inverse + log-determinant of a symmetric positive definite p x p matrix (Cholesky) */
static int spd_inverse(const double *F, double *G, double *logdet, int p) {
  double L[MAXP * MAXP], Li[MAXP * MAXP];
  memset(L, 0, sizeof(L));
  *logdet = 0.0;
  for (int j = 0; j < p; ++j) {
    double d = F[j * p + j];
    for (int k = 0; k < j; ++k) d -= L[j * p + k] * L[j * p + k];
    if (!(d > 0.0)) return 1;
    L[j * p + j] = sqrt(d);
    *logdet += log(d);
    for (int i = j + 1; i < p; ++i) {
      double s = F[i * p + j];
      for (int k = 0; k < j; ++k) s -= L[i * p + k] * L[j * p + k];
      L[i * p + j] = s / L[j * p + j];
    }
  }
  memset(Li, 0, sizeof(Li));
  for (int c = 0; c < p; ++c)
    for (int i = c; i < p; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= L[i * p + k] * Li[k * p + c];
      Li[i * p + c] = s / L[i * p + i];
    }
  matmul(G, Li, Li, p, p, p, 1, 0);
  return 0;
}

typedef struct {
  double v[MAXP], M[MAXM * MAXP], F[MAXP * MAXP], G[MAXP * MAXP], K[MAXM * MAXP], A[MAXM * MAXM], w[MAXP];
  double af[MAXM], Pf[MAXM * MAXM], logdet, quad;
} upd_t;

static int update_step(int m, int p, const double *y, const double *a, const double *P, const double *Z, const double *H,
                       const double *d, upd_t *u) {
  double t1[MAXM * MAXM], t2[MAXM * MAXM];
  for (int i = 0; i < p; ++i) {
    double s = y[i] - (d ? d[i] : 0.0);
    for (int k = 0; k < m; ++k) s -= Z[i * m + k] * a[k];
    u->v[i] = s;
  }
  matmul(u->M, P, Z, m, m, p, 0, 1);
  matmul(u->F, Z, u->M, p, m, p, 0, 0);
  for (int i = 0; i < p * p; ++i) u->F[i] += H[i];
  if (spd_inverse(u->F, u->G, &u->logdet, p)) return 1;
  matmul(u->K, u->M, u->G, m, p, p, 0, 0);
  matmul(u->w, u->G, u->v, p, p, 1, 0, 0);
  u->quad = 0.0;
  for (int i = 0; i < p; ++i) u->quad += u->v[i] * u->w[i];
  matmul(u->A, u->K, Z, m, p, m, 0, 0);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) u->A[i * m + j] = (i == j ? 1.0 : 0.0) - u->A[i * m + j];
  for (int i = 0; i < m; ++i) {
    double s = a[i];
    for (int k = 0; k < p; ++k) s += u->K[i * p + k] * u->v[k];
    u->af[i] = s;
  }
  matmul(t1, u->A, P, m, m, m, 0, 0);
  matmul(u->Pf, t1, u->A, m, m, m, 0, 1);
  matmul(t1, u->K, H, m, p, p, 0, 0);
  matmul(t2, t1, u->K, m, p, m, 0, 1);
  for (int i = 0; i < m * m; ++i) u->Pf[i] += t2[i];
  return 0;
}

/* One draw.  y[n*p] (NaN rows = missing, all-or-nothing), C = R Q R^T.  tape: n*(m+m*m) doubles scratch.
 * Outputs: *ll; if grads != 0: ga0[m] gP0[m*m] gT[m*m] gZ[p*m] gH[p*p] gC[m*m] gc[m] gd[p].  Returns info. */
static int kalman_one(int n, int m, int p, const double *y, const double *a0, const double *P0, const double *T,
                      const double *Z, const double *H, const double *C, const double *c, const double *d,
                      double ll_const, double *tape, double *ll_out, int grads, double *ga0, double *gP0, double *gT,
                      double *gZ, double *gH, double *gC, double *gc, double *gd) {
  double a[MAXM], P[MAXM * MAXM], t1[MAXM * MAXM], t2[MAXM * MAXM];
  upd_t u;
  const int mm = m * m;
  memcpy(a, a0, sizeof(double) * m);
  memcpy(P, P0, sizeof(double) * mm);
  double ll = 0.0;
  for (int t = 0; t < n; ++t) {
    memcpy(tape + (size_t)t * (m + mm), a, sizeof(double) * m);
    memcpy(tape + (size_t)t * (m + mm) + m, P, sizeof(double) * mm);
    const double *yt = y + (size_t)t * p;
    int nmiss = 0;
    for (int i = 0; i < p; ++i) nmiss += isnan(yt[i]) ? 1 : 0;
    if (nmiss == 0) {
      if (update_step(m, p, yt, a, P, Z, H, d, &u)) { *ll_out = NAN; return t + 1; }
      ll += -0.5 * (ll_const + u.logdet + u.quad);
    } else {
      if (nmiss != p) { *ll_out = NAN; return -(t + 1); }
      memcpy(u.af, a, sizeof(double) * m);
      memcpy(u.Pf, P, sizeof(double) * mm);
    }
    for (int i = 0; i < m; ++i) {
      double s = c ? c[i] : 0.0;
      for (int k = 0; k < m; ++k) s += T[i * m + k] * u.af[k];
      a[i] = s;
    }
    matmul(t1, T, u.Pf, m, m, m, 0, 0);
    matmul(t2, t1, T, m, m, m, 0, 1);
    for (int i = 0; i < mm; ++i) t2[i] += C[i];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) P[i * m + j] = 0.5 * (t2[i * m + j] + t2[j * m + i]);
  }
  *ll_out = ll;
  if (!grads) return 0;

  double ab[MAXM], Pb[MAXM * MAXM], afb[MAXM], Pfb[MAXM * MAXM], Ps[MAXM * MAXM], Ab[MAXM * MAXM];
  double Kb[MAXM * MAXP], Mb[MAXM * MAXP], Fb[MAXP * MAXP], vb[MAXP], t3[MAXM * MAXM], q1[MAXP * MAXP];
  memset(ab, 0, sizeof(ab)); memset(Pb, 0, sizeof(Pb));
  memset(gT, 0, sizeof(double) * mm); memset(gZ, 0, sizeof(double) * p * m); memset(gH, 0, sizeof(double) * p * p);
  memset(gC, 0, sizeof(double) * mm); memset(gc, 0, sizeof(double) * m); memset(gd, 0, sizeof(double) * p);
  for (int t = n - 1; t >= 0; --t) {
    const double *at = tape + (size_t)t * (m + mm), *Pt = at + m;
    const double *yt = y + (size_t)t * p;
    int observed = 1;
    for (int i = 0; i < p; ++i) if (isnan(yt[i])) observed = 0;
    if (observed) update_step(m, p, yt, at, Pt, Z, H, d, &u);
    else { memcpy(u.af, at, sizeof(double) * m); memcpy(u.Pf, Pt, sizeof(double) * mm); }
    /* predict adjoint */
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Ps[i * m + j] = 0.5 * (Pb[i * m + j] + Pb[j * m + i]);
    for (int i = 0; i < mm; ++i) gC[i] += Ps[i];
    for (int i = 0; i < m; ++i) gc[i] += ab[i];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) t1[i * m + j] = u.Pf[i * m + j] + u.Pf[j * m + i];
    matmul(t2, T, t1, m, m, m, 0, 0);
    matmul(t3, Ps, t2, m, m, m, 0, 0);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) gT[i * m + j] += t3[i * m + j] + ab[i] * u.af[j];
    matmul(afb, T, ab, m, m, 1, 1, 0);
    matmul(t2, Ps, T, m, m, m, 0, 0);
    matmul(Pfb, T, t2, m, m, m, 1, 0);
    if (!observed) { memcpy(ab, afb, sizeof(double) * m); memcpy(Pb, Pfb, sizeof(double) * mm); continue; }
    /* update adjoint */
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) t1[i * m + j] = Pt[i * m + j] + Pt[j * m + i];
    matmul(t2, u.A, t1, m, m, m, 0, 0);
    matmul(Ab, Pfb, t2, m, m, m, 0, 0);
    matmul(t2, Pfb, u.A, m, m, m, 0, 0);
    matmul(Pb, u.A, t2, m, m, m, 1, 0);
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) q1[i * p + j] = H[i * p + j] + H[j * p + i];
    matmul(t1, u.K, q1, m, p, p, 0, 0);
    matmul(Kb, Pfb, t1, m, m, p, 0, 0);
    matmul(t1, Ab, Z, m, m, p, 0, 1);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < p; ++j) Kb[i * p + j] += afb[i] * u.v[j] - t1[i * p + j];
    matmul(t1, Pfb, u.K, m, m, p, 0, 0);
    matmul(t2, u.K, t1, p, m, p, 1, 0);
    for (int i = 0; i < p * p; ++i) gH[i] += t2[i];
    matmul(t2, u.K, Ab, p, m, m, 1, 0);
    for (int i = 0; i < p * m; ++i) gZ[i] -= t2[i];
    matmul(vb, u.K, afb, p, m, 1, 1, 0);
    for (int i = 0; i < p; ++i) vb[i] -= u.w[i];
    matmul(t1, u.K, Kb, p, m, p, 1, 0);
    matmul(t2, t1, u.G, p, p, p, 0, 1);
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) Fb[i * p + j] = -0.5 * (u.G[j * p + i] - u.w[i] * u.w[j]) - t2[i * p + j];
    matmul(Mb, Kb, u.G, m, p, p, 0, 1);
    matmul(t1, Z, Fb, m, p, p, 1, 0);
    for (int i = 0; i < m * p; ++i) Mb[i] += t1[i];
    matmul(t1, Fb, u.M, p, p, m, 0, 1);
    matmul(t2, Mb, Pt, p, m, m, 1, 0);
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < m; ++j) gZ[i * m + j] += t1[i * m + j] + t2[i * m + j] - vb[i] * at[j];
    for (int i = 0; i < p * p; ++i) gH[i] += Fb[i];
    matmul(t1, Mb, Z, m, p, m, 0, 0);
    for (int i = 0; i < mm; ++i) Pb[i] += t1[i];
    matmul(t1, Z, vb, m, p, 1, 1, 0);
    for (int i = 0; i < m; ++i) ab[i] = afb[i] - t1[i];
    for (int i = 0; i < p; ++i) gd[i] -= vb[i];
  }
  memcpy(ga0, ab, sizeof(double) * m);
  memcpy(gP0, Pb, sizeof(double) * mm);
  return 0;
}

/* Batched over draws with OpenMP.  Per-draw arrays are dense [B, ...]; y, Z, H shared.  C = R Q R^T per draw.
 * grads layout per draw: [a0 m | P0 mm | T mm | Z pm | H pp | C mm | c m | d p].  Returns #draws with info != 0. */
int kalman_c_batch(long B, int n, int m, int p, const double *y, const double *a0, const double *P0, const double *T,
                   const double *Z, const double *H, const double *C, double ll_const, int want_grads, double *ll,
                   double *grads, int nthreads) {
  if (m > MAXM || p > MAXP) return -1;
  const int mm = m * m;
  const int gsz = m + mm + mm + p * m + p * p + mm + m + p;
  int bad = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : bad)
  {
    double *tape = (double *)malloc(sizeof(double) * (size_t)n * (m + mm));
#pragma omp for schedule(static)
    for (long b = 0; b < B; ++b) {
      double *g = grads ? grads + (size_t)b * gsz : 0;
      int info = kalman_one(n, m, p, y, a0 + b * m, P0 + b * mm, T + b * mm, Z, H, C + b * mm, 0, 0, ll_const, tape,
                            ll + b, want_grads, g, g + m, g + m + mm, g + m + 2 * mm, g + m + 2 * mm + p * m,
                            g + m + 2 * mm + p * m + p * p, g + m + 3 * mm + p * m + p * p,
                            g + 2 * m + 3 * mm + p * m + p * p);
      bad += (info != 0);
    }
    free(tape);
  }
  return bad;
}

int kalman_c_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
