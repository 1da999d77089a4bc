// kf_dare.cuh - steady-state covariance for filter_type="steady_state".
//
// The reference calls scipy.linalg.solve_discrete_are(T^T, Z^T, R Q R^T, H) once per logp evaluation
// (reference pymc_statespace/filters/kalman_filter.py:384, utils/pytensor_scipy.py:28-35) and differentiates it
// with the Kao & Hennequin adjoint (utils/pytensor_scipy.py:39-60).  SciPy's QZ solver has no batched GPU
// analogue, so the published fixed point is computed here by
//   (1) a short run of the Riccati recursion from P = alpha*I (every iterate >= P_ss, hence every gain is
//       stabilising), which needs only F = Z P Z^T + H to be invertible (H = 0 is fine: BayesianARMA), then
//   (2) Newton-Hewer iterations  P <- Lyapunov(T(I-KZ), C + T K H K^T T^T)  (quadratic convergence), each
//       Lyapunov equation solved by squared-Smith doubling.
// Written over the same execution-context policy as kf_core.cuh.
#pragma once
#include "kf_core.cuh"

namespace kfb {

// X <- sum_k Ak^k X Ak^kT (Ak destroyed).  Returns false if it does not converge (spectral radius >= 1).
template <class X_, class TM>
KFB_HD bool smith_doubling_ctx(X_& x, TM& Ak, TM& X, TM& S1, int m) {
  for (int it = 0; it < 64; ++it) {
    double mx = 0.0;
    KFB_FOR(i, m * m) mx = fmax(mx, fabs(Ak[i]));
    mx = x.reduce_max(mx);
    if (!(mx < 1.0e150)) return false;
    if (mx < 1.0e-11) return true;
    gemm<false, false, 0>(x, S1, Ak, X, m, m, m);
    gemm<false, true, 1>(x, X, S1, Ak, m, m, m);
    gemm<false, false, 0>(x, S1, Ak, Ak, m, m, m);
    KFB_FOR(i, m * m) Ak[i] = S1[i];
    x.sync();
  }
  return false;
}

// Solves P = T P T^T - T P Z^T (Z P Z^T + H)^-1 Z P T^T + C.  Outputs Pss[m*m], Gss[p*p] = (Z Pss Z^T + H)^-1
// (global memory).  Returns 0 ok, 1 = not converged / F not positive definite.
template <class X>
KFB_HD int dare_unit(X& x, const double* Tg, const double* Zg, const double* Hg, const double* Cg, double* Pss,
                     double* Gss) {
  const int m = x.m(), p = x.p();
  Params<X> prm(x);
  typename X::template Buf<SZ_MM> C(x), P(x), Pf(x), Pn(x), Ak(x), Rhs(x);
  typename X::template Buf<SZ_M> a(x), af(x), c0(x);
  typename X::template Buf<SZ_P> y0(x);
  UpdTmp<X> tmp(x);
  load_or_zero(x, prm.T, Tg, m * m);
  load_or_zero(x, prm.Z, Zg, p * m);
  load_or_zero(x, prm.H, Hg, p * p);
  load_or_zero(x, C, Cg, m * m);
  KFB_FOR(i, p) { prm.d[i] = 0.0; y0[i] = 0.0; }
  KFB_FOR(i, m) { a[i] = 0.0; c0[i] = 0.0; }
  double scale = 1.0;
  KFB_FOR(i, m * m) scale = fmax(scale, fabs(C[i]));
  KFB_FOR(i, p * p) scale = fmax(scale, fabs(prm.H[i]));
  x.sync();
  scale = x.reduce_max(scale);
  KFB_FOR(idx, m * m) P[idx] = (x.div_m(idx) * (m + 1) == idx) ? 1.0e4 * scale : 0.0;
  x.sync();
  int info = 0;
  const double* yp = &y0[0];
  // (1) Riccati recursion from above
  const int n0 = 2 * m + 24;
  for (int k = 0; k < n0; ++k) {
    StepStat st = update_observed<MK_STD>(x, prm, yp, 0.0, a, P, tmp, af, Pf, (LogAcc*)nullptr, false);
    if (!x.all_ok(st.ok)) info = 1;
    predict(x, prm.T, C, c0, af, Pf, a, P, tmp.S1, tmp.S2);
  }
  // (2) Newton-Hewer
  bool converged = false;
  for (int it = 0; it < 60 && !converged && info == 0; ++it) {
    StepStat st = update_observed<MK_STD>(x, prm, yp, 0.0, a, P, tmp, af, Pf, (LogAcc*)nullptr, false);
    if (!x.all_ok(st.ok)) { info = 1; break; }
    // Ak = T (I - K Z) ; Rhs = C + T (K H K^T) T^T
    gemm<false, false, 0>(x, Ak, prm.T, tmp.A, m, m, m);
    gemm<false, true, 0>(x, tmp.S1, tmp.KH, tmp.K, m, p, m);  // K H K^T   (tmp.KH = K H from update_observed)
    gemm<false, false, 0>(x, tmp.S2, prm.T, tmp.S1, m, m, m);
    KFB_FOR(i, m * m) Rhs[i] = C[i];
    x.sync();
    gemm<false, true, 1>(x, Rhs, tmp.S2, prm.T, m, m, m);
    KFB_FOR(idx, m * m) {  // symmetrise: keeps the iteration on the symmetric manifold
      const int i = x.div_m(idx), j = idx - i * m;
      Pn[idx] = 0.5 * (Rhs[idx] + Rhs[j * m + i]);
    }
    x.sync();
    if (!smith_doubling_ctx(x, Ak, Pn, tmp.S1, m)) { info = 1; break; }
    double diff = 0.0, mag = 0.0;
    KFB_FOR(i, m * m) {
      diff = fmax(diff, fabs(Pn[i] - P[i]));
      mag = fmax(mag, fabs(Pn[i]));
    }
    diff = x.reduce_max(diff);
    mag = x.reduce_max(mag);
    KFB_FOR(idx, m * m) {
      const int i = x.div_m(idx), j = idx - i * m;
      P[idx] = 0.5 * (Pn[idx] + Pn[j * m + i]);
    }
    x.sync();
    converged = diff <= 4.0e-15 * mag;
  }
  if (!converged) info = 1;
  // Gss = (Z Pss Z^T + H)^-1
  StepStat st = update_observed<MK_STD>(x, prm, yp, 0.0, a, P, tmp, af, Pf, (LogAcc*)nullptr, false);
  if (!x.all_ok(st.ok)) info = 1;
  KFB_FOR(i, m * m) Pss[i] = info ? nan("") : P[i];
  KFB_FOR(i, p * p) Gss[i] = info ? nan("") : tmp.Fi[i];
  return info;
}

// Adjoint of (Pss, Gss) wrt (T, Z, H, C): reference utils/pytensor_scipy.py:39-60 with A = T^T, B = Z^T,
// Q = C, R = H, plus the chain through Gss = (Z Pss Z^T + H)^-1.  Accumulates (+=) into gT, gZ, gH, gC.
template <class X>
KFB_HD void dare_adjoint_unit(X& x, const double* Tg, const double* Zg, const double* Hg, const double* Pssg,
                              const double* Gssg, const double* gPss, const double* gGss, double* gT, double* gZ,
                              double* gH, double* gC) {
  const int m = x.m(), p = x.p();
  Params<X> prm(x);
  typename X::template Buf<SZ_MM> P(x), Pf(x), Xb(x), Ak(x), S(x), W(x);
  typename X::template Buf<SZ_M> a(x), af(x);
  typename X::template Buf<SZ_P> y0(x);
  typename X::template Buf<SZ_PP> Fb(x), Q1(x);
  typename X::template Buf<SZ_MP> Kp(x), ZP(x);
  UpdTmp<X> tmp(x);
  load_or_zero(x, prm.T, Tg, m * m);
  load_or_zero(x, prm.Z, Zg, p * m);
  load_or_zero(x, prm.H, Hg, p * p);
  load_or_zero(x, P, Pssg, m * m);
  load_or_zero(x, prm.Gss, Gssg, p * p);
  load_or_zero(x, Xb, gPss, m * m);
  load_or_zero(x, Q1, gGss, p * p);
  KFB_FOR(i, p) { prm.d[i] = 0.0; y0[i] = 0.0; }
  KFB_FOR(i, m) a[i] = 0.0;
  x.sync();
  // Fb = -Gss^T Gssb Gss^T
  gemm<true, false, 0>(x, tmp.F, prm.Gss, Q1, p, p, p);
  gemm<false, true, 0>(x, Fb, tmp.F, prm.Gss, p, p, p);
  KFB_FOR(i, p * p) Fb[i] = -Fb[i];
  x.sync();
  // Xb += Z^T Fb Z ; gZ += Fb Z P^T + Fb^T Z P ; gH += Fb
  gemm<false, false, 0>(x, ZP, Fb, prm.Z, p, p, m);      // Fb Z      (p x m, stored in an m*p buffer)
  gemm<true, false, 1>(x, Xb, prm.Z, ZP, m, p, m);        // Xb += Z^T (Fb Z)
  gemm<false, true, 0>(x, Kp, ZP, P, p, m, m);            // (Fb Z) P^T
  if (gZ) KFB_FOR(i, p * m) gZ[i] += Kp[i];
  x.sync();
  gemm<true, false, 0>(x, ZP, Fb, prm.Z, p, p, m);        // Fb^T Z
  gemm<false, false, 0>(x, Kp, ZP, P, p, m, m);           // (Fb^T Z) P
  if (gZ) KFB_FOR(i, p * m) gZ[i] += Kp[i];
  if (gH) KFB_FOR(i, p * p) gH[i] += Fb[i];
  x.sync();
  // closed loop at Pss
  const double* yp = &y0[0];
  update_observed<MK_STD>(x, prm, yp, 0.0, a, P, tmp, af, Pf, (LogAcc*)nullptr, false);
  gemm<false, false, 0>(x, W, prm.T, tmp.A, m, m, m);     // At = T (I - K Z)   (= (A - B K)^T of the reference)
  gemm<false, false, 0>(x, Kp, prm.T, tmp.K, m, m, p);    // Kp = T K  (m x p) = K_ref^T
  // S = At^T S At + sym(Xb)
  KFB_FOR(idx, m * m) {
    const int i = x.div_m(idx), j = idx - i * m;
    Ak[idx] = W[j * m + i];
    S[idx] = 0.5 * (Xb[idx] + Xb[j * m + i]);
  }
  x.sync();
  smith_doubling_ctx(x, Ak, S, tmp.S1, m);
  // gC += S ; gT += 2 S At P ; gZ += -2 Kp^T S At P ; gH += Kp^T S Kp
  if (gC) KFB_FOR(i, m * m) gC[i] += S[i];
  gemm<false, false, 0>(x, tmp.S1, S, W, m, m, m);        // S At
  gemm<false, false, 0>(x, tmp.S2, tmp.S1, P, m, m, m);   // S At P
  if (gT) KFB_FOR(i, m * m) gT[i] += 2.0 * tmp.S2[i];
  x.sync();
  gemm<true, false, 0>(x, ZP, Kp, tmp.S2, p, m, m);       // Kp^T (S At P)   (p x m)
  if (gZ) KFB_FOR(i, p * m) gZ[i] -= 2.0 * ZP[i];
  gemm<false, false, 0>(x, tmp.Mm, S, Kp, m, m, p);       // S Kp
  gemm<true, false, 0>(x, Q1, Kp, tmp.Mm, p, m, p);       // Kp^T S Kp
  if (gH) KFB_FOR(i, p * p) gH[i] += Q1[i];
  x.sync();
}

inline int dare_arena_doubles(int m, int p) {
  const int mm = m * m, mp = m * p, pp = p * p;
  const int params = mm + mp + pp + p + pp;
  const int upd = 3 * p + 3 * mp + 4 * pp + 3 * mm;
  return params + upd + 6 * mm + 3 * m + p + 2 * pp + 2 * mp + 64;
}

}  // namespace kfb
