// kf_rowsU.cuh - fused UnivariateFilter programs (reference kalman_filter.py:444-505) for mid-size systems
// (k_states 5..8, k_endog <= 3): loglik-only forward (+ tape) and the adjoint without Z-bar - the theta-level hot path of
// BASELINE.json configs[2] ("cholesky and univariate filters").
//
// Mapping: 8 lanes per unit, lane r owns ROW r of every m x m quantity (4 units per warp); the state vector a, the
// design rows Z and every per-observation scalar are REPLICATED in the registers of all lanes of the unit, so the p
// sequential scalar updates of a step need ONE exchange each:
//   forward   Mv = P z_i: every lane computes its element, all-gather through shared memory (m doubles), then
//             F, v, K = Mv / F, a += K v (replicated) and the rank-one downdate of the lane's row P[r][:] -= K[r] K[:] F;
//   predict   S1 = T Pf needs every row of Pf (all-gather, m^2 doubles), S2 = C + S1 T^T reads T from shared memory,
//             P' = sym(S2) needs column r of S2 (transpose exchange);
//   adjoint   only the SYMMETRIC part of the filtered-covariance cotangent is ever used (Kb and K^T Pfb K see
//             Pfb + Pfb^T; the step above applies sym), so the lane carries row r of Qh = sym(Pfb) and the update
//             Pfb += Mvb z^T becomes Qh += sym(Mvb z^T) with Mvb replicated - no transposes in the reverse sweep.  The
//             per-observation exchange is (Kb[r], (Qh K)[r]) = 2m doubles.  At t = 0 the antisymmetric part of
//             sum_i Mvb_i z_i^T is tracked as well, so that P0-bar comes out in the entry-wise "generic-op gauge"
//              exactly as the generic kernels and autograd of the oracle produce it.
// The generic sub-warp kernels (CoopCtxT<M,P,8>) run the same filter through ~20 / ~60 synchronised primitive calls
// per step and recompute the inner updates O(p^2) times in the adjoint; this version needs p + 2 / 2p + 3 warp
// synchronisations.  Units of a warp take identical control flow: shared observation stream, static matrices.
#pragma once
#include "kf_core.cuh"

namespace kfb {

template <int M, int P>
struct RowsULayout {
  static constexpr int G = 8;
  static constexpr int MM = M * M, ME = M + (M & 1);
  static constexpr int KT = M + (M * (M + 1)) / 2, KTP = (KT + 1) & ~1;
  static constexpr int T = 0, X = T + MM, S = X + MM, Ex = S + MM, av = Ex + 2 * 2 * ME, tp = av + ME,
                       END_FWD = tp, END_BWD = tp + 2 * KTP;
  // unit stride == 4 (mod 16) doubles: the four units of a warp hit disjoint bank groups on broadcast reads
  static constexpr int stride(int n) { return n + ((4 - (n % 16)) + 16) % 16; }
  static constexpr int fwd_doubles = stride(END_FWD), bwd_doubles = stride(END_BWD);
};

// all-gather helper: every lane of the unit reads the m doubles at p (broadcast loads)
template <int M>
__device__ __forceinline__ void rowsU_read_vec(const double* p, double (&out)[M]) {
#pragma unroll
  for (int k = 0; k < M; ++k) out[k] = p[k];
}

// One inner update of the univariate filter on observation i, replicated state.  Returns live (F != 0).
// In: Pr (lane's row of P), a (replicated).  Out: Mv (replicated), F, rF, v; Pr, a updated.
template <int M>
__device__ __forceinline__ bool rowsU_inner(double* ex, unsigned mask, int r, bool act, const double (&z)[M], double hd,
                                            double yi, double dd, double (&Pr)[M], double (&a)[M], double (&Mv)[M],
                                            double& F, double& rF, double& v) {
  double mv = 0.0;
#pragma unroll
  for (int k = 0; k < M; ++k) mv = fma(Pr[k], z[k], mv);
  if (act) ex[r] = mv;
  __syncwarp(mask);
  rowsU_read_vec<M>(ex, Mv);
  v = yi - dd;
  F = hd;
#pragma unroll
  for (int k = 0; k < M; ++k) {
    v = fma(-z[k], a[k], v);
    F = fma(z[k], Mv[k], F);
  }
  const bool live = (F != 0.0);
  rF = live ? 1.0 / F : 0.0;
  const double kr = mv * rF;  // the lane's own element of K (never index the replicated arrays with the run-time row)
#pragma unroll
  for (int k = 0; k < M; ++k) {
    const double kk = Mv[k] * rF;
    Pr[k] = fma(-kr * kk, F, Pr[k]);  // P - K K^T F  (:477, not Joseph)
    a[k] = fma(kk, v, a[k]);
  }
  return live;
}

// ------------------------------------------------------------------------------------------------ forward
template <int M, int P>
__device__ void rowsU_forward(const KfArgs& A, long long u, double* sm, int l, unsigned mask) {
  using L = RowsULayout<M, P>;
  constexpr int KT = L::KT;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const bool act = l < M;         // lanes M..7 of the unit (and the warp's spare lanes) shadow row 0 without storing
  const int r = act ? l : 0;
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  const double* Cp = A.C.p + draw * A.C.bs;
  const double* P0p = A.P0.p + draw * A.P0.bs;
  const double* a0p = A.a0.p + draw * A.a0.bs;
  if (l < L::G)
    for (int k = l; k < M * M; k += L::G) sm[L::T + k] = Tp[k];
  double Tr[M], Cr[M], Pr[M], a[M], z[P][M], hd[P], dd[P];
#pragma unroll
  for (int j = 0; j < M; ++j) {
    Tr[j] = Tp[r * M + j];
    Cr[j] = Cp[r * M + j];
    Pr[j] = P0p[r * M + j];
    a[j] = a0p[j];
  }
#pragma unroll
  for (int i = 0; i < P; ++i) {
#pragma unroll
    for (int k = 0; k < M; ++k) z[i][k] = Zp[i * M + k];
    hd[i] = Hp[i * P + i];  // only diag(H) is used (:493)
    dd[i] = A.d.p ? A.d_sign * A.d.p[draw * A.d.bs + i] : 0.0;
  }
  const double cr = (act && A.c.p) ? A.c.p[draw * A.c.bs + r] : 0.0;
  __syncwarp(mask);

  const double* y = A.y.p;
  LogAcc acc;
  double qsum = 0.0;
  int cnt = 0, info = 0;
  double* tp = A.tape ? A.tape + u * (long long)(n - 1) * KT : nullptr;
  double yn[P];
#pragma unroll
  for (int i = 0; i < P; ++i) yn[i] = y[i];
  int slot = 0;
  for (int t = 0; t < n; ++t) {
    double yt[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      yt[i] = yn[i];
      yn[i] = y[(long long)(t + 1 < n ? t + 1 : t) * P + i];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
      if (kf_isnan(yt[i])) continue;  // uniform: one observation stream for the whole warp
      double Mv[M], F, rF, v;
      const bool live = rowsU_inner<M>(sm + L::Ex + slot * 2 * L::ME, mask, r, act, z[i], hd[i], yt[i], dd[i], Pr, a, Mv, F, rF, v);
      slot ^= 1;
      if (live) {
        if (!(F > 0.0) && info == 0) info = t + 1;
        if (F > 0.0 && F < 1.0e300) acc.mul(F);
        qsum = fma(v * v, rF, qsum);
        cnt += 1;
      }
    }
    // ---- predict: a' = T af + c ; P' = sym(T Pf T^T + C)
    double an = cr;
#pragma unroll
    for (int k = 0; k < M; ++k) an = fma(Tr[k], a[k], an);
    if (act) {
      sm[L::av + r] = an;
#pragma unroll
      for (int j = 0; j < M; ++j) sm[L::X + r * M + j] = Pr[j];
    }
    __syncwarp(mask);
    double S1[M], S2[M];
#pragma unroll
    for (int j = 0; j < M; ++j) S1[j] = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
#pragma unroll
      for (int j = 0; j < M; ++j) S1[j] = fma(Tr[k], sm[L::X + k * M + j], S1[j]);  // (T Pf)[r][j]
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s = Cr[j];
#pragma unroll
      for (int k = 0; k < M; ++k) s = fma(S1[k], sm[L::T + j * M + k], s);  // + S1 T^T
      S2[j] = s;
    }
    if (act) {
#pragma unroll
      for (int j = 0; j < M; ++j) sm[L::S + r * M + j] = S2[j];
    }
    rowsU_read_vec<M>(sm + L::av, a);
    __syncwarp(mask);
#pragma unroll
    for (int j = 0; j < M; ++j) Pr[j] = 0.5 * (S2[j] + sm[L::S + j * M + r]);
    if (tp && t + 1 < n) {
      if (act) {
        tp[r] = an;
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (j >= r) tp[M + r * M - (r * (r - 1)) / 2 + (j - r)] = Pr[j];
      }
      tp += KT;
    }
    __syncwarp(mask);  // X / S / av are rewritten by the next step
  }
  if (l == 0) {
    double ll = -0.5 * ((double)cnt * KF_LOG_2PI + qsum + acc.value());
    if (info != 0) ll = nan("");
    if (A.loglik) A.loglik[u] = ll;
    if (A.info) A.info[u] = info;
  }
}

// ------------------------------------------------------------------------------------------------ adjoint
template <int KT>
__device__ __forceinline__ void rowsU_tape_prefetch(double* dst, const double* src, int l) {
#ifdef __CUDA_ARCH__
  if (l < 8) {
    const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll
    for (int k0 = 0; k0 < KT; k0 += 8) {
      const int k = k0 + l;
      if (k < KT) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + (unsigned)(k * 8)), "l"(src + k) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}

template <int M, int P>
__device__ void rowsU_backward(const KfArgs& A, long long u, double* sm, int l, unsigned mask) {
  using L = RowsULayout<M, P>;
  constexpr int KT = L::KT;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const bool act = l < M;
  const int r = act ? l : 0;
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  const double* tape = A.tape + u * (long long)(n - 1) * KT;
  if (n >= 2) rowsU_tape_prefetch<KT>(sm + L::tp + ((n - 1) & 1) * L::KTP, tape + (long long)(n - 2) * KT, l);
  if (l < L::G)
    for (int k = l; k < M * M; k += L::G) sm[L::T + k] = Tp[k];
  double z[P][M], hd[P], dd[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
#pragma unroll
    for (int k = 0; k < M; ++k) z[i][k] = Zp[i * M + k];
    hd[i] = Hp[i * P + i];
    dd[i] = A.d.p ? A.d_sign * A.d.p[draw * A.d.bs + i] : 0.0;
  }
  __syncwarp(mask);
  const double* y = A.y.p;
  const double gl = A.g_loglik ? A.g_loglik[u] : 1.0;
  // running cotangents: ab (replicated), row r of Qh = sym(P-bar); accumulators: rows of Tb, Cb; cb[r]; Hb, db (replicated)
  double ab[M], Qh[M], Tb[M], Cb[M], Ah[M], cb = 0.0, Hb[P], db[P];
#pragma unroll
  for (int j = 0; j < M; ++j) ab[j] = Qh[j] = Tb[j] = Cb[j] = Ah[j] = 0.0;
#pragma unroll
  for (int i = 0; i < P; ++i) Hb[i] = db[i] = 0.0;
  double zr[P];  // the lane's own element of each design row (the replicated arrays are never indexed with the run-time row)
#pragma unroll
  for (int i = 0; i < P; ++i) zr[i] = Zp[i * M + r];
  double abr = 0.0;  // own element of ab
  double yn[P];
#pragma unroll
  for (int i = 0; i < P; ++i) yn[i] = y[(long long)(n - 1) * P + i];
  int slot = 0;
  for (int t = n - 1; t >= 0; --t) {
    double a[M], Pr[M];
    if (t == 0) {
#pragma unroll
      for (int j = 0; j < M; ++j) {
        a[j] = A.a0.p[draw * A.a0.bs + j];
        Pr[j] = A.P0.p[draw * A.P0.bs + r * M + j];
      }
    } else {
#ifdef __CUDA_ARCH__
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
      __syncwarp(mask);
      const double* tq = sm + L::tp + (t & 1) * L::KTP;
#pragma unroll
      for (int j = 0; j < M; ++j) {
        a[j] = tq[j];
        const int lo = r < j ? r : j, hi = r < j ? j : r;
        Pr[j] = tq[M + lo * M - (lo * (lo - 1)) / 2 + (hi - lo)];
      }
      if (t >= 2) rowsU_tape_prefetch<KT>(sm + L::tp + ((t - 1) & 1) * L::KTP, tape + (long long)(t - 2) * KT, l);
    }
    double yt[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      yt[i] = yn[i];
      yn[i] = y[(long long)(t > 0 ? t - 1 : 0) * P + i];
    }
    const double lb = gl + (A.g_ll_obs ? A.g_ll_obs[u * n + t] : 0.0);
    // ---- forward recompute of the p inner updates, keeping what their adjoints need
    double Mv[P][M], Fv[P], rFv[P], vv[P];
    bool obs[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      obs[i] = !kf_isnan(yt[i]);
      Fv[i] = 1.0; rFv[i] = 0.0; vv[i] = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) Mv[i][k] = 0.0;
      if (!obs[i]) continue;
      rowsU_inner<M>(sm + L::Ex + slot * 2 * L::ME, mask, r, act, z[i], hd[i], yt[i], dd[i], Pr, a, Mv[i], Fv[i], rFv[i], vv[i]);
      slot ^= 1;
    }
    // (a, Pr) are now the filtered moments af, Pf
    // ---- adjoint of predict: Cb += Ps ; cb += ab ; W = Ps T ; Tb += ab af^T + W (Pf + Pf^T) ; afb = T^T ab ; Qh = T^T W
    if (act) {
#pragma unroll
      for (int j = 0; j < M; ++j) sm[L::X + r * M + j] = Pr[j];
    }
    double W[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      Cb[j] += Qh[j];
      W[j] = 0.0;
    }
    cb += abr;
#pragma unroll
    for (int k = 0; k < M; ++k) {
#pragma unroll
      for (int j = 0; j < M; ++j) W[j] = fma(Qh[k], sm[L::T + k * M + j], W[j]);  // (Ps T)[r][j]
    }
    double afr = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) afr = fma(sm[L::T + k * M + r], ab[k], afr);  // (T^T ab)[r]
    if (act) {
#pragma unroll
      for (int j = 0; j < M; ++j) sm[L::S + r * M + j] = W[j];
      sm[L::av + r] = afr;
    }
    __syncwarp(mask);
#pragma unroll
    for (int j = 0; j < M; ++j) Tb[j] = fma(abr, a[j], Tb[j]);
    if (t > 0) {
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const double w2 = W[k] + W[k];  // rank-one downdates of a symmetric tape entry: Pf is exactly symmetric, Pf + Pf^T = 2 Pf
#pragma unroll
        for (int j = 0; j < M; ++j) Tb[j] = fma(w2, sm[L::X + k * M + j], Tb[j]);
      }
    } else {  // t = 0: P0 is the caller's matrix, possibly non-symmetric
#pragma unroll
      for (int k = 0; k < M; ++k) {
#pragma unroll
        for (int j = 0; j < M; ++j) Tb[j] = fma(W[k], sm[L::X + k * M + j] + sm[L::X + j * M + k], Tb[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) s = fma(sm[L::T + k * M + r], sm[L::S + k * M + j], s);  // (T^T W)[r][j]
      Qh[j] = s;
    }
    double afb[M];
    rowsU_read_vec<M>(sm + L::av, afb);
    // ---- reverse the p scalar updates
    const double lq = -0.5 * lb;
#pragma unroll
    for (int i = P - 1; i >= 0; --i) {
      if (!obs[i]) continue;
      const double F = Fv[i], rF = rFv[i], v = vv[i];
      double Kv[M];
#pragma unroll
      for (int k = 0; k < M; ++k) Kv[k] = Mv[i][k] * rF;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) s = fma(Qh[k], Kv[k], s);              // (sym(Pfb) K)[r]
      const double kb = fma(afr, v, -2.0 * s * F);                        // Kb[r] = afb[r] v - ((Pfb + Pfb^T) K)[r] F
      double* ex = sm + L::Ex + slot * 2 * L::ME;
      slot ^= 1;
      if (act) {
        ex[r] = kb;
        ex[L::ME + r] = s;
      }
      __syncwarp(mask);
      double Kb[M], Sv[M];
      rowsU_read_vec<M>(ex, Kb);
      rowsU_read_vec<M>(ex + L::ME, Sv);
      double ktab = 0.0, kpk = 0.0, kbk = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) {
        ktab = fma(Kv[k], afb[k], ktab);
        kbk = fma(Kb[k], Kv[k], kbk);
        kpk = fma(Kv[k], Sv[k], kpk);
      }
      const double vbar = ktab + lq * 2.0 * v * rF;
      const double Fbar = -kpk + lq * (rF - v * v * rF * rF) - kbk * rF;
      double Mvb[M];
#pragma unroll
      for (int k = 0; k < M; ++k) Mvb[k] = fma(z[i][k], Fbar, Kb[k] * rF);  // Mvb = Kb / F + z Fbar
      Hb[i] += Fbar;
      db[i] -= vbar;
      const double mr = fma(zr[i], Fbar, kb * rF);                         // own element of Mvb
#pragma unroll
      for (int k = 0; k < M; ++k) {
        Qh[k] = fma(0.5, fma(mr, z[i][k], zr[i] * Mvb[k]), Qh[k]);             // + sym(Mvb z^T)
        if (t == 0) Ah[k] = fma(0.5, fma(mr, z[i][k], -(zr[i] * Mvb[k])), Ah[k]);  // antisymmetric part (P0-bar gauge)
        afb[k] = fma(-z[i][k], vbar, afb[k]);
      }
      afr = fma(-zr[i], vbar, afr);
    }
    abr = afr;
#pragma unroll
    for (int k = 0; k < M; ++k) ab[k] = afb[k];
    __syncwarp(mask);  // X / S / av / Ex are rewritten by the next step
  }
  if (act) {
    if (A.ga0) A.ga0[u * M + r] = abr;
    if (A.gc) A.gc[u * M + r] = cb;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      if (A.gP0) A.gP0[u * M * M + r * M + j] = Qh[j] + Ah[j];
      if (A.gT) A.gT[u * M * M + r * M + j] = Tb[j];
      if (A.gC) A.gC[u * M * M + r * M + j] = Cb[j];
    }
  }
  if (l == 0) {
#pragma unroll
    for (int i = 0; i < P; ++i) {
      if (A.gd) A.gd[u * P + i] = A.d_sign * db[i];
      if (A.gH) {
#pragma unroll
        for (int j = 0; j < P; ++j) A.gH[u * P * P + i * P + j] = (i == j) ? Hb[i] : 0.0;
      }
    }
  }
}

}  // namespace kfb
