// kf_rows.cuh - fused "row per lane" programs for mid-size systems (k_states 5..8, k_endog <= 3, MK_STD).
//
// Same mathematics as kf_pred.cuh (one-step-predictor form and its adjoint) but written directly for the sub-warp
// mapping instead of through the generic primitive-per-phase abstraction: 8 lanes per unit, lane i owns ROW i of every
// m x m / m x p quantity.  Everything a lane can compute from its own rows stays in registers; only what OTHER lanes
// must read goes through shared memory, and the p x p inverse, w = F^-1 v, K^T Kb, v-bar, F-bar are computed redundantly
// by every lane in registers.  The generic CoopCtxT path needs ~14 (forward) / ~40 (adjoint) warp-synchronised phases
// per filter step, each paying shared-memory + 32-cycle DFMA latency with only 8-12 warps per SM to hide it; this
// version needs 5 / 8, uses ~4 KB instead of ~6.4 KB of shared memory per unit (3 instead of 2 CTAs per SM in the
// adjoint) and keeps the gradient accumulators in registers.
// Units of a warp must take identical control flow: shared observation stream, static matrices (enforced by the launcher).
#pragma once
#include "kf_core.cuh"

namespace kfb {

template <int M, int P>
struct RowsLayout {
  static constexpr int MM = M * M, MP = M * P, PP = P * P;
  // forward + adjoint share the first block
  static constexpr int T = 0, Z = T + MM, H = Z + MP, Pm = H + PP + (PP & 1), Mm = Pm + MM, Kp = Mm + MP, Lm = Kp + MP,
                       F = Lm + MM, a = F + PP + (PP & 1), v = a + M + (M & 1), END_COMMON = v + P + (P & 1);
  // forward only
  static constexpr int C = END_COMMON, S2 = C + MM, END_FWD = S2 + MM;
  // adjoint only
  static constexpr int Pb = END_COMMON, Ps = Pb + MM, X = Ps + MM, W = X + MM, Lb = W + MM, Kb = Lb + MM, TMb = Kb + MP,
                       Mb = TMb + MP, PK = Mb + MP, ab = PK + MP, tp = ab + M + (M & 1),
                       END_BWD = tp + M + (M * (M + 1)) / 2 + 1;
  static constexpr int fwd_doubles = (END_FWD + 1) & ~1, bwd_doubles = (END_BWD + 1) & ~1;
};

// ------------------------------------------------------------------------------------------------ shared pieces
// Phases A, B, D of a step for an observed row: v, Mm, F | TM | (F^-1, w, quad) Kp, Lm.   Leaves in shared memory:
// v, Mm, F, Kp, Lm; in registers of lane i: TM row, Kp row, Lm row, and (every lane) Fi, w.
template <int M, int P>
struct RowGain {
  double TM[P], Kp[P], Lm[M], Fi[P * P], w[P], piv[P], quad;
  bool ok;
};

template <int M, int P>
__device__ __forceinline__ void rows_gain(double* sm, const double* yt, double d_sign, double di, int lane, unsigned mask,
                                          RowGain<M, P>& g) {
  using L = RowsLayout<M, P>;
  const bool act = lane < M;
  const int i = act ? lane : 0;
  // ---- phase A: v (lanes < P), Mm row
  if (lane < P) {
    double s = yt[lane] - d_sign * di;
#pragma unroll
    for (int k = 0; k < M; ++k) s = fma(-sm[L::Z + lane * M + k], sm[L::a + k], s);
    sm[L::v + lane] = s;
  }
  if (act) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) s = fma(sm[L::Pm + i * M + k], sm[L::Z + j * M + k], s);
      sm[L::Mm + i * P + j] = s;
    }
  }
  __syncwarp(mask);
  // ---- phase B: F rows (lanes < P), TM row (registers)
  if (lane < P) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      double s = sm[L::H + lane * P + j];
#pragma unroll
      for (int k = 0; k < M; ++k) s = fma(sm[L::Z + lane * M + k], sm[L::Mm + k * P + j], s);
      sm[L::F + lane * P + j] = s;
    }
  }
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) s = fma(sm[L::T + i * M + k], sm[L::Mm + k * P + j], s);
    g.TM[j] = s;
  }
  __syncwarp(mask);
  // ---- phase D: every lane inverts F in registers; Kp, Lm rows
  double Fr[P * P], Lr[P * P], Lir[P * P];
#pragma unroll
  for (int k = 0; k < P * P; ++k) Fr[k] = sm[L::F + k];
  g.ok = ldl_inverse(Fr, g.Fi, Lr, Lir, g.piv, P);
  double q = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) s = fma(g.Fi[j * P + k], sm[L::v + k], s);
    g.w[j] = s;
    q = fma(sm[L::v + j], s, q);
  }
  g.quad = q;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) s = fma(g.TM[k], g.Fi[k * P + j], s);
    g.Kp[j] = s;
  }
#pragma unroll
  for (int j = 0; j < M; ++j) {
    double s = sm[L::T + i * M + j];
#pragma unroll
    for (int k = 0; k < P; ++k) s = fma(-g.Kp[k], sm[L::Z + k * M + j], s);
    g.Lm[j] = s;
  }
  if (act) {
#pragma unroll
    for (int j = 0; j < P; ++j) sm[L::Kp + i * P + j] = g.Kp[j];
#pragma unroll
    for (int j = 0; j < M; ++j) sm[L::Lm + i * M + j] = g.Lm[j];
  }
  __syncwarp(mask);
}

template <int P>
__device__ __forceinline__ int rows_count_missing(const double* yt) {
  int nm = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) nm += (yt[i] != yt[i]) ? 1 : 0;
  return nm;
}

// ------------------------------------------------------------------------------------------------ forward
template <int M, int P>
__device__ void rows_forward(const KfArgs& A, long long u, double* sm, int lane, unsigned mask) {
  using L = RowsLayout<M, P>;
  constexpr int KT = M + (M * (M + 1)) / 2;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const bool act = lane < M;
  const int i = act ? lane : 0;
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  const double* Cp = A.C.p + draw * A.C.bs;
  const double* P0p = A.P0.p + draw * A.P0.bs;
  const double* a0p = A.a0.p + draw * A.a0.bs;
  for (int k = lane; k < M * M; k += 8) {
    sm[L::T + k] = Tp[k];
    sm[L::C + k] = Cp[k];
    sm[L::Pm + k] = P0p[k];
  }
  for (int k = lane; k < P * M; k += 8) sm[L::Z + k] = Zp[k];
  for (int k = lane; k < P * P; k += 8) sm[L::H + k] = Hp[k];
  if (act) sm[L::a + i] = a0p[i];
  const double ci = (act && A.c.p) ? A.c.p[draw * A.c.bs + i] : 0.0;
  const double di = (lane < P && A.d.p) ? A.d.p[draw * A.d.bs + lane] : 0.0;
  __syncwarp(mask);

  const double* y = A.y.p;
  LogAcc acc;
  double llsum = 0.0;
  int info = 0;
  double* tp = A.tape ? A.tape + u * (long long)(n - 1) * KT : nullptr;
  RowGain<M, P> g;

  for (int t = 0; t < n; ++t) {
    const double* yt = y + (long long)t * P;
    const int nm = rows_count_missing<P>(yt);
    double an = ci, S1[M], S2[M];
    if (nm == 0) {
      rows_gain<M, P>(sm, yt, A.d_sign, di, lane, mask, g);
      if (!g.ok && info == 0) info = t + 1;
      if (g.ok) {
#pragma unroll
        for (int k = 0; k < P; ++k) acc.mul(g.piv[k]);
      }
      llsum += -0.5 * (A.ll_const + g.quad);
      // ---- phase E: a' row, S2 row = C + (L P) L^T + (Kp H) Kp^T
#pragma unroll
      for (int k = 0; k < M; ++k) an = fma(sm[L::T + i * M + k], sm[L::a + k], an);
#pragma unroll
      for (int k = 0; k < P; ++k) an = fma(g.Kp[k], sm[L::v + k], an);
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(g.Lm[k], sm[L::Pm + k * M + j], s);
        S1[j] = s;
      }
      double KH[P];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(g.Kp[k], sm[L::H + k * P + j], s);
        KH[j] = s;
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = sm[L::C + i * M + j];
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(S1[k], sm[L::Lm + j * M + k], s);
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(KH[k], sm[L::Kp + j * P + k], s);
        S2[j] = s;
      }
    } else {
      if (nm != P && info == 0) info = -(t + 1);
#pragma unroll
      for (int k = 0; k < M; ++k) an = fma(sm[L::T + i * M + k], sm[L::a + k], an);
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(sm[L::T + i * M + k], sm[L::Pm + k * M + j], s);
        S1[j] = s;
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = sm[L::C + i * M + j];
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(S1[k], sm[L::T + j * M + k], s);
        S2[j] = s;
      }
    }
    if (act) {
#pragma unroll
      for (int j = 0; j < M; ++j) sm[L::S2 + i * M + j] = S2[j];
    }
    __syncwarp(mask);
    // ---- phase F: P' = sym(S2), a' ; tape
    if (act) {
#pragma unroll
      for (int j = 0; j < M; ++j) sm[L::Pm + i * M + j] = 0.5 * (S2[j] + sm[L::S2 + j * M + i]);
      sm[L::a + i] = an;
    }
    __syncwarp(mask);
    if (tp && t + 1 < n) {
      if (act) {
        tp[i] = an;
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (j >= i) tp[M + i * M - (i * (i - 1)) / 2 + (j - i)] = sm[L::Pm + i * M + j];
      }
      tp += KT;
    }
  }
  if (lane == 0) {
    double ll = llsum - 0.5 * acc.value();
    if (info != 0) ll = nan("");
    if (A.loglik) A.loglik[u] = ll;
    if (A.info) A.info[u] = info;
  }
}

// ------------------------------------------------------------------------------------------------ adjoint
template <int M, int P>
__device__ void rows_backward(const KfArgs& A, long long u, double* sm, int lane, unsigned mask) {
  using L = RowsLayout<M, P>;
  constexpr int KT = M + (M * (M + 1)) / 2;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const bool act = lane < M;
  const int i = act ? lane : 0;
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  for (int k = lane; k < M * M; k += 8) {
    sm[L::T + k] = Tp[k];
    sm[L::Pb + k] = 0.0;
  }
  for (int k = lane; k < P * M; k += 8) sm[L::Z + k] = Zp[k];
  for (int k = lane; k < P * P; k += 8) sm[L::H + k] = Hp[k];
  if (act) sm[L::ab + i] = 0.0;
  const double di = (lane < P && A.d.p) ? A.d.p[draw * A.d.bs + lane] : 0.0;
  __syncwarp(mask);

  const double* y = A.y.p;
  const double gl = A.g_loglik ? A.g_loglik[u] : 1.0;
  const bool need_Z = (A.gZ != nullptr), need_H = (A.gH != nullptr);
  // gradient accumulators: lane i holds row i (Tb, Cb), lanes < P hold rows of Zb, Hb; element i of cb, db
  double Tb[M], Cb[M], Zb[M], Hb[P], cb = 0.0, db = 0.0, abi = 0.0;
#pragma unroll
  for (int j = 0; j < M; ++j) Tb[j] = Cb[j] = Zb[j] = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) Hb[j] = 0.0;
  const double* tp = A.tape + u * (long long)(n - 1) * KT + (long long)(n - 2) * KT;
  RowGain<M, P> g;

  for (int t = n - 1; t >= 0; --t) {
    // ---- predicted moments of step t -> shared memory
    if (t == 0) {
      const double* P0p = A.P0.p + draw * A.P0.bs;
      for (int k = lane; k < M * M; k += 8) sm[L::Pm + k] = P0p[k];
      if (act) sm[L::a + i] = A.a0.p[draw * A.a0.bs + i];
    } else {
      for (int k = lane; k < KT; k += 8) sm[L::tp + k] = tp[k];
      tp -= KT;
      __syncwarp(mask);
      if (act) {
        sm[L::a + i] = sm[L::tp + i];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          const int lo = i < j ? i : j, hi = i < j ? j : i;
          sm[L::Pm + i * M + j] = sm[L::tp + M + lo * M - (lo * (lo - 1)) / 2 + (hi - lo)];
        }
      }
    }
    __syncwarp(mask);
    const double* yt = y + (long long)t * P;
    const double lb = gl + (A.g_ll_obs ? A.g_ll_obs[u * n + t] : 0.0);
    const bool observed = (rows_count_missing<P>(yt) == 0);
    if (observed) {
      rows_gain<M, P>(sm, yt, A.d_sign, di, lane, mask, g);
    } else {
#pragma unroll
      for (int j = 0; j < M; ++j) g.Lm[j] = sm[L::T + i * M + j];
      if (act) {
#pragma unroll
        for (int j = 0; j < M; ++j) sm[L::Lm + i * M + j] = g.Lm[j];
      }
      __syncwarp(mask);
    }
    // ---- phase 1: Ps row, X = L (P + P^T) row ; Cb, cb
    double Ps[M];
#pragma unroll
    for (int j = 0; j < M; ++j) Ps[j] = 0.5 * (sm[L::Pb + i * M + j] + sm[L::Pb + j * M + i]);
    abi = sm[L::ab + i];
    if (act) {
#pragma unroll
      for (int j = 0; j < M; ++j) {
        sm[L::Ps + i * M + j] = Ps[j];
        Cb[j] += Ps[j];
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(g.Lm[k], sm[L::Pm + k * M + j] + sm[L::Pm + j * M + k], s);
        sm[L::X + i * M + j] = s;
      }
      cb += abi;
    }
    __syncwarp(mask);
    // ---- phase 2: Lb = Ps X, W = Ps L, T^T ab, (observed) PK = Ps Kp, Kb
    double Lb[M], PK[P], Kb[P], abn = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s = 0.0, s2 = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) {
        s = fma(Ps[k], sm[L::X + k * M + j], s);
        s2 = fma(Ps[k], sm[L::Lm + k * M + j], s2);
      }
      Lb[j] = s;
      Tb[j] += fma(abi, sm[L::a + j], s);  // Tb += ab a^T + Lb
      if (act) sm[L::W + i * M + j] = s2;
    }
#pragma unroll
    for (int k = 0; k < M; ++k) abn = fma(sm[L::T + k * M + i], sm[L::ab + k], abn);
    if (observed) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(Ps[k], sm[L::Kp + k * P + j], s);
        PK[j] = s;
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s = abi * sm[L::v + j];
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(PK[k], sm[L::H + k * P + j] + sm[L::H + j * P + k], s);
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(-Lb[k], sm[L::Z + j * M + k], s);
        Kb[j] = s;
      }
      if (act) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
          sm[L::Kb + i * P + j] = Kb[j];
          if (need_H) sm[L::PK + i * P + j] = PK[j];
        }
        if (need_Z) {
#pragma unroll
          for (int j = 0; j < M; ++j) sm[L::Lb + i * M + j] = Lb[j];
        }
      }
    }
    __syncwarp(mask);
    // ---- phase 3: Pb' = L^T W (registers) ; (observed) every lane: K^T Kb, vb, Fb ; TMb row, Tb += TMb Mm^T
    double Pbn[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) s = fma(sm[L::Lm + k * M + i], sm[L::W + k * M + j], s);
      Pbn[j] = s;
    }
    double vb[P], Fb[P * P];
    if (observed) {
      double Q1[P * P];
#pragma unroll
      for (int a2 = 0; a2 < P; ++a2) {
        double s = -lb * g.w[a2];
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(sm[L::Kp + k * P + a2], sm[L::ab + k], s);
        vb[a2] = s;
#pragma unroll
        for (int b2 = 0; b2 < P; ++b2) {
          double q = 0.0;
#pragma unroll
          for (int k = 0; k < M; ++k) q = fma(sm[L::Kp + k * P + a2], sm[L::Kb + k * P + b2], q);
          Q1[a2 * P + b2] = q;
        }
      }
#pragma unroll
      for (int a2 = 0; a2 < P; ++a2)
#pragma unroll
        for (int b2 = 0; b2 < P; ++b2) {
          double s = -0.5 * lb * (g.Fi[b2 * P + a2] - g.w[a2] * g.w[b2]);
#pragma unroll
          for (int k = 0; k < P; ++k) s = fma(-Q1[a2 * P + k], g.Fi[b2 * P + k], s);
          Fb[a2 * P + b2] = s;
        }
      double TMb[P];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(Kb[k], g.Fi[j * P + k], s);
        TMb[j] = s;
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = Tb[j];
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(TMb[k], sm[L::Mm + j * P + k], s);
        Tb[j] = s;
      }
      if (act) {
#pragma unroll
        for (int j = 0; j < P; ++j) sm[L::TMb + i * P + j] = TMb[j];
      }
    }
    __syncwarp(mask);
    // ---- phase 4: (observed) Mb = T^T TMb + Z^T Fb ; Pb' += Mb Z ; ab' = T^T ab - Z^T vb ; store Pb', ab'
    double Mb[P];
    if (observed) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(sm[L::T + k * M + i], sm[L::TMb + k * P + j], s);
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(sm[L::Z + k * M + i], Fb[k * P + j], s);
        Mb[j] = s;
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
#pragma unroll
        for (int k = 0; k < P; ++k) Pbn[j] = fma(Mb[k], sm[L::Z + k * M + j], Pbn[j]);
      }
#pragma unroll
      for (int k = 0; k < P; ++k) abn = fma(-sm[L::Z + k * M + i], vb[k], abn);
#pragma unroll
      for (int q = 0; q < P; ++q)
        if (lane == q) db = fma(-A.d_sign, vb[q], db);
      if (need_Z && act) {
#pragma unroll
        for (int j = 0; j < P; ++j) sm[L::Mb + i * P + j] = Mb[j];
      }
    }
    __syncwarp(mask);  // every lane has finished reading Pb, ab, Lm, W of this step
    if (act) {
#pragma unroll
      for (int j = 0; j < M; ++j) sm[L::Pb + i * M + j] = Pbn[j];
      sm[L::ab + i] = abn;
    }
    // ---- optional cotangents that need cross-row reductions (lanes < P own the rows of Zb, Hb)
    if (observed && (need_Z || need_H)) {
#pragma unroll
      for (int q = 0; q < P; ++q) {  // static index q instead of vb[lane] / Fb[lane * P + k]: keeps both in registers
        if (lane != q) continue;
        if (need_Z) {
#pragma unroll
          for (int j = 0; j < M; ++j) {
            double s = fma(-vb[q], sm[L::a + j], Zb[j]);
#pragma unroll
            for (int k = 0; k < M; ++k) {
              s = fma(-sm[L::Kp + k * P + q], sm[L::Lb + k * M + j], s);   // - Kp^T Lb
              s = fma(sm[L::Mb + k * P + q], sm[L::Pm + k * M + j], s);    // + Mb^T P
            }
#pragma unroll
            for (int k = 0; k < P; ++k) s = fma(Fb[q * P + k], sm[L::Mm + j * P + k], s);  // + Fb Mm^T
            Zb[j] = s;
          }
        }
        if (need_H) {
#pragma unroll
          for (int j = 0; j < P; ++j) {
            double s = Hb[j] + Fb[q * P + j];
#pragma unroll
            for (int k = 0; k < M; ++k) s = fma(sm[L::Kp + k * P + q], sm[L::PK + k * P + j], s);  // + Kp^T Ps Kp
            Hb[j] = s;
          }
        }
      }
    }
    __syncwarp(mask);
  }
  // ---- write-out (row i by lane i)
  if (act) {
    if (A.ga0) A.ga0[u * M + i] = sm[L::ab + i];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      if (A.gP0) A.gP0[u * M * M + i * M + j] = sm[L::Pb + i * M + j];
      if (A.gT) A.gT[u * M * M + i * M + j] = Tb[j];
      if (A.gC) A.gC[u * M * M + i * M + j] = Cb[j];
    }
    if (A.gc) A.gc[u * M + i] = cb;
  }
  if (lane < P) {
    if (A.gd) A.gd[u * P + lane] = db;
#pragma unroll
    for (int j = 0; j < M; ++j)
      if (A.gZ) A.gZ[u * P * M + lane * M + j] = Zb[j];
#pragma unroll
    for (int j = 0; j < P; ++j)
      if (A.gH) A.gH[u * P * P + lane * P + j] = Hb[j];
  }
}

}  // namespace kfb
