// kf_rows.cuh - fused "row block per lane" programs for mid-size systems (k_states 5..8, k_endog <= 3, MK_STD).
//
// Same mathematics as kf_pred.cuh (one-step-predictor form and its adjoint) but written directly for a sub-warp
// mapping instead of through the generic primitive-per-phase abstraction: G lanes per unit, lane l owns the
// R = ceil(M / G) ROWS l*R .. l*R+R-1 of every m x m / m x p quantity.  Everything a lane can compute from its own rows
// stays in registers; only what OTHER lanes must read goes through shared memory, and the p x p inverse, w = F^-1 v,
// K^T Kb, v-bar, F-bar are computed redundantly by every lane in registers.  The generic CoopCtxT path needs ~14 (forward)
// / ~40 (adjoint) warp-synchronised phases per filter step; this version needs 5 / 8.
//
// Why R > 1: with one row per lane every multiply-add needs one shared-memory operand (the other comes from the lane's
// registers), and ncu shows the kernels bound by shared-memory wavefronts (90 % / 80 % of the LSU peak, fp64 pipe 32 %,
// profiles/r1_ncu_rows.md).  With R rows per lane every broadcast operand feeds R multiply-adds and a warp holds 32/G
// units, so the wavefronts per multiply-add drop by ~R * (8/G).
// Units of a warp must take identical control flow: shared observation stream, static matrices (enforced by the launcher).
#pragma once
#include "kf_core.cuh"

namespace kfb {

template <int M, int P, int G>
struct RowsCfg {
  static constexpr int R = (M + G - 1) / G;  // rows per lane
  static constexpr int UPW = 32 / G;         // units per warp (lanes >= UPW*G idle)
  static_assert(G >= P && G <= 32, "lanes 0..P-1 own the rows of v, F, Zb, Hb");
};

// NZ = the adjoint also produces Z-bar.  Without it Lb / Mb are never exchanged and nothing reads Pm after phase 1, so
// W lives in Pm's slot; TMb (written in phase 3) always reuses X's slot (last read in phase 2).  At m = 6, p = 3 that
// is 422 doubles per unit = 8 resident units-of-8 per SM instead of 6.
// ST = steady-state filter: one more m x p exchange buffer (TM rows, for Gss-bar).
template <int M, int P, bool NZ = true, bool ST = false>
struct RowsLayout {
  static constexpr int MM = M * M, MP = M * P, PP = P * P, MPE = MP + (MP & 1);
  static constexpr int KT = M + (M * (M + 1)) / 2, KTP = (KT + 1) & ~1;
  // forward + adjoint share the first block
  static constexpr int T = 0, Z = T + MM, H = Z + MP, Pm = H + PP + (PP & 1), Mm = Pm + MM, Kp = Mm + MPE,
                       Lm = Kp + MPE, a = Lm + MM, END_COMMON = a + M + (M & 1);
  // forward only
  static constexpr int S2 = END_COMMON, END_FWD = S2 + MM;
  // adjoint only (tp: double-buffered cp.async landing zone for the packed tape entry)
  // (W = Ps L has a slot of its own: round 1 formed X = L (P + P^T) there first and kept W in Pm's slot; L-bar = Ps X is now
  //  computed as W (P + P^T) from the lane's own W rows, so P is still being read when W is stored)
  static constexpr int Pb = END_COMMON, X = Pb + MM, W = NZ ? X + MM : X, Lb = X + 2 * MM,
                       Kb = NZ ? Lb + MM : X + MM, Mb = Kb + MPE, PK = NZ ? Mb + MPE : Kb + MPE, TMs = PK + MPE,
                       ab = TMs + (ST ? MPE : 0),
                       Cb = ab + M + (M & 1), tp = Cb + MM, TMb = tp + 2 * KTP, END_BWD = TMb + MPE;
  // unit stride == 6 (mod 16) doubles = 48 bytes modulo the 128-byte bank row: with 16-byte accesses served a quarter
  // warp (two units of 4 lanes) at a time, both the lanes' row blocks (96 bytes apart) and their column blocks (16 bytes
  // apart) of two neighbouring units then fall on distinct bank groups.  Measured: a stride == 0 (mod 4) gave 8-way
  // replays on row accesses, == 2 (mod 16) 2-way replays on every column access.
  static constexpr int stride(int n) { return n + ((6 - (n % 16)) + 16) % 16; }
  static constexpr int fwd_doubles = stride(END_FWD), bwd_doubles = stride(END_BWD);
};

// rows owned by lane l: r[q] (may be >= M: inactive), c[q] = clamped index that is always safe to read
template <int M, int R>
struct RowIdx {
  int r[R], c[R];
  bool a[R];
  __device__ __forceinline__ explicit RowIdx(int l) {
#pragma unroll
    for (int q = 0; q < R; ++q) {
      r[q] = l * R + q;
      a[q] = r[q] < M;
      c[q] = a[q] ? r[q] : 0;
    }
  }
};

// out[q] = p[c[q]]: the lane's R consecutive elements of a length-M vector (its block of a COLUMN-indexed access).
// Two rows per lane and even M: one aligned 16-byte load instead of two 8-byte loads that replay on the same banks.
template <int M, int R>
__device__ __forceinline__ void lane_block(const double* p, const RowIdx<M, R>& rw, double (&out)[R]) {
  if constexpr (R == 2 && (M % 2) == 0) {
    const double2 v2 = *reinterpret_cast<const double2*>(p + rw.c[0]);
    out[0] = v2.x;
    out[1] = v2.y;
  } else {
#pragma unroll
    for (int q = 0; q < R; ++q) out[q] = p[rw.c[q]];
  }
}

// base[c[q] * W + j] += v[q][j] on the lane's own rows (an accumulator the lane keeps in shared memory, not registers)
template <int M, int R, int W>
__device__ __forceinline__ void accumulate_rows(double* base, const RowIdx<M, R>& rw, const double (&v)[R][W]) {
  if constexpr (R == 2 && (M % 2) == 0) {
    if (rw.a[0]) {
      double2* d = reinterpret_cast<double2*>(base + rw.r[0] * W);
#pragma unroll
      for (int e = 0; e < W; ++e) {
        double2 c = d[e];
        c.x += v[(2 * e) / W][(2 * e) % W];
        c.y += v[(2 * e + 1) / W][(2 * e + 1) % W];
        d[e] = c;
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < R; ++q)
      if (rw.a[q]) {
#pragma unroll
        for (int j = 0; j < W; ++j) base[rw.r[q] * W + j] += v[q][j];
      }
  }
}

// p[r[q]] = v[q] for the lane's active rows (inverse of lane_block)
template <int M, int R>
__device__ __forceinline__ void store_block(double* p, const RowIdx<M, R>& rw, const double (&v)[R]) {
  if constexpr (R == 2 && (M % 2) == 0) {
    if (rw.a[0]) *reinterpret_cast<double2*>(p + rw.r[0]) = make_double2(v[0], v[1]);
  } else {
#pragma unroll
    for (int q = 0; q < R; ++q)
      if (rw.a[q]) p[rw.r[q]] = v[q];
  }
}

// base[r[q] * W + j] = v[q][j] for the lane's active rows.  Two rows per lane and even M: the lane's block is 2W
// contiguous doubles starting on a 16-byte boundary -> W 16-byte stores (rows of width 3 would otherwise go out as
// 8-byte stores that replay on the banks).
template <int M, int R, int W>
__device__ __forceinline__ void store_rows(double* base, const RowIdx<M, R>& rw, const double (&v)[R][W]) {
  if constexpr (R == 2 && (M % 2) == 0) {
    if (rw.a[0]) {
      double2* d = reinterpret_cast<double2*>(base + rw.r[0] * W);
#pragma unroll
      for (int e = 0; e < W; ++e) d[e] = make_double2(v[(2 * e) / W][(2 * e) % W], v[(2 * e + 1) / W][(2 * e + 1) % W]);
    }
  } else {
#pragma unroll
    for (int q = 0; q < R; ++q)
      if (rw.a[q]) {
#pragma unroll
        for (int j = 0; j < W; ++j) base[rw.r[q] * W + j] = v[q][j];
      }
  }
}

// ------------------------------------------------------------------------------------------------ shared pieces
// Phases A, B of a step for an observed row: v, Mm | TM, F, F^-1, w, quad, Kp, Lm.   Leaves in shared memory: Mm, Kp, Lm;
// in registers of the lane: its TM, Kp, Lm rows, and (every lane, redundantly: p <= 3) v, F^-1, w.  v and F cost every
// lane 18 + 54 multiply-adds at m = 6, p = 3, but their operands (Z, a, Mm) are broadcast loads the phase needs anyway,
// where rows of v / F owned by lanes 0..p-1 cost 8-way replayed row loads of Z, a store / sync / load round trip and a
// third warp synchronisation.
template <int M, int P, int R>
struct RowGain {
  double TM[R][P], Kp[R][P], Lm[R][M], Fi[P * P], v[P], w[P], piv[P], quad;
  double Lc[P * P], Lic[P * P];  // MK_CHOLS only: Cholesky factor of F and its inverse (dead registers elsewhere)
  bool ok;
};

// tr(q, k) = T[row q of this lane][k] (registers in the forward kernel, shared memory in the adjoint)
// Pr = the lane's rows of the predicted covariance (registers).
// MK_STEADY: the gain matrix is the fixed Gss = (Z Pss Z^T + H)^-1 instead of F^-1 (F is still factorised for log det).
template <int M, int P, int R, int MK, class TR>
__device__ __forceinline__ void rows_gain(double* sm, const double* yt, double d_sign, const double (&dv)[P], unsigned mask,
                                          const RowIdx<M, R>& rw, TR tr, const double (&Pr)[R][M], const double (&Gss)[P * P],
                                          RowGain<M, P, R>& g, bool full_det = false) {
  using L = RowsLayout<M, P>;
  // ---- phase A: v (every lane), Mm rows
  {
    double av[M], Mr[R][P];
#pragma unroll
    for (int k = 0; k < M; ++k) av[k] = sm[L::a + k];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      double s[R], sv = yt[j] - d_sign * dv[j];
#pragma unroll
      for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const double z = sm[L::Z + j * M + k];
        sv = fma(-z, av[k], sv);
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = fma(Pr[q][k], z, s[q]);
      }
      g.v[j] = sv;
#pragma unroll
      for (int q = 0; q < R; ++q) Mr[q][j] = s[q];
    }
    store_rows<M, R, P>(sm + L::Mm, rw, Mr);
  }
  __syncwarp(mask);
  // ---- phase B: TM rows, F (every lane), F^-1, w, quad, Kp, Lm rows
  double Fr[P * P], Lr[P * P], Lir[P * P];
#pragma unroll
  for (int k = 0; k < P * P; ++k) Fr[k] = sm[L::H + k];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double s[R];
#pragma unroll
    for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double b = sm[L::Mm + k * P + j];
#pragma unroll
      for (int q = 0; q < R; ++q) s[q] = fma(tr(q, k), b, s[q]);
#pragma unroll
      for (int i = 0; i < P; ++i) Fr[i * P + j] = fma(sm[L::Z + i * M + k], b, Fr[i * P + j]);
    }
#pragma unroll
    for (int q = 0; q < R; ++q) g.TM[q][j] = s[q];
  }
  if (MK == MK_CHOLS) {
    // the as-coded CholeskyFilter for k_endog > 1: gain matrix Gk[k][i] = Li[i][k] / L_ii, w = Gk^T v
    g.ok = chol_factor_piv(Fr, g.Lc, g.Lic, g.piv, P);
#pragma unroll
    for (int k = 0; k < P; ++k)
#pragma unroll
      for (int i = 0; i < P; ++i) g.Fi[k * P + i] = g.Lic[i * P + k] / g.Lc[i * P + i];
  } else {
    g.ok = ldl_inverse(Fr, g.Fi, Lr, Lir, g.piv, P);
    if (MK == MK_STD && P > 1 && full_det && g.ok) g.ok = lu_pivots(Fr, Lr, g.piv, P);  // det of the full matrix (t = 0)
  }
  double qd = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k)
      s = fma(MK == MK_STEADY ? Gss[j * P + k] : MK == MK_CHOLS ? g.Fi[k * P + j] : g.Fi[j * P + k], g.v[k], s);
    g.w[j] = s;
    qd = fma(g.v[j], s, qd);
  }
  g.quad = qd;
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int j = 0; j < P; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < P; ++k) s = fma(g.TM[q][k], MK == MK_STEADY ? Gss[k * P + j] : g.Fi[k * P + j], s);
      g.Kp[q][j] = s;
    }
#pragma unroll
  for (int j = 0; j < M; ++j) {
    double s[R];
#pragma unroll
    for (int q = 0; q < R; ++q) s[q] = tr(q, j);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const double z = sm[L::Z + k * M + j];
#pragma unroll
      for (int q = 0; q < R; ++q) s[q] = fma(-g.Kp[q][k], z, s[q]);
    }
#pragma unroll
    for (int q = 0; q < R; ++q) g.Lm[q][j] = s[q];
  }
  store_rows<M, R, P>(sm + L::Kp, rw, g.Kp);
  store_rows<M, R, M>(sm + L::Lm, rw, g.Lm);
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
template <int M, int P, int G, int MK>
__device__ void rows_forward(const KfArgs& A, long long u, double* sm, int l, unsigned mask) {
  using L = RowsLayout<M, P>;
  constexpr int R = RowsCfg<M, P, G>::R;
  constexpr int KT = L::KT;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const RowIdx<M, R> rw(l);
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  const double* Cp = A.C.p + draw * A.C.bs;
  const double* P0p = (MK == MK_STEADY) ? A.Pss.p + draw * A.Pss.bs : A.P0.p + draw * A.P0.bs;  // steady: starts at Pss
  const double* a0p = A.a0.p + draw * A.a0.bs;
  double Gss[P * P];
#pragma unroll
  for (int k = 0; k < P * P; ++k) Gss[k] = (MK == MK_STEADY) ? A.Gss.p[draw * A.Gss.bs + k] : 0.0;
  if (l < G) {  // idle lanes of the warp (l == G) shadow a unit without owning anything
    for (int k = l; k < M * M; k += G) {
      sm[L::T + k] = Tp[k];
      sm[L::Pm + k] = P0p[k];
    }
    for (int k = l; k < P * M; k += G) sm[L::Z + k] = Zp[k];
    for (int k = l; k < P * P; k += G) sm[L::H + k] = Hp[k];
  }
  // static rows the lane owns: registers
  double Tr[R][M], Cr[R][M], Pr[R][M], ci[R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
      Tr[q][j] = Tp[rw.c[q] * M + j];
      Cr[q][j] = Cp[rw.c[q] * M + j];
      Pr[q][j] = P0p[rw.c[q] * M + j];
    }
    ci[q] = (rw.a[q] && A.c.p) ? A.c.p[draw * A.c.bs + rw.r[q]] : 0.0;
    if (rw.a[q]) sm[L::a + rw.r[q]] = a0p[rw.r[q]];
  }
  double dv[P];
#pragma unroll
  for (int j = 0; j < P; ++j) dv[j] = A.d.p ? A.d.p[draw * A.d.bs + j] : 0.0;
  __syncwarp(mask);
  auto tr = [&](int q, int k) { return Tr[q][k]; };

  const double* y = A.y.p;
  LogAcc acc;
  double llsum = 0.0;
  int info = 0;
  double* tp = A.tape ? A.tape + u * (long long)(n - 1) * KT : nullptr;
  RowGain<M, P, R> g;

  double yt[P], ynx[P];  // y[t] and, loaded one step ahead so its latency hides behind a whole step, y[t+1]
#pragma unroll
  for (int j = 0; j < P; ++j) ynx[j] = y[j];
  for (int t = 0; t < n; ++t) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      yt[j] = ynx[j];
      ynx[j] = y[(long long)(t + 1 < n ? t + 1 : t) * P + j];
    }
    const int nm = rows_count_missing<P>(yt);
    double an[R], S1[R][M], S2[R][M];
#pragma unroll
    for (int q = 0; q < R; ++q) an[q] = ci[q];
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double ak = sm[L::a + k];
#pragma unroll
      for (int q = 0; q < R; ++q) an[q] = fma(Tr[q][k], ak, an[q]);
    }
    if (nm == 0) {
      rows_gain<M, P, R, MK>(sm, yt, A.d_sign, dv, mask, rw, tr, Pr, Gss, g, t == 0);
      if (!g.ok && info == 0) info = t + 1;
      if (g.ok) {
#pragma unroll
        for (int k = 0; k < P; ++k) acc.mul(g.piv[k]);
      }
      llsum += -0.5 * (A.ll_const + g.quad);
      // ---- phase E: a' rows, S2 rows = C + (L P) L^T + (Kp H) Kp^T
#pragma unroll
      for (int k = 0; k < P; ++k) {
#pragma unroll
        for (int q = 0; q < R; ++q) an[q] = fma(g.Kp[q][k], g.v[k], an[q]);
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::Pm + k * M + j];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(g.Lm[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) S1[q][j] = s[q];
      }
      double KH[R][P];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const double b = sm[L::H + k * P + j];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(g.Kp[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) KH[q][j] = s[q];
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = Cr[q][j];
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::Lm + j * M + k];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(S1[q][k], b, s[q]);
        }
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const double b = sm[L::Kp + j * P + k];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(KH[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) S2[q][j] = s[q];
      }
    } else {
      if (nm != P && info == 0) info = -(t + 1);
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::Pm + k * M + j];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(Tr[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) S1[q][j] = s[q];
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = Cr[q][j];
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::T + j * M + k];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(S1[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) S2[q][j] = s[q];
      }
    }
    store_rows<M, R, M>(sm + L::S2, rw, S2);
    __syncwarp(mask);
    // ---- phase F: P' = sym(S2), a' ; tape
    const bool taped = tp && t + 1 < n;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double col[R];
      lane_block<M, R>(sm + L::S2 + j * M, rw, col);
#pragma unroll
      for (int q = 0; q < R; ++q) Pr[q][j] = 0.5 * (S2[q][j] + col[q]);
    }
#pragma unroll
    for (int q = 0; q < R; ++q)
      if (rw.a[q]) {
        const int i = rw.r[q];
        if (taped) tp[i] = an[q];
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (taped && j >= i) tp[M + i * M - (i * (i - 1)) / 2 + (j - i)] = Pr[q][j];
        sm[L::a + i] = an[q];
      }
    store_rows<M, R, M>(sm + L::Pm, rw, Pr);
    if (taped) tp += KT;
    __syncwarp(mask);
  }
  if (l == 0) {
    double ll = llsum - 0.5 * acc.value();
    if (info != 0) ll = nan("");
    if (A.loglik) A.loglik[u] = ll;
    if (MK == MK_STEADY && A.dare_info && A.dare_info[u / A.n_series] != 0) info = KF_INFO_DARE_FAILED;
    if (A.info) A.info[u] = info;
  }
}

// ------------------------------------------------------------------------------------------------ forward, all outputs
// The reference's six outputs (kalman_filter.py:166-193) on the same mapping: two-stage form
//     A = I - K Z,  P_f = A P A^T + K H K^T,  a_f = a + K v ;  P' = sym(T P_f T^T + C),  a' = T a_f + c
// (rows_gain with T := I yields the filter gain K and A).  MK_STD / MK_STEADY.  Rows go out as 16-byte stores, a unit's
// matrix as one contiguous run.  (Before: the generic sub-warp kernels, 0.73 TB/s written at k_states 6.)
template <int M, int P, int G, int MK>
__device__ void rows_forward_full(const KfArgs& A, long long u, double* sm, int l, unsigned mask) {
  using L = RowsLayout<M, P>;
  constexpr int R = RowsCfg<M, P, G>::R;
  constexpr int KT = L::KT;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const RowIdx<M, R> rw(l);
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  const double* Cp = A.C.p + draw * A.C.bs;
  const double* P0g = A.P0.p + draw * A.P0.bs;
  const double* P0p = (MK == MK_STEADY) ? A.Pss.p + draw * A.Pss.bs : P0g;  // steady: the recursion starts at Pss
  const double* a0p = A.a0.p + draw * A.a0.bs;
  double Gss[P * P];
#pragma unroll
  for (int k = 0; k < P * P; ++k) Gss[k] = (MK == MK_STEADY) ? A.Gss.p[draw * A.Gss.bs + k] : 0.0;
  if (l < G) {
    for (int k = l; k < M * M; k += G) {
      sm[L::T + k] = Tp[k];
      sm[L::Pm + k] = P0p[k];
    }
    for (int k = l; k < P * M; k += G) sm[L::Z + k] = Zp[k];
    for (int k = l; k < P * P; k += G) sm[L::H + k] = Hp[k];
  }
  double Tr[R][M], Cr[R][M], Pr[R][M], ci[R], af[R];
  {
    double r0[R][M];  // row 0 of the predicted moments = the caller's a0 / P0 (steady state: reported, not used, :397)
#pragma unroll
    for (int q = 0; q < R; ++q) {
#pragma unroll
      for (int j = 0; j < M; ++j) {
        Tr[q][j] = Tp[rw.c[q] * M + j];
        Cr[q][j] = Cp[rw.c[q] * M + j];
        Pr[q][j] = P0p[rw.c[q] * M + j];
        r0[q][j] = P0g[rw.c[q] * M + j];
      }
      ci[q] = (rw.a[q] && A.c.p) ? A.c.p[draw * A.c.bs + rw.r[q]] : 0.0;
      af[q] = a0p[rw.c[q]];
      if (rw.a[q]) sm[L::a + rw.r[q]] = af[q];
    }
    if (l < G) {
      if (A.ps) store_block<M, R>(A.ps + u * (long long)(n + 1) * M, rw, af);
      if (A.pc) store_rows<M, R, M>(A.pc + u * (long long)(n + 1) * M * M, rw, r0);
    }
  }
  double dv[P];
#pragma unroll
  for (int j = 0; j < P; ++j) dv[j] = A.d.p ? A.d.p[draw * A.d.bs + j] : 0.0;
  __syncwarp(mask);
  auto idt = [&](int q, int k) { return rw.c[q] == k ? 1.0 : 0.0; };

  const double* y = A.y.p;
  double llsum = 0.0;
  int info = 0;
  double* tp = A.tape ? A.tape + u * (long long)(n - 1) * KT : nullptr;
  RowGain<M, P, R> g;
  double yt[P], ynx[P];
#pragma unroll
  for (int j = 0; j < P; ++j) ynx[j] = y[j];
  for (int t = 0; t < n; ++t) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      yt[j] = ynx[j];
      ynx[j] = y[(long long)(t + 1 < n ? t + 1 : t) * P + j];
    }
    const int nm = rows_count_missing<P>(yt);
    double S1[R][M], S2[R][M], ll = 0.0;
#pragma unroll
    for (int q = 0; q < R; ++q) af[q] = sm[L::a + rw.c[q]];
    if (nm == 0) {
      rows_gain<M, P, R, MK>(sm, yt, A.d_sign, dv, mask, rw, idt, Pr, Gss, g, t == 0);  // K (in Kp), A = I - K Z (in Lm)
      if (!g.ok && info == 0) info = t + 1;
      double ld = 0.0;
#pragma unroll
      for (int k = 0; k < P; ++k) ld += log(g.piv[k]);
      ll = g.ok ? -0.5 * (A.ll_const + ld + g.quad) : nan("");
#pragma unroll
      for (int k = 0; k < P; ++k) {
#pragma unroll
        for (int q = 0; q < R; ++q) af[q] = fma(g.Kp[q][k], g.v[k], af[q]);
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {  // S1 = A P
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::Pm + k * M + j];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(g.Lm[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) S1[q][j] = s[q];
      }
      double KH[R][P];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const double b = sm[L::H + k * P + j];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(g.Kp[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) KH[q][j] = s[q];
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {  // P_f = S1 A^T + (K H) K^T
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::Lm + j * M + k];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(S1[q][k], b, s[q]);
        }
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const double b = sm[L::Kp + j * P + k];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(KH[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) Pr[q][j] = s[q];
      }
      __syncwarp(mask);  // every lane has read P (S1) and A, K: P_f may replace P
      store_rows<M, R, M>(sm + L::Pm, rw, Pr);
    } else if (nm != P && info == 0) {
      info = -(t + 1);
    }
    llsum += ll;
    if (l < G) {
      if (A.fs) store_block<M, R>(A.fs + (u * (long long)n + t) * M, rw, af);
      if (A.fc) store_rows<M, R, M>(A.fc + (u * (long long)n + t) * M * M, rw, Pr);
    }
    store_block<M, R>(sm + L::a, rw, af);
    if (l == 0 && A.ll_obs) A.ll_obs[u * (long long)n + t] = ll;
    __syncwarp(mask);  // a_f and P_f visible
    // ---- predict: a' = T a_f + c ; P' = sym(T P_f T^T + C)
    double an[R];
#pragma unroll
    for (int q = 0; q < R; ++q) an[q] = ci[q];
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double ak = sm[L::a + k];
#pragma unroll
      for (int q = 0; q < R; ++q) an[q] = fma(Tr[q][k], ak, an[q]);
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s[R];
#pragma unroll
      for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const double b = sm[L::Pm + k * M + j];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = fma(Tr[q][k], b, s[q]);
      }
#pragma unroll
      for (int q = 0; q < R; ++q) S1[q][j] = s[q];
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s[R];
#pragma unroll
      for (int q = 0; q < R; ++q) s[q] = Cr[q][j];
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const double b = sm[L::T + j * M + k];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = fma(S1[q][k], b, s[q]);
      }
#pragma unroll
      for (int q = 0; q < R; ++q) S2[q][j] = s[q];
    }
    store_rows<M, R, M>(sm + L::S2, rw, S2);
    __syncwarp(mask);
    const bool taped = tp && t + 1 < n;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double col[R];
      lane_block<M, R>(sm + L::S2 + j * M, rw, col);
#pragma unroll
      for (int q = 0; q < R; ++q) Pr[q][j] = 0.5 * (S2[q][j] + col[q]);
    }
#pragma unroll
    for (int q = 0; q < R; ++q)
      if (rw.a[q]) {
        const int i = rw.r[q];
        if (taped) tp[i] = an[q];
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (taped && j >= i) tp[M + i * M - (i * (i - 1)) / 2 + (j - i)] = Pr[q][j];
      }
    store_block<M, R>(sm + L::a, rw, an);
    store_rows<M, R, M>(sm + L::Pm, rw, Pr);
    if (l < G) {
      if (A.ps) store_block<M, R>(A.ps + (u * (long long)(n + 1) + t + 1) * M, rw, an);
      if (A.pc) store_rows<M, R, M>(A.pc + (u * (long long)(n + 1) + t + 1) * M * M, rw, Pr);
    }
    if (taped) tp += KT;
    __syncwarp(mask);
  }
  if (l == 0) {
    if (info != 0) llsum = nan("");
    if (A.loglik) A.loglik[u] = llsum;
    if (MK == MK_STEADY && A.dare_info && A.dare_info[u / A.n_series] != 0) info = KF_INFO_DARE_FAILED;
    if (A.info) A.info[u] = info;
  }
}


// ------------------------------------------------------------------------------------------------ adjoint
// Asynchronous copy of one packed tape entry (KT doubles) into shared memory, G lanes cooperating (LDGSTS).
template <int KT, int G>
__device__ __forceinline__ void rows_tape_prefetch(double* dst, const double* src, int l) {
#ifdef __CUDA_ARCH__
  if (l < G) {
    const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll
    for (int k0 = 0; k0 < KT; k0 += G) {
      const int k = k0 + l;
      if (k < KT) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + (unsigned)(k * 8)), "l"(src + k) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
#else
  if (l < G)
    for (int k = l; k < KT; k += G) dst[k] = src[k];
#endif
}

__device__ __forceinline__ void rows_tape_wait() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// NEED_Z: the caller wants Z-bar (never the case for the reference's models, whose design matrix is constant): the
// Z-bar accumulators and the Lb / Mb exchanges exist only in that instantiation.
template <int M, int P, int G, int MK, bool NEED_Z>
__device__ void rows_backward(const KfArgs& A, long long u, double* sm, int l, unsigned mask) {
  using L = RowsLayout<M, P, NEED_Z, MK == MK_STEADY || MK == MK_CHOLS>;
  constexpr int R = RowsCfg<M, P, G>::R;
  constexpr int KT = L::KT;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const RowIdx<M, R> rw(l);
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  const double* tape = A.tape + u * (long long)(n - 1) * KT;  // entry t-1 = predicted moments of step t
  if (n >= 2) rows_tape_prefetch<KT, G>(sm + L::tp + ((n - 1) & 1) * L::KTP, tape + (long long)(n - 2) * KT, l);
  if (l < G) {
    for (int k = l; k < M * M; k += G) {
      sm[L::T + k] = Tp[k];
      sm[L::Pb + k] = 0.0;
      sm[L::Cb + k] = 0.0;
    }
    for (int k = l; k < P * M; k += G) sm[L::Z + k] = Zp[k];
    for (int k = l; k < P * P; k += G) sm[L::H + k] = Hp[k];
  }
#pragma unroll
  for (int q = 0; q < R; ++q)
    if (rw.a[q]) sm[L::ab + rw.r[q]] = 0.0;
  double dv[P];
#pragma unroll
  for (int j = 0; j < P; ++j) dv[j] = A.d.p ? A.d.p[draw * A.d.bs + j] : 0.0;
  __syncwarp(mask);
  auto tr = [&](int q, int k) { return sm[L::T + rw.c[q] * M + k]; };

  const double* y = A.y.p;
  const double gl = A.g_loglik ? A.g_loglik[u] : 1.0;
  constexpr bool need_Z = NEED_Z;
  const bool need_H = (A.gH != nullptr) || (MK == MK_STEADY);
  double Gss[P * P], Gb[P * P];  // steady state: fixed gain matrix and its cotangent (every lane holds all of it)
#pragma unroll
  for (int k = 0; k < P * P; ++k) {
    Gss[k] = (MK == MK_STEADY) ? A.Gss.p[draw * A.Gss.bs + k] : 0.0;
    Gb[k] = 0.0;
  }
  // gradient accumulators: the lane's rows of Tb (registers) and Cb (shared memory: touched once per step, and the
  // registers are all taken); lanes < P hold rows of Zb, Hb; elements of cb (rows), db (lane)
  double Tb[R][M], Zb[NEED_Z ? M : 1], Hb[P], cb[R], db = 0.0;
#pragma unroll
  for (int q = 0; q < R; ++q) {
    cb[q] = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) Tb[q][j] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < (NEED_Z ? M : 1); ++j) Zb[j] = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) Hb[j] = 0.0;
  RowGain<M, P, R> g;
  double Pr[R][M];
  double Pbr[R][M], abi[R];  // the lane's rows of Pb / elements of ab as of the previous step (their columns go via shared memory)
#pragma unroll
  for (int q = 0; q < R; ++q) {
    abi[q] = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) Pbr[q][j] = 0.0;
  }
  double yt[P], ynx[P];  // y[t] and, loaded one step ahead, y[t-1]
#pragma unroll
  for (int j = 0; j < P; ++j) ynx[j] = y[(long long)(n - 1) * P + j];

  for (int t = n - 1; t >= 0; --t) {
    // ---- predicted moments of step t -> shared memory (and the lane's rows of P in registers)
    if (t == 0) {
      const double* P0p = (MK == MK_STEADY) ? A.Pss.p + draw * A.Pss.bs : A.P0.p + draw * A.P0.bs;
#pragma unroll
      for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int j = 0; j < M; ++j) Pr[q][j] = P0p[rw.c[q] * M + j];
        if (rw.a[q]) sm[L::a + rw.r[q]] = A.a0.p[draw * A.a0.bs + rw.r[q]];
      }
      store_rows<M, R, M>(sm + L::Pm, rw, Pr);
    } else {
      rows_tape_wait();
      __syncwarp(mask);
      const double* tq = sm + L::tp + (t & 1) * L::KTP;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int i = rw.c[q];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          const int lo = i < j ? i : j, hi = i < j ? j : i;
          Pr[q][j] = tq[M + lo * M - (lo * (lo - 1)) / 2 + (hi - lo)];
        }
        if (rw.a[q]) sm[L::a + i] = tq[i];
      }
      store_rows<M, R, M>(sm + L::Pm, rw, Pr);
      if (t >= 2) rows_tape_prefetch<KT, G>(sm + L::tp + ((t - 1) & 1) * L::KTP, tape + (long long)(t - 2) * KT, l);
    }
    __syncwarp(mask);
#pragma unroll
    for (int j = 0; j < P; ++j) {
      yt[j] = ynx[j];
      ynx[j] = y[(long long)(t > 0 ? t - 1 : 0) * P + j];
    }
    const double lb = gl + (A.g_ll_obs ? A.g_ll_obs[u * n + t] : 0.0);
    const bool observed = (rows_count_missing<P>(yt) == 0);
    if (observed) {
      rows_gain<M, P, R, MK>(sm, yt, A.d_sign, dv, mask, rw, tr, Pr, Gss, g);
      if (MK == MK_STEADY || MK == MK_CHOLS) store_rows<M, R, P>(sm + L::TMs, rw, g.TM);  // read by every lane in phase 3 (after syncs)
    } else {
#pragma unroll
      for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int j = 0; j < M; ++j) g.Lm[q][j] = sm[L::T + rw.c[q] * M + j];
      }
      store_rows<M, R, M>(sm + L::Lm, rw, g.Lm);
      __syncwarp(mask);
    }
    // ---- phase 1: Ps rows (registers) ; Cb, cb
    double Ps[R][M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double col[R];
      lane_block<M, R>(sm + L::Pb + j * M, rw, col);
#pragma unroll
      for (int q = 0; q < R; ++q) Ps[q][j] = 0.5 * (Pbr[q][j] + col[q]);
    }
    accumulate_rows<M, R, M>(sm + L::Cb, rw, Ps);
#pragma unroll
    for (int q = 0; q < R; ++q) cb[q] += abi[q];
    // ---- phase 2: W = Ps L, Lb = Ps L (P + P^T) = W (P + P^T) from the lane's own W rows (round 1: X = L (P + P^T) rows
    //      through shared memory, then Ps X: one m^3 product, one row store and one sync more per step), T^T ab,
    //      (observed) PK = Ps Kp, Kb
    double Lb[R][M], PK[R][P], Kb[R][P], abn[R];
    {
    double Wr[R][M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s2[R];
#pragma unroll
      for (int q = 0; q < R; ++q) s2[q] = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const double bl = sm[L::Lm + k * M + j];
#pragma unroll
        for (int q = 0; q < R; ++q) s2[q] = fma(Ps[q][k], bl, s2[q]);
      }
#pragma unroll
      for (int q = 0; q < R; ++q) Wr[q][j] = s2[q];
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s[R];
#pragma unroll
      for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const double b = sm[L::Pm + k * M + j] + sm[L::Pm + j * M + k];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = fma(Wr[q][k], b, s[q]);
      }
      const double aj = sm[L::a + j];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        Lb[q][j] = s[q];
        Tb[q][j] += fma(abi[q], aj, s[q]);  // Tb += ab a^T + Lb
      }
    }
    store_rows<M, R, M>(sm + L::W, rw, Wr);
    }
#pragma unroll
    for (int q = 0; q < R; ++q) abn[q] = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double abk = sm[L::ab + k];
      double tc[R];
      lane_block<M, R>(sm + L::T + k * M, rw, tc);
#pragma unroll
      for (int q = 0; q < R; ++q) abn[q] = fma(tc[q], abk, abn[q]);
    }
    if (observed) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::Kp + k * P + j];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(Ps[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) PK[q][j] = s[q];
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = abi[q] * g.v[j];
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const double b = sm[L::H + k * P + j] + sm[L::H + j * P + k];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(PK[q][k], b, s[q]);
        }
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double z = sm[L::Z + j * M + k];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(-Lb[q][k], z, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) Kb[q][j] = s[q];
      }
      store_rows<M, R, P>(sm + L::Kb, rw, Kb);
      if (need_H) store_rows<M, R, P>(sm + L::PK, rw, PK);
      if (need_Z) store_rows<M, R, M>(sm + L::Lb, rw, Lb);
    }
    __syncwarp(mask);
    // ---- phase 3: Pb' = L^T W (registers) ; (observed) every lane: K^T Kb, vb, Fb ; TMb rows, Tb += TMb Mm^T
    double Pbn[R][M];
    {
      double Lc[R][M];  // the lane's COLUMNS of L
#pragma unroll
      for (int k = 0; k < M; ++k) {
        double col[R];
        lane_block<M, R>(sm + L::Lm + k * M, rw, col);
#pragma unroll
        for (int q = 0; q < R; ++q) Lc[q][k] = col[q];
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::W + k * M + j];
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(Lc[q][k], b, s[q]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) Pbn[q][j] = s[q];
      }
    }
    double vb[P], Fb[P * P];
    if (observed) {
      if (MK == MK_STEADY) {
        // vb = Kp^T ab - lb sym(Gss) v ; Gss-bar += TM^T Kb - lb/2 v v^T ; Fb = -lb/2 F^-T
#pragma unroll
        for (int a2 = 0; a2 < P; ++a2) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < M; ++k) s = fma(sm[L::Kp + k * P + a2], sm[L::ab + k], s);
#pragma unroll
          for (int b2 = 0; b2 < P; ++b2) s = fma(-0.5 * lb * (Gss[a2 * P + b2] + Gss[b2 * P + a2]), g.v[b2], s);
          vb[a2] = s;
#pragma unroll
          for (int b2 = 0; b2 < P; ++b2) {
            double q1 = Gb[a2 * P + b2];
#pragma unroll
            for (int k = 0; k < M; ++k) q1 = fma(sm[L::TMs + k * P + a2], sm[L::Kb + k * P + b2], q1);
            Gb[a2 * P + b2] = fma(-0.5 * lb * g.v[a2], g.v[b2], q1);
            Fb[a2 * P + b2] = -0.5 * lb * g.Fi[b2 * P + a2];
          }
        }
      } else if (MK == MK_CHOLS) {
        // vb = Kp^T ab - lb/2 (Gk + Gk^T) v ; Gk-bar = TM^T Kb - lb/2 v v^T ; Fb = adjoint of the as-coded gain matrix
        // and of - sum log L_ii through the Cholesky factor (chols_adjoint, kf_core.cuh)
        double Gkb[P * P], W1[P * P], W2[P * P];
#pragma unroll
        for (int a2 = 0; a2 < P; ++a2) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < M; ++k) s = fma(sm[L::Kp + k * P + a2], sm[L::ab + k], s);
#pragma unroll
          for (int b2 = 0; b2 < P; ++b2) s = fma(-0.5 * lb * (g.Fi[a2 * P + b2] + g.Fi[b2 * P + a2]), g.v[b2], s);
          vb[a2] = s;
#pragma unroll
          for (int b2 = 0; b2 < P; ++b2) {
            double q1 = 0.0;
#pragma unroll
            for (int k = 0; k < M; ++k) q1 = fma(sm[L::TMs + k * P + a2], sm[L::Kb + k * P + b2], q1);
            Gkb[a2 * P + b2] = fma(-0.5 * lb * g.v[a2], g.v[b2], q1);
          }
        }
        chols_adjoint(Gkb, g.Lc, g.Lic, g.Fi, lb, W1, W2, Fb, P);
      } else {
        double Q1[P * P];
#pragma unroll
        for (int a2 = 0; a2 < P; ++a2) {
          double s = -lb * g.w[a2];
#pragma unroll
          for (int k = 0; k < M; ++k) s = fma(sm[L::Kp + k * P + a2], sm[L::ab + k], s);
          vb[a2] = s;
#pragma unroll
          for (int b2 = 0; b2 < P; ++b2) {
            double q1 = 0.0;
#pragma unroll
            for (int k = 0; k < M; ++k) q1 = fma(sm[L::Kp + k * P + a2], sm[L::Kb + k * P + b2], q1);
            Q1[a2 * P + b2] = q1;
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
      }
      double TMb[R][P];  // TMb = Kb G^T, G = F^-1 or Gss
#pragma unroll
      for (int q = 0; q < R; ++q)
#pragma unroll
        for (int j = 0; j < P; ++j) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < P; ++k) s = fma(Kb[q][k], MK == MK_STEADY ? Gss[j * P + k] : g.Fi[j * P + k], s);
          TMb[q][j] = s;
        }
#pragma unroll
      for (int j = 0; j < M; ++j) {
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const double b = sm[L::Mm + j * P + k];
#pragma unroll
          for (int q = 0; q < R; ++q) Tb[q][j] = fma(TMb[q][k], b, Tb[q][j]);
        }
      }
      store_rows<M, R, P>(sm + L::TMb, rw, TMb);
    }
    __syncwarp(mask);
    // ---- phase 4: (observed) Mb = T^T TMb + Z^T Fb ; Pb' += Mb Z ; ab' = T^T ab - Z^T vb ; store Pb', ab'
    if (observed) {
      double Mb[R][P], Zc[R][P];  // Zc: the lane's columns of Z
#pragma unroll
      for (int k = 0; k < P; ++k) {
        double col[R];
        lane_block<M, R>(sm + L::Z + k * M, rw, col);
#pragma unroll
        for (int q = 0; q < R; ++q) Zc[q][k] = col[q];
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        double s[R];
#pragma unroll
        for (int q = 0; q < R; ++q) s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const double b = sm[L::TMb + k * P + j];
          double tc[R];
          lane_block<M, R>(sm + L::T + k * M, rw, tc);
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(tc[q], b, s[q]);
        }
#pragma unroll
        for (int k = 0; k < P; ++k)
#pragma unroll
          for (int q = 0; q < R; ++q) s[q] = fma(Zc[q][k], Fb[k * P + j], s[q]);
#pragma unroll
        for (int q = 0; q < R; ++q) Mb[q][j] = s[q];
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const double z = sm[L::Z + k * M + j];
#pragma unroll
          for (int q = 0; q < R; ++q) Pbn[q][j] = fma(Mb[q][k], z, Pbn[q][j]);
        }
      }
#pragma unroll
      for (int k = 0; k < P; ++k)
#pragma unroll
        for (int q = 0; q < R; ++q) abn[q] = fma(-Zc[q][k], vb[k], abn[q]);
#pragma unroll
      for (int k = 0; k < P; ++k)
        if (l == k) db = fma(-A.d_sign, vb[k], db);
      if (need_Z) store_rows<M, R, P>(sm + L::Mb, rw, Mb);
    }
    __syncwarp(mask);  // every lane has finished reading Pb, ab, Lm, W of this step
    store_rows<M, R, M>(sm + L::Pb, rw, Pbn);
    store_block<M, R>(sm + L::ab, rw, abn);
#pragma unroll
    for (int q = 0; q < R; ++q) {
      abi[q] = abn[q];
#pragma unroll
      for (int j = 0; j < M; ++j) Pbr[q][j] = Pbn[q][j];
    }
    // ---- optional cotangents that need cross-row reductions (lanes < P own the rows of Zb, Hb)
    if (observed && (need_Z || need_H)) {
#pragma unroll
      for (int e = 0; e < P; ++e) {  // static index e instead of vb[l] / Fb[l * P + k]: keeps both in registers
        if (l != e) continue;
        if constexpr (NEED_Z) {
#pragma unroll
          for (int j = 0; j < M; ++j) {
            double s = fma(-vb[e], sm[L::a + j], Zb[j]);
#pragma unroll
            for (int k = 0; k < M; ++k) {
              s = fma(-sm[L::Kp + k * P + e], sm[L::Lb + k * M + j], s);  // - Kp^T Lb
              s = fma(sm[L::Mb + k * P + e], sm[L::Pm + k * M + j], s);   // + Mb^T P
            }
#pragma unroll
            for (int k = 0; k < P; ++k) s = fma(Fb[e * P + k], sm[L::Mm + j * P + k], s);  // + Fb Mm^T
            Zb[j] = s;
          }
        }
        if (need_H) {
#pragma unroll
          for (int j = 0; j < P; ++j) {
            double s = Hb[j] + Fb[e * P + j];
#pragma unroll
            for (int k = 0; k < M; ++k) s = fma(sm[L::Kp + k * P + e], sm[L::PK + k * P + j], s);  // + Kp^T Ps Kp
            Hb[j] = s;
          }
        }
      }
    }
    __syncwarp(mask);
  }
  // ---- write-out (each lane its rows)
#pragma unroll
  for (int q = 0; q < R; ++q)
    if (rw.a[q]) {
      const int i = rw.r[q];
      if (A.ga0) A.ga0[u * M + i] = sm[L::ab + i];
#pragma unroll
      for (int j = 0; j < M; ++j) {
        if (MK == MK_STEADY) {
          if (A.gPss) A.gPss[u * M * M + i * M + j] = sm[L::Pb + i * M + j];
          if (A.gP0) A.gP0[u * M * M + i * M + j] = 0.0;
        } else if (A.gP0) {
          A.gP0[u * M * M + i * M + j] = sm[L::Pb + i * M + j];
        }
        if (A.gT) A.gT[u * M * M + i * M + j] = Tb[q][j];
        if (A.gC) A.gC[u * M * M + i * M + j] = sm[L::Cb + i * M + j];
      }
      if (A.gc) A.gc[u * M + i] = cb[q];
    }
  if (l < P) {
    if (A.gd) A.gd[u * P + l] = db;
#pragma unroll
    for (int j = 0; j < (NEED_Z ? M : 0); ++j) A.gZ[u * P * M + l * M + j] = Zb[j];
#pragma unroll
    for (int j = 0; j < P; ++j)
      if (A.gH) A.gH[u * P * P + l * P + j] = Hb[j];
  }
  if (MK == MK_STEADY && l == 0 && A.gGss) {
#pragma unroll
    for (int k = 0; k < P * P; ++k) A.gGss[u * P * P + k] = Gb[k];
  }
}

}  // namespace kfb
