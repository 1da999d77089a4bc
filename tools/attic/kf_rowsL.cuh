// kf_rowsL.cuh - fused "row per lane, warp per unit" programs for LARGE systems (16 < k_states <= 32, even; k_endog <= 3;
// MK_STD; static matrices; shared observation stream): config 4 of BASELINE.json (trend + seasonal, k_states = 30).
//
// Same mathematics as kf_pred.cuh / kf_rows.cuh.  Lane i owns row i of every m x m / m x p quantity.  A row of 30
// doubles is 60 registers, so - unlike kf_rows.cuh - only the ACCUMULATORS of a product live in registers:
//     acc[j] += a_k * B[k][j]    for k = 0..m-1,
// with B[k][:] read as broadcast 16-byte loads (one wavefront serves two multiply-adds of all 32 lanes: the shared-memory
// pipe and the fp64 pipe are balanced at 1 wavefront : 2 DFMA warp-instructions) and a_k, the lane's own element,
// fetched six at a time from shared memory (row source: three conflict-free 16-byte loads; column source - for
// products with L^T / T^T on the left - six 8-byte loads that are contiguous across lanes).  The k loop is NOT unrolled
// (a fully unrolled step would be ~10k instructions per product chain and thrash the instruction cache).
// The 2x2-register-tile CTA kernel this replaces (kf_coopT_kernel<30,1,256>) needed one shared-memory operand per
// multiply-add and ~40 CTA barriers per step: shared-memory bound at 24 % of the fp64 peak.
#pragma once
#include "kf_core.cuh"
#include "kf_rows.cuh"

namespace kfb {

// NEED_T: the caller wants T-bar.  Structural models have a constant T: then X = L (P + P^T) is never stored (only
// X Z^T is exchanged) and the Lb = Ps X product is skipped.
template <int M, int P, bool NEED_T>
struct RowsLLayout {
  static_assert(M % 2 == 0 && M <= 32 && P <= 3, "even k_states <= 32");
  static constexpr int MM = M * M, MP = M * P, PP = P * P, MPE = MP + (MP & 1), ME = M + (M & 1);
  static constexpr int KT = M + (M * (M + 1)) / 2, KTP = (KT + 1) & ~1;
  static constexpr int T = 0, Z = T + MM, H = Z + MPE, Pm = H + PP + (PP & 1), Mm = Pm + MM, Kp = Mm + MPE, Lm = Kp + MPE,
                       a = Lm + MM, END_COMMON = a + ME;
  // forward only: L^T (the right operand of X L^T must be row-contiguous) and X / S2 (own rows only)
  static constexpr int LmT = END_COMMON, X = LmT + MM, END_FWD = X + MM;
  // adjoint only: W lives in Pm's slot (nothing reads Pm after X is formed)
  static constexpr int Pb = END_COMMON, W = Pm, Kb = Pb + MM, TMb = Kb + MPE, PK = TMb + MPE, lz = PK + MPE, ab = lz + MPE,
                       tp = ab + ME, Xb = tp + KTP, END_BWD = Xb + (NEED_T ? MM : 0);
  static constexpr int fwd_doubles = (END_FWD + 1) & ~1, bwd_doubles = (END_BWD + 1) & ~1;
};

// left-operand sources for rowmul: KC consecutive elements of the lane's row / column of a row-major m x m matrix
template <int M>
struct RowSrc {
  const double* p;  // &A[i][0]
  template <int KC>
  __device__ __forceinline__ void load(int k0, double (&a)[KC]) const {
    const double2* r = reinterpret_cast<const double2*>(p + k0);
#pragma unroll
    for (int e = 0; e < KC / 2; ++e) {
      const double2 v = r[e];
      a[2 * e] = v.x;
      a[2 * e + 1] = v.y;
    }
  }
};
template <int M>
struct ColSrc {
  const double* p;  // &A[0][i]
  template <int KC>
  __device__ __forceinline__ void load(int k0, double (&a)[KC]) const {
#pragma unroll
    for (int e = 0; e < KC; ++e) a[e] = p[(k0 + e) * M];
  }
};

// acc[j] += sum_k a_k B[k][j]   (B row-major m x m in shared memory, 16-byte aligned rows)
template <int M, class SRC>
__device__ __forceinline__ void rowmul(double (&acc)[M], const SRC src, const double* B) {
  constexpr int KC = (M % 6 == 0) ? 6 : ((M % 4 == 0) ? 4 : 2);
#pragma unroll 1
  for (int k0 = 0; k0 < M; k0 += KC) {
    double a[KC];
    src.template load<KC>(k0, a);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const double2* b = reinterpret_cast<const double2*>(B + (k0 + kk) * M);
#pragma unroll
      for (int j = 0; j < M / 2; ++j) {
        const double2 bv = b[j];
        acc[2 * j] = fma(a[kk], bv.x, acc[2 * j]);
        acc[2 * j + 1] = fma(a[kk], bv.y, acc[2 * j + 1]);
      }
    }
  }
}

template <int M>
__device__ __forceinline__ void store_row(double* dst, const double (&v)[M]) {
  double2* d = reinterpret_cast<double2*>(dst);
#pragma unroll
  for (int j = 0; j < M / 2; ++j) d[j] = make_double2(v[2 * j], v[2 * j + 1]);
}

template <int M, int P>
struct RowLGain {
  double Kp[P], Fi[P * P], v[P], w[P], piv[P], quad;
  bool ok;
};

// v, Mm | TM, F, F^-1, w, quad, Kp, Lm for an observed step.  Leaves Mm, Kp, Lm in shared memory (visible after the
// trailing sync); returns the lane's Kp row and (every lane) v, F^-1, w.
template <int M, int P, class LAY>
__device__ __forceinline__ void rowsL_gain(double* sm, const double (&yt)[P], double d_sign, const double (&dv)[P], int i, bool act,
                                           unsigned mask, RowLGain<M, P>& g) {
  using L = LAY;
  // ---- A: Mm row (own P row x Z rows), v (every lane)
  {
    double Mr[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      Mr[j] = 0.0;
      g.v[j] = yt[j] - d_sign * dv[j];
    }
    const double2* pr = reinterpret_cast<const double2*>(sm + L::Pm + i * M);
    const double2* av = reinterpret_cast<const double2*>(sm + L::a);
#pragma unroll
    for (int k = 0; k < M / 2; ++k) {
      const double2 pk = pr[k], ak = av[k];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const double2 z = *reinterpret_cast<const double2*>(sm + L::Z + j * M + 2 * k);
        Mr[j] = fma(pk.x, z.x, Mr[j]);
        Mr[j] = fma(pk.y, z.y, Mr[j]);
        g.v[j] = fma(-z.x, ak.x, g.v[j]);
        g.v[j] = fma(-z.y, ak.y, g.v[j]);
      }
    }
    if (act) {
#pragma unroll
      for (int j = 0; j < P; ++j) sm[L::Mm + i * P + j] = Mr[j];
    }
  }
  __syncwarp(mask);
  // ---- B: TM row, F (every lane), inverse, w, quad, Kp row, Lm row
  double Fr[P * P], Lr[P * P], Lir[P * P], TM[P];
#pragma unroll
  for (int k = 0; k < P * P; ++k) Fr[k] = sm[L::H + k];
#pragma unroll
  for (int j = 0; j < P; ++j) TM[j] = 0.0;
  {
    const double2* tr = reinterpret_cast<const double2*>(sm + L::T + i * M);
#pragma unroll
    for (int k = 0; k < M / 2; ++k) {
      const double2 tk = tr[k];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const double m0 = sm[L::Mm + (2 * k) * P + j], m1 = sm[L::Mm + (2 * k + 1) * P + j];
        TM[j] = fma(tk.x, m0, TM[j]);
        TM[j] = fma(tk.y, m1, TM[j]);
#pragma unroll
        for (int e = 0; e < P; ++e) {
          const double2 z = *reinterpret_cast<const double2*>(sm + L::Z + e * M + 2 * k);
          Fr[e * P + j] = fma(z.x, m0, Fr[e * P + j]);
          Fr[e * P + j] = fma(z.y, m1, Fr[e * P + j]);
        }
      }
    }
  }
  g.ok = ldl_inverse(Fr, g.Fi, Lr, Lir, g.piv, P);
  double qd = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) s = fma(g.Fi[j * P + k], g.v[k], s);
    g.w[j] = s;
    qd = fma(g.v[j], s, qd);
  }
  g.quad = qd;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) s = fma(TM[k], g.Fi[k * P + j], s);
    g.Kp[j] = s;
  }
  {
    const double2* tr = reinterpret_cast<const double2*>(sm + L::T + i * M);
    double2* lr = reinterpret_cast<double2*>(sm + L::Lm + i * M);
#pragma unroll
    for (int k = 0; k < M / 2; ++k) {
      double2 lk = tr[k];
#pragma unroll
      for (int e = 0; e < P; ++e) {
        const double2 z = *reinterpret_cast<const double2*>(sm + L::Z + e * M + 2 * k);
        lk.x = fma(-g.Kp[e], z.x, lk.x);
        lk.y = fma(-g.Kp[e], z.y, lk.y);
      }
      if (act) lr[k] = lk;
    }
  }
  if (act) {
#pragma unroll
    for (int j = 0; j < P; ++j) sm[L::Kp + i * P + j] = g.Kp[j];
  }
  __syncwarp(mask);
}

// ------------------------------------------------------------------------------------------------ forward
template <int M, int P>
__device__ void rowsL_forward(const KfArgs& A, long long u, double* sm, int lane, unsigned mask) {
  using L = RowsLLayout<M, P, false>;
  constexpr int KT = L::KT;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const bool act = lane < M;
  const int i = act ? lane : 0;
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  const double* Cp = A.C.p + draw * A.C.bs;
  const double* P0p = A.P0.p + draw * A.P0.bs;
  for (int k = lane; k < M * M; k += 32) {
    sm[L::T + k] = Tp[k];
    sm[L::Pm + k] = P0p[k];
  }
  for (int k = lane; k < P * M; k += 32) sm[L::Z + k] = Zp[k];
  for (int k = lane; k < P * P; k += 32) sm[L::H + k] = Hp[k];
  if (act) sm[L::a + i] = A.a0.p[draw * A.a0.bs + i];
  double Cr[M];  // the lane's row of C = R Q R^T (static): registers
#pragma unroll
  for (int j = 0; j < M; ++j) Cr[j] = Cp[i * M + j];
  const double ci = (act && A.c.p) ? A.c.p[draw * A.c.bs + i] : 0.0;
  double dv[P];
#pragma unroll
  for (int j = 0; j < P; ++j) dv[j] = A.d.p ? A.d.p[draw * A.d.bs + j] : 0.0;
  __syncwarp(mask);

  const double* y = A.y.p;
  LogAcc acc;
  double llsum = 0.0;
  int info = 0;
  double* tp = A.tape ? A.tape + u * (long long)(n - 1) * KT : nullptr;
  RowLGain<M, P> g;
  double yt[P], ynx[P];
#pragma unroll
  for (int j = 0; j < P; ++j) ynx[j] = y[j];

  for (int t = 0; t < n; ++t) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      yt[j] = ynx[j];
      ynx[j] = y[(long long)(t + 1 < n ? t + 1 : t) * P + j];
    }
    int nm = 0;
#pragma unroll
    for (int j = 0; j < P; ++j) nm += (yt[j] != yt[j]) ? 1 : 0;
    const bool observed = (nm == 0);
    // a' = T a + c (+ Kp v)
    double an = ci;
    {
      const double2* tr = reinterpret_cast<const double2*>(sm + L::T + i * M);
      const double2* av = reinterpret_cast<const double2*>(sm + L::a);
#pragma unroll
      for (int k = 0; k < M / 2; ++k) {
        const double2 tk = tr[k], ak = av[k];
        an = fma(tk.x, ak.x, an);
        an = fma(tk.y, ak.y, an);
      }
    }
    if (observed) {
      rowsL_gain<M, P, L>(sm, yt, A.d_sign, dv, i, act, mask, g);
      if (!g.ok && info == 0) info = t + 1;
      if (g.ok) {
#pragma unroll
        for (int k = 0; k < P; ++k) acc.mul(g.piv[k]);
      }
      llsum += -0.5 * (A.ll_const + g.quad);
#pragma unroll
      for (int k = 0; k < P; ++k) an = fma(g.Kp[k], g.v[k], an);
    } else if (nm != P && info == 0) {
      info = -(t + 1);
    }
    const double* Lsrc = observed ? sm + L::Lm : sm + L::T;  // L = T when nothing is observed
    // ---- C: L^T row i = column i of L ; X row = L row x P
    {
      double2* lt = reinterpret_cast<double2*>(sm + L::LmT + i * M);
#pragma unroll
      for (int k = 0; k < M / 2; ++k) {
        const double2 c2 = make_double2(Lsrc[(2 * k) * M + i], Lsrc[(2 * k + 1) * M + i]);
        if (act) lt[k] = c2;
      }
    }
    double S[M];
#pragma unroll
    for (int j = 0; j < M; ++j) S[j] = 0.0;
    rowmul<M>(S, RowSrc<M>{Lsrc + i * M}, sm + L::Pm);
    if (act) store_row<M>(sm + L::X + i * M, S);
    __syncwarp(mask);
    // ---- D: S2 row = C row + X row x L^T (+ (Kp H) Kp^T)
#pragma unroll
    for (int j = 0; j < M; ++j) S[j] = Cr[j];
    rowmul<M>(S, RowSrc<M>{act ? sm + L::X + i * M : sm + L::T}, sm + L::LmT);  // idle lanes: a row nobody writes
    if (observed) {
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
#pragma unroll
        for (int k = 0; k < P; ++k) S[j] = fma(KH[k], sm[L::Kp + j * P + k], S[j]);
      }
    }
    if (act) store_row<M>(sm + L::X + i * M, S);  // own row: nobody else reads X rows
    __syncwarp(mask);
    // ---- E: P' = sym(S2), a' ; tape
    const bool taped = tp && t + 1 < n;
#pragma unroll
    for (int j = 0; j < M; ++j) S[j] = 0.5 * (S[j] + sm[L::X + j * M + i]);
    if (act) {
      store_row<M>(sm + L::Pm + i * M, S);
      sm[L::a + i] = an;
      if (taped) {
        tp[i] = an;
        double* trow = tp + M + i * M - (i * (i - 1)) / 2 - i;
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (j >= i) trow[j] = S[j];
      }
    }
    if (taped) tp += KT;
    __syncwarp(mask);
  }
  if (lane == 0) {
    double ll = llsum - 0.5 * acc.value();
    if (info != 0) ll = nan("");
    if (A.loglik) A.loglik[u] = ll;
    if (A.info) A.info[u] = info;
  }
}

// ------------------------------------------------------------------------------------------------ adjoint
template <int M, int P, bool NEED_T>
__device__ void rowsL_backward(const KfArgs& A, long long u, double* sm, int lane, unsigned mask) {
  using L = RowsLLayout<M, P, NEED_T>;
  constexpr int KT = L::KT;
  const int n = A.n;
  const long long draw = u / A.n_series;
  const bool act = lane < M;
  const int i = act ? lane : 0;
  const double* Tp = A.T.p + draw * A.T.bs;
  const double* Zp = A.Z.p + draw * A.Z.bs;
  const double* Hp = A.H.p + draw * A.H.bs;
  const double* tape = A.tape + u * (long long)(n - 1) * KT;  // entry t-1 = predicted moments of step t
  if (n >= 2) rows_tape_prefetch<KT, 32>(sm + L::tp, tape + (long long)(n - 2) * KT, lane);
  for (int k = lane; k < M * M; k += 32) {
    sm[L::T + k] = Tp[k];
    sm[L::Pb + k] = 0.0;
  }
  for (int k = lane; k < P * M; k += 32) sm[L::Z + k] = Zp[k];
  for (int k = lane; k < P * P; k += 32) sm[L::H + k] = Hp[k];
  if (act) sm[L::ab + i] = 0.0;
  double dv[P];
#pragma unroll
  for (int j = 0; j < P; ++j) dv[j] = A.d.p ? A.d.p[draw * A.d.bs + j] : 0.0;
  __syncwarp(mask);

  const double* y = A.y.p;
  const double gl = A.g_loglik ? A.g_loglik[u] : 1.0;
  const bool need_H = (A.gH != nullptr);
  // gradient accumulators: the lane's rows of Cb (and Tb) in registers; lanes < P hold rows of Hb; cb (row), db (lane)
  double Cb[M], Tb[NEED_T ? M : 1], Hb[P], cb = 0.0, db = 0.0, abi = 0.0;
#pragma unroll
  for (int j = 0; j < M; ++j) Cb[j] = 0.0;
#pragma unroll
  for (int j = 0; j < (NEED_T ? M : 1); ++j) Tb[j] = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) Hb[j] = 0.0;
  RowLGain<M, P> g;
  double yt[P], ynx[P];
#pragma unroll
  for (int j = 0; j < P; ++j) ynx[j] = y[(long long)(n - 1) * P + j];

  for (int t = n - 1; t >= 0; --t) {
    // ---- predicted moments of step t -> shared memory
    if (t == 0) {
      const double* P0p = A.P0.p + draw * A.P0.bs;
      for (int k = lane; k < M * M; k += 32) sm[L::Pm + k] = P0p[k];
      if (act) sm[L::a + i] = A.a0.p[draw * A.a0.bs + i];
    } else {
      rows_tape_wait();
      __syncwarp(mask);
      const double* tq = sm + L::tp;
      if (act) {
        sm[L::a + i] = tq[i];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          const int lo = i < j ? i : j, hi = i < j ? j : i;
          sm[L::Pm + i * M + j] = tq[M + lo * M - (lo * (lo - 1)) / 2 + (hi - lo)];
        }
      }
      __syncwarp(mask);  // every lane has read the staging buffer: refill it for step t-1
      if (t >= 2) rows_tape_prefetch<KT, 32>(sm + L::tp, tape + (long long)(t - 2) * KT, lane);
    }
    __syncwarp(mask);
#pragma unroll
    for (int j = 0; j < P; ++j) {
      yt[j] = ynx[j];
      ynx[j] = y[(long long)(t > 0 ? t - 1 : 0) * P + j];
    }
    const double lb = gl + (A.g_ll_obs ? A.g_ll_obs[u * n + t] : 0.0);
    bool observed = true;
#pragma unroll
    for (int j = 0; j < P; ++j) observed = observed && (yt[j] == yt[j]);
    if (observed) rowsL_gain<M, P, L>(sm, yt, A.d_sign, dv, i, act, mask, g);
    const double* Lsrc = observed ? sm + L::Lm : sm + L::T;
    if (t == 0) {  // P0 may be any matrix: X needs P + P^T (for t >= 1 the taped P is symmetric: P + P^T = 2 P)
      double S0[M];
#pragma unroll
      for (int j = 0; j < M; ++j) S0[j] = 0.5 * (sm[L::Pm + i * M + j] + sm[L::Pm + j * M + i]);
      __syncwarp(mask);
      if (act) store_row<M>(sm + L::Pm + i * M, S0);
      __syncwarp(mask);
    }
    // ---- 1a: X row = 2 L P ; lz = X Z^T ; Ps row = sym(Pb) row
    double lbz[P];
    cb += abi;
    {
      double X[M];
#pragma unroll
      for (int j = 0; j < M; ++j) X[j] = 0.0;
      rowmul<M>(X, RowSrc<M>{Lsrc + i * M}, sm + L::Pm);
#pragma unroll
      for (int j = 0; j < M; ++j) X[j] *= 2.0;
      if (NEED_T) {
        if (act) store_row<M>(sm + L::Xb + i * M, X);
      } else if (observed) {
#pragma unroll
        for (int e = 0; e < P; ++e) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < M; ++j) s = fma(X[j], sm[L::Z + e * M + j], s);
          if (act) sm[L::lz + i * P + e] = s;
        }
      }
    }
    {
      double Ps[M];
#pragma unroll
      for (int j = 0; j < M; ++j) {
        Ps[j] = 0.5 * (sm[L::Pb + i * M + j] + sm[L::Pb + j * M + i]);
        Cb[j] += Ps[j];
      }
      __syncwarp(mask);  // every lane has read its row and column of Pb (and Pm)
      // ---- 1b: Pb := Ps (from here on the lane reads its Ps row back from shared memory: registers are scarce)
      if (act) store_row<M>(sm + L::Pb + i * M, Ps);
    }
    // (idle lanes 30, 31 shadow row 0 without owning it: give them a row nobody writes, so they never race with lane 0)
    const double* Psr = act ? sm + L::Pb + i * M : sm + L::T;
    // ---- 2: W row = Ps row x L ; (NEED_T) Lb row = Ps row x X ; PK, Kb, T^T ab
    {
      double Wr[M];
#pragma unroll
      for (int j = 0; j < M; ++j) Wr[j] = 0.0;
      rowmul<M>(Wr, RowSrc<M>{Psr}, Lsrc);                 // own row of Pb (just written by this lane)
      if (act) store_row<M>(sm + L::W + i * M, Wr);        // Pm's slot: all reads of Pm are behind the sync above
    }
#pragma unroll
    for (int e = 0; e < P; ++e) lbz[e] = 0.0;
    if (NEED_T) {
      double Lb[M];
#pragma unroll
      for (int j = 0; j < M; ++j) Lb[j] = 0.0;
      rowmul<M>(Lb, RowSrc<M>{Psr}, sm + L::Xb);
#pragma unroll
      for (int j = 0; j < M; ++j) {
        Tb[NEED_T ? j : 0] += fma(abi, sm[L::a + j], Lb[j]);  // Tb += ab a^T + Lb
        if (observed) {
#pragma unroll
          for (int e = 0; e < P; ++e) lbz[e] = fma(Lb[j], sm[L::Z + e * M + j], lbz[e]);
        }
      }
    } else if (observed) {
#pragma unroll
      for (int k = 0; k < M; ++k) {
#pragma unroll
        for (int e = 0; e < P; ++e) lbz[e] = fma(Psr[k], sm[L::lz + k * P + e], lbz[e]);
      }
    }
    double abn = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) abn = fma(sm[L::T + k * M + i], sm[L::ab + k], abn);
    double PK[P], Kb[P];
    if (observed) {
#pragma unroll
      for (int e = 0; e < P; ++e) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(Psr[k], sm[L::Kp + k * P + e], s);
        PK[e] = s;
      }
#pragma unroll
      for (int e = 0; e < P; ++e) {
        double s = abi * g.v[e] - lbz[e];
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(PK[k], sm[L::H + k * P + e] + sm[L::H + e * P + k], s);
        Kb[e] = s;
      }
      if (act) {
#pragma unroll
        for (int e = 0; e < P; ++e) {
          sm[L::Kb + i * P + e] = Kb[e];
          if (need_H) sm[L::PK + i * P + e] = PK[e];
        }
      }
    }
    __syncwarp(mask);
    // ---- 3: Pb' row = (L^T) row x W ; (observed) every lane: K^T Kb, vb, Fb ; TMb row, Tb += TMb Mm^T
    double Pbn[M];
#pragma unroll
    for (int j = 0; j < M; ++j) Pbn[j] = 0.0;
    rowmul<M>(Pbn, ColSrc<M>{Lsrc + i}, sm + L::W);
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
      double TMb[P];
#pragma unroll
      for (int e = 0; e < P; ++e) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(Kb[k], g.Fi[e * P + k], s);
        TMb[e] = s;
      }
      if (NEED_T) {
#pragma unroll
        for (int j = 0; j < M; ++j) {
#pragma unroll
          for (int k = 0; k < P; ++k) Tb[NEED_T ? j : 0] = fma(TMb[k], sm[L::Mm + j * P + k], Tb[NEED_T ? j : 0]);
        }
      }
      if (act) {
#pragma unroll
        for (int e = 0; e < P; ++e) sm[L::TMb + i * P + e] = TMb[e];
      }
    }
    __syncwarp(mask);
    // ---- 4: (observed) Mb = T^T TMb + Z^T Fb ; Pb' += Mb Z ; ab' = T^T ab - Z^T vb ; store Pb', ab'
    if (observed) {
      double Mb[P];
#pragma unroll
      for (int e = 0; e < P; ++e) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(sm[L::T + k * M + i], sm[L::TMb + k * P + e], s);
#pragma unroll
        for (int k = 0; k < P; ++k) s = fma(sm[L::Z + k * M + i], Fb[k * P + e], s);
        Mb[e] = s;
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
#pragma unroll
        for (int k = 0; k < P; ++k) Pbn[j] = fma(Mb[k], sm[L::Z + k * M + j], Pbn[j]);
      }
#pragma unroll
      for (int k = 0; k < P; ++k) abn = fma(-sm[L::Z + k * M + i], vb[k], abn);
#pragma unroll
      for (int k = 0; k < P; ++k)
        if (lane == k) db = fma(-A.d_sign, vb[k], db);
      if (need_H) {
#pragma unroll
        for (int e = 0; e < P; ++e) {
          if (lane != e) continue;
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
    // Pb rows are read by their owner only (phase 2) and ab was last read in phase 3 (behind the sync above)
    if (act) {
      store_row<M>(sm + L::Pb + i * M, Pbn);
      sm[L::ab + i] = abn;
    }
    abi = abn;
    __syncwarp(mask);
  }
  // ---- write-out (row i by lane i)
  if (act) {
    if (A.ga0) A.ga0[u * M + i] = abi;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      if (A.gP0) A.gP0[u * M * M + i * M + j] = sm[L::Pb + i * M + j];
      if (NEED_T && A.gT) A.gT[u * M * M + i * M + j] = Tb[NEED_T ? j : 0];
      if (A.gC) A.gC[u * M * M + i * M + j] = Cb[j];
    }
    if (A.gc) A.gc[u * M + i] = cb;
  }
  if (lane < P) {
    if (A.gd) A.gd[u * P + lane] = db;
#pragma unroll
    for (int j = 0; j < P; ++j)
      if (A.gH) A.gH[u * P * P + lane * P + j] = Hb[j];
  }
}

}  // namespace kfb
