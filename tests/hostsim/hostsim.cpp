// hostsim.cpp - TEST-ONLY host build of the device step programs in
// pymc_statespace_b200/csrc/kf_core.cuh.  The container that builds this repo has no GPU, so the
// Kalman/adjoint math is debugged here against the oracle before it is run on a B200.  This file is
// NOT part of the product: the package never loads it, and it cannot be reached from any public API.
//
//   g++ -O1 -shared -fPIC -I pymc_statespace_b200/csrc tests/hostsim/hostsim.cpp -o tests/hostsim/_hostsim.so
#include <vector>

#include "kf_ctx.cuh"
#include "kf_dare.cuh"
#include "kf_p1.cuh"
#include "kf_pred.cuh"
#include "kf_smooth.cuh"

using namespace kfb;

static int g_p1 = 0;    // 1: adjoint through kf_p1.cuh (k_endog = 1, standard family) after the predictor-form forward
static int g_pred = 0;  // 1: run the one-step-predictor programs (kf_pred.cuh) where they exist

template <int MK, class X>
static void run_both(X& x, KfArgs& A, int do_bwd) {
  constexpr bool PRED = (MK == MK_STD || MK == MK_STEADY);
  if (A.ll_obs || A.fs) forward_unit<MK, true>(x, A, 0);
  else if (g_pred && PRED) forward_unit_pred<MK == MK_STEADY ? MK_STEADY : MK_STD>(x, A, 0);
  else forward_unit<MK, false>(x, A, 0);
  if (do_bwd) {
    if (g_pred && PRED) backward_unit_pred<MK == MK_STEADY ? MK_STEADY : MK_STD>(x, A, 0);
    else backward_unit<MK>(x, A, 0);
  }
}

template <int M>
static void run_p1(ThreadCtx<M, 1>& x, KfArgs& A, int do_bwd) {
  const bool zu = (A.struct_flags & 1) != 0, h0 = (A.struct_flags & 2) != 0;
  if (g_p1 == 1 && (A.struct_flags & 7) == 7 && !A.gZ && !A.gH) {  // + companion T (+ complete data: compressed tape)
    const bool ct = (A.struct_flags & 8) != 0;
    constexpr int KTA_ = p1::Dim<M>::KTA;  // reduced recursion: (a_t[0], leading block of P_t)
    if (ct) p1::forward_unit_p1<M, true, 4, true>(A, 0, true, A.y.p, A.tape, (long long)KTA_ * 32);
    else p1::forward_unit_p1<M, true, 2, true>(A, 0, true, A.y.p, x.tape_base(A, 0), x.tape_step(A));
    if (!do_bwd) return;
    if (ct) {
      p1::DirectTape<M, KTA_> tc{A.tape + (long long)(A.n - 2) * KTA_ * 32, (long long)KTA_ * 32, 32};
      if (A.g_ll_obs) p1::backward_unit_p1<M, false, false, true, p1::DirectTape<M, KTA_>, 4, true>(A, 0, true, A.y.p, tc);
      else p1::backward_unit_p1<M, false, false, false, p1::DirectTape<M, KTA_>, 4, true>(A, 0, true, A.y.p, tc);
    } else {
      p1::DirectTape<M> tc{x.tape_base(A, 0) + (long long)(A.n - 2) * x.tape_step(A), x.tape_step(A), x.tape_elem(A)};
      if (A.g_ll_obs) p1::backward_unit_p1<M, false, false, true, p1::DirectTape<M>, 2, true>(A, 0, true, A.y.p, tc);
      else p1::backward_unit_p1<M, false, false, false, p1::DirectTape<M>, 2, true>(A, 0, true, A.y.p, tc);
    }
    return;
  }
  if (g_p1 == 2) forward_unit_pred<MK_STD>(x, A, 0);  // generic forward + specialised adjoint
  else if (zu && h0) p1::forward_unit_p1<M, true, true, true>(A, 0, true, A.y.p, x.tape_base(A, 0), x.tape_step(A));
  else if (zu) p1::forward_unit_p1<M, true, true, false>(A, 0, true, A.y.p, x.tape_base(A, 0), x.tape_step(A));
  else p1::forward_unit_p1<M, true>(A, 0, true, A.y.p, x.tape_base(A, 0), x.tape_step(A));
  if (!do_bwd) return;
  if (zu && !A.gZ && !A.gH) {
    p1::DirectTape<M> tp2{x.tape_base(A, 0) + (long long)(A.n - 2) * x.tape_step(A), x.tape_step(A), x.tape_elem(A)};
    if (A.g_ll_obs) {
      if (h0) p1::backward_unit_p1<M, false, false, true, p1::DirectTape<M>, true, true>(A, 0, true, A.y.p, tp2);
      else p1::backward_unit_p1<M, false, false, true, p1::DirectTape<M>, true, false>(A, 0, true, A.y.p, tp2);
    } else {
      if (h0) p1::backward_unit_p1<M, false, false, false, p1::DirectTape<M>, true, true>(A, 0, true, A.y.p, tp2);
      else p1::backward_unit_p1<M, false, false, false, p1::DirectTape<M>, true, false>(A, 0, true, A.y.p, tp2);
    }
    return;
  }
  p1::DirectTape<M> tape{x.tape_base(A, 0) + (long long)(A.n - 2) * x.tape_step(A), x.tape_step(A), x.tape_elem(A)};
  const bool z = A.gZ != nullptr, hh = A.gH != nullptr;
#define KFB_P1RUN(ZZ, HH, GG) p1::backward_unit_p1<M, ZZ, HH, GG>(A, 0, true, A.y.p, tape)
  if (A.g_ll_obs) {
    if (z) KFB_P1RUN(true, true, true);
    else if (hh) KFB_P1RUN(false, true, true);
    else KFB_P1RUN(false, false, true);
  } else {
    if (z) KFB_P1RUN(true, true, false);
    else if (hh) KFB_P1RUN(false, true, false);
    else KFB_P1RUN(false, false, false);
  }
#undef KFB_P1RUN
}

template <class X>
static void run_kind(X& x, KfArgs& A, int do_bwd) {
  if (A.math_kind == MK_STD) run_both<MK_STD>(x, A, do_bwd);
  else if (A.math_kind == MK_UNIV) run_both<MK_UNIV>(x, A, do_bwd);
  else if (A.math_kind == MK_STEADY) run_both<MK_STEADY>(x, A, do_bwd);
  else if (A.math_kind == MK_CHOLS) run_both<MK_CHOLS>(x, A, do_bwd);
}

extern "C" int hostsim_run(int mk, int m, int p, int n, const double* y, const double* a0, const double* P0,
                           const double* T, const double* Z, const double* H, const double* C, const double* c,
                           const double* d, const double* Pss, const double* Gss, const long long* ts /*T,Z,H,C,c,d*/,
                           double ll_const, double d_sign, int static_dims, double* loglik, double* ll_obs, double* fs,
                           double* ps, double* fc, double* pc, int* info, int do_bwd, const double* g_loglik,
                           const double* g_ll_obs, double* ga0, double* gP0, double* gT, double* gZ, double* gH,
                           double* gC, double* gc, double* gd, double* gPss, double* gGss) {
  KfArgs A;
  std::memset(&A, 0, sizeof(A));
  A.U = 1; A.n_series = 1; A.n = n; A.m = m; A.p = p; A.math_kind = mk;
  A.y = {y, 0, 0}; A.a0 = {a0, 0, 0}; A.P0 = {P0, 0, 0};
  A.T = {T, 0, ts[0]}; A.Z = {Z, 0, ts[1]}; A.H = {H, 0, ts[2]}; A.C = {C, 0, ts[3]};
  A.c = {c, 0, ts[4]}; A.d = {d, 0, ts[5]};
  A.Pss = {Pss, 0, 0}; A.Gss = {Gss, 0, 0};
  A.ll_const = ll_const; A.d_sign = d_sign;
  A.loglik = loglik; A.ll_obs = ll_obs; A.fs = fs; A.ps = ps; A.fc = fc; A.pc = pc; A.info = info;
  std::vector<double> tape((size_t)(n > 1 ? n - 1 : 1) * tape_width(m) * 32);  // ThreadCtx layout: whole warps of units
  A.tape = tape.data();
  A.g_loglik = g_loglik; A.g_ll_obs = g_ll_obs;
  A.ga0 = ga0; A.gP0 = gP0; A.gT = gT; A.gZ = gZ; A.gH = gH; A.gC = gC; A.gc = gc; A.gd = gd;
  A.gPss = gPss; A.gGss = gGss;

  g_pred = (static_dims >> 1) & 1;
  A.struct_flags = (static_dims >> 4) & 15;  // bit 4: Z = [1, 0, ..], bit 5: H = 0, bit 6: companion T, bit 7: no missing y
  g_p1 = (static_dims >> 2) & 3;  // 1: specialised forward + adjoint, 2: generic forward + specialised adjoint
  static_dims &= 1;
  if (g_p1) {
    if (p != 1 || mk != MK_STD || ts[0] || ts[1] || ts[2] || ts[3] || ts[4] || ts[5]) return 7;
#define KFB_P1CASE(MM)                \
  if (m == MM) {                      \
    ThreadCtx<MM, 1> x{nullptr, nullptr, 0, 1}; \
    run_p1<MM>(x, A, do_bwd);         \
    return 0;                         \
  }
    KFB_P1CASE(1) KFB_P1CASE(2) KFB_P1CASE(3) KFB_P1CASE(4)
#undef KFB_P1CASE
    return 2;
  }
  const bool tv_any = ts[0] || ts[1] || ts[2] || ts[3] || ts[4] || ts[5];
  if (static_dims) {
#define KFB_CASE(MM, PP)                                   \
  if (m == MM && p == PP) {                                \
    if (tv_any) { /* the time-varying instantiation (standard filter only, kf_thread_inst.inc pick_kind) */ \
      if (mk != MK_STD) return 7;                          \
      ThreadCtx<MM, PP, true> x{nullptr, nullptr, 0, 1};   \
      run_both<MK_STD>(x, A, do_bwd);                      \
      return 0;                                            \
    }                                                      \
    ThreadCtx<MM, PP> x{nullptr, nullptr, 0, 1};                         \
    run_kind(x, A, do_bwd);                                \
    return 0;                                              \
  }
    KFB_CASE(1, 1) KFB_CASE(2, 1) KFB_CASE(2, 2) KFB_CASE(3, 1) KFB_CASE(3, 2) KFB_CASE(3, 3)
    KFB_CASE(4, 1) KFB_CASE(4, 2) KFB_CASE(4, 3)
#undef KFB_CASE
    return 2;
  }
  const int cap = coop_arena_doubles(m, p, true) + coop_arena_doubles(m, p, false);
  std::vector<double> arena((size_t)cap);
  CoopCtx x;
  x.set_dims(m, p); x.lane_ = 0; x.G_ = 1; x.arena = arena.data(); x.cap = cap; x.overflow = false;
  x.off = 0;
  run_kind(x, A, 0);
  const int fwd_used = x.off;
  if (fwd_used > coop_arena_doubles(m, p, false)) return 3;
  if (do_bwd) {
    x.off = 0;  // the adjoint re-uses the arena; the tape written by the forward pass lives in A.tape
    KfArgs B = A;
    B.loglik = nullptr; B.ll_obs = nullptr; B.fs = B.ps = B.fc = B.pc = nullptr; B.info = nullptr;
    if (A.math_kind == MK_STD) { if (g_pred) backward_unit_pred<MK_STD>(x, B, 0); else backward_unit<MK_STD>(x, B, 0); }
    else if (A.math_kind == MK_UNIV) backward_unit<MK_UNIV>(x, B, 0);
    else if (A.math_kind == MK_CHOLS) backward_unit<MK_CHOLS>(x, B, 0);
    else { if (g_pred) backward_unit_pred<MK_STEADY>(x, B, 0); else backward_unit<MK_STEADY>(x, B, 0); }
    if (x.off > coop_arena_doubles(m, p, true)) return 4;
  }
  return x.overflow ? 5 : 0;
}

extern "C" int hostsim_dare(int m, int p, const double* T, const double* Z, const double* H, const double* C, double* Pss,
                            double* Gss, int do_bwd, const double* gPss, const double* gGss, double* gT, double* gZ,
                            double* gH, double* gC) {
  const int cap = dare_arena_doubles(m, p);
  std::vector<double> arena((size_t)cap);
  CoopCtx x;
  x.set_dims(m, p); x.lane_ = 0; x.G_ = 1; x.arena = arena.data(); x.cap = cap; x.overflow = false; x.off = 0;
  x.red = x.bump(34);
  int info = dare_unit(x, T, Z, H, C, Pss, Gss);
  if (x.overflow) return 5;
  if (do_bwd && info == 0) {
    x.off = 34;
    dare_adjoint_unit(x, T, Z, H, Pss, Gss, gPss, gGss, gT, gZ, gH, gC);
    if (x.overflow) return 6;
  }
  return info;
}

extern "C" int hostsim_smoother(int n, int m, const double* T, const double* C, const double* fs, const double* fc, double* ss,
                                double* sc) {
  SmoothArgs S;
  S.U = 1; S.n_series = 1; S.n = n; S.m = m;
  S.T = {T, 0, 0}; S.C = {C, 0, 0}; S.fs = fs; S.fc = fc; S.ss = ss; S.sc = sc;
  const int cap = smoother_arena_doubles(m);
  std::vector<double> arena((size_t)cap);
  CoopCtx x;
  x.set_dims(m, 1); x.lane_ = 0; x.G_ = 1; x.arena = arena.data(); x.cap = cap; x.overflow = false; x.off = 0;
  x.red = x.bump(34);
  smoother_unit(x, S, 0);
  return x.overflow ? 5 : 0;
}
