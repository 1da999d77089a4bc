// kf_api.cu - the extern "C" boundary declared in include/kfb200.h.
#include <atomic>
#include <cstdio>

#include "../../include/kfb200.h"
#include "kf_aux.cuh"
#include "kf_kernels.cuh"

namespace kfb {

static std::atomic<long long> g_launches{0};
static thread_local const char* g_cuda_err = "";
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

thread_launch_fn thread_launcher_m1(int p, int mk, bool tv);
thread_launch_fn thread_launcher_m2(int p, int mk, bool tv);
thread_launch_fn thread_launcher_m3(int p, int mk, bool tv);
thread_launch_fn thread_launcher_m4(int p, int mk, bool tv);

bool p1_adjoint_supported(int m, int p, int mk);
cudaError_t launch_p1_adjoint(const KfArgs& A, int y_smem_doubles, int bulk_ok, cudaStream_t s);
cudaError_t launch_p1_forward(const KfArgs& A, int y_smem_doubles, int bulk_ok, cudaStream_t s);

coopT_launch_fn coopT_launcher_m5(int p, int mk);
coopT_launch_fn coopT_launcher_m6(int p, int mk);
coopT_launch_fn coopT_launcher_m7(int p, int mk);
coopT_launch_fn coopT_launcher_m8(int p, int mk);
coopT_launch_fn coopT_launcher_m10(int p, int mk);
coopT_launch_fn coopT_launcher_m12(int p, int mk);
coopT_launch_fn coopT_launcher_m14(int p, int mk);
coopT_launch_fn coopT_launcher_m16(int p, int mk);
coopT_launch_fn coopT_launcher_m18(int p, int mk);
coopT_launch_fn coopT_launcher_m20(int p, int mk);
coopT_launch_fn coopT_launcher_m22(int p, int mk);
coopT_launch_fn coopT_launcher_m24(int p, int mk);
coopT_launch_fn coopT_launcher_m26(int p, int mk);
coopT_launch_fn coopT_launcher_m28(int p, int mk);
coopT_launch_fn coopT_launcher_m30(int p, int mk);
coopT_launch_fn coopT_launcher_m32(int p, int mk);

dare_launch_fn dareD_launcher_m10(int p);
dare_launch_fn dareD_launcher_m12(int p);
dare_launch_fn dareD_launcher_m14(int p);
dare_launch_fn dareD_launcher_m16(int p);
dare_launch_fn dareD_launcher_m18(int p);
dare_launch_fn dareD_launcher_m20(int p);
dare_launch_fn dareD_launcher_m22(int p);
dare_launch_fn dareD_launcher_m24(int p);
dare_launch_fn dareD_launcher_m26(int p);
dare_launch_fn dareD_launcher_m28(int p);
dare_launch_fn dareD_launcher_m30(int p);
dare_launch_fn dareD_launcher_m32(int p);

dare_launch_fn find_dareD_launcher(int m, int p) {
  switch (m) {
    case 10: return dareD_launcher_m10(p);
    case 12: return dareD_launcher_m12(p);
    case 14: return dareD_launcher_m14(p);
    case 16: return dareD_launcher_m16(p);
    case 18: return dareD_launcher_m18(p);
    case 20: return dareD_launcher_m20(p);
    case 22: return dareD_launcher_m22(p);
    case 24: return dareD_launcher_m24(p);
    case 26: return dareD_launcher_m26(p);
    case 28: return dareD_launcher_m28(p);
    case 30: return dareD_launcher_m30(p);
    case 32: return dareD_launcher_m32(p);
    default: return nullptr;
  }
}

coopT_launch_fn find_coopT_launcher(int m, int p, int mk) {
  switch (m) {
    case 5: return coopT_launcher_m5(p, mk);
    case 6: return coopT_launcher_m6(p, mk);
    case 7: return coopT_launcher_m7(p, mk);
    case 8: return coopT_launcher_m8(p, mk);
    case 10: return coopT_launcher_m10(p, mk);
    case 12: return coopT_launcher_m12(p, mk);
    case 14: return coopT_launcher_m14(p, mk);
    case 16: return coopT_launcher_m16(p, mk);
    case 18: return coopT_launcher_m18(p, mk);
    case 20: return coopT_launcher_m20(p, mk);
    case 22: return coopT_launcher_m22(p, mk);
    case 24: return coopT_launcher_m24(p, mk);
    case 26: return coopT_launcher_m26(p, mk);
    case 28: return coopT_launcher_m28(p, mk);
    case 30: return coopT_launcher_m30(p, mk);
    case 32: return coopT_launcher_m32(p, mk);
    default: return nullptr;
  }
}

smoother_launch_fn smoother_thread_m1();
smoother_launch_fn smoother_thread_m2();
smoother_launch_fn smoother_thread_m3();
smoother_launch_fn smoother_thread_m4();

smoother_launch_fn find_smoother_thread_launcher(int m) {
  switch (m) {
    case 1: return smoother_thread_m1();
    case 2: return smoother_thread_m2();
    case 3: return smoother_thread_m3();
    case 4: return smoother_thread_m4();
    default: return nullptr;
  }
}

thread_launch_fn find_thread_launcher(int m, int p, int mk, bool tv) {
  switch (m) {
    case 1: return thread_launcher_m1(p, mk, tv);
    case 2: return thread_launcher_m2(p, mk, tv);
    case 3: return thread_launcher_m3(p, mk, tv);
    case 4: return thread_launcher_m4(p, mk, tv);
    default: return nullptr;
  }
}

struct Plan {
  int mk;
  double ll_const, d_sign;
  bool tv_any, use_thread;
  bool compressed;   // k_endog = 1 kernels with all four structure promises: tape entries hold a_t and the leading block of P_t
  long long U, nD;
  int nTC;           // time entries of C = R Q R^T (1 or n)
  long long C_bs;    // 0 if R and Q are shared by all draws
  size_t off_C, off_Pss, off_Gss, off_dinfo, off_tape, off_gC, off_gPss, off_gGss, total;
};

static kfb_status make_plan(const kfb_desc* d, bool save, Plan* pl) {
  if (!d || d->n_draws <= 0 || d->n_series <= 0 || d->n <= 0 || d->m <= 0 || d->p <= 0 || d->r <= 0)
    return KFB_ERR_INVALID_ARG;
  const bool corrected = (d->flags & KFB_FLAG_CORRECTED) != 0;
  const double l2pi = KF_LOG_2PI;
  switch (d->filter_kind) {
    case KFB_STANDARD:
      pl->mk = MK_STD; pl->ll_const = corrected ? d->p * l2pi : l2pi; pl->d_sign = 1.0; break;
    case KFB_SINGLE:
      if (d->p != 1) return KFB_ERR_INVALID_ARG;  // assert_data_is_1d, kalman_filter.py:19,329
      pl->mk = MK_STD; pl->ll_const = l2pi; pl->d_sign = corrected ? 1.0 : -1.0; break;
    case KFB_CHOLESKY:
      // as coded the filter is exact only for p = 1; p > 1 strict = bug-compatible MK_CHOLS
      pl->mk = (d->p != 1 && !corrected) ? MK_CHOLS : MK_STD; pl->ll_const = d->p * l2pi; pl->d_sign = 1.0; break;
    case KFB_UNIVARIATE:
      pl->mk = MK_UNIV; pl->ll_const = 0.0; pl->d_sign = 1.0; break;
    case KFB_STEADY_STATE:
      pl->mk = MK_STEADY; pl->ll_const = corrected ? d->p * l2pi : l2pi; pl->d_sign = corrected ? 1.0 : 0.0; break;
    default: return KFB_ERR_INVALID_ARG;
  }
  pl->tv_any = d->T_ts || d->Z_ts || d->R_ts || d->H_ts || d->Q_ts || d->c_ts || d->d_ts;
  if (pl->tv_any && (pl->mk == MK_UNIV || pl->mk == MK_STEADY)) return KFB_ERR_UNSUPPORTED;
  pl->U = d->n_draws * d->n_series;
  pl->nD = d->n_draws;
  pl->nTC = (d->R_ts || d->Q_ts) ? d->n : 1;
  const bool C_batched = d->R_bs || d->Q_bs;
  const long long nDC = C_batched ? d->n_draws : 1;
  pl->C_bs = C_batched ? (long long)pl->nTC * d->m * d->m : 0;
  pl->use_thread = !(d->flags & KFB_FLAG_FORCE_COOP) && find_thread_launcher(d->m, d->p, pl->mk, pl->tv_any) != nullptr;
  size_t off = 0;
  auto take = [&](size_t doubles) { size_t o = off; off += ((doubles * 8 + 255) / 256) * 256; return o; };
  pl->off_C = take((size_t)nDC * pl->nTC * d->m * d->m);
  pl->off_Pss = pl->off_Gss = pl->off_gPss = pl->off_gGss = 0;
  pl->off_dinfo = 0;
  if (pl->mk == MK_STEADY) {
    pl->off_Pss = take((size_t)pl->nD * d->m * d->m);
    pl->off_Gss = take((size_t)pl->nD * d->p * d->p);
    pl->off_dinfo = take((size_t)(pl->nD + 1) / 2);
  }
  // decided from the descriptor alone, so that workspace sizing, the forward pass and the adjoint always agree
  constexpr uint32_t kPromises = KFB_FLAG_Z_UNIT0 | KFB_FLAG_H_ZERO | KFB_FLAG_T_COMPANION | KFB_FLAG_NO_MISSING;
  pl->compressed = pl->use_thread && !pl->tv_any && d->y_bs == 0 && d->n_series == 1 && (d->flags & kPromises) == kPromises &&
                   !(d->flags & KFB_FLAG_GENERIC_ADJOINT) && p1_adjoint_supported(d->m, d->p, pl->mk);
  pl->off_tape = pl->off_gC = 0;
  if (save) {
    const size_t entry = pl->compressed ? (size_t)(1 + ((d->m - 1) * d->m) / 2) : (size_t)tape_width(d->m);
    pl->off_tape = take((size_t)tape_units_padded(pl->U) * (d->n > 1 ? d->n - 1 : 0) * entry);
    pl->off_gC = take((size_t)pl->U * pl->nTC * d->m * d->m);
    if (pl->mk == MK_STEADY) {
      pl->off_gPss = take((size_t)pl->U * d->m * d->m);
      pl->off_gGss = take((size_t)pl->U * d->p * d->p);
    }
  }
  pl->total = off;
  return KFB_OK;
}

static void fill_args(const kfb_desc* d, const kfb_inputs* in, const Plan& pl, char* ws, KfArgs* A) {
  std::memset(A, 0, sizeof(*A));
  A->U = pl.U; A->n_series = d->n_series; A->n = d->n; A->m = d->m; A->p = d->p; A->math_kind = pl.mk;
  A->y = {in->y, d->y_bs, 0};
  A->a0 = {in->a0, d->a0_bs, 0};
  A->P0 = {in->P0, d->P0_bs, 0};
  A->T = {in->T, d->T_bs, d->T_ts};
  A->Z = {in->Z, d->Z_bs, d->Z_ts};
  A->H = {in->H, d->H_bs, d->H_ts};
  A->C = {(const double*)(ws + pl.off_C), pl.C_bs, pl.nTC > 1 ? (long long)d->m * d->m : 0};
  A->c = {in->c, d->c_bs, d->c_ts};
  A->d = {in->d, d->d_bs, d->d_ts};
  if (pl.mk == MK_STEADY) {
    A->Pss = {(const double*)(ws + pl.off_Pss), (long long)d->m * d->m, 0};  // one per draw
    A->Gss = {(const double*)(ws + pl.off_Gss), (long long)d->p * d->p, 0};
    A->dare_info = (const int*)(ws + pl.off_dinfo);
  }
  A->ll_const = pl.ll_const;
  A->d_sign = pl.d_sign;
  if (d->p == 1 && !d->Z_ts && !d->H_ts)
    A->struct_flags = ((d->flags & KFB_FLAG_Z_UNIT0) ? 1 : 0) | ((d->flags & KFB_FLAG_H_ZERO) ? 2 : 0) |
                      ((d->flags & KFB_FLAG_T_COMPANION) ? 4 : 0);
}

static kfb_status cuda_fail(cudaError_t e) {
  g_cuda_err = cudaGetErrorString(e);
  return KFB_ERR_CUDA;
}

static kfb_status launch_main(const kfb_desc* d, const Plan& pl, const KfArgs& A, bool bwd, cudaStream_t s) {
  cudaError_t e;
  if (pl.use_thread) {
    // shared observation stream -> stage it in shared memory once per CTA
    int ysm = 0, bulk_ok = 0;
    const long long ydoubles = (long long)d->n * d->p;
    if (d->y_bs == 0 && d->n_series == 1 && ydoubles * 8 <= 96 * 1024) {
      ysm = (int)ydoubles;
      bulk_ok = ((uintptr_t)A.y.p % 16 == 0) ? 1 : 0;
    }
    // k_endog = 1, shared observations: the specialised adjoint with the TMA tape ring (kf_p1.cu)
    const bool p1_ok = !pl.tv_any && d->y_bs == 0 && d->n_series == 1 && !(d->flags & KFB_FLAG_GENERIC_ADJOINT) &&
                       p1_adjoint_supported(d->m, d->p, pl.mk);
    const bool full = A.ll_obs || A.fs || A.ps || A.fc || A.pc;
    // all four structure promises on the k_endog = 1 kernels: compressed tape entries (kf_p1.cuh, ZU == 3) - decided from
    // the descriptor alone, so that the forward pass and the adjoint always agree on the format
    KfArgs B = A;
    if (pl.compressed && p1_ok && (A.struct_flags & 7) == 7) B.struct_flags |= 8;
    if ((B.struct_flags & 8) && bwd && (A.gZ || A.gH)) return KFB_ERR_UNSUPPORTED;  // Z, H were promised constant
    if (p1_ok && bwd)
      e = launch_p1_adjoint(B, ysm, bulk_ok, s);
    else if (p1_ok && !full)
      e = launch_p1_forward(B, ysm, bulk_ok, s);
    else
      e = find_thread_launcher(d->m, d->p, pl.mk, pl.tv_any)(B, bwd, ysm, bulk_ok, s);
  } else {
    // sub-warp kernels with compile-time dims need uniform control flow across the units of a warp:
    // static matrices and ONE observation stream shared by every unit
    coopT_launch_fn ft = nullptr;
    if (!pl.tv_any && !(d->flags & KFB_FLAG_FORCE_COOP) && d->y_bs == 0 && d->n_series == 1)
      ft = find_coopT_launcher(d->m, d->p, pl.mk);
    e = ft ? ft(A, bwd, s) : launch_coop(A, bwd, s);
    if (e == cudaErrorInvalidConfiguration) return KFB_ERR_UNSUPPORTED;
  }
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

}  // namespace kfb

using namespace kfb;

extern "C" {
#pragma GCC visibility push(default)

int32_t kfb_version(void) { return KFB_VERSION; }

const char* kfb_status_string(kfb_status s) {
  switch (s) {
    case KFB_OK: return "ok";
    case KFB_ERR_INVALID_ARG: return "invalid argument";
    case KFB_ERR_UNSUPPORTED: return "unsupported configuration";
    case KFB_ERR_WORKSPACE: return "workspace missing or too small";
    case KFB_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}

const char* kfb_last_cuda_error(void) { return g_cuda_err; }

int64_t kfb_launch_count(void) { return g_launches.load(); }

kfb_status kfb_workspace_bytes(const kfb_desc* desc, int32_t save_for_backward, size_t* bytes) {
  if (!bytes) return KFB_ERR_INVALID_ARG;
  Plan pl;
  kfb_status st = make_plan(desc, save_for_backward != 0, &pl);
  if (st != KFB_OK) return st;
  *bytes = pl.total;
  return KFB_OK;
}

kfb_status kfb_forward(const kfb_desc* desc, const kfb_inputs* in, const kfb_outputs* out, void* workspace,
                       size_t workspace_bytes, int32_t save_for_backward, void* stream) {
  if (!in || !out || !in->y || !in->a0 || !in->P0 || !in->T || !in->Z || !in->R || !in->H || !in->Q)
    return KFB_ERR_INVALID_ARG;
  Plan pl;
  kfb_status st = make_plan(desc, save_for_backward != 0, &pl);
  if (st != KFB_OK) return st;
  if (!workspace || workspace_bytes < pl.total) return KFB_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  KfArgs A;
  fill_args(desc, in, pl, ws, &A);
  A.loglik = out->loglik; A.ll_obs = out->ll_obs; A.fs = out->filtered_states; A.ps = out->predicted_states;
  A.fc = out->filtered_covs; A.pc = out->predicted_covs; A.info = out->info;
  A.tape = save_for_backward ? (double*)(ws + pl.off_tape) : nullptr;
  const long long nDC = pl.C_bs ? desc->n_draws : 1;
  cudaError_t e = launch_rqr_forward(nDC, pl.nTC, desc->m, desc->r, MatArg{in->R, desc->R_bs, desc->R_ts},
                                     MatArg{in->Q, desc->Q_bs, desc->Q_ts}, (double*)(ws + pl.off_C), s);
  if (e != cudaSuccess) return cuda_fail(e);
  if (pl.mk == MK_STEADY) {
    // P_steady = DARE(T^T, Z^T, R Q R^T, H), F_inv = (Z P_steady Z^T + H)^-1   (kalman_filter.py:384-386)
    DareArgs D;
    std::memset(&D, 0, sizeof(D));
    D.nD = desc->n_draws; D.U = pl.U; D.n_series = desc->n_series; D.m = desc->m; D.p = desc->p;
    D.T = MatArg{in->T, desc->T_bs, 0}; D.Z = MatArg{in->Z, desc->Z_bs, 0}; D.H = MatArg{in->H, desc->H_bs, 0};
    D.C = A.C;
    D.Pss = (double*)(ws + pl.off_Pss); D.Gss = (double*)(ws + pl.off_Gss); D.info = (int*)(ws + pl.off_dinfo);
    const dare_launch_fn fast = (desc->flags & KFB_FLAG_FORCE_COOP) ? nullptr : find_dareD_launcher(desc->m, desc->p);
    e = fast ? fast(D, s) : launch_dare(D, false, s);  // even k_states 18..32, k_endog 1: tensor-core mapping (kf_rowsD.cuh)
    if (e == cudaErrorInvalidConfiguration) return KFB_ERR_UNSUPPORTED;
    if (e != cudaSuccess) return cuda_fail(e);
  }
  return launch_main(desc, pl, A, false, s);
}

kfb_status kfb_backward(const kfb_desc* desc, const kfb_inputs* in, const kfb_cotangents* cot, const kfb_grads* g,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (!in || !g || !in->y || !in->a0 || !in->P0 || !in->T || !in->Z || !in->R || !in->H || !in->Q)
    return KFB_ERR_INVALID_ARG;
  Plan pl;
  kfb_status st = make_plan(desc, true, &pl);
  if (st != KFB_OK) return st;
  if (!workspace || workspace_bytes < pl.total) return KFB_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  KfArgs A;
  fill_args(desc, in, pl, ws, &A);
  A.tape = (double*)(ws + pl.off_tape);
  A.g_loglik = cot ? cot->g_loglik : nullptr;
  A.g_ll_obs = cot ? cot->g_ll_obs : nullptr;
  A.ga0 = g->a0; A.gP0 = g->P0; A.gT = g->T; A.gZ = g->Z; A.gH = g->H; A.gc = g->c; A.gd = g->d;
  const bool want_C = g->R || g->Q;
  A.gC = want_C ? (double*)(ws + pl.off_gC) : nullptr;
  if (pl.mk == MK_STEADY) {
    A.gPss = (double*)(ws + pl.off_gPss);
    A.gGss = (double*)(ws + pl.off_gGss);
  }
  st = launch_main(desc, pl, A, true, s);
  if (st != KFB_OK) return st;
  if (pl.mk == MK_STEADY) {
    // chain (P_steady-bar, F_inv-bar) through the DARE (utils/pytensor_scipy.py:39-60) into T, Z, H, C
    DareArgs D;
    std::memset(&D, 0, sizeof(D));
    D.nD = desc->n_draws; D.U = pl.U; D.n_series = desc->n_series; D.m = desc->m; D.p = desc->p;
    D.T = MatArg{in->T, desc->T_bs, 0}; D.Z = MatArg{in->Z, desc->Z_bs, 0}; D.H = MatArg{in->H, desc->H_bs, 0};
    D.C = A.C;
    D.Pss = (double*)(ws + pl.off_Pss); D.Gss = (double*)(ws + pl.off_Gss);
    D.gPss = A.gPss; D.gGss = A.gGss; D.gT = g->T; D.gZ = g->Z; D.gH = g->H; D.gC = A.gC;
    cudaError_t e = launch_dare(D, true, s);
    if (e == cudaErrorInvalidConfiguration) return KFB_ERR_UNSUPPORTED;
    if (e != cudaSuccess) return cuda_fail(e);
  }
  if (want_C) {
    cudaError_t e = launch_rqr_backward(pl.U, desc->n_series, pl.nTC, desc->R_ts ? desc->n : 1, desc->Q_ts ? desc->n : 1,
                                        desc->m, desc->r, MatArg{in->R, desc->R_bs, desc->R_ts},
                                        MatArg{in->Q, desc->Q_bs, desc->Q_ts}, A.gC, g->R, g->Q, 0, s);
    if (e != cudaSuccess) return cuda_fail(e);
  }
  return KFB_OK;
}

kfb_status kfb_smoother(int64_t n_draws, int64_t n_series, int32_t n, int32_t m, int32_t r, const double* T, int64_t T_bs,
                        const double* R, int64_t R_bs, const double* Q, int64_t Q_bs, const double* filtered_states,
                        const double* filtered_covs, double* smoothed_states, double* smoothed_covs, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (n_draws <= 0 || n_series <= 0 || n <= 0 || m <= 0 || r <= 0 || !T || !R || !Q || !filtered_states || !filtered_covs ||
      !smoothed_states || !smoothed_covs)
    return KFB_ERR_INVALID_ARG;
  const bool C_batched = R_bs || Q_bs;
  const long long nDC = C_batched ? n_draws : 1;
  const size_t need = (size_t)nDC * m * m * sizeof(double);
  if (!workspace || workspace_bytes < need) return KFB_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = launch_rqr_forward(nDC, 1, m, r, MatArg{R, R_bs, 0}, MatArg{Q, Q_bs, 0}, (double*)workspace, s);
  if (e != cudaSuccess) return cuda_fail(e);
  SmoothArgs S;
  S.U = n_draws * n_series; S.n_series = n_series; S.n = n; S.m = m;
  S.T = MatArg{T, T_bs, 0};
  S.C = MatArg{(const double*)workspace, C_batched ? (long long)m * m : 0, 0};
  S.fs = filtered_states; S.fc = filtered_covs; S.ss = smoothed_states; S.sc = smoothed_covs;
  const smoother_launch_fn per_thread = find_smoother_thread_launcher(m);
  e = per_thread ? per_thread(S, s) : launch_smoother(S, s);
  if (e == cudaErrorInvalidConfiguration) return KFB_ERR_UNSUPPORTED;
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_lyapunov_forward(int64_t B, int32_t m, int32_t r, const double* A, int64_t A_bs, const double* R,
                                int64_t R_bs, const double* Q, int64_t Q_bs, double* X, int32_t* info, void* stream) {
  if (B <= 0 || m <= 0 || r <= 0 || !A || !R || !Q || !X) return KFB_ERR_INVALID_ARG;
  cudaError_t e = launch_lyapunov_forward(B, m, r, MatArg{A, A_bs, 0}, MatArg{R, R_bs, 0}, MatArg{Q, Q_bs, 0}, X, info,
                                          (cudaStream_t)stream);
  if (e == cudaErrorInvalidConfiguration) return KFB_ERR_UNSUPPORTED;
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_lyapunov_backward(int64_t B, int32_t m, int32_t r, const double* A, int64_t A_bs, const double* R,
                                 int64_t R_bs, const double* Q, int64_t Q_bs, const double* X, const double* Xbar,
                                 double* Abar, double* Rbar, double* Qbar, void* stream) {
  if (B <= 0 || m <= 0 || r <= 0 || !A || !R || !Q || !X || !Xbar) return KFB_ERR_INVALID_ARG;
  cudaError_t e = launch_lyapunov_backward(B, m, r, MatArg{A, A_bs, 0}, MatArg{R, R_bs, 0}, MatArg{Q, Q_bs, 0}, X, Xbar,
                                           Abar, Rbar, Qbar, (cudaStream_t)stream);
  if (e == cudaErrorInvalidConfiguration) return KFB_ERR_UNSUPPORTED;
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_scatter_forward(int64_t B, int32_t n_theta, int32_t block, int32_t n_map, const double* theta,
                               const double* base, const int32_t* src_idx, const int32_t* dst_idx, double* dst,
                               void* stream) {
  if (B <= 0 || n_theta <= 0 || block <= 0 || n_map < 0 || !theta || !base || !dst) return KFB_ERR_INVALID_ARG;
  if (n_map > 0 && (!src_idx || !dst_idx)) return KFB_ERR_INVALID_ARG;
  cudaError_t e = launch_scatter_forward(B, n_theta, block, n_map, theta, base, src_idx, dst_idx, dst,
                                         (cudaStream_t)stream);
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_scatter_backward(int64_t B, int32_t n_theta, int32_t block, int32_t n_map, const double* gdst,
                                const int32_t* src_idx, const int32_t* dst_idx, double* gtheta, void* stream) {
  if (B <= 0 || n_theta <= 0 || block <= 0 || n_map < 0 || !gdst || !gtheta) return KFB_ERR_INVALID_ARG;
  if (n_map > 0 && (!src_idx || !dst_idx)) return KFB_ERR_INVALID_ARG;
  cudaError_t e =
      launch_scatter_backward(B, n_theta, block, n_map, gdst, src_idx, dst_idx, gtheta, (cudaStream_t)stream);
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

static bool scatter_segs_ok(int32_t n_seg, const kfb_scatter_seg* segs, bool forward) {
  if (n_seg <= 0 || n_seg > KFB_MAX_SCATTER_SEGMENTS || !segs) return false;
  for (int q = 0; q < n_seg; ++q) {
    const kfb_scatter_seg& g = segs[q];
    if (g.block <= 0 || g.n_map < 0 || !g.data || (forward && !g.base)) return false;
    if (g.n_map > 0 && (!g.src_idx || !g.dst_idx)) return false;
  }
  return true;
}

kfb_status kfb_scatter_forward_multi(int64_t B, int32_t n_theta, int32_t n_seg, const kfb_scatter_seg* segs,
                                     const double* theta, void* stream) {
  if (B <= 0 || n_theta <= 0 || !theta || !scatter_segs_ok(n_seg, segs, true)) return KFB_ERR_INVALID_ARG;
  cudaError_t e = launch_scatter_forward_multi(B, n_theta, n_seg, segs, theta, (cudaStream_t)stream);
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_scatter_backward_multi(int64_t B, int32_t n_theta, int32_t n_seg, const kfb_scatter_seg* segs,
                                      double* gtheta, void* stream) {
  if (B <= 0 || n_theta <= 0 || !gtheta || !scatter_segs_ok(n_seg, segs, false)) return KFB_ERR_INVALID_ARG;
  cudaError_t e = launch_scatter_backward_multi(B, n_theta, n_seg, segs, gtheta, (cudaStream_t)stream);
  if (e == cudaErrorInvalidConfiguration) return KFB_ERR_UNSUPPORTED;  // mapped elements exceed shared memory
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_simulate(int64_t n_draws, int64_t sims_per_draw, int32_t n, int32_t m, int32_t p, int32_t r, const double* T,
                        int64_t T_bs, const double* Z, int64_t Z_bs, const double* R, int64_t R_bs, const double* H,
                        int64_t H_bs, const double* Q, int64_t Q_bs, const double* x0, int64_t x0_bs,
                        const double* z_state, const double* z_obs, double* states, double* obs, int32_t* info,
                        void* stream) {
  if (n_draws <= 0 || sims_per_draw <= 0 || n <= 0 || m <= 0 || p <= 0 || r <= 0 || !T || !Z || !R || !H || !Q || !z_state ||
      !z_obs || !states || !obs)
    return KFB_ERR_INVALID_ARG;
  if (m > 32 || p > 32 || r > 32) return KFB_ERR_UNSUPPORTED;
  cudaError_t e = launch_simulate(n_draws * sims_per_draw, sims_per_draw, n, m, p, r, MatArg{T, T_bs, 0}, MatArg{Z, Z_bs, 0},
                                  MatArg{R, R_bs, 0}, MatArg{H, H_bs, 0}, MatArg{Q, Q_bs, 0}, x0, x0_bs, z_state, z_obs,
                                  states, obs, info, (cudaStream_t)stream);
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_mvn_draws(int64_t n_units, int64_t sims_per_unit, int32_t n, int32_t k, const double* mus, const double* covs,
                         const double* z, const double* jitter, double* out, int32_t* info, void* stream) {
  if (n_units <= 0 || sims_per_unit <= 0 || n <= 0 || k <= 0 || !mus || !covs || !z || !out) return KFB_ERR_INVALID_ARG;
  if (k > 32) return KFB_ERR_UNSUPPORTED;
  cudaError_t e = launch_mvn_draws(n_units * sims_per_unit, sims_per_unit, n, k, mus, covs, z, jitter, out, info,
                                   (cudaStream_t)stream);
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_fp64_peak_distinct(int32_t iters, int32_t blocks, int32_t threads, double* sink, double* h_flops,
                                  void* stream) {
  if (iters <= 0 || blocks <= 0 || threads <= 0 || threads > 1024 || !sink) return KFB_ERR_INVALID_ARG;
  cudaError_t e = launch_fp64_peak_distinct(iters, blocks, threads, sink, (cudaStream_t)stream);
  if (h_flops) *h_flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

kfb_status kfb_fp64_peak(int32_t iters, int32_t blocks, int32_t threads, double* sink, double* h_flops, void* stream) {
  if (iters <= 0 || blocks <= 0 || threads <= 0 || threads > 1024 || !sink) return KFB_ERR_INVALID_ARG;
  cudaError_t e = launch_fp64_peak(iters, blocks, threads, sink, (cudaStream_t)stream);
  if (h_flops) *h_flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  return e == cudaSuccess ? KFB_OK : cuda_fail(e);
}

#pragma GCC visibility pop
}  // extern "C"
