// kf_aux.cuh - launchers of the once-per-draw helper kernels (kf_aux.cu)
#pragma once
#include <cuda_runtime.h>

#include "../../include/kfb200.h"
#include "kf_core.cuh"

namespace kfb {
void count_launch();
cudaError_t launch_rqr_forward(long long nD, int nT, int m, int r, MatArg R, MatArg Q, double* C, cudaStream_t s);
cudaError_t launch_rqr_backward(long long U, long long n_series, int nTC, int nTR, int nTQ, int m, int r, MatArg R,
                                MatArg Q, const double* Cb, double* Rb, double* Qb, int accumulate, cudaStream_t s);
cudaError_t launch_lyapunov_forward(long long B, int m, int r, MatArg A, MatArg R, MatArg Q, double* X, int* info,
                                    cudaStream_t s);
cudaError_t launch_lyapunov_backward(long long B, int m, int r, MatArg A, MatArg R, MatArg Q, const double* X,
                                     const double* Xbar, double* Abar, double* Rbar, double* Qbar, cudaStream_t s);
cudaError_t launch_scatter_forward(long long B, int n_theta, int block, int n_map, const double* theta,
                                   const double* base, const int* src_idx, const int* dst_idx, double* dst,
                                   cudaStream_t s);
cudaError_t launch_scatter_backward(long long B, int n_theta, int block, int n_map, const double* gdst,
                                    const int* src_idx, const int* dst_idx, double* gtheta, cudaStream_t s);
cudaError_t launch_scatter_forward_multi(long long B, int n_theta, int n_seg, const kfb_scatter_seg* segs,
                                         const double* theta, cudaStream_t s);
cudaError_t launch_scatter_backward_multi(long long B, int n_theta, int n_seg, const kfb_scatter_seg* segs,
                                          double* gtheta, cudaStream_t s);
cudaError_t launch_simulate(long long n_sims, long long sims_per_draw, int n, int m, int p, int r, MatArg T, MatArg Z,
                            MatArg R, MatArg H, MatArg Q, const double* x0, long long x0_bs, const double* z_state,
                            const double* z_obs, double* states, double* obs, int* info, cudaStream_t s);
cudaError_t launch_mvn_draws(long long n_sims, long long sims_per_unit, int n, int k, const double* mus, const double* covs,
                             const double* z, const double* jitter, double* out, int* info, cudaStream_t s);
cudaError_t launch_fp64_peak_distinct(int iters, int blocks, int threads, double* sink, cudaStream_t s);
cudaError_t launch_fp64_peak(int iters, int blocks, int threads, double* sink, cudaStream_t s);
}  // namespace kfb
