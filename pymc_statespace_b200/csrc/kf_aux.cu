// kf_aux.cu - once-per-draw helpers around the recursion:
//   C = R Q R^T and its adjoint            (predict, reference kalman_filter.py:219)
//   stationary P0 = Lyapunov(T, R Q R^T)   (reference models/SARIMAX.py:100-107, models/VARMAX.py:143-150)
//   theta -> matrices scatter              (reference models/*.py update())
//   FP64 FMA peak probe                    (roofline denominator)
#include "kf_aux.cuh"

namespace kfb {

// ------------------------------------------------------------------ C = R Q R^T
__global__ void rqr_forward_kernel(long long nD, int nT, int m, int r, MatArg R, MatArg Q, double* __restrict__ C) {
  const long long total = nD * nT * m * m;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % m);
    const int i = (int)((idx / m) % m);
    const int t = (int)((idx / ((long long)m * m)) % nT);
    const long long d = idx / ((long long)m * m * nT);
    const double* Rp = R.p + d * R.bs + t * R.ts;
    const double* Qp = Q.p + d * Q.bs + t * Q.ts;
    double s = 0.0;
    for (int k = 0; k < r; ++k) {
      double rq = 0.0;
      for (int l = 0; l < r; ++l) rq = fma(Rp[i * r + l], Qp[l * r + k], rq);
      s = fma(rq, Rp[j * r + k], s);
    }
    C[idx] = s;
  }
}

// Cb[U, nTC, m, m] -> Rb[U, nTR, m, r] , Qb[U, nTQ, r, r]   (Rb = Cb R Q^T + Cb^T R Q ; Qb = R^T Cb R)
__global__ void rqr_backward_kernel(long long U, long long n_series, int nTC, int nTR, int nTQ, int m, int r, MatArg R,
                                    MatArg Q, const double* __restrict__ Cb, double* __restrict__ Rb,
                                    double* __restrict__ Qb, int accumulate) {
  const long long nR = Rb ? U * nTR * m * r : 0;
  const long long nQ = Qb ? U * nTQ * r * r : 0;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nR + nQ;
       idx += (long long)gridDim.x * blockDim.x) {
    if (idx < nR) {
      const int j = (int)(idx % r);
      const int i = (int)((idx / r) % m);
      const int to = (int)((idx / ((long long)m * r)) % nTR);
      const long long u = idx / ((long long)m * r * nTR);
      const long long d = u / n_series;
      double s = 0.0;
      const int t0 = (nTR > 1) ? to : 0, t1 = (nTR > 1) ? to + 1 : nTC;
      for (int t = t0; t < t1; ++t) {
        const double* Rp = R.p + d * R.bs + t * R.ts;
        const double* Qp = Q.p + d * Q.bs + t * Q.ts;
        const double* Cp = Cb + (u * nTC + t) * m * m;
        for (int k = 0; k < m; ++k) {
          double rq1 = 0.0, rq2 = 0.0;  // (R Q^T)[k][j], (R Q)[k][j]
          for (int l = 0; l < r; ++l) {
            rq1 = fma(Rp[k * r + l], Qp[j * r + l], rq1);
            rq2 = fma(Rp[k * r + l], Qp[l * r + j], rq2);
          }
          s = fma(Cp[i * m + k], rq1, s);
          s = fma(Cp[k * m + i], rq2, s);
        }
      }
      Rb[idx] = accumulate ? Rb[idx] + s : s;
    } else {
      const long long q = idx - nR;
      const int b = (int)(q % r);
      const int a = (int)((q / r) % r);
      const int to = (int)((q / ((long long)r * r)) % nTQ);
      const long long u = q / ((long long)r * r * nTQ);
      const long long d = u / n_series;
      double s = 0.0;
      const int t0 = (nTQ > 1) ? to : 0, t1 = (nTQ > 1) ? to + 1 : nTC;
      for (int t = t0; t < t1; ++t) {
        const double* Rp = R.p + d * R.bs + t * R.ts;
        const double* Cp = Cb + (u * nTC + t) * m * m;
        for (int i = 0; i < m; ++i) {
          double cr = 0.0;
          for (int j = 0; j < m; ++j) cr = fma(Cp[i * m + j], Rp[j * r + b], cr);
          s = fma(Rp[i * r + a], cr, s);
        }
      }
      Qb[q] = accumulate ? Qb[q] + s : s;
    }
  }
}

cudaError_t launch_rqr_forward(long long nD, int nT, int m, int r, MatArg R, MatArg Q, double* C, cudaStream_t s) {
  const long long total = nD * nT * m * m;
  const int block = 256;
  const unsigned grid = (unsigned)min((total + block - 1) / block, (long long)148 * 32);
  rqr_forward_kernel<<<grid, block, 0, s>>>(nD, nT, m, r, R, Q, C);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_rqr_backward(long long U, long long n_series, int nTC, int nTR, int nTQ, int m, int r, MatArg R,
                                MatArg Q, const double* Cb, double* Rb, double* Qb, int accumulate, cudaStream_t s) {
  const long long total = (Rb ? U * nTR * m * r : 0) + (Qb ? U * nTQ * r * r : 0);
  if (total == 0) return cudaSuccess;
  const int block = 256;
  const unsigned grid = (unsigned)min((total + block - 1) / block, (long long)148 * 32);
  rqr_backward_kernel<<<grid, block, 0, s>>>(U, n_series, nTC, nTR, nTQ, m, r, R, Q, Cb, Rb, Qb, accumulate);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------ Lyapunov by squared-Smith doubling
// X = sum_k A^k C A^kT :  X <- X + Ak X Ak^T ; Ak <- Ak Ak  (quadratic convergence for rho(A) < 1).
// One warp per draw, matrices in shared memory.  Lanes stride over the m*m outputs.
struct WarpMat {
  double* v;
  __device__ double& operator[](int i) { return v[i]; }
  __device__ const double& operator[](int i) const { return v[i]; }
};

__device__ __forceinline__ void wmm(int lane, double* C, const double* A, const double* B, int m, bool ta, bool tb,
                                    bool acc) {
  for (int idx = lane; idx < m * m; idx += 32) {
    const int i = idx / m, j = idx - i * m;
    double s = acc ? C[idx] : 0.0;
    for (int k = 0; k < m; ++k) s = fma(ta ? A[k * m + i] : A[i * m + k], tb ? B[j * m + k] : B[k * m + j], s);
    C[idx] = s;
  }
  __syncwarp();
}

// returns true if converged.  Ak is destroyed; X holds the solution; S scratch.
__device__ bool smith_doubling(int lane, double* Ak, double* X, double* S, int m) {
  for (int it = 0; it < 64; ++it) {
    double mx = 0.0;
    for (int idx = lane; idx < m * m; idx += 32) mx = fmax(mx, fabs(Ak[idx]));
    bool bad = !(mx < 1.0e150);
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      bad = bad || __shfl_xor_sync(0xffffffffu, (int)bad, o);
    }
    if (bad) return false;
    if (mx < 1.0e-11) return true;
    wmm(lane, S, Ak, X, m, false, false, false);  // S = Ak X
    wmm(lane, X, S, Ak, m, false, true, true);    // X += S Ak^T
    wmm(lane, S, Ak, Ak, m, false, false, false); // S = Ak Ak
    for (int idx = lane; idx < m * m; idx += 32) Ak[idx] = S[idx];
    __syncwarp();
  }
  return false;
}

__global__ void lyapunov_forward_kernel(long long B, int m, int r, MatArg A, MatArg R, MatArg Q, double* __restrict__ Xo,
                                        int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  double* Ak = sm + (size_t)warp * 3 * m * m;
  double* X = Ak + m * m;
  double* S = X + m * m;
  const double* Ap = A.p + b * A.bs;
  const double* Rp = R.p + b * R.bs;
  const double* Qp = Q.p + b * Q.bs;
  for (int idx = lane; idx < m * m; idx += 32) {
    Ak[idx] = Ap[idx];
    const int i = idx / m, j = idx - i * m;
    double s = 0.0;
    for (int k = 0; k < r; ++k) {
      double rq = 0.0;
      for (int l = 0; l < r; ++l) rq = fma(Rp[i * r + l], Qp[l * r + k], rq);
      s = fma(rq, Rp[j * r + k], s);
    }
    X[idx] = s;
  }
  __syncwarp();
  const bool ok = smith_doubling(lane, Ak, X, S, m);
  for (int idx = lane; idx < m * m; idx += 32) Xo[b * m * m + idx] = ok ? X[idx] : nan("");
  if (lane == 0 && info) info[b] = ok ? 0 : 1;
}

// S = A^T S A + Xbar ;  Abar += S A X^T + S^T A X ;  Cbar = S -> Rbar += S R Q^T + S^T R Q ; Qbar += R^T S R
__global__ void lyapunov_backward_kernel(long long B, int m, int r, MatArg A, MatArg R, MatArg Q,
                                         const double* __restrict__ Xs, const double* __restrict__ Xbar,
                                         double* __restrict__ Abar, double* __restrict__ Rbar,
                                         double* __restrict__ Qbar) {
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  double* Ak = sm + (size_t)warp * 4 * m * m;
  double* S = Ak + m * m;
  double* W = S + m * m;
  double* W2 = W + m * m;
  const double* Ap = A.p + b * A.bs;
  const double* Rp = R.p + b * R.bs;
  const double* Qp = Q.p + b * Q.bs;
  const double* Xp = Xs + b * m * m;
  for (int idx = lane; idx < m * m; idx += 32) {
    const int i = idx / m, j = idx - i * m;
    Ak[idx] = Ap[j * m + i];  // A^T
    S[idx] = Xbar[b * m * m + idx];
  }
  __syncwarp();
  smith_doubling(lane, Ak, S, W, m);
  if (Abar) {
    // W = A X^T ; Abar += S W ; W2 = A X ; Abar += S^T W2
    for (int idx = lane; idx < m * m; idx += 32) {
      const int i = idx / m, j = idx - i * m;
      double s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < m; ++k) {
        s1 = fma(Ap[i * m + k], Xp[j * m + k], s1);
        s2 = fma(Ap[i * m + k], Xp[k * m + j], s2);
      }
      W[idx] = s1;
      W2[idx] = s2;
    }
    __syncwarp();
    for (int idx = lane; idx < m * m; idx += 32) {
      const int i = idx / m, j = idx - i * m;
      double s = Abar[b * m * m + idx];
      for (int k = 0; k < m; ++k) {
        s = fma(S[i * m + k], W[k * m + j], s);
        s = fma(S[k * m + i], W2[k * m + j], s);
      }
      Abar[b * m * m + idx] = s;
    }
  }
  if (Rbar) {
    for (int idx = lane; idx < m * r; idx += 32) {
      const int i = idx / r, j = idx - i * r;
      double s = Rbar[b * m * r + idx];
      for (int k = 0; k < m; ++k) {
        double rq1 = 0.0, rq2 = 0.0;
        for (int l = 0; l < r; ++l) {
          rq1 = fma(Rp[k * r + l], Qp[j * r + l], rq1);
          rq2 = fma(Rp[k * r + l], Qp[l * r + j], rq2);
        }
        s = fma(S[i * m + k], rq1, s);
        s = fma(S[k * m + i], rq2, s);
      }
      Rbar[b * m * r + idx] = s;
    }
  }
  if (Qbar) {
    for (int idx = lane; idx < r * r; idx += 32) {
      const int a = idx / r, c = idx - a * r;
      double s = Qbar[b * r * r + idx];
      for (int i = 0; i < m; ++i) {
        double cr = 0.0;
        for (int j = 0; j < m; ++j) cr = fma(S[i * m + j], Rp[j * r + c], cr);
        s = fma(Rp[i * r + a], cr, s);
      }
      Qbar[b * r * r + idx] = s;
    }
  }
}

// ---- thread-per-draw variants for small k_states: everything in registers, zero synchronisation ----
template <int M>
__device__ __forceinline__ bool smith_doubling_reg(double (&Ak)[M * M], double (&X)[M * M]) {
  double S[M * M];
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < M * M; ++i) mx = fmax(mx, fabs(Ak[i]));
    if (!(mx < 1.0e150)) return false;
    if (mx < 1.0e-11) return true;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(Ak[i * M + k], X[k * M + j], s);
        S[i * M + j] = s;
      }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = X[i * M + j];
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(S[i * M + k], Ak[j * M + k], s);
        X[i * M + j] = s;
      }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fma(Ak[i * M + k], Ak[k * M + j], s);
        S[i * M + j] = s;
      }
#pragma unroll
    for (int i = 0; i < M * M; ++i) Ak[i] = S[i];
  }
  return false;
}

template <int M>
__global__ void __launch_bounds__(128)
    lyapunov_forward_thread_kernel(long long B, int r, MatArg A, MatArg R, MatArg Q, double* __restrict__ Xo,
                                   int* __restrict__ info) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* Ap = A.p + b * A.bs;
  const double* Rp = R.p + b * R.bs;
  const double* Qp = Q.p + b * Q.bs;
  double Ak[M * M], X[M * M];
#pragma unroll
  for (int i = 0; i < M * M; ++i) Ak[i] = Ap[i];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double s = 0.0;
      for (int k = 0; k < r; ++k) {
        double rq = 0.0;
        for (int l = 0; l < r; ++l) rq = fma(Rp[i * r + l], Qp[l * r + k], rq);
        s = fma(rq, Rp[j * r + k], s);
      }
      X[i * M + j] = s;
    }
  const bool ok = smith_doubling_reg<M>(Ak, X);
#pragma unroll
  for (int i = 0; i < M * M; ++i) Xo[b * M * M + i] = ok ? X[i] : nan("");
  if (info) info[b] = ok ? 0 : 1;
}

template <int M>
__global__ void __launch_bounds__(128)
    lyapunov_backward_thread_kernel(long long B, int r, MatArg A, MatArg R, MatArg Q, const double* __restrict__ Xs,
                                    const double* __restrict__ Xbar, double* __restrict__ Abar,
                                    double* __restrict__ Rbar, double* __restrict__ Qbar) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* Ap = A.p + b * A.bs;
  const double* Rp = R.p + b * R.bs;
  const double* Qp = Q.p + b * Q.bs;
  double Ak[M * M], S[M * M], Am[M * M], Xm[M * M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      Am[i * M + j] = Ap[i * M + j];
      Ak[j * M + i] = Am[i * M + j];  // A^T
      S[i * M + j] = Xbar[b * M * M + i * M + j];
      Xm[i * M + j] = Xs[b * M * M + i * M + j];
    }
  smith_doubling_reg<M>(Ak, S);
  if (Abar) {
    double W[M * M], W2[M * M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          s1 = fma(Am[i * M + k], Xm[j * M + k], s1);
          s2 = fma(Am[i * M + k], Xm[k * M + j], s2);
        }
        W[i * M + j] = s1;
        W2[i * M + j] = s2;
      }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = Abar[b * M * M + i * M + j];
#pragma unroll
        for (int k = 0; k < M; ++k) {
          s = fma(S[i * M + k], W[k * M + j], s);
          s = fma(S[k * M + i], W2[k * M + j], s);
        }
        Abar[b * M * M + i * M + j] = s;
      }
  }
  if (Rbar) {
#pragma unroll
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < r; ++j) {
        double s = Rbar[b * M * r + i * r + j];
#pragma unroll
        for (int k = 0; k < M; ++k) {
          double rq1 = 0.0, rq2 = 0.0;
          for (int l = 0; l < r; ++l) {
            rq1 = fma(Rp[k * r + l], Qp[j * r + l], rq1);
            rq2 = fma(Rp[k * r + l], Qp[l * r + j], rq2);
          }
          s = fma(S[i * M + k], rq1, s);
          s = fma(S[k * M + i], rq2, s);
        }
        Rbar[b * M * r + i * r + j] = s;
      }
  }
  if (Qbar) {
    for (int a = 0; a < r; ++a)
      for (int c = 0; c < r; ++c) {
        double s = Qbar[b * r * r + a * r + c];
#pragma unroll
        for (int i = 0; i < M; ++i) {
          double cr = 0.0;
#pragma unroll
          for (int j = 0; j < M; ++j) cr = fma(S[i * M + j], Rp[j * r + c], cr);
          s = fma(Rp[i * r + a], cr, s);
        }
        Qbar[b * r * r + a * r + c] = s;
      }
  }
}

template <int M>
static cudaError_t lyap_thread_fwd(long long B, int r, MatArg A, MatArg R, MatArg Q, double* X, int* info,
                                   cudaStream_t s) {
  lyapunov_forward_thread_kernel<M><<<(unsigned)((B + 127) / 128), 128, 0, s>>>(B, r, A, R, Q, X, info);
  count_launch();
  return cudaGetLastError();
}
template <int M>
static cudaError_t lyap_thread_bwd(long long B, int r, MatArg A, MatArg R, MatArg Q, const double* X, const double* Xbar,
                                   double* Abar, double* Rbar, double* Qbar, cudaStream_t s) {
  lyapunov_backward_thread_kernel<M><<<(unsigned)((B + 127) / 128), 128, 0, s>>>(B, r, A, R, Q, X, Xbar, Abar, Rbar, Qbar);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_lyapunov_forward(long long B, int m, int r, MatArg A, MatArg R, MatArg Q, double* X, int* info,
                                    cudaStream_t s) {
  switch (m) {
    case 1: return lyap_thread_fwd<1>(B, r, A, R, Q, X, info, s);
    case 2: return lyap_thread_fwd<2>(B, r, A, R, Q, X, info, s);
    case 3: return lyap_thread_fwd<3>(B, r, A, R, Q, X, info, s);
    case 4: return lyap_thread_fwd<4>(B, r, A, R, Q, X, info, s);
    default: break;
  }
  const size_t per_warp = (size_t)3 * m * m * sizeof(double);
  int warps = 4;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  if (per_warp * warps > 227 * 1024) return cudaErrorInvalidConfiguration;
  const size_t smem = per_warp * warps;
  cudaFuncSetAttribute(lyapunov_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lyapunov_forward_kernel<<<(unsigned)((B + warps - 1) / warps), warps * 32, smem, s>>>(B, m, r, A, R, Q, X, info);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_lyapunov_backward(long long B, int m, int r, MatArg A, MatArg R, MatArg Q, const double* X,
                                     const double* Xbar, double* Abar, double* Rbar, double* Qbar, cudaStream_t s) {
  switch (m) {
    case 1: return lyap_thread_bwd<1>(B, r, A, R, Q, X, Xbar, Abar, Rbar, Qbar, s);
    case 2: return lyap_thread_bwd<2>(B, r, A, R, Q, X, Xbar, Abar, Rbar, Qbar, s);
    case 3: return lyap_thread_bwd<3>(B, r, A, R, Q, X, Xbar, Abar, Rbar, Qbar, s);
    case 4: return lyap_thread_bwd<4>(B, r, A, R, Q, X, Xbar, Abar, Rbar, Qbar, s);
    default: break;
  }
  const size_t per_warp = (size_t)4 * m * m * sizeof(double);
  int warps = 4;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  if (per_warp * warps > 227 * 1024) return cudaErrorInvalidConfiguration;
  const size_t smem = per_warp * warps;
  cudaFuncSetAttribute(lyapunov_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lyapunov_backward_kernel<<<(unsigned)((B + warps - 1) / warps), warps * 32, smem, s>>>(B, m, r, A, R, Q, X, Xbar, Abar,
                                                                                      Rbar, Qbar);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------ theta -> packed matrices
__global__ void scatter_forward_kernel(long long B, int n_theta, int block, int n_map, const double* __restrict__ theta,
                                       const double* __restrict__ base, const int* __restrict__ src_idx,
                                       const int* __restrict__ dst_idx, double* __restrict__ dst) {
  extern __shared__ int inv[];  // inv[e] = theta index written into element e, or -1 (last writer wins)
  for (int e = threadIdx.x; e < block; e += blockDim.x) inv[e] = -1;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < n_map; ++k) inv[dst_idx[k]] = src_idx[k];
  __syncthreads();
  const long long total = B * block;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(idx % block);
    const long long b = idx / block;
    const int j = inv[e];
    dst[idx] = (j >= 0) ? theta[b * n_theta + j] : base[e];
  }
}

__global__ void scatter_backward_kernel(long long B, int n_theta, int block, int n_map, const double* __restrict__ gdst,
                                        const int* __restrict__ src_idx, const int* __restrict__ dst_idx,
                                        double* __restrict__ gtheta) {
  extern __shared__ int inv[];  // inv[e] = map entry k that owns element e (the last writer), or -1
  for (int e = threadIdx.x; e < block; e += blockDim.x) inv[e] = -1;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < n_map; ++k) inv[dst_idx[k]] = k;
  __syncthreads();
  const long long total = B * n_theta;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % n_theta);
    const long long b = idx / n_theta;
    double s = gtheta[idx];
    for (int k = 0; k < n_map; ++k)
      if (src_idx[k] == j && inv[dst_idx[k]] == k) s += gdst[b * block + dst_idx[k]];
    gtheta[idx] = s;
  }
}

cudaError_t launch_scatter_forward(long long B, int n_theta, int block, int n_map, const double* theta,
                                   const double* base, const int* src_idx, const int* dst_idx, double* dst,
                                   cudaStream_t s) {
  const long long total = B * block;
  const unsigned grid = (unsigned)min((total + 255) / 256, (long long)148 * 32);
  scatter_forward_kernel<<<grid, 256, block * sizeof(int), s>>>(B, n_theta, block, n_map, theta, base, src_idx, dst_idx, dst);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_scatter_backward(long long B, int n_theta, int block, int n_map, const double* gdst,
                                    const int* src_idx, const int* dst_idx, double* gtheta, cudaStream_t s) {
  const long long total = B * n_theta;
  const unsigned grid = (unsigned)min((total + 255) / 256, (long long)148 * 32);
  scatter_backward_kernel<<<grid, 256, block * sizeof(int), s>>>(B, n_theta, block, n_map, gdst, src_idx, dst_idx, gtheta);
  count_launch();
  return cudaGetLastError();
}

// ---- several matrices per launch (kfb_scatter_*_multi)
struct ScatterSegs {
  int n;
  kfb_scatter_seg s[KFB_MAX_SCATTER_SEGMENTS];
};

// grid.y = segment; each slice is scatter_forward_kernel for its matrix
__global__ void scatter_forward_multi_kernel(long long B, int n_theta, const double* __restrict__ theta,
                                             const __grid_constant__ ScatterSegs S) {
  extern __shared__ int inv[];
  const kfb_scatter_seg& g = S.s[blockIdx.y];
  const int block = g.block;
  for (int e = threadIdx.x; e < block; e += blockDim.x) inv[e] = -1;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < g.n_map; ++k) inv[g.dst_idx[k]] = g.src_idx[k];
  __syncthreads();
  const long long total = B * block;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(idx % block);
    const long long b = idx / block;
    const int j = inv[e];
    g.data[idx] = (j >= 0) ? theta[b * n_theta + j] : g.base[e];
  }
}

// one thread per (draw, theta entry): sums the owning map entries of every segment and WRITES gtheta
__global__ void scatter_backward_multi_kernel(long long B, int n_theta, double* __restrict__ gtheta,
                                              const __grid_constant__ ScatterSegs S) {
  extern __shared__ int inv[];  // per segment: inv[e] = map entry that owns element e (the last writer), or -1
  int off = 0;
  for (int q = 0; q < S.n; ++q) {
    for (int e = threadIdx.x; e < S.s[q].block; e += blockDim.x) inv[off + e] = -1;
    off += S.s[q].block;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    off = 0;
    for (int q = 0; q < S.n; ++q) {
      for (int k = 0; k < S.s[q].n_map; ++k) inv[off + S.s[q].dst_idx[k]] = k;
      off += S.s[q].block;
    }
  }
  __syncthreads();
  const long long total = B * n_theta;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % n_theta);
    const long long b = idx / n_theta;
    double s = 0.0;
    off = 0;
    for (int q = 0; q < S.n; ++q) {
      const kfb_scatter_seg& g = S.s[q];
      for (int k = 0; k < g.n_map; ++k)
        if (g.src_idx[k] == j && inv[off + g.dst_idx[k]] == k) s += g.data[b * g.block + g.dst_idx[k]];
      off += g.block;
    }
    gtheta[idx] = s;
  }
}

cudaError_t launch_scatter_forward_multi(long long B, int n_theta, int n_seg, const kfb_scatter_seg* segs,
                                         const double* theta, cudaStream_t s) {
  ScatterSegs S;
  S.n = n_seg;
  int maxblock = 1;
  for (int q = 0; q < n_seg; ++q) {
    S.s[q] = segs[q];
    maxblock = max(maxblock, segs[q].block);
  }
  const long long total = B * maxblock;
  const unsigned gx = (unsigned)min((total + 255) / 256, (long long)148 * 16);
  scatter_forward_multi_kernel<<<dim3(gx, n_seg), 256, maxblock * sizeof(int), s>>>(B, n_theta, theta, S);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_scatter_backward_multi(long long B, int n_theta, int n_seg, const kfb_scatter_seg* segs,
                                          double* gtheta, cudaStream_t s) {
  ScatterSegs S;
  S.n = n_seg;
  int sum = 0;
  for (int q = 0; q < n_seg; ++q) {
    S.s[q] = segs[q];
    sum += segs[q].block;
  }
  const long long total = B * n_theta;
  const unsigned grid = (unsigned)min((total + 255) / 256, (long long)148 * 32);
  const size_t smem = (size_t)sum * sizeof(int);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;  // > 58k mapped elements: caller uses the per-matrix kernel
  if (smem > 48 * 1024) {                                        // e.g. T and P0 of a k_states ~ 80 model mapped
    cudaError_t e = cudaFuncSetAttribute(scatter_backward_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  scatter_backward_multi_kernel<<<grid, 256, smem, s>>>(B, n_theta, gtheta, S);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------ FP64 FMA peak probe
__global__ void fp64_peak_kernel(int iters, double* __restrict__ sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
      a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) sink[0] = s;  // never true; keeps the chain alive
}

// Same probe with three DISTINCT register operands per FMA (no operand-reuse): what real matrix code looks like to
// the register file.  8 accumulators x (own multiplier, own addend).
__global__ void fp64_peak_distinct_kernel(int iters, double* __restrict__ sink) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-9 + i;
    b[i] = 1.0 + 1e-7 * (i + 1);
    c[i] = 1e-9 * (i + 1);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b[(i + k) & 7], c[(i + 3 * k) & 7]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456) sink[0] = s;
}

cudaError_t launch_fp64_peak_distinct(int iters, int blocks, int threads, double* sink, cudaStream_t s) {
  fp64_peak_distinct_kernel<<<blocks, threads, 0, s>>>(iters, sink);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_fp64_peak(int iters, int blocks, int threads, double* sink, cudaStream_t s) {
  fp64_peak_kernel<<<blocks, threads, 0, s>>>(iters, sink);
  count_launch();
  return cudaGetLastError();
}

}  // namespace kfb
