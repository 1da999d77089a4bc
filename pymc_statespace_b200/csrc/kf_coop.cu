// kf_coop.cu - cooperative (shared-memory) kernels: any (k_states, k_endog), time-varying matrices.
#include "kf_kernels.cuh"

namespace kfb {

template <int MK, bool WARP>
static cudaError_t launch_coop_kind(const KfArgs& A, bool bwd, int arena, int block, unsigned grid, size_t smem,
                                    cudaStream_t s) {
  const bool full = A.ll_obs || A.fs || A.ps || A.fc || A.pc;
#define KFB_LAUNCH(MODE)                                                                                          \
  do {                                                                                                            \
    cudaFuncSetAttribute(kf_coop_kernel<MK, MODE, WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    kf_coop_kernel<MK, MODE, WARP><<<grid, block, smem, s>>>(A, arena);                                           \
  } while (0)
  if (bwd) KFB_LAUNCH(2);
  else if (full) KFB_LAUNCH(1);
  else KFB_LAUNCH(0);
#undef KFB_LAUNCH
  count_launch();
  return cudaGetLastError();
}

template <bool WARP>
static cudaError_t launch_coop_mode(const KfArgs& A, bool bwd, int arena, int block, unsigned grid, size_t smem,
                                    cudaStream_t s) {
  switch (A.math_kind) {
    case MK_STD: return launch_coop_kind<MK_STD, WARP>(A, bwd, arena, block, grid, smem, s);
    case MK_UNIV: return launch_coop_kind<MK_UNIV, WARP>(A, bwd, arena, block, grid, smem, s);
    case MK_STEADY: return launch_coop_kind<MK_STEADY, WARP>(A, bwd, arena, block, grid, smem, s);
    case MK_CHOLS: return launch_coop_kind<MK_CHOLS, WARP>(A, bwd, arena, block, grid, smem, s);
    default: return cudaErrorInvalidValue;
  }
}

// Warp-per-unit for small systems (a 32-lane warp covers the m*m outputs of a product in a few rounds and 4+ arenas
// fit in one SM's shared memory); one 256-thread CTA per unit from k_states = 16 up (>= 256 outputs per product).
cudaError_t launch_coop(const KfArgs& A, bool bwd, cudaStream_t s) {
  int arena = coop_arena_doubles(A.m, A.p, bwd);
  arena = (arena + 1) & ~1;  // keep every arena 16-byte aligned
  const size_t arena_bytes = (size_t)arena * sizeof(double);
  const size_t smem_max = 227 * 1024;
  if (arena_bytes > smem_max) return cudaErrorInvalidConfiguration;
  if (arena_bytes * 4 <= smem_max && A.m < 16) {
    int warps = 4;
    // more warps per CTA only helps when the arena is tiny; keep CTAs small so many are resident
    const int block = warps * 32;
    const unsigned grid = (unsigned)((A.U + warps - 1) / warps);
    return launch_coop_mode<true>(A, bwd, arena, block, grid, arena_bytes * warps, s);
  }
  const int block = 256;
  return launch_coop_mode<false>(A, bwd, arena, block, (unsigned)A.U, arena_bytes, s);
}

cudaError_t launch_smoother(const SmoothArgs& S, cudaStream_t s) {
  const int arena = (smoother_arena_doubles(S.m) + 1) & ~1;
  const size_t arena_bytes = (size_t)arena * sizeof(double);
  const size_t smem_max = 227 * 1024;
  if (arena_bytes > smem_max) return cudaErrorInvalidConfiguration;
  if (arena_bytes * 4 <= smem_max && S.m < 16) {
    const int warps = 4;
    const size_t smem = arena_bytes * warps;
    cudaFuncSetAttribute(kf_smoother_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kf_smoother_kernel<true><<<(unsigned)((S.U + warps - 1) / warps), warps * 32, smem, s>>>(S, arena);
  } else {
    cudaFuncSetAttribute(kf_smoother_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)arena_bytes);
    kf_smoother_kernel<false><<<(unsigned)S.U, 256, arena_bytes, s>>>(S, arena);
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_dare(const DareArgs& D, bool bwd, cudaStream_t s) {
  int arena = (dare_arena_doubles(D.m, D.p) + 1) & ~1;
  const size_t arena_bytes = (size_t)arena * sizeof(double);
  const size_t smem_max = 227 * 1024;
  if (arena_bytes > smem_max) return cudaErrorInvalidConfiguration;
  const long long count = bwd ? D.U : D.nD;
  if (arena_bytes * 4 <= smem_max) {
    const int warps = 4;
    const unsigned grid = (unsigned)((count + warps - 1) / warps);
    const size_t smem = arena_bytes * warps;
    if (bwd) {
      cudaFuncSetAttribute(kf_dare_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kf_dare_kernel<true, true><<<grid, warps * 32, smem, s>>>(D, arena);
    } else {
      cudaFuncSetAttribute(kf_dare_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kf_dare_kernel<false, true><<<grid, warps * 32, smem, s>>>(D, arena);
    }
  } else {
    if (bwd) {
      cudaFuncSetAttribute(kf_dare_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)arena_bytes);
      kf_dare_kernel<true, false><<<(unsigned)count, 256, arena_bytes, s>>>(D, arena);
    } else {
      cudaFuncSetAttribute(kf_dare_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)arena_bytes);
      kf_dare_kernel<false, false><<<(unsigned)count, 256, arena_bytes, s>>>(D, arena);
    }
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace kfb
