// kf_ctx.cuh - execution contexts for kf_core.cuh (see the header comment there).
#pragma once
#include "kf_core.cuh"

namespace kfb {

// ---------------------------------------------------------------------------
// ThreadCtx: one unit per thread; compile-time dims; all buffers are registers.
// Tape layout [t-1][k][U]: consecutive threads (units) touch consecutive doubles -> every tape
// load/store of a warp is one fully-coalesced 256-byte transaction.
// ---------------------------------------------------------------------------
template <int N>
struct RegBuf {
  double v[N > 0 ? N : 1];
  template <class X>
  KFB_HD explicit RegBuf(X&) {}
  KFB_HD double& operator[](int i) { return v[i]; }
  KFB_HD const double& operator[](int i) const { return v[i]; }
};

template <int M, int P>
struct ThreadCtx {
  static constexpr bool TV = false;
  template <int SZ>
  using Buf = RegBuf<(SZ == SZ_M ? M : SZ == SZ_P ? P : SZ == SZ_MM ? M * M : SZ == SZ_MP ? M * P : SZ == SZ_PP ? P * P : M + (M * (M + 1)) / 2)>;
  const double* y_smem;  // observations staged in shared memory (shared y) or nullptr
  KFB_HD static constexpr int m() { return M; }
  KFB_HD static constexpr int p() { return P; }
  KFB_HD static constexpr int lane() { return 0; }
  KFB_HD static constexpr int G() { return 1; }
  KFB_HD void sync() const {}
  KFB_HD const double* y_base(const KfArgs& A, long long series) const {
    return y_smem ? y_smem : A.y.p + series * A.y.bs;
  }
  // tape entry of step t (t >= 1) = tape_base + (t-1) * tape_step; element k at [k * tape_elem]
  KFB_HD double* tape_base(const KfArgs& A, long long u) const { return A.tape + u; }
  KFB_HD long long tape_step(const KfArgs& A) const { return (long long)tape_width(M) * A.U; }
  KFB_HD long long tape_elem(const KfArgs& A) const { return A.U; }
};

// ---------------------------------------------------------------------------
// CoopCtx: G lanes cooperate on one unit; run-time dims; buffers are bump-allocated from a per-unit
// shared-memory arena.  G == 32: one warp per unit (sync = __syncwarp); G == blockDim.x: one CTA per
// unit (sync = __syncthreads).  Tape layout [u][t-1][k]: the G lanes read/write consecutive doubles.
// ---------------------------------------------------------------------------
struct CoopCtx {
  static constexpr bool TV = true;
  int m_, p_, lane_, G_;
  double* arena;
  int off, cap;
  bool overflow;

  KFB_HD int size_of(int sz) const {
    return sz == SZ_M ? m_ : sz == SZ_P ? p_ : sz == SZ_MM ? m_ * m_ : sz == SZ_MP ? m_ * p_ : sz == SZ_PP ? p_ * p_ : tape_width(m_);
  }
  KFB_HD double* bump(int cnt) {
    double* r = arena + off;
    off += cnt;
    if (off > cap) { overflow = true; r = arena; }
    return r;
  }
  template <int SZ>
  struct Buf {
    double* v;
    KFB_HD explicit Buf(CoopCtx& x) : v(x.bump(x.size_of(SZ))) {}
    KFB_HD double& operator[](int i) { return v[i]; }
    KFB_HD const double& operator[](int i) const { return v[i]; }
  };
  KFB_HD int m() const { return m_; }
  KFB_HD int p() const { return p_; }
  KFB_HD int lane() const { return lane_; }
  KFB_HD int G() const { return G_; }
  KFB_HD void sync() const {
#if defined(__CUDA_ARCH__)
    if (G_ <= 32) __syncwarp();
    else __syncthreads();
#endif
  }
  KFB_HD const double* y_base(const KfArgs& A, long long series) const { return A.y.p + series * A.y.bs; }
  KFB_HD double* tape_base(const KfArgs& A, long long u) const {
    return A.tape + u * (long long)(A.n - 1) * tape_width(m_);
  }
  KFB_HD long long tape_step(const KfArgs&) const { return tape_width(m_); }
  KFB_HD long long tape_elem(const KfArgs&) const { return 1; }
};

// doubles of arena one unit needs (upper bound of what forward_unit / backward_unit bump-allocate)
inline int coop_arena_doubles(int m, int p, bool backward) {
  const int mm = m * m, mp = m * p, pp = p * p;
  const int params = mm + mp + pp + p + pp;
  const int upd = 3 * p + 3 * mp + 4 * pp + 3 * mm;
  if (!backward) return params + 3 * mm + 5 * m + upd;
  return params + 8 * mm + 8 * m + 4 * mp + 4 * pp + 2 * p + upd + m + (m * (m + 1)) / 2;
}

}  // namespace kfb
