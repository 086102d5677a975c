// npde vector field for ARBITRARY inducing locations Z (m <= 32*JPL): lane-sliced kernel.
//
// One WARP owns one (particle, trajectory) pair; lane l owns the inducing points j = l, l+32, ... (JPL of them): their
// locations, the projected values W_j = (A U)_j and the gradient accumulators live in that lane's registers, every lane
// evaluates its own kernel values and the two-component sums (f, J^T a) are completed with a shuffle butterfly.  This is
// the path for non-grid Z and for large grids (BASELINE config 5: 16 x 16 = 256 points -> JPL = 8), where the
// (particle, trajectory) count alone cannot fill the machine and W does not fit one thread's registers.
// Same maths as npde_sep.cuh without the factorisation: k_j(x) = 2^-( (c0 (x0 - z_j0))^2 + (c1 (x1 - z_j1))^2 ).
#pragma once
#include "npde_sep.cuh"

namespace bode {

__device__ __forceinline__ float warp_sum_gen(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int JPL>
struct GenField {
  static constexpr int G = 32;
  static constexpr int MAX_THREADS = 256;
  float2 W[JPL], gW[JPL], Zs[JPL];     // Zs = (c0 z_j0, c1 z_j1); lanes past m hold W = 0
  int lane;

  static __device__ __forceinline__ void prologue(const NpdeKParams& prm, float* smem) {
    // W = A U for the CTA's particles (same routine as the separable field)
    const int m = prm.m, m2 = 2 * m, nout = prm.ppc * m2;
    float* Us = smem;
    float* Ws = smem + nout;
    const int p0 = blockIdx.x * prm.ppc;
    if (prm.Wpre) {                                  // split mode: the projection was done for all particles by proj_W_kernel
      for (int idx = threadIdx.x; idx < nout; idx += blockDim.x) {
        const int q = idx / m2, r = idx - q * m2;
        Ws[idx] = (p0 + q < prm.P) ? __ldg(prm.Wpre + (long long)(p0 + q) * m2 + r) : 0.f;
      }
      __syncthreads();
      return;
    }
    for (int idx = threadIdx.x; idx < nout; idx += blockDim.x) {
      const int q = idx / m2, r = idx - q * m2;
      Us[idx] = (p0 + q < prm.P) ? __ldg(prm.U + (long long)(p0 + q) * prm.U_stride + r) : 0.f;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nout; idx += blockDim.x) {
      const int q = idx / m2, r = idx - q * m2, j = r >> 1, d = r & 1;
      const float* Uq = Us + q * m2 + d;
      float acc = 0.f;
      if (prm.AT) {                                  // threads run over j: AT[k][j] is one cache line per warp load, A[j][k] sixteen
        const float* Atj = prm.AT + j;
        for (int k = 0; k < m; ++k) acc = fmaf(__ldg(Atj + (long long)k * m), Uq[2 * k], acc);
      } else {
        const float* Aj = prm.A + (long long)j * m;
        for (int k = 0; k < m; ++k) acc = fmaf(__ldg(Aj + k), Uq[2 * k], acc);
      }
      Ws[idx] = acc;
    }
    __syncthreads();
  }
  static __device__ __forceinline__ float2 lik_weight(const NpdeKParams& prm, int p) {
    const float2 ls = *reinterpret_cast<const float2*>(prm.logsn + (long long)p * prm.logsn_stride);
    return f2(expf(-2.f * ls.x), expf(-2.f * ls.y));
  }
  template <int INJ>
  static __device__ __forceinline__ void epilogue(const NpdeKParams& prm, float* smem, const GenField& fld, bool active, int pl, int n,
                                                  int pairl, int lane_, float r2x, float r2y);

  __device__ __forceinline__ void load(const NpdeKParams& prm, const float* smem, int pl, int, int lane_) {
    lane = lane_;
    const float* Wp = smem + (prm.ppc + pl) * 2 * prm.m;
#pragma unroll
    for (int i = 0; i < JPL; ++i) {
      const int j = lane + 32 * i;
      const bool ok = j < prm.m;
      W[i] = ok ? f2(Wp[2 * j], Wp[2 * j + 1]) : f2(0.f, 0.f);
      Zs[i] = ok ? f2(prm.c0 * __ldg(prm.Z + 2 * j), prm.c1 * __ldg(prm.Z + 2 * j + 1)) : f2(0.f, 0.f);
    }
  }
  __device__ __forceinline__ void store_gW(const NpdeKParams& prm, float* gp, int) const {
#pragma unroll
    for (int i = 0; i < JPL; ++i) {
      const int j = lane + 32 * i;
      if (j < prm.m) { gp[2 * j] = gW[i].x; gp[2 * j + 1] = gW[i].y; }
    }
  }
  __device__ __forceinline__ void zero_grad() {
#pragma unroll
    for (int i = 0; i < JPL; ++i) gW[i] = f2(0.f, 0.f);
  }

  __device__ __forceinline__ float2 eval(const NpdeKParams& prm, float2 x) const {
    const float u0 = prm.c0 * x.x, u1 = prm.c1 * x.y;
    float2 f = f2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < JPL; ++i) {
      const float d0 = u0 - Zs[i].x, d1 = u1 - Zs[i].y;
      const float k = ex2(-fmaf(d1, d1, d0 * d0));
      f = fma2(k, W[i], f);
    }
    return f2(warp_sum_gen(f.x), warp_sum_gen(f.y));
  }

  template <bool WITH_F>
  __device__ __forceinline__ float2 vjp(const NpdeKParams& prm, float2 x, float2 a, float wg, float2* fout) {
    const float u0 = prm.c0 * x.x, u1 = prm.c1 * x.y;
    const float2 aw = f2(a.x * wg, a.y * wg);
    float sx = 0.f, sy = 0.f;
    float2 f = f2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < JPL; ++i) {
      const float d0 = u0 - Zs[i].x, d1 = u1 - Zs[i].y;
      const float k = ex2(-fmaf(d1, d1, d0 * d0));
      const float c = fmaf(a.y, W[i].y, a.x * W[i].x) * k;
      gW[i] = fma2(k, aw, gW[i]);
      sx = fmaf(c, d0, sx);
      sy = fmaf(c, d1, sy);
      if (WITH_F) f = fma2(k, W[i], f);
    }
    if (WITH_F) *fout = f2(warp_sum_gen(f.x), warp_sum_gen(f.y));
    return f2(-prm.k0 * warp_sum_gen(sx), -prm.k1 * warp_sum_gen(sy));
  }
};

}  // namespace bode
