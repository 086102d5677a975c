// npde vector field on a TENSOR GRID of inducing points too large for one thread (7x7 .. 16x16; BASELINE config 5: 16 x 16):
// row-sliced separable kernel.
//
// The squared-exponential kernel factorises over the grid axes, k_ab(x) = kx_a(x0) ky_b(x1) with Z[a*My+b] = (gx[a], gy[b])
// (gp.py:315-318), so an evaluation needs Mx + My exponentials instead of Mx*My and one multiply-add per inducing point instead
// of the five instructions of the general-Z kernel (npde_gen.cuh).  SIXTEEN lanes own one (particle, trajectory) pair -- two
// pairs per warp: lane a holds row a of W (My float2 values, as packed pairs for FFMA2) and of the gradient accumulator, computes
// kx_a and -- acting as column index b = a -- ky_b and ky_b (u1 - gy_b); the column factors are exchanged through a double-buffered
// shared-memory line (one store, one group barrier, My/4 vector loads, instead of My shuffles) and the sums over a are completed
// with a 4-round butterfly inside the half-warp.  Per lane and evaluation (My = 16): ~95 SASS instructions forward and ~120 reverse
// for TWO pairs; the general-Z kernel needs 68 / 120-140 for one.
//   f(x)        = sum_a kx_a sum_b ky_b W_ab
//   (J^T a)_0   = -k0 sum_a kx_a dx_a sum_b ky_b (a . W_ab)          dx_a = c0 (x0 - gx_a)
//   (J^T a)_1   = -k1 sum_a kx_a      sum_b ky_b dy_b (a . W_ab)     dy_b = c1 (x1 - gy_b)
//   gW_ab      += wg a kx_a ky_b
#pragma once
#include "npde_gen.cuh"

namespace bode {

__device__ __forceinline__ float* row_xch() {
  __shared__ __align__(16) float x[2 * 2 * 256];   // [parity][ky | ky*dy][thread]
  return x;
}

template <int MY>
struct RowField {
  static constexpr int G = 16;
  static constexpr int MAX_THREADS = 160;   // N <= 10 trajectories (two particles per CTA for odd N <= 5)
#ifndef BODE_ROW_MIN_BLOCKS
#define BODE_ROW_MIN_BLOCKS 3   // measured on c5 (16 x 16, 2048 particles): 3 -> 128 registers, 0.423 ms; 4 -> 96 + spills, 0.497; 5 -> 1.0
#endif
  static constexpr int MIN_BLOCKS = BODE_ROW_MIN_BLOCKS;
  f32x2 W[MY], gW[MY];   // row a = lane of W / gW, (x, y) components packed; zero beyond the grid
  float gxa, gyb;        // c0 gx[lane], c1 gy[lane]
  unsigned mask;         // the 16 lanes of this pair
  int a, mx, my;
  mutable int par;

  static __device__ __forceinline__ void prologue(const NpdeKParams& prm, float* smem) { GenField<1>::prologue(prm, smem); }
  static __device__ __forceinline__ float2 lik_weight(const NpdeKParams& prm, int p) { return GenField<1>::lik_weight(prm, p); }
  template <int INJ>
  static __device__ __forceinline__ void epilogue(const NpdeKParams& prm, float* smem, const RowField& fld, bool active, int pl, int n,
                                                  int pairl, int lane_, float r2x, float r2y);

  __device__ __forceinline__ void load(const NpdeKParams& prm, const float* smem, int pl, int, int lane_) {
    a = lane_;
    mx = prm.gmx;
    my = prm.gmy;
    par = 0;
    mask = 0xffffu << (threadIdx.x & 16);
    const float* Wp = smem + (prm.ppc + pl) * 2 * prm.m;
#pragma unroll
    for (int b = 0; b < MY; ++b) {
      const bool ok = a < mx && b < my;
      const int j = a * my + b;
      W[b] = ok ? pk(Wp[2 * j], Wp[2 * j + 1]) : pk(0.f, 0.f);
    }
    gxa = a < mx ? prm.gxs[a] : 0.f;
    gyb = a < my ? prm.gys[a] : 0.f;
  }
  __device__ __forceinline__ void store_gW(const NpdeKParams&, float* gp, int) const {
#pragma unroll
    for (int b = 0; b < MY; ++b) {
      if (a < mx && b < my) {
        float g0, g1;
        upk(gW[b], g0, g1);
        gp[2 * (a * my + b)] = g0;
        gp[2 * (a * my + b) + 1] = g1;
      }
    }
  }
  __device__ __forceinline__ void zero_grad() {
#pragma unroll
    for (int b = 0; b < MY; ++b) gW[b] = pk(0.f, 0.f);
  }
  __device__ __forceinline__ float gsum(float v) const {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
  }

  __device__ __forceinline__ float2 eval(const NpdeKParams& prm, float2 x) const {
    const float dx = fmaf(prm.c0, x.x, -gxa), dy = fmaf(prm.c1, x.y, -gyb);
    const float kx = ex2(-dx * dx), kyo = ex2(-dy * dy);
    float* xb = row_xch() + par * 512;
    par ^= 1;
    xb[threadIdx.x] = kyo;
    __syncwarp(mask);
    const float4* kp = reinterpret_cast<const float4*>(xb + (threadIdx.x & ~15));
    f32x2 q0 = pk(0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;          // four independent chains
#pragma unroll
    for (int b4 = 0; b4 < MY / 4; ++b4) {
      const float4 k = kp[b4];
      q0 = fma2x(W[4 * b4 + 0], pk(k.x, k.x), q0);
      q1 = fma2x(W[4 * b4 + 1], pk(k.y, k.y), q1);
      q2 = fma2x(W[4 * b4 + 2], pk(k.z, k.z), q2);
      q3 = fma2x(W[4 * b4 + 3], pk(k.w, k.w), q3);
    }
    float qx, qy, rx, ry, sx, sy, tx, ty;
    upk(q0, qx, qy);
    upk(q1, rx, ry);
    upk(q2, sx, sy);
    upk(q3, tx, ty);
    return f2(gsum(kx * ((qx + rx) + (sx + tx))), gsum(kx * ((qy + ry) + (sy + ty))));
  }

  template <bool WITH_F>
  __device__ __forceinline__ float2 vjp(const NpdeKParams& prm, float2 x, float2 av, float wg, float2* fout) {
    const float dx = fmaf(prm.c0, x.x, -gxa), dy = fmaf(prm.c1, x.y, -gyb);
    const float kx = ex2(-dx * dx), kyo = ex2(-dy * dy);
    float* xb = row_xch() + par * 512;
    par ^= 1;
    xb[threadIdx.x] = kyo;
    xb[256 + threadIdx.x] = kyo * dy;
    __syncwarp(mask);
    const float4* kp = reinterpret_cast<const float4*>(xb + (threadIdx.x & ~15));
    const float4* dp = reinterpret_cast<const float4*>(xb + 256 + (threadIdx.x & ~15));
    const f32x2 kxaw = pk(kx * (av.x * wg), kx * (av.y * wg));
    // sum_b ky_b (a . W_ab) = a . q with q = sum_b ky_b W_ab (the forward contraction), likewise r = sum_b ky_b dy_b W_ab:
    // three FFMA2 per inducing point (q, r, gW) instead of five scalar instructions
    f32x2 q[2] = {pk(0.f, 0.f), pk(0.f, 0.f)}, r[2] = {pk(0.f, 0.f), pk(0.f, 0.f)};          // two independent chains each
#pragma unroll
    for (int b4 = 0; b4 < MY / 4; ++b4) {
      const float4 k4 = kp[b4], d4 = dp[b4];
      const float kk[4] = {k4.x, k4.y, k4.z, k4.w}, kd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = 4 * b4 + i;
        q[i & 1] = fma2x(W[b], pk(kk[i], kk[i]), q[i & 1]);
        r[i & 1] = fma2x(W[b], pk(kd[i], kd[i]), r[i & 1]);
        gW[b] = fma2x(kxaw, pk(kk[i], kk[i]), gW[b]);
      }
    }
    float qx, qy, q1x, q1y, rx, ry, r1x, r1y;
    upk(q[0], qx, qy);
    upk(q[1], q1x, q1y);
    upk(r[0], rx, ry);
    upk(r[1], r1x, r1y);
    qx += q1x; qy += q1y; rx += r1x; ry += r1y;
    if (WITH_F) *fout = f2(gsum(kx * qx), gsum(kx * qy));
    const float S1 = fmaf(av.y, qy, av.x * qx), S2 = fmaf(av.y, ry, av.x * rx);
    return f2(-prm.k0 * gsum(kx * dx * S1), -prm.k1 * gsum(kx * S2));
  }
};

}  // namespace bode
