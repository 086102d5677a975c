// MLP neural-ODE vector field (notebooks/jai/nn.ipynb cell 4: Linear(2,H) - ELU - Linear(H,H) - ELU - Linear(H,2)),
// one independent set of weights per particle / chain, plugged into the same fused solver template as the npde field.
//
// Mapping: one WARP owns one (particle, trajectory) pair (G = 32 lanes), a CTA owns one particle (N warps).  Lane l owns
// hidden units {l, l+32, ...}: its rows of W1/b1/b2, its columns of W3 and its rows of the W2-gradient accumulator live
// in registers; the particle's W2 (H x H) sits in shared memory with pitch H+1 so that both the row access of the
// forward product and the column access of the transposed product are bank-conflict free.  The ODE state and adjoint
// are replicated in every lane of the warp.
// theta layout per particle (PyTorch parameters() order of the notebook's nn.Sequential):
//   [ W1 (H x 2) | b1 (H) | W2 (H x H) | b2 (H) | W3 (2 x H) | b3 (2) ]          d = H^2 + 6H + 2
// Closure (nn.ipynb cell 10 bayesian_closure): sum_rows sum (X - x)^2 + reg * sum theta^2.
#pragma once
#include "npde_sep.cuh"

namespace bode {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int H>
struct MlpField {
  static constexpr int G = 32;
  static constexpr int MAX_THREADS = 256;
  static constexpr int R = (H + 31) / 32;      // hidden units per lane
  static constexpr int PITCH = H + 1;
  static constexpr int D = H * H + 6 * H + 2;
  static constexpr int DP = (D + 3) / 4 * 4;
  static constexpr int oW1 = 0, ob1 = 2 * H, oW2 = 3 * H, ob2 = 3 * H + H * H, oW3 = 4 * H + H * H, ob3 = 6 * H + H * H;

  float W1[R][2], b1[R], b2[R], W3[2][R], b3[2];
  float gW1[R][2], gb1[R], gb2[R], gW3[2][R], gb3[2];
  float gW2[R][H];
  const float* W2s;     // shared: [H][PITCH]
  float* hs;            // shared, per warp: [2][H]  (hidden activations | layer-2 cotangent)
  int lane;

  // shared-memory carve-up for one CTA (ppc == 1)
  static __host__ __device__ constexpr int smem_floats(int N) { return DP + H * PITCH + 3 + DP + N * 2 * H + 2 * N + 8; }
  static __device__ __forceinline__ float* s_theta(float* sm) { return sm; }
  static __device__ __forceinline__ float* s_W2(float* sm) { return sm + DP; }
  static __device__ __forceinline__ float* s_gacc(float* sm) { return sm + DP + ((H * PITCH + 3) / 4) * 4; }
  static __device__ __forceinline__ float* s_hs(float* sm) { return s_gacc(sm) + DP; }
  static __device__ __forceinline__ float* s_red(float* sm, int N) { return s_hs(sm) + N * 2 * H; }

  static __device__ __forceinline__ void prologue(const NpdeKParams& prm, float* sm) {
    const int p = blockIdx.x;
    float* th = s_theta(sm);
    float* w2 = s_W2(sm);
    for (int i = threadIdx.x; i < D; i += blockDim.x) th[i] = (p < prm.P) ? __ldg(prm.U + (long long)p * prm.U_stride + i) : 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) w2[(i / H) * PITCH + (i % H)] = th[oW2 + i];
    __syncthreads();
  }

  static __device__ __forceinline__ float2 lik_weight(const NpdeKParams& prm, int) { return f2(2.f * prm.lik_w, 2.f * prm.lik_w); }

  __device__ __forceinline__ void load(const NpdeKParams&, float* sm, int, int pairl, int lane_) {
    lane = lane_;
    const float* th = s_theta(sm);
    W2s = s_W2(sm);
    hs = s_hs(sm) + pairl * 2 * H;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = lane + 32 * r;
      const bool ok = row < H;
      W1[r][0] = ok ? th[oW1 + 2 * row] : 0.f;
      W1[r][1] = ok ? th[oW1 + 2 * row + 1] : 0.f;
      b1[r] = ok ? th[ob1 + row] : 0.f;
      b2[r] = ok ? th[ob2 + row] : 0.f;
      W3[0][r] = ok ? th[oW3 + row] : 0.f;
      W3[1][r] = ok ? th[oW3 + H + row] : 0.f;
    }
    b3[0] = th[ob3];
    b3[1] = th[ob3 + 1];
  }

  __device__ __forceinline__ void zero_grad() {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      gW1[r][0] = gW1[r][1] = gb1[r] = gb2[r] = gW3[0][r] = gW3[1][r] = 0.f;
#pragma unroll
      for (int k = 0; k < H; ++k) gW2[r][k] = 0.f;
    }
    gb3[0] = gb3[1] = 0.f;
  }

  static __device__ __forceinline__ float elu(float z) { return z > 0.f ? z : ex2(z * 1.4426950408889634f) - 1.f; }
  static __device__ __forceinline__ float delu(float z, float h) { return z > 0.f ? 1.f : h + 1.f; }

  // hidden layers for state x; leaves h1 in hs[0][.]
  __device__ __forceinline__ void hidden(float2 x, float (&z1)[R], float (&h1)[R], float (&z2)[R], float (&h2)[R]) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      z1[r] = fmaf(W1[r][0], x.x, fmaf(W1[r][1], x.y, b1[r]));
      h1[r] = elu(z1[r]);
      const int row = lane + 32 * r;
      if (row < H) hs[row] = h1[r];
      z2[r] = b2[r];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < H; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(hs + k);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = min(lane + 32 * r, H - 1);
        const float* w = W2s + row * PITCH + k;
        z2[r] = fmaf(w[0], hv.x, z2[r]);
        z2[r] = fmaf(w[1], hv.y, z2[r]);
        z2[r] = fmaf(w[2], hv.z, z2[r]);
        z2[r] = fmaf(w[3], hv.w, z2[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) h2[r] = (lane + 32 * r < H) ? elu(z2[r]) : 0.f;
  }

  __device__ __forceinline__ float2 output(const float (&h2)[R]) const {
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      o0 = fmaf(W3[0][r], h2[r], o0);
      o1 = fmaf(W3[1][r], h2[r], o1);
    }
    return f2(warp_sum(o0) + b3[0], warp_sum(o1) + b3[1]);
  }

  __device__ __forceinline__ float2 eval(const NpdeKParams&, float2 x) const {
    float z1[R], h1[R], z2[R], h2[R];
    hidden(x, z1, h1, z2, h2);
    const float2 o = output(h2);
    __syncwarp();                      // hs is rewritten by the next evaluation
    return o;
  }

  // J(x)^T a; parameter cotangents accumulate with weight wg.  WITH_F also returns f(x).
  template <bool WITH_F>
  __device__ __forceinline__ float2 vjp(const NpdeKParams&, float2 x, float2 a, float wg, float2* fout) {
    float z1[R], h1[R], z2[R], h2[R];
    hidden(x, z1, h1, z2, h2);
    if (WITH_F) *fout = output(h2);
    float* gs = hs + H;
    float gz2[R];
    const float aw0 = a.x * wg, aw1 = a.y * wg;
    gb3[0] += aw0;
    gb3[1] += aw1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float gh2 = fmaf(W3[0][r], a.x, W3[1][r] * a.y);
      gW3[0][r] = fmaf(aw0, h2[r], gW3[0][r]);
      gW3[1][r] = fmaf(aw1, h2[r], gW3[1][r]);
      gz2[r] = (lane + 32 * r < H) ? gh2 * delu(z2[r], h2[r]) : 0.f;
      gb2[r] = fmaf(wg, gz2[r], gb2[r]);
      const int row = lane + 32 * r;
      if (row < H) gs[row] = gz2[r];
    }
    __syncwarp();
    // gW2[row][k] += wg * gz2[row] * h1[k]      and      gh1[kk] = sum_row W2[row][kk] gz2[row]
    float gh1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) gh1[r] = 0.f;
#pragma unroll
    for (int k = 0; k < H; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(hs + k);
      const float4 gv = *reinterpret_cast<const float4*>(gs + k);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float gw = wg * gz2[r];
        gW2[r][k + 0] = fmaf(gw, hv.x, gW2[r][k + 0]);
        gW2[r][k + 1] = fmaf(gw, hv.y, gW2[r][k + 1]);
        gW2[r][k + 2] = fmaf(gw, hv.z, gW2[r][k + 2]);
        gW2[r][k + 3] = fmaf(gw, hv.w, gW2[r][k + 3]);
        const int kk = min(lane + 32 * r, H - 1);
        gh1[r] = fmaf(W2s[(k + 0) * PITCH + kk], gv.x, gh1[r]);
        gh1[r] = fmaf(W2s[(k + 1) * PITCH + kk], gv.y, gh1[r]);
        gh1[r] = fmaf(W2s[(k + 2) * PITCH + kk], gv.z, gh1[r]);
        gh1[r] = fmaf(W2s[(k + 3) * PITCH + kk], gv.w, gh1[r]);
      }
    }
    float ax = 0.f, ay = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float gz1 = (lane + 32 * r < H) ? gh1[r] * delu(z1[r], h1[r]) : 0.f;
      const float gw = wg * gz1;
      gW1[r][0] = fmaf(gw, x.x, gW1[r][0]);
      gW1[r][1] = fmaf(gw, x.y, gW1[r][1]);
      gb1[r] += gw;
      ax = fmaf(W1[r][0], gz1, ax);
      ay = fmaf(W1[r][1], gz1, ay);
    }
    ax = warp_sum(ax);
    ay = warp_sum(ay);
    __syncwarp();
    return f2(ax, ay);
  }

  // add this warp's parameter cotangents into the CTA accumulator (called by one warp at a time)
  __device__ __forceinline__ void add_grad(float* ga, bool first) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = lane + 32 * r;
      if (row >= H) continue;
      auto put = [&](int i, float v) { ga[i] = first ? v : ga[i] + v; };
      put(oW1 + 2 * row, gW1[r][0]);
      put(oW1 + 2 * row + 1, gW1[r][1]);
      put(ob1 + row, gb1[r]);
      put(ob2 + row, gb2[r]);
      put(oW3 + row, gW3[0][r]);
      put(oW3 + H + row, gW3[1][r]);
#pragma unroll
      for (int k = 0; k < H; ++k) put(oW2 + row * H + k, gW2[r][k]);
    }
    if (lane == 0) {
      ga[ob3] = first ? gb3[0] : ga[ob3] + gb3[0];
      ga[ob3 + 1] = first ? gb3[1] : ga[ob3 + 1] + gb3[1];
    }
  }

  // sum over trajectories in a fixed order, add the prior gradient 2 reg theta, closure values
  template <int INJ>
  static __device__ __forceinline__ void epilogue(const NpdeKParams& prm, float* sm, const MlpField& fld, bool active, int pl, int n,
                                                  int pairl, int lane_, float r2x, float r2y) {
    float* th = s_theta(sm);
    float* ga = s_gacc(sm);
    float* red = s_red(sm, prm.N);
    const int p = blockIdx.x;
    for (int nn = 0; nn < prm.N; ++nn) {
      __syncthreads();
      if (active && n == nn) fld.add_grad(ga, nn == 0);
    }
    if (active && lane_ == 0) { red[2 * n] = r2x; red[2 * n + 1] = r2y; }
    __syncthreads();
    if (p >= prm.P) return;
    const float reg2 = prm.add_prior ? 2.f * prm.reg : 0.f;
    float ss = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      const float t = th[i];
      prm.gU[(long long)p * prm.gU_stride + i] = prm.scale * fmaf(reg2, t, ga[i]);
      ss = fmaf(t, t, ss);
    }
    if (INJ == INJ_LIK) {
      ss = warp_sum(ss);
      __syncthreads();
      float* tmp = ga;                                   // reuse as scratch for the cross-warp sum
      if ((threadIdx.x & 31) == 0) tmp[threadIdx.x >> 5] = ss;
      __syncthreads();
      if (threadIdx.x == 0) {
        float tot = 0.f, sq = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += tmp[w];
        for (int nn = 0; nn < prm.N; ++nn) sq += red[2 * nn] + red[2 * nn + 1];
        prm.loss[p] = prm.scale * (prm.lik_w * sq + (prm.add_prior ? prm.reg * tot : 0.f));
        prm.sqerr[p] = sq;
      }
    }
  }
};

}  // namespace bode
