// MLP neural-ODE field (nn.ipynb cell 4, H = 64) with the 64 x 64 hidden layer on the TENSOR CORES.
//
// The N trajectories of a particle share its weights, so the three W2 products of one RHS evaluation / VJP are small GEMMs over
// the trajectory batch (padded to the MMA's N = 8):
//     forward    Z2 [64 x 8]  = W2   [64 x 64] . H1 [64 x 8]          (hidden(): every evaluation, and the recomputation inside a VJP)
//     backward   G1 [64 x 8]  = W2^T [64 x 64] . G2 [64 x 8]          (cotangent of h1)
//                gW2[64 x 64] += wg G2 [64 x 8] . H1^T [8 x 64]       (summed over the trajectories by the contraction itself)
// issued as mma.sync.m16n8k8 tf32 with the 3xTF32 split (x = hi + lo, hi = top 19 bits: A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, ~2^-21
// relative like the SVGD kernels), fp32 accumulation.  A CTA is one particle, warp n is trajectory n for all element-wise work
// (lane l owns hidden units l and l + 32, exactly as MlpField), and warps 0..3 double as the MMA warps: warp w owns the 16-row tile
// w of every product -- its slice of the gW2 accumulator lives in registers (32 instead of the 128 per lane that made the FP32-pipe
// kernel spill); the W2 and W2^T fragments sit pre-split (hi / lo) in shared memory in fragment order, one conflict-free LDS.128
// per MMA operand (keeping them in registers cost 64 registers and held the kernel to one CTA per SM).
// Activations travel through shared memory in [trajectory][unit] rows of pitch 68 (conflict-free for the writers and for the
// fragment loads): hidden() = 2 CTA barriers, a VJP = 4.
//
// LOCK-STEP REQUIREMENT: every warp of the CTA must evaluate the field the same number of times -- true for the fixed-grid solvers
// and for dopri5 with the pooled controller (one controller per particle, Dopri5Params::pool); the per-pair controller keeps
// MlpField.  Needs 4 <= N <= 8 (four MMA warps, MMA N = 8).  Same interface as MlpField, so the solver templates are unchanged.
#pragma once
#include "mlp_field.cuh"

namespace bode {

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split3(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

struct MlpTcField {
  static constexpr int H = 64;
  static constexpr int G = 32;
  static constexpr int MAX_THREADS = 256;
  static constexpr int R = 2;
  static constexpr int D = H * H + 6 * H + 2;
  static constexpr int DP = (D + 3) / 4 * 4;
  static constexpr int oW1 = 0, ob1 = 2 * H, oW2 = 3 * H, ob2 = 3 * H + H * H, oW3 = 4 * H + H * H, ob3 = 6 * H + H * H;
  static constexpr int BP = 68;                 // staging pitch: rows = trajectory (8), columns = hidden unit (64)
  static constexpr int STG = 8 * BP;            // one staging array
  static constexpr int WTF = 4 * 8 * 32 * 4;    // W2 or W2^T fragments, one of (hi, lo): [tile][kstep][lane][4]

  // element-wise role
  float W1[R][2], b1[R], b2[R], W3[2][R], b3[2];
  float gW1[R][2], gb1[R], gb2[R], gW3[2][R], gb3[2];
  // MMA role (warps 0..3)
  float gW2c[8][4];
  float *Bh, *Bl, *Gh, *Gl, *Zs;
  const float *WFh, *WFl, *WTh, *WTl;
  int lane, warp;

  // shared-memory carve-up for one CTA (ppc == 1)
  static __host__ __device__ constexpr int smem_floats(int N) { return 2 * DP + 5 * STG + 4 * WTF + 2 * N + 8; }
  static __device__ __forceinline__ float* s_theta(float* sm) { return sm; }
  static __device__ __forceinline__ float* s_gacc(float* sm) { return sm + DP; }
  static __device__ __forceinline__ float* s_stage(float* sm) { return sm + 2 * DP; }
  static __device__ __forceinline__ float* s_wt(float* sm) { return sm + 2 * DP + 5 * STG; }
  static __device__ __forceinline__ float* s_red(float* sm, int) { return sm + 2 * DP + 5 * STG + 4 * WTF; }

  static __device__ __forceinline__ void prologue(const NpdeKParams& prm, float* sm) {
    const int p = blockIdx.x;
    float* th = s_theta(sm);
    for (int i = threadIdx.x; i < D; i += blockDim.x) th[i] = (p < prm.P) ? __ldg(prm.U + (long long)p * prm.U_stride + i) : 0.f;
    float* st = s_stage(sm);
    for (int i = threadIdx.x; i < 5 * STG; i += blockDim.x) st[i] = 0.f;       // rows of absent trajectories (n >= N) stay zero
    __syncthreads();
    // A fragments, pre-split: product 1 A(row = i, col = k) = W2[i][k]; product 2 A(row = k, col = i) = W2[i][k]
    float* wfh = s_wt(sm);
    float* wfl = wfh + WTF;
    float* wth = wfl + WTF;
    float* wtl = wth + WTF;
    for (int idx = threadIdx.x; idx < 4 * 8 * 32; idx += blockDim.x) {
      const int ln = idx & 31, ks = (idx >> 5) & 7, w = idx >> 8;
      const int g = ln >> 2, t = ln & 3;
      const int rr[4] = {16 * w + g, 16 * w + g + 8, 16 * w + g, 16 * w + g + 8};
      const int cc[4] = {8 * ks + t, 8 * ks + t, 8 * ks + t + 4, 8 * ks + t + 4};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hi, lo;
        split3(th[oW2 + rr[j] * H + cc[j]], hi, lo);
        wfh[idx * 4 + j] = __uint_as_float(hi);
        wfl[idx * 4 + j] = __uint_as_float(lo);
        split3(th[oW2 + cc[j] * H + rr[j]], hi, lo);
        wth[idx * 4 + j] = __uint_as_float(hi);
        wtl[idx * 4 + j] = __uint_as_float(lo);
      }
    }
    __syncthreads();
  }

  static __device__ __forceinline__ float2 lik_weight(const NpdeKParams& prm, int) { return f2(2.f * prm.lik_w, 2.f * prm.lik_w); }

  __device__ __forceinline__ void load(const NpdeKParams&, float* sm, int, int pairl, int lane_) {
    lane = lane_;
    warp = pairl;                                 // ppc == 1, G == 32: the pair index inside the CTA is the warp = the trajectory
    const float* th = s_theta(sm);
    float* st = s_stage(sm);
    Bh = st; Bl = st + STG; Gh = st + 2 * STG; Gl = st + 3 * STG; Zs = st + 4 * STG;
    WFh = s_wt(sm); WFl = WFh + WTF; WTh = WFl + WTF; WTl = WTh + WTF;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = lane + 32 * r;
      W1[r][0] = th[oW1 + 2 * row];
      W1[r][1] = th[oW1 + 2 * row + 1];
      b1[r] = th[ob1 + row];
      b2[r] = th[ob2 + row];
      W3[0][r] = th[oW3 + row];
      W3[1][r] = th[oW3 + H + row];
    }
    b3[0] = th[ob3];
    b3[1] = th[ob3 + 1];
  }

  __device__ __forceinline__ void zero_grad() {
#pragma unroll
    for (int r = 0; r < R; ++r) gW1[r][0] = gW1[r][1] = gb1[r] = gb2[r] = gW3[0][r] = gW3[1][r] = 0.f;
    gb3[0] = gb3[1] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) gW2c[nt][0] = gW2c[nt][1] = gW2c[nt][2] = gW2c[nt][3] = 0.f;
    warp = 8;                                     // until load(): no MMA role
  }

  static __device__ __forceinline__ float elu(float z) { return z > 0.f ? z : ex2(z * 1.4426950408889634f) - 1.f; }
  static __device__ __forceinline__ float delu(float z, float h) { return z > 0.f ? 1.f : h + 1.f; }

  // one 16 x 8 output tile = sum over 64 k of A (pre-split fragments in shared memory) times the staged activations B[n][k]
  __device__ __forceinline__ void tile_product(const float* fh, const float* fl, const float* bh, const float* bl, float (&d)[4]) const {
    const int g = lane >> 2, t = lane & 3;
    d[0] = d[1] = d[2] = d[3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t b0h = __float_as_uint(bh[g * BP + 8 * ks + t]), b1h = __float_as_uint(bh[g * BP + 8 * ks + t + 4]);
      const uint32_t b0l = __float_as_uint(bl[g * BP + 8 * ks + t]), b1l = __float_as_uint(bl[g * BP + 8 * ks + t + 4]);
      const float4 vh = *reinterpret_cast<const float4*>(fh + ((warp * 8 + ks) * 32 + lane) * 4);
      const float4 vl = *reinterpret_cast<const float4*>(fl + ((warp * 8 + ks) * 32 + lane) * 4);
      const uint32_t ah[4] = {__float_as_uint(vh.x), __float_as_uint(vh.y), __float_as_uint(vh.z), __float_as_uint(vh.w)};
      const uint32_t al[4] = {__float_as_uint(vl.x), __float_as_uint(vl.y), __float_as_uint(vl.z), __float_as_uint(vl.w)};
      mma_tf32(d, al, b0h, b1h);
      mma_tf32(d, ah, b0l, b1l);
      mma_tf32(d, ah, b0h, b1h);
    }
  }
  __device__ __forceinline__ void store_tile(const float (&d)[4]) const {
    const int g = lane >> 2, t = lane & 3;
    Zs[(2 * t) * BP + 16 * warp + g] = d[0];
    Zs[(2 * t + 1) * BP + 16 * warp + g] = d[1];
    Zs[(2 * t) * BP + 16 * warp + g + 8] = d[2];
    Zs[(2 * t + 1) * BP + 16 * warp + g + 8] = d[3];
  }
  __device__ __forceinline__ void stage(float* dh, float* dl, int n, const float (&v)[R]) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      uint32_t hi, lo;
      split3(v[r], hi, lo);
      dh[n * BP + lane + 32 * r] = __uint_as_float(hi);
      dl[n * BP + lane + 32 * r] = __uint_as_float(lo);
    }
  }

  // hidden layers for state x (trajectory n = warp); leaves H1 staged in Bh / Bl for the VJP's weight-gradient product
  __device__ __forceinline__ void hidden(int n, float2 x, float (&z1)[R], float (&h1)[R], float (&z2)[R], float (&h2)[R]) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      z1[r] = fmaf(W1[r][0], x.x, fmaf(W1[r][1], x.y, b1[r]));
      h1[r] = elu(z1[r]);
    }
    stage(Bh, Bl, n, h1);
    __syncthreads();
    if (warp < 4) {
      float d[4];
      tile_product(WFh, WFl, Bh, Bl, d);
      store_tile(d);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      z2[r] = b2[r] + Zs[n * BP + lane + 32 * r];
      h2[r] = elu(z2[r]);
    }
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
    hidden(warp, x, z1, h1, z2, h2);
    return output(h2);
  }

  // J(x)^T a; parameter cotangents accumulate with weight wg.  WITH_F also returns f(x).
  template <bool WITH_F>
  __device__ __forceinline__ float2 vjp(const NpdeKParams&, float2 x, float2 a, float wg, float2* fout) {
    const int n = warp;
    float z1[R], h1[R], z2[R], h2[R];
    hidden(n, x, z1, h1, z2, h2);
    if (WITH_F) *fout = output(h2);
    float gz2[R];
    const float aw0 = a.x * wg, aw1 = a.y * wg;
    gb3[0] += aw0;
    gb3[1] += aw1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float gh2 = fmaf(W3[0][r], a.x, W3[1][r] * a.y);
      gW3[0][r] = fmaf(aw0, h2[r], gW3[0][r]);
      gW3[1][r] = fmaf(aw1, h2[r], gW3[1][r]);
      gz2[r] = gh2 * delu(z2[r], h2[r]);
      gb2[r] = fmaf(wg, gz2[r], gb2[r]);
    }
    stage(Gh, Gl, n, gz2);
    __syncthreads();
    if (warp < 4) {
      const int g = lane >> 2, t = lane & 3;
      // G1 = W2^T G2
      float d[4];
      tile_product(WTh, WTl, Gh, Gl, d);
      store_tile(d);
      // gW2[16 w + ., :] += wg G2[16 w + ., n] H1[:, n]^T    (A = G2 rows of this tile over the trajectories, K = n)
      uint32_t ah[4], al[4];
      {
        const int ro[4] = {16 * warp + g, 16 * warp + g + 8, 16 * warp + g, 16 * warp + g + 8};
        const int no[4] = {t, t, t + 4, t + 4};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = Gh[no[j] * BP + ro[j]] + Gl[no[j] * BP + ro[j]];          // hi + lo == the staged value exactly
          if (wg != 1.f) v *= wg;
          split3(v, ah[j], al[j]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t b0h = __float_as_uint(Bh[t * BP + 8 * nt + g]), b1h = __float_as_uint(Bh[(t + 4) * BP + 8 * nt + g]);
        const uint32_t b0l = __float_as_uint(Bl[t * BP + 8 * nt + g]), b1l = __float_as_uint(Bl[(t + 4) * BP + 8 * nt + g]);
        mma_tf32(gW2c[nt], al, b0h, b1h);
        mma_tf32(gW2c[nt], ah, b0l, b1l);
        mma_tf32(gW2c[nt], ah, b0h, b1h);
      }
    }
    __syncthreads();
    float ax = 0.f, ay = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float gz1 = Zs[n * BP + lane + 32 * r] * delu(z1[r], h1[r]);
      const float gw = wg * gz1;
      gW1[r][0] = fmaf(gw, x.x, gW1[r][0]);
      gW1[r][1] = fmaf(gw, x.y, gW1[r][1]);
      gb1[r] += gw;
      ax = fmaf(W1[r][0], gz1, ax);
      ay = fmaf(W1[r][1], gz1, ay);
    }
    return f2(warp_sum(ax), warp_sum(ay));
  }

  // add this warp's cotangents of everything but W2 into the CTA accumulator (called by one warp at a time)
  __device__ __forceinline__ void add_grad(float* ga, bool first) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = lane + 32 * r;
      auto put = [&](int i, float v) { ga[i] = first ? v : ga[i] + v; };
      put(oW1 + 2 * row, gW1[r][0]);
      put(oW1 + 2 * row + 1, gW1[r][1]);
      put(ob1 + row, gb1[r]);
      put(ob2 + row, gb2[r]);
      put(oW3 + row, gW3[0][r]);
      put(oW3 + H + row, gW3[1][r]);
    }
    if (lane == 0) {
      ga[ob3] = first ? gb3[0] : ga[ob3] + gb3[0];
      ga[ob3 + 1] = first ? gb3[1] : ga[ob3 + 1] + gb3[1];
    }
  }
  // the MMA warps hold gW2 already summed over the trajectories
  __device__ __forceinline__ void store_gW2(float* ga) const {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      ga[oW2 + (16 * warp + g) * H + 8 * nt + 2 * t] = gW2c[nt][0];
      ga[oW2 + (16 * warp + g) * H + 8 * nt + 2 * t + 1] = gW2c[nt][1];
      ga[oW2 + (16 * warp + g + 8) * H + 8 * nt + 2 * t] = gW2c[nt][2];
      ga[oW2 + (16 * warp + g + 8) * H + 8 * nt + 2 * t + 1] = gW2c[nt][3];
    }
  }

  // sum over trajectories in a fixed order, add the prior gradient 2 reg theta, closure values
  template <int INJ>
  static __device__ __forceinline__ void epilogue(const NpdeKParams& prm, float* sm, const MlpTcField& fld, bool active, int pl, int n,
                                                  int pairl, int lane_, float r2x, float r2y) {
    float* th = s_theta(sm);
    float* ga = s_gacc(sm);
    float* red = s_red(sm, prm.N);
    const int p = blockIdx.x;
    for (int nn = 0; nn < prm.N; ++nn) {
      __syncthreads();
      if (active && n == nn) fld.add_grad(ga, nn == 0);
    }
    if (active && pairl < 4) fld.store_gW2(ga);
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
