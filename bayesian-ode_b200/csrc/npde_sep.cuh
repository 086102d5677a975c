// npde vector field on a tensor (MX x MY) inducing grid: register-resident separable kernel.
//
// One thread owns one (particle, trajectory) pair: the projected inducing values
// W = A U_p (MX*MY float2) and the gradient accumulator gW live in registers for the
// whole solve, the ODE state never leaves registers, and the RBF kernel factorises
//   k_ab(x) = 2^-(c0 x0 - gxs_a)^2 * 2^-(c1 x1 - gys_b)^2
// so one RHS evaluation costs MX+MY SFU ex2 instead of MX*MY.
// Reference semantics: KernelRegression.forward gp.py:69-71 (f = K(x,Z) Kzz^-1 L U) and its
// reverse-mode derivative.
#pragma once
#include "common.cuh"

namespace bode {

struct NpdeKParams {
  int P, N, S, T, ppc, m;
  int y0_stride;  // 2N when y0 is per particle, 0 when shared
  int add_prior;
  float sign, scale;
  float c0, c1;   // sqrt(log2(e)/2) / ell_d
  float k0, k1;   // 2 ln2 c_d     (d kappa / dx = -k * delta * kappa)
  float gxs[16], gys[16];  // c0*gx[a], c1*gy[b]
  const float *U, *logsn, *A, *Ksym, *y0, *dt, *Y, *gout, *adj_dt, *Z;
  const int *obs_ptr, *adj_ptr;
  float2* ck;
  long long npairs;
  long long U_stride, logsn_stride, gU_stride, glogsn_stride;
  int stage_off;      // float offset in dynamic shared memory of the staged dt[S] | obs_ptr[S+1] | Y[N,T,2]
  float reg, lik_w;   // MLP closure: lik_w * sum (X - x)^2 + reg * sum theta^2
  float *sol, *loss, *sqerr, *gU, *glogsn, *gy0;
};

enum { INJ_LIK = 0, INJ_GOUT = 1 };

__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 operator*(float s, float2 a) { return f2(s * a.x, s * a.y); }
__device__ __forceinline__ float2 fma2(float s, float2 a, float2 b) { return f2(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y)); }

template <int MX, int MY>
struct SepField {
  static constexpr int G = 1;              // lanes cooperating on one (particle, trajectory) pair
  static constexpr int MAX_THREADS = 256;
  float2 W[MX][MY];
  float2 gW[MX][MY];

  static __device__ __forceinline__ void prologue(const NpdeKParams& prm, float* smem);
  static __device__ __forceinline__ float2 lik_weight(const NpdeKParams& prm, int p) {
    const float2 ls = *reinterpret_cast<const float2*>(prm.logsn + (long long)p * prm.logsn_stride);
    return f2(expf(-2.f * ls.x), expf(-2.f * ls.y));
  }
  template <int INJ>
  static __device__ __forceinline__ void epilogue(const NpdeKParams& prm, float* smem, const SepField& fld, bool active, int pl,
                                                  int n, int pairl, int lane, float r2x, float r2y);
  __device__ __forceinline__ void load(const NpdeKParams& prm, const float* smem, int pl, int, int lane) {
    load_W(prm, smem + (prm.ppc + pl) * 2 * prm.m, lane);      // Ws sits after Us in the CTA's shared memory
  }
  __device__ __forceinline__ void load_W(const NpdeKParams&, const float* Wp, int) {
#pragma unroll
    for (int a = 0; a < MX; ++a)
#pragma unroll
      for (int b = 0; b < MY; ++b) W[a][b] = f2(Wp[2 * (a * MY + b)], Wp[2 * (a * MY + b) + 1]);
  }
  __device__ __forceinline__ void store_gW(const NpdeKParams&, float* gp, int) const {
#pragma unroll
    for (int a = 0; a < MX; ++a)
#pragma unroll
      for (int b = 0; b < MY; ++b) {
        gp[2 * (a * MY + b)] = gW[a][b].x;
        gp[2 * (a * MY + b) + 1] = gW[a][b].y;
      }
  }

  __device__ __forceinline__ void zero_grad() {
#pragma unroll
    for (int a = 0; a < MX; ++a)
#pragma unroll
      for (int b = 0; b < MY; ++b) gW[a][b] = f2(0.f, 0.f);
  }

  // f(x) = sum_ab kx_a ky_b W_ab
  __device__ __forceinline__ float2 eval(const NpdeKParams& prm, float2 x) const {
    const float u0 = prm.c0 * x.x, u1 = prm.c1 * x.y;
    float ky[MY];
#pragma unroll
    for (int b = 0; b < MY; ++b) {
      const float d = u1 - prm.gys[b];
      ky[b] = ex2(-d * d);
    }
    float2 f = f2(0.f, 0.f);
#pragma unroll
    for (int a = 0; a < MX; ++a) {
      const float d = u0 - prm.gxs[a];
      const float kx = ex2(-d * d);
      float2 t = f2(0.f, 0.f);
#pragma unroll
      for (int b = 0; b < MY; ++b) t = fma2(ky[b], W[a][b], t);
      f = fma2(kx, t, f);
    }
    return f;
  }

  // Returns J(x)^T a and accumulates gW += wg * kappa(x) (x) a.   If WITH_F also returns f(x).
  template <bool WITH_F>
  __device__ __forceinline__ float2 vjp(const NpdeKParams& prm, float2 x, float2 a, float wg, float2* fout) {
    const float u0 = prm.c0 * x.x, u1 = prm.c1 * x.y;
    float ky[MY], dy[MY], ub[MY];
#pragma unroll
    for (int b = 0; b < MY; ++b) {
      dy[b] = u1 - prm.gys[b];
      ky[b] = ex2(-dy[b] * dy[b]);
      ub[b] = 0.f;
    }
    const float aw0 = a.x * wg, aw1 = a.y * wg;
    float sx = 0.f;
    float2 f = f2(0.f, 0.f);
#pragma unroll
    for (int ia = 0; ia < MX; ++ia) {
      const float dx = u0 - prm.gxs[ia];
      const float kx = ex2(-dx * dx);
      const float ka0 = kx * aw0, ka1 = kx * aw1;
      float s = 0.f;
      float2 t = f2(0.f, 0.f);
#pragma unroll
      for (int b = 0; b < MY; ++b) {
        const float2 w = W[ia][b];
        const float c = fmaf(a.y, w.y, a.x * w.x);
        s = fmaf(c, ky[b], s);
        ub[b] = fmaf(c, kx, ub[b]);
        gW[ia][b].x = fmaf(ka0, ky[b], gW[ia][b].x);
        gW[ia][b].y = fmaf(ka1, ky[b], gW[ia][b].y);
        if (WITH_F) t = fma2(ky[b], w, t);
      }
      sx = fmaf(kx * dx, s, sx);
      if (WITH_F) f = fma2(kx, t, f);
    }
    float sy = 0.f;
#pragma unroll
    for (int b = 0; b < MY; ++b) sy = fmaf(ky[b] * dy[b], ub[b], sy);
    if (WITH_F) *fout = f;
    return f2(-prm.k0 * sx, -prm.k1 * sy);
  }
};

}  // namespace bode
