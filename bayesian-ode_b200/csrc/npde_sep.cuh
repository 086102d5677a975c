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
  int gmx, gmy;            // grid size for the row-sliced field (npde_row.cuh)
  const float *U, *logsn, *A, *Ksym, *y0, *dt, *Y, *gout, *adj_dt, *Z;
  const float* AT;    // optional transpose of A (coalesced projection), may be null
  const float* Wpre;  // split mode (npde_proj.cu): W = A U precomputed for all particles [P][2m]; the epilogue then leaves sum_n gW in gU
  int split;
  const int *obs_ptr, *adj_ptr;
  float2* ck;
  long long npairs;
  long long U_stride, logsn_stride, gU_stride, glogsn_stride;
  int stage_off;      // float offset in dynamic shared memory of the staged dt[S] | obs_ptr[S+1] | Y[N,T,2]
  int a_off;          // pair kernels: float offset of the staged A[m,m] | Ksym[m,m]
  float reg, lik_w;   // MLP closure: lik_w * sum (X - x)^2 + reg * sum theta^2
  float *sol, *loss, *sqerr, *gU, *glogsn, *gy0;
  float *VH, *VC;     // optional: SVGD operand tiles that receive vsign * (gU | glogsn) of every particle (svgd_tiles.cuh)
  float vsign;
};

enum { INJ_LIK = 0, INJ_GOUT = 1 };

__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 operator*(float s, float2 a) { return f2(s * a.x, s * a.y); }
__device__ __forceinline__ float2 fma2(float s, float2 a, float2 b) { return f2(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y)); }

// Packed FP32 pairs: sm_100a's FFMA2 (PTX fma.rn.f32x2) retires two FMAs per issue slot; ptxas folds a (x, x) pair
// into a scalar-broadcast operand, so a broadcast costs no extra instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2x(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float hsum(f32x2 v) {
  float lo, hi;
  upk(v, lo, hi);
  return lo + hi;
}

template <int MX, int MY>
struct SepField {
  static constexpr int G = 1;              // lanes cooperating on one (particle, trajectory) pair
  static constexpr int MAX_THREADS = 256;
  f32x2 W[MX][MY];      // (W_ab0, W_ab1): both output components of one inducing point in one FFMA2 operand
  f32x2 gW[MX][MY];

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
      for (int b = 0; b < MY; ++b) W[a][b] = pk(Wp[2 * (a * MY + b)], Wp[2 * (a * MY + b) + 1]);
  }
  __device__ __forceinline__ void store_gW(const NpdeKParams&, float* gp, int) const {
#pragma unroll
    for (int a = 0; a < MX; ++a)
#pragma unroll
      for (int b = 0; b < MY; ++b) upk(gW[a][b], gp[2 * (a * MY + b)], gp[2 * (a * MY + b) + 1]);
  }

  __device__ __forceinline__ void zero_grad() {
#pragma unroll
    for (int a = 0; a < MX; ++a)
#pragma unroll
      for (int b = 0; b < MY; ++b) gW[a][b] = pk(0.f, 0.f);
  }

  // f(x) = sum_ab kx_a ky_b W_ab
  __device__ __forceinline__ float2 eval(const NpdeKParams& prm, float2 x) const {
    float ky[MY];
#pragma unroll
    for (int b = 0; b < MY; ++b) {
      const float d = fmaf(prm.c1, x.y, -prm.gys[b]);
      ky[b] = ex2(-d * d);
    }
    f32x2 f = pk(0.f, 0.f);
#pragma unroll
    for (int a = 0; a < MX; ++a) {
      const float d = fmaf(prm.c0, x.x, -prm.gxs[a]);
      const float kx = ex2(-d * d);
      f32x2 t = pk(0.f, 0.f);
#pragma unroll
      for (int b = 0; b < MY; ++b) t = fma2x(W[a][b], pk(ky[b], ky[b]), t);
      f = fma2x(t, pk(kx, kx), f);
    }
    float2 r;
    upk(f, r.x, r.y);
    return r;
  }

  // Returns J(x)^T a and accumulates gW += wg * kappa(x) (x) a.   If WITH_F also returns f(x).
  // The two Jacobian columns  df/dx0 = -k0 sum_a (kx_a dx_a) T_a,  df/dx1 = -k1 sum_a kx_a T'_a  with
  // T_a = sum_b ky_b W_ab,  T'_a = sum_b (ky_b dy_b) W_ab  do not depend on the adjoint: only four FMAs sit on the serial chain.
  template <bool WITH_F>
  __device__ __forceinline__ float2 vjp(const NpdeKParams& prm, float2 x, float2 a, float wg, float2* fout) {
    float ky[MY], kdy[MY];
#pragma unroll
    for (int b = 0; b < MY; ++b) {
      const float d = fmaf(prm.c1, x.y, -prm.gys[b]);
      ky[b] = ex2(-d * d);
      kdy[b] = ky[b] * d;
    }
    const f32x2 aw = pk(a.x * wg, a.y * wg);
    f32x2 jx = pk(0.f, 0.f), jy = pk(0.f, 0.f), f = pk(0.f, 0.f);
#pragma unroll
    for (int ia = 0; ia < MX; ++ia) {
      const float dx = fmaf(prm.c0, x.x, -prm.gxs[ia]);
      const float kx = ex2(-dx * dx);
      const float kdx = kx * dx;
      const f32x2 kaw = fma2x(aw, pk(kx, kx), pk(0.f, 0.f));
      f32x2 t = pk(0.f, 0.f), tp = pk(0.f, 0.f);
#pragma unroll
      for (int b = 0; b < MY; ++b) {
        t = fma2x(W[ia][b], pk(ky[b], ky[b]), t);
        tp = fma2x(W[ia][b], pk(kdy[b], kdy[b]), tp);
        gW[ia][b] = fma2x(kaw, pk(ky[b], ky[b]), gW[ia][b]);
      }
      jx = fma2x(t, pk(kdx, kdx), jx);
      jy = fma2x(tp, pk(kx, kx), jy);
      if (WITH_F) f = fma2x(t, pk(kx, kx), f);
    }
    if (WITH_F) upk(f, fout->x, fout->y);
    float jx0, jx1, jy0, jy1;
    upk(jx, jx0, jx1);
    upk(jy, jy0, jy1);
    return f2(-prm.k0 * fmaf(a.y, jx1, a.x * jx0), -prm.k1 * fmaf(a.y, jy1, a.x * jy0));
  }
};

}  // namespace bode
