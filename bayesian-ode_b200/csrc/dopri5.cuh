// Adaptive Dormand-Prince 5(4) forward solve, one independent step-size controller per (particle, trajectory) pair
// -- the "batched-step variant" of torchdiffeq's Dopri5Solver (dopri5.py:58-122): the reference notebook integrates one
// trajectory row per odeint call (nn.ipynb cell 10), so every pair reproduces the control flow of its own reference call:
//   tableau / DPS_C_MID            dopri5.py:11-36
//   initial step (order 4)         misc.py:84-143  (dopri5.py:79-82: ANY user first_step becomes 0.01)
//   stage loop, FSAL, error        rk_common.py:22-61
//   error ratio, accept            misc.py:146-157, dopri5.py:109     (mean over the elements of the state tensor)
//   step-size controller           misc.py:160-170 (float64 t, dt; applied after accepted AND rejected steps)
//   dense output                   interp.py:5-65  (outputs are interpolated; dt is never clipped to hit t[i])
// t and dt are float64 like the reference's controller; the state is fp32.
//
// Controller granularity (Dopri5Params::pool).  torchdiffeq keeps ONE controller per odeint call and pools the error over every
// element of y0 (misc.py:146-157: mean over the whole tensor; the initial-step norms of misc.py:116-143 likewise).  The notebook
// integrates one trajectory row per call (nn.ipynb cell 10), gp.py integrates all N rows in one call (gp.py:346, 452):
//   pool = 0  one controller per (particle, trajectory) pair   == one reference call per row
//   pool = 1  one controller per particle, error pooled over its N trajectories x 2 components  == one reference call with y0 [N, 2]
// In pooled mode the N pairs of a particle advance in lock-step (same dt, same accept / reject); their squared error terms meet
// in shared memory (CtrlPool) and every thread of the particle adds them in the same fixed order.
#pragma once
#include "npde_solve.cuh"

namespace bode {

struct Dopri5Params {
  const double* t;        // [T] (already sign-normalised: increasing)
  float rtol, atol;
  int user_first_step;    // dopri5.py:81-82
  double safety, ifactor, dfactor;
  int max_num_steps;
  int* stats;             // [P*N][3]: accepted, rejected, status bits (1 max_num_steps, 2 dt underflow, 4 non-finite state)
  int pool;               // 0: controller per pair, 1: per particle (see above)
  // Tuple states (misc.py:175-182): the N trajectories of a particle are the concatenation of ngroups state tensors, group g =
  // trajectories [gend[g-1], gend[g]).  torchdiffeq pools the error PER TENSOR and takes the maximum over the tuple (accept iff
  // every tensor's mean ratio <= 1, dopri5.py:108-109; _optimal_step_size uses max(ratios), misc.py:161; the initial-step norms
  // likewise, misc.py:125-141).  ngroups <= 1: one tensor.
  int ngroups;
  int gend[4];
};

__device__ __forceinline__ float rms2(float2 v) { return sqrtf(0.5f * (v.x * v.x + v.y * v.y)); }
__device__ __forceinline__ float ss2(float2 v) { return v.x * v.x + v.y * v.y; }
__device__ __forceinline__ float2 div2(float2 a, float2 b) { return f2(a.x / b.x, a.y / b.y); }

// Pooled controller: sum of one float per pair (and OR of one flag) over the N pairs of a particle.  Geometry comes from plan()
// (npde.cu) / the MLP launchers: G == 1 -> 32-thread CTAs holding floor(32 / N) particles, so a particle's lanes share a warp and
// synchronise with __syncwarp on their own lane mask (particles of one warp take different numbers of attempts); G == 32 -> one
// particle per CTA (N warps), __syncthreads.  Two alternating buffers: one barrier per reduction.
template <int G>
struct CtrlPool {
  float (*buf)[2][64];
  int flip, pairl, first, N, ng;
  int ge[4];
  unsigned mask;
  __device__ __forceinline__ void init(float (*b)[2][64], int pairl_, int pl, int N_, const Dopri5Params& dp) {
    buf = b; flip = 0; pairl = pairl_; N = N_; first = pl * N_;
    ng = dp.ngroups > 1 ? dp.ngroups : 1;
#pragma unroll
    for (int g = 0; g < 4; ++g) ge[g] = dp.ngroups > 1 ? dp.gend[g] : N_;
    mask = G == 1 ? (((N_ >= 32 ? 0u : (1u << N_)) - 1u) << (first & 31)) : 0xffffffffu;
  }
  // gm[g] = mean over the elements of state tensor g of v (v = this pair's sum over its 2 components); *any = OR of the flags
  __device__ __forceinline__ void means(float v, int lane, int flag, int* any, float (&gm)[4]) {
    float(*b)[64] = buf[flip];
    flip ^= 1;
    if (lane == 0) { b[0][pairl] = v; b[1][pairl] = flag ? 1.f : 0.f; }
    if (G == 1) __syncwarp(mask); else __syncthreads();
    float f = 0.f;
    int n = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float s = 0.f;
      const int n0 = n;
      if (g < ng)
        for (; n < ge[g]; ++n) { s += b[0][first + n]; f += b[1][first + n]; }
      gm[g] = (g < ng && n > n0) ? s / (2.f * (float)(n - n0)) : 0.f;
    }
    *any = f != 0.f;
  }
  __device__ __forceinline__ float max_mean(float v, int lane, int flag, int* any) {
    float gm[4];
    means(v, lane, flag, any, gm);
    return fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));      // unused groups are 0 and every mean is >= 0
  }
};

// misc.py:84-143 (_select_initial_step, order 4) in the state dtype; pooled mode forms the norms per state tensor and combines them
// as the reference does for tuples (max of the norms; h0 from the largest d0 / d1 ratio).
template <class Field, int G>
__device__ __forceinline__ double dopri5_first_dt(const NpdeKParams& prm, const Dopri5Params& dp, const Field& fld, CtrlPool<G>& cp, int lane,
                                                  float2 y, float2 f, float sg) {
  const float2 scale = f2(dp.atol + fabsf(y.x) * dp.rtol, dp.atol + fabsf(y.y) * dp.rtol);
  float d0, d1, r01;
  int any = 0;
  if (dp.pool) {
    float g0[4], g1[4];
    cp.means(ss2(div2(y, scale)), lane, 0, &any, g0);
    cp.means(ss2(div2(f, scale)), lane, 0, &any, g1);
    d0 = d1 = r01 = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < cp.ng) {
        const float a = sqrtf(g0[g]), b = sqrtf(g1[g]);
        d0 = fmaxf(d0, a);
        d1 = fmaxf(d1, b);
        r01 = fmaxf(r01, a / b);
      }
  } else {
    d0 = rms2(div2(y, scale));
    d1 = rms2(div2(f, scale));
    r01 = 0.f;
  }
  // one tensor: 0.01 * d0 / d1 (misc.py:128 evaluates left to right); a tuple: 0.01 * max_g(d0_g / d1_g)
  const float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : ((dp.pool && cp.ng > 1) ? 0.01f * r01 : 0.01f * d0 / d1);
  const float2 f1 = sg * fld.eval(prm, fma2(h0, f, y));
  float d2 = rms2(div2(f1 - f, scale)) / h0;
  if (dp.pool) d2 = sqrtf(cp.max_mean(ss2(div2(f1 - f, scale)), lane, 0, &any)) / h0;
  float h1;
  if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
  else h1 = powf(0.01f / fmaxf(d1, d2), 1.f / 5.f);
  return (double)fminf(100.f * h0, h1);
}

template <class Field>
__global__ void __launch_bounds__(Field::MAX_THREADS) dopri5_fwd_kernel(const __grid_constant__ NpdeKParams prm,
                                                                       const __grid_constant__ Dopri5Params dp) {
  extern __shared__ __align__(16) float smem[];
  constexpr int G = Field::G;
  Field::prologue(prm, smem);
  const int tid = threadIdx.x;
  const int pairl = tid / G, lane = tid % G;
  const int pl = pairl / prm.N, n = pairl % prm.N;
  const int p = blockIdx.x * prm.ppc + pl;
  __shared__ float pool_buf[2][2][64];
  if (pl >= prm.ppc || p >= prm.P) return;
  CtrlPool<G> cp;
  cp.init(pool_buf, pairl, pl, prm.N, dp);
  int any = 0;
  Field fld;
  fld.load(prm, smem, pl, pairl, lane);
  const long long pair = (long long)p * prm.N + n;
  const long long PN = (long long)prm.P * prm.N;
  float2* sol = reinterpret_cast<float2*>(prm.sol);
  const float sg = prm.sign;
  // tableau (dopri5.py:11-31)
  const float B21 = 1.f / 5, B31 = 3.f / 40, B32 = 9.f / 40, B41 = 44.f / 45, B42 = -56.f / 15, B43 = 32.f / 9;
  const float B51 = (float)(19372.0 / 6561), B52 = (float)(-25360.0 / 2187), B53 = (float)(64448.0 / 6561), B54 = (float)(-212.0 / 729);
  const float B61 = (float)(9017.0 / 3168), B62 = (float)(-355.0 / 33), B63 = (float)(46732.0 / 5247), B64 = (float)(49.0 / 176),
              B65 = (float)(-5103.0 / 18656);
  const float C1 = (float)(35.0 / 384), C3 = (float)(500.0 / 1113), C4 = (float)(125.0 / 192), C5 = (float)(-2187.0 / 6784), C6 = (float)(11.0 / 84);
  const float E1 = (float)(35.0 / 384 - 1951.0 / 21600), E3 = (float)(500.0 / 1113 - 22642.0 / 50085), E4 = (float)(125.0 / 192 - 451.0 / 720),
              E5 = (float)(-2187.0 / 6784 + 12231.0 / 42400), E6 = (float)(11.0 / 84 - 649.0 / 6300), E7 = (float)(-1.0 / 60);
  const float M1 = (float)(6025192743.0 / 30085553152.0 / 2), M3 = (float)(51252292925.0 / 65400821598.0 / 2),
              M4 = (float)(-2691868925.0 / 45128329728.0 / 2), M5 = (float)(187940372067.0 / 1594534317056.0 / 2),
              M6 = (float)(-1776094331.0 / 19743644256.0 / 2), M7 = (float)(11237099.0 / 235043384.0 / 2);

  float2 y = reinterpret_cast<const float2*>(prm.y0)[(prm.y0_stride ? (long long)p * prm.N : 0) + n];
  if (lane == 0) sol[pair] = y;
  int n_acc = 0, n_rej = 0, status = 0;
  if (prm.T > 1) {
    float2 f = sg * fld.eval(prm, y);
    // ---- first step (misc.py:84-143), state dtype arithmetic
    double dt;
    if (dp.user_first_step) {
      dt = 0.01;
    } else {
      dt = dopri5_first_dt<Field, G>(prm, dp, fld, cp, lane, y, f, sg);
    }
    double t0 = dp.t[0], t1 = dp.t[0];
    float2 ca = y, cb = y, cc = y, cd = y, ce = y;       // interp_coeff = [y0]*5 (dopri5.py:83)
    for (int i = 1; i < prm.T; ++i) {
      const double next_t = dp.t[i];
      int n_steps = 0;
      while (next_t > t1) {
        if (n_steps >= dp.max_num_steps) { status |= 1; break; }
        const double ts = t1;                            // start of the attempted step
        if (!(ts + dt > ts)) { status |= 2; break; }
        const bool nonfin = !(fabsf(y.x) <= 3.4028234e38f && fabsf(y.y) <= 3.4028234e38f);
        if (nonfin && !dp.pool) { status |= 4; break; }     // pooled: reported after the reduction below, by the whole particle
        const float h = (float)dt;
        const float2 k1 = f;
        const float2 k2 = sg * fld.eval(prm, fma2(h * B21, k1, y));
        const float2 k3 = sg * fld.eval(prm, fma2(h * B32, k2, fma2(h * B31, k1, y)));
        const float2 k4 = sg * fld.eval(prm, fma2(h * B43, k3, fma2(h * B42, k2, fma2(h * B41, k1, y))));
        const float2 k5 = sg * fld.eval(prm, fma2(h * B54, k4, fma2(h * B53, k3, fma2(h * B52, k2, fma2(h * B51, k1, y)))));
        const float2 k6 = sg * fld.eval(prm, fma2(h * B65, k5, fma2(h * B64, k4, fma2(h * B63, k3, fma2(h * B62, k2, fma2(h * B61, k1, y))))));
        const float2 y1 = fma2(h * C6, k6, fma2(h * C5, k5, fma2(h * C4, k4, fma2(h * C3, k3, fma2(h * C1, k1, y)))));
        const float2 k7 = sg * fld.eval(prm, y1);        // FSAL: f1 = k[-1]
        const float2 err = fma2(h * E7, k7, fma2(h * E6, k6, fma2(h * E5, k5, fma2(h * E4, k4, fma2(h * E3, k3, (h * E1) * k1)))));
        const float2 tol = f2(dp.atol + dp.rtol * fmaxf(fabsf(y.x), fabsf(y1.x)), dp.atol + dp.rtol * fmaxf(fabsf(y.y), fabsf(y1.y)));
        const float2 er = div2(err, tol);
        float ratio = 0.5f * (er.x * er.x + er.y * er.y);
        if (dp.pool) {                                     // misc.py:151-157: mean over ALL elements of the state tensor
          ratio = cp.max_mean(er.x * er.x + er.y * er.y, lane, nonfin, &any);
          if (any) { status |= 4; break; }
        }
        const bool accept = ratio <= 1.f;
        if (accept) {
          const float2 ymid = fma2(h * M7, k7, fma2(h * M6, k6, fma2(h * M5, k5, fma2(h * M4, k4, fma2(h * M3, k3, fma2(h * M1, k1, y))))));
          // interp.py:21-35
          ca = fma2(16.f, ymid, fma2(-8.f, y1, fma2(-8.f, y, fma2(2.f * h, k7, (-2.f * h) * k1))));
          cb = fma2(-32.f, ymid, fma2(14.f, y1, fma2(18.f, y, fma2(-3.f * h, k7, (5.f * h) * k1))));
          cc = fma2(16.f, ymid, fma2(-5.f, y1, fma2(-11.f, y, fma2(h, k7, (-4.f * h) * k1))));
          cd = h * k1;
          ce = y;
          y = y1;
          f = k7;
          t0 = ts;
          t1 = ts + dt;
          ++n_acc;
        } else {
          t0 = ts;                                         // rk_state.t0 := start of the rejected attempt (dopri5.py:121)
          ++n_rej;
        }
        // misc.py:160-170
        double factor;
        if (ratio == 0.f) {
          dt = dt * dp.ifactor;
        } else {
          const double dfac = ratio < 1.f ? 1.0 : dp.dfactor;
          const double er5 = pow((double)sqrtf(ratio), 0.2);     // sqrt in the ratio's dtype, then ** (1/5) in float64
          factor = fmax(1.0 / dp.ifactor, fmin(er5 / dp.safety, 1.0 / dfac));
          dt = dt / factor;
        }
        ++n_steps;
      }
      if (status) break;
      // interp.py:38-65 in the state dtype
      const float ft0 = (float)t0, ft1 = (float)t1, ft = (float)next_t;
      const float x = (ft - ft0) / (ft1 - ft0);
      const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
      const float2 out = ((((x4 * ca) + (x3 * cb)) + (x2 * cc)) + (x * cd)) + ce;
      if (lane == 0) sol[(long long)i * PN + pair] = out;
    }
  }
  if (lane == 0 && dp.stats) {
    dp.stats[pair * 3 + 0] = n_acc;
    dp.stats[pair * 3 + 1] = n_rej;
    dp.stats[pair * 3 + 2] = status;
  }
}

}  // namespace bode

namespace bode {

// ------------------------------------------------------------------ dopri5: fused adaptive solve + closure + gradient
// Gradient definition (SURVEY.md hard part 6): the exact reverse of the ACCEPTED-step recursion and of the dense-output
// evaluation with the accepted step sizes frozen -- validated against the reference's odeint_adjoint(dopri5)
// (oracle/dopri5.py::solve_and_grad, 5e-6 relative at tight tolerance; autograd through the reference's controller is
// numerically meaningless, 1e3 x larger).  Forward records, per accepted step, its size and the six stage points; per
// output the interpolated value, the step it belongs to and x = (t - t0)/(t1 - t0).  FSAL couples the steps: k7 of
// step k is k1 of step k+1, so its cotangent is carried backwards.
struct Dopri5Rec {
  float2* steps;     // [max_rec][7][npairs]: y2..y6, y1, (h, 0)
  float4* outs;      // [T][npairs]: (out.x, out.y, x, step index as float)
  int max_rec;
};

template <class Field, int INJ>
__global__ void __launch_bounds__(Field::MAX_THREADS) dopri5_grad_kernel(const __grid_constant__ NpdeKParams prm,
                                                                        const __grid_constant__ Dopri5Params dp,
                                                                        const __grid_constant__ Dopri5Rec rec) {
  extern __shared__ __align__(16) float smem[];
  constexpr int G = Field::G;
  Field::prologue(prm, smem);
  const int tid = threadIdx.x;
  const int pairl = tid / G, lane = tid % G;
  const int N = prm.N;
  const int pl = pairl / N, n = pairl % N;
  const int p = blockIdx.x * prm.ppc + pl;
  const bool active = pl < prm.ppc && p < prm.P;
  __shared__ float pool_buf[2][2][64];
  CtrlPool<G> cp;
  cp.init(pool_buf, active ? pairl : 0, active ? pl : 0, N, dp);
  int any = 0;
  float r2x = 0.f, r2y = 0.f;
  Field fld;
  fld.zero_grad();
  const float sg = prm.sign;
  const float B21 = 1.f / 5, B31 = 3.f / 40, B32 = 9.f / 40, B41 = 44.f / 45, B42 = -56.f / 15, B43 = 32.f / 9;
  const float B51 = (float)(19372.0 / 6561), B52 = (float)(-25360.0 / 2187), B53 = (float)(64448.0 / 6561), B54 = (float)(-212.0 / 729);
  const float B61 = (float)(9017.0 / 3168), B62 = (float)(-355.0 / 33), B63 = (float)(46732.0 / 5247), B64 = (float)(49.0 / 176),
              B65 = (float)(-5103.0 / 18656);
  const float C1 = (float)(35.0 / 384), C3 = (float)(500.0 / 1113), C4 = (float)(125.0 / 192), C5 = (float)(-2187.0 / 6784), C6 = (float)(11.0 / 84);
  const float E1 = (float)(35.0 / 384 - 1951.0 / 21600), E3 = (float)(500.0 / 1113 - 22642.0 / 50085), E4 = (float)(125.0 / 192 - 451.0 / 720),
              E5 = (float)(-2187.0 / 6784 + 12231.0 / 42400), E6 = (float)(11.0 / 84 - 649.0 / 6300), E7 = (float)(-1.0 / 60);
  const float M1 = (float)(6025192743.0 / 30085553152.0 / 2), M3 = (float)(51252292925.0 / 65400821598.0 / 2),
              M4 = (float)(-2691868925.0 / 45128329728.0 / 2), M5 = (float)(187940372067.0 / 1594534317056.0 / 2),
              M6 = (float)(-1776094331.0 / 19743644256.0 / 2), M7 = (float)(11237099.0 / 235043384.0 / 2);
  int n_acc = 0, n_rej = 0, status = 0;
  if (active) {
    fld.load(prm, smem, pl, pairl, lane);
    const long long pair = (long long)p * N + n;
    const long long PN = (long long)prm.P * N;
    const float2* Y2 = reinterpret_cast<const float2*>(prm.Y) + (long long)n * prm.T;
    const float2* go = reinterpret_cast<const float2*>(prm.gout) + pair;
    float2 e2inv = f2(0.f, 0.f);
    if (INJ == INJ_LIK) e2inv = Field::lik_weight(prm, p);
    float2* rs = rec.steps + pair;
    float4* ro = rec.outs + pair;
    const long long stride = prm.npairs;
    const float2 yinit = reinterpret_cast<const float2*>(prm.y0)[(prm.y0_stride ? (long long)p * N : 0) + n];
    float2 y = yinit;
    if (INJ == INJ_LIK) {
      const float2 r = __ldg(Y2) - y;
      r2x = r.x * r.x;
      r2y = r.y * r.y;
    }
    // ---------------- forward (same control flow as dopri5_fwd_kernel) with recording
    if (prm.T > 1) {
      float2 f = sg * fld.eval(prm, y);
      double dt;
      if (dp.user_first_step) {
        dt = 0.01;
      } else {
        dt = dopri5_first_dt<Field, G>(prm, dp, fld, cp, lane, y, f, sg);
      }
      double t0 = dp.t[0], t1 = dp.t[0];
      float2 ca = y, cb = y, cc = y, cd = y, ce = y;
      for (int i = 1; i < prm.T; ++i) {
        const double next_t = dp.t[i];
        int n_steps = 0;
        while (next_t > t1) {
          if (n_steps >= dp.max_num_steps) { status |= 1; break; }
          const double ts = t1;
          if (!(ts + dt > ts)) { status |= 2; break; }
          const bool nonfin = !(fabsf(y.x) <= 3.4028234e38f && fabsf(y.y) <= 3.4028234e38f);
          if (nonfin && !dp.pool) { status |= 4; break; }
          const float h = (float)dt;
          const float2 k1 = f;
          const float2 p2 = fma2(h * B21, k1, y);
          const float2 k2 = sg * fld.eval(prm, p2);
          const float2 p3 = fma2(h * B32, k2, fma2(h * B31, k1, y));
          const float2 k3 = sg * fld.eval(prm, p3);
          const float2 p4 = fma2(h * B43, k3, fma2(h * B42, k2, fma2(h * B41, k1, y)));
          const float2 k4 = sg * fld.eval(prm, p4);
          const float2 p5 = fma2(h * B54, k4, fma2(h * B53, k3, fma2(h * B52, k2, fma2(h * B51, k1, y))));
          const float2 k5 = sg * fld.eval(prm, p5);
          const float2 p6 = fma2(h * B65, k5, fma2(h * B64, k4, fma2(h * B63, k3, fma2(h * B62, k2, fma2(h * B61, k1, y)))));
          const float2 k6 = sg * fld.eval(prm, p6);
          const float2 y1 = fma2(h * C6, k6, fma2(h * C5, k5, fma2(h * C4, k4, fma2(h * C3, k3, fma2(h * C1, k1, y)))));
          const float2 k7 = sg * fld.eval(prm, y1);
          const float2 err = fma2(h * E7, k7, fma2(h * E6, k6, fma2(h * E5, k5, fma2(h * E4, k4, fma2(h * E3, k3, (h * E1) * k1)))));
          const float2 tol = f2(dp.atol + dp.rtol * fmaxf(fabsf(y.x), fabsf(y1.x)), dp.atol + dp.rtol * fmaxf(fabsf(y.y), fabsf(y1.y)));
          const float2 er = div2(err, tol);
          float ratio = 0.5f * (er.x * er.x + er.y * er.y);
          if (dp.pool) {
            ratio = cp.max_mean(er.x * er.x + er.y * er.y, lane, nonfin, &any);
            if (any) { status |= 4; break; }
          }
          if (ratio <= 1.f) {
            const float2 ymid = fma2(h * M7, k7, fma2(h * M6, k6, fma2(h * M5, k5, fma2(h * M4, k4, fma2(h * M3, k3, fma2(h * M1, k1, y))))));
            ca = fma2(16.f, ymid, fma2(-8.f, y1, fma2(-8.f, y, fma2(2.f * h, k7, (-2.f * h) * k1))));
            cb = fma2(-32.f, ymid, fma2(14.f, y1, fma2(18.f, y, fma2(-3.f * h, k7, (5.f * h) * k1))));
            cc = fma2(16.f, ymid, fma2(-5.f, y1, fma2(-11.f, y, fma2(h, k7, (-4.f * h) * k1))));
            cd = h * k1;
            ce = y;
            if (n_acc < rec.max_rec) {
              if (lane == 0) {
                float2* q = rs + (long long)n_acc * 7 * stride;
                q[0] = p2; q[stride] = p3; q[2 * stride] = p4; q[3 * stride] = p5; q[4 * stride] = p6; q[5 * stride] = y1;
                q[6 * stride] = f2(h, 0.f);
              }
            } else {
              status |= 8;                                   // more accepted steps than the record buffer holds
            }
            y = y1; f = k7; t0 = ts; t1 = ts + dt;
            ++n_acc;
          } else {
            t0 = ts;
            ++n_rej;
          }
          if (ratio == 0.f) {
            dt = dt * dp.ifactor;
          } else {
            const double dfac = ratio < 1.f ? 1.0 : dp.dfactor;
            dt = dt / fmax(1.0 / dp.ifactor, fmin(pow((double)sqrtf(ratio), 0.2) / dp.safety, 1.0 / dfac));
          }
          ++n_steps;
        }
        if (status) break;
        const float ft0 = (float)t0, ft1 = (float)t1, ft = (float)next_t;
        const float x = (ft - ft0) / (ft1 - ft0);
        const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
        const float2 out = ((((x4 * ca) + (x3 * cb)) + (x2 * cc)) + (x * cd)) + ce;
        if (lane == 0) ro[(long long)i * stride] = make_float4(out.x, out.y, x, (float)(n_acc - 1));
        if (INJ == INJ_LIK) {
          const float2 r = __ldg(Y2 + i) - out;
          r2x = fmaf(r.x, r.x, r2x);
          r2y = fmaf(r.y, r.y, r2y);
        }
      }
    }
    // ---------------- backward over the accepted steps
    float2 abar = f2(0.f, 0.f), k7bar = f2(0.f, 0.f);
    if (!status) {
      if (G > 1) __syncwarp();
      int io = prm.T - 1;
      for (int k = n_acc - 1; k >= 0; --k) {
        const float2* q = rs + (long long)k * 7 * stride;
        const float2 p2 = q[0], p3 = q[stride], p4 = q[2 * stride], p5 = q[3 * stride], p6 = q[4 * stride], y1 = q[5 * stride];
        const float h = q[6 * stride].x;
        float2 kb1 = f2(0.f, 0.f), kb2 = kb1, kb3 = kb1, kb4 = kb1, kb5 = kb1, kb6 = kb1, kb7 = k7bar;
        float2 y0bar = f2(0.f, 0.f), y1bar = abar;
        while (io >= 1) {
          const float4 o = ro[(long long)io * stride];
          if ((int)o.w != k) break;
          float2 gi;
          if (INJ == INJ_LIK) {
            const float2 r = __ldg(Y2 + io) - f2(o.x, o.y);
            gi = f2(-r.x * e2inv.x, -r.y * e2inv.y);
          } else {
            gi = __ldg(go + (long long)io * PN);
          }
          const float x = o.z, x2 = x * x, x3 = x2 * x, x4 = x3 * x;
          const float w1 = -2.f * x4 + 5.f * x3 - 4.f * x2 + x, w7 = 2.f * x4 - 3.f * x3 + x2, wy0 = -8.f * x4 + 18.f * x3 - 11.f * x2 + 1.f,
                      wy1 = -8.f * x4 + 14.f * x3 - 5.f * x2, wm = 16.f * x4 - 32.f * x3 + 16.f * x2;
          const float hm = h * wm;
          kb1 = fma2(w1 * h + hm * M1, gi, kb1);
          kb3 = fma2(hm * M3, gi, kb3);
          kb4 = fma2(hm * M4, gi, kb4);
          kb5 = fma2(hm * M5, gi, kb5);
          kb6 = fma2(hm * M6, gi, kb6);
          kb7 = fma2(w7 * h + hm * M7, gi, kb7);
          y0bar = fma2(wy0 + wm, gi, y0bar);
          y1bar = fma2(wy1, gi, y1bar);
          --io;
        }
        float2 v = fld.template vjp<false>(prm, y1, sg * kb7, 1.f, nullptr);         // k7 = f(y1)
        y1bar = y1bar + v;
        y0bar = y0bar + y1bar;
        kb1 = fma2(h * C1, y1bar, kb1); kb3 = fma2(h * C3, y1bar, kb3); kb4 = fma2(h * C4, y1bar, kb4);
        kb5 = fma2(h * C5, y1bar, kb5); kb6 = fma2(h * C6, y1bar, kb6);
        v = fld.template vjp<false>(prm, p6, sg * kb6, 1.f, nullptr);
        y0bar = y0bar + v;
        kb1 = fma2(h * B61, v, kb1); kb2 = fma2(h * B62, v, kb2); kb3 = fma2(h * B63, v, kb3); kb4 = fma2(h * B64, v, kb4); kb5 = fma2(h * B65, v, kb5);
        v = fld.template vjp<false>(prm, p5, sg * kb5, 1.f, nullptr);
        y0bar = y0bar + v;
        kb1 = fma2(h * B51, v, kb1); kb2 = fma2(h * B52, v, kb2); kb3 = fma2(h * B53, v, kb3); kb4 = fma2(h * B54, v, kb4);
        v = fld.template vjp<false>(prm, p4, sg * kb4, 1.f, nullptr);
        y0bar = y0bar + v;
        kb1 = fma2(h * B41, v, kb1); kb2 = fma2(h * B42, v, kb2); kb3 = fma2(h * B43, v, kb3);
        v = fld.template vjp<false>(prm, p3, sg * kb3, 1.f, nullptr);
        y0bar = y0bar + v;
        kb1 = fma2(h * B31, v, kb1); kb2 = fma2(h * B32, v, kb2);
        v = fld.template vjp<false>(prm, p2, sg * kb2, 1.f, nullptr);
        y0bar = y0bar + v;
        kb1 = fma2(h * B21, v, kb1);
        abar = y0bar;
        k7bar = kb1;
        if (k == 0) {
          v = fld.template vjp<false>(prm, yinit, sg * kb1, 1.f, nullptr);           // k1 of the first step = f(y0)
          abar = abar + v;
        }
      }
      // output 0 is y0 itself
      if (INJ == INJ_LIK) {
        const float2 r = __ldg(Y2) - yinit;
        abar = f2(fmaf(-r.x, e2inv.x, abar.x), fmaf(-r.y, e2inv.y, abar.y));
      } else {
        abar = abar + __ldg(go);
      }
    }
    if (prm.gy0 != nullptr && lane == 0) reinterpret_cast<float2*>(prm.gy0)[pair] = prm.scale * abar;
    if (lane == 0 && dp.stats) {
      dp.stats[pair * 3 + 0] = n_acc;
      dp.stats[pair * 3 + 1] = n_rej;
      dp.stats[pair * 3 + 2] = status;
    }
  }
  Field::template epilogue<INJ>(prm, smem, fld, active, pl, n, pairl, lane, r2x, r2y);
}

}  // namespace bode
