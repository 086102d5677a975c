// Component-split separable npde kernels: TWO lanes per (particle, trajectory) pair.
//
// The tensor-grid kernel factorises, k_ab(x) = kx_a(x0) ky_b(x1), and the 2-D state splits the same way:
// lane d of a pair owns state component y_d, adjoint component a_d, the d-th column of W = A U_p and of the
// gradient accumulator (M*M floats each, stored [own axis][partner axis]), and evaluates only the M kernel
// factors of ITS axis; the partner's factors arrive with M warp shuffles.  Compared with one thread per pair
// (npde_sep.cuh) this doubles the number of resident warps at a fixed particle count (the 4096-particle
// configuration only has 20 480 pairs for 148 SMs), halves the per-thread register footprint, and keeps the
// SFU work identical (2M ex2 per RHS evaluation per pair).
//
// The reverse sweep recomputes the Jacobian row of the lane's own output component,
//   df_d/dx_d   = -k_d  sum_i (k_i delta_i) T_i ,   T_i  = sum_j W_ij k'_j
//   df_d/dx_d'  = -k_d' sum_i  k_i          T'_i ,  T'_i = sum_j W_ij (k'_j delta'_j)
// (k, delta: own-axis factors / scaled offsets, primes: the partner axis), none of which depends on the incoming
// adjoint, so only two multiplies per stage sit on the serial adjoint chain; J^T a is completed with one shuffle.
// Reference semantics are those of npde_sep.cuh / npde_solve.cuh: gp.py:69-71 (field), solvers.py:79-99 +
// rk_common.py:72-78 (fixed-grid RK), gp.py:342-353 (closure), adjoint.py:23-102 (continuous adjoint).
#pragma once
#include "npde_solve.cuh"
#include "svgd_tiles.cuh"

namespace bode {

constexpr unsigned FULL_MASK = 0xffffffffu;

template <int M>
struct PairField {
  static constexpr int G = 2;
  static constexpr int MAX_THREADS = M <= 5 ? 384 : 320;   // 12 warps (three per scheduler) where the registers allow it
  static constexpr int MP = (M + 1) / 2;
  // i indexes the lane's own axis, j the partner's.  W is held as pairs over j (zero padded when M is odd), the
  // gradient accumulator as pairs over i, so that every contraction below is an FFMA2 with one broadcast operand.
  f32x2 Wp[M][MP];    // (W[i][2jp], W[i][2jp+1])
  f32x2 gWp[MP][M];   // (gW[2ip][j], gW[2ip+1][j])
  float gm[M], gt[M];   // scaled grid coordinates c_d * g_d[i] of the own / partner axis
  float cm, ct;         // coordinate scales
  float nkm, nkt;       // -k_own, -k_partner  (d kappa / dx = -k delta kappa)

  __device__ __forceinline__ void load(const NpdeKParams& prm, const float* Wp_, int d) {
    auto w = [&](int i, int j) { return j < M ? (d ? Wp_[2 * (j * M + i) + 1] : Wp_[2 * (i * M + j)]) : 0.f; };
#pragma unroll
    for (int i = 0; i < M; ++i) {
      gm[i] = d ? prm.gys[i] : prm.gxs[i];
      gt[i] = d ? prm.gxs[i] : prm.gys[i];
#pragma unroll
      for (int jp = 0; jp < MP; ++jp) Wp[i][jp] = pk(w(i, 2 * jp), w(i, 2 * jp + 1));
    }
#pragma unroll
    for (int ip = 0; ip < MP; ++ip)
#pragma unroll
      for (int j = 0; j < M; ++j) gWp[ip][j] = pk(0.f, 0.f);
    cm = d ? prm.c1 : prm.c0;
    ct = d ? prm.c0 : prm.c1;
    nkm = d ? -prm.k1 : -prm.k0;
    nkt = d ? -prm.k0 : -prm.k1;
  }
  // npde_epilogue's contract: write this lane's share of gW[m][2]
  __device__ __forceinline__ void store_gW(const NpdeKParams&, float* gp, int d) const {
#pragma unroll
    for (int ip = 0; ip < MP; ++ip)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        float g[2];
        upk(gWp[ip][j], g[0], g[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * ip + h;
          if (i < M) {
            if (d) gp[2 * (j * M + i) + 1] = g[h];
            else gp[2 * (i * M + j)] = g[h];
          }
        }
      }
  }

  // f_d(x), x_d = ym (the partner lane holds x_d').  The contraction over the lane's OWN axis, Q_j = sum_i k_i W_ij, runs while the
  // partner's factors are still in flight through the shuffle unit; only sum_j k'_j Q_j waits for them.
  __device__ __forceinline__ float eval(float ym) const {
    float km[M], kt[2 * MP];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const float dl = fmaf(cm, ym, -gm[i]);
      km[i] = ex2(-dl * dl);
    }
#pragma unroll
    for (int j = 0; j < M; ++j) kt[j] = __shfl_xor_sync(FULL_MASK, km[j], 1);
    if (M & 1) kt[M] = 0.f;
    f32x2 q[MP];
#pragma unroll
    for (int jp = 0; jp < MP; ++jp) q[jp] = pk(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const f32x2 kb = pk(km[i], km[i]);
#pragma unroll
      for (int jp = 0; jp < MP; ++jp) q[jp] = fma2x(Wp[i][jp], kb, q[jp]);
    }
    f32x2 facc = pk(0.f, 0.f);
#pragma unroll
    for (int jp = 0; jp < MP; ++jp) facc = fma2x(q[jp], pk(kt[2 * jp], kt[2 * jp + 1]), facc);
    return hsum(facc);
  }

  // Component d of J(x)^T a, given this lane's a_d; accumulates gW += wg a_d kappa(x).  WITH_F: also f_d(x).
  template <bool WITH_F>
  __device__ __forceinline__ float vjp(float ym, float yt, float a, float wg, float* fout) {
    float km[2 * MP], kdm[M], kt[M], kdt[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const float dl = fmaf(cm, ym, -gm[i]);
      km[i] = ex2(-dl * dl);
      kdm[i] = km[i] * dl;
    }
    if (M & 1) km[M] = 0.f;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      kt[j] = __shfl_xor_sync(FULL_MASK, km[j], 1);
      kdt[j] = kt[j] * fmaf(ct, yt, -gt[j]);
    }
    const float aw = a * wg;
    // (R_j, Q_j) = sum_i (k_i delta_i, k_i) W_ij: own-axis quantities only, issued under the shuffle latency
    f32x2 qr[M];
#pragma unroll
    for (int j = 0; j < M; ++j) qr[j] = pk(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const f32x2 kk = pk(kdm[i], km[i]);
#pragma unroll
      for (int j = 0; j < M; ++j) {
        float wlo, whi;
        upk(Wp[i][j >> 1], wlo, whi);
        const float w = (j & 1) ? whi : wlo;
        qr[j] = fma2x(kk, pk(w, w), qr[j]);
      }
    }
    f32x2 smst = pk(0.f, 0.f);           // (sm, st) = (sum_ij kdm_i W_ij kt_j, sum_ij km_i W_ij kdt_j)
    float f = 0.f;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      smst = fma2x(qr[j], pk(kt[j], kdt[j]), smst);
      if (WITH_F) {
        float rj, qj;
        upk(qr[j], rj, qj);
        f = fmaf(qj, kt[j], f);
      }
    }
    float sm, st;
    upk(smst, sm, st);
#pragma unroll
    for (int ip = 0; ip < MP; ++ip) {
      const f32x2 kaw = pk(km[2 * ip] * aw, km[2 * ip + 1] * aw);
#pragma unroll
      for (int j = 0; j < M; ++j) gWp[ip][j] = fma2x(kaw, pk(kt[j], kt[j]), gWp[ip][j]);
    }
    if (WITH_F) *fout = f;
    const float recv = __shfl_xor_sync(FULL_MASK, (nkt * a) * st, 1);
    return fmaf(nkm * a, sm, recv);
  }
};

// ------------------------------------------------------------------ one forward step on the lane's component
template <int METHOD, bool STORE, int M>
__device__ __forceinline__ float pair_step_fwd(const NpdeKParams& prm, const PairField<M>& fld, float y, float dt, float* ckp,
                                               long long stride, bool act) {
  // `act` only predicates the checkpoint stores: every lane of the warp must reach the shuffles in eval() together
  const float sg = prm.sign;
  if (METHOD == BODE_EULER) {
    if (STORE && act) ckp[0] = y;
    return fmaf(dt, sg * fld.eval(y), y);
  } else if (METHOD == BODE_MIDPOINT) {
    if (STORE && act) ckp[0] = y;
    const float k1 = sg * fld.eval(y);
    const float y2 = fmaf(0.5f * dt, k1, y);
    if (STORE && act) ckp[stride] = y2;
    return fmaf(dt, sg * fld.eval(y2), y);
  } else {  // 3/8 rule, rk_common.py:72-78
    if (STORE && act) ckp[0] = y;
    const float k1 = sg * fld.eval(y);
    const float dt3 = dt * (1.f / 3.f);
    const float y2 = fmaf(dt3, k1, y);
    if (STORE && act) ckp[stride] = y2;
    const float k2 = sg * fld.eval(y2);
    const float y3 = fmaf(dt, k2, fmaf(-dt3, k1, y));
    if (STORE && act) ckp[2 * stride] = y3;
    const float k3 = sg * fld.eval(y3);
    const float y4 = fmaf(dt, (k1 - k2) + k3, y);
    if (STORE && act) ckp[3 * stride] = y4;
    const float k4 = sg * fld.eval(y4);
    return fmaf(dt * 0.125f, (k1 + k4) + 3.f * (k2 + k3), y);
  }
}

// stage points of one step as (own, partner) components
template <int METHOD>
__device__ __forceinline__ void pair_load_stage_points(float2 (&ys)[Stages<METHOD>::value], const float2* ckp, long long stride, int d) {
#pragma unroll
  for (int i = 0; i < Stages<METHOD>::value; ++i) {
    const float2 v = ckp[(long long)i * stride];
    ys[i] = d ? f2(v.y, v.x) : v;
  }
}

// ------------------------------------------------------------------ discrete adjoint of one step
template <int METHOD, int M>
__device__ __forceinline__ float pair_step_bwd(const NpdeKParams& prm, PairField<M>& fld, float a, float dt,
                                               const float2 (&ys)[Stages<METHOD>::value], float* y_start) {
  const float sg = prm.sign;
  *y_start = ys[0].x;
  if (METHOD == BODE_EULER) {
    return a + fld.template vjp<false>(ys[0].x, ys[0].y, (sg * dt) * a, 1.f, nullptr);
  } else if (METHOD == BODE_MIDPOINT) {
    const float v2 = fld.template vjp<false>(ys[1].x, ys[1].y, (sg * dt) * a, 1.f, nullptr);
    const float v1 = fld.template vjp<false>(ys[0].x, ys[0].y, (sg * 0.5f * dt) * v2, 1.f, nullptr);
    return (a + v2) + v1;
  } else {
    constexpr int I2 = METHOD == BODE_RK4 ? 2 : 0, I3 = METHOD == BODE_RK4 ? 3 : 0;
    const float dt3 = dt * (1.f / 3.f);
    float kb1 = (dt * 0.125f) * a, kb2 = (dt * 0.375f) * a, kb3 = kb2;
    float yb = a;
    const float v4 = fld.template vjp<false>(ys[I3].x, ys[I3].y, sg * kb1, 1.f, nullptr);  // kb4 == initial kb1
    yb += v4;
    kb1 = fmaf(dt, v4, kb1);
    kb2 = fmaf(-dt, v4, kb2);
    kb3 = fmaf(dt, v4, kb3);
    const float v3 = fld.template vjp<false>(ys[I2].x, ys[I2].y, sg * kb3, 1.f, nullptr);
    yb += v3;
    kb1 = fmaf(-dt3, v3, kb1);
    kb2 = fmaf(dt, v3, kb2);
    const float v2 = fld.template vjp<false>(ys[1].x, ys[1].y, sg * kb2, 1.f, nullptr);
    yb += v2;
    kb1 = fmaf(dt3, v2, kb1);
    const float v1 = fld.template vjp<false>(ys[0].x, ys[0].y, sg * kb1, 1.f, nullptr);
    return yb + v1;
  }
}

// ------------------------------------------------------------------ one step of the augmented reverse solve (adjoint.py:32-55)
template <int METHOD, int M>
__device__ __forceinline__ void pair_step_aug(PairField<M>& fld, float& y, float& a, float dt, float s_in) {
  float f1, f2_, f3, f4;
  auto partner = [](float v) { return __shfl_xor_sync(FULL_MASK, v, 1); };
  if (METHOD == BODE_EULER) {
    const float j1 = fld.template vjp<true>(y, partner(y), a, -s_in * dt, &f1);
    y = fmaf(s_in * dt, f1, y);
    a = fmaf(-s_in * dt, j1, a);
  } else if (METHOD == BODE_MIDPOINT) {
    const float j1 = fld.template vjp<true>(y, partner(y), a, 0.f, &f1);
    const float h = 0.5f * dt * s_in;
    const float y2 = fmaf(h, f1, y), a2 = fmaf(-h, j1, a);
    const float j2 = fld.template vjp<true>(y2, partner(y2), a2, -s_in * dt, &f2_);
    y = fmaf(s_in * dt, f2_, y);
    a = fmaf(-s_in * dt, j2, a);
  } else {
    const float h = s_in * dt, h3 = h * (1.f / 3.f);
    const float j1 = fld.template vjp<true>(y, partner(y), a, -h * 0.125f, &f1);
    const float y2 = fmaf(h3, f1, y), a2 = fmaf(-h3, j1, a);
    const float j2 = fld.template vjp<true>(y2, partner(y2), a2, -h * 0.375f, &f2_);
    const float y3 = fmaf(h, f2_, fmaf(-h3, f1, y)), a3 = fmaf(-h, j2, fmaf(h3, j1, a));
    const float j3 = fld.template vjp<true>(y3, partner(y3), a3, -h * 0.375f, &f3);
    const float y4 = fmaf(h, (f1 - f2_) + f3, y), a4 = fmaf(-h, (j1 - j2) + j3, a);
    const float j4 = fld.template vjp<true>(y4, partner(y4), a4, -h * 0.125f, &f4);
    y = fmaf(h * 0.125f, (f1 + f4) + 3.f * (f2_ + f3), y);
    a = fmaf(-h * 0.125f, (j1 + j4) + 3.f * (j2 + j3), a);
  }
}

// ------------------------------------------------------------------ CTA prologue / epilogue of the pair kernels
// Same maths as project_W / npde_epilogue (npde_solve.cuh), but A = sf^2 Kzz^-1 L and Ksym are staged in shared memory first, so
// the projection W = A U, the back-projection gU = A^T gW and the prior read shared memory instead of chains of dependent
// global loads; every global load of the prologue is issued before the first barrier (one memory latency in total).
// 4-byte asynchronous global -> shared copies: every staging load of the prologue is in flight at once and none passes through
// registers (the register-staged loops paid one memory latency per loop).  src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void stage4(float* dst_smem, const void* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  const int nbytes = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void stage_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// Issues the copies of U (this CTA's particles), A = sf^2 Kzz^-1 L and Ksym; pair_project() waits for them.
__device__ __forceinline__ void pair_stage_issue(const NpdeKParams& prm, float* smem, bool with_prior) {
  const int m = prm.m, m2 = 2 * m, nout = prm.ppc * m2, mm = m * m;
  float* Us = smem;
  float* As = smem + prm.a_off;
  float* Ks = As + mm;
  const int p0 = blockIdx.x * prm.ppc;
  for (int idx = threadIdx.x; idx < nout; idx += blockDim.x) {
    const int q = idx / m2, r = idx - q * m2;
    const bool ok = p0 + q < prm.P;
    stage4(Us + idx, prm.U + (ok ? (long long)(p0 + q) * prm.U_stride + r : 0), ok);
  }
  for (int i = threadIdx.x; i < mm; i += blockDim.x) stage4(As + i, prm.A + i, true);
  if (with_prior)
    for (int i = threadIdx.x; i < mm; i += blockDim.x) stage4(Ks + i, prm.Ksym + i, true);
}

// Same maths as project_W (npde_solve.cuh), but A and U are read from shared memory, so the projection W = A U, the
// back-projection gU = A^T gW and the prior are not chains of dependent global loads.  `pred` is AND-reduced over the CTA by
// the closing barrier (the caller's "one observation per step" test rides on it).
__device__ __forceinline__ int pair_project(const NpdeKParams& prm, float* smem, int pred) {
  const int m = prm.m, m2 = 2 * m, nout = prm.ppc * m2;
  float* Us = smem;
  float* Ws = smem + nout;
  float* As = smem + prm.a_off;
  stage_wait_all();
  __syncthreads();
  for (int idx = threadIdx.x; idx < nout; idx += blockDim.x) {
    const int q = idx / m2, r = idx - q * m2, j = r >> 1, d = r & 1;
    const float* Uq = Us + q * m2 + d;
    const float* Aj = As + j * m;
    float acc = 0.f;
    for (int k = 0; k < m; ++k) acc = fmaf(Aj[k], Uq[2 * k], acc);
    Ws[idx] = acc;
  }
  return __syncthreads_and(pred);
}

template <int INJ, int M>
__device__ __forceinline__ void pair_epilogue(const NpdeKParams& prm, float* smem, const PairField<M>& fld, bool active, int pl, int n,
                                              int d, float r2x, float r2y) {
  const int tid = threadIdx.x;
  const int m = prm.m, m2 = 2 * m, mm = m * m;
  const int N = prm.N, ppc = prm.ppc;
  float* Us = smem;                      // [ppc][m2]   U of this CTA's particles
  float* Ws = Us + ppc * m2;             // [ppc][m2]   W = A U, later sum_n gW
  float* gWs = Ws + ppc * m2;            // [N][ppc][m2]
  float* red = gWs + N * ppc * m2;       // [ppc*N][2]  sum of squared residuals per pair
  const float* As = smem + prm.a_off;
  const float* Ks = As + mm;
  __syncthreads();   // everyone is done reading Ws
  if (active) {
    fld.store_gW(prm, gWs + (n * ppc + pl) * m2, d);
    if (d == 0) {
      red[(pl * N + n) * 2 + 0] = r2x;
      red[(pl * N + n) * 2 + 1] = r2y;
    }
  }
  __syncthreads();
  const int nout = ppc * m2;
  for (int idx = tid; idx < nout; idx += blockDim.x) {
    float acc = 0.f;
    for (int nn = 0; nn < N; ++nn) acc += gWs[nn * nout + idx];
    Ws[idx] = acc;
  }
  __syncthreads();
  float* pri = gWs;                      // reuse: prior partials [ppc][m2]
  // SVGD score tiles (svgd_tiles.cuh): vsign * (gU | glogsn) of the CTA's particles, collected in the free part of gWs and written
  // four particles at a time (vector stores) for the aligned quads inside the CTA's particle range, element stores at its ends
  const bool tiles = INJ == INJ_LIK && prm.VH != nullptr;
  const bool tiles4 = tiles && (N - 1) * nout >= ppc * (m2 + 2);
  float* gv = gWs + nout;                // [ppc][m2 + 2]
  for (int idx = tid; idx < nout; idx += blockDim.x) {
    const int q = idx / m2, r = idx - q * m2, k = r >> 1, dd = r & 1;
    const int pp = blockIdx.x * ppc + q;
    if (pp >= prm.P) { pri[idx] = 0.f; continue; }
    float acc = 0.f;
    const float* Wq = Ws + q * m2 + dd;
    for (int j = 0; j < m; ++j) acc = fmaf(As[j * m + k], Wq[2 * j], acc);
    float pr = 0.f;
    if (prm.add_prior) {
      const float* Uq = Us + q * m2 + dd;
      for (int j = 0; j < m; ++j) pr = fmaf(Ks[k * m + j], Uq[2 * j], pr);
      acc += pr;
      pr *= 0.5f * Us[idx];
    }
    pri[idx] = pr;
    const float gval = prm.scale * acc;
    prm.gU[(long long)pp * prm.gU_stride + r] = gval;
    if (tiles) {
      if (tiles4) gv[q * (m2 + 2) + r] = prm.vsign * gval;
      else score_tile_store(prm.VH, prm.VC, pp, r, prm.vsign * gval);
    }
  }
  if (INJ == INJ_LIK) {
    __syncthreads();
    if (tid < ppc) {
      const int pp = blockIdx.x * ppc + tid;
      if (pp < prm.P) {
        float sx = 0.f, sy = 0.f, pr = 0.f;
        for (int nn = 0; nn < N; ++nn) {
          sx += red[(tid * N + nn) * 2 + 0];
          sy += red[(tid * N + nn) * 2 + 1];
        }
        for (int j = 0; j < m2; ++j) pr += pri[tid * m2 + j];
        const float2 ls = *reinterpret_cast<const float2*>(prm.logsn + (long long)pp * prm.logsn_stride);
        const float ex = expf(-2.f * ls.x), ey = expf(-2.f * ls.y);
        const float nt = (float)N * (float)prm.T;
        prm.loss[pp] = prm.scale * (0.5f * (sx * ex + sy * ey) + nt * (ls.x + ls.y) + pr);
        prm.sqerr[pp] = sx + sy;
        const float gl0 = prm.scale * (nt - sx * ex), gl1 = prm.scale * (nt - sy * ey);
        prm.glogsn[(long long)pp * prm.glogsn_stride + 0] = gl0;
        prm.glogsn[(long long)pp * prm.glogsn_stride + 1] = gl1;
        if (tiles) {
          if (tiles4) {
            gv[tid * (m2 + 2) + m2] = prm.vsign * gl0;
            gv[tid * (m2 + 2) + m2 + 1] = prm.vsign * gl1;
          } else {
            score_tile_store(prm.VH, prm.VC, pp, m2, prm.vsign * gl0);
            score_tile_store(prm.VH, prm.VC, pp, m2 + 1, prm.vsign * gl1);
          }
        }
      }
    }
    if (tiles4) {
      __syncthreads();
      const int nf = m2 + 2;
      const int p0 = blockIdx.x * ppc, p1 = min(p0 + ppc, prm.P);          // this CTA's particles [p0, p1)
      const int jq0 = p0 >> 2, nq = ((p1 + 3) >> 2) - jq0;                   // the quads they touch
      for (int idx = tid; idx < nq * nf; idx += blockDim.x) {
        const int qi = idx / nf, f = idx - qi * nf;
        const int pp0 = 4 * (jq0 + qi);
        if (pp0 >= p0 && pp0 + 4 <= p1) {
          const float* g0 = gv + (pp0 - p0) * nf + f;
          const float v[4] = {g0[0], g0[nf], g0[2 * nf], g0[3 * nf]};
          score_tile_store4(prm.VH, prm.VC, jq0 + qi, f, v);
        } else {
          for (int e = 0; e < 4; ++e)
            if (pp0 + e >= p0 && pp0 + e < p1) score_tile_store(prm.VH, prm.VC, pp0 + e, f, gv[(pp0 + e - p0) * nf + f]);
        }
      }
    }
  }
}

// Thread -> (particle slot, trajectory, component).  Padding threads (past the CTA's pairs or past P) shadow a valid
// pair so that every warp-wide shuffle is executed by all 32 lanes; they never write.
struct PairIds {
  int pl, n, d, p;
  bool active;
  long long pair;
};
__device__ __forceinline__ PairIds pair_ids(const NpdeKParams& prm) {
  PairIds id;
  const int tid = threadIdx.x;
  id.d = tid & 1;
  int pairl = tid >> 1;
  const int npl = prm.ppc * prm.N;
  bool ok = pairl < npl;
  if (!ok) pairl = npl - 1;
  id.pl = pairl / prm.N;
  id.n = pairl - id.pl * prm.N;
  id.p = blockIdx.x * prm.ppc + id.pl;
  if (id.p >= prm.P) {
    ok = false;
    id.p = prm.P - 1;
    id.pl = id.p - blockIdx.x * prm.ppc;   // >= 0: the CTA's first particle always exists
  }
  id.active = ok;
  id.pair = (long long)id.p * prm.N + id.n;
  return id;
}

// ------------------------------------------------------------------ forward-only kernel: sol[T,P,N,2]
template <int M, int METHOD>
__global__ void __launch_bounds__(PairField<M>::MAX_THREADS) npde_pair_fwd_kernel(const __grid_constant__ NpdeKParams prm) {
  extern __shared__ __align__(16) float smem[];
  pair_stage_issue(prm, smem, false);
  pair_project(prm, smem, 1);
  const PairIds id = pair_ids(prm);
  PairField<M> fld;
  fld.load(prm, smem + (prm.ppc + id.pl) * 2 * prm.m, id.d);
  const long long PN2 = 2ll * prm.P * prm.N;
  float y = __ldg(prm.y0 + (prm.y0_stride ? 2ll * id.p * prm.N : 0) + 2 * id.n + id.d);
  float* sol = prm.sol + 2 * id.pair + id.d;
  if (id.active) sol[0] = y;
  for (int s = 0; s < prm.S; ++s) {
    y = pair_step_fwd<METHOD, false>(prm, fld, y, __ldg(prm.dt + s), nullptr, 0, false);
    const int j1 = __ldg(prm.obs_ptr + s + 1);
    for (int j = __ldg(prm.obs_ptr + s); j < j1; ++j)
      if (id.active) sol[(long long)j * PN2] = y;
  }
}

// ------------------------------------------------------------------ fused forward + closure + gradient kernel
// Per-lane context of the solve; OBS1 = every solver step emits exactly one observation (obs_ptr[s] == s + 1: the solver grid IS
// the observation grid, the default of FixedGridODESolver, solvers.py:50-51), which removes the observation loops and their
// index loads from both sweeps.
struct PairCtx {
  const float* sdt;
  const int* sptr;
  const float* Yd;       // Y[n][j][d] at Yd[2 j]
  const float* go;       // gout[j][pair][d]
  float* ckf;
  const float2* ck2;
  long long stride, PN2;
  float e2inv, y0;
  int d;
  bool active;
};

template <int M, int METHOD, int INJ, int ADJ, bool OBS1>
__device__ __forceinline__ void pair_solve_and_reverse(const NpdeKParams& prm, PairField<M>& fld, const PairCtx& c, float& a_out,
                                                       float& r2_out) {
  constexpr int STG = Stages<METHOD>::value;
  const float* sdt = c.sdt;
  const int* sptr = c.sptr;
  const float* Yd = c.Yd;
  const long long stride = c.stride, PN2 = c.PN2;
  const float e2inv = c.e2inv;
  const bool active = c.active;
  float* ckf = c.ckf;
  float r2 = 0.f;

  // ---------------- forward
  float y = c.y0;
  if (INJ == INJ_LIK) {
    const float r = Yd[0] - y;
    r2 = r * r;
  }
  if (ADJ == BODE_GRAD_ADJOINT && active) ckf[0] = y;
  for (int s = 0; s < prm.S; ++s) {
    if (ADJ == BODE_GRAD_DISCRETE) {
      y = pair_step_fwd<METHOD, true>(prm, fld, y, sdt[s], ckf + 2ll * s * STG * stride, 2 * stride, active);
    } else {
      y = pair_step_fwd<METHOD, false>(prm, fld, y, sdt[s], nullptr, 0, false);
    }
    if (OBS1) {
      if (INJ == INJ_LIK) {
        const float r = Yd[2 * (s + 1)] - y;
        r2 = fmaf(r, r, r2);
      }
      if (ADJ == BODE_GRAD_ADJOINT && active) ckf[2ll * (s + 1) * stride] = y;
    } else {
      const int j1 = sptr[s + 1];
      for (int j = sptr[s]; j < j1; ++j) {
        if (INJ == INJ_LIK) {
          const float r = Yd[2 * j] - y;
          r2 = fmaf(r, r, r2);
        }
        if (ADJ == BODE_GRAD_ADJOINT && active) ckf[2ll * j * stride] = y;
      }
    }
  }

  // ---------------- backward
  __syncwarp();   // the partner lane's checkpoint stores are read below
  float a = 0.f;
  if (ADJ == BODE_GRAD_DISCRETE) {
    float yend = y;
    float2 ys[STG], yn[STG];
    if (prm.S > 0) pair_load_stage_points<METHOD>(ys, c.ck2 + (long long)(prm.S - 1) * STG * stride, stride, c.d);
    for (int s = prm.S - 1; s >= 0; --s) {
      // software prefetch: the stage points of step s-1 travel from L2 while step s is differentiated
      if (s > 0) pair_load_stage_points<METHOD>(yn, c.ck2 + (long long)(s - 1) * STG * stride, stride, c.d);
      if (OBS1) {
        if (INJ == INJ_LIK) a = fmaf(yend - Yd[2 * (s + 1)], e2inv, a);
        else a += __ldg(c.go + (long long)(s + 1) * PN2);
      } else {
        const int j0 = sptr[s];
        for (int j = sptr[s + 1] - 1; j >= j0; --j) {
          if (INJ == INJ_LIK) a = fmaf(yend - Yd[2 * j], e2inv, a);
          else a += __ldg(c.go + (long long)j * PN2);
        }
      }
      a = pair_step_bwd<METHOD>(prm, fld, a, sdt[s], ys, &yend);
#pragma unroll
      for (int i = 0; i < STG; ++i) ys[i] = yn[i];
    }
    if (INJ == INJ_LIK) a = fmaf(yend - Yd[0], e2inv, a);
    else a += __ldg(c.go);
  } else {
    // continuous adjoint, adjoint.py:57-95: restart from the stored forward value at every t[i]
    const float s_in = -prm.sign;
    {
      const float yT = ckf[2ll * (prm.T - 1) * stride];
      if (INJ == INJ_LIK) a = (yT - Yd[2 * (prm.T - 1)]) * e2inv;
      else a = __ldg(c.go + (long long)(prm.T - 1) * PN2);
    }
    for (int i = prm.T - 1; i >= 1; --i) {
      float yy = ckf[2ll * i * stride];
      const int q1 = __ldg(prm.adj_ptr + i);
      for (int q = __ldg(prm.adj_ptr + i - 1); q < q1; ++q) pair_step_aug<METHOD>(fld, yy, a, __ldg(prm.adj_dt + q), s_in);
      const float yp = ckf[2ll * (i - 1) * stride];
      if (INJ == INJ_LIK) a = fmaf(yp - Yd[2 * (i - 1)], e2inv, a);
      else a += __ldg(c.go + (long long)(i - 1) * PN2);
    }
  }
  a_out = a;
  r2_out = r2;
}

template <int M, int METHOD, int INJ, int ADJ>
__global__ void __launch_bounds__(PairField<M>::MAX_THREADS) npde_pair_grad_kernel(const __grid_constant__ NpdeKParams prm) {
  extern __shared__ __align__(16) float smem[];
  const int N = prm.N;
  // solver grid, observations, U / A / Ksym: one batch of asynchronous copies, one memory latency
  float* sdt = smem + prm.stage_off;
  int* sptr = reinterpret_cast<int*>(sdt + prm.S);
  float* sY = sdt + ((2 * prm.S + 2) & ~1);
  pair_stage_issue(prm, smem, prm.add_prior != 0);
  for (int i = threadIdx.x; i < prm.S; i += blockDim.x) stage4(sdt + i, prm.dt + i, true);
  if (prm.S > 0) {
    for (int i = threadIdx.x; i <= prm.S; i += blockDim.x) stage4(reinterpret_cast<float*>(sptr + i), prm.obs_ptr + i, true);
  } else if (threadIdx.x == 0) {
    sptr[0] = 1;
  }
  if (INJ == INJ_LIK)
    for (int i = threadIdx.x; i < 2 * N * prm.T; i += blockDim.x) stage4(sY + i, prm.Y + i, true);
  stage_wait_all();
  __syncthreads();
  int one_obs = ADJ == BODE_GRAD_DISCRETE ? 1 : 0;          // the fast path is instantiated for the discrete adjoint only
  for (int i = threadIdx.x; i <= prm.S; i += blockDim.x) one_obs &= (sptr[i] == i + 1);
  one_obs = pair_project(prm, smem, one_obs);

  const PairIds id = pair_ids(prm);
  const int d = id.d;
  PairField<M> fld;
  fld.load(prm, smem + (prm.ppc + id.pl) * 2 * prm.m, d);
  PairCtx c;
  c.sdt = sdt;
  c.sptr = sptr;
  c.Yd = sY + 2 * id.n * prm.T + d;
  c.go = prm.gout + 2 * id.pair + d;
  c.e2inv = 0.f;                                            // dL/dx_d = -e2inv (Y_d - x_d)
  if (INJ == INJ_LIK) c.e2inv = expf(-2.f * __ldg(prm.logsn + (long long)id.p * prm.logsn_stride + d));
  c.stride = prm.npairs;                                    // checkpoint slot stride in float2
  c.PN2 = 2ll * prm.P * N;
  c.ckf = reinterpret_cast<float*>(prm.ck) + 2 * id.pair + d;
  c.ck2 = prm.ck + id.pair;
  c.y0 = __ldg(prm.y0 + (prm.y0_stride ? 2ll * id.p * N : 0) + 2 * id.n + d);
  c.d = d;
  c.active = id.active;

  float a, r2;
  if (ADJ == BODE_GRAD_DISCRETE && one_obs) pair_solve_and_reverse<M, METHOD, INJ, ADJ, ADJ == BODE_GRAD_DISCRETE>(prm, fld, c, a, r2);
  else pair_solve_and_reverse<M, METHOD, INJ, ADJ, false>(prm, fld, c, a, r2);
  if (prm.gy0 != nullptr && id.active) prm.gy0[2 * id.pair + d] = prm.scale * a;

  const float r2o = __shfl_xor_sync(FULL_MASK, r2, 1);
  pair_epilogue<INJ>(prm, smem, fld, id.active, id.pl, id.n, d, r2, r2o);
}

}  // namespace bode
