// HAMCMC2 / HAMCMC3 / HAMCMC4 (samplers/langevin.py:1109-1470): the variants of the L-BFGS-preconditioned Langevin sampler that
// form the (s, y) pairs from CONTIGUOUS samples.  Same product-form BFGS as hamcmc.cu (_compute_vector_prod, :717-860, is
// inherited by all three); what differs is the bookkeeping, restated from oracle/samplers.py::HAMCMCContiguous (pinned to
// reference runs of all three, tests/golden/hamcmc_contiguous.npz).  M = memory + 1 (:645), one CTA per chain:
//   warm-up  M plain Langevin steps, each storing (theta AFTER the update, gradient BEFORE it) (:1180-1203); when the M-th lands the
//            pairs are formed, with no curvature filter: variant 2 from entries 1..M-2 -> 2..M-1 (:1173-1178), variant 3 from
//            0..M-3 -> 1..M-2 (:1355-1359), variant 4 from 0..M-2 -> 1..M-1 (:1462-1466)
//   step     base = the OLDEST stored theta (variant 2, :1209) or the NEWEST (3 and 4, :1368); theta_new = base - lr Hg - lr Sn; the
//            new (theta, grad) is appended, ONE pair is added -- between the last two entries (variants 2 and 4, :1131-1134,
//            :1424-1427) or between the two entries before the last (variant 3, :1315-1318) -- and the oldest entry of every list
//            is dropped.
// Parity: tests/test_samplers_gpu.py::test_hamcmc_contiguous_variants_match_reference_runs (reference runs of all three, B200).
// A metric step before the window is full (the reference indexes an empty list there and raises) changes nothing and sets
// status bit 2; the host raises RuntimeError.
#include <cstdlib>
#include "common.cuh"
#include "hamcmc_sliced.cuh"

namespace bode {

struct HamcmcContigArgs {
  int P, d, M, variant;            // M = memory + 1
  float *hist_theta, *hist_grad;   // [P][M][d]      ring, meta[1] = index of the oldest entry
  float *pair_s, *pair_y;          // [P][M-1][d]    ring of K pairs (K = M-2 for variants 2 and 3, M-1 for 4), meta[3] = oldest
  float* work;                     // [P][4(M-1)+2][d]  u, v, p, q, z, z2
  int* meta;                       // [P][4]  n_hist, head, K, pair_head
  float* theta; long long ld_theta;
  const float* grad; long long ld_grad;
  const float* xi;                 // [P][d] injected standard normals or null
  float lr, H_gamma, trust_reg;
  int mode, update_metric, add_noise;
  unsigned long long seed; unsigned int step;
  int* status;
};

// Dot products accumulate in float64 (the product of two floats is exact there, so a length-d dot is rounded ONCE, when it is
// returned): the product-form BFGS recursion subtracts projections of nearly parallel vectors and pairs whose curvature differs
// by 1e7 are common on the 16 x 16 npde posterior -- fp32 accumulation cost two digits of the update there.  The vectors stay fp32
// (HBM traffic unchanged); ~50 dots of d <= 514 per chain and step are nothing against the FP64 rate.  Fixed order: deterministic.
__device__ __forceinline__ double hc_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ float hc_dot(const float* a, const float* b, int d, double* red) {
  double acc = 0.0;
  for (int e = threadIdx.x; e < d; e += blockDim.x) acc = fma((double)a[e], (double)b[e], acc);
  return (float)hc_block_sum(acc, red);
}

// the Philox stream of hamcmc.cu / samplers.cu
__device__ __forceinline__ float hc_philox_normal(unsigned long long seed, unsigned int idx, unsigned int step) {
  unsigned int c0 = idx >> 1, c1 = step, c2 = 0x4a3cu, c3 = 0x5eedu, a = (unsigned int)seed, b = (unsigned int)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
    a += 0x9E3779B9u; b += 0xBB67AE85u;
  }
  const float u1 = (float)(c0 >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
  const float u2 = (float)(c1 >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
  const float r = sqrtf(-2.f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return (idx & 1) ? r * s : r * c;
}

// Every thread owns the elements e = threadIdx.x + k blockDim.x of every length-d vector: elementwise updates need no barrier,
// only the dot products synchronise the CTA.
__global__ void __launch_bounds__(128) hamcmc_contig_generic_kernel(const HamcmcContigArgs a) {
  __shared__ double red[8];
  const int p = blockIdx.x, d = a.d, M = a.M;
  float* ht = a.hist_theta + (long long)p * M * d;
  float* hg = a.hist_grad + (long long)p * M * d;
  float* ps = a.pair_s + (long long)p * (M - 1) * d;
  float* py = a.pair_y + (long long)p * (M - 1) * d;
  float* wk = a.work + (long long)p * (4 * (M - 1) + 2) * d;
  float *U = wk, *V = wk + (long long)(M - 1) * d, *Pp = wk + 2ll * (M - 1) * d, *Q = wk + 3ll * (M - 1) * d;
  float *z = wk + 4ll * (M - 1) * d, *z2 = z + d;
  int* meta = a.meta + 4 * p;
  float* th = a.theta + (long long)p * a.ld_theta;
  const float* g = a.grad + (long long)p * a.ld_grad;
  const float nscale = rsqrtf(0.5f * a.lr);
  int bad = 0;
  auto noise_at = [&](int e) -> float {
    const float x = a.xi ? a.xi[(long long)p * d + e] : hc_philox_normal(a.seed, (unsigned)(p * d + e), a.step);
    return x * nscale;
  };
  const int n_hist0 = meta[0], head = meta[1], K0 = meta[2], phead = meta[3];

  if (a.mode == 0) {
    // ---------------- step_without_metric (:1180-1203) + _add_to_memory
    int n_hist = n_hist0, K = K0;
    const bool store = a.update_metric && n_hist < M;
    for (int e = threadIdx.x; e < d; e += blockDim.x) {
      float t = th[e];
      bad |= !(fabsf(t) <= 3.4028234e38f);
      t = fmaf(-a.lr, g[e], t);
      if (a.add_noise) t = fmaf(-a.lr, noise_at(e), t);
      th[e] = t;
      if (store) {
        ht[(long long)n_hist * d + e] = t;          // theta AFTER the update, gradient from BEFORE it
        hg[(long long)n_hist * d + e] = g[e];
      }
    }
    if (store) ++n_hist;
    if (store && n_hist == M) {
      // the window just became full: contiguous pairs, no curvature filter (own elements only: no barrier needed)
      const int first = a.variant == 2 ? 1 : 0;
      K = a.variant == 4 ? M - 1 : M - 2;
      for (int i = 0; i < K; ++i)
        for (int e = threadIdx.x; e < d; e += blockDim.x) {
          const float s = ht[(long long)(first + i + 1) * d + e] - ht[(long long)(first + i) * d + e];
          const float y = hg[(long long)(first + i + 1) * d + e] - hg[(long long)(first + i) * d + e] + a.trust_reg * s;
          ps[(long long)i * d + e] = s;
          py[(long long)i * d + e] = y;
        }
    }
    __syncthreads();                                 // every thread has read meta
    if (threadIdx.x == 0) { meta[0] = n_hist; meta[1] = 0; meta[2] = K; meta[3] = 0; }
  } else {
    // ---------------- metric step (:1205-1238 / :1364-1397)
    if (n_hist0 < M) {                                                    // window not full: nothing to build the metric from
      if (threadIdx.x == 0 && a.status) atomicOr(a.status, 2);
      return;
    }
    const int K = K0;
    const int newest = (head + M - 1) % M, prev = (head + M - 2) % M;
    const float B0 = 1.f / a.H_gamma, C0 = sqrtf(B0), S0 = rsqrtf(B0);
    const float* base = ht + (long long)(a.variant == 2 ? head : newest) * d;
    int nu = 0;
    for (int i = 0; i < K; ++i) {
      const float* s = ps + (long long)((phead + i) % K) * d;
      const float* y = py + (long long)((phead + i) % K) * d;
      const float sy = hc_dot(s, y, d, red);
      if (sy < 0.f) continue;                                           // :825-829
      if (nu == 0) {
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = B0 * s[e];
      } else {
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = s[e];
        for (int j = nu - 1; j >= 0; --j) {                             // C^T z (:760-777)
          const float c = hc_dot(z, V + (long long)j * d, d, red);
          for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = fmaf(-c, U[(long long)j * d + e], z[e]);
        }
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] *= C0 * C0;   // C0 applied by C^T and again by C (:751,776)
        for (int j = 0; j < nu; ++j) {                                  // C z (:750-757)
          const float c = hc_dot(z, U + (long long)j * d, d, red);
          for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = fmaf(-c, V[(long long)j * d + e], z[e]);
        }
      }
      const float sBs = hc_dot(s, z, d, red);
      const float cq = sqrtf(sy / sBs), cu = sqrtf(sBs / sy), isy = 1.f / sy, isBs = 1.f / sBs;
      for (int e = threadIdx.x; e < d; e += blockDim.x) {
        Q[(long long)nu * d + e] = cq * z[e] - y[e];
        Pp[(long long)nu * d + e] = s[e] * isy;
        U[(long long)nu * d + e] = cu + z[e];                            // scalar + vector (:846)
        V[(long long)nu * d + e] = s[e] * isBs;
      }
      ++nu;
    }
    // Hg = S (S^T g)   (:808-815) ; Sn = S n (:856)
    for (int e = threadIdx.x; e < d; e += blockDim.x) { z[e] = g[e]; z2[e] = S0 * noise_at(e); }
    if (nu == 0) {
      for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = z[e] / B0;
    } else {
      for (int j = nu - 1; j >= 0; --j) {
        const float c = hc_dot(z, Q + (long long)j * d, d, red);
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = fmaf(-c, Pp[(long long)j * d + e], z[e]);
      }
      for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] *= S0 * S0;
      for (int j = 0; j < nu; ++j) {
        const float c = hc_dot(z, Pp + (long long)j * d, d, red);
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = fmaf(-c, Q[(long long)j * d + e], z[e]);
      }
    }
    for (int j = 0; j < nu; ++j) {
      const float c = hc_dot(z2, Pp + (long long)j * d, d, red);
      for (int e = threadIdx.x; e < d; e += blockDim.x) z2[e] = fmaf(-c, Q[(long long)j * d + e], z2[e]);
    }
    // theta_new = base - lr Hg - lr Sn; the pair that joins the window; then the rings move (own elements only)
    const float* tn = ht + (long long)newest * d;
    const float* gn = hg + (long long)newest * d;
    const float* tp = ht + (long long)prev * d;
    const float* gp = hg + (long long)prev * d;
    for (int e = threadIdx.x; e < d; e += blockDim.x) {
      float t = fmaf(-a.lr, z[e], base[e]);
      if (a.add_noise) t = fmaf(-a.lr, z2[e], t);
      bad |= !(fabsf(t) <= 3.4028234e38f);
      float s, y;
      if (a.variant == 3) {                                              // between the two entries BEFORE the new one
        s = tn[e] - tp[e];
        y = gn[e] - gp[e] + a.trust_reg * s;
      } else {                                                           // between the newest stored entry and the new one
        s = t - tn[e];
        y = g[e] - gn[e] + a.trust_reg * s;
      }
      th[e] = t;
      if (K > 0) {                                                       // append + pop(0): the oldest pair is replaced
        ps[(long long)phead * d + e] = s;
        py[(long long)phead * d + e] = y;
      }
      ht[(long long)head * d + e] = t;                                   // history: append new, pop oldest
      hg[(long long)head * d + e] = g[e];
    }
    __syncthreads();                                                     // every thread has read meta
    if (threadIdx.x == 0) { meta[1] = (head + 1) % M; meta[3] = K > 0 ? (phead + 1) % K : 0; }
  }
  if (bad && a.status) atomicOr(a.status, 1);
}


// ------------------------------------------------------------------ register-sliced kernel (d <= NT * EPT), see hamcmc.cu
template <int EPT, int NT>
__global__ void __launch_bounds__(NT) hamcmc_contig_kernel(const HamcmcContigArgs a) {
  __shared__ double red[2 * (NT / 32)];
  extern __shared__ __align__(16) float wsm[];          // u, v, p, q: [4][M-1][d]
  const int p = blockIdx.x, d = a.d, M = a.M;
  float* ht = a.hist_theta + (long long)p * M * d;
  float* hg = a.hist_grad + (long long)p * M * d;
  float* ps = a.pair_s + (long long)p * (M - 1) * d;
  float* py = a.pair_y + (long long)p * (M - 1) * d;
  int* meta = a.meta + 4 * p;
  float* th = a.theta + (long long)p * a.ld_theta;
  const float* g = a.grad + (long long)p * a.ld_grad;
  const float nscale = rsqrtf(0.5f * a.lr);
  Sliced<EPT, NT> sl{(int)threadIdx.x, d, red, 0};
  const int tid = threadIdx.x;
  int bad = 0;
  auto noise_at = [&](int e) -> float {
    const float x = a.xi ? a.xi[(long long)p * d + e] : hc_philox_normal(a.seed, (unsigned)(p * d + e), a.step);
    return x * nscale;
  };
  const int n_hist0 = meta[0], head = meta[1], K0 = meta[2], phead = meta[3];
  float gv[EPT];
  sl.load(gv, g);

  if (a.mode == 0) {
    // ---------------- step_without_metric (:1180-1203) + _add_to_memory
    int n_hist = n_hist0, K = K0;
    const bool store = a.update_metric && n_hist < M;
    float tv[EPT];
    sl.load(tv, th);
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      if (!sl.ok(i)) continue;
      float t = tv[i];
      bad |= !(fabsf(t) <= 3.4028234e38f);
      t = fmaf(-a.lr, gv[i], t);
      if (a.add_noise) t = fmaf(-a.lr, noise_at(tid + NT * i), t);
      tv[i] = t;
    }
    sl.store(th, tv);
    if (store) {
      sl.store(ht + (long long)n_hist * d, tv);          // theta AFTER the update, gradient from BEFORE it
      sl.store(hg + (long long)n_hist * d, gv);
      ++n_hist;
    }
    if (store && n_hist == M) {
      // the window just became full: contiguous pairs, no curvature filter (own elements only: no barrier needed)
      const int first = a.variant == 2 ? 1 : 0;
      K = a.variant == 4 ? M - 1 : M - 2;
      for (int i = 0; i < K; ++i) {
        float t1[EPT], t0[EPT], g1[EPT], g0[EPT], sv[EPT], yv[EPT];
        sl.load(t1, ht + (long long)(first + i + 1) * d);
        sl.load(t0, ht + (long long)(first + i) * d);
        sl.load(g1, hg + (long long)(first + i + 1) * d);
        sl.load(g0, hg + (long long)(first + i) * d);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
          sv[k] = t1[k] - t0[k];
          yv[k] = g1[k] - g0[k] + a.trust_reg * sv[k];
        }
        sl.store(ps + (long long)i * d, sv);
        sl.store(py + (long long)i * d, yv);
      }
    }
    __syncthreads();                                 // every thread has read meta
    if (tid == 0) { meta[0] = n_hist; meta[1] = 0; meta[2] = K; meta[3] = 0; }
  } else {
    // ---------------- metric step (:1205-1238 / :1364-1397)
    if (n_hist0 < M) {                                                    // window not full: nothing to build the metric from
      if (tid == 0 && a.status) atomicOr(a.status, 2);
      return;
    }
    const int K = K0;
    const int newest = (head + M - 1) % M, prev = (head + M - 2) % M;
    const float S0 = rsqrtf(1.f / a.H_gamma);
    float bv[EPT], tn[EPT], gn[EPT], tp[EPT], gp[EPT], z[EPT], z2[EPT];
    sl.load(bv, ht + (long long)(a.variant == 2 ? head : newest) * d);   // requested now, used after the recursion
    sl.load(tn, ht + (long long)newest * d);
    sl.load(gn, hg + (long long)newest * d);
    if (a.variant == 3) {
      sl.load(tp, ht + (long long)prev * d);
      sl.load(gp, hg + (long long)prev * d);
    }
#pragma unroll
    for (int i = 0; i < EPT; ++i) z2[i] = sl.ok(i) ? S0 * noise_at(tid + NT * i) : 0.f;
    sliced_metric(sl, ps, py, K, phead, d, M, wsm, a.H_gamma, gv, z, z2);
    // theta_new = base - lr Hg - lr Sn; the pair that joins the window; then the rings move (own elements only)
    float tv[EPT], sv[EPT], yv[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      float t = fmaf(-a.lr, z[k], bv[k]);
      if (a.add_noise) t = fmaf(-a.lr, z2[k], t);
      if (sl.ok(k)) bad |= !(fabsf(t) <= 3.4028234e38f);
      if (a.variant == 3) {                                              // between the two entries BEFORE the new one
        sv[k] = tn[k] - tp[k];
        yv[k] = gn[k] - gp[k] + a.trust_reg * sv[k];
      } else {                                                           // between the newest stored entry and the new one
        sv[k] = t - tn[k];
        yv[k] = gv[k] - gn[k] + a.trust_reg * sv[k];
      }
      tv[k] = t;
    }
    sl.store(th, tv);
    if (K > 0) {                                                         // append + pop(0): the oldest pair is replaced
      sl.store(ps + (long long)phead * d, sv);
      sl.store(py + (long long)phead * d, yv);
    }
    sl.store(ht + (long long)head * d, tv);                              // history: append new, pop oldest
    sl.store(hg + (long long)head * d, gv);
    __syncthreads();                                                     // every thread has read meta
    if (tid == 0) { meta[1] = (head + 1) % M; meta[3] = K > 0 ? (phead + 1) % K : 0; }
  }
  if (bad && a.status) atomicOr(a.status, 1);
}

template <int EPT, int NT>
static int launch_contig_sliced(const HamcmcContigArgs& a, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    int e = check_cuda(cudaFuncSetAttribute(hamcmc_contig_kernel<EPT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024),
                       "hamcmc contiguous-variant smem attr");
    if (e != BODE_OK) return e;
    attr_set = true;
  }
  hamcmc_contig_kernel<EPT, NT><<<a.P, NT, smem, st>>>(a);
  return check_cuda(cudaGetLastError(), "hamcmc contiguous-variant launch");
}

}  // namespace bode

using namespace bode;

/* floats needed for: which 0 -> hist_theta (= hist_grad), 1 -> pair_s (= pair_y), 2 -> work */
extern "C" size_t bode_hamcmc_contig_floats(int32_t P, int32_t d, int32_t memory, int32_t which) {
  const size_t M = (size_t)memory + 1, pd = (size_t)P * d;
  if (which == 0) return pd * M;
  if (which == 1) return pd * (M - 1);
  return pd * (4 * (M - 1) + 2);
}

extern "C" int bode_hamcmc_contig_step(int32_t variant, int32_t P, int32_t d, int32_t memory, float* hist_theta, float* hist_grad,
                                       float* pair_s, float* pair_y, float* work, int32_t* meta, float* theta, int64_t ld_theta,
                                       const float* grad, int64_t ld_grad, const float* xi, float lr, float H_gamma, float trust_reg,
                                       int32_t metric_step, int32_t update_metric, int32_t add_noise, uint64_t seed, uint32_t step,
                                       int32_t* status, bode_stream_t stream) {
  BODE_REQUIRE(variant >= 2 && variant <= 4, "variant must be 2, 3 or 4 (HAMCMC2 / HAMCMC3 / HAMCMC4)");
  BODE_REQUIRE(P > 0 && d > 0 && memory >= 1, "bad sizes P=%d d=%d memory=%d", P, d, memory);
  BODE_REQUIRE(hist_theta && hist_grad && pair_s && pair_y && work && meta && theta && grad, "null pointer");
  BODE_REQUIRE(lr > 0 && H_gamma > 0, "lr and H_gamma must be positive");
  HamcmcContigArgs a = {};
  a.P = P; a.d = d; a.M = memory + 1; a.variant = variant; a.hist_theta = hist_theta; a.hist_grad = hist_grad; a.pair_s = pair_s;
  a.pair_y = pair_y; a.work = work; a.meta = meta; a.theta = theta; a.ld_theta = ld_theta; a.grad = grad; a.ld_grad = ld_grad;
  a.xi = xi; a.lr = lr; a.H_gamma = H_gamma; a.trust_reg = trust_reg; a.mode = metric_step ? 1 : 0; a.update_metric = update_metric;
  a.add_noise = add_noise; a.seed = seed; a.step = step; a.status = status;
  const size_t wbytes = sizeof(float) * (size_t)(4 * memory) * d;
  const cudaStream_t st = (cudaStream_t)stream;
  if (wbytes <= 100 * 1024 && !getenv("BODE_HAMCMC_GENERIC")) {          // as in bode_hamcmc_step
    if (d <= 32) return launch_contig_sliced<1, 32>(a, wbytes, st);
    if (d <= 64) return launch_contig_sliced<1, 64>(a, wbytes, st);
    if (d <= 128) return launch_contig_sliced<1, 128>(a, wbytes, st);
    if (d <= 256) return launch_contig_sliced<2, 128>(a, wbytes, st);
    if (d <= 384) return launch_contig_sliced<3, 128>(a, wbytes, st);
    if (d <= 640) return launch_contig_sliced<5, 128>(a, wbytes, st);
    if (d <= 1024) return launch_contig_sliced<8, 128>(a, wbytes, st);
  }
  hamcmc_contig_generic_kernel<<<P, 128, 0, st>>>(a);
  return check_cuda(cudaGetLastError(), "hamcmc contiguous-variant launch");
}
