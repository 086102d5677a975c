// HAMCMC: L-BFGS-preconditioned Langevin (samplers/langevin.py:619-1107; Simsekli et al. 2016 via the product-form BFGS of
// Zhang & Sutton), one CTA per chain, bug-compatible with the reference (see oracle/samplers.py::HAMCMC for the list):
//   warm-up  step_without_metric  :941-964  theta <- theta - lr g - lr n, history append (:902-939) and the start-up pairs
//   metric   step                 :966-1000 base = params[M-1]; (Hg, Sn) = _compute_vector_prod(g, n) (:717-860);
//                                           theta_new = base - lr Hg - lr Sn; _update_metric_vars (:862-900)
// The reference materialises d x d outer products and B0 = eye(d)/H_gamma; here every operator is applied as dot + axpy over
// length-d vectors (O(K^2 d) per step), each thread owning a fixed strided slice so elementwise updates need no barrier and
// only the dot products synchronise the CTA.
#include <cstdlib>
#include "common.cuh"
#include "hamcmc_sliced.cuh"

namespace bode {

struct HamcmcArgs {
  int P, d, M;                     // M = memory + 1
  float *hist_theta, *hist_grad;   // [P][2M-1][d]
  float *pair_s, *pair_y;          // [P][M-1][d]
  float* work;                     // [P][4(M-1)+2][d]  u, v, p, q, z, z2
  int* meta;                       // [P][4]  n_hist, head, K, pair_head
  float* theta; long long ld_theta;
  const float* grad; long long ld_grad;
  const float* xi;                 // [P][d] injected standard normals or null
  float lr, H_gamma, trust_reg;
  int mode, add_params, add_noise;
  unsigned long long seed; unsigned int step;
  int* status;
};

// Dot products accumulate in float64 (the product of two floats is exact there, so a length-d dot is rounded ONCE, when it is
// returned): the product-form BFGS recursion subtracts projections of nearly parallel vectors and pairs whose curvature differs
// by 1e7 are common on the 16 x 16 npde posterior -- fp32 accumulation cost two digits of the update there.  The vectors stay fp32
// (HBM traffic unchanged); ~50 dots of d <= 514 per chain and step are nothing against the FP64 rate.  Fixed order: deterministic.
__device__ __forceinline__ double block_sum(double v, double* red) {
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

__device__ __forceinline__ float bdot(const float* a, const float* b, int d, double* red) {
  double acc = 0.0;
  for (int e = threadIdx.x; e < d; e += blockDim.x) acc = fma((double)a[e], (double)b[e], acc);
  return (float)block_sum(acc, red);
}

// minimal Philox for in-kernel noise (same generator as samplers.cu)
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned int idx, unsigned int step) {
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

__global__ void __launch_bounds__(128) hamcmc_generic_kernel(const HamcmcArgs a) {
  __shared__ double red[8];
  const int p = blockIdx.x, d = a.d, M = a.M, cap = 2 * M - 1;
  float* ht = a.hist_theta + (long long)p * cap * d;
  float* hg = a.hist_grad + (long long)p * cap * d;
  float* ps = a.pair_s + (long long)p * (M - 1) * d;
  float* py = a.pair_y + (long long)p * (M - 1) * d;
  float* wk = a.work + (long long)p * (4 * (M - 1) + 2) * d;
  float *U = wk, *V = wk + (M - 1) * d, *Pp = wk + 2 * (M - 1) * d, *Q = wk + 3 * (M - 1) * d;
  float *z = wk + 4 * (M - 1) * d, *z2 = z + d;
  int* meta = a.meta + 4 * p;
  float* th = a.theta + (long long)p * a.ld_theta;
  const float* g = a.grad + (long long)p * a.ld_grad;
  const float nscale = rsqrtf(0.5f * a.lr);
  int bad = 0;
  auto noise_at = [&](int e) -> float {
    const float x = a.xi ? a.xi[(long long)p * d + e] : philox_normal(a.seed, (unsigned)(p * d + e), a.step);
    return x * nscale;
  };

  if (a.mode == 0) {
    // ---------------- step_without_metric (:941-964)
    int n_hist = meta[0];
    for (int e = threadIdx.x; e < d; e += blockDim.x) {
      float t = th[e];
      bad |= !(fabsf(t) <= 3.4028234e38f);
      t = fmaf(-a.lr, g[e], t);
      if (a.add_noise) t = fmaf(-a.lr, noise_at(e), t);
      th[e] = t;
      if (a.add_params && n_hist < cap) {
        ht[(long long)n_hist * d + e] = t;          // theta AFTER the update, gradient from BEFORE it (:954-961)
        hg[(long long)n_hist * d + e] = g[e];
      }
    }
    if (a.add_params && n_hist < cap) ++n_hist;
    int K = meta[2];
    if (a.add_params && n_hist == cap && meta[0] == cap - 1) {
      // history just became full: start-up pairs i <-> i+M with the 1e-4 curvature filter (:924-935)
      __syncthreads();
      K = 0;
      for (int i = 0; i < M - 1; ++i) {
        double sy = 0.0, ss = 0.0;
        for (int e = threadIdx.x; e < d; e += blockDim.x) {
          const float s = ht[(long long)(i + M) * d + e] - ht[(long long)i * d + e];
          const float y = hg[(long long)(i + M) * d + e] - hg[(long long)i * d + e] + a.trust_reg * s;
          z[e] = s; z2[e] = y;
          sy = fma((double)s, (double)y, sy); ss = fma((double)s, (double)s, ss);
        }
        sy = block_sum(sy, red);
        ss = block_sum(ss, red);
        if (sy > 1e-4 * ss) {
          for (int e = threadIdx.x; e < d; e += blockDim.x) { ps[(long long)K * d + e] = z[e]; py[(long long)K * d + e] = z2[e]; }
          ++K;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { meta[0] = n_hist; meta[1] = 0; meta[2] = K; meta[3] = 0; }
  } else {
    // ---------------- metric step (:966-1000)
    const int head = meta[1], K = meta[2], phead = meta[3];
    const float B0 = 1.f / a.H_gamma, C0 = sqrtf(B0), S0 = rsqrtf(B0);
    const float* base = ht + (long long)((head + M - 1) % cap) * d;
    const float* gbase = hg + (long long)((head + M - 1) % cap) * d;
    int nu = 0;
    for (int i = 0; i < K; ++i) {
      const float* s = ps + (long long)((phead + i) % K) * d;
      const float* y = py + (long long)((phead + i) % K) * d;
      const float sy = bdot(s, y, d, red);
      if (sy < 0.f) continue;                                           // :825-829
      if (nu == 0) {
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = B0 * s[e];
      } else {
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = s[e];
        for (int j = nu - 1; j >= 0; --j) {                             // C^T z (:760-777)
          const float c = bdot(z, V + (long long)j * d, d, red);
          for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = fmaf(-c, U[(long long)j * d + e], z[e]);
        }
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] *= C0 * C0;   // C0 applied by C^T and again by C (:751,776)
        for (int j = 0; j < nu; ++j) {                                  // C z (:750-757)
          const float c = bdot(z, U + (long long)j * d, d, red);
          for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = fmaf(-c, V[(long long)j * d + e], z[e]);
        }
      }
      const float sBs = bdot(s, z, d, red);
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
        const float c = bdot(z, Q + (long long)j * d, d, red);
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = fmaf(-c, Pp[(long long)j * d + e], z[e]);
      }
      for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] *= S0 * S0;
      for (int j = 0; j < nu; ++j) {
        const float c = bdot(z, Pp + (long long)j * d, d, red);
        for (int e = threadIdx.x; e < d; e += blockDim.x) z[e] = fmaf(-c, Q[(long long)j * d + e], z[e]);
      }
    }
    for (int j = 0; j < nu; ++j) {
      const float c = bdot(z2, Pp + (long long)j * d, d, red);
      for (int e = threadIdx.x; e < d; e += blockDim.x) z2[e] = fmaf(-c, Q[(long long)j * d + e], z2[e]);
    }
    // theta_new = base - lr Hg - lr Sn ; s/y of the refresh pair (:862-873)
    double sy = 0.0, ss = 0.0;
    for (int e = threadIdx.x; e < d; e += blockDim.x) {
      float t = fmaf(-a.lr, z[e], base[e]);
      if (a.add_noise) t = fmaf(-a.lr, z2[e], t);
      bad |= !(fabsf(t) <= 3.4028234e38f);
      const float s = t - base[e];
      const float y = g[e] - gbase[e] + a.trust_reg * s;
      z[e] = s; z2[e] = y;
      th[e] = t;
      sy = fma((double)s, (double)y, sy); ss = fma((double)s, (double)s, ss);
    }
    sy = block_sum(sy, red);
    ss = block_sum(ss, red);
    int np_head = phead;
    if (sy > 1e-8 * ss && K > 0) {                                      // append + pop(0): the oldest pair is replaced
      for (int e = threadIdx.x; e < d; e += blockDim.x) { ps[(long long)phead * d + e] = z[e]; py[(long long)phead * d + e] = z2[e]; }
      np_head = (phead + 1) % K;
    }
    __syncthreads();                                                     // base/gbase fully consumed before the ring moves
    for (int e = threadIdx.x; e < d; e += blockDim.x) {                  // history: append new, pop oldest
      ht[(long long)head * d + e] = th[e];
      hg[(long long)head * d + e] = g[e];
    }
    if (threadIdx.x == 0) { meta[1] = (head + 1) % cap; meta[3] = np_head; }
  }
  if (bad && a.status) atomicOr(a.status, 1);
}


// ------------------------------------------------------------------ register-sliced kernel (d <= NT * EPT)
// The ~50 dot + axpy pairs of a metric step are ONE dependent chain per sampler chain.  In the generic kernel above every link is a
// strided loop over global memory: ncu (profiles/ncu_summary_r02.md, hamcmc) counted 76 M warp instructions per launch at d = 514 --
// 17 per element and link, most of them 64-bit address arithmetic -- for 0.18 ms per step, 4 % of the HBM roofline.  Here thread t
// owns elements t, t + NT, ... (EPT of them, compile time): the running vectors z, z2 and the current pair (s, y) stay in
// REGISTERS, the product-form vectors u, v, p, q (rebuilt on every step, never read by another launch) in shared memory at
// constant offsets from one base, and a block-wide sum costs one barrier (the partial sums alternate between two buffers).
// Element ownership, accumulation order (per-thread sequential in float64, xor tree, warps in order) and every elementwise
// expression are those of the generic kernel.
template <int EPT, int NT>
__global__ void __launch_bounds__(NT) hamcmc_kernel(const HamcmcArgs a) {
  __shared__ double red[2 * (NT / 32)];
  extern __shared__ __align__(16) float wsm[];          // u, v, p, q: [4][M-1][d]
  const int p = blockIdx.x, d = a.d, M = a.M, cap = 2 * M - 1;
  float* ht = a.hist_theta + (long long)p * cap * d;
  float* hg = a.hist_grad + (long long)p * cap * d;
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
    const float x = a.xi ? a.xi[(long long)p * d + e] : philox_normal(a.seed, (unsigned)(p * d + e), a.step);
    return x * nscale;
  };
  float gv[EPT];
  sl.load(gv, g);

  if (a.mode == 0) {
    // ---------------- step_without_metric (:941-964)
    const int n_hist0 = meta[0];
    int n_hist = n_hist0;
    const bool keep = a.add_params && n_hist < cap;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      if (!sl.ok(i)) continue;
      const int e = tid + NT * i;
      float t = th[e];
      bad |= !(fabsf(t) <= 3.4028234e38f);
      t = fmaf(-a.lr, gv[i], t);
      if (a.add_noise) t = fmaf(-a.lr, noise_at(e), t);
      th[e] = t;
      if (keep) {
        ht[(long long)n_hist * d + e] = t;          // theta AFTER the update, gradient from BEFORE it (:954-961)
        hg[(long long)n_hist * d + e] = gv[i];
      }
    }
    if (keep) ++n_hist;
    int K = meta[2];
    if (a.add_params && n_hist == cap && n_hist0 == cap - 1) {
      // history just became full: start-up pairs i <-> i+M with the 1e-4 curvature filter (:924-935); own elements only, so the
      // slot written above is read back by the thread that wrote it
      K = 0;
      for (int i = 0; i < M - 1; ++i) {
        float t1[EPT], t0[EPT], g1[EPT], g0[EPT], sv[EPT], yv[EPT];
        sl.load(t1, ht + (long long)(i + M) * d);
        sl.load(t0, ht + (long long)i * d);
        sl.load(g1, hg + (long long)(i + M) * d);
        sl.load(g0, hg + (long long)i * d);
        double sy = 0.0, ss = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
          sv[k] = t1[k] - t0[k];
          yv[k] = g1[k] - g0[k] + a.trust_reg * sv[k];
          if (sl.ok(k)) { sy = fma((double)sv[k], (double)yv[k], sy); ss = fma((double)sv[k], (double)sv[k], ss); }
        }
        sy = sl.bsum(sy);
        ss = sl.bsum(ss);
        if (sy > 1e-4 * ss) {
          sl.store(ps + (long long)K * d, sv);
          sl.store(py + (long long)K * d, yv);
          ++K;
        }
      }
    }
    __syncthreads();                                  // every thread has read meta
    if (tid == 0) { meta[0] = n_hist; meta[1] = 0; meta[2] = K; meta[3] = 0; }
  } else {
    // ---------------- metric step (:966-1000)
    const int head = meta[1], K = meta[2], phead = meta[3];
    const float S0 = rsqrtf(1.f / a.H_gamma);
    const float* base = ht + (long long)((head + M - 1) % cap) * d;
    const float* gbase = hg + (long long)((head + M - 1) % cap) * d;
    float bv[EPT], gbv[EPT], z[EPT], z2[EPT];
    sl.load(bv, base);                                // requested now, used after the recursion
    sl.load(gbv, gbase);
#pragma unroll
    for (int i = 0; i < EPT; ++i) z2[i] = sl.ok(i) ? S0 * noise_at(tid + NT * i) : 0.f;
    sliced_metric(sl, ps, py, K, phead, d, M, wsm, a.H_gamma, gv, z, z2);
    // theta_new = base - lr Hg - lr Sn ; s/y of the refresh pair (:862-873)
    double sy = 0.0, ss = 0.0;
    float tv[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      float t = fmaf(-a.lr, z[k], bv[k]);
      if (a.add_noise) t = fmaf(-a.lr, z2[k], t);
      if (sl.ok(k)) bad |= !(fabsf(t) <= 3.4028234e38f);
      const float s = t - bv[k];
      const float y = gv[k] - gbv[k] + a.trust_reg * s;
      z[k] = s; z2[k] = y; tv[k] = t;
      if (sl.ok(k)) { sy = fma((double)s, (double)y, sy); ss = fma((double)s, (double)s, ss); }
    }
    sl.store(th, tv);
    sy = sl.bsum(sy);
    ss = sl.bsum(ss);
    int np_head = phead;
    if (sy > 1e-8 * ss && K > 0) {                                      // append + pop(0): the oldest pair is replaced
      sl.store(ps + (long long)phead * d, z);
      sl.store(py + (long long)phead * d, z2);
      np_head = (phead + 1) % K;
    }
    sl.store(ht + (long long)head * d, tv);                             // history: append new, pop oldest (own elements: base was
    sl.store(hg + (long long)head * d, gv);                             // read into registers by the same thread)
    __syncthreads();                                                    // every thread has read meta
    if (tid == 0) { meta[1] = (head + 1) % cap; meta[3] = np_head; }
  }
  if (bad && a.status) atomicOr(a.status, 1);
}

template <int EPT, int NT>
static int launch_sliced(const HamcmcArgs& a, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    int e = check_cuda(cudaFuncSetAttribute(hamcmc_kernel<EPT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024), "hamcmc smem attr");
    if (e != BODE_OK) return e;
    attr_set = true;
  }
  hamcmc_kernel<EPT, NT><<<a.P, NT, smem, st>>>(a);
  return check_cuda(cudaGetLastError(), "hamcmc launch");
}

}  // namespace bode

using namespace bode;

/* floats needed for: which 0 -> hist_theta (= hist_grad), 1 -> pair_s (= pair_y), 2 -> work */
extern "C" size_t bode_hamcmc_floats(int32_t P, int32_t d, int32_t memory, int32_t which) {
  const size_t M = (size_t)memory + 1, pd = (size_t)P * d;
  if (which == 0) return pd * (2 * M - 1);
  if (which == 1) return pd * (M - 1);
  return pd * (4 * (M - 1) + 2);
}

extern "C" int bode_hamcmc_step(int32_t P, int32_t d, int32_t memory, float* hist_theta, float* hist_grad, float* pair_s,
                                float* pair_y, float* work, int32_t* meta, float* theta, int64_t ld_theta, const float* grad,
                                int64_t ld_grad, const float* xi, float lr, float H_gamma, float trust_reg, int32_t metric_step,
                                int32_t add_params, int32_t add_noise, uint64_t seed, uint32_t step, int32_t* status,
                                bode_stream_t stream) {
  BODE_REQUIRE(P > 0 && d > 0 && memory >= 1, "bad sizes P=%d d=%d memory=%d", P, d, memory);
  BODE_REQUIRE(hist_theta && hist_grad && pair_s && pair_y && work && meta && theta && grad, "null pointer");
  BODE_REQUIRE(lr > 0 && H_gamma > 0, "lr and H_gamma must be positive");
  HamcmcArgs a = {};
  a.P = P; a.d = d; a.M = memory + 1; a.hist_theta = hist_theta; a.hist_grad = hist_grad; a.pair_s = pair_s; a.pair_y = pair_y;
  a.work = work; a.meta = meta; a.theta = theta; a.ld_theta = ld_theta; a.grad = grad; a.ld_grad = ld_grad; a.xi = xi;
  a.lr = lr; a.H_gamma = H_gamma; a.trust_reg = trust_reg; a.mode = metric_step ? 1 : 0; a.add_params = add_params;
  a.add_noise = add_noise; a.seed = seed; a.step = step; a.status = status;
  // d <= 1024 with the product-form vectors (4 * memory * d floats) within 100 KB of shared memory: the register-sliced kernel;
  // anything larger: the generic kernel over the global work buffer (BODE_HAMCMC_GENERIC=1 forces it: the test that compares the
  // two).  Measured at d = 514, 2048 chains (c5): 128 threads x 5 elements 0.116 ms, 64 x 9 0.122, 32 x 17 0.165, generic 0.187.
  const size_t wbytes = sizeof(float) * (size_t)(4 * memory) * d;
  const cudaStream_t st = (cudaStream_t)stream;
  if (wbytes <= 100 * 1024 && !getenv("BODE_HAMCMC_GENERIC")) {
    if (d <= 32) return launch_sliced<1, 32>(a, wbytes, st);
    if (d <= 64) return launch_sliced<1, 64>(a, wbytes, st);
    if (d <= 128) return launch_sliced<1, 128>(a, wbytes, st);
    if (d <= 256) return launch_sliced<2, 128>(a, wbytes, st);
    if (d <= 384) return launch_sliced<3, 128>(a, wbytes, st);
    if (d <= 640) return launch_sliced<5, 128>(a, wbytes, st);
    if (d <= 1024) return launch_sliced<8, 128>(a, wbytes, st);
  }
  hamcmc_generic_kernel<<<P, 128, 0, st>>>(a);
  return check_cuda(cudaGetLastError(), "hamcmc launch");
}
