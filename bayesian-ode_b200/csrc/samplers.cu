// Fused elementwise SG-MCMC parameter updates over flat theta[P*d] buffers (HBM-bound: one read of every
// operand, one write of every result, 128-bit accesses, grid-stride over a multiple of the SM count).
//
//   SGLD    samplers/langevin.py:173-202     p <- p - lr (g + n),            n = xi / sqrt(lr/2)
//   pSGLD   samplers/langevin.py:457-500     V <- a V + (1-a) g^2 ; G = 1/(lam + sqrt V) ; p <- p - lr (G g + sqrt(G) n)
//   aSGHMC  samplers/hamiltonian.py:38-99    adaptive SGHMC (burn-in statistics, stale tau_inv, optional resample)
//   axpy    SVGD update, p <- p + alpha x    (samplers/stein.py: the wrapped optimiser descends -phi)
//
// Noise: either injected (xi pointers, bit-parity with a reference run that used the same draws) or generated in
// kernel by counter-based Philox4x32-10 keyed on (seed, step) with the element index as counter.
// A non-finite parameter on entry sets *status |= 1 (langevin.py:184-185 raises ValueError; the host raises
// after the launch, asynchronously safe).
#include "common.cuh"

namespace bode {

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f); }

__device__ __forceinline__ float4 normal4(const Philox& rng, uint32_t idx4, uint32_t step, uint32_t stream) {
  const uint4 r = rng(idx4, step, stream, 0x5eedu);
  const float r0 = sqrtf(-2.f * __logf(u01(r.x))), r1 = sqrtf(-2.f * __logf(u01(r.z)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
  __sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

struct SamplerArgs {
  float *p, *g;
  float *s0, *s1, *s2, *s3;      // sampler state (V | tau, gbar, vhat, mom)
  const float *xi, *xi2;         // injected standard normals (may be null)
  long long n;
  float lr, alpha, lambda, mom_decay;
  int burn_in, resample, add_noise;
  uint64_t seed;
  uint32_t step;
  int* status;
  const bode_sampler_ctl* ctl;   // optional device-resident overrides (CUDA-graph replay with a changing schedule)
};

__device__ __forceinline__ SamplerArgs resolve(SamplerArgs a) {
  if (a.ctl) {
    a.lr = a.ctl->lr;
    a.step = a.ctl->step;
    a.burn_in = a.ctl->burn_in;
    a.resample = a.ctl->resample;
  }
  return a;
}

__device__ __forceinline__ float4 ld4(const float* p, long long i) { return *reinterpret_cast<const float4*>(p + i); }
__device__ __forceinline__ void st4(float* p, long long i, float4 v) { *reinterpret_cast<float4*>(p + i) = v; }
__device__ __forceinline__ bool bad(float x) { return !(fabsf(x) <= 3.4028234e38f); }

enum { K_SGLD = 0, K_PSGLD = 1, K_ASGHMC = 2 };

template <int KIND>
__device__ __forceinline__ void update1(const SamplerArgs& a, float& p, float g, float& s0, float& s1, float& s2, float& s3,
                                        float xi, float xi2) {
  if (KIND == K_SGLD) {
    const float n = a.add_noise ? xi * rsqrtf(0.5f * a.lr) : 0.f;
    p = fmaf(-a.lr, g + n, p);
  } else if (KIND == K_PSGLD) {
    s0 = a.alpha * s0 + (1.f - a.alpha) * g * g;
    const float G = 1.f / (a.lambda + sqrtf(s0));
    const float n = a.add_noise ? xi * rsqrtf(0.5f * a.lr) : 0.f;
    p = fmaf(-a.lr, G * g + sqrtf(G) * n, p);
  } else {
    // s0 = tau, s1 = gbar, s2 = v_hat, s3 = momentum            (hamiltonian.py:55-99)
    const float tau_inv = 1.f / (s0 + 1.f);                      // from the OLD tau (hamiltonian.py:70)
    if (a.burn_in) {
      s0 += -s0 * (s1 * s1 / (s2 + a.lambda)) + 1.f;
      s1 += -s1 * tau_inv + tau_inv * g;
      s2 += -s2 * tau_inv + tau_inv * (g * g);
    }
    const float minv = 1.f / (sqrtf(s2) + a.lambda);
    if (a.resample) s3 = xi2 * fminf(1.f / minv, 10.f);
    const float lr2 = a.lr * a.lr;
    const float sigma = sqrtf(fmaxf(2.f * lr2 * a.mom_decay * minv - lr2 * lr2, 1e-16f));
    s3 += -lr2 * minv * g - a.mom_decay * s3;
    if (a.add_noise) s3 += xi * sigma;
    p += s3;
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) sampler_kernel(const SamplerArgs a_in) {
  const SamplerArgs a = resolve(a_in);
  const Philox rng(a.seed);
  const long long n4 = a.n >> 2;
  int flag = 0;
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
    const long long i = i4 << 2;
    float4 p = ld4(a.p, i);
    const float4 g = ld4(a.g, i);
    float4 s0 = make_float4(0, 0, 0, 0), s1 = s0, s2 = s0, s3 = s0, xi = s0, xi2 = s0;
    if (KIND >= K_PSGLD) s0 = ld4(a.s0, i);
    if (KIND == K_ASGHMC) { s1 = ld4(a.s1, i); s2 = ld4(a.s2, i); s3 = ld4(a.s3, i); }
    if (a.add_noise) xi = a.xi ? ld4(a.xi, i) : normal4(rng, (uint32_t)i4, a.step, (uint32_t)(i4 >> 32) * 2u);
    if (KIND == K_ASGHMC && a.resample) xi2 = a.xi2 ? ld4(a.xi2, i) : normal4(rng, (uint32_t)i4, a.step, (uint32_t)(i4 >> 32) * 2u + 1u);
    flag |= bad(p.x) | bad(p.y) | bad(p.z) | bad(p.w);
    update1<KIND>(a, p.x, g.x, s0.x, s1.x, s2.x, s3.x, xi.x, xi2.x);
    update1<KIND>(a, p.y, g.y, s0.y, s1.y, s2.y, s3.y, xi.y, xi2.y);
    update1<KIND>(a, p.z, g.z, s0.z, s1.z, s2.z, s3.z, xi.z, xi2.z);
    update1<KIND>(a, p.w, g.w, s0.w, s1.w, s2.w, s3.w, xi.w, xi2.w);
    st4(a.p, i, p);
    if (KIND >= K_PSGLD) st4(a.s0, i, s0);
    if (KIND == K_ASGHMC) {
      if (a.burn_in) { st4(a.s1, i, s1); st4(a.s2, i, s2); }
      st4(a.s3, i, s3);
    }
  }
  // scalar tail (n not a multiple of 4): one thread per element
  const long long tail0 = n4 << 2;
  const long long t = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < a.n) {
    float p = a.p[t], s0 = 0, s1 = 0, s2 = 0, s3 = 0, xi = 0, xi2 = 0;
    const float g = a.g[t];
    if (KIND >= K_PSGLD) s0 = a.s0[t];
    if (KIND == K_ASGHMC) { s1 = a.s1[t]; s2 = a.s2[t]; s3 = a.s3[t]; }
    if (a.add_noise) xi = a.xi ? a.xi[t] : normal4(rng, (uint32_t)t, a.step, 0x7a11u).x;
    if (KIND == K_ASGHMC && a.resample) xi2 = a.xi2 ? a.xi2[t] : normal4(rng, (uint32_t)t, a.step, 0x7a12u).x;
    flag |= bad(p);
    update1<KIND>(a, p, g, s0, s1, s2, s3, xi, xi2);
    a.p[t] = p;
    if (KIND >= K_PSGLD) a.s0[t] = s0;
    if (KIND == K_ASGHMC) { a.s1[t] = s1; a.s2[t] = s2; a.s3[t] = s3; }
  }
  if (flag && a.status) atomicOr(a.status, 1);
}

__global__ void __launch_bounds__(256) axpy_kernel(float* p, const float* x, float alpha, long long n, int* status,
                                                   const bode_sampler_ctl* ctl) {
  if (ctl) alpha = ctl->lr;
  const long long n4 = n >> 2;
  int flag = 0;
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
    float4 a = ld4(p, i4 << 2);
    const float4 b = ld4(x, i4 << 2);
    flag |= bad(a.x) | bad(a.y) | bad(a.z) | bad(a.w);
    a.x = fmaf(alpha, b.x, a.x); a.y = fmaf(alpha, b.y, a.y); a.z = fmaf(alpha, b.z, a.z); a.w = fmaf(alpha, b.w, a.w);
    st4(p, i4 << 2, a);
  }
  const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) { flag |= bad(p[t]); p[t] = fmaf(alpha, x[t], p[t]); }
  if (flag && status) atomicOr(status, 1);
}

__global__ void fill_normal_kernel(float* out, long long n, uint64_t seed, uint32_t step) {
  const Philox rng(seed);
  const long long n4 = (n + 3) >> 2;
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
    const float4 v = normal4(rng, (uint32_t)i4, step, (uint32_t)(i4 >> 32) * 2u);
    const long long i = i4 << 2;
    if (i + 0 < n) out[i + 0] = v.x;
    if (i + 1 < n) out[i + 1] = v.y;
    if (i + 2 < n) out[i + 2] = v.z;
    if (i + 3 < n) out[i + 3] = v.w;
  }
}

__global__ void schedule_kernel(bode_sampler_ctl* ctl, int kind, double lr0, double gamma, double t0, double alpha,
                                uint32_t burn_in_iters, uint32_t resample_every) {
  const uint32_t it = ctl->next_iter;
  ctl->step = it;
  ctl->lr = (float)(kind == 1 ? lr0 / pow(t0 + alpha * (double)it, gamma) : lr0);
  const int burn = it < burn_in_iters;
  ctl->burn_in = burn;
  ctl->resample = (!burn && resample_every > 0 && ((it + 1u) % resample_every) == 0u) ? 1 : 0;
  ctl->next_iter = it + 1u;
}

static int grid_for(long long n, int* grid) {
  int dev = 0, sms = 0;
  BODE_CUDA(cudaGetDevice(&dev));
  BODE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sms * 8;          // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  *grid = (int)blocks;
  return BODE_OK;
}

static int check_ptrs(const void* p, const void* g, long long n) {
  BODE_REQUIRE(p && g && n > 0, "null p/g or n <= 0");
  BODE_REQUIRE(((uintptr_t)p & 15) == 0 && ((uintptr_t)g & 15) == 0, "p and g must be 16-byte aligned");
  return BODE_OK;
}


// ---------------------------------------------------------------- MALA accept / reject (samplers/langevin.py:57-95)
// One warp per chain: the two proposal terms are length-d sums (warp-shuffle reduction, fixed order), the decision is
// accepted = isfinite(log_alpha) && log(u) < log_alpha (:88).  theta_prev == NULL reproduces the reference as it runs:
// its saved state aliases the parameter it then updates in place (:45, :60), so both terms see theta_prev == theta_new and
// a rejection restores nothing.  With theta_prev the textbook ratio is used and rejected chains are restored in place.
__global__ void __launch_bounds__(256) mala_accept_kernel(const float* __restrict__ theta_prev, long long ld_prev, float* __restrict__ theta,
                                                          long long ld_theta, const float* __restrict__ g_prev, long long ld_gp,
                                                          const float* __restrict__ g_new, long long ld_gn,
                                                          const float* __restrict__ loss_prev, const float* __restrict__ loss_new,
                                                          const float* __restrict__ log_u, int P, int d, float lr, int restore,
                                                          uint64_t seed, uint32_t step, float* __restrict__ log_alpha_out,
                                                          int* __restrict__ accepted_out) {
  const int chain = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (chain >= P) return;
  const float* tp = theta_prev ? theta_prev + (long long)chain * ld_prev : nullptr;
  float* tn = theta + (long long)chain * ld_theta;
  const float* gp = g_prev + (long long)chain * ld_gp;
  const float* gn = g_new + (long long)chain * ld_gn;
  float rev = 0.f, fwd = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float dth = tp ? tp[k] - tn[k] : 0.f;         // theta_prev - theta_new
    const float r = fmaf(lr, gn[k], dth), f = fmaf(lr, gp[k], -dth);
    rev = fmaf(r, r, rev);
    fwd = fmaf(f, f, fwd);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rev += __shfl_xor_sync(0xffffffffu, rev, o);
    fwd += __shfl_xor_sync(0xffffffffu, fwd, o);
  }
  const float c = -1.f / (4.f * lr);
  const float la = (loss_prev[chain] - loss_new[chain]) + c * rev - c * fwd;
  float lu;
  if (log_u) lu = log_u[chain];
  else {
    const Philox rng(seed);
    lu = __logf(u01(rng((uint32_t)chain, step, 0x3a1au, 0x5eedu).x));
  }
  const bool acc = (fabsf(la) <= 3.4028234e38f) && (lu < la);
  if (lane == 0) {
    if (log_alpha_out) log_alpha_out[chain] = la;
    accepted_out[chain] = acc ? 1 : 0;
  }
  if (!acc && restore && tp)
    for (int k = lane; k < d; k += 32) tn[k] = tp[k];
}

}  // namespace bode

using namespace bode;

extern "C" int bode_sgld_step(float* p, const float* g, const float* xi, int64_t n, float lr, int32_t add_noise,
                              uint64_t seed, uint32_t step, int32_t* status, const bode_sampler_ctl* ctl, bode_stream_t stream) {
  int st = check_ptrs(p, g, n);
  if (st != BODE_OK) return st;
  SamplerArgs a = {};
  a.p = p; a.g = const_cast<float*>(g); a.xi = xi; a.n = n; a.lr = lr; a.add_noise = add_noise; a.seed = seed; a.step = step;
  a.status = status; a.ctl = ctl;
  int grid;
  st = grid_for(n, &grid);
  if (st != BODE_OK) return st;
  sampler_kernel<K_SGLD><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_cuda(cudaGetLastError(), "sgld launch");
}

extern "C" int bode_psgld_step(float* p, const float* g, float* V, const float* xi, int64_t n, float lr, float alpha,
                               float lambda, int32_t add_noise, uint64_t seed, uint32_t step, int32_t* status,
                               const bode_sampler_ctl* ctl, bode_stream_t stream) {
  int st = check_ptrs(p, g, n);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(V && ((uintptr_t)V & 15) == 0, "V must be a 16-byte aligned device pointer");
  SamplerArgs a = {};
  a.p = p; a.g = const_cast<float*>(g); a.s0 = V; a.xi = xi; a.n = n; a.lr = lr; a.alpha = alpha; a.lambda = lambda;
  a.add_noise = add_noise; a.seed = seed; a.step = step; a.status = status; a.ctl = ctl;
  int grid;
  st = grid_for(n, &grid);
  if (st != BODE_OK) return st;
  sampler_kernel<K_PSGLD><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_cuda(cudaGetLastError(), "psgld launch");
}

extern "C" int bode_asghmc_step(float* p, const float* g, float* tau, float* gbar, float* vhat, float* mom, const float* xi,
                                const float* xi_resample, int64_t n, float lr, float mom_decay, float lambda, int32_t burn_in,
                                int32_t resample, int32_t add_noise, uint64_t seed, uint32_t step, int32_t* status,
                                const bode_sampler_ctl* ctl, bode_stream_t stream) {
  int st = check_ptrs(p, g, n);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(tau && gbar && vhat && mom, "null aSGHMC state");
  BODE_REQUIRE((((uintptr_t)tau | (uintptr_t)gbar | (uintptr_t)vhat | (uintptr_t)mom) & 15) == 0, "state must be 16-byte aligned");
  SamplerArgs a = {};
  a.p = p; a.g = const_cast<float*>(g); a.s0 = tau; a.s1 = gbar; a.s2 = vhat; a.s3 = mom; a.xi = xi; a.xi2 = xi_resample;
  a.n = n; a.lr = lr; a.mom_decay = mom_decay; a.lambda = lambda; a.burn_in = burn_in; a.resample = resample;
  a.add_noise = add_noise; a.seed = seed; a.step = step; a.status = status; a.ctl = ctl;
  int grid;
  st = grid_for(n, &grid);
  if (st != BODE_OK) return st;
  sampler_kernel<K_ASGHMC><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_cuda(cudaGetLastError(), "asghmc launch");
}

extern "C" int bode_mala_accept(const float* theta_prev, int64_t ld_prev, float* theta, int64_t ld_theta, const float* grad_prev,
                                int64_t ld_gprev, const float* grad_new, int64_t ld_gnew, const float* loss_prev, const float* loss_new,
                                const float* log_u, int32_t P, int32_t d, float lr, int32_t restore, uint64_t seed, uint32_t step,
                                float* log_alpha, int32_t* accepted, bode_stream_t stream) {
  BODE_REQUIRE(theta && grad_prev && grad_new && loss_prev && loss_new && accepted, "null pointer");
  BODE_REQUIRE(P > 0 && d > 0 && lr > 0.f, "P, d, lr must be positive");
  BODE_REQUIRE(ld_theta >= d && ld_gprev >= d && ld_gnew >= d && (!theta_prev || ld_prev >= d), "row strides must be >= d");
  const long long threads = (long long)P * 32;
  mala_accept_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(theta_prev, ld_prev, theta, ld_theta, grad_prev, ld_gprev,
                                                                                    grad_new, ld_gnew, loss_prev, loss_new, log_u, P, d, lr,
                                                                                    restore, seed, step, log_alpha, accepted);
  return check_cuda(cudaGetLastError(), "mala_accept launch");
}

extern "C" int bode_axpy(float* p, const float* x, float alpha, int64_t n, int32_t* status, const bode_sampler_ctl* ctl,
                         bode_stream_t stream) {
  int st = check_ptrs(p, x, n);
  if (st != BODE_OK) return st;
  int grid;
  st = grid_for(n, &grid);
  if (st != BODE_OK) return st;
  axpy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, x, alpha, n, status, ctl);
  return check_cuda(cudaGetLastError(), "axpy launch");
}

extern "C" int bode_sampler_schedule(bode_sampler_ctl* ctl, int32_t kind, double lr0, double gamma, double t0, double alpha,
                                     uint32_t burn_in_iters, uint32_t resample_every, bode_stream_t stream) {
  BODE_REQUIRE(ctl, "null ctl");
  schedule_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctl, kind, lr0, gamma, t0, alpha, burn_in_iters, resample_every);
  return check_cuda(cudaGetLastError(), "schedule launch");
}

extern "C" int bode_fill_normal(float* out, int64_t n, uint64_t seed, uint32_t step, bode_stream_t stream) {
  BODE_REQUIRE(out && n > 0, "null out or n <= 0");
  int grid;
  int st = grid_for(n, &grid);
  if (st != BODE_OK) return st;
  fill_normal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n, seed, step);
  return check_cuda(cudaGetLastError(), "fill_normal launch");
}
