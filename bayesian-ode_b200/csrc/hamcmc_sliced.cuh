// Register-sliced building blocks shared by hamcmc.cu and hamcmc_contig.cu: thread t owns elements t, t + NT, ... (EPT of them,
// compile time) of every length-d vector of its chain.
#pragma once
#include "common.cuh"

namespace bode {

template <int EPT, int NT>
struct Sliced {
  int tid, d;
  double* red;      // [2][NT / 32]
  int par;
  __device__ __forceinline__ bool ok(int i) const { return tid + NT * i < d; }
  __device__ __forceinline__ double bsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    constexpr int NW = NT / 32;
    if (NW == 1) return 0.0 + v;
    if ((tid & 31) == 0) red[par * NW + (tid >> 5)] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += red[par * NW + i];
    par ^= 1;
    return t;
  }
  __device__ __forceinline__ void load(float (&r)[EPT], const float* src) const {
#pragma unroll
    for (int i = 0; i < EPT; ++i) r[i] = ok(i) ? src[tid + NT * i] : 0.f;
  }
  __device__ __forceinline__ void store(float* dst, const float (&r)[EPT]) const {
#pragma unroll
    for (int i = 0; i < EPT; ++i)
      if (ok(i)) dst[tid + NT * i] = r[i];
  }
  __device__ __forceinline__ float dot(const float (&x)[EPT], const float (&y)[EPT]) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < EPT; ++i)
      if (ok(i)) acc = fma((double)x[i], (double)y[i], acc);
    return (float)bsum(acc);
  }
  // z <- z - (z . x) y   with x, y in (shared) memory
  __device__ __forceinline__ void project(float (&z)[EPT], const float* x, const float* y) {
    float xv[EPT], yv[EPT];
    load(xv, x);
    load(yv, y);
    const float c = dot(z, xv);
#pragma unroll
    for (int i = 0; i < EPT; ++i) z[i] = fmaf(-c, yv[i], z[i]);
  }
};

// (Hg, S n) of langevin.py:717-860 from the K stored curvature pairs (ring starting at phead): builds the product-form vectors
// u, v, p, q in shared memory, then z <- S S^T g and z2 <- S z2 (z2 enters holding S0 * noise).  Same operation order as the
// generic kernels.
template <int EPT, int NT>
__device__ __forceinline__ void sliced_metric(Sliced<EPT, NT>& sl, const float* ps, const float* py, int K, int phead, int d, int M,
                                              float* wsm, float H_gamma, const float (&gv)[EPT], float (&z)[EPT], float (&z2)[EPT]) {
  float *U = wsm, *V = wsm + (M - 1) * d, *Pp = wsm + 2 * (M - 1) * d, *Q = wsm + 3 * (M - 1) * d;
  const float B0 = 1.f / H_gamma, C0 = sqrtf(B0), S0 = rsqrtf(B0);
  const int tid = sl.tid;
  int nu = 0;
  int slot = phead;
  for (int i = 0; i < K; ++i) {
    float sv[EPT], yv[EPT];
    sl.load(sv, ps + (long long)slot * d);
    sl.load(yv, py + (long long)slot * d);
    slot = slot + 1 == K ? 0 : slot + 1;
    const float sy = sl.dot(sv, yv);
    if (sy < 0.f) continue;                                           // :825-829
    if (nu == 0) {
#pragma unroll
      for (int k = 0; k < EPT; ++k) z[k] = B0 * sv[k];
    } else {
#pragma unroll
      for (int k = 0; k < EPT; ++k) z[k] = sv[k];
      for (int j = nu - 1; j >= 0; --j) sl.project(z, V + j * d, U + j * d);      // C^T z (:760-777)
#pragma unroll
      for (int k = 0; k < EPT; ++k) z[k] *= C0 * C0;                  // C0 applied by C^T and again by C (:751,776)
      for (int j = 0; j < nu; ++j) sl.project(z, U + j * d, V + j * d);           // C z (:750-757)
    }
    const float sBs = sl.dot(sv, z);
    const float cq = sqrtf(sy / sBs), cu = sqrtf(sBs / sy), isy = 1.f / sy, isBs = 1.f / sBs;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      if (!sl.ok(k)) continue;
      const int e = tid + NT * k;
      Q[nu * d + e] = cq * z[k] - yv[k];
      Pp[nu * d + e] = sv[k] * isy;
      U[nu * d + e] = cu + z[k];                                       // scalar + vector (:846)
      V[nu * d + e] = sv[k] * isBs;
    }
    ++nu;
  }
  // Hg = S (S^T g)   (:808-815) ; Sn = S n (:856)
#pragma unroll
  for (int k = 0; k < EPT; ++k) z[k] = gv[k];
  if (nu == 0) {
#pragma unroll
    for (int k = 0; k < EPT; ++k) z[k] = z[k] / B0;
  } else {
    for (int j = nu - 1; j >= 0; --j) sl.project(z, Q + j * d, Pp + j * d);
#pragma unroll
    for (int k = 0; k < EPT; ++k) z[k] *= S0 * S0;
    for (int j = 0; j < nu; ++j) sl.project(z, Pp + j * d, Q + j * d);
  }
  for (int j = 0; j < nu; ++j) sl.project(z2, Pp + j * d, Q + j * d);
}

}  // namespace bode
