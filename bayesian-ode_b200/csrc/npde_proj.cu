// Projections of the npde closure for LARGE inducing sets (general-Z path, m >= 64; BASELINE config 5: m = 256) as panel GEMMs over
// all particles instead of per-CTA loops inside the fused solve:
//   proj_W_kernel     W[p] = A U[p]                                   (gp.py:70: K(x,Z) (Kzz^-1 L U), hoisted to once per solve)
//   proj_back_kernel  gU[p] = scale (A^T gW[p] + Ksym U[p]),  loss[p] += scale/2 tr(U[p]^T Ksym U[p])     (gp.py:350 prior)
// Inside the solve kernel every CTA (one particle) streamed all of A three times from L2 -- 768 KB per particle, 1.5 GB at 2048
// particles -- and the prologue + epilogue were 1.27 ms of a 1.62 ms launch (0.5 ms after the loads were made coalesced).  Here a CTA
// takes a PANEL of 16 particles (32 columns (p, d)) into shared memory and thread r owns output row r for all 32 columns: one
// coalesced load of the matrix element per 32 FMAs, the panel values are shared-memory broadcasts.  Summation order over the
// contracted index is ascending, as in the in-kernel loops, so W is bit-identical to the fused path.
#include "npde_sep.cuh"

namespace bode {

constexpr int PROJ_PP = 16;            // particles per panel
constexpr int PROJ_C = 2 * PROJ_PP;    // columns per panel

__device__ __forceinline__ void load_panel(float* dst, const float* src, long long ld, int p0, int P, int m) {
  // dst[k][c], c = 2 pl + d  <-  src[(p0 + pl) ld + 2 k + d]; global reads run along a particle's row (coalesced)
  const int m2 = 2 * m;
  for (int idx = threadIdx.x; idx < PROJ_PP * m2; idx += blockDim.x) {
    const int pl = idx / m2, r = idx - pl * m2;
    const float v = (p0 + pl < P) ? __ldg(src + (long long)(p0 + pl) * ld + r) : 0.f;
    dst[(r >> 1) * PROJ_C + 2 * pl + (r & 1)] = v;
  }
}

__global__ void __launch_bounds__(256) proj_W_kernel(const float* __restrict__ AT, const float* __restrict__ U, long long U_stride, int P, int m,
                                                     float* __restrict__ W) {
  extern __shared__ __align__(16) float ps[];
  const int p0 = blockIdx.x * PROJ_PP;
  load_panel(ps, U, U_stride, p0, P, m);
  __syncthreads();
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    float acc[PROJ_C];
#pragma unroll
    for (int c = 0; c < PROJ_C; ++c) acc[c] = 0.f;
    for (int k = 0; k < m; ++k) {
      const float a = __ldg(AT + (long long)k * m + j);
      const float4* b = reinterpret_cast<const float4*>(ps + k * PROJ_C);
#pragma unroll
      for (int q = 0; q < PROJ_C / 4; ++q) {
        const float4 v = b[q];
        acc[4 * q + 0] = fmaf(a, v.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(a, v.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(a, v.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(a, v.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int pl = 0; pl < PROJ_PP; ++pl)
      if (p0 + pl < P) *reinterpret_cast<float2*>(W + (long long)(p0 + pl) * 2 * m + 2 * j) = make_float2(acc[2 * pl], acc[2 * pl + 1]);
  }
}

// gW arrives in the gU buffer (the solve kernel's split-mode epilogue) and is overwritten by gU: the panel is staged first
__global__ void __launch_bounds__(256) proj_back_kernel(const float* __restrict__ A, const float* __restrict__ Ksym, const float* __restrict__ U,
                                                        long long U_stride, float* __restrict__ gU, long long gU_stride, float* __restrict__ loss,
                                                        float scale, int add_prior, int P, int m) {
  extern __shared__ __align__(16) float ps[];
  float* Gs = ps;
  float* Us = ps + (size_t)m * PROJ_C;
  const int p0 = blockIdx.x * PROJ_PP;
  load_panel(Gs, gU, gU_stride, p0, P, m);
  if (add_prior) load_panel(Us, U, U_stride, p0, P, m);
  __syncthreads();
  float accg[PROJ_C], accp[PROJ_C];
  const int k = threadIdx.x;                      // m <= 256 == blockDim.x: one output row per thread
#pragma unroll
  for (int c = 0; c < PROJ_C; ++c) accg[c] = accp[c] = 0.f;
  if (k < m) {
    for (int j = 0; j < m; ++j) {
      const float a = __ldg(A + (long long)j * m + k);
      const float4* g = reinterpret_cast<const float4*>(Gs + j * PROJ_C);
#pragma unroll
      for (int q = 0; q < PROJ_C / 4; ++q) {
        const float4 v = g[q];
        accg[4 * q + 0] = fmaf(a, v.x, accg[4 * q + 0]);
        accg[4 * q + 1] = fmaf(a, v.y, accg[4 * q + 1]);
        accg[4 * q + 2] = fmaf(a, v.z, accg[4 * q + 2]);
        accg[4 * q + 3] = fmaf(a, v.w, accg[4 * q + 3]);
      }
      if (add_prior) {
        const float s = __ldg(Ksym + (long long)j * m + k);          // symmetric: row j, column k
        const float4* u = reinterpret_cast<const float4*>(Us + j * PROJ_C);
#pragma unroll
        for (int q = 0; q < PROJ_C / 4; ++q) {
          const float4 v = u[q];
          accp[4 * q + 0] = fmaf(s, v.x, accp[4 * q + 0]);
          accp[4 * q + 1] = fmaf(s, v.y, accp[4 * q + 1]);
          accp[4 * q + 2] = fmaf(s, v.z, accp[4 * q + 2]);
          accp[4 * q + 3] = fmaf(s, v.w, accp[4 * q + 3]);
        }
      }
    }
  }
  __syncthreads();                                // every thread has finished reading the gW panel
  if (k < m) {
#pragma unroll
    for (int pl = 0; pl < PROJ_PP; ++pl) {
      if (p0 + pl < P)
        *reinterpret_cast<float2*>(gU + (long long)(p0 + pl) * gU_stride + 2 * k) =
            make_float2(scale * (accg[2 * pl] + accp[2 * pl]), scale * (accg[2 * pl + 1] + accp[2 * pl + 1]));
      // prior partials 1/2 U (Ksym U), parked in the (consumed) gW panel in the [k][c] layout
      Gs[k * PROJ_C + 2 * pl] = add_prior ? 0.5f * Us[k * PROJ_C + 2 * pl] * accp[2 * pl] : 0.f;
      Gs[k * PROJ_C + 2 * pl + 1] = add_prior ? 0.5f * Us[k * PROJ_C + 2 * pl + 1] * accp[2 * pl + 1] : 0.f;
    }
  }
  __syncthreads();
  if (add_prior && loss && threadIdx.x < PROJ_PP && p0 + threadIdx.x < P) {
    const int pl = threadIdx.x;
    float pr = 0.f;
    for (int kk = 0; kk < m; ++kk) {             // r = 2 kk + d ascending, the order of the fused epilogue
      pr += Gs[kk * PROJ_C + 2 * pl];
      pr += Gs[kk * PROJ_C + 2 * pl + 1];
    }
    loss[p0 + pl] = fmaf(scale, pr, loss[p0 + pl]);
  }
}

int launch_proj_W(const float* AT, const float* U, long long U_stride, int P, int m, float* W, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)m * PROJ_C;
  proj_W_kernel<<<(P + PROJ_PP - 1) / PROJ_PP, 256, smem, st>>>(AT, U, U_stride, P, m, W);
  return check_cuda(cudaGetLastError(), "proj_W launch");
}

int launch_proj_back(const float* A, const float* Ksym, const float* U, long long U_stride, float* gU, long long gU_stride, float* loss,
                     float scale, int add_prior, int P, int m, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)m * PROJ_C * 2;
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    int e = check_cuda(cudaFuncSetAttribute(proj_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * PROJ_C * (int)sizeof(float)),
                       "proj_back smem attr");
    if (e != BODE_OK) return e;
    attr_set = true;
  }
  proj_back_kernel<<<(P + PROJ_PP - 1) / PROJ_PP, 256, smem, st>>>(A, Ksym, U, U_stride, gU, gU_stride, loss, scale, add_prior, P, m);
  return check_cuda(cudaGetLastError(), "proj_back launch");
}

}  // namespace bode
