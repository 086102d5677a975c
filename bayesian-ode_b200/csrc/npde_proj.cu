// Projections of the npde closure for LARGE inducing sets (general-Z path, m >= 64; BASELINE config 5: m = 256) as panel GEMMs over
// all particles instead of per-CTA loops inside the fused solve:
//   proj_W_kernel     W[p] = A U[p]                                   (gp.py:70: K(x,Z) (Kzz^-1 L U), hoisted to once per solve)
//   proj_back_kernel  gU[p] = scale (A^T gW[p] + Ksym U[p]),  loss[p] += scale/2 tr(U[p]^T Ksym U[p])     (gp.py:350 prior)
// Inside the solve kernel every CTA (one particle) streamed all of A three times from L2 -- 768 KB per particle, 1.5 GB at 2048
// particles -- and the prologue + epilogue were 1.27 ms of a 1.62 ms launch (0.5 ms after the loads were made coalesced).  Here a CTA
// takes a PANEL of 16 particles (32 columns (p, d)) into shared memory and a thread owns output row r for 16 of the columns: one
// coalesced load of the matrix element per 16 FMAs, the panel values are shared-memory broadcasts.  Summation order over the
// contracted index is ascending, as in the in-kernel loops, so W is bit-identical to the fused path.
#include "npde_sep.cuh"

namespace bode {

constexpr int PROJ_PP = 16;            // particles per panel
constexpr int PROJ_C = 2 * PROJ_PP;    // columns per panel
constexpr int PROJ_LD = PROJ_C + 4;    // panel row stride in shared memory: the transposing fill then hits 2 banks per warp store, not 16

constexpr int PF = 8;                  // matrix elements in flight per thread
constexpr int PROJ_T = 512;            // threads: row = tid & 255, column half = tid >> 8 (16 warps per SM hide the shared-memory latency)
constexpr int PROJ_H = PROJ_C / 2;     // columns per thread

// acc[c] += a * row[c] over this thread's 16 columns (row: shared memory, the same address for the whole warp)
__device__ __forceinline__ void axpy_panel(float (&acc)[PROJ_H], float a, const float* row) {
  const float4* b = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int q = 0; q < PROJ_H / 4; ++q) {
    const float4 v = b[q];
    acc[4 * q + 0] = fmaf(a, v.x, acc[4 * q + 0]);
    acc[4 * q + 1] = fmaf(a, v.y, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(a, v.z, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(a, v.w, acc[4 * q + 3]);
  }
}

__device__ __forceinline__ void load_panel(float* dst, const float* src, long long ld, int p0, int P, int m) {
  // dst[k][c], c = 2 pl + d  <-  src[(p0 + pl) ld + 2 k + d]; global reads run along a particle's row (coalesced)
  const int m2 = 2 * m;
  for (int idx = threadIdx.x; idx < PROJ_PP * m2; idx += blockDim.x) {
    const int pl = idx / m2, r = idx - pl * m2;
    const float v = (p0 + pl < P) ? __ldg(src + (long long)(p0 + pl) * ld + r) : 0.f;
    dst[(r >> 1) * PROJ_LD + 2 * pl + (r & 1)] = v;
  }
}

__global__ void __launch_bounds__(PROJ_T) proj_W_kernel(const float* __restrict__ AT, const float* __restrict__ U, long long U_stride, int P, int m,
                                                     float* __restrict__ W) {
  extern __shared__ __align__(16) float ps[];
  const int p0 = blockIdx.x * PROJ_PP;
  load_panel(ps, U, U_stride, p0, P, m);
  __syncthreads();
  const int j = threadIdx.x & 255, h = threadIdx.x >> 8;          // m <= 256
  if (j < m) {
    const float* pc = ps + h * PROJ_H;
    float acc[PROJ_H];
#pragma unroll
    for (int c = 0; c < PROJ_H; ++c) acc[c] = 0.f;
    // PF matrix elements are requested together: one L2 round trip per PF contracted indices
    for (int k0 = 0; k0 < m; k0 += PF) {
      float a[PF];
#pragma unroll
      for (int u = 0; u < PF; ++u) a[u] = (k0 + u < m) ? __ldg(AT + (long long)(k0 + u) * m + j) : 0.f;
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        if (k0 + u >= m) break;
        axpy_panel(acc, a[u], pc + (k0 + u) * PROJ_LD);
      }
    }
#pragma unroll
    for (int q = 0; q < PROJ_H / 2; ++q) {
      const int pl = h * (PROJ_PP / 2) + q;
      if (p0 + pl < P) *reinterpret_cast<float2*>(W + (long long)(p0 + pl) * 2 * m + 2 * j) = make_float2(acc[2 * q], acc[2 * q + 1]);
    }
  }
}

// gW arrives in the gU buffer (the solve kernel's split-mode epilogue) and is overwritten by gU: the panel is staged first
__global__ void __launch_bounds__(PROJ_T) proj_back_kernel(const float* __restrict__ A, const float* __restrict__ Ksym, const float* __restrict__ U,
                                                        long long U_stride, float* __restrict__ gU, long long gU_stride, float* __restrict__ loss,
                                                        float scale, int add_prior, int P, int m) {
  extern __shared__ __align__(16) float ps[];
  float* Gs = ps;
  float* Us = ps + (size_t)m * PROJ_LD;
  const int p0 = blockIdx.x * PROJ_PP;
  load_panel(Gs, gU, gU_stride, p0, P, m);
  if (add_prior) load_panel(Us, U, U_stride, p0, P, m);
  __syncthreads();
  float accg[PROJ_H], accp[PROJ_H];
  const int k = threadIdx.x & 255, h = threadIdx.x >> 8;          // m <= 256: one output row and one half of the columns per thread
  const float* gc = Gs + h * PROJ_H;
  const float* uc = Us + h * PROJ_H;
#pragma unroll
  for (int c = 0; c < PROJ_H; ++c) accg[c] = accp[c] = 0.f;
  if (k < m) {
    for (int j0 = 0; j0 < m; j0 += PF) {
      float a[PF], s[PF];
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        a[u] = (j0 + u < m) ? __ldg(A + (long long)(j0 + u) * m + k) : 0.f;
        s[u] = (add_prior && j0 + u < m) ? __ldg(Ksym + (long long)(j0 + u) * m + k) : 0.f;          // symmetric: row j, column k
      }
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        if (j0 + u >= m) break;
        axpy_panel(accg, a[u], gc + (j0 + u) * PROJ_LD);
        if (add_prior) axpy_panel(accp, s[u], uc + (j0 + u) * PROJ_LD);
      }
    }
  }
  __syncthreads();                                // every thread has finished reading the gW panel
  if (k < m) {
#pragma unroll
    for (int q = 0; q < PROJ_H / 2; ++q) {
      const int pl = h * (PROJ_PP / 2) + q;
      if (p0 + pl < P)
        *reinterpret_cast<float2*>(gU + (long long)(p0 + pl) * gU_stride + 2 * k) =
            make_float2(scale * (accg[2 * q] + accp[2 * q]), scale * (accg[2 * q + 1] + accp[2 * q + 1]));
      // prior partials 1/2 U (Ksym U), parked in the (consumed) gW panel in the [k][c] layout
      Gs[k * PROJ_LD + 2 * pl] = add_prior ? 0.5f * Us[k * PROJ_LD + 2 * pl] * accp[2 * q] : 0.f;
      Gs[k * PROJ_LD + 2 * pl + 1] = add_prior ? 0.5f * Us[k * PROJ_LD + 2 * pl + 1] * accp[2 * q + 1] : 0.f;
    }
  }
  __syncthreads();
  if (add_prior && loss && threadIdx.x < PROJ_PP && p0 + threadIdx.x < P) {
    const int pl = threadIdx.x;
    float pr = 0.f;
    for (int kk = 0; kk < m; ++kk) {             // r = 2 kk + d ascending, the order of the fused epilogue
      pr += Gs[kk * PROJ_LD + 2 * pl];
      pr += Gs[kk * PROJ_LD + 2 * pl + 1];
    }
    loss[p0 + pl] = fmaf(scale, pr, loss[p0 + pl]);
  }
}

int launch_proj_W(const float* AT, const float* U, long long U_stride, int P, int m, float* W, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)m * PROJ_LD;
  proj_W_kernel<<<(P + PROJ_PP - 1) / PROJ_PP, PROJ_T, smem, st>>>(AT, U, U_stride, P, m, W);
  return check_cuda(cudaGetLastError(), "proj_W launch");
}

int launch_proj_back(const float* A, const float* Ksym, const float* U, long long U_stride, float* gU, long long gU_stride, float* loss,
                     float scale, int add_prior, int P, int m, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)m * PROJ_LD * 2;
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    int e = check_cuda(cudaFuncSetAttribute(proj_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * PROJ_LD * (int)sizeof(float)),
                       "proj_back smem attr");
    if (e != BODE_OK) return e;
    attr_set = true;
  }
  proj_back_kernel<<<(P + PROJ_PP - 1) / PROJ_PP, PROJ_T, smem, st>>>(A, Ksym, U, U_stride, gU, gU_stride, loss, scale, add_prior, P, m);
  return check_cuda(cudaGetLastError(), "proj_back launch");
}

}  // namespace bode
