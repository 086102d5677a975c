// Projections of the npde closure for LARGE inducing sets (m >= 64; BASELINE config 5: m = 256) as panel GEMMs over all particles
// instead of per-CTA loops inside the fused solve:
//   proj_W_kernel     W[p] = A U[p]                                   (gp.py:70: K(x,Z) (Kzz^-1 L U), hoisted to once per solve)
//   proj_back_kernel  gU[p] = scale (A^T gW[p] + Ksym U[p]),  loss[p] += scale/2 tr(U[p]^T Ksym U[p])     (gp.py:350 prior)
// Inside the solve kernel every CTA (one particle) streamed all of A three times from L2 -- 768 KB per particle, 1.5 GB at 2048
// particles -- and the prologue + epilogue were 1.27 ms of a 1.62 ms launch (0.5 ms after the loads were made coalesced).  Here a CTA
// takes a PANEL of 16 particles (32 columns (p, d)) into shared memory and a thread owns a 4-row x 8-column register tile: per
// contracted index one 16-byte shared-memory load of the matrix (4 rows; the matrix streams from L2 through a 4-stage cp.async
// pipeline of 8-row tiles) and two 16-byte loads of the panel (8 columns) feed 16 FFMA2.
// (A first version with 1 x 16 tiles was bound by shared-memory bandwidth: a 16-byte broadcast load still costs four wavefronts,
// 64 bytes per 16 FMAs.)  Summation order over the contracted index is ascending, as in the in-kernel loops, and fma.rn.f32x2 rounds
// each half like fmaf, so W is bit-identical to the fused path.
#include "npde_sep.cuh"

namespace bode {

constexpr int PROJ_PP = 16;            // particles per panel
constexpr int PROJ_C = 2 * PROJ_PP;    // columns per panel
constexpr int PROJ_LD = PROJ_C + 4;    // panel row stride in shared memory: the transposing fill then hits 2 banks per warp store, not 16
constexpr int PROJ_T = 256;            // threads: row group = tid & 63 (rows 4 rg .. 4 rg + 3), column group = tid >> 6 (4 particles)
constexpr int TR = 4, TQ = 4;          // register tile: TR rows x TQ particles (2 TQ columns)
constexpr int KT = 8;                  // contracted indices per pipeline stage
constexpr int NST = 4;                 // cp.async stages in flight: the matrix streams from L2 under the FMAs
constexpr int TILE = KT * 256;         // floats per staged matrix tile ([KT][256], zero filled past m)

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = ok ? 16 : 0;                                 // src-size 0: nothing is read, the destination is zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// rows k0 .. k0+KT-1 of the row-major [m][m] matrix M_ -> dst[KT][256]
__device__ __forceinline__ void load_tile(float* dst, const float* M_, int k0, int m, bool vec) {
  if (vec) {
#pragma unroll
    for (int i = 0; i < KT * 64 / PROJ_T; ++i) {
      const int slot = threadIdx.x + PROJ_T * i, row = slot >> 6, col = 4 * (slot & 63);
      const bool ok = k0 + row < m && col < m;
      cp_async16(dst + row * 256 + col, ok ? M_ + (long long)(k0 + row) * m + col : M_, ok);
    }
  } else {
#pragma unroll
    for (int i = 0; i < KT * 256 / PROJ_T; ++i) {
      const int idx = threadIdx.x + PROJ_T * i, row = idx >> 8, col = idx & 255;
      const bool ok = k0 + row < m && col < m;
      cp_async4(dst + row * 256 + col, ok ? M_ + (long long)(k0 + row) * m + col : M_, ok);
    }
  }
}

__device__ __forceinline__ void load_panel(float* dst, const float* src, long long ld, int p0, int P, int m) {
  // dst[k][c], c = 2 pl + d  <-  src[(p0 + pl) ld + 2 k + d]; global reads run along a particle's row (coalesced), eight requests
  // in flight per thread (one request per trip cost a memory latency for each of the 32 trips: a quarter of the kernel in ncu)
  const int m2 = 2 * m, total = PROJ_PP * m2;
  constexpr int UB = 8;
  for (int base = threadIdx.x; base < total; base += UB * PROJ_T) {
    float v[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int idx = base + u * PROJ_T;
      const int pl = idx / m2, r = idx - pl * m2;
      v[u] = (idx < total && p0 + pl < P) ? __ldg(src + (long long)(p0 + pl) * ld + r) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int idx = base + u * PROJ_T;
      const int pl = idx / m2, r = idx - pl * m2;
      if (idx < total) dst[(r >> 1) * PROJ_LD + 2 * pl + (r & 1)] = v[u];
    }
  }
}

// acc[i][q] += a_i * (b[2q], b[2q+1]) : 16 FFMA2
__device__ __forceinline__ void tile_fma(f32x2 (&acc)[TR][TQ], float4 a, const float* brow) {
  const ulonglong2* b = reinterpret_cast<const ulonglong2*>(brow);
  const ulonglong2 b0 = b[0], b1 = b[1];
  const f32x2 bq[TQ] = {b0.x, b0.y, b1.x, b1.y};
  const float av[TR] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < TR; ++i) {
    const f32x2 aa = pk(av[i], av[i]);
#pragma unroll
    for (int q = 0; q < TQ; ++q) acc[i][q] = fma2x(aa, bq[q], acc[i][q]);
  }
}

__global__ void __launch_bounds__(PROJ_T) proj_W_kernel(const float* __restrict__ AT, const float* __restrict__ U, long long U_stride, int P, int m,
                                                        float* __restrict__ W) {
  extern __shared__ __align__(16) float ps[];
  float* As = ps + (size_t)m * PROJ_LD;                  // [NST][KT][256]
  const int p0 = blockIdx.x * PROJ_PP;
  const bool vec = (m & 3) == 0;
  const int ntile = (m + KT - 1) / KT;
  for (int s = 0; s < NST - 1; ++s) {
    if (s < ntile) load_tile(As + s * TILE, AT, s * KT, m, vec);
    cp_commit();
  }
  load_panel(ps, U, U_stride, p0, P, m);
  const int r0 = TR * (threadIdx.x & 63), cg = threadIdx.x >> 6;          // m <= 256
  const float* pc = ps + 2 * TQ * cg;
  f32x2 acc[TR][TQ];
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int q = 0; q < TQ; ++q) acc[i][q] = pk(0.f, 0.f);
  // W[j] = sum_k A[j][k] U[k] = sum_k AT[k][j] U[k]: row k of AT, columns r0 .. r0+3
  for (int t = 0; t < ntile; ++t) {
    cp_wait<NST - 2>();
    __syncthreads();                                     // tile t (and, at t = 0, the panel) visible; tile t-1 consumed by everyone
    if (t + NST - 1 < ntile) load_tile(As + ((t + NST - 1) % NST) * TILE, AT, (t + NST - 1) * KT, m, vec);
    cp_commit();
    const float* at = As + (t % NST) * TILE + r0;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const int k = t * KT + kk;
      if (k < m) tile_fma(acc, *reinterpret_cast<const float4*>(at + kk * 256), pc + k * PROJ_LD);
    }
  }
  if (r0 >= m) return;
#pragma unroll
  for (int q = 0; q < TQ; ++q) {
    const int pl = TQ * cg + q;
    if (p0 + pl >= P) continue;
    float* out = W + (long long)(p0 + pl) * 2 * m + 2 * r0;
#pragma unroll
    for (int i = 0; i < TR; ++i)
      if (r0 + i < m) *reinterpret_cast<f32x2*>(out + 2 * i) = acc[i][q];
  }
}

// gW arrives in the gU buffer (the solve kernel's split-mode epilogue) and is overwritten by gU: the panel is staged first
__global__ void __launch_bounds__(PROJ_T) proj_back_kernel(const float* __restrict__ A, const float* __restrict__ Ksym, const float* __restrict__ U,
                                                           long long U_stride, float* __restrict__ gU, long long gU_stride, float* __restrict__ loss,
                                                           float scale, int add_prior, int P, int m) {
  extern __shared__ __align__(16) float ps[];
  float* Gs = ps;
  float* Us = ps + (size_t)m * PROJ_LD;
  float* As = ps + 2 * (size_t)m * PROJ_LD;              // [NST][2][KT][256]: A tile, Ksym tile
  const int p0 = blockIdx.x * PROJ_PP;
  const bool vec = (m & 3) == 0;
  const int ntile = (m + KT - 1) / KT;
  for (int s = 0; s < NST - 1; ++s) {
    if (s < ntile) {
      load_tile(As + (2 * s) * TILE, A, s * KT, m, vec);
      if (add_prior) load_tile(As + (2 * s + 1) * TILE, Ksym, s * KT, m, vec);
    }
    cp_commit();
  }
  load_panel(Gs, gU, gU_stride, p0, P, m);
  if (add_prior) load_panel(Us, U, U_stride, p0, P, m);
  const int r0 = TR * (threadIdx.x & 63), cg = threadIdx.x >> 6;
  const bool rows = r0 < m;
  const float* gc = Gs + 2 * TQ * cg;
  const float* uc = Us + 2 * TQ * cg;
  f32x2 accg[TR][TQ], accp[TR][TQ];
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int q = 0; q < TQ; ++q) accg[i][q] = accp[i][q] = pk(0.f, 0.f);
  // gU[k] = sum_j A[j][k] gW[j] + sum_j Ksym[j][k] U[j]   (Ksym symmetric: row j, columns r0 .. r0+3)
  for (int t = 0; t < ntile; ++t) {
    cp_wait<NST - 2>();
    __syncthreads();
    if (t + NST - 1 < ntile) {
      const int sn = (t + NST - 1) % NST;
      load_tile(As + (2 * sn) * TILE, A, (t + NST - 1) * KT, m, vec);
      if (add_prior) load_tile(As + (2 * sn + 1) * TILE, Ksym, (t + NST - 1) * KT, m, vec);
    }
    cp_commit();
    const float* at = As + (2 * (t % NST)) * TILE + r0;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const int j = t * KT + kk;
      if (j < m) {
        tile_fma(accg, *reinterpret_cast<const float4*>(at + kk * 256), gc + j * PROJ_LD);
        if (add_prior) tile_fma(accp, *reinterpret_cast<const float4*>(at + TILE + kk * 256), uc + j * PROJ_LD);
      }
    }
  }
  __syncthreads();                                // every thread has finished reading the gW panel
  if (rows) {
#pragma unroll
    for (int q = 0; q < TQ; ++q) {
      const int pl = TQ * cg + q;
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const int k = r0 + i;
        if (k >= m) continue;
        float g0, g1, q0, q1;
        upk(accg[i][q], g0, g1);
        upk(accp[i][q], q0, q1);
        if (p0 + pl < P)
          *reinterpret_cast<float2*>(gU + (long long)(p0 + pl) * gU_stride + 2 * k) = make_float2(scale * (g0 + q0), scale * (g1 + q1));
        // prior partials 1/2 U (Ksym U), parked in the (consumed) gW panel in the [k][c] layout
        Gs[k * PROJ_LD + 2 * pl] = add_prior ? 0.5f * Us[k * PROJ_LD + 2 * pl] * q0 : 0.f;
        Gs[k * PROJ_LD + 2 * pl + 1] = add_prior ? 0.5f * Us[k * PROJ_LD + 2 * pl + 1] * q1 : 0.f;
      }
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
  const size_t smem = sizeof(float) * ((size_t)m * PROJ_LD + NST * TILE);
  static bool attr_w = false;
  if (!attr_w) {
    int e = check_cuda(cudaFuncSetAttribute(proj_W_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * (256 * PROJ_LD + NST * TILE))),
                       "proj_W smem attr");
    if (e != BODE_OK) return e;
    attr_w = true;
  }
  proj_W_kernel<<<(P + PROJ_PP - 1) / PROJ_PP, PROJ_T, smem, st>>>(AT, U, U_stride, P, m, W);
  return check_cuda(cudaGetLastError(), "proj_W launch");
}

int launch_proj_back(const float* A, const float* Ksym, const float* U, long long U_stride, float* gU, long long gU_stride, float* loss,
                     float scale, int add_prior, int P, int m, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)m * PROJ_LD * 2 + 2 * NST * TILE);
  static bool attr_set = false;
  if (!attr_set) {
    int e = check_cuda(cudaFuncSetAttribute(proj_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(sizeof(float) * (2 * 256 * PROJ_LD + 2 * NST * TILE))),
                       "proj_back smem attr");
    if (e != BODE_OK) return e;
    attr_set = true;
  }
  proj_back_kernel<<<(P + PROJ_PP - 1) / PROJ_PP, PROJ_T, smem, st>>>(A, Ksym, U, U_stride, gU, gU_stride, loss, scale, add_prior, P, m);
  return check_cuda(cudaGetLastError(), "proj_back launch");
}

}  // namespace bode
