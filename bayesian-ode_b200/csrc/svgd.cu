// SVGD interaction step (samplers/stein.py:12-34 RBFKernel, :75-86 phi), particle-sharded.
//
//   d2_ij  = ||x_i - x_j||^2                               (torch.cdist(X, Y) ** 2, stein.py:22)
//   h      = median(d2 over ALL n*n entries) / (2 ln(n+1)) (np.median, stein.py:25-26; even count -> mean of the
//                                                           two middle order statistics)
//   gamma  = 1 / (1e-8 + 2 h)                              (stein.py:31, sigma = sqrt(h))
//   K_ij   = exp(-gamma d2_ij)
//   phi_i  = (1/n) [ sum_j K_ij s_j + 2 gamma ( (sum_j K_ij) x_i - sum_j K_ij x_j ) ]      
//
// Rows i are the LOCAL particles of this rank, columns j run over all n gathered particles.  The median is an
// exact radix select on the fp32 bit patterns of d2 (order statistic choice is bit-exact and deterministic:
// integer histograms only); with several ranks the per-pass histograms are summed by the caller (one small
// all-reduce per pass) before bode_svgd_select_digit runs, so every rank selects the same element.
#include "common.cuh"
#include "svgd_state.cuh"
#include <map>
#include <string.h>
#include <cooperative_groups.h>

namespace bode {

// tensor-core path (svgd_tc.cu)
int svgd_tc_supported(int d);
int svgd_tc_colmean(const float* X, long long ld, int n, int d, float* mu, SelState* sel, unsigned long long total, cudaStream_t st);
int svgd_tc_gram(const float* Xr, long long ldr, int nr, int row_offset, const float* Xc, long long ldc, int nc, int d, const float* mu,
                 float* D2, unsigned int* maxbits, cudaStream_t st);
int svgd_tc_phi(const float* D2, int nr, int nc, const float* Xc, long long ldx, const float* Gc, long long ldg, int d, const float* mu,
                const float* gam, float gsign, int jsplit, float* part, cudaStream_t st);
int svgd_tc_combine(const float* part, int jsplit, int nr, int d, const float* Xr, long long ldr, const float* mu, const float* gam,
                    float inv_n, float* phi, long long ldp, float* theta, long long ldt, float step, cudaStream_t st);
// pipelined tensor-core path with resident A tiles, bulk-copied operands and the median window (svgd_tc2.cu)
int svgd_tc2_supported(int d, int nc);
int svgd_tc2_d2_tiled(int nr, int nc);
size_t svgd_tc2_carved_bytes(int nr, int nc);
int svgd_tc2_gram(const float* Xr, long long ldr, int nr, int row_offset, const float* Xc, long long ldc, int nc, int d, float* mu,
                  void* ops_base, float* D2, SelState* st, unsigned long long total, int sms, int stages, cudaStream_t stream);
int svgd_tc2_window_select(SelState* st, void* ops_base, int nr, int nc, const PeerInfo& peer, int one_cta, cudaStream_t stream);
static int g_select_ctas = 0;   // bode_svgd_set_select_ctas
unsigned long long* svgd_tc2_table(void* ops_base, int nr, int nc);
int svgd_tc2_phi(const float* D2, int nr, int nc, const float* Xc, long long ldx, const float* Gc, long long ldg, int d, const float* mu,
                 const float* gam, float gsign, void* ops_base, int* jsplit_out, float* part, int sms, int stages, const float* Xr,
                 long long ldr, float inv_n, float* phi, long long ldp, float* theta, long long ldt, float step, cudaStream_t stream);
int svgd_tc2_set_gram_split(int js);
static int g_tensor_cores = 1;

// ---------------------------------------------------------------- squared distances (difference form, fp32)
constexpr int TS = 64;   // tile of 64 x 64 pairs per CTA, 4 x 4 per thread

__global__ void __launch_bounds__(256) sqdist_kernel(const float* __restrict__ Xr, long long ldr, int nr,
                                                     const float* __restrict__ Xc, long long ldc, int nc, int d,
                                                     float* __restrict__ D2, unsigned int* __restrict__ maxbits) {
  extern __shared__ float sm[];
  const int ldk = d | 1;                         // odd row pitch: conflict-free column reads
  float* A = sm;                                 // [TS][ldk] rows
  float* B = sm + TS * ldk;                      // [TS][ldk] cols
  const int r0 = blockIdx.y * TS, c0 = blockIdx.x * TS;
  for (int idx = threadIdx.x; idx < TS * d; idx += blockDim.x) {
    const int i = idx / d, k = idx - i * d;
    A[i * ldk + k] = (r0 + i < nr) ? __ldg(Xr + (long long)(r0 + i) * ldr + k) : 0.f;
    B[i * ldk + k] = (c0 + i < nc) ? __ldg(Xc + (long long)(c0 + i) * ldc + k) : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k = 0; k < d; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = A[(ty + 16 * u) * ldk + k]; b[u] = B[(tx + 16 * u) * ldk + k]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) { const float t = a[u] - b[v]; acc[u][v] = fmaf(t, t, acc[u][v]); }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int r = r0 + ty + 16 * u, c = c0 + tx + 16 * v;
      if (r < nr && c < nc) D2[(long long)r * nc + c] = acc[u][v];
    }
  // largest d2 of the job (non-negative floats order like their bit patterns): positions the hot window of pass 0
  float mx = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) mx = fmaxf(mx, acc[u][v]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(maxbits, __float_as_uint(mx));
}

// ---------------------------------------------------------------- exact median by radix select
__global__ void select_init_kernel(SelState* st, unsigned long long* hist, unsigned long long total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->prefix[0] = st->prefix[1] = 0u;
    st->maxbits = 0u;
    st->hit = 0u;
    st->rank[0] = (total - 1) / 2;             // lower middle (0-based, ascending)
    st->rank[1] = total / 2;                   // upper middle; equal to rank[0] when total is odd
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * 2048; i += gridDim.x * blockDim.x) hist[i] = 0ull;
}

// Pass 0 (top 11 bits, every element participates): d2 values cluster in a few dozen exponent/mantissa buckets just
// below the maximum, so plain shared atomics serialise.  The HOT buckets [top-HOTB+1, top] get one counter PER LANE
// (column = lane id -> bank = lane: conflict-free within a warp); anything below the window takes the ordinary path.
constexpr int HOTB = 64;

__global__ void __launch_bounds__(512) hist0_kernel(const float* __restrict__ D2, long long n, const SelState* __restrict__ st,
                                                    unsigned long long* __restrict__ hist) {
  __shared__ unsigned int hot[HOTB][32];
  __shared__ unsigned int cold[2048];
  if (st->hit) return;                              // the window already resolved the median (svgd_state.cuh)
  for (int i = threadIdx.x; i < HOTB * 32; i += blockDim.x) (&hot[0][0])[i] = 0u;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) cold[i] = 0u;
  __syncthreads();
  const int top = (int)(st->maxbits >> 20), base = top - (HOTB - 1);
  const int lane = threadIdx.x & 31;
  const long long n4 = n >> 2;
  const uint4* D4 = reinterpret_cast<const uint4*>(D2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 q = __ldg(D4 + i);
    const unsigned int v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int dig = (int)(v[u] >> 20), h = dig - base;
      if (h >= 0 && h < HOTB) atomicAdd(&hot[h][lane], 1u);
      else atomicAdd(&cold[dig & 2047], 1u);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - (n4 << 2))) {          // scalar tail
    const int dig = (int)(__float_as_uint(__ldg(D2 + (n4 << 2) + threadIdx.x)) >> 20);
    atomicAdd(&cold[dig & 2047], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    unsigned int c = cold[i];
    const int h = i - base;
    if (h >= 0 && h < HOTB) {
#pragma unroll 8
      for (int l = 0; l < 32; ++l) c += hot[h][l];
    }
    if (c) atomicAdd(hist + i, (unsigned long long)c);
  }
}

// Passes 1, 2 over bits [shift, shift+nbits): histogram of the (few) elements whose higher bits equal a running prefix
__global__ void __launch_bounds__(512) hist_kernel(const float* __restrict__ D2, long long n, const SelState* __restrict__ st,
                                                   int shift, int nbits, unsigned int himask, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[2][2048];
  if (st->hit) return;
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) (&sh[0][0])[i] = 0u;
  __syncthreads();
  const unsigned int pa = st->prefix[0], pb = st->prefix[1];
  const bool two = pa != pb;
  const unsigned int dmask = (1u << nbits) - 1u;
  const long long n4 = n >> 2;
  const uint4* D4 = reinterpret_cast<const uint4*>(D2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 q = __ldg(D4 + i);
    const unsigned int v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned int hi = v[u] & himask, dig = (v[u] >> shift) & dmask;
      if (hi == pa) atomicAdd(&sh[0][dig], 1u);
      else if (two && hi == pb) atomicAdd(&sh[1][dig], 1u);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - (n4 << 2))) {          // scalar tail
    const unsigned int v = __float_as_uint(__ldg(D2 + (n4 << 2) + threadIdx.x));
    const unsigned int hi = v & himask, dig = (v >> shift) & dmask;
    if (hi == pa) atomicAdd(&sh[0][dig], 1u);
    else if (two && hi == pb) atomicAdd(&sh[1][dig], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) {
    const unsigned int c = (&sh[0][0])[i];
    if (c) atomicAdd(hist + i, (unsigned long long)c);
  }
}

// one CTA of 1024 threads: pick the digit holding each remaining rank (block-wide exclusive scan of the 2048 counts),
// extend the prefixes, clear the histograms
__global__ void __launch_bounds__(1024) select_digit_kernel(SelState* st, unsigned long long* hist, int shift, int nbits) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned int newp[2];
  __shared__ unsigned long long newr[2];
  if (st->hit) return;
  const int nb = 1 << nbits;
  const bool two = st->prefix[0] != st->prefix[1];
  const unsigned long long rank0 = st->rank[0], rank1 = st->rank[1];
  const unsigned int pre0 = st->prefix[0], pre1 = st->prefix[1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int which = 0; which < 2; ++which) {
    const unsigned long long* h = hist + ((which == 1 && two) ? 2048 : 0);
    const unsigned long long r = which ? rank1 : rank0;
    const int i0 = 2 * threadIdx.x;
    const unsigned long long c0 = i0 < nb ? h[i0] : 0ull, c1 = i0 + 1 < nb ? h[i0 + 1] : 0ull;
    unsigned long long incl = c0 + c1;                          // inclusive scan over threads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      unsigned long long w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      wsum[lane] = w;                                           // inclusive over warps
    }
    __syncthreads();
    const unsigned long long excl = incl - (c0 + c1) + (wid ? wsum[wid - 1] : 0ull);   // count below bin i0
    if (r >= excl && r < excl + c0) { newp[which] = (which ? pre1 : pre0) | ((unsigned)i0 << shift); newr[which] = r - excl; }
    else if (r >= excl + c0 && r < excl + c0 + c1) { newp[which] = (which ? pre1 : pre0) | ((unsigned)(i0 + 1) << shift); newr[which] = r - excl - c0; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    st->prefix[0] = newp[0]; st->prefix[1] = newp[1];
    st->rank[0] = newr[0]; st->rank[1] = newr[1];
  }
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) hist[i] = 0ull;
}

__device__ __forceinline__ void gamma_dev(SelState* st, int n, float sigma_fixed, int arm_window, float* out);

// Single-rank fallback in ONE cooperative launch: when the median window missed (first call, large step) the three radix
// passes run back to back with grid-wide barriers; when it hit, every CTA returns at once, so a captured CUDA graph pays one
// near-empty launch instead of seven.  1024 threads per CTA (the select scan needs them), grid = co-resident CTAs.
__device__ __forceinline__ void hist_pass_dev(const float* __restrict__ D2, long long n, unsigned int pa, unsigned int pb, int shift, int nbits,
                                              unsigned int himask, unsigned int* sh /*[2][2048]*/, unsigned long long* __restrict__ hist) {
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const bool two = pa != pb;
  const unsigned int dmask = (1u << nbits) - 1u;
  const long long n4 = n >> 2;
  const uint4* D4 = reinterpret_cast<const uint4*>(D2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 q = __ldg(D4 + i);
    const unsigned int v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned int hi = v[u] & himask, dig = (v[u] >> shift) & dmask;
      if (hi == pa) atomicAdd(&sh[dig], 1u);
      else if (two && hi == pb) atomicAdd(&sh[2048 + dig], 1u);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - (n4 << 2))) {          // scalar tail
    const unsigned int v = __float_as_uint(__ldg(D2 + (n4 << 2) + threadIdx.x));
    const unsigned int hi = v & himask, dig = (v >> shift) & dmask;
    if (hi == pa) atomicAdd(&sh[dig], 1u);
    else if (two && hi == pb) atomicAdd(&sh[2048 + dig], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) {
    const unsigned int c = sh[i];
    if (c) atomicAdd(hist + i, (unsigned long long)c);
  }
}

__device__ __forceinline__ void select_digit_dev(SelState* st, unsigned long long* hist, int shift, int nbits, unsigned long long* wsum /*[32]*/,
                                                 unsigned int* newp /*[2]*/, unsigned long long* newr /*[2]*/) {
  const int nb = 1 << nbits;
  const bool two = st->prefix[0] != st->prefix[1];
  const unsigned long long rank0 = st->rank[0], rank1 = st->rank[1];
  const unsigned int pre0 = st->prefix[0], pre1 = st->prefix[1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int which = 0; which < 2; ++which) {
    const unsigned long long* h = hist + ((which == 1 && two) ? 2048 : 0);
    const unsigned long long r = which ? rank1 : rank0;
    const int i0 = 2 * threadIdx.x;
    const unsigned long long c0 = i0 < nb ? h[i0] : 0ull, c1 = i0 + 1 < nb ? h[i0 + 1] : 0ull;
    unsigned long long incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      unsigned long long w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const unsigned long long excl = incl - (c0 + c1) + (wid ? wsum[wid - 1] : 0ull);
    if (r >= excl && r < excl + c0) { newp[which] = (which ? pre1 : pre0) | ((unsigned)i0 << shift); newr[which] = r - excl; }
    else if (r >= excl + c0 && r < excl + c0 + c1) { newp[which] = (which ? pre1 : pre0) | ((unsigned)(i0 + 1) << shift); newr[which] = r - excl - c0; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    st->prefix[0] = newp[0]; st->prefix[1] = newp[1];
    st->rank[0] = newr[0]; st->rank[1] = newr[1];
  }
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) hist[i] = 0ull;
}

// Several ranks: `hist` holds this rank's counts; after every histogram pass block 0 meets the peers at a flag barrier, sums all
// ranks' histograms (read over NVLink) into `hsum` and selects on the sums -- every rank extends the same prefix.
__global__ void __launch_bounds__(1024) radix_fallback_kernel(const float* __restrict__ D2, long long n, SelState* st, unsigned long long* hist,
                                                              unsigned long long* hsum, int n_total, int arm_window, float* med_gamma,
                                                              const PeerInfo peer) {
  __shared__ unsigned int sh[2 * 2048];
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned int newp[2];
  __shared__ unsigned long long newr[2];
  if (st->hit) {                                     // uniform over the grid: written before this launch
    if (med_gamma && blockIdx.x == 0 && threadIdx.x == 0) gamma_dev(st, n_total, 0.f, arm_window, med_gamma);
    return;
  }
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int sh_[3] = {20, 9, 0}, nb_[3] = {11, 11, 9};
  for (int pass = 0; pass < 3; ++pass) {
    const unsigned int himask = pass == 0 ? 0u : (0xffffffffu << (sh_[pass] + nb_[pass]));
    // pass 0: himask = 0 and both prefixes are 0, so every element lands in sh[0][digit]
    hist_pass_dev(D2, n, st->prefix[0], st->prefix[1], sh_[pass], nb_[pass], himask, sh, hist);
    __threadfence();
    grid.sync();
    if (blockIdx.x == 0) {
      if (peer.world > 1) {
        if (threadIdx.x == 0) peer_barrier(peer);                   // every rank's histogram of this pass is complete
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) {
          unsigned long long acc = 0ull;
          for (int q = 0; q < peer.world; ++q)
            acc += peer_ld_u64(reinterpret_cast<const unsigned long long*>(peer.base[q] + peer.hist_off) + i);
          hsum[i] = acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) peer_barrier(peer);                   // ... and has been read by everyone: clear mine
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) hist[i] = 0ull;
        select_digit_dev(st, hsum, sh_[pass], nb_[pass], wsum, newp, newr);
      } else {
        select_digit_dev(st, hist, sh_[pass], nb_[pass], wsum, newp, newr);
      }
    }
    __threadfence();
    grid.sync();
  }
  if (med_gamma && blockIdx.x == 0 && threadIdx.x == 0) gamma_dev(st, n_total, 0.f, arm_window, med_gamma);
}

// out[0] = median, out[1] = gamma   (stein.py:25-31)
__device__ __forceinline__ void gamma_dev(SelState* st, int n, float sigma_fixed, int arm_window, float* out) {
  const float a = __uint_as_float(st->prefix[0]), b = __uint_as_float(st->prefix[1]);
  // arm the window of the next call around the lower middle element just selected (svgd_state.cuh)
  if (arm_window) {
    st->win_lo = st->prefix[0] > WIN_HALF ? st->prefix[0] - WIN_HALF : 0u;
    st->win_valid = 1u;
  }
  const float med = 0.5f * (a + b);
  double s2;
  if (sigma_fixed > 0.f) s2 = (double)sigma_fixed * (double)sigma_fixed;
  else s2 = (double)med / (2.0 * log((double)n + 1.0));
  out[0] = med;
  out[1] = (float)(1.0 / (1e-8 + 2.0 * s2));
  st->maxbits = 0u;   // end of this selection: the next operand pass raises it again (pipelined path has no separate reset launch)
}
__global__ void gamma_kernel(SelState* st, int n, float sigma_fixed, int arm_window, float* out) { gamma_dev(st, n, sigma_fixed, arm_window, out); }

// ---------------------------------------------------------------- phi partials: K[rows, jslice] @ [S | X | 1]
// CTA = 64 rows x F features (F = 16*FT >= 2d+1; the trailing ones column yields the row sums), 128 threads, each an
// 8 x FT register tile; per j the thread issues 2 LDS.128 (8 kernel values) + FT LDS for 8*FT FFMA.
constexpr int PR = 64;      // rows per CTA
constexpr int PJ = 32;      // columns per smem stage
constexpr int KP = PR + 4;  // pitch of the transposed kernel tile (16-byte aligned rows)
constexpr int MAXSPLIT = 16;

template <int FT>
__global__ void __launch_bounds__(128, 3) phi_partial_kernel(const float* __restrict__ D2, int nr, int nc,
                                                          const float* __restrict__ Xc, long long ldx,
                                                          const float* __restrict__ Sc, long long lds, int d,
                                                          const float* __restrict__ gam, int jsplit, float ssign,
                                                          float* __restrict__ part) {
  extern __shared__ __align__(16) float sm[];
  constexpr int F = 16 * FT;
  float* Ks = sm;                                // [PJ][KP]   K^T tile
  float* Vs = sm + PJ * KP;                      // [PJ][F]    [sign*G | X | 1 | 0..]
  const float ngamma = -gam[1] * 1.4426950408889634f;
  const int r0 = blockIdx.x * PR;
  const int jper = ((nc + jsplit - 1) / jsplit + PJ - 1) / PJ * PJ;
  const int jbeg = blockIdx.y * jper, jend = min(nc, jbeg + jper);
  const int fg = threadIdx.x & 15, rg = threadIdx.x >> 4;
  float acc[8][FT] = {};
  for (int j0 = jbeg; j0 < jend; j0 += PJ) {
    // all global loads of the stage are issued before the first use (fixed trip counts, fully unrolled)
    float kreg[PR * PJ / 128], vreg[4 * FT];
#pragma unroll
    for (int q = 0; q < PR * PJ / 128; ++q) {
      const int idx = threadIdx.x + 128 * q, i = idx / PJ, j = idx - i * PJ;
      kreg[q] = (r0 + i < nr && j0 + j < jend) ? __ldg(D2 + (long long)(r0 + i) * nc + j0 + j) : INFINITY;
    }
#pragma unroll
    for (int q = 0; q < 4 * FT; ++q) {
      const int idx = threadIdx.x + 128 * q, j = idx / F, c = idx - j * F;
      float v = 0.f;
      if (j0 + j < jend) {
        if (c < d) v = ssign * __ldg(Sc + (long long)(j0 + j) * lds + c);
        else if (c < 2 * d) v = __ldg(Xc + (long long)(j0 + j) * ldx + (c - d));
        else if (c == 2 * d) v = 1.f;
      }
      vreg[q] = v;
    }
#pragma unroll
    for (int q = 0; q < PR * PJ / 128; ++q) {
      const int idx = threadIdx.x + 128 * q, i = idx / PJ, j = idx - i * PJ;
      Ks[j * KP + i] = ex2(ngamma * kreg[q]);                 // 2^(-inf) = 0 for out-of-range pairs
    }
#pragma unroll
    for (int q = 0; q < 4 * FT; ++q) Vs[threadIdx.x + 128 * q] = vreg[q];
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < PJ; ++j) {
      const float4 ka = *reinterpret_cast<const float4*>(Ks + j * KP + 8 * rg);
      const float4 kb = *reinterpret_cast<const float4*>(Ks + j * KP + 8 * rg + 4);
      const float k[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
      float v[FT];
#pragma unroll
      for (int c = 0; c < FT; ++c) v[c] = Vs[j * F + fg * FT + c];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < FT; ++c) acc[u][c] = fmaf(k[u], v[c], acc[u][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int r = r0 + 8 * rg + u;
    if (r >= nr) continue;
    float* dst = part + ((long long)blockIdx.y * nr + r) * (2 * d + 1);
#pragma unroll
    for (int c = 0; c < FT; ++c) {
      const int f = fg * FT + c;
      if (f <= 2 * d) dst[f] = acc[u][c];
    }
  }
}

// phi = (KS + 2 gamma (rowsum * x_i - KX)) / n ; optional fused update theta_i += step * phi_i
__global__ void phi_combine_kernel(const float* __restrict__ part, int jsplit, int nr, int d,
                                   const float* __restrict__ Xr, long long ldr, const float* __restrict__ gam, float inv_n,
                                   float* __restrict__ phi, long long ldp, float* __restrict__ theta, long long ldt, float step) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nr * d) return;
  const int r = (int)(idx / d), c = (int)(idx - (long long)r * d);
  float ks = 0.f, kx = 0.f, rs = 0.f;
  for (int s = 0; s < jsplit; ++s) {
    const float* p = part + ((long long)s * nr + r) * (2 * d + 1);
    ks += p[c];
    kx += p[d + c];
    rs += p[2 * d];
  }
  const float x = Xr[(long long)r * ldr + c];
  const float ph = (ks + 2.f * gam[1] * (rs * x - kx)) * inv_n;
  if (phi) phi[(long long)r * ldp + c] = ph;
  if (theta) theta[(long long)r * ldt + c] = fmaf(step, ph, theta[(long long)r * ldt + c]);
}

}  // namespace bode

using namespace bode;

extern "C" size_t bode_svgd_workspace_bytes(int32_t n_rows, int32_t n_cols, int32_t d) {
  const size_t jsplit = MAXSPLIT;
  size_t b = 0;
  b += (size_t)n_rows * n_cols * sizeof(float);                 // d2
  b += jsplit * n_rows * (2 * (size_t)d + 1) * sizeof(float);   // phi partials (last column: row sums)
  b += 2 * 2048 * sizeof(unsigned long long) + 256;             // histograms + select state
  b += 256;                                                     // column means (tensor-core path)
  b += 256;                                                     // peer barrier flags (svgd_state.cuh)
  b += svgd_tc2_carved_bytes(n_rows, n_cols);                   // pre-split operands + median window table
  b += 2 * ((((size_t)n_cols * d * sizeof(float)) + 255) / 256 * 256) + 256;   // gathered positions / scores (bode_svgd_peer_gather) + tickets
  return b + 1024;
}

namespace {
struct Ws {
  float* d2; float* part; unsigned long long* hist; SelState* st; float* mu; PeerFlags* flags; void* ops;
  float* gath[2]; unsigned int* tickets;
};
Ws carve(void* ws, int nr, int nc, int d) {
  Ws w;
  char* p = (char*)ws;
  w.d2 = (float*)p; p += (((size_t)nr * nc * sizeof(float)) + 255) / 256 * 256;
  w.part = (float*)p; p += ((MAXSPLIT * (size_t)nr * (2 * d + 1) * sizeof(float)) + 255) / 256 * 256;
  w.hist = (unsigned long long*)p; p += 2 * 2048 * sizeof(unsigned long long);
  w.st = (SelState*)p; p += 256;
  w.mu = (float*)p; p += 256;
  w.flags = (PeerFlags*)p; p += 256;
  w.ops = p;
  p += svgd_tc2_carved_bytes(nr, nc);
  p = (char*)(((uintptr_t)p + 255) / 256 * 256);
  const size_t gb = (((size_t)nc * d * sizeof(float)) + 255) / 256 * 256;
  w.gath[0] = (float*)p; p += gb;
  w.gath[1] = (float*)p; p += gb;
  w.tickets = (unsigned int*)p;
  return w;
}
// workspace -> peer mapping (bode_svgd_set_peers); absent = single rank
std::map<void*, PeerInfo> g_peers;
PeerInfo peers_of(void* workspace) {
  auto it = g_peers.find(workspace);
  if (it != g_peers.end()) return it->second;
  PeerInfo p;
  memset(&p, 0, sizeof(p));
  return p;
}
}  // namespace

/* d2[rows, cols] for the local rows, and reset of the select state; total = number of entries the median runs over
 * (n*n for the whole job).  hist_out receives the device address of the 2x2048 uint64 histogram block so a multi-rank
 * caller can all-reduce it between bode_svgd_hist_pass and bode_svgd_select_digit. */
/* Peer-mapped workspaces (several ranks on one node): bases[q] = address, in THIS process, of rank q's workspace (bases[rank] ==
 * workspace).  From then on bode_svgd_window_select and bode_svgd_radix_fallback on this workspace read the peers' window tables /
 * histograms directly and synchronise through flag barriers, so the exact distributed median needs no collective launches.
 * world <= 1 removes the mapping. */
extern "C" int bode_svgd_set_peers(void* workspace, int32_t n_rows, int32_t n_cols, int32_t d, void* const* bases, int32_t rank, int32_t world) {
  BODE_REQUIRE(workspace, "null workspace");
  if (world <= 1) {
    g_peers.erase(workspace);
    return BODE_OK;
  }
  BODE_REQUIRE(world <= MAX_PEERS && rank >= 0 && rank < world && bases, "bad peer description (at most %d ranks)", MAX_PEERS);
  BODE_REQUIRE(bases[rank] == workspace, "bases[rank] must be this rank's own workspace");
  Ws w = carve(workspace, n_rows, n_cols, d);
  PeerInfo p;
  memset(&p, 0, sizeof(p));
  for (int q = 0; q < world; ++q) {
    BODE_REQUIRE(bases[q], "null peer base");
    p.base[q] = (unsigned char*)bases[q];
  }
  p.rank = rank;
  p.world = world;
  p.hist_off = (unsigned long long)((char*)w.hist - (char*)workspace);
  p.table_off = (unsigned long long)((char*)svgd_tc2_table(w.ops, n_rows, n_cols) - (char*)workspace);
  p.flag_off = (unsigned long long)((char*)w.flags - (char*)workspace);
  g_peers[workspace] = p;
  return BODE_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// All-gather over peer memory (the SVGD data-path exchange of SURVEY.md 8(e), without a collective launch).  Every rank PUSHES
// its rows into the gather buffer of every rank's workspace (posted NVLink writes) between two flag barriers:
//   entry  every rank has reached this gather, i.e. finished the kernels of the previous step that read the buffer;
//   exit   every rank's rows have landed everywhere (release / acquire at system scope).
// The barriers use their own flag block per buffer (block 0 belongs to the median kernels, which run on another stream).  The
// first CTA to start signals the entry, the last CTA to finish pushing signals the exit and waits: the kernel does not retire
// before the gathered buffer is complete, so the consumers simply follow it in stream order.
__device__ __forceinline__ void peer_signal(const PeerInfo& p, int fb, unsigned int e) {
  __threadfence_system();
  for (int q = 0; q < p.world; ++q) {
    unsigned int* f = &reinterpret_cast<PeerFlags*>(p.base[q] + p.flag_off + 64 * fb)->arrived[p.rank];
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(e) : "memory");
  }
}
__device__ __forceinline__ void peer_wait(const PeerInfo& p, PeerFlags* mine, unsigned int e) {
  for (int q = 0; q < p.world; ++q) {
    unsigned int v;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(&mine->arrived[q]) : "memory");
    } while ((int)(v - e) < 0 && ++spins < (1ll << 24));
    if ((int)(v - e) < 0) mine->timed_out = 1u;
  }
}

__global__ void __launch_bounds__(256) peer_gather_kernel(const float* __restrict__ src, long long ld, int n_rows, int d, const PeerInfo peer,
                                                          unsigned long long gath_off, int fb, unsigned int* tickets) {
  PeerFlags* mine = reinterpret_cast<PeerFlags*>(peer.base[peer.rank] + peer.flag_off + 64 * fb);
  __shared__ unsigned int e_s;
  if (threadIdx.x == 0) {
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&mine->epoch) + 1u;
    if (atomicAdd(tickets + 2 * fb, 1u) == 0u) peer_signal(peer, fb, e);
    peer_wait(peer, mine, e);
    e_s = e;
  }
  __syncthreads();
  const long long n = (long long)n_rows * d;
  const unsigned long long dst_off = gath_off + (unsigned long long)peer.rank * n * sizeof(float);
  if (ld == d && (n & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n >> 2); i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(s4 + i);
      for (int q = 0; q < peer.world; ++q) reinterpret_cast<float4*>(peer.base[q] + dst_off)[i] = v;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const float v = __ldg(src + (i / d) * ld + (i % d));
      for (int q = 0; q < peer.world; ++q) reinterpret_cast<float*>(peer.base[q] + dst_off)[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(tickets + 2 * fb + 1, 1u) == gridDim.x - 1) {            // every CTA of this rank has pushed (and fenced) its part
      const unsigned int e = e_s + 1u;
      tickets[2 * fb] = 0u;
      tickets[2 * fb + 1] = 0u;
      mine->epoch = e;
      peer_signal(peer, fb, e);
      peer_wait(peer, mine, e);
    }
  }
}

/* which = 0: particle positions, 1: scores.  rows[n_rows, d] (leading dimension ld) of this rank -> *gathered_out = the
 * [n_cols, d] buffer (contiguous, rank-major) inside this rank's workspace, complete when the launch retires.  Needs
 * bode_svgd_set_peers and n_cols == world * n_rows; every rank must issue the same sequence of gathers per buffer. */
extern "C" int bode_svgd_peer_gather(int32_t which, const float* rows, int64_t ld, int32_t n_rows, int32_t n_cols, int32_t d, void* workspace,
                                     float** gathered_out, bode_stream_t stream) {
  BODE_REQUIRE(workspace && rows && gathered_out && (which == 0 || which == 1), "bad args");
  const PeerInfo peer = peers_of(workspace);
  BODE_REQUIRE(peer.world > 1, "workspace has no peers (bode_svgd_set_peers)");
  BODE_REQUIRE((long long)peer.world * n_rows == n_cols, "n_cols must be world * n_rows");
  Ws w = carve(workspace, n_rows, n_cols, d);
  const unsigned long long off = (unsigned long long)((char*)w.gath[which] - (char*)workspace);
  const long long n4 = ((long long)n_rows * d + 3) / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 32) blocks = 32;                      // all CTAs poll the entry flags: keep them few and co-resident
  if (blocks < 1) blocks = 1;
  peer_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rows, ld, n_rows, d, peer, off, 1 + which, w.tickets);
  BODE_CUDA(cudaGetLastError());
  *gathered_out = w.gath[which];
  return BODE_OK;
}

/* Host-side check of the flag barriers (median kernels, position gather, score gather): *timed_out = 1 when one of them gave up
 * waiting for a peer (PeerFlags::timed_out, svgd_state.cuh) since bode_svgd_workspace_init -- the results of that step are then
 * invalid.  Blocking copy: call it after synchronising the streams that use the workspace. */
extern "C" int bode_svgd_peer_status(void* workspace, int32_t n_rows, int32_t n_cols, int32_t d, int32_t* timed_out) {
  BODE_REQUIRE(workspace && timed_out, "null pointer");
  Ws w = carve(workspace, n_rows, n_cols, d);
  unsigned int h[64];
  static_assert(sizeof(PeerFlags) <= 64 && sizeof(h) == 256, "three 64-byte flag blocks inside the 256-byte flag region");
  BODE_CUDA(cudaMemcpy(h, w.flags, sizeof(h), cudaMemcpyDeviceToHost));
  int t = 0;
  for (int k = 0; k < 3; ++k) {
    PeerFlags f;
    memcpy(&f, reinterpret_cast<const char*>(h) + 64 * k, sizeof(f));
    t |= f.timed_out != 0u;
  }
  *timed_out = t;
  return BODE_OK;
}

extern "C" int bode_svgd_set_gram_split(int32_t column_splits) { return svgd_tc2_set_gram_split(column_splits); }

extern "C" int bode_svgd_staged_supported(int32_t n_cols, int32_t d) {
  return (g_tensor_cores && svgd_tc_supported(d) && svgd_tc2_supported(d, n_cols)) ? 1 : 0;
}

extern "C" int bode_svgd_d2_tiled(int32_t n_rows, int32_t n_cols, int32_t d) {
  return (bode_svgd_staged_supported(n_cols, d) && svgd_tc2_d2_tiled(n_rows, n_cols)) ? 1 : 0;
}

/* stages: BODE_SVGD_PREPARE (column means, selection-state reset, pre-split operands: needs only the positions) and / or
 * BODE_SVGD_COMPUTE (the Gram kernel).  On shapes bode_svgd_staged_supported rejects, PREPARE is empty and COMPUTE does it all. */
extern "C" int bode_svgd_sqdist_staged(int32_t stages, const float* Xrows, int64_t ld_rows, int32_t n_rows, const float* Xcols,
                                       int64_t ld_cols, int32_t n_cols, int32_t d, int32_t row_offset, uint64_t total_entries,
                                       void* workspace, size_t workspace_bytes, void** hist_out, bode_stream_t stream) {
  BODE_REQUIRE(Xrows && Xcols && workspace, "null pointer");
  BODE_REQUIRE((stages & ~3) == 0 && stages != 0, "stages must be a combination of BODE_SVGD_PREPARE | BODE_SVGD_COMPUTE");
  if (!bode_svgd_staged_supported(n_cols, d)) {
    if (!(stages & BODE_SVGD_COMPUTE)) return BODE_OK;
    stages = BODE_SVGD_PREPARE | BODE_SVGD_COMPUTE;
  }
  BODE_REQUIRE(n_rows > 0 && n_cols > 0 && d > 0 && d <= 512, "bad sizes");
  BODE_REQUIRE(workspace_bytes >= bode_svgd_workspace_bytes(n_rows, n_cols, d), "workspace too small");
  BODE_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  // unrelated row and column sets: no entry is a cdist(x, x) diagonal.  The kernels force d2 = 0 where row + row_offset == column,
  // which for the documented -1 would have hit the first sub-diagonal; move the "diagonal" out of every matrix instead.
  if (row_offset < 0) row_offset = -(1 << 30);
  Ws w = carve(workspace, n_rows, n_cols, d);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = 2 * (size_t)TS * (d | 1) * sizeof(float);
  if (smem > 48 * 1024) BODE_CUDA(cudaFuncSetAttribute(sqdist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((n_cols + TS - 1) / TS, (n_rows + TS - 1) / TS);
  if (g_tensor_cores && svgd_tc_supported(d)) {
    // 3xTF32 Gram on tcgen05: d2 = |xc_i|^2 + |xc_j|^2 - 2 xc_i.xc_j with xc centred on a reference point inside the cloud.
    int e = BODE_OK;
    if (svgd_tc2_supported(d, n_cols)) {
      // pipelined path: the operand kernel forms the centre itself and starts the selection (no column-mean launch)
      const int sms = bode_device_sm_count();
      if (sms < 0) return BODE_ERR_CUDA;
      e = svgd_tc2_gram(Xrows, ld_rows, n_rows, row_offset, Xcols, ld_cols, n_cols, d, w.mu, w.ops, w.d2, w.st, total_entries, sms, stages, st);
      if (e != BODE_OK) return e;
      if (hist_out) *hist_out = w.hist;
      return BODE_OK;
    }
    // centred on the mean of ALL particles; the column-mean kernel also resets the selection state
    e = svgd_tc_colmean(Xcols, ld_cols, n_cols, d, w.mu, w.st, total_entries, st);
    if (e != BODE_OK) return e;
    e = svgd_tc_gram(Xrows, ld_rows, n_rows, row_offset, Xcols, ld_cols, n_cols, d, w.mu, w.d2, &w.st->maxbits, st);
    if (e != BODE_OK) return e;
  } else {
    select_init_kernel<<<4, 1024, 0, st>>>(w.st, w.hist, total_entries);
    BODE_CUDA(cudaGetLastError());
    sqdist_kernel<<<grid, 256, smem, st>>>(Xrows, ld_rows, n_rows, Xcols, ld_cols, n_cols, d, w.d2, &w.st->maxbits);
    BODE_CUDA(cudaGetLastError());
  }
  if (hist_out) *hist_out = w.hist;
  return BODE_OK;
}

extern "C" int bode_svgd_sqdist(const float* Xrows, int64_t ld_rows, int32_t n_rows, const float* Xcols, int64_t ld_cols,
                                int32_t n_cols, int32_t d, int32_t row_offset, uint64_t total_entries, void* workspace,
                                size_t workspace_bytes, void** hist_out, bode_stream_t stream) {
  return bode_svgd_sqdist_staged(BODE_SVGD_PREPARE | BODE_SVGD_COMPUTE, Xrows, ld_rows, n_rows, Xcols, ld_cols, n_cols, d, row_offset,
                                 total_entries, workspace, workspace_bytes, hist_out, stream);
}

/* Zero the persistent selection state (median window disarmed, table cleared).  Call once after allocating a workspace. */
extern "C" int bode_svgd_workspace_init(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, size_t workspace_bytes, bode_stream_t stream) {
  BODE_REQUIRE(workspace && n_rows > 0 && n_cols > 0 && d > 0, "bad args");
  BODE_REQUIRE(workspace_bytes >= bode_svgd_workspace_bytes(n_rows, n_cols, d), "workspace too small");
  Ws w = carve(workspace, n_rows, n_cols, d);
  BODE_CUDA(cudaMemsetAsync(w.hist, 0, 2 * 2048 * sizeof(unsigned long long), (cudaStream_t)stream));
  BODE_CUDA(cudaMemsetAsync(w.st, 0, 256, (cudaStream_t)stream));
  BODE_CUDA(cudaMemsetAsync(w.flags, 0, 256, (cudaStream_t)stream));
  BODE_CUDA(cudaMemsetAsync(w.tickets, 0, 256, (cudaStream_t)stream));
  BODE_CUDA(cudaMemsetAsync(svgd_tc2_table(w.ops, n_rows, n_cols), 0, 2 * (size_t)(WIN_TABLE + 1) * sizeof(unsigned long long) + 512, (cudaStream_t)stream));
  return BODE_OK;
}

/* Disarm the median window: the next selection takes the radix passes (what the first call of a run, or a step that moves the
 * median by more than the window's +-0.2 %, pays).  Used to MEASURE that path (bench.py median_fallback_ms) and by callers that
 * replace the particles wholesale.  Every rank of a peer-mapped job must call it (all ranks must take the same path). */
extern "C" int bode_svgd_window_disarm(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, bode_stream_t stream) {
  BODE_REQUIRE(workspace && n_rows > 0 && n_cols > 0 && d > 0, "bad args");
  Ws w = carve(workspace, n_rows, n_cols, d);
  BODE_CUDA(cudaMemsetAsync(&w.st->win_valid, 0, sizeof(unsigned int), (cudaStream_t)stream));
  return BODE_OK;
}

/* Median window (svgd_state.cuh): *table_out = device address of WIN_TABLE+1 uint64 counters filled by bode_svgd_sqdist
 * (multi-rank callers all-reduce them); bode_svgd_window_select then reads the median off the table when both middle
 * ranks fall inside the window, which turns the following bode_svgd_hist_pass / bode_svgd_select_digit calls into no-ops. */
extern "C" int bode_svgd_window_table(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, void** table_out, size_t* count_out) {
  BODE_REQUIRE(workspace && table_out && count_out, "null pointer");
  Ws w = carve(workspace, n_rows, n_cols, d);
  *table_out = svgd_tc2_table(w.ops, n_rows, n_cols);
  *count_out = (size_t)WIN_TABLE + 1;
  return BODE_OK;
}

extern "C" int bode_svgd_window_select(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, bode_stream_t stream) {
  BODE_REQUIRE(workspace, "null workspace");
  Ws w = carve(workspace, n_rows, n_cols, d);
  return svgd_tc2_window_select(w.st, w.ops, n_rows, n_cols, peers_of(workspace), g_select_ctas > 0 ? 1 : 0, (cudaStream_t)stream);
}

/* 1 (default): Gram and K@[S|X|1] on tcgen05 tensor cores (3xTF32) when d <= 56; 0: FP32-pipe kernels */
extern "C" int bode_svgd_set_tensor_cores(int32_t on) {
  const int old = g_tensor_cores;
  g_tensor_cores = on ? 1 : 0;
  return old;
}

/* pass = 0,1,2 over bit windows [20,31), [9,20), [0,9) of the fp32 pattern (d2 >= 0 so bit 31 is clear) */
static void window(int pass, int* shift, int* nbits, unsigned int* himask) {
  const int sh[3] = {20, 9, 0}, nb[3] = {11, 11, 9};
  *shift = sh[pass]; *nbits = nb[pass];
  *himask = pass == 0 ? 0u : (0xffffffffu << (sh[pass] + nb[pass]));
}

extern "C" int bode_svgd_hist_pass(int32_t pass, int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, bode_stream_t stream) {
  BODE_REQUIRE(pass >= 0 && pass < 3 && workspace, "bad pass/workspace");
  Ws w = carve(workspace, n_rows, n_cols, d);
  int shift, nbits; unsigned int himask;
  window(pass, &shift, &nbits, &himask);
  const long long n = (long long)n_rows * n_cols;
  int sms = bode_device_sm_count();
  if (sms < 0) return BODE_ERR_CUDA;
  long long blocks = (n / 4 + 511) / 512;
  if (blocks > 4LL * sms) blocks = 4LL * sms;
  if (blocks < 1) blocks = 1;
  if (pass == 0) hist0_kernel<<<(int)blocks, 512, 0, (cudaStream_t)stream>>>(w.d2, n, w.st, w.hist);
  else hist_kernel<<<(int)blocks, 512, 0, (cudaStream_t)stream>>>(w.d2, n, w.st, shift, nbits, himask, w.hist);
  return check_cuda(cudaGetLastError(), "hist launch");
}

extern "C" int bode_svgd_select_digit(int32_t pass, int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, bode_stream_t stream) {
  BODE_REQUIRE(pass >= 0 && pass < 3 && workspace, "bad pass/workspace");
  Ws w = carve(workspace, n_rows, n_cols, d);
  int shift, nbits; unsigned int himask;
  window(pass, &shift, &nbits, &himask);
  select_digit_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(w.st, w.hist, shift, nbits);
  return check_cuda(cudaGetLastError(), "select launch");
}

/* Single-rank fallback of the exact median: the three radix passes in one cooperative launch (returns immediately when
 * bode_svgd_window_select resolved the median).  With med_gamma != NULL the launch also does bode_svgd_gamma's work for the
 * median heuristic (median -> gamma, next window armed).  Multi-rank callers keep the per-pass calls with their all-reduces. */
/* Upper bound on the CTAs of the cooperative fallback launch (0 = one per SM slot).  A cooperative grid only starts when ALL its
 * CTAs can be resident: launched beside a kernel that fills most SMs (the fused solve in the overlapped SVGD step) a full-size grid
 * waits for that kernel to end -- even when the launch is the no-op it is after a window hit (measured on c3: the median chain ended
 * 7 us after the solve instead of 15 us before it).  Returns the previous bound. */
extern "C" int bode_svgd_set_select_ctas(int32_t n) {
  const int old = g_select_ctas;
  g_select_ctas = n > 0 ? n : 0;
  return old;
}

extern "C" int bode_svgd_radix_fallback(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, int32_t n_total, float* med_gamma,
                                        bode_stream_t stream) {
  BODE_REQUIRE(workspace, "null workspace");
  Ws w = carve(workspace, n_rows, n_cols, d);
  static int grid_blocks = 0;
  if (grid_blocks == 0) {
    int per_sm = 0;
    BODE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, radix_fallback_kernel, 1024, 0));
    const int sms = bode_device_sm_count();
    if (sms < 0) return BODE_ERR_CUDA;
    BODE_REQUIRE(per_sm >= 1, "radix_fallback_kernel does not fit an SM");
    grid_blocks = sms * per_sm;
  }
  const float* d2 = w.d2;
  long long n = (long long)n_rows * n_cols;
  SelState* stp = w.st;
  unsigned long long* hist = w.hist;
  int ntot = n_total;
  int arm = (g_tensor_cores && svgd_tc_supported(d) && svgd_tc2_supported(d, n_cols)) ? 1 : 0;
  float* mg = med_gamma;
  PeerInfo peer = peers_of(workspace);
  unsigned long long* hsum = svgd_tc2_table(w.ops, n_rows, n_cols) + WIN_TABLE + 1;   // spare half of the table block
  BODE_REQUIRE(peer.world <= 1 || arm, "peer-mapped workspaces need the pipelined tensor-core path (column count a multiple of 4, d <= 55)");
  void* args[] = {(void*)&d2, (void*)&n, (void*)&stp, (void*)&hist, (void*)&hsum, (void*)&ntot, (void*)&arm, (void*)&mg, (void*)&peer};
  const int gb = (g_select_ctas > 0 && g_select_ctas < grid_blocks) ? g_select_ctas : grid_blocks;
  BODE_CUDA(cudaLaunchCooperativeKernel((const void*)radix_fallback_kernel, dim3(gb), dim3(1024), args, 0, (cudaStream_t)stream));
  return BODE_OK;
}

/* med_gamma[0] = median(d2), med_gamma[1] = gamma; sigma > 0 fixes the bandwidth (RBFKernel(sigma), stein.py:13-16,28) */
extern "C" int bode_svgd_gamma(int32_t n_total, float sigma, int32_t n_rows, int32_t n_cols, int32_t d, void* workspace,
                               float* med_gamma, bode_stream_t stream) {
  BODE_REQUIRE(workspace && med_gamma && n_total > 0, "bad args");
  Ws w = carve(workspace, n_rows, n_cols, d);
  const int arm = (sigma <= 0.f && g_tensor_cores && svgd_tc_supported(d) && svgd_tc2_supported(d, n_cols)) ? 1 : 0;
  gamma_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(w.st, n_total, sigma, arm, med_gamma);
  return check_cuda(cudaGetLastError(), "gamma launch");
}

/* phi for the local rows (uses d2 left in the workspace by bode_svgd_sqdist).  Scols holds score_sign * score, so the
 * gradient of the negative log posterior can be passed as is with score_sign = -1 (score = -grad loss).  phi may be NULL; when theta != NULL the
 * update theta_i += step * phi_i is fused (the wrapped optimiser of stein.py descends -phi with lr = step). */
extern "C" int bode_svgd_phi_staged(int32_t stages, const float* Xrows, int64_t ld_rows, int32_t n_rows, const float* Xcols, int64_t ld_xc,
                                    const float* Scols, int64_t ld_sc, float score_sign, int32_t n_cols, int32_t d, int32_t n_total,
                                    const float* med_gamma, void* workspace, float* phi, int64_t ld_phi, float* theta,
                                    int64_t ld_theta, float step, bode_stream_t stream) {
  BODE_REQUIRE(Xrows && Xcols && med_gamma && workspace, "null pointer");
  BODE_REQUIRE((stages & ~7) == 0 && stages != 0,
               "stages must be a combination of BODE_SVGD_PREPARE | BODE_SVGD_COMPUTE | BODE_SVGD_PREPARE_POSITIONS");
  BODE_REQUIRE(Scols || !(stages & BODE_SVGD_PREPARE), "null scores");
  if (!bode_svgd_staged_supported(n_cols, d)) {
    if (!(stages & BODE_SVGD_COMPUTE)) return BODE_OK;
    BODE_REQUIRE(Scols, "null scores");
    stages = BODE_SVGD_PREPARE | BODE_SVGD_COMPUTE;
  }
  BODE_REQUIRE(d > 0 && 2 * d + 1 <= 256, "svgd phi kernel supports d <= 127 (got %d)", d);
  Ws w = carve(workspace, n_rows, n_cols, d);
  cudaStream_t st = (cudaStream_t)stream;
  const int rb = (n_rows + PR - 1) / PR;
  int sms = bode_device_sm_count();
  if (sms < 0) return BODE_ERR_CUDA;
  int jsplit = (3 * sms) / rb;                          // one full wave at 3 resident CTAs per SM, never a tail wave
  if (jsplit > MAXSPLIT) jsplit = MAXSPLIT;
  if (jsplit > (n_cols + PJ - 1) / PJ) jsplit = (n_cols + PJ - 1) / PJ;
  if (jsplit < 1) jsplit = 1;
  if (g_tensor_cores && svgd_tc_supported(d)) {
    if (svgd_tc2_supported(d, n_cols)) {
      int js2 = 1;
      // the split-K combine and the update theta += step * phi run in the tail of the phi kernel (last CTA of each row block)
      return svgd_tc2_phi(w.d2, n_rows, n_cols, Xcols, ld_xc, Scols, ld_sc, d, w.mu, med_gamma, score_sign, w.ops, &js2, w.part, sms, stages,
                          Xrows, ld_rows, 1.f / (float)n_total, phi, ld_phi, theta, ld_theta, step, st);
    }
    int js = (2 * sms) / ((n_rows + 127) / 128);        // ~2 CTAs per SM, one wave
    if (js > MAXSPLIT) js = MAXSPLIT;
    if (js > (n_cols + 31) / 32) js = (n_cols + 31) / 32;
    if (js < 1) js = 1;
    int e = svgd_tc_phi(w.d2, n_rows, n_cols, Xcols, ld_xc, Scols, ld_sc, d, w.mu, med_gamma, score_sign, js, w.part, st);
    if (e != BODE_OK) return e;
    return svgd_tc_combine(w.part, js, n_rows, d, Xrows, ld_rows, w.mu, med_gamma, 1.f / (float)n_total, phi, ld_phi, theta, ld_theta,
                           step, st);
  }
  dim3 grid(rb, jsplit);
  const int ft = (2 * d + 1 + 15) / 16;
#define BODE_PHI(C)                                                                                              \
  {                                                                                                              \
    const size_t smem = ((size_t)PJ * KP + (size_t)PJ * 16 * C) * sizeof(float);                                 \
    phi_partial_kernel<C><<<grid, 128, smem, st>>>(w.d2, n_rows, n_cols, Xcols, ld_xc, Scols, ld_sc, d, med_gamma, jsplit, score_sign, w.part); \
  }
  if (ft <= 2) BODE_PHI(2)
  else if (ft <= 4) BODE_PHI(4)
  else if (ft <= 7) BODE_PHI(7)
  else if (ft <= 10) BODE_PHI(10)
  else BODE_PHI(16)
#undef BODE_PHI
  BODE_CUDA(cudaGetLastError());
  const long long tot = (long long)n_rows * d;
  phi_combine_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(w.part, jsplit, n_rows, d, Xrows, ld_rows, med_gamma,
                                                               1.f / (float)n_total, phi, ld_phi, theta, ld_theta, step);
  return check_cuda(cudaGetLastError(), "phi combine launch");
}

/* Score-tile fusion.  arm: from now until bode_svgd_disarm_score_tiles, every fused likelihood closure of the component-split npde
 * kernels (bode_npde_nlp_grad on a 3x3 .. 6x6 grid) over exactly n_cols particles with 2m + 2 == d parameters also writes
 * score_sign * gradient into this workspace's phi operand tiles, so that the interaction needs only
 * bode_svgd_phi_staged(BODE_SVGD_PREPARE_POSITIONS) (any time after the operands' PREPARE) and ..._staged(BODE_SVGD_COMPUTE).
 * Returns 1 when armed, 0 when the shapes are outside the pipelined tensor-core path (nothing changes then).
 * disarm: returns how many closure launches wrote the tiles since arm (0: the caller must run BODE_SVGD_PREPARE itself). */
namespace bode {
void npde_arm_score_tiles(float* VH, float* VC, float sign, int P, int d);
int npde_disarm_score_tiles();
void svgd_tc2_v_tiles(void* ops_base, int nr, int nc, float** VH, float** VC);
}
extern "C" int bode_svgd_arm_score_tiles(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, float score_sign) {
  if (!workspace || n_rows != n_cols || (n_cols & 3) || !bode_svgd_staged_supported(n_cols, d)) return 0;
  if (!(g_tensor_cores && svgd_tc_supported(d) && svgd_tc2_supported(d, n_cols))) return 0;
  if (peers_of(workspace).world > 1) return 0;
  Ws w = carve(workspace, n_rows, n_cols, d);
  float *VH = nullptr, *VC = nullptr;
  svgd_tc2_v_tiles(w.ops, n_rows, n_cols, &VH, &VC);
  npde_arm_score_tiles(VH, VC, score_sign, n_cols, d);
  return 1;
}
extern "C" int bode_svgd_disarm_score_tiles(void) { return npde_disarm_score_tiles(); }

extern "C" int bode_svgd_phi(const float* Xrows, int64_t ld_rows, int32_t n_rows, const float* Xcols, int64_t ld_xc,
                             const float* Scols, int64_t ld_sc, float score_sign, int32_t n_cols, int32_t d, int32_t n_total,
                             const float* med_gamma, void* workspace, float* phi, int64_t ld_phi, float* theta,
                             int64_t ld_theta, float step, bode_stream_t stream) {
  return bode_svgd_phi_staged(BODE_SVGD_PREPARE | BODE_SVGD_COMPUTE, Xrows, ld_rows, n_rows, Xcols, ld_xc, Scols, ld_sc, score_sign, n_cols,
                              d, n_total, med_gamma, workspace, phi, ld_phi, theta, ld_theta, step, stream);
}
