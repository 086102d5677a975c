// Pipelined SVGD contractions on tcgen05 (3xTF32 Gram; K @ V as TF32 hi.hi + ONE bf16 product for both correction terms; TMEM
// accumulators) for shapes whose column count is a multiple of 4.
//
//   prep_x_kernel / prep_v_kernel   centre, split (hi = top 19 bits, lo = x - hi) and store the operands ONCE per step in global
//                                   memory, already in the canonical no-swizzle K-major UMMA layout, so that every later
//                                   operand tile is one contiguous block moved by a 1-D bulk copy (TMA engine, UBLKCP).
//   gram2_kernel                    d2 = |xc_i|^2 + |xc_j|^2 - 2 xc_i.xc_j   (stein.py:22).  One CTA keeps its 128-row A tile
//                                   resident and walks 128-column B tiles through a 2-slot ring; the accumulator is
//                                   double-buffered in TMEM (2 x 128 columns) so tile t+1's MMAs run under tile t's epilogue.
//                                   The epilogue also counts the window of the exact median selection (svgd_state.cuh) and
//                                   hands the d2 tile to the TMA engine (tensor stores from swizzled per-warp staging); d2 is
//                                   row-major, or [rows/128][cols/32][128][32] stage tiles when both edges are whole tiles.
//   phi2_kernel                     part[i,:] = sum_j 2^(-g d2_ij) [ -grad_j | xc_j | 1 ]   (stein.py:75-86).  d2 tiles arrive
//                                   by 2-D TMA six stages ahead, V^T tiles by bulk copy, the exp/split of stage t runs
//                                   under the MMAs of stage t-1.
// Warp 16 issues the MMAs (its issue loop is back-pressured by the tensor core), warp 17 the bulk / TMA copies; warps 0-15 (four per
// SM sub-partition) own the TMEM lanes / operand generation.  Inside the tile loops the warps meet only through mbarriers.
#include "tc_ptx.cuh"
#include "svgd_tiles.cuh"
#include "svgd_state.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cooperative_groups.h>

namespace bode {

constexpr int KP2 = 56;                       // K (feature dim) padded to a multiple of 8
constexpr int KCH2 = KP2 / 4;                 // 16-byte chunks along K
constexpr int BLK = 128;                      // rows per operand block / tile edge
constexpr uint32_t BLK_BYTES = BLK * KP2 * 4; // 28672: one hi (or lo) operand tile
constexpr int NF2 = 112;                      // padded feature count of V = [-G | X - mu | 1 | 0]: multiple of 16, >= 2d+1
constexpr int PK2 = 32;                       // j per phi stage
constexpr uint32_t VST_BYTES = PK2 * NF2 * 4; // 14336: one hi (or lo) V^T stage
constexpr int NWORK = 512;                    // worker threads (16 warps: 4 per SM sub-partition); warp 16 is the control warp
constexpr int NWARP = NWORK / 32;
constexpr int NTHR_PHI = NWORK + 64;            // warp 16 issues the MMAs, warp 17 the TMA / bulk loads

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- operand preparation
// XH/XL[blk][kc][r/8][r%8][4] : K-major core matrices of the centred rows, zero padded to a multiple of 128 rows.
// One CTA per 128-row block; norms[r] = |xc_r|^2 summed chunk by chunk in a fixed order.
// Centre: d2 is translation invariant, the centre only keeps |xc|^2 small against the cancellation in |xi|^2 + |xj|^2 - 2 xi.xj.
// Every CTA therefore forms the SAME reference point itself -- the mean of the first min(n_ref, 128) rows of the (gathered)
// column set, summed in a fixed order, 26 KB of L2 reads -- instead of waiting for a separate column-mean launch over all
// particles.  CTA 0 of the column launch publishes it as mu (read by prep_v / the combine kernel, which need a point that is
// consistent, not a particular one: phi is invariant to it as well) and starts the order-statistic selection (ranks of the two
// middle elements, stein.py:25-26); `sel` is null for the row-operand launch.
constexpr int CREF = 128;
__global__ void __launch_bounds__(256) prep_x_kernel(const float* __restrict__ X, long long ld, int n, int d, const float* __restrict__ Xref,
                                                     long long ldref, int n_ref, float* __restrict__ mu, float* __restrict__ XH,
                                                     float* __restrict__ XL, float* __restrict__ norms, SelState* sel,
                                                     unsigned long long total) {
  __shared__ float part[KCH2][BLK + 1];
  __shared__ float cpart[4][64];
  __shared__ float ctr[64];
  const int blk = blockIdx.x;
  {
    const int k = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int nr = n_ref < CREF ? n_ref : CREF;
    float acc = 0.f;
    if (k < d)
      for (int r = g * (CREF / 4); r < (g + 1) * (CREF / 4) && r < nr; ++r) acc += __ldg(Xref + (long long)r * ldref + k);
    cpart[g][k] = acc;
    __syncthreads();
    if (threadIdx.x < 64) {
      const float c = (k < d) ? ((cpart[0][k] + cpart[1][k]) + (cpart[2][k] + cpart[3][k])) / (float)nr : 0.f;
      ctr[k] = c;
      if (blk == 0 && sel && k < d) mu[k] = c;
    }
    if (blk == 0 && threadIdx.x == 0 && sel) {       // start of a selection: same reset as select_init_kernel (svgd.cu); maxbits is
      sel->prefix[0] = sel->prefix[1] = 0u;          // re-armed by gamma_kernel / workspace_init (other CTAs are raising it now)
      sel->hit = 0u;
      sel->rank[0] = (total - 1) / 2;
      sel->rank[1] = total / 2;
    }
    __syncthreads();
  }
  unsigned int* maxbits = sel ? &sel->maxbits : nullptr;
  for (int idx = threadIdx.x; idx < BLK * KCH2; idx += 256) {
    const int r = idx % BLK, kc = idx / BLK;
    const long long row = (long long)blk * BLK + r;
    float h[4], l[4], s = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * kc + e;
      const float v = (row < n && k < d) ? __ldg(X + row * ld + k) - ctr[k] : 0.f;
      s = fmaf(v, v, s);
      split_tf32(v, h[e], l[e]);
    }
    part[kc][r] = s;
    const long long off = (long long)blk * (BLK_BYTES / 4) + (long long)(kc * (BLK / 8) + (r >> 3)) * 32 + (r & 7) * 4;
    *reinterpret_cast<float4*>(XH + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(XL + off) = make_float4(l[0], l[1], l[2], l[3]);
  }
  __syncthreads();
  if (threadIdx.x < BLK) {
    float s = 0.f;
#pragma unroll
    for (int kc = 0; kc < KCH2; ++kc) s += part[kc][threadIdx.x];
    norms[(long long)blk * BLK + threadIdx.x] = s;
    // d2_ij <= (|xc_i| + |xc_j|)^2 <= 4 max |xc|^2: positions the hot buckets of the radix fallback's first pass (svgd.cu)
    float m = 4.f * s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && maxbits) atomicMax(maxbits, __float_as_uint(m));
  }
}

// VH[j/4][f][j%4] : K-major (K = particle index j) tf32 core matrices of V^T (hi parts), zero padded to n_pad particles (multiple of 32).
// VC: the B operand of the CORRECTION product.  K V = K_hi V_hi + (K_hi V_lo + K_lo V_hi) + O(2^-22); the bracket only needs 2^-9
// relative accuracy, so it is ONE bf16 product of twice the K extent, [bf16 K | bf16 K_lo] . [bf16 V_lo ; bf16 V_hi]: per 32-particle
// stage 64 bf16 rows -- rows 0..31 = V_lo[j], rows 32..63 = V_hi[j] -- as 8 K-chunks of 8 bf16 (16 bytes) x 112 features, i.e. the
// same 14336 bytes, LBO and SBO as a tf32 stage.  8 instead of 12 MMA slots per stage, error ~2^-19.
__device__ __forceinline__ uint2 pack_bf16x4(const float (&v)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);   // .x = low half = lower k
  return make_uint2(*reinterpret_cast<const unsigned int*>(&a), *reinterpret_cast<const unsigned int*>(&b));
}
__global__ void __launch_bounds__(256) prep_v_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ G, long long ldg,
                                                     int n, int d, const float* __restrict__ mu, float gsign, int n_pad,
                                                     float* __restrict__ VH, float* __restrict__ VC, int f_lo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(n_pad / 4) * NF2) return;
  const int f = (int)(idx % NF2);
  const long long jq = idx / NF2;
  // f_lo = d: the score columns are written by the closure kernel (svgd_tiles.cuh); only their zero padding (particles >= n,
  // whole quads: n % 4 == 0 on this path) is kept up here
  if (f < f_lo && 4 * jq < n) return;
  float h[4], l[4], full[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const long long j = 4 * jq + e;
    float v = 0.f;
    if (j < n) {
      if (f < d) v = gsign * __ldg(G + j * ldg + f);
      else if (f < 2 * d) v = __ldg(X + j * ldx + (f - d)) - __ldg(mu + f - d);
      else if (f == 2 * d) v = 1.f;
    }
    full[e] = v;
    split_tf32(v, h[e], l[e]);
  }
  *reinterpret_cast<float4*>(VH + idx * 4) = make_float4(h[0], h[1], h[2], h[3]);
  // stage s = j / 32; inside it this thread's 4 particles are half (jq % 2) of K-chunk (jq % 8) / 2 (V_lo) and of chunk 4 + that (V_hi)
  const long long s = jq >> 3;
  const int c = (int)(jq & 7) >> 1, half = (int)(jq & 1);
  unsigned char* base = reinterpret_cast<unsigned char*>(VC) + s * VST_BYTES + (long long)f * 16 + half * 8;
  *reinterpret_cast<uint2*>(base + (long long)c * (NF2 * 16)) = pack_bf16x4(l);
  *reinterpret_cast<uint2*>(base + (long long)(c + 4) * (NF2 * 16)) = pack_bf16x4(full);
}

// ---------------------------------------------------------------- Gram tiles + window count
// One 32 x 32 chunk of a Gram tile: d2 values, window count, staged coalesced store.  FULL: the tile lies inside the matrix
// (no bounds checks); DIAG: the tile intersects the diagonal (cdist(x, x) = 0 is forced there).
// The d2 tile leaves the SM through the TMA engine: each warp stages 32 rows x 16 columns (2 KB, 64-byte rows, 64-byte swizzle so
// that the lanes' 16-byte stores fall into distinct banks) and one lane issues a 2-D tensor store, which also clips the ragged
// edges.  (The same data written with st.global -- 8 row segments of 64 bytes per warp instruction -- kept the LSU busy for 40 % of
// the kernel: 158 -> 94 us at 4096 x 32768 with the stores removed.)
// `policy` != 0: an L2 evict-first policy (l2_evict_first()).  A d2 block that cannot stay in L2 anyway (several ranks: 134 MB and
// up) is streamed through it, so that it does not push out what the kernels running beside the Gram pass keep there -- the fused
// solve's stage checkpoints (its launch took 103 instead of 91 us next to an unhinted 4096 x 8192 Gram pass) and the V operand tiles.
__device__ __forceinline__ uint64_t l2_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* src_smem, int c0, int c1, uint64_t policy) {
  if (policy)
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(tmap), "r"(c0),
                 "r"(c1), "r"(smem_u32(src_smem)), "l"(policy)
                 : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1),
                 "r"(smem_u32(src_smem))
                 : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the same store into the tiled d2 layout [row block][column stage][128][32] (see svgd_tc2_d2_tiled)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tmap, const void* src_smem, int c0, int c1, int c2, int c3, uint64_t policy) {
  if (policy)
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4}], [%5], %6;" ::"l"(tmap),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src_smem)), "l"(policy)
                 : "memory");
  else
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(tmap), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3), "r"(smem_u32(src_smem))
                 : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <bool FULL, bool DIAG>
__device__ __forceinline__ void gram2_chunk(const float (&s)[32], const float* ncs, unsigned char* stage, const CUtensorMap* tmD2, float nrow,
                                            int lane, int dcol, int row, int nr, int colbase, int nc, int grow0, unsigned int wlo,
                                            unsigned int wspan, unsigned long long* table, unsigned int& below, bool tiled, uint64_t policy) {
  const uint32_t sw = (uint32_t)(lane >> 1) & 3u;                // 64-byte swizzle: 16-byte chunk c of row r sits at chunk c ^ ((r / 2) % 4)
  unsigned char* myrow = stage + lane * 64;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    // Entries inside the window are rare (a fraction of a percent): the straight-line code only records WHICH of the lane's 16
    // entries fell inside (one bit each); the reductions into the table are issued afterwards, off the common path.  (One
    // predicated `red` per entry compiled to a BSSY / BRA / BSYNC region per entry: 17 instructions per entry instead of 8.)
    unsigned int hits = 0u;
    float o[16];
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const float4 nj = *reinterpret_cast<const float4*>(ncs + 16 * hf + 4 * k4);
      const float njv[4] = {nj.x, nj.y, nj.z, nj.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = 16 * hf + 4 * k4 + e;
        float v = fmaxf(fmaf(-2.f, s[c], nrow + njv[e]), 0.f);
        if (DIAG && c == dcol) v = 0.f;                          // cdist(x, x) = 0 on the diagonal
        o[4 * k4 + e] = v;
        unsigned int off = __float_as_uint(v) - wlo;             // wraps (sign bit set) for entries below the window
        if (!FULL && !(row < nr && colbase + c < nc)) off = 0x7fffffffu;      // outside the matrix: neither below nor inside
        below += off >> 31;
        hits |= (off <= wspan) ? (1u << (4 * k4 + e)) : 0u;
      }
    }
    if (lane == 0) bulk_wait_read0();                            // the previous store has read the staging buffer (long ago, usually)
    __syncwarp();
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4)
      *reinterpret_cast<float4*>(myrow + (((uint32_t)k4 ^ sw) << 4)) = make_float4(o[4 * k4], o[4 * k4 + 1], o[4 * k4 + 2], o[4 * k4 + 3]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      if (tiled) tma_store_4d(tmD2, stage, 16 * hf, grow0 & (BLK - 1), colbase / PK2, grow0 / BLK, policy);   // a chunk is one column stage wide
      else tma_store_2d(tmD2, stage, colbase + 16 * hf, grow0, policy);
    }
    while (hits) {                                               // the lane's own staged row still holds the values
      const int b = __ffs(hits) - 1;
      hits &= hits - 1u;
      const float v = *reinterpret_cast<const float*>(myrow + ((((uint32_t)b >> 2) ^ sw) << 4) + ((b & 3) << 2));
      atomicAdd(table + (__float_as_uint(v) - wlo), 1ull);
    }
  }
}

struct Gram2Smem {
  static constexpr uint32_t A = 0;                                  // hi | lo
  static constexpr uint32_t B = 2 * BLK_BYTES;                      // 2 slots x (hi | lo)
  static constexpr uint32_t STAGE = B + 4 * BLK_BYTES;              // per warp 32 rows x 64 bytes (16 columns at a time), 64-byte swizzle
  static constexpr uint32_t STAGE_WARP = 32 * 64;
  static constexpr uint32_t NCS = STAGE + NWARP * STAGE_WARP;       // per warp 32 column norms
  static constexpr uint32_t BARS = NCS + NWARP * 32 * 4;            // barA, barB[2], barS[2], barE[2]
  static constexpr uint32_t TSLOT = BARS + 7 * 8;
  static constexpr uint32_t TOTAL = TSLOT + 16 + 1024;              // + slack to align the dynamic base to 1024 bytes
};
static_assert(Gram2Smem::STAGE % 1024 == 0, "swizzled staging buffers need an aligned base");

__global__ void __launch_bounds__(NTHR_PHI, 1) gram2_kernel(const __grid_constant__ CUtensorMap tmD2, const float* __restrict__ XrH, const float* __restrict__ XrL,
                                                        const float* __restrict__ nrm_r, int nr, int row_offset,
                                                        const float* __restrict__ XcH, const float* __restrict__ XcL,
                                                        const float* __restrict__ nrm_c, int nc, int tiles_per_cta,
                                                        SelState* __restrict__ st, unsigned long long* __restrict__ table, int tiled, int stream_d2) {
  extern __shared__ unsigned char sm_raw_g[];
  unsigned char* sm = sm_raw_g + ((1024u - (smem_u32(sm_raw_g) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Gram2Smem::BARS);
  uint64_t* barA = bars;
  uint64_t* barB = bars + 1;
  uint64_t* barS = bars + 3;      // MMAs of a tile retired: accumulator readable, B slot free
  uint64_t* barE = bars + 5;      // all 16 worker warps have pulled the accumulator of a tile into registers (count NWARP)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + Gram2Smem::TSLOT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rb = blockIdx.y;
  const int nct = (nc + BLK - 1) / BLK;
  const int ct0 = blockIdx.x * tiles_per_cta;
  const int nt = min(tiles_per_cta, nct - ct0);
  if (nt <= 0) return;

  if (warp == 0) tmem_alloc(tslot, 256);
  if (tid == 0) {
    mbar_init(barA, 1);
    mbar_init(barB + 0, 1);
    mbar_init(barB + 1, 1);
    mbar_init(barS + 0, 1);
    mbar_init(barS + 1, 1);
    mbar_init(barE + 0, NWARP);
    mbar_init(barE + 1, NWARP);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  constexpr uint32_t idesc = idesc_tf32(BLK, BLK, 0, 0);
  constexpr uint32_t LBO = BLK * 16, SBO = 128;

  // No CTA-wide barrier in the tile loop: the MMA thread, the loader thread and the 16 epilogue warps meet only through mbarriers.
  if (warp == NWARP + 1) {
    // ---------------- loader warp (lane 0): A once, B tiles two ahead; a B slot is free when the MMAs that read it retired
    if (lane == 0) {
      auto load_b = [&](int t) {
        const int slot = t & 1;
        unsigned char* dst = sm + Gram2Smem::B + slot * 2 * BLK_BYTES;
        const long long src = (long long)(ct0 + t) * (BLK_BYTES / 4);
        mbar_expect_tx(barB + slot, 2 * BLK_BYTES);
        bulk_g2s(dst, XcH + src, BLK_BYTES, barB + slot);
        bulk_g2s(dst + BLK_BYTES, XcL + src, BLK_BYTES, barB + slot);
      };
      mbar_expect_tx(barA, 2 * BLK_BYTES);
      bulk_g2s(sm + Gram2Smem::A, XrH + (long long)rb * (BLK_BYTES / 4), BLK_BYTES, barA);
      bulk_g2s(sm + Gram2Smem::A + BLK_BYTES, XrL + (long long)rb * (BLK_BYTES / 4), BLK_BYTES, barA);
      load_b(0);
      if (nt > 1) load_b(1);
      for (int t = 0; t + 2 < nt; ++t) {
        mbar_wait(barS + (t & 1), (t >> 1) & 1);
        load_b(t + 2);
      }
    }
  } else if (warp == NWARP) {
    // ---------------- MMA warp (all lanes, one elected issuer): 21 MMAs per tile into accumulator t & 1, free once every worker
    // warp has read tile t-2
    {
      const uint64_t dAh = smem_desc(smem_u32(sm + Gram2Smem::A), LBO, SBO), dAl = smem_desc(smem_u32(sm + Gram2Smem::A) + BLK_BYTES, LBO, SBO);
      const uint64_t dB0h = smem_desc(smem_u32(sm + Gram2Smem::B), LBO, SBO), dB0l = smem_desc(smem_u32(sm + Gram2Smem::B) + BLK_BYTES, LBO, SBO);
      const uint64_t dB1h = smem_desc(smem_u32(sm + Gram2Smem::B) + 2 * BLK_BYTES, LBO, SBO),
                     dB1l = smem_desc(smem_u32(sm + Gram2Smem::B) + 3 * BLK_BYTES, LBO, SBO);
      constexpr uint64_t KSTEP = (2 * LBO) >> 4;       // a K step only adds to the 14-bit start-address field
      mbar_wait(barA, 0);
      for (int t = 0; t < nt; ++t) {
        const int slot = t & 1;
        mbar_wait(barB + slot, (t >> 1) & 1);
        if (t >= 2) mbar_wait(barE + slot, ((t - 2) >> 1) & 1);
        tc_fence_after();
        const uint64_t bh = slot ? dB1h : dB0h, bl = slot ? dB1l : dB0l;
        const uint32_t dst = tmem + slot * BLK;
#pragma unroll
        for (int ks = 0; ks < KP2 / 8; ++ks) umma_tf32_w(dst, dAh + ks * KSTEP, bh + ks * KSTEP, idesc, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < KP2 / 8; ++ks) umma_tf32_w(dst, dAh + ks * KSTEP, bl + ks * KSTEP, idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < KP2 / 8; ++ks) umma_tf32_w(dst, dAl + ks * KSTEP, bh + ks * KSTEP, idesc, 1u);
        umma_commit_w(barS + slot);
      }
    }
  } else {
    // ---------------- worker warps: epilogue.  Warp w reads TMEM lanes 32 (w % 4) .. +31 (tile rows), columns 32 (w / 4) .. +31
    const int q = warp & 3, cq = warp >> 2;
    const int rl = 32 * q + lane;
    const int row = rb * BLK + rl;
    const float nrow = __ldg(nrm_r + row);
    unsigned char* stage = sm + Gram2Smem::STAGE + warp * Gram2Smem::STAGE_WARP;
    float* ncs = reinterpret_cast<float*>(sm + Gram2Smem::NCS) + warp * 32;
    const bool win = st->win_valid != 0;
    // disarmed: wlo = 0x80000000 makes every offset wrap to >= 0x80000000 - 0x7f800000 > 0 with the sign bit clear only for
    // bit patterns >= 0x80000000 (none: d2 >= 0), i.e. nothing falls inside; the "below" count is simply not published
    const unsigned int wlo = win ? st->win_lo : 0x80000000u;
    const unsigned int wspan = win ? WIN_SPAN : 0u;
    unsigned int below = 0;
    const uint64_t policy = stream_d2 ? l2_evict_first() : 0ull;
    const int cb = 32 * cq;                                        // first tile column of this warp
    float nreg = __ldg(nrm_c + ct0 * BLK + cb + lane);
    for (int t = 0; t < nt; ++t) {
      const int c0 = (ct0 + t) * BLK;
      ncs[lane] = nreg;
      if (t + 1 < nt) nreg = __ldg(nrm_c + c0 + BLK + cb + lane);  // next tile's column norms travel during this tile
      mbar_wait(barS + (t & 1), (t >> 1) & 1);
      tc_fence_after();
      float s[32];
      tmem_ld32(tmem + (t & 1) * BLK + cb + ((uint32_t)(32 * q) << 16), s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(barE + (t & 1));                  // this warp's part of the accumulator is in registers
      const bool full = rb * BLK + BLK <= nr && c0 + BLK <= nc;
      const int dcol = row + row_offset - (c0 + cb);               // chunk-local column of the diagonal entry (if 0..31)
      const int dtile = rb * BLK + 32 * q + row_offset - (c0 + cb);  // warp-uniform: diagonal crosses this chunk iff -31 <= dtile <= 31
      const bool diag = dtile > -32 && dtile < 32;
      const int grow0 = rb * BLK + 32 * q, colbase = c0 + cb;
      if (full && !diag) gram2_chunk<true, false>(s, ncs, stage, &tmD2, nrow, lane, dcol, row, nr, colbase, nc, grow0, wlo, wspan, table, below, tiled != 0, policy);
      else if (full) gram2_chunk<true, true>(s, ncs, stage, &tmD2, nrow, lane, dcol, row, nr, colbase, nc, grow0, wlo, wspan, table, below, tiled != 0, policy);
      else gram2_chunk<false, true>(s, ncs, stage, &tmD2, nrow, lane, dcol, row, nr, colbase, nc, grow0, wlo, wspan, table, below, tiled != 0, policy);
    }
    if (lane == 0) bulk_wait_all0();                               // this warp's tensor stores are complete before the CTA retires
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0 && win && below) atomicAdd(table + WIN_TABLE, (unsigned long long)below);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------- median from the window table
// (the cluster kernel follows the block-scan helper)

__device__ __forceinline__ unsigned long long block_incl_scan_u64(unsigned long long v, unsigned long long* wsum, unsigned long long* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();                       // wsum may still be read from a previous call
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
  if (total) *total = wsum[31];
  return incl + (wid ? wsum[wid - 1] : 0ull);
}

// One thread-block cluster of 16 CTAs x 1024 threads, two adjacent bins per thread (16 x 2048 = 32768 bins; bin 32768 and the
// "below" counter are handled by CTA 0): ONE global round trip for the table, segment totals exchanged through distributed
// shared memory, two cluster barriers.  The 33-CTA ticket version this replaces needed ~8 dependent global round trips
// (table -> totals -> ticket -> totals -> segment -> segment) and took 10.9 us for 262 KB.
constexpr int WSEL_CLUSTER = 16;
static_assert(WSEL_CLUSTER * 2048 == (int)WIN_SPAN, "two bins per thread must tile the window");

// Several ranks (peer.world > 1): the table is the SUM of every rank's table, read straight from the peers' workspaces over NVLink
// behind a flag barrier (svgd_state.cuh) -- no all-reduce launch; every rank computes the same sums, hence the same selection.
__global__ void __launch_bounds__(1024) window_select_kernel(SelState* st, unsigned long long* table, const PeerInfo peer) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const bool multi = peer.world > 1;
  if (multi) {                                                        // every rank's Gram pass has filled its table
    if (cl.block_rank() == 0 && threadIdx.x == 0) peer_barrier(peer);
    cl.sync();
  }
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long segtot[WSEL_CLUSTER];
  __shared__ unsigned long long mytot;
  __shared__ unsigned int found[2];                                   // meaningful in CTA 0
  const int tid = threadIdx.x, c = (int)cl.block_rank();
  const bool armed = st->win_valid != 0;
  const unsigned long long r0 = st->rank[0], r1 = st->rank[1];
  ulonglong2* t2 = reinterpret_cast<ulonglong2*>(table) + c * 1024 + tid;
  ulonglong2 v;
  unsigned long long last = 0ull, below = 0ull;
  if (!multi) {
    v = *t2;
    *t2 = make_ulonglong2(0ull, 0ull);                                // ready for the next call
    if (tid == 0) {                                                   // every CTA needs `below`; CTA 0 clears it after the barrier
      last = table[WIN_SPAN];
      below = table[WIN_TABLE];
    }
  } else {
    // all peers' loads are in flight together (one NVLink round trip, not one per rank), then summed in rank order
    unsigned long long px[MAX_PEERS], py[MAX_PEERS];
#pragma unroll
    for (int q = 0; q < MAX_PEERS; ++q) {
      px[q] = py[q] = 0ull;
      if (q < peer.world) {
        const unsigned long long* tq = reinterpret_cast<const unsigned long long*>(peer.base[q] + peer.table_off);
        px[q] = peer_ld_u64(tq + 2 * (c * 1024 + tid));
        py[q] = peer_ld_u64(tq + 2 * (c * 1024 + tid) + 1);
      }
    }
    v = make_ulonglong2(0ull, 0ull);
#pragma unroll
    for (int q = 0; q < MAX_PEERS; ++q) {
      v.x += px[q];
      v.y += py[q];
    }
    if (tid == 0)
      for (int q = 0; q < peer.world; ++q) {
        const unsigned long long* tq = reinterpret_cast<const unsigned long long*>(peer.base[q] + peer.table_off);
        last += peer_ld_u64(tq + WIN_SPAN);
        below += peer_ld_u64(tq + WIN_TABLE);
      }
  }
  if (tid == 0 && c == 0) found[0] = found[1] = 0xffffffffu;
  if (!armed) v = make_ulonglong2(0ull, 0ull);
  const unsigned long long mine = v.x + v.y;
  const unsigned long long incl = block_incl_scan_u64(mine, wsum, &mytot);
  __syncthreads();
  if (tid < WSEL_CLUSTER) cl.map_shared_rank(segtot, tid)[c] = mytot;
  __shared__ unsigned long long below_s, last_s;
  if (tid == 0) {
    below_s = armed ? below : 0ull;
    last_s = armed ? last : 0ull;
  }
  cl.sync();
  if (!multi && c == 0 && tid == 0) table[WIN_SPAN] = table[WIN_TABLE] = 0ull;
  unsigned long long run = below_s;
  for (int i = 0; i < c; ++i) run += segtot[i];
  const unsigned long long e0 = run + incl - mine, e1 = e0 + v.x;
  const unsigned int b0 = (unsigned int)(c * 2048 + 2 * tid);
  unsigned int* f0 = cl.map_shared_rank(found, 0);
  const unsigned long long rr[2] = {r0, r1};
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    if (v.x && rr[which] >= e0 && rr[which] < e1) f0[which] = b0;
    else if (v.y && rr[which] >= e1 && rr[which] < e1 + v.y) f0[which] = b0 + 1;
  }
  if (c == WSEL_CLUSTER - 1 && tid == 0 && last_s) {                  // the closing bin of the window
    unsigned long long tot = below_s;
    for (int i = 0; i < WSEL_CLUSTER; ++i) tot += segtot[i];
#pragma unroll
    for (int which = 0; which < 2; ++which)
      if (rr[which] >= tot && rr[which] < tot + last_s) f0[which] = WIN_SPAN;
  }
  cl.sync();
  if (c == 0 && tid == 0) {
    const bool hit = armed && found[0] != 0xffffffffu && found[1] != 0xffffffffu;
    if (hit) {
      st->prefix[0] = st->win_lo + found[0];
      st->prefix[1] = st->win_lo + found[1];
    }
    st->hit = hit ? 1u : 0u;
  }
  if (multi) {                                                        // every rank has read my table: clear it for the next call
    if (c == 0 && tid == 0) peer_barrier(peer);
    cl.sync();
    *t2 = make_ulonglong2(0ull, 0ull);
    if (c == 0 && tid == 0) table[WIN_SPAN] = table[WIN_TABLE] = 0ull;
  }
}

// The same selection by a cluster of FOUR CTAs (single rank), eight bins per thread.  A 16-CTA cluster needs 16 free SMs inside
// one GPC; beside the fused solve of the overlapped SVGD step (108 of 148 SMs busy) no GPC has that many, so the launch waited for
// the solve to end and put the median chain -- and phi behind it -- after the solve (tools/step_timeline.py: Gram pass done at
// 76 us, selection done at 100 us, solve at 97 us).  40 free SMs over 8 GPCs always leave one GPC with five.  (One CTA instead is no
// answer: a single SM needs ~30 us to move the 262 KB table twice.)
constexpr int WSEL4 = 4;
constexpr int WSEL4_V = (int)WIN_SPAN / (WSEL4 * 1024) / 2;            // 16-byte vectors (two bins each) per thread: 4
__global__ void __launch_bounds__(1024) window_select4_kernel(SelState* st, unsigned long long* table) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long segtot[WSEL4];
  __shared__ unsigned long long mytot, below_s, last_s;
  __shared__ unsigned int found[2];                                   // meaningful in CTA 0
  const int tid = threadIdx.x, c = (int)cl.block_rank();
  const bool armed = st->win_valid != 0;
  const unsigned long long rr[2] = {st->rank[0], st->rank[1]};
  ulonglong2* t2 = reinterpret_cast<ulonglong2*>(table) + (c * 1024 + tid) * WSEL4_V;
  ulonglong2 v[WSEL4_V];
#pragma unroll
  for (int i = 0; i < WSEL4_V; ++i) v[i] = t2[i];
#pragma unroll
  for (int i = 0; i < WSEL4_V; ++i) t2[i] = make_ulonglong2(0ull, 0ull);   // ready for the next call
  if (tid == 0) {                                                     // every CTA needs `below`; CTA 0 clears it after the barrier
    below_s = armed ? table[WIN_TABLE] : 0ull;
    last_s = armed ? table[WIN_SPAN] : 0ull;
    if (c == 0) found[0] = found[1] = 0xffffffffu;
  }
  unsigned long long mine = 0ull;
#pragma unroll
  for (int i = 0; i < WSEL4_V; ++i) {
    if (!armed) v[i] = make_ulonglong2(0ull, 0ull);
    mine += v[i].x + v[i].y;
  }
  const unsigned long long incl = block_incl_scan_u64(mine, wsum, &mytot);
  __syncthreads();
  if (tid < WSEL4) cl.map_shared_rank(segtot, tid)[c] = mytot;
  cl.sync();
  if (c == 0 && tid == 0) table[WIN_SPAN] = table[WIN_TABLE] = 0ull;
  unsigned long long run = below_s;
  for (int i = 0; i < c; ++i) run += segtot[i];
  run += incl - mine;                                                 // entries before this thread's first bin
  const unsigned int b0 = (unsigned int)((c * 1024 + tid) * 2 * WSEL4_V);
  unsigned int* f0 = cl.map_shared_rank(found, 0);
  if (mine) {
#pragma unroll
    for (int i = 0; i < WSEL4_V; ++i) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        if (rr[which] >= run && rr[which] < run + v[i].x) f0[which] = b0 + 2 * i;
        else if (rr[which] >= run + v[i].x && rr[which] < run + v[i].x + v[i].y) f0[which] = b0 + 2 * i + 1;
      }
      run += v[i].x + v[i].y;
    }
  }
  if (c == WSEL4 - 1 && tid == 0 && last_s) {                         // the closing bin of the window
    unsigned long long tot = below_s;
    for (int i = 0; i < WSEL4; ++i) tot += segtot[i];
#pragma unroll
    for (int which = 0; which < 2; ++which)
      if (rr[which] >= tot && rr[which] < tot + last_s) f0[which] = WIN_SPAN;
  }
  cl.sync();
  if (c == 0 && tid == 0) {
    const bool hit = armed && found[0] != 0xffffffffu && found[1] != 0xffffffffu;
    if (hit) {
      st->prefix[0] = st->win_lo + found[0];
      st->prefix[1] = st->win_lo + found[1];
    }
    st->hit = hit ? 1u : 0u;
  }
}

// ---------------------------------------------------------------- phi partials
struct Phi2Smem {
  // d2 ring: NRAW slots x [128][32] floats, TMA tiles with 128-byte swizzle.  Six stages (96 KB per CTA, 12 MB over the GPU) in
  // flight: with three, the worker warps spent 37 % of their samples waiting for tiles once the d2 block no longer fits in L2
  // (several ranks: 4096 x 32768 floats = 537 MB per rank).
  static constexpr int NRAW = 6;
  static constexpr uint32_t RAW = 0;
  static constexpr uint32_t RAW_SLOT = BLK * PK2 * 4;                  // 16384 (1024-byte aligned)
  static constexpr uint32_t V = RAW + NRAW * RAW_SLOT;                 // 3 slots x (hi | lo); the K tile (A operand) lives in TMEM
  static constexpr uint32_t BARS = V + 6 * VST_BYTES;                  // barM[3], barR[NRAW], barKV[6], barF[2], barD[2]
  static constexpr uint32_t TSLOT = BARS + 24 * 8;
  static constexpr uint32_t TOTAL = TSLOT + 16 + 1024;                 // + slack to align the dynamic base to 1024 bytes
};

// 2-D tiled TMA load (SASS UTMALDG): box {32 columns, 128 rows} of the row-major d2 matrix, 128-byte swizzle
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar, uint64_t policy) {
  if (policy)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar,
                                            uint64_t policy) {
  if (policy)
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
  else
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (one row per TMEM lane, K 32-bit columns) never touches shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
               "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

__device__ __forceinline__ void tmem_ld4_add(uint32_t taddr, float* acc) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] += __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st4u(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// D[tmem] += A[tmem, bf16 pairs per 32-bit column] * B[smem desc, bf16], fp32 accumulate; warp-uniform issue (see umma_tf32_ts_w)
__device__ __forceinline__ void umma_bf16_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor, kind::f16: c = F32 [4,6), a = BF16 (1) [7,10), b = BF16 (1) [10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Arguments of the fused combine.
constexpr int ACC_PITCH = NF2 + 1;   // odd pitch: a warp's 32 rows hit 32 banks
struct Phi2Combine {
  const float* Xr;            // local rows (positions)
  long long ldr;
  const float* mu;
  float inv_n;
  float* phi;
  long long ldp;
  float* theta;
  long long ldt;
  float step;
};

__global__ void __launch_bounds__(NTHR_PHI, 1) phi2_kernel(const __grid_constant__ CUtensorMap tmD2, int nr, int nc, const float* __restrict__ VH,
                                                       const float* __restrict__ VC, int d, const float* __restrict__ gam, int jsplit,
                                                       float* __restrict__ part, const Phi2Combine cmb, int tiled, int stream_d2) {
  extern __shared__ unsigned char sm_raw[];
  // 1024-byte alignment for the swizzled TMA tiles, computed as an OFFSET so the pointer stays in the shared address space
  unsigned char* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Phi2Smem::BARS);
  // An mbarrier wait costs ~250 cycles on this part even when the phase is already complete (clock64 trace of one CTA), so the
  // MMA thread gets ONE barrier per stage: barKV completes when all 16 worker warps have written K(t) AND V(t) has landed.
  uint64_t* barM = bars;          // [3] MMAs of stage t retired: K slot t % 3 and V slot t % 3 are free
  constexpr int NRAW = Phi2Smem::NRAW;
  uint64_t* barR = bars + 3;      // [NRAW] d2 tile landed
  uint64_t* barKV = bars + 3 + NRAW;   // [6] stage t ready for the tensor core (count NWARP + 1 arrivals, + the V bytes)
  // Two-level accumulation.  The tensor core adds every MMA into its fp32 accumulator with TRUNCATION, so a long chain into one
  // growing sum is biased by ~half an ulp of that sum per MMA: measured 2e-5 (4096 columns) -> 5e-4 (32768 columns) relative error
  // of phi on B200 (tools/phi_accuracy.py), linear in the number of accumulated stages.  The MMAs therefore accumulate CHUNKS of
  // ACC_G stages into two alternating TMEM accumulators (small partial sums: small ulps) and the worker warps drain each finished
  // chunk into fp32 registers with round-to-nearest adds (28 columns of its row per thread, +7 % worker instructions).
  uint64_t* barF = barKV + 6;          // [2] chunk accumulator complete (tcgen05.commit after the chunk's last MMA)
  uint64_t* barD = barF + 2;           // [2] chunk accumulator drained by all NWARP worker warps: the MMA warp may overwrite it
  constexpr int ACC_G = 8;
  constexpr uint32_t ACC_B = 384;      // second accumulator: TMEM columns [384, 496)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + Phi2Smem::TSLOT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * BLK;
  const int nst_all = (nc + PK2 - 1) / PK2;
  const int per = (nst_all + jsplit - 1) / jsplit;
  const int s0 = blockIdx.y * per;
  const int nst = min(per, nst_all - s0);                              // stages of this CTA (may be <= 0)

  // TMEM columns: [0,112) and [384,496) the two chunk accumulators; slot s = 0..2: [128 + 64 s, +32) K_hi (tf32), [+32, +48) bf16 pairs of K, [+48, +64) of K_lo
  if (warp == 0) tmem_alloc(tslot, 512);
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(barM + i, 1);
    for (int i = 0; i < NRAW; ++i) mbar_init(barR + i, 1);
    for (int i = 0; i < 6; ++i) mbar_init(barKV + i, NWARP + 1);
    for (int i = 0; i < 2; ++i) { mbar_init(barF + i, 1); mbar_init(barD + i, NWARP); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  constexpr uint32_t idesc = idesc_tf32(BLK, NF2, 0, 0), idesc_c = idesc_bf16(BLK, NF2);
  constexpr uint32_t B_LBO = NF2 * 16, SBO = 128;

  // No CTA-wide barrier inside the stage loop: the warps only meet through mbarriers, so a slow warp delays nobody but the MMA
  // that needs its rows.
  if (warp == NWARP + 1) {
    // ---------------- loader warp (lane 0): d2 tiles by 2-D TMA NRAW stages ahead, V^T tiles by bulk copy two stages ahead.
    // It only waits for slots to drain (barKV: the workers consumed raw(t); barM: the MMAs that read V(t-1) retired), so the
    // MMA-issuing thread never spends time on copies.
    if (lane == 0 && nst > 0) {
      const uint64_t policy = stream_d2 ? l2_evict_first() : 0ull;     // d2 is read once: do not let it displace the V tiles in L2
      auto load_v = [&](int t) {
        const int slot = t % 3;
        unsigned char* dst = sm + Phi2Smem::V + slot * 2 * VST_BYTES;
        const long long src = (long long)(s0 + t) * (VST_BYTES / 4);
        mbar_expect_tx(barKV + t % 6, 2 * VST_BYTES);
        bulk_g2s(dst, VH + src, VST_BYTES, barKV + t % 6);
        bulk_g2s(dst + VST_BYTES, VC + src, VST_BYTES, barKV + t % 6);
      };
      auto load_raw = [&](int t) {
        if (t < nst) {
          const int slot = t % NRAW;
          mbar_expect_tx(barR + slot, Phi2Smem::RAW_SLOT);             // out-of-range box elements are zero-filled and counted
          if (tiled) tma_load_4d(sm + Phi2Smem::RAW + slot * Phi2Smem::RAW_SLOT, &tmD2, 0, 0, s0 + t, blockIdx.x, barR + slot, policy);
          else tma_load_2d(sm + Phi2Smem::RAW + slot * Phi2Smem::RAW_SLOT, &tmD2, (s0 + t) * PK2, r0, barR + slot, policy);
        }
      };
      for (int t = 0; t < NRAW; ++t) load_raw(t);
      load_v(0);
      if (nst > 1) load_v(1);
      for (int t = 0; t < nst; ++t) {
        if (t + 2 < nst) {
          if (t >= 1) mbar_wait(barM + (t - 1) % 3, ((t - 1) / 3) & 1);      // V(t+2) reuses the slot of V(t-1)
          load_v(t + 2);
        }
        if (t + NRAW < nst) {
          mbar_wait(barKV + t % 6, (t / 6) & 1);                       // raw(t) consumed by every worker warp
          load_raw(t + NRAW);
        }
      }
    }
  } else if (warp == NWARP) {
    // ---------------- MMA warp: all 32 lanes run the loop, one elected lane issues (see umma_tf32_ts_w)
    if (nst > 0) {
      const uint64_t dV = smem_desc(smem_u32(sm + Phi2Smem::V), B_LBO, SBO);
      constexpr uint64_t BK = (2 * B_LBO) >> 4;
      mbar_wait(barKV + 0, 0);
      for (int t = 0; t < nst; ++t) {
        tc_fence_after();
        const int c = t / ACC_G;                                       // chunk of this stage and its accumulator
        const uint32_t acc = tmem + ((c & 1) ? ACC_B : 0u);
        const bool first = (t % ACC_G) == 0;
        if (first && c >= 2) {                                         // chunk c-2 (same accumulator) has been drained
          mbar_wait(barD + (c & 1), ((c >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t ah = tmem + 128 + (t % 3) * 64, ac = ah + 32;         // K_hi (tf32) / [bf16 K | bf16 K_lo] of this stage in TMEM
        const uint64_t bh = dV + (uint64_t)((t % 3) * ((2 * VST_BYTES) >> 4)), bc = bh + (VST_BYTES >> 4);
#pragma unroll
        for (int ks = 0; ks < PK2 / 8; ++ks) umma_tf32_ts_w(acc, ah + ks * 8, bh + ks * BK, idesc, (!first || ks > 0) ? 1u : 0u);
        // the wait for the NEXT stage is taken while the MMAs of this one are queued on the tensor core
        if (t + 1 < nst) mbar_wait(barKV + (t + 1) % 6, ((t + 1) / 6) & 1);
        // corrections K_hi V_lo + K_lo V_hi: ONE bf16 product over 64 K rows = four K = 16 MMAs (8 TMEM columns / 2 chunks each)
#pragma unroll
        for (int q = 0; q < 4; ++q) umma_bf16_ts_w(acc, ac + q * 8, bc + q * BK, idesc_c, 1u);
        if ((t % ACC_G) == ACC_G - 1 || t == nst - 1) umma_commit_w(barF + (c & 1));
        umma_commit_w(barM + t % 3);
      }
    }
  } else {
    const float ngamma = -gam[1] * 1.4426950408889634f;
    const int rl = tid & (BLK - 1), qd = tid >> 7;                     // tile row, which 8 of the stage's 32 columns
    const int row = r0 + rl;
    const bool rows_full = r0 + BLK <= nr;
    const uint32_t klane = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 128 + 8 * qd;   // this thread's row (TMEM lane), its 8 K columns of slot 0
    // 128-byte swizzle of the TMA tile: 16-byte chunk c of row r sits at chunk c ^ (r % 8)
    const uint32_t roff0 = rl * 128 + (((2 * qd) ^ (rl & 7)) << 4), roff1 = rl * 128 + (((2 * qd + 1) ^ (rl & 7)) << 4);
    // this thread's share of the fp32 result: row rl, columns [28 qd, 28 qd + 28) (4 threads per row: the 4 warps of a TMEM lane quarter)
    float racc[28];
#pragma unroll
    for (int i = 0; i < 28; ++i) racc[i] = 0.f;
    const uint32_t alane = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 28 * qd;
    int next_drain = 0;
    auto drain = [&](int c) {                                          // chunk c's accumulator -> registers (round-to-nearest adds)
      mbar_wait(barF + (c & 1), (c >> 1) & 1);
      tc_fence_after();
      const uint32_t a0 = alane + ((c & 1) ? ACC_B : 0u);
#pragma unroll
      for (int i = 0; i < 7; ++i) tmem_ld4_add(a0 + 4 * i, racc + 4 * i);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(barD + (c & 1));
    };
    for (int t = 0; t < nst; ++t) {
      if ((next_drain + 1) * ACC_G + 1 <= t) drain(next_drain++);      // one stage after the chunk closed: its MMAs have retired
      mbar_wait(barR + (t % NRAW), (t / NRAW) & 1);                    // d2 tile of this stage has landed
      const unsigned char* raw = sm + Phi2Smem::RAW + (t % NRAW) * Phi2Smem::RAW_SLOT;
      const float4 dv0 = *reinterpret_cast<const float4*>(raw + roff0), dv1 = *reinterpret_cast<const float4*>(raw + roff1);
      float kv[8] = {ex2(ngamma * dv0.x), ex2(ngamma * dv0.y), ex2(ngamma * dv0.z), ex2(ngamma * dv0.w),
                     ex2(ngamma * dv1.x), ex2(ngamma * dv1.y), ex2(ngamma * dv1.z), ex2(ngamma * dv1.w)};
      if (!(rows_full && (s0 + t) * PK2 + PK2 <= nc)) {                // ragged edge: zero the entries outside the matrix
        const int j0 = (s0 + t) * PK2 + 8 * qd;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!(row < nr && j0 + e < nc)) kv[e] = 0.f;
      }
      float h[8], l[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_tf32(kv[e], h[e], l[e]);
      // correction operand: bf16 pairs (low half = lower k) of K at columns 32 + 4 qd .., of K_lo at 48 + 4 qd ..
      uint32_t ck[4], cl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(kv[2 * e], kv[2 * e + 1]), b = __floats2bfloat162_rn(l[2 * e], l[2 * e + 1]);
        ck[e] = *reinterpret_cast<const unsigned int*>(&a);
        cl[e] = *reinterpret_cast<const unsigned int*>(&b);
      }
      if (t >= 3) {                                                    // K slot t % 3 was read by stage t-3's MMAs
        mbar_wait(barM + t % 3, ((t - 3) / 3) & 1);
        tc_fence_after();
      }
      tmem_st8(klane + (t % 3) * 64, h);
      tmem_st4u(klane - 8 * qd + (t % 3) * 64 + 32 + 4 * qd, ck);
      tmem_st4u(klane - 8 * qd + (t % 3) * 64 + 48 + 4 * qd, cl);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(barKV + t % 6);
    }
    // ---------------- epilogue: drain the remaining chunks, then every thread parks its 28 columns of the partial tile in shared
    // memory (the drained d2 / V pipeline buffers: all MMAs have retired, hence every warp has consumed its last raw tile) for the
    // cluster-wide split-K reduction below.
    float* acc_s = reinterpret_cast<float*>(sm);                       // [BLK][ACC_PITCH]
    if (nst > 0) {
      const int nchunks = (nst + ACC_G - 1) / ACC_G;
      while (next_drain < nchunks) drain(next_drain++);
    }
#pragma unroll
    for (int i = 0; i < 28; ++i)
      if (28 * qd + i <= 2 * d) acc_s[rl * ACC_PITCH + 28 * qd + i] = racc[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
  // ---------------- fused combine (stein.py:84-86).  The jsplit CTAs of a row block form one thread-block cluster; after the
  // cluster barrier CTA y sums rows [y BLK / jsplit, (y + 1) BLK / jsplit) of all jsplit partial tiles through distributed shared
  // memory in a fixed order, forms phi = (KS + 2 gamma (rowsum (x_i - mu) - K(X - mu))) / n and applies theta += step phi.
  // No partials in global memory, no second launch.
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  {
    const float* acc_s = reinterpret_cast<const float*>(sm);
    const int y = blockIdx.y;
    const int rbeg = (y * BLK) / jsplit, rend = ((y + 1) * BLK) / jsplit;
    const float g2 = 2.f * gam[1];
    const float* rem[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) rem[q] = cl.map_shared_rank(acc_s, q < jsplit ? q : 0);
    const int tot = (rend - rbeg) * d;
    for (int idx = tid; idx < tot; idx += blockDim.x) {
      const int rl = rbeg + idx / d, c = idx % d;
      const long long r = r0 + rl;
      if (r >= nr) break;
      float ks = 0.f, kx = 0.f, rs = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < jsplit) {
          const float* p = rem[q] + rl * ACC_PITCH;
          ks += p[c];
          kx += p[d + c];
          rs += p[2 * d];
        }
      const float x = cmb.Xr[r * cmb.ldr + c];
      const float ph = (ks + g2 * (rs * (x - cmb.mu[c]) - kx)) * cmb.inv_n;
      if (cmb.phi) cmb.phi[r * cmb.ldp + c] = ph;
      if (cmb.theta) cmb.theta[r * cmb.ldt + c] = fmaf(cmb.step, ph, cmb.theta[r * cmb.ldt + c]);
    }
  }
  cl.sync();                                                           // nobody leaves while its tile is still being read
}

// ---------------------------------------------------------------- host launchers (called from svgd.cu)
size_t svgd_tc2_operand_bytes(int nr, int nc) {
  const size_t nrp = (size_t)(nr + BLK - 1) / BLK * BLK, ncp = (size_t)(nc + BLK - 1) / BLK * BLK;
  size_t b = 0;
  b += 2 * nrp * KP2 * 4 + nrp * 4;             // row operands hi/lo + norms
  b += 2 * ncp * KP2 * 4 + ncp * 4;             // column operands hi/lo + norms
  b += 2 * ncp * NF2 * 4;                       // V^T hi/lo
  b += 2 * (size_t)(WIN_TABLE + 1) * 8 + 512;   // window table + below counter, scratch copy + segment totals + ticket
  return b + 1024;
}

struct Tc2Ops {
  float *XrH, *XrL, *nrm_r, *XcH, *XcL, *nrm_c, *VH, *VC;
  unsigned long long* table;
};
Tc2Ops svgd_tc2_carve(void* base, int nr, int nc) {
  const size_t nrp = (size_t)(nr + BLK - 1) / BLK * BLK, ncp = (size_t)(nc + BLK - 1) / BLK * BLK;
  Tc2Ops o;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) / 256 * 256; return q; };
  o.XrH = (float*)take(nrp * KP2 * 4); o.XrL = (float*)take(nrp * KP2 * 4); o.nrm_r = (float*)take(nrp * 4);
  o.XcH = (float*)take(ncp * KP2 * 4); o.XcL = (float*)take(ncp * KP2 * 4); o.nrm_c = (float*)take(ncp * 4);
  o.VH = (float*)take(ncp * NF2 * 4); o.VC = (float*)take(ncp * NF2 * 4);
  o.table = (unsigned long long*)take(2 * (size_t)(WIN_TABLE + 1) * 8 + 512);
  return o;
}
size_t svgd_tc2_carved_bytes(int nr, int nc) {
  const size_t nrp = (size_t)(nr + BLK - 1) / BLK * BLK, ncp = (size_t)(nc + BLK - 1) / BLK * BLK;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  return 2 * al(nrp * KP2 * 4) + al(nrp * 4) + 2 * al(ncp * KP2 * 4) + al(ncp * 4) + 2 * al(ncp * NF2 * 4) + al(2 * (size_t)(WIN_TABLE + 1) * 8 + 512);
}

int svgd_tc2_supported(int d, int nc) { return d >= 1 && d <= 55 && (nc & 3) == 0 && nc >= 4; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Row-major d2[nr][nc] as a 2-D tensor: box = bx columns x by rows
static int encode_d2_map(CUtensorMap* tm, const float* D2, int nr, int nc, int bx, int by, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = encode_tiled_fn();
  BODE_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)nc, (cuuint64_t)nr};
  const cuuint64_t gstr[1] = {(cuuint64_t)nc * 4};
  const cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)D2, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BODE_REQUIRE(cr == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
  return BODE_OK;
}

// d2 layout inside the workspace.  Row-major [nr][nc] in general; when both edges are whole 128 x 128 Gram tiles the block is stored as
// [nr / 128][nc / 32][128][32], i.e. every 16 KB tile a K@V stage consumes is CONTIGUOUS.  A stage tile of the row-major matrix
// is 128 segments of 128 bytes, one per DRAM page: once the block has left L2 (several ranks: 537 MB per rank at 8 x 4096
// particles) those reads ran at 3.2 TB/s.  The order statistics do not care about the order of the entries.
int svgd_tc2_d2_tiled(int nr, int nc) { return (nr % BLK == 0 && nc % BLK == 0) ? 1 : 0; }
// the d2 block is larger than what L2 can keep between the Gram pass and the K@V pass (126 MB, shared with the solve's checkpoints)
static int d2_streams(int nr, int nc) { return (size_t)nr * nc * 4 > ((size_t)96 << 20) ? 1 : 0; }
static int encode_d2_map_tiled(CUtensorMap* tm, const float* D2, int nr, int nc, int bx, int by, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = encode_tiled_fn();
  BODE_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t nst = (cuuint64_t)(nc / PK2);
  const cuuint64_t gdim[4] = {(cuuint64_t)PK2, (cuuint64_t)BLK, nst, (cuuint64_t)(nr / BLK)};
  const cuuint64_t gstr[3] = {(cuuint64_t)PK2 * 4, (cuuint64_t)BLK * PK2 * 4, (cuuint64_t)BLK * PK2 * 4 * nst};
  const cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)D2, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BODE_REQUIRE(cr == CUDA_SUCCESS, "cuTensorMapEncodeTiled (tiled d2) failed (%d)", (int)cr);
  return BODE_OK;
}

static int g_gram_split = 0;   // 0 = one wave over all SMs; > 0: column splits per row block (finer CTAs when the Gram pass shares the GPU)
int svgd_tc2_set_gram_split(int js) {
  const int old = g_gram_split;
  g_gram_split = js > 0 ? js : 0;
  return old;
}

static int split_for(int blocks_y, int units, int sms) {
  int js = sms / (blocks_y > 0 ? blocks_y : 1);
  if (js < 1) js = 1;
  if (js > 16) js = 16;
  if (js > units) js = units;
  return js;
}

// Column tiles per Gram CTA (one CTA per SM is resident).  The grid is (column chunks, row blocks); its cost is
// waves x (tiles per CTA + a fixed prologue of about one tile: TMEM allocation, A tile, pipeline fill).  One chunk count per
// row block rarely fills the last wave (32 row blocks x 4 chunks = 128 CTAs on 148 SMs: 64 tile times for 8192 tiles, where
// 9 chunks of 29 tiles take 2 x 29 = 58), so the choice is made by enumeration.
static int gram_tiles_per_cta(int nrb, int nct, int sms) {
  int best = nct;
  long long best_cost = -1;
  for (int tp = 1; tp <= nct; ++tp) {
    const long long chunks = (nct + tp - 1) / tp;
    const long long waves = (chunks * nrb + sms - 1) / sms;
    const long long cost = waves * (tp + 1);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = tp;
    }
  }
  return best;
}

// stages: bit 0 = operand preparation (pre-split, centred column operands; needs only the positions), bit 1 = the Gram kernel
int svgd_tc2_gram(const float* Xr, long long ldr, int nr, int row_offset, const float* Xc, long long ldc, int nc, int d, float* mu,
                  void* ops_base, float* D2, SelState* st, unsigned long long total, int sms, int stages, cudaStream_t stream) {
  const Tc2Ops o = svgd_tc2_carve(ops_base, nr, nc);
  const int nrp = (nr + BLK - 1) / BLK * BLK, ncp = (nc + BLK - 1) / BLK * BLK;
  const float *rH = o.XrH, *rL = o.XrL, *rN = o.nrm_r;
  // the local rows usually ARE a 128-aligned block of the gathered columns: reuse the column operands
  const bool alias = row_offset >= 0 && (row_offset % BLK) == 0 && ldr == ldc && Xr == Xc + (long long)row_offset * ldc &&
                     (row_offset + nrp <= ncp) && (nr % BLK == 0 || row_offset + nr == nc);
  if (alias) {
    rH = o.XcH + (long long)(row_offset / BLK) * (BLK_BYTES / 4);
    rL = o.XcL + (long long)(row_offset / BLK) * (BLK_BYTES / 4);
    rN = o.nrm_c + row_offset;
  }
  if (stages & 1) {
    prep_x_kernel<<<ncp / BLK, 256, 0, stream>>>(Xc, ldc, nc, d, Xc, ldc, nc, mu, o.XcH, o.XcL, o.nrm_c, st, total);
    if (!alias) prep_x_kernel<<<nrp / BLK, 256, 0, stream>>>(Xr, ldr, nr, d, Xc, ldc, nc, mu, o.XrH, o.XrL, o.nrm_r, nullptr, 0ull);
    BODE_CUDA(cudaGetLastError());
  }
  if (!(stages & 2)) return BODE_OK;
  static bool attr_set = false;
  if (!attr_set) {
    BODE_CUDA(cudaFuncSetAttribute(gram2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gram2Smem::TOTAL));
    attr_set = true;
  }
  const int nrb = nrp / BLK, nct = ncp / BLK;
  int tiles_per;
  if (g_gram_split > 0) {
    const int js = g_gram_split > nct ? nct : g_gram_split;
    tiles_per = (nct + js - 1) / js;
  } else {
    tiles_per = gram_tiles_per_cta(nrb, nct, sms);
  }
  dim3 grid((nct + tiles_per - 1) / tiles_per, nrb);
  CUtensorMap tm;                                                     // d2 tile stores: 16 columns x 32 rows per warp, 64-byte swizzle
  const int tiled = svgd_tc2_d2_tiled(nr, nc);
  if (int rc = tiled ? encode_d2_map_tiled(&tm, D2, nr, nc, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B)
                     : encode_d2_map(&tm, D2, nr, nc, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B))
    return rc;
  gram2_kernel<<<grid, NTHR_PHI, Gram2Smem::TOTAL, stream>>>(tm, rH, rL, rN, nr, row_offset, o.XcH, o.XcL, o.nrm_c, nc, tiles_per, st, o.table,
                                                             tiled, d2_streams(nr, nc));
  return check_cuda(cudaGetLastError(), "gram2 launch");
}

int svgd_tc2_window_select(SelState* st, void* ops_base, int nr, int nc, const PeerInfo& peer, int one_cta, cudaStream_t stream) {
  const Tc2Ops o = svgd_tc2_carve(ops_base, nr, nc);
  if (one_cta && peer.world <= 1) {                      // small cluster: schedulable beside a kernel that fills most SMs
    cudaLaunchConfig_t cfg4 = {};
    cfg4.gridDim = dim3(WSEL4);
    cfg4.blockDim = dim3(1024);
    cfg4.dynamicSmemBytes = 0;
    cfg4.stream = stream;
    cudaLaunchAttribute at4[1];
    at4[0].id = cudaLaunchAttributeClusterDimension;
    at4[0].val.clusterDim.x = WSEL4;
    at4[0].val.clusterDim.y = 1;
    at4[0].val.clusterDim.z = 1;
    cfg4.attrs = at4;
    cfg4.numAttrs = 1;
    BODE_CUDA(cudaLaunchKernelEx(&cfg4, window_select4_kernel, st, o.table));
    return BODE_OK;
  }
  static bool attr_set = false;
  if (!attr_set) {
    BODE_CUDA(cudaFuncSetAttribute(window_select_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(WSEL_CLUSTER);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = WSEL_CLUSTER;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  BODE_CUDA(cudaLaunchKernelEx(&cfg, window_select_kernel, st, o.table, peer));
  return BODE_OK;
}

unsigned long long* svgd_tc2_table(void* ops_base, int nr, int nc) { return svgd_tc2_carve(ops_base, nr, nc).table; }
void svgd_tc2_v_tiles(void* ops_base, int nr, int nc, float** VH, float** VC) {
  const Tc2Ops o = svgd_tc2_carve(ops_base, nr, nc);
  *VH = o.VH;
  *VC = o.VC;
}
static_assert(NF2 == SV_NF && PK2 == SV_PK && VST_BYTES == SV_VST_BYTES, "svgd_tiles.cuh must describe the layout prep_v_kernel writes");

int svgd_tc2_phi(const float* D2, int nr, int nc, const float* Xc, long long ldx, const float* Gc, long long ldg, int d, const float* mu,
                 const float* gam, float gsign, void* ops_base, int* jsplit_out, float* part, int sms, int stages, const float* Xr,
                 long long ldr, float inv_n, float* phi, long long ldp, float* theta, long long ldt, float step, cudaStream_t stream) {
  const Tc2Ops o = svgd_tc2_carve(ops_base, nr, nc);
  const int ncp = (nc + BLK - 1) / BLK * BLK;
  const int nrb = (nr + BLK - 1) / BLK, nst = (nc + PK2 - 1) / PK2;
  int js = split_for(nrb, nst, sms);
  if (js > 8) js = 8;                                                  // portable cluster size
  *jsplit_out = js;
  if (stages & 5) {   // V^T = [-G | X - mu | 1] operand tiles: needs positions and scores, not d2 or gamma; bit 2 alone: without the scores
    prep_v_kernel<<<(int)(((long long)(ncp / 4) * NF2 + 255) / 256), 256, 0, stream>>>(Xc, ldx, Gc, ldg, nc, d, mu, gsign, ncp, o.VH, o.VC,
                                                                                          (stages & 1) ? 0 : d);
    BODE_CUDA(cudaGetLastError());
  }
  if (!(stages & 2)) return BODE_OK;
  static bool attr_set = false;
  if (!attr_set) {
    BODE_CUDA(cudaFuncSetAttribute(phi2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Phi2Smem::TOTAL));
    attr_set = true;
  }
  // TMA descriptor of the row-major d2[nr][nc] matrix: box = 32 columns x 128 rows, 128-byte swizzle, zero fill outside
  CUtensorMap tm;
  const int tiled = svgd_tc2_d2_tiled(nr, nc);
  if (int rc = tiled ? encode_d2_map_tiled(&tm, D2, nr, nc, PK2, BLK, CU_TENSOR_MAP_SWIZZLE_128B)
                     : encode_d2_map(&tm, D2, nr, nc, PK2, BLK, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  dim3 grid(nrb, js);
  Phi2Combine cmb;
  cmb.Xr = Xr; cmb.ldr = ldr; cmb.mu = mu; cmb.inv_n = inv_n; cmb.phi = phi; cmb.ldp = ldp; cmb.theta = theta; cmb.ldt = ldt; cmb.step = step;
  static_assert((size_t)BLK * ACC_PITCH * 4 <= Phi2Smem::BARS, "the partial tile must fit the drained pipeline buffers");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NTHR_PHI);
  cfg.dynamicSmemBytes = Phi2Smem::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;                      // the split-K CTAs of a row block share one cluster
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = js;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  BODE_CUDA(cudaLaunchKernelEx(&cfg, phi2_kernel, tm, nr, nc, (const float*)o.VH, (const float*)o.VC, d, gam, js, part, cmb, tiled, d2_streams(nr, nc)));
  return check_cuda(cudaGetLastError(), "phi2 launch");
}

}  // namespace bode
