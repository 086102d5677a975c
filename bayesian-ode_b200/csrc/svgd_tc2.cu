// Pipelined SVGD contractions on tcgen05 (3xTF32, TMEM accumulators) for shapes whose column count is a multiple of 4.
//
//   prep_x_kernel / prep_v_kernel   centre, split (hi = top 19 bits, lo = x - hi) and store the operands ONCE per step in global
//                                   memory, already in the canonical no-swizzle K-major UMMA layout, so that every later
//                                   operand tile is one contiguous block moved by a 1-D bulk copy (TMA engine, UBLKCP).
//   gram2_kernel                    d2 = |xc_i|^2 + |xc_j|^2 - 2 xc_i.xc_j   (stein.py:22).  One CTA keeps its 128-row A tile
//                                   resident and walks 128-column B tiles through a 2-slot ring; the accumulator is
//                                   double-buffered in TMEM (2 x 128 columns) so tile t+1's MMAs run under tile t's epilogue.
//                                   The epilogue also counts the window of the exact median selection (svgd_state.cuh).
//   phi2_kernel                     part[i,:] = sum_j 2^(-g d2_ij) [ -grad_j | xc_j | 1 ]   (stein.py:75-86).  d2 tiles arrive
//                                   by cp.async three stages ahead, V^T tiles by bulk copy, the exp/split of stage t runs
//                                   under the MMAs of stage t-1.
// A seventeenth "control" warp issues every bulk copy and MMA; warps 0-15 (four per SM sub-partition) own the TMEM lanes /
// operand generation.
#include "tc_ptx.cuh"
#include "svgd_state.cuh"

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
constexpr int NTHR = NWORK + 32;

// ---------------------------------------------------------------- operand preparation
// XH/XL[blk][kc][r/8][r%8][4] : K-major core matrices of the centred rows, zero padded to n_pad rows (multiple of 128)
__global__ void __launch_bounds__(256) prep_x_kernel(const float* __restrict__ X, long long ld, int n, int d, const float* __restrict__ mu,
                                                     int n_pad, float* __restrict__ XH, float* __restrict__ XL, float* __restrict__ norms) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long tot = (long long)n_pad * KCH2;
  if (idx < tot) {
    const int r = (int)(idx % BLK);
    const long long q = idx / BLK;
    const int kc = (int)(q % KCH2);
    const long long blk = q / KCH2;
    const long long row = blk * BLK + r;
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * kc + e;
      const float v = (row < n && k < d) ? __ldg(X + row * ld + k) - __ldg(mu + k) : 0.f;
      split_tf32(v, h[e], l[e]);
    }
    const long long off = blk * (BLK_BYTES / 4) + (long long)(kc * (BLK / 8) + (r >> 3)) * 32 + (r & 7) * 4;
    *reinterpret_cast<float4*>(XH + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(XL + off) = make_float4(l[0], l[1], l[2], l[3]);
  }
  if (idx < n_pad) {                       // |xc|^2, summed in feature order
    float s = 0.f;
    if (idx < n)
      for (int k = 0; k < d; ++k) {
        const float v = __ldg(X + idx * ld + k) - __ldg(mu + k);
        s = fmaf(v, v, s);
      }
    norms[idx] = s;
  }
}

// VH/VL[j/4][f][j%4] : K-major (K = particle index j) core matrices of V^T, zero padded to n_pad particles (multiple of 32)
__global__ void __launch_bounds__(256) prep_v_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ G, long long ldg,
                                                     int n, int d, const float* __restrict__ mu, float gsign, int n_pad,
                                                     float* __restrict__ VH, float* __restrict__ VL) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(n_pad / 4) * NF2) return;
  const int f = (int)(idx % NF2);
  const long long jq = idx / NF2;
  float h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const long long j = 4 * jq + e;
    float v = 0.f;
    if (j < n) {
      if (f < d) v = gsign * __ldg(G + j * ldg + f);
      else if (f < 2 * d) v = __ldg(X + j * ldx + (f - d)) - __ldg(mu + f - d);
      else if (f == 2 * d) v = 1.f;
    }
    split_tf32(v, h[e], l[e]);
  }
  *reinterpret_cast<float4*>(VH + idx * 4) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(VL + idx * 4) = make_float4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------- Gram tiles + window count
struct Gram2Smem {
  static constexpr uint32_t A = 0;                                  // hi | lo
  static constexpr uint32_t B = 2 * BLK_BYTES;                      // 2 slots x (hi | lo)
  static constexpr uint32_t STAGE = B + 4 * BLK_BYTES;              // per warp [32][20] floats (16 columns at a time)
  static constexpr uint32_t NCS = STAGE + NWARP * 32 * 20 * 4;      // per warp 32 column norms
  static constexpr uint32_t BARS = NCS + NWARP * 32 * 4;            // barA, barB[2], barS[2]
  static constexpr uint32_t TSLOT = BARS + 5 * 8;
  static constexpr uint32_t TOTAL = TSLOT + 16;
};

__global__ void __launch_bounds__(NTHR, 1) gram2_kernel(const float* __restrict__ XrH, const float* __restrict__ XrL,
                                                        const float* __restrict__ nrm_r, int nr, int row_offset,
                                                        const float* __restrict__ XcH, const float* __restrict__ XcL,
                                                        const float* __restrict__ nrm_c, int nc, int tiles_per_cta,
                                                        float* __restrict__ D2, SelState* __restrict__ st,
                                                        unsigned long long* __restrict__ table) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Gram2Smem::BARS);
  uint64_t* barA = bars;
  uint64_t* barB = bars + 1;
  uint64_t* barS = bars + 3;
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
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  constexpr uint32_t idesc = idesc_tf32(BLK, BLK, 0, 0);
  constexpr uint32_t LBO = BLK * 16, SBO = 128;

  if (warp == NWARP) {
    // ---------------- control warp: bulk copies + MMA issue (lane 0)
    auto load_b = [&](int t) {
        const int slot = t & 1;
        unsigned char* dst = sm + Gram2Smem::B + slot * 2 * BLK_BYTES;
        const long long src = (long long)(ct0 + t) * (BLK_BYTES / 4);
        mbar_expect_tx(barB + slot, 2 * BLK_BYTES);
        bulk_g2s(dst, XcH + src, BLK_BYTES, barB + slot);
        bulk_g2s(dst + BLK_BYTES, XcL + src, BLK_BYTES, barB + slot);
      };
      auto issue = [&](int t) {
        const int slot = t & 1;
        const uint32_t ah = smem_u32(sm + Gram2Smem::A), al = ah + BLK_BYTES;
        const uint32_t bh = smem_u32(sm + Gram2Smem::B + slot * 2 * BLK_BYTES), bl = bh + BLK_BYTES;
        uint32_t acc = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a0 = pass == 2 ? al : ah, b0 = pass == 1 ? bl : bh;
#pragma unroll 1
          for (int ks = 0; ks < KP2 / 8; ++ks) {
            umma_tf32(tmem + slot * BLK, smem_desc(a0 + ks * 2 * LBO, LBO, SBO), smem_desc(b0 + ks * 2 * LBO, LBO, SBO), idesc, acc);
            acc = 1;
          }
        }
        umma_commit(barS + slot);
      };
    if (lane == 0) {
      mbar_expect_tx(barA, 2 * BLK_BYTES);
      bulk_g2s(sm + Gram2Smem::A, XrH + (long long)rb * (BLK_BYTES / 4), BLK_BYTES, barA);
      bulk_g2s(sm + Gram2Smem::A + BLK_BYTES, XrL + (long long)rb * (BLK_BYTES / 4), BLK_BYTES, barA);
      load_b(0);
      if (nt > 1) load_b(1);
      mbar_wait(barA, 0);
      mbar_wait(barB + 0, 0);
      tc_fence_after();
      issue(0);
    }
    __syncwarp();
    for (int t = 0; t < nt; ++t) {
      if (lane == 0) {
        if (t + 1 < nt) {                       // S[(t+1)&1] was released by the barrier that ended iteration t-1
          mbar_wait(barB + ((t + 1) & 1), ((t + 1) >> 1) & 1);
          tc_fence_after();
          issue(t + 1);
        }
        mbar_wait(barS + (t & 1), (t >> 1) & 1);            // tile t's MMAs are done: its B slot is free
        if (t + 2 < nt) load_b(t + 2);
      }
      __syncwarp();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  } else {
    // ---------------- worker warps: epilogue.  Warp w reads TMEM lanes 32 (w % 4) .. +31 (tile rows), columns 32 (w / 4) .. +31
    const int q = warp & 3, cq = warp >> 2;
    const int rl = 32 * q + lane;
    const int row = rb * BLK + rl;
    const float nrow = __ldg(nrm_r + row);
    float* stage = reinterpret_cast<float*>(sm + Gram2Smem::STAGE) + warp * 32 * 20;
    float* ncs = reinterpret_cast<float*>(sm + Gram2Smem::NCS) + warp * 32;
    const bool win = st->win_valid != 0;
    const unsigned int wlo = win ? st->win_lo : 0xffffffffu;     // disarmed: nothing is below, nothing is inside
    const unsigned int wspan = win ? WIN_SPAN : 0u;
    unsigned int below = 0;
    float mx = 0.f;
    const int cb = 32 * cq;                                        // first tile column of this warp
    for (int t = 0; t < nt; ++t) {
      const int c0 = (ct0 + t) * BLK;
      ncs[lane] = __ldg(nrm_c + c0 + cb + lane);                   // in flight while the MMAs of this tile finish
      mbar_wait(barS + (t & 1), (t >> 1) & 1);
      tc_fence_after();
      float s[32];
      tmem_ld32(tmem + (t & 1) * BLK + cb + ((uint32_t)(32 * q) << 16), s);
      __syncwarp();
      const bool full = rb * BLK + BLK <= nr && c0 + BLK <= nc;
      const int dcol = row + row_offset - (c0 + cb);               // tile-local column of the diagonal entry (if 0..31)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 nj = *reinterpret_cast<const float4*>(ncs + 16 * hf + 4 * k4);
          const float njv[4] = {nj.x, nj.y, nj.z, nj.w};
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 16 * hf + 4 * k4 + e;
            float v = fmaxf(fmaf(-2.f, s[c], nrow + njv[e]), 0.f);
            if (c == dcol) v = 0.f;                                // cdist(x, x) = 0 on the diagonal
            o[e] = v;
            if (full || (row < nr && c0 + cb + c < nc)) {
              mx = fmaxf(mx, v);
              const unsigned int off = __float_as_uint(v) - wlo;   // wraps for entries below the window
              below += (int)off < 0;
              if (off <= wspan) atomicAdd(table + off, 1ull);
            }
          }
          *reinterpret_cast<float4*>(stage + lane * 20 + 4 * k4) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        // coalesced rows: 4 lanes cover the 64 bytes of one row of the 16-column half
#pragma unroll
        for (int rr = 0; rr < 32; rr += 8) {
          const int r2 = rr + (lane >> 2), c4 = lane & 3;
          const int grow = rb * BLK + 32 * q + r2, gcol = c0 + cb + 16 * hf + 4 * c4;
          if (full || (grow < nr && gcol < nc))
            *reinterpret_cast<float4*>(D2 + (long long)grow * nc + gcol) = *reinterpret_cast<const float4*>(stage + r2 * 20 + 4 * c4);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      below += __shfl_xor_sync(0xffffffffu, below, o);
    }
    if (lane == 0) {
      atomicMax(&st->maxbits, __float_as_uint(mx));
      if (win && below) atomicAdd(table + WIN_TABLE, (unsigned long long)below);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------- median from the window table (one CTA)
__global__ void __launch_bounds__(1024) window_select_kernel(SelState* st, unsigned long long* table) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned int found[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int PER = (WIN_TABLE + 1023) / 1024;             // bins per thread (contiguous)
  const bool armed = st->win_valid != 0;
  unsigned long long sum = 0;
  for (int i = 0; i < PER; ++i) {
    const int b = tid * PER + i;
    if (armed && b < (int)WIN_TABLE) sum += table[b];
  }
  unsigned long long incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[wid] = incl;
  if (tid < 2) found[tid] = 0xffffffffu;
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
  const unsigned long long below = armed ? table[WIN_TABLE] : 0ull;
  unsigned long long excl = below + incl - sum + (wid ? wsum[wid - 1] : 0ull);     // entries strictly below this thread's first bin
  const unsigned long long r0 = st->rank[0], r1 = st->rank[1];
  if (sum) {
    for (int i = 0; i < PER; ++i) {
      const int b = tid * PER + i;
      const unsigned long long c = b < (int)WIN_TABLE ? table[b] : 0ull;
      if (c) {
        if (r0 >= excl && r0 < excl + c) found[0] = b;
        if (r1 >= excl && r1 < excl + c) found[1] = b;
      }
      excl += c;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const bool hit = armed && found[0] != 0xffffffffu && found[1] != 0xffffffffu;
    if (hit) {
      st->prefix[0] = st->win_lo + found[0];
      st->prefix[1] = st->win_lo + found[1];
    }
    st->hit = hit ? 1u : 0u;
  }
  for (int b = tid; b <= (int)WIN_TABLE; b += 1024) table[b] = 0ull;      // ready for the next call
}

// ---------------------------------------------------------------- phi partials
struct Phi2Smem {
  static constexpr uint32_t RAW = 0;                                   // 3 slots x [128][36] floats (d2 tile, row-major)
  static constexpr uint32_t RAW_SLOT = BLK * 36 * 4;
  static constexpr uint32_t K = RAW + 3 * RAW_SLOT;                    // 2 slots x (hi | lo) x [8 kc][16][8][4]
  static constexpr uint32_t K_HALF = BLK * PK2 * 4;
  static constexpr uint32_t V = K + 4 * K_HALF;                        // 3 slots x (hi | lo)
  static constexpr uint32_t BARS = V + 6 * VST_BYTES;                  // barM[2], barV[3]
  static constexpr uint32_t TSLOT = BARS + 5 * 8;
  static constexpr uint32_t TOTAL = TSLOT + 16;
};

__global__ void __launch_bounds__(NTHR, 1) phi2_kernel(const float* __restrict__ D2, int nr, int nc, const float* __restrict__ VH,
                                                       const float* __restrict__ VL, int d, const float* __restrict__ gam, int jsplit,
                                                       float* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Phi2Smem::BARS);
  uint64_t* barM = bars;
  uint64_t* barV = bars + 2;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + Phi2Smem::TSLOT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * BLK;
  const int nst_all = (nc + PK2 - 1) / PK2;
  const int per = (nst_all + jsplit - 1) / jsplit;
  const int s0 = blockIdx.y * per;
  const int nst = min(per, nst_all - s0);                              // stages of this CTA (may be <= 0)

  if (warp == 0) tmem_alloc(tslot, 128);
  if (tid == 0) {
    mbar_init(barM + 0, 1);
    mbar_init(barM + 1, 1);
    mbar_init(barV + 0, 1);
    mbar_init(barV + 1, 1);
    mbar_init(barV + 2, 1);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  constexpr uint32_t idesc = idesc_tf32(BLK, NF2, 0, 0);
  constexpr uint32_t A_LBO = BLK * 16, B_LBO = NF2 * 16, SBO = 128;

  if (warp == NWARP) {
    auto load_v = [&](int t) {
      const int slot = t % 3;
      unsigned char* dst = sm + Phi2Smem::V + slot * 2 * VST_BYTES;
      const long long src = (long long)(s0 + t) * (VST_BYTES / 4);
      mbar_expect_tx(barV + slot, 2 * VST_BYTES);
      bulk_g2s(dst, VH + src, VST_BYTES, barV + slot);
      bulk_g2s(dst + VST_BYTES, VL + src, VST_BYTES, barV + slot);
    };
    if (lane == 0 && nst > 0) {
      load_v(0);
      if (nst > 1) load_v(1);
    }
    __syncwarp();
    for (int t = 0; t < nst; ++t) {
      tc_fence_before();
      __syncthreads();                                                 // K(t) is in shared memory
      tc_fence_after();
      if (lane == 0) {
        mbar_wait(barV + (t % 3), (t / 3) & 1);
        const int ks_ = t & 1;
        const uint32_t ah = smem_u32(sm + Phi2Smem::K + ks_ * 2 * Phi2Smem::K_HALF), al = ah + Phi2Smem::K_HALF;
        const uint32_t bh = smem_u32(sm + Phi2Smem::V + (t % 3) * 2 * VST_BYTES), bl = bh + VST_BYTES;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a0 = pass == 2 ? al : ah, b0 = pass == 1 ? bl : bh;
#pragma unroll
          for (int ks = 0; ks < PK2 / 8; ++ks)
            umma_tf32(tmem, smem_desc(a0 + ks * 2 * A_LBO, A_LBO, SBO), smem_desc(b0 + ks * 2 * B_LBO, B_LBO, SBO), idesc,
                      (t > 0 || pass > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(barM + (t & 1));
        // V(t+2) reuses the slot of V(t-1): wait for stage t-1's MMAs
        if (t + 2 < nst) {
          if (t >= 1) mbar_wait(barM + ((t - 1) & 1), ((t - 1) >> 1) & 1);
          load_v(t + 2);
        }
      }
      __syncwarp();
    }
  } else {
    const float ngamma = -gam[1] * 1.4426950408889634f;
    const int rl = tid & (BLK - 1), qd = tid >> 7;                     // tile row, which 8 of the stage's 32 columns
    const int row = r0 + rl;
    const bool rows_full = r0 + BLK <= nr;
    auto load_raw = [&](int t) {                                       // 128 x 32 floats of d2, 2 x 16 B per thread, coalesced
      if (t < nst) {
        float* dst = reinterpret_cast<float*>(sm + Phi2Smem::RAW + (t % 3) * Phi2Smem::RAW_SLOT);
        const int j0 = (s0 + t) * PK2;
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          const int idx = tid + NWORK * qq, rr = idx >> 3, c4 = idx & 7;
          const int gr = min(r0 + rr, nr - 1), gc = min(j0 + 4 * c4, nc - 4);
          cp_async16(dst + rr * 36 + 4 * c4, D2 + (long long)gr * nc + gc);
        }
      }
      cp_async_commit();
    };
    load_raw(0);
    load_raw(1);
    for (int t = 0; t < nst; ++t) {
      load_raw(t + 2);
      cp_async_wait<2>();
      named_bar_sync(1, NWORK);                                        // raw(t) complete for every worker
      if (t >= 2) {                                                    // K slot t&1 was read by stage t-2's MMAs
        mbar_wait(barM + (t & 1), ((t - 2) >> 1) & 1);
        tc_fence_after();
      }
      const float* raw = reinterpret_cast<const float*>(sm + Phi2Smem::RAW + (t % 3) * Phi2Smem::RAW_SLOT) + rl * 36 + 8 * qd;
      unsigned char* kh = sm + Phi2Smem::K + (t & 1) * 2 * Phi2Smem::K_HALF;
      unsigned char* kl = kh + Phi2Smem::K_HALF;
      const int j0 = (s0 + t) * PK2 + 8 * qd;
      const bool full = rows_full && (s0 + t) * PK2 + PK2 <= nc;
#pragma unroll
      for (int kq = 0; kq < 2; ++kq) {
        const float4 dv = *reinterpret_cast<const float4*>(raw + 4 * kq);
        const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float kv = ex2(ngamma * dd[e]);
          if (!full && !(row < nr && j0 + 4 * kq + e < nc)) kv = 0.f;
          split_tf32(kv, h[e], l[e]);
        }
        const uint32_t off = kmajor_off<BLK>(rl, 2 * qd + kq);
        *reinterpret_cast<float4*>(kh + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(kl + off) = make_float4(l[0], l[1], l[2], l[3]);
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
    cp_async_wait<0>();
    // ---------------- epilogue: one TMEM lane per thread = one output row; warps 4..7 take the upper feature half
    const int half = qd;
    if (nst > 0 && warp < 8) {
      mbar_wait(barM + ((nst - 1) & 1), ((nst - 1) >> 1) & 1);
      tc_fence_after();
      const int cbeg = half * (NF2 / 2);
#pragma unroll 1
      for (int cb = cbeg; cb < cbeg + NF2 / 2; cb += 8) {
        float s[8];
        tmem_ld8(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cb, s);
        if (row < nr) {
          float* dst = part + ((long long)blockIdx.y * nr + row) * (2 * d + 1);
#pragma unroll
          for (int qq = 0; qq < 8; ++qq)
            if (cb + qq <= 2 * d) dst[cb + qq] = s[qq];
        }
      }
    } else if (nst <= 0 && row < nr && half == 0) {
      float* dst = part + ((long long)blockIdx.y * nr + row) * (2 * d + 1);
      for (int f = 0; f <= 2 * d; ++f) dst[f] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------- host launchers (called from svgd.cu)
size_t svgd_tc2_operand_bytes(int nr, int nc) {
  const size_t nrp = (size_t)(nr + BLK - 1) / BLK * BLK, ncp = (size_t)(nc + BLK - 1) / BLK * BLK;
  size_t b = 0;
  b += 2 * nrp * KP2 * 4 + nrp * 4;             // row operands hi/lo + norms
  b += 2 * ncp * KP2 * 4 + ncp * 4;             // column operands hi/lo + norms
  b += 2 * ncp * NF2 * 4;                       // V^T hi/lo
  b += (size_t)(WIN_TABLE + 1) * 8;             // window table + below counter
  return b + 1024;
}

struct Tc2Ops {
  float *XrH, *XrL, *nrm_r, *XcH, *XcL, *nrm_c, *VH, *VL;
  unsigned long long* table;
};
Tc2Ops svgd_tc2_carve(void* base, int nr, int nc) {
  const size_t nrp = (size_t)(nr + BLK - 1) / BLK * BLK, ncp = (size_t)(nc + BLK - 1) / BLK * BLK;
  Tc2Ops o;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) / 256 * 256; return q; };
  o.XrH = (float*)take(nrp * KP2 * 4); o.XrL = (float*)take(nrp * KP2 * 4); o.nrm_r = (float*)take(nrp * 4);
  o.XcH = (float*)take(ncp * KP2 * 4); o.XcL = (float*)take(ncp * KP2 * 4); o.nrm_c = (float*)take(ncp * 4);
  o.VH = (float*)take(ncp * NF2 * 4); o.VL = (float*)take(ncp * NF2 * 4);
  o.table = (unsigned long long*)take((size_t)(WIN_TABLE + 1) * 8);
  return o;
}
size_t svgd_tc2_carved_bytes(int nr, int nc) {
  const size_t nrp = (size_t)(nr + BLK - 1) / BLK * BLK, ncp = (size_t)(nc + BLK - 1) / BLK * BLK;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  return 2 * al(nrp * KP2 * 4) + al(nrp * 4) + 2 * al(ncp * KP2 * 4) + al(ncp * 4) + 2 * al(ncp * NF2 * 4) + al((size_t)(WIN_TABLE + 1) * 8);
}

int svgd_tc2_supported(int d, int nc) { return d >= 1 && d <= 55 && (nc & 3) == 0 && nc >= 4; }

static int split_for(int blocks_y, int units, int sms) {
  int js = sms / (blocks_y > 0 ? blocks_y : 1);
  if (js < 1) js = 1;
  if (js > 16) js = 16;
  if (js > units) js = units;
  return js;
}

int svgd_tc2_gram(const float* Xr, long long ldr, int nr, int row_offset, const float* Xc, long long ldc, int nc, int d, const float* mu,
                  void* ops_base, float* D2, SelState* st, int sms, cudaStream_t stream) {
  const Tc2Ops o = svgd_tc2_carve(ops_base, nr, nc);
  const int nrp = (nr + BLK - 1) / BLK * BLK, ncp = (nc + BLK - 1) / BLK * BLK;
  prep_x_kernel<<<(int)(((long long)nrp * KCH2 + 255) / 256), 256, 0, stream>>>(Xr, ldr, nr, d, mu, nrp, o.XrH, o.XrL, o.nrm_r);
  prep_x_kernel<<<(int)(((long long)ncp * KCH2 + 255) / 256), 256, 0, stream>>>(Xc, ldc, nc, d, mu, ncp, o.XcH, o.XcL, o.nrm_c);
  BODE_CUDA(cudaGetLastError());
  BODE_CUDA(cudaFuncSetAttribute(gram2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gram2Smem::TOTAL));
  const int nrb = nrp / BLK, nct = ncp / BLK;
  const int js = split_for(nrb, nct, sms);
  const int tiles_per = (nct + js - 1) / js;
  dim3 grid((nct + tiles_per - 1) / tiles_per, nrb);
  gram2_kernel<<<grid, NTHR, Gram2Smem::TOTAL, stream>>>(o.XrH, o.XrL, o.nrm_r, nr, row_offset, o.XcH, o.XcL, o.nrm_c, nc, tiles_per, D2, st,
                                                         o.table);
  return check_cuda(cudaGetLastError(), "gram2 launch");
}

int svgd_tc2_window_select(SelState* st, void* ops_base, int nr, int nc, cudaStream_t stream) {
  const Tc2Ops o = svgd_tc2_carve(ops_base, nr, nc);
  window_select_kernel<<<1, 1024, 0, stream>>>(st, o.table);
  return check_cuda(cudaGetLastError(), "window select launch");
}

unsigned long long* svgd_tc2_table(void* ops_base, int nr, int nc) { return svgd_tc2_carve(ops_base, nr, nc).table; }

int svgd_tc2_phi(const float* D2, int nr, int nc, const float* Xc, long long ldx, const float* Gc, long long ldg, int d, const float* mu,
                 const float* gam, float gsign, void* ops_base, int* jsplit_out, float* part, int sms, cudaStream_t stream) {
  const Tc2Ops o = svgd_tc2_carve(ops_base, nr, nc);
  const int ncp = (nc + BLK - 1) / BLK * BLK;
  prep_v_kernel<<<(int)(((long long)(ncp / 4) * NF2 + 255) / 256), 256, 0, stream>>>(Xc, ldx, Gc, ldg, nc, d, mu, gsign, ncp, o.VH, o.VL);
  BODE_CUDA(cudaGetLastError());
  BODE_CUDA(cudaFuncSetAttribute(phi2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Phi2Smem::TOTAL));
  const int nrb = (nr + BLK - 1) / BLK, nst = (nc + PK2 - 1) / PK2;
  const int js = split_for(nrb, nst, sms);
  *jsplit_out = js;
  dim3 grid(nrb, js);
  phi2_kernel<<<grid, NTHR, Phi2Smem::TOTAL, stream>>>(D2, nr, nc, o.VH, o.VL, d, gam, js, part);
  return check_cuda(cudaGetLastError(), "phi2 launch");
}

}  // namespace bode
