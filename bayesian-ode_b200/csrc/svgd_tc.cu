// SVGD contractions on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
//   gram_d2_tc_kernel   d2_ij = |xc_i|^2 + |xc_j|^2 - 2 xc_i.xc_j      (xc = x - column mean; stein.py:22 cdist^2)
//   phi_tc_kernel       part[i, :] = sum_j 2^(-g d2_ij) [ -grad_j | xc_j | 1 ]   (stein.py:75-86, K @ [S | X | 1])
//
// fp32 parity needs more than TF32's 10-bit mantissa, so every fp32 operand is split x = hi + lo (hi = top 19 bits) and each
// product is issued as three MMAs  hi*hi + hi*lo + lo*hi  (3xTF32, ~2^-21 relative).  Operand tiles are produced by the
// CUDA cores straight into shared memory in the canonical no-swizzle UMMA layout (8-row x 16-byte core matrices), one
// elected thread issues the MMAs, completion comes back through tcgen05.commit on an mbarrier, the epilogue reads the
// accumulators with tcgen05.ld (one TMEM lane = one output row per thread).  No TMA: the A operand of the second
// contraction (exp of a distance tile) does not exist in memory, and the first one needs the centred/split transform.
#include "tc_ptx.cuh"
#include "svgd_state.cuh"

namespace bode {

// ---------------------------------------------------------------- column means (deterministic tree per column)
__global__ void __launch_bounds__(256) colmean_kernel(const float* __restrict__ X, long long ld, int n, int d, float* __restrict__ mu,
                                                      SelState* st, unsigned long long total) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  if (c == 0 && threadIdx.x == 0 && st) {        // start of a selection: same reset as select_init_kernel (svgd.cu)
    st->prefix[0] = st->prefix[1] = 0u;
    st->maxbits = 0u;
    st->hit = 0u;
    st->rank[0] = (total - 1) / 2;
    st->rank[1] = total / 2;
  }
  float acc = 0.f;
  for (int r = threadIdx.x; r < n; r += 256) acc += X[(long long)r * ld + c];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) mu[c] = red[0] / (float)n;
}

// ---------------------------------------------------------------- d2 tile = Gram on tensor cores
constexpr int GT = 128;       // output tile per CTA: GT rows x GN columns (two CTAs per SM by shared memory)
constexpr int GN = 64;
constexpr int GTH = 256;      // threads per CTA: two warpgroups share operand generation and split the epilogue columns

template <int KP>             // K padded to a multiple of 8 (d <= KP)
__global__ void __launch_bounds__(GTH) gram_d2_tc_kernel(const float* __restrict__ Xr, long long ldr, int nr, int row_offset,
                                                         const float* __restrict__ Xc, long long ldc, int nc, int d,
                                                         const float* __restrict__ mu, float* __restrict__ D2,
                                                         unsigned int* __restrict__ maxbits) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr int KCH = KP / 4;                       // 16-byte chunks along K
  constexpr uint32_t TILE_A = GT * KP * 4, TILE_B = GN * KP * 4;
  unsigned char* sAh = smraw;
  unsigned char* sAl = sAh + TILE_A;
  unsigned char* sBh = sAl + TILE_A;
  unsigned char* sBl = sBh + TILE_B;
  float* ni = reinterpret_cast<float*>(sBl + TILE_B);
  float* nj = ni + GT;
  uint64_t* bar = reinterpret_cast<uint64_t*>(nj + GN);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.y * GT, c0 = blockIdx.x * GN;

  if (warp == 0) tmem_alloc(tslot, GN);
  if (tid == 0) mbar_init(bar, 1);
  // operand tiles.  Fast path (d and the row strides multiples of 4): every 16-byte chunk of the canonical layout is one
  // aligned float4 of a row, so the loads are fully coalesced; the row norms are then summed from shared memory.
  const bool vec = ((d & 3) == 0) && ((ldr & 3) == 0) && ((ldc & 3) == 0) && ((((uintptr_t)Xr | (uintptr_t)Xc | (uintptr_t)mu) & 15) == 0);
  if (vec) {
    const int dch = d >> 2;
    // all loads of the CTA's operand rows are issued before the first use (fixed trip counts: KCH and KCH/2 per thread)
    constexpr int NA = (GT * KCH + GTH - 1) / GTH, NB = (GN * KCH + GTH - 1) / GTH;
    float4 xa[NA], xb[NB], ma[NA], mb[NB];
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int idx = min(tid + GTH * q, GT * KCH - 1), row = idx / KCH, kc = idx - row * KCH;
      const bool ok = kc < dch && r0 + row < nr;
      const int rr = min(r0 + row, nr - 1), kk = min(kc, dch - 1);
      xa[q] = __ldg(reinterpret_cast<const float4*>(Xr + (long long)rr * ldr) + kk);
      ma[q] = __ldg(reinterpret_cast<const float4*>(mu) + kk);
      if (!ok) { xa[q] = make_float4(0.f, 0.f, 0.f, 0.f); ma[q] = xa[q]; }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int idx = min(tid + GTH * q, GN * KCH - 1), row = idx / KCH, kc = idx - row * KCH;
      const bool ok = kc < dch && c0 + row < nc;
      const int rr = min(c0 + row, nc - 1), kk = min(kc, dch - 1);
      xb[q] = __ldg(reinterpret_cast<const float4*>(Xc + (long long)rr * ldc) + kk);
      mb[q] = __ldg(reinterpret_cast<const float4*>(mu) + kk);
      if (!ok) { xb[q] = make_float4(0.f, 0.f, 0.f, 0.f); mb[q] = xb[q]; }
    }
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int idx = tid + GTH * q, row = idx / KCH, kc = idx - row * KCH;
      if (idx >= GT * KCH) break;
      float h[4], l[4];
      split_tf32(xa[q].x - ma[q].x, h[0], l[0]); split_tf32(xa[q].y - ma[q].y, h[1], l[1]);
      split_tf32(xa[q].z - ma[q].z, h[2], l[2]); split_tf32(xa[q].w - ma[q].w, h[3], l[3]);
      const uint32_t off = kmajor_off<GT>(row, kc);
      *reinterpret_cast<float4*>(sAh + off) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(sAl + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int idx = tid + GTH * q, row = idx / KCH, kc = idx - row * KCH;
      if (idx >= GN * KCH) break;
      float h[4], l[4];
      split_tf32(xb[q].x - mb[q].x, h[0], l[0]); split_tf32(xb[q].y - mb[q].y, h[1], l[1]);
      split_tf32(xb[q].z - mb[q].z, h[2], l[2]); split_tf32(xb[q].w - mb[q].w, h[3], l[3]);
      const uint32_t off = kmajor_off<GN>(row, kc);
      *reinterpret_cast<float4*>(sBh + off) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(sBl + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
    __syncthreads();
    if (tid < GT + GN) {                  // threads 0..127: |xc|^2 of the A rows, 128..191: of the B rows
      const bool isA = tid < GT;
      const int rr = isA ? tid : tid - GT;
      float sa = 0.f;
      for (int kc = 0; kc < KCH; ++kc) {
        const uint32_t off = isA ? kmajor_off<GT>(rr, kc) : kmajor_off<GN>(rr, kc);
        const float4 h = *reinterpret_cast<const float4*>((isA ? sAh : sBh) + off);
        const float4 l = *reinterpret_cast<const float4*>((isA ? sAl : sBl) + off);
        const float a0 = h.x + l.x, a1 = h.y + l.y, a2 = h.z + l.z, a3 = h.w + l.w;
        sa = fmaf(a0, a0, sa); sa = fmaf(a1, a1, sa); sa = fmaf(a2, a2, sa); sa = fmaf(a3, a3, sa);
      }
      if (isA) ni[rr] = sa; else nj[rr] = sa;
    }
  } else if (tid < GT) {
    // general path: thread t owns row t of both tiles
    float sa = 0.f, sb = 0.f;
    const bool va = r0 + tid < nr, vb = tid < GN && c0 + tid < nc;
    const float* xa = Xr + (long long)(r0 + tid) * ldr;
    const float* xb = Xc + (long long)(c0 + tid) * ldc;
#pragma unroll 2
    for (int kc = 0; kc < KCH; ++kc) {
      float a[4], b[4], ah[4], al[4], bh[4], bl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = 4 * kc + q;
        const float m = k < d ? __ldg(mu + k) : 0.f;
        a[q] = (va && k < d) ? __ldg(xa + k) - m : 0.f;
        b[q] = (vb && k < d) ? __ldg(xb + k) - m : 0.f;
        sa = fmaf(a[q], a[q], sa);
        sb = fmaf(b[q], b[q], sb);
        split_tf32(a[q], ah[q], al[q]);
        split_tf32(b[q], bh[q], bl[q]);
      }
      const uint32_t off = kmajor_off<GT>(tid, kc);
      *reinterpret_cast<float4*>(sAh + off) = make_float4(ah[0], ah[1], ah[2], ah[3]);
      *reinterpret_cast<float4*>(sAl + off) = make_float4(al[0], al[1], al[2], al[3]);
      if (tid < GN) {
        const uint32_t offb = kmajor_off<GN>(tid, kc);
        *reinterpret_cast<float4*>(sBh + offb) = make_float4(bh[0], bh[1], bh[2], bh[3]);
        *reinterpret_cast<float4*>(sBl + offb) = make_float4(bl[0], bl[1], bl[2], bl[3]);
      }
    }
    ni[tid] = sa;
    if (tid < GN) nj[tid] = sb;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    constexpr uint32_t idesc = idesc_tf32(GT, GN, 0, 0);
    constexpr uint32_t LBO = GT * 16, LBOB = GN * 16, SBO = 128;
    uint32_t acc = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t a0 = smem_u32(pass == 2 ? sAl : sAh), b0 = smem_u32(pass == 1 ? sBl : sBh);
#pragma unroll 1
      for (int ks = 0; ks < KP / 8; ++ks) {
        umma_tf32(tmem, smem_desc(a0 + ks * 2 * LBO, LBO, SBO), smem_desc(b0 + ks * 2 * LBOB, LBOB, SBO), idesc, acc);
        acc = 1;
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  // epilogue: TMEM lane (row) = 32*warp + lane -> d2 values -> shared staging (operand tiles are dead) -> coalesced rows
  float* stage = reinterpret_cast<float*>(smraw);                  // [GT][GN+4]
  constexpr int SP = GN + 4;
  const int rl = 32 * (warp & 3) + lane;                           // TMEM lane = tile row; warps 4..7 take the upper column half
  const int row = r0 + rl;
  const float nrow = ni[rl];
  float mx = 0.f;
  const int cbeg = (warp >> 2) * (GN / 2);
#pragma unroll 1
  for (int cb = cbeg; cb < cbeg + GN / 2; cb += 8) {
    float s[8];
    tmem_ld8(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cb, s);
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int col = c0 + cb + q;
      float v = fmaxf(fmaf(-2.f, s[q], nrow + nj[cb + q]), 0.f);
      if (row + row_offset == col) v = 0.f;                     // cdist(x, x) = 0 on the diagonal
      o[q] = v;
      if (row < nr && col < nc) mx = fmaxf(mx, v);
    }
    *reinterpret_cast<float4*>(stage + rl * SP + cb) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(stage + rl * SP + cb + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
  __syncthreads();
  if ((nc & 3) == 0 && c0 + GN <= nc) {
    for (int idx = tid; idx < GT * (GN / 4); idx += GTH) {
      const int rr = idx / (GN / 4), c4 = idx - rr * (GN / 4);
      if (r0 + rr < nr)
        *reinterpret_cast<float4*>(D2 + (long long)(r0 + rr) * nc + c0 + 4 * c4) = *reinterpret_cast<const float4*>(stage + rr * SP + 4 * c4);
    }
  } else {
    for (int idx = tid; idx < GT * GN; idx += GTH) {
      const int rr = idx / GN, cc = idx - rr * GN;
      if (r0 + rr < nr && c0 + cc < nc) D2[(long long)(r0 + rr) * nc + c0 + cc] = stage[rr * SP + cc];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) atomicMax(maxbits, __float_as_uint(mx));
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, GN);
}

// ---------------------------------------------------------------- phi partials on tensor cores
constexpr int PT = 128;       // rows per CTA
constexpr int PK = 32;        // j per stage (4 MMA k-steps)
constexpr int PTH = 256;      // threads per CTA (two warpgroups: each builds half of the K tile, half of the epilogue columns)

template <int NF>             // padded feature count, multiple of 16, >= 2d+1
__global__ void __launch_bounds__(PTH) phi_tc_kernel(const float* __restrict__ D2, int nr, int nc, const float* __restrict__ Xc,
                                                     long long ldx, const float* __restrict__ Gc, long long ldg, int d,
                                                     const float* __restrict__ mu, const float* __restrict__ gam, float gsign,
                                                     int jsplit, float* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr uint32_t A_B = PT * PK * 4, B_B = PK * NF * 4;
  unsigned char* sAh = smraw;
  unsigned char* sAl = sAh + A_B;
  unsigned char* sBh = sAl + A_B;
  unsigned char* sBl = sBh + B_B;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBl + B_B);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  constexpr uint32_t TCOLS = NF <= 32 ? 32 : (NF <= 64 ? 64 : (NF <= 128 ? 128 : 256));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r0 = blockIdx.x * PT;
  const int jper = ((nc + jsplit - 1) / jsplit + PK - 1) / PK * PK;
  const int jbeg = blockIdx.y * jper, jend = min(nc, jbeg + jper);
  const float ngamma = -gam[1] * 1.4426950408889634f;

  if (warp == 0) tmem_alloc(tslot, TCOLS);
  if (tid == 0) mbar_init(bar, 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  constexpr uint32_t idesc = idesc_tf32(PT, NF, 0, 0);              // A and B both K-major (B holds V^T)
  constexpr uint32_t A_LBO = PT * 16, A_SBO = 128;                   // [kchunk][row/8][row%8][16B]
  constexpr uint32_t B_LBO = NF * 16, B_SBO = 128;
  uint32_t phase = 0, acc = 0;
  const int rl = tid & (PT - 1), half = tid >> 7;                    // tile row and which half of the stage's columns
  const int row = r0 + rl;
  constexpr int NV = NF * (PK / 4);                                  // 16-byte chunks of V^T per stage
  constexpr int NQ = (NV + PTH - 1) / PTH;
  const float* vsrc[NQ];
  long long vld[NQ];
  float vscale[NQ], vadd[NQ];
  int vjc[NQ], voff[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int slot = tid + PTH * q;
    const int jc = slot / NF, f = slot - jc * NF;
    vjc[q] = jc;
    voff[q] = slot < NV ? (int)kmajor_off<NF>(f, jc) : -1;
    vsrc[q] = Gc; vld[q] = 0; vscale[q] = 0.f; vadd[q] = 0.f;            // padding / ones columns read a dummy word, scale 0
    if (slot < NV) {
      if (f < d) { vsrc[q] = Gc + f; vld[q] = ldg; vscale[q] = gsign; }
      else if (f < 2 * d) { vsrc[q] = Xc + (f - d); vld[q] = ldx; vscale[q] = 1.f; vadd[q] = -__ldg(mu + f - d); }
      else if (f == 2 * d) { vadd[q] = 1.f; }
    }
  }
  const bool avec = (nc & 3) == 0;
  for (int j0 = jbeg; j0 < jend; j0 += PK) {
    // ---- A: K[row][j0..j0+31] = 2^(-g d2), split hi/lo
    constexpr int KCH2 = PK / 8;                                      // chunks of 4 columns built by each half
    float4 dv[KCH2];
#pragma unroll
    for (int kq = 0; kq < KCH2; ++kq) {
      const int kc = half * KCH2 + kq;
      const int j = j0 + 4 * kc;
      if (avec && j + 4 <= jend) {
        dv[kq] = __ldg(reinterpret_cast<const float4*>(D2 + (long long)min(row, nr - 1) * nc + j));
        if (row >= nr) dv[kq] = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
      } else {
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = (row < nr && j + q < jend) ? __ldg(D2 + (long long)row * nc + j + q) : INFINITY;
        dv[kq] = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    // ---- B: V^T[f][j] = [gsign*G | X - mu | 1 | 0]^T, K-major like A (thread <-> (feature f, 4 consecutive j)); the
    //      per-slot source pointer / scale / offset are loop invariant, so the loads are branch free and all in flight
    float4 vv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float t[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + 4 * vjc[q] + e;
        const bool ok = j < jend;
        const float raw = __ldg(vsrc[q] + (long long)min(j, jend - 1) * vld[q]);     // always a valid address: no branch
        t[e] = ok ? fmaf(vscale[q], raw, vadd[q]) : 0.f;
      }
      vv[q] = make_float4(t[0], t[1], t[2], t[3]);
    }
    if (acc) {                                                       // previous stage's MMAs must have drained smem
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
    }
#pragma unroll
    for (int kq = 0; kq < KCH2; ++kq) {
      const int kc = half * KCH2 + kq;
      const float k4[4] = {ex2(ngamma * dv[kq].x), ex2(ngamma * dv[kq].y), ex2(ngamma * dv[kq].z), ex2(ngamma * dv[kq].w)};
      float h[4], l[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) split_tf32(k4[q], h[q], l[q]);
      const uint32_t off = kmajor_off<PT>(rl, kc);
      *reinterpret_cast<float4*>(sAh + off) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(sAl + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (voff[q] >= 0) {
        float h[4], l[4];
        split_tf32(vv[q].x, h[0], l[0]); split_tf32(vv[q].y, h[1], l[1]); split_tf32(vv[q].z, h[2], l[2]); split_tf32(vv[q].w, h[3], l[3]);
        *reinterpret_cast<float4*>(sBh + voff[q]) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(sBl + voff[q]) = make_float4(l[0], l[1], l[2], l[3]);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = smem_u32(pass == 2 ? sAl : sAh), b0 = smem_u32(pass == 1 ? sBl : sBh);
#pragma unroll
        for (int ks = 0; ks < PK / 8; ++ks) {
          umma_tf32(tmem, smem_desc(a0 + ks * 2 * A_LBO, A_LBO, A_SBO), smem_desc(b0 + ks * 2 * B_LBO, B_LBO, B_SBO), idesc, acc);
          acc = 1;
        }
      }
      umma_commit(bar);
    }
    acc = 1;
  }
  if (acc) {
    mbar_wait(bar, phase);
    tc_fence_after();
  }
  // ---- epilogue: one TMEM lane per thread = one output row
  if (jbeg < jend) {
    const int cbeg = half * (NF / 2);                              // warps 4..7 read the upper half of the feature columns
#pragma unroll 1
    for (int cb = cbeg; cb < cbeg + NF / 2; cb += 8) {
      float s[8];
      tmem_ld8(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cb, s);
      if (row < nr) {
        float* dst = part + ((long long)blockIdx.y * nr + row) * (2 * d + 1);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (cb + q <= 2 * d) dst[cb + q] = s[q];
      }
    }
  } else if (row < nr && half == 0) {
    float* dst = part + ((long long)blockIdx.y * nr + row) * (2 * d + 1);
    for (int f = 0; f <= 2 * d; ++f) dst[f] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// combine with centred positions: phi = (KS + 2 g (rowsum (x_i - mu) - K(X - mu))) / n
__global__ void phi_combine_tc_kernel(const float* __restrict__ part, int jsplit, int nr, int d, const float* __restrict__ Xr,
                                      long long ldr, const float* __restrict__ mu, const float* __restrict__ gam, float inv_n,
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
  const float ph = (ks + 2.f * gam[1] * (rs * (x - mu[c]) - kx)) * inv_n;
  if (phi) phi[(long long)r * ldp + c] = ph;
  if (theta) theta[(long long)r * ldt + c] = fmaf(step, ph, theta[(long long)r * ldt + c]);
}

// ---------------------------------------------------------------- host launchers (called from svgd.cu)
int svgd_tc_supported(int d) { return d >= 1 && d <= 56; }

int svgd_tc_colmean(const float* X, long long ld, int n, int d, float* mu, SelState* sel, unsigned long long total, cudaStream_t st) {
  colmean_kernel<<<d, 256, 0, st>>>(X, ld, n, d, mu, sel, total);
  return check_cuda(cudaGetLastError(), "colmean launch");
}

int svgd_tc_gram(const float* Xr, long long ldr, int nr, int row_offset, const float* Xc, long long ldc, int nc, int d, const float* mu,
                 float* D2, unsigned int* maxbits, cudaStream_t st) {
  constexpr int KP = 56;
  const size_t smem = 2 * (size_t)(GT + GN) * KP * 4 + (GT + GN) * 4 + 64;
  BODE_CUDA(cudaFuncSetAttribute(gram_d2_tc_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((nc + GN - 1) / GN, (nr + GT - 1) / GT);
  gram_d2_tc_kernel<KP><<<grid, GTH, smem, st>>>(Xr, ldr, nr, row_offset, Xc, ldc, nc, d, mu, D2, maxbits);
  return check_cuda(cudaGetLastError(), "gram tc launch");
}

int svgd_tc_phi(const float* D2, int nr, int nc, const float* Xc, long long ldx, const float* Gc, long long ldg, int d, const float* mu,
                const float* gam, float gsign, int jsplit, float* part, cudaStream_t st) {
  constexpr int NF = 112;
  const size_t smem = 2 * (size_t)PT * PK * 4 + 2 * (size_t)PK * NF * 4 + 64;
  BODE_CUDA(cudaFuncSetAttribute(phi_tc_kernel<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((nr + PT - 1) / PT, jsplit);
  phi_tc_kernel<NF><<<grid, PTH, smem, st>>>(D2, nr, nc, Xc, ldx, Gc, ldg, d, mu, gam, gsign, jsplit, part);
  return check_cuda(cudaGetLastError(), "phi tc launch");
}

int svgd_tc_combine(const float* part, int jsplit, int nr, int d, const float* Xr, long long ldr, const float* mu, const float* gam,
                    float inv_n, float* phi, long long ldp, float* theta, long long ldt, float step, cudaStream_t st) {
  const long long tot = (long long)nr * d;
  phi_combine_tc_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(part, jsplit, nr, d, Xr, ldr, mu, gam, inv_n, phi, ldp, theta, ldt, step);
  return check_cuda(cudaGetLastError(), "phi combine tc launch");
}

}  // namespace bode
