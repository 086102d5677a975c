// Layout of the SVGD phi operand V^T = [-G | X - mu | 1] (svgd_tc2.cu, prep_v_kernel), shared with the npde closure kernel that
// writes the SCORE columns itself when the sampler asks for it (bode_svgd_arm_score_tiles): the scores are the last thing the fused
// solve produces and the first thing the interaction needs, and a separate operand-preparation launch between the two sat on the
// critical path of every SVGD step (7 us of 137 on c3).
//   VH[j / 4][f][j % 4]                       tf32 hi part of V[j][f]  (K-major core matrices, K = particle index j)
//   VC: per 32-particle stage 8 K-chunks x SV_NF features x 16 bytes; chunk c holds bf16 V_lo of particles 8c .. 8c+7 of the stage,
//       chunk 4 + c bf16 V itself (the one-product correction operand, see prep_v_kernel)
#pragma once
#include <cuda_bf16.h>

namespace bode {

constexpr int SV_NF = 112;                             // padded feature count (>= 2 d + 1, multiple of 16)
constexpr int SV_PK = 32;                              // particles per phi stage
constexpr unsigned int SV_VST_BYTES = SV_PK * SV_NF * 4;

__device__ __forceinline__ void score_tile_store(float* VH, float* VC, long long j, int f, float v) {
  const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);      // split_tf32 (tc_ptx.cuh)
  const float lo = v - hi;
  const long long jq = j >> 2;
  const int e = (int)(j & 3);
  VH[(jq * SV_NF + f) * 4 + e] = hi;
  const long long s = jq >> 3;
  const int c = (int)(jq & 7) >> 1, half = (int)(jq & 1);
  unsigned char* base = reinterpret_cast<unsigned char*>(VC) + s * SV_VST_BYTES + (long long)f * 16 + half * 8 + 2 * e;
  *reinterpret_cast<__nv_bfloat16*>(base + (long long)c * (SV_NF * 16)) = __float2bfloat16_rn(lo);
  *reinterpret_cast<__nv_bfloat16*>(base + (long long)(c + 4) * (SV_NF * 16)) = __float2bfloat16_rn(v);
}

// The same for the four particles 4 jq .. 4 jq + 3 at once: one 16-byte and two 8-byte stores, consecutive f -> consecutive addresses
// (the write pattern of prep_v_kernel).  Element stores from every CTA at the end of the solve were 21 MB of partial-sector L2
// writes in one burst: +4 us on the c3 step instead of -7.
__device__ __forceinline__ void score_tile_store4(float* VH, float* VC, long long jq, int f, const float (&v)[4]) {
  float hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hi[e] = __uint_as_float(__float_as_uint(v[e]) & 0xFFFFE000u);
    lo[e] = v[e] - hi[e];
  }
  *reinterpret_cast<float4*>(VH + (jq * SV_NF + f) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  const long long s = jq >> 3;
  const int c = (int)(jq & 7) >> 1, half = (int)(jq & 1);
  unsigned char* base = reinterpret_cast<unsigned char*>(VC) + s * SV_VST_BYTES + (long long)f * 16 + half * 8;
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(lo[0], lo[1]), l1 = __floats2bfloat162_rn(lo[2], lo[3]);
  const __nv_bfloat162 f0 = __floats2bfloat162_rn(v[0], v[1]), f1 = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(base + (long long)c * (SV_NF * 16)) =
      make_uint2(*reinterpret_cast<const unsigned int*>(&l0), *reinterpret_cast<const unsigned int*>(&l1));
  *reinterpret_cast<uint2*>(base + (long long)(c + 4) * (SV_NF * 16)) =
      make_uint2(*reinterpret_cast<const unsigned int*>(&f0), *reinterpret_cast<const unsigned int*>(&f1));
}

}  // namespace bode
