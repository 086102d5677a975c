// Register-resident FMA / MUFU.EX2 chains: the measured non-tensor fp32 and SFU peaks that the integrator kernel's
// roofline is quoted against (MEASURED_PEAKS.json carries only HBM and bf16 tensor figures).
#include "common.cuh"

namespace bode {

__global__ void __launch_bounds__(256) peak_fma_kernel(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;      // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256) peak_ex2_kernel(float* out, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -(float)(threadIdx.x + i) * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -ex2(x[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;
}

}  // namespace bode

using namespace bode;

/* kind 0: dependent-FMA chains (16 per thread) -> flops = 2*16*iters per thread; kind 1: MUFU.EX2 chains (8 per thread).
 * Launches `ctas` CTAs of 256 threads; the caller times it with events and divides. */
extern "C" int bode_peak_kernel(int32_t kind, int32_t ctas, int32_t iters, float* scratch, bode_stream_t stream) {
  BODE_REQUIRE(scratch && ctas > 0 && iters > 0, "bad args");
  if (kind == 0) peak_fma_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(scratch, iters, 0.999f, 1e-3f);
  else peak_ex2_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(scratch, iters);
  return check_cuda(cudaGetLastError(), "peak kernel launch");
}
