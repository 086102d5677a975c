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

__global__ void stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

}  // namespace bode

using namespace bode;

/* Measurement aid: one thread writes the GPU's nanosecond timer to *slot when the stream reaches this point.  Unlike an event
 * record it is a kernel node, so it can sit inside a captured CUDA graph (tools/step_timeline.py). */
extern "C" int bode_stamp(unsigned long long* slot, bode_stream_t stream) {
  BODE_REQUIRE(slot, "null slot");
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slot);
  return check_cuda(cudaGetLastError(), "stamp launch");
}

/* kind 0: dependent-FMA chains (16 per thread) -> flops = 2*16*iters per thread; kind 1: MUFU.EX2 chains (8 per thread).
 * Launches `ctas` CTAs of 256 threads; the caller times it with events and divides. */
extern "C" int bode_peak_kernel(int32_t kind, int32_t ctas, int32_t iters, float* scratch, bode_stream_t stream) {
  BODE_REQUIRE(scratch && ctas > 0 && iters > 0, "bad args");
  if (kind == 0) peak_fma_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(scratch, iters, 0.999f, 1e-3f);
  else peak_ex2_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(scratch, iters);
  return check_cuda(cudaGetLastError(), "peak kernel launch");
}
