// Developer microbenchmark (not product code): issue rate / latency of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a,
// alone and mixed with ALU / MUFU / SHFL work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE, int ILP>
__global__ void k(float* out, int iters, float s) {
  float a[ILP]; u64 A[ILP]; unsigned q[ILP];
  const float m = s, c = 1e-9f;
  const u64 M2 = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s), C2 = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x + i; A[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); q[i] = threadIdx.x * 7 + i; }
  for (int t = 0; t < iters; ++t) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) a[i] = fma1(a[i], m, c);                       // FFMA
      if (MODE == 1) A[i] = fma2(A[i], M2, C2);                     // FFMA2
      if (MODE == 2) { A[i] = fma2(A[i], M2, C2); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(t), "r"(i)); }   // FFMA2 + ALU
      if (MODE == 3) { a[i] = fma1(a[i], m, c); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(t), "r"(i)); }     // FFMA + ALU
      if (MODE == 4) { A[i] = fma2(A[i], M2, C2); a[i] = fma1(a[i], m, c); }                                                         // FFMA2 + FFMA
      if (MODE == 5) a[i] = ex2(a[i]);                                                                                               // MUFU
      if (MODE == 6) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);                                                                   // SHFL
      if (MODE == 7) { a[i] = fma1(a[i], m, c); q[i] = __float_as_uint(ex2(__uint_as_float(q[i]))); }                                // FFMA + MUFU
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += a[i] + __uint_as_float((unsigned)A[i]) + __uint_as_float((unsigned)(A[i] >> 32)) + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int ILP>
void run(const char* name, int warps_per_sm, int ops_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  float* out; cudaMalloc(&out, sizeof(float) * sms * warps_per_sm * 32);
  dim3 grid(sms), block(warps_per_sm * 32);
  k<MODE, ILP><<<grid, block>>>(out, 100, 1.0000001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE, ILP><<<grid, block>>>(out, iters, 1.0000001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // warp-instructions per SMSP per cycle assuming max clock
  const double cyc = ms * 1e-3 * clk * 1e3;
  const double winst = (double)iters * ILP * ops_per_iter * warps_per_sm / 4.0;
  printf("%-14s ILP=%d warps/SM=%2d  %.3f ms  %.3f warp-instr/clk/SMSP (at %d MHz)  cyc/iter/warp=%.2f\n", name, ILP, warps_per_sm, ms, winst / cyc, clk / 1000, cyc / iters);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    if (w == 4) { run<0, 1>("FFMA lat", 4, 1); run<1, 1>("FFMA2 lat", 4, 1); run<5, 1>("MUFU lat", 4, 1); run<6, 1>("SHFL lat", 4, 1); }
    switch (w) {
      case 4: run<0, 8>("FFMA", 4, 1); run<1, 8>("FFMA2", 4, 1); run<2, 8>("FFMA2+LOP3", 4, 2); run<3, 8>("FFMA+LOP3", 4, 2); run<4, 8>("FFMA2+FFMA", 4, 2); run<5, 8>("MUFU", 4, 1); run<6, 8>("SHFL", 4, 1); run<7, 8>("FFMA+MUFU", 4, 2); break;
      case 8: run<0, 8>("FFMA", 8, 1); run<1, 8>("FFMA2", 8, 1); run<2, 8>("FFMA2+LOP3", 8, 2); run<3, 8>("FFMA+LOP3", 8, 2); run<4, 8>("FFMA2+FFMA", 8, 2); run<5, 8>("MUFU", 8, 1); run<6, 8>("SHFL", 8, 1); run<7, 8>("FFMA+MUFU", 8, 2); break;
      case 16: run<0, 8>("FFMA", 16, 1); run<1, 8>("FFMA2", 16, 1); run<2, 8>("FFMA2+LOP3", 16, 2); run<3, 8>("FFMA+LOP3", 16, 2); run<4, 8>("FFMA2+FFMA", 16, 2); run<5, 8>("MUFU", 16, 1); run<6, 8>("SHFL", 16, 1); run<7, 8>("FFMA+MUFU", 16, 2); break;
      case 32: run<0, 8>("FFMA", 32, 1); run<1, 8>("FFMA2", 32, 1); run<2, 8>("FFMA2+LOP3", 32, 2); run<4, 8>("FFMA2+FFMA", 32, 2); run<6, 8>("SHFL", 32, 1); break;
    }
  }
  return 0;
}
