// Instantiates the separable npde kernels for one grid size BODE_M (compiled once per size).
#include "npde_solve.cuh"
#include "npde_pair.cuh"
#include "dopri5.cuh"
#include <string.h>

namespace bode {

#define BODE_CAT_(a, b) a##b
#define BODE_CAT(a, b) BODE_CAT_(a, b)

template <int METHOD>
static int launch_fwd_m(const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  using F = SepField<BODE_M, BODE_M>;
  npde_fwd_kernel<F, METHOD><<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), "npde_fwd_kernel launch");
}

template <int METHOD, int INJ, int ADJ>
static int launch_grad_m(const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  using F = SepField<BODE_M, BODE_M>;
  npde_grad_kernel<F, METHOD, INJ, ADJ><<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), "npde_grad_kernel launch");
}

int BODE_CAT(launch_sep_dopri5_, BODE_M)(const NpdeKParams& prm, const Dopri5Params& dp, dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st) {
  dopri5_fwd_kernel<SepField<BODE_M, BODE_M>><<<grid, block, smem, st>>>(prm, dp);
  return check_cuda(cudaGetLastError(), "dopri5 launch");
}

int BODE_CAT(launch_sep_dopri5_grad_, BODE_M)(const NpdeKParams& prm, const Dopri5Params& dp, const Dopri5Rec& rec, int inj, dim3 grid,
                                              dim3 block, size_t smem, cudaStream_t st) {
  if (inj == INJ_LIK) dopri5_grad_kernel<SepField<BODE_M, BODE_M>, INJ_LIK><<<grid, block, smem, st>>>(prm, dp, rec);
  else dopri5_grad_kernel<SepField<BODE_M, BODE_M>, INJ_GOUT><<<grid, block, smem, st>>>(prm, dp, rec);
  return check_cuda(cudaGetLastError(), "dopri5 grad launch");
}

int BODE_CAT(launch_sep_fwd_, BODE_M)(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return launch_fwd_m<BODE_EULER>(prm, grid, block, smem, st);
    case BODE_MIDPOINT: return launch_fwd_m<BODE_MIDPOINT>(prm, grid, block, smem, st);
    default: return launch_fwd_m<BODE_RK4>(prm, grid, block, smem, st);
  }
}

template <int METHOD>
static int launch_grad_ma(const NpdeKParams& prm, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (inj == INJ_LIK) {
    if (adj == BODE_GRAD_DISCRETE) return launch_grad_m<METHOD, INJ_LIK, BODE_GRAD_DISCRETE>(prm, grid, block, smem, st);
    return launch_grad_m<METHOD, INJ_LIK, BODE_GRAD_ADJOINT>(prm, grid, block, smem, st);
  }
  if (adj == BODE_GRAD_DISCRETE) return launch_grad_m<METHOD, INJ_GOUT, BODE_GRAD_DISCRETE>(prm, grid, block, smem, st);
  return launch_grad_m<METHOD, INJ_GOUT, BODE_GRAD_ADJOINT>(prm, grid, block, smem, st);
}

int BODE_CAT(launch_sep_grad_, BODE_M)(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return launch_grad_ma<BODE_EULER>(prm, inj, adj, grid, block, smem, st);
    case BODE_MIDPOINT: return launch_grad_ma<BODE_MIDPOINT>(prm, inj, adj, grid, block, smem, st);
    default: return launch_grad_ma<BODE_RK4>(prm, inj, adj, grid, block, smem, st);
  }
}

// ---- component-split kernels (two lanes per pair): the fixed-step path for square grids
template <class K>
static int launch_big_smem(K kernel, const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const char* what) {
  if (smem > 48 * 1024) {
    int e = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what);
    if (e != BODE_OK) return e;
  }
  kernel<<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), what);
}

int BODE_CAT(launch_pair_fwd_, BODE_M)(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return launch_big_smem(npde_pair_fwd_kernel<BODE_M, BODE_EULER>, prm, grid, block, smem, st, "npde_pair_fwd_kernel");
    case BODE_MIDPOINT: return launch_big_smem(npde_pair_fwd_kernel<BODE_M, BODE_MIDPOINT>, prm, grid, block, smem, st, "npde_pair_fwd_kernel");
    default: return launch_big_smem(npde_pair_fwd_kernel<BODE_M, BODE_RK4>, prm, grid, block, smem, st, "npde_pair_fwd_kernel");
  }
}

template <int METHOD>
static int launch_pair_grad_ma(const NpdeKParams& prm, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  const char* w = "npde_pair_grad_kernel";
  if (inj == INJ_LIK) {
    if (adj == BODE_GRAD_DISCRETE) return launch_big_smem(npde_pair_grad_kernel<BODE_M, METHOD, INJ_LIK, BODE_GRAD_DISCRETE>, prm, grid, block, smem, st, w);
    return launch_big_smem(npde_pair_grad_kernel<BODE_M, METHOD, INJ_LIK, BODE_GRAD_ADJOINT>, prm, grid, block, smem, st, w);
  }
  if (adj == BODE_GRAD_DISCRETE) return launch_big_smem(npde_pair_grad_kernel<BODE_M, METHOD, INJ_GOUT, BODE_GRAD_DISCRETE>, prm, grid, block, smem, st, w);
  return launch_big_smem(npde_pair_grad_kernel<BODE_M, METHOD, INJ_GOUT, BODE_GRAD_ADJOINT>, prm, grid, block, smem, st, w);
}

int BODE_CAT(launch_pair_grad_, BODE_M)(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return launch_pair_grad_ma<BODE_EULER>(prm, inj, adj, grid, block, smem, st);
    case BODE_MIDPOINT: return launch_pair_grad_ma<BODE_MIDPOINT>(prm, inj, adj, grid, block, smem, st);
    default: return launch_pair_grad_ma<BODE_RK4>(prm, inj, adj, grid, block, smem, st);
  }
}

}  // namespace bode
