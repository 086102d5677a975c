// Instantiates the solver kernels for the MLP field with hidden width BODE_H (compiled once per width).
// BODE_MF / BODE_SFX select another field type and symbol suffix (mlp_h64tc.cu: the tensor-core field MlpTcField, suffix 64tc).
#include "npde_solve.cuh"
#include "mlp_field.cuh"
#include "mlp_tc.cuh"
#include "dopri5.cuh"

namespace bode {

#define BODE_CAT_(a, b) a##b
#define BODE_CAT(a, b) BODE_CAT_(a, b)
#ifdef BODE_MF
using MF = BODE_MF;
#else
using MF = MlpField<BODE_H>;
#define BODE_SFX BODE_H
#endif

template <int METHOD>
static int mlp_launch_fwd(const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    int e = check_cuda(cudaFuncSetAttribute(npde_fwd_kernel<MF, METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
    if (e != BODE_OK) return e;
  }
  npde_fwd_kernel<MF, METHOD><<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), "mlp fwd launch");
}

template <int METHOD, int INJ, int ADJ>
static int mlp_launch_grad(const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    int e = check_cuda(cudaFuncSetAttribute(npde_grad_kernel<MF, METHOD, INJ, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
    if (e != BODE_OK) return e;
  }
  npde_grad_kernel<MF, METHOD, INJ, ADJ><<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), "mlp grad launch");
}

int BODE_CAT(launch_mlp_dopri5_, BODE_SFX)(const NpdeKParams& prm, const Dopri5Params& dp, dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st) {
  if (smem > 48 * 1024) {
    int e = check_cuda(cudaFuncSetAttribute(dopri5_fwd_kernel<MF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
    if (e != BODE_OK) return e;
  }
  dopri5_fwd_kernel<MF><<<grid, block, smem, st>>>(prm, dp);
  return check_cuda(cudaGetLastError(), "mlp dopri5 launch");
}

int BODE_CAT(launch_mlp_dopri5_grad_, BODE_SFX)(const NpdeKParams& prm, const Dopri5Params& dp, const Dopri5Rec& rec, int inj, dim3 grid,
                                              dim3 block, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    int e = check_cuda(cudaFuncSetAttribute(dopri5_grad_kernel<MF, INJ_LIK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
    if (e != BODE_OK) return e;
    e = check_cuda(cudaFuncSetAttribute(dopri5_grad_kernel<MF, INJ_GOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
    if (e != BODE_OK) return e;
  }
  if (inj == INJ_LIK) dopri5_grad_kernel<MF, INJ_LIK><<<grid, block, smem, st>>>(prm, dp, rec);
  else dopri5_grad_kernel<MF, INJ_GOUT><<<grid, block, smem, st>>>(prm, dp, rec);
  return check_cuda(cudaGetLastError(), "mlp dopri5 grad launch");
}

size_t BODE_CAT(mlp_smem_bytes_, BODE_SFX)(int N) { return sizeof(float) * (size_t)MF::smem_floats(N); }

int BODE_CAT(launch_mlp_fwd_, BODE_SFX)(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return mlp_launch_fwd<BODE_EULER>(prm, grid, block, smem, st);
    case BODE_MIDPOINT: return mlp_launch_fwd<BODE_MIDPOINT>(prm, grid, block, smem, st);
    default: return mlp_launch_fwd<BODE_RK4>(prm, grid, block, smem, st);
  }
}

template <int METHOD>
static int mlp_launch_grad_m(const NpdeKParams& prm, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (inj == INJ_LIK) {
    if (adj == BODE_GRAD_DISCRETE) return mlp_launch_grad<METHOD, INJ_LIK, BODE_GRAD_DISCRETE>(prm, grid, block, smem, st);
    return mlp_launch_grad<METHOD, INJ_LIK, BODE_GRAD_ADJOINT>(prm, grid, block, smem, st);
  }
  if (adj == BODE_GRAD_DISCRETE) return mlp_launch_grad<METHOD, INJ_GOUT, BODE_GRAD_DISCRETE>(prm, grid, block, smem, st);
  return mlp_launch_grad<METHOD, INJ_GOUT, BODE_GRAD_ADJOINT>(prm, grid, block, smem, st);
}

int BODE_CAT(launch_mlp_grad_, BODE_SFX)(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return mlp_launch_grad_m<BODE_EULER>(prm, inj, adj, grid, block, smem, st);
    case BODE_MIDPOINT: return mlp_launch_grad_m<BODE_MIDPOINT>(prm, inj, adj, grid, block, smem, st);
    default: return mlp_launch_grad_m<BODE_RK4>(prm, inj, adj, grid, block, smem, st);
  }
}

}  // namespace bode
