// Instantiates the solver kernels for the general-Z npde field with BODE_JPL inducing points per lane.
#include "npde_solve.cuh"
#include "npde_gen.cuh"
#include "dopri5.cuh"

namespace bode {

template <int JPL>
template <int INJ>
__device__ __forceinline__ void GenField<JPL>::epilogue(const NpdeKParams& prm, float* smem, const GenField& fld, bool active, int pl,
                                                        int n, int pairl, int lane_, float r2x, float r2y) {
  npde_epilogue<INJ>(prm, smem, fld, active, pl, n, pairl, lane_, r2x, r2y);
}

#define BODE_CAT_(a, b) a##b
#define BODE_CAT(a, b) BODE_CAT_(a, b)
using GF = GenField<BODE_JPL>;

int BODE_CAT(launch_gen_dopri5_, BODE_JPL)(const NpdeKParams& prm, const Dopri5Params& dp, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  dopri5_fwd_kernel<GF><<<grid, block, smem, st>>>(prm, dp);
  return check_cuda(cudaGetLastError(), "gen dopri5 launch");
}

int BODE_CAT(launch_gen_dopri5_grad_, BODE_JPL)(const NpdeKParams& prm, const Dopri5Params& dp, const Dopri5Rec& rec, int inj, dim3 grid,
                                                dim3 block, size_t smem, cudaStream_t st) {
  if (inj == INJ_LIK) dopri5_grad_kernel<GF, INJ_LIK><<<grid, block, smem, st>>>(prm, dp, rec);
  else dopri5_grad_kernel<GF, INJ_GOUT><<<grid, block, smem, st>>>(prm, dp, rec);
  return check_cuda(cudaGetLastError(), "gen dopri5 grad launch");
}

template <int METHOD>
static int gen_fwd(const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  npde_fwd_kernel<GF, METHOD><<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), "gen fwd launch");
}
template <int METHOD, int INJ, int ADJ>
static int gen_grad(const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  npde_grad_kernel<GF, METHOD, INJ, ADJ><<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), "gen grad launch");
}

int BODE_CAT(launch_gen_fwd_, BODE_JPL)(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return gen_fwd<BODE_EULER>(prm, grid, block, smem, st);
    case BODE_MIDPOINT: return gen_fwd<BODE_MIDPOINT>(prm, grid, block, smem, st);
    default: return gen_fwd<BODE_RK4>(prm, grid, block, smem, st);
  }
}

template <int METHOD>
static int gen_grad_m(const NpdeKParams& prm, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (inj == INJ_LIK) {
    if (adj == BODE_GRAD_DISCRETE) return gen_grad<METHOD, INJ_LIK, BODE_GRAD_DISCRETE>(prm, grid, block, smem, st);
    return gen_grad<METHOD, INJ_LIK, BODE_GRAD_ADJOINT>(prm, grid, block, smem, st);
  }
  if (adj == BODE_GRAD_DISCRETE) return gen_grad<METHOD, INJ_GOUT, BODE_GRAD_DISCRETE>(prm, grid, block, smem, st);
  return gen_grad<METHOD, INJ_GOUT, BODE_GRAD_ADJOINT>(prm, grid, block, smem, st);
}

int BODE_CAT(launch_gen_grad_, BODE_JPL)(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return gen_grad_m<BODE_EULER>(prm, inj, adj, grid, block, smem, st);
    case BODE_MIDPOINT: return gen_grad_m<BODE_MIDPOINT>(prm, inj, adj, grid, block, smem, st);
    default: return gen_grad_m<BODE_RK4>(prm, inj, adj, grid, block, smem, st);
  }
}

}  // namespace bode
