// Instantiates the fixed-step solver kernels for the row-sliced separable field with BODE_ROW_MY columns per lane.
#include "npde_solve.cuh"
#include "npde_row.cuh"

namespace bode {

template <int MY>
template <int INJ>
__device__ __forceinline__ void RowField<MY>::epilogue(const NpdeKParams& prm, float* smem, const RowField& fld, bool active, int pl,
                                                       int n, int pairl, int lane_, float r2x, float r2y) {
  npde_epilogue<INJ>(prm, smem, fld, active, pl, n, pairl, lane_, r2x, r2y);
}

#define BODE_CAT_(a, b) a##b
#define BODE_CAT(a, b) BODE_CAT_(a, b)
using RF = RowField<BODE_ROW_MY>;

template <int METHOD>
static int row_fwd(const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  npde_fwd_kernel<RF, METHOD><<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), "row fwd launch");
}
template <int METHOD, int INJ, int ADJ>
static int row_grad(const NpdeKParams& prm, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  npde_grad_kernel<RF, METHOD, INJ, ADJ><<<grid, block, smem, st>>>(prm);
  return check_cuda(cudaGetLastError(), "row grad launch");
}

int BODE_CAT(launch_row_fwd_, BODE_ROW_MY)(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return row_fwd<BODE_EULER>(prm, grid, block, smem, st);
    case BODE_MIDPOINT: return row_fwd<BODE_MIDPOINT>(prm, grid, block, smem, st);
    default: return row_fwd<BODE_RK4>(prm, grid, block, smem, st);
  }
}

template <int METHOD>
static int row_grad_m(const NpdeKParams& prm, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  if (inj == INJ_LIK) {
    if (adj == BODE_GRAD_DISCRETE) return row_grad<METHOD, INJ_LIK, BODE_GRAD_DISCRETE>(prm, grid, block, smem, st);
    return row_grad<METHOD, INJ_LIK, BODE_GRAD_ADJOINT>(prm, grid, block, smem, st);
  }
  if (adj == BODE_GRAD_DISCRETE) return row_grad<METHOD, INJ_GOUT, BODE_GRAD_DISCRETE>(prm, grid, block, smem, st);
  return row_grad<METHOD, INJ_GOUT, BODE_GRAD_ADJOINT>(prm, grid, block, smem, st);
}

int BODE_CAT(launch_row_grad_, BODE_ROW_MY)(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem,
                                            cudaStream_t st) {
  switch (method) {
    case BODE_EULER: return row_grad_m<BODE_EULER>(prm, inj, adj, grid, block, smem, st);
    case BODE_MIDPOINT: return row_grad_m<BODE_MIDPOINT>(prm, inj, adj, grid, block, smem, st);
    default: return row_grad_m<BODE_RK4>(prm, inj, adj, grid, block, smem, st);
  }
}

}  // namespace bode
