// Host side of the MLP neural-ODE entry points (argument checks, parameter packing, dispatch on the hidden width).
#include "npde_solve.cuh"
#include "dopri5.cuh"
#include <string.h>

int fill_dopri5(bode::Dopri5Params& dp, const bode_dopri5_opts* o, int N);
int carve_dopri5_rec(bode::Dopri5Rec& rec, float* scratch, size_t scratch_n, long long npairs, int T, int max_rec);

namespace bode {
#define BODE_DECL_MLP(H)                                                                                                  \
  size_t mlp_smem_bytes_##H(int N);                                                                                       \
  int launch_mlp_fwd_##H(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st);         \
  int launch_mlp_grad_##H(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_mlp_dopri5_##H(const NpdeKParams& prm, const Dopri5Params& dp, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_mlp_dopri5_grad_##H(const NpdeKParams& prm, const Dopri5Params& dp, const Dopri5Rec& rec, int inj, dim3 grid, dim3 block, \
                                 size_t smem, cudaStream_t st);
BODE_DECL_MLP(20)
BODE_DECL_MLP(64)
BODE_DECL_MLP(64tc)

// H = 64: the 64 x 64 layer runs on the tensor cores (mlp_tc.cuh) whenever the trajectories of a particle advance in lock-step --
// fixed-grid solvers, or dopri5 with the pooled controller -- and 4 <= N <= 8.  0 forces the FP32-pipe kernels (MlpField).
static int g_mlp_tc = 1;
static bool mlp_use_tc(const bode_mlp_field* f, int N, bool lockstep) { return g_mlp_tc && f->H == 64 && N >= 4 && N <= 8 && lockstep; }

static int mlp_fill(NpdeKParams& prm, const bode_mlp_field* f, const bode_grid* g, int method, int N, const float* y0, int y0_batched) {
  BODE_REQUIRE(f && g, "null field/grid");
  BODE_REQUIRE(f->H == 20 || f->H == 64, "MLP field is built for hidden widths 20 and 64 (got %d)", f->H);
  BODE_REQUIRE(f->P > 0 && N > 0 && N <= 8, "need P > 0 and 1 <= N <= 8 trajectories (got P=%d N=%d)", f->P, N);
  BODE_REQUIRE(g->S >= 0 && g->T >= 1, "bad grid S=%d T=%d", g->S, g->T);
  BODE_REQUIRE(method >= BODE_EULER && method <= BODE_DOPRI5, "unknown method %d", method);
  const int d = f->H * f->H + 6 * f->H + 2;
  BODE_REQUIRE(f->theta && y0 && f->theta_stride >= d, "null theta/y0 or theta_stride < d=%d", d);
  BODE_REQUIRE(method == BODE_DOPRI5 || g->S == 0 || (g->dt && g->obs_ptr), "null dt/obs_ptr");
  memset(&prm, 0, sizeof(prm));
  prm.P = f->P; prm.N = N; prm.S = g->S; prm.T = g->T; prm.m = 0; prm.ppc = 1;
  prm.y0_stride = y0_batched ? 2 * N : 0;
  prm.sign = g->sign; prm.scale = 1.f;
  prm.U = f->theta; prm.U_stride = f->theta_stride; prm.y0 = y0; prm.dt = g->dt; prm.obs_ptr = g->obs_ptr;
  prm.adj_dt = g->adj_dt; prm.adj_ptr = g->adj_ptr;
  return BODE_OK;
}

static size_t mlp_smem(int H, int N, bool tc = false) { return tc ? mlp_smem_bytes_64tc(N) : (H == 20 ? mlp_smem_bytes_20(N) : mlp_smem_bytes_64(N)); }

static int mlp_grad(const bode_mlp_field* f, const bode_grid* g, int method, int grad_mode, int inj, int N, NpdeKParams& prm,
                    float* scratch, size_t scratch_n, cudaStream_t st) {
  BODE_REQUIRE(grad_mode == BODE_GRAD_DISCRETE || grad_mode == BODE_GRAD_ADJOINT, "unknown grad_mode %d", grad_mode);
  BODE_REQUIRE(grad_mode != BODE_GRAD_ADJOINT || g->T == 1 || (g->adj_dt && g->adj_ptr), "ADJOINT needs adj_dt/adj_ptr");
  const size_t need = bode_npde_scratch_floats(f->P, N, g->S, g->T, method, grad_mode);
  BODE_REQUIRE(scratch && scratch_n >= need, "scratch too small: have %zu floats, need %zu", scratch_n, need);
  prm.ck = reinterpret_cast<float2*>(scratch);
  prm.npairs = (long long)f->P * N;
  const dim3 grid(f->P), block(32 * N);
  const bool tc = mlp_use_tc(f, N, true);
  prm.stage_off = (int)((mlp_smem(f->H, N, tc) / sizeof(float) + 3) & ~(size_t)3);
  const size_t smem = sizeof(float) * ((size_t)prm.stage_off + ((2 * (size_t)prm.S + 2) & ~(size_t)1) + 2 * (size_t)N * prm.T + 2);
  if (tc) return launch_mlp_grad_64tc(prm, method, inj, grad_mode, grid, block, smem, st);
  if (f->H == 20) return launch_mlp_grad_20(prm, method, inj, grad_mode, grid, block, smem, st);
  return launch_mlp_grad_64(prm, method, inj, grad_mode, grid, block, smem, st);
}
}  // namespace bode

using namespace bode;

/* 1 (default): tensor-core kernels for H = 64 where the trajectories advance in lock-step; 0: FP32-pipe kernels.  Returns the old value. */
extern "C" int bode_mlp_set_tensor_cores(int32_t on) {
  const int old = g_mlp_tc;
  g_mlp_tc = on ? 1 : 0;
  return old;
}

extern "C" int bode_mlp_odeint(const bode_mlp_field* f, const bode_grid* g, int32_t method, int32_t N, const float* y0,
                               int32_t y0_batched, float* sol, bode_stream_t stream) {
  NpdeKParams prm;
  int st = mlp_fill(prm, f, g, method, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(sol, "null sol");
  prm.sol = sol;
  const dim3 grid(f->P), block(32 * N);
  const bool tc = mlp_use_tc(f, N, true);
  const size_t smem = mlp_smem(f->H, N, tc);
  if (tc) return launch_mlp_fwd_64tc(prm, method, grid, block, smem, (cudaStream_t)stream);
  if (f->H == 20) return launch_mlp_fwd_20(prm, method, grid, block, smem, (cudaStream_t)stream);
  return launch_mlp_fwd_64(prm, method, grid, block, smem, (cudaStream_t)stream);
}

extern "C" int bode_mlp_dopri5(const bode_mlp_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                               const float* y0, int32_t y0_batched, float* sol, bode_stream_t stream) {
  bode_grid g = {};
  g.S = 0; g.T = T; g.sign = sign;
  NpdeKParams prm;
  int st = mlp_fill(prm, f, &g, BODE_DOPRI5, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(sol, "null sol");
  prm.sol = sol;
  Dopri5Params dp;
  st = fill_dopri5(dp, o, N);
  if (st != BODE_OK) return st;
  const dim3 grid(f->P), block(32 * N);
  const bool tc = mlp_use_tc(f, N, dp.pool != 0);
  const size_t smem = mlp_smem(f->H, N, tc);
  if (tc) return launch_mlp_dopri5_64tc(prm, dp, grid, block, smem, (cudaStream_t)stream);
  if (f->H == 20) return launch_mlp_dopri5_20(prm, dp, grid, block, smem, (cudaStream_t)stream);
  return launch_mlp_dopri5_64(prm, dp, grid, block, smem, (cudaStream_t)stream);
}

static int mlp_dopri5_grad(const bode_mlp_field* f, const bode_dopri5_opts* o, int T, int N, NpdeKParams& prm, int inj, float* scratch,
                           size_t scratch_n, int max_rec, cudaStream_t st) {
  Dopri5Params dp;
  int e = fill_dopri5(dp, o, N);
  if (e != BODE_OK) return e;
  Dopri5Rec rec;
  e = carve_dopri5_rec(rec, scratch, scratch_n, (long long)f->P * N, T, max_rec);
  if (e != BODE_OK) return e;
  prm.npairs = (long long)f->P * N;
  const dim3 grid(f->P), block(32 * N);
  const bool tc = mlp_use_tc(f, N, dp.pool != 0);
  const size_t smem = mlp_smem(f->H, N, tc);
  if (tc) return launch_mlp_dopri5_grad_64tc(prm, dp, rec, inj, grid, block, smem, st);
  if (f->H == 20) return launch_mlp_dopri5_grad_20(prm, dp, rec, inj, grid, block, smem, st);
  return launch_mlp_dopri5_grad_64(prm, dp, rec, inj, grid, block, smem, st);
}

extern "C" int bode_mlp_dopri5_backward(const bode_mlp_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                                        const float* y0, int32_t y0_batched, const float* gout, float* gtheta, int64_t gtheta_stride,
                                        float* gy0, float* scratch, size_t scratch_n, int32_t max_rec_steps, bode_stream_t stream) {
  bode_grid g = {};
  g.S = 0; g.T = T; g.sign = sign;
  NpdeKParams prm;
  int st = mlp_fill(prm, f, &g, BODE_DOPRI5, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(gout && gtheta, "null gout/gtheta");
  prm.gout = gout; prm.gU = gtheta; prm.gU_stride = gtheta_stride; prm.gy0 = gy0; prm.add_prior = 0;
  return mlp_dopri5_grad(f, o, T, N, prm, INJ_GOUT, scratch, scratch_n, max_rec_steps, (cudaStream_t)stream);
}

extern "C" int bode_mlp_dopri5_sse_grad(const bode_mlp_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                                        const float* y0, int32_t y0_batched, const float* X, float lik_w, float reg, float scale,
                                        int32_t add_prior, float* loss, float* sqerr, float* gtheta, int64_t gtheta_stride,
                                        float* scratch, size_t scratch_n, int32_t max_rec_steps, bode_stream_t stream) {
  bode_grid g = {};
  g.S = 0; g.T = T; g.sign = sign;
  NpdeKParams prm;
  int st = mlp_fill(prm, f, &g, BODE_DOPRI5, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(X && loss && sqerr && gtheta, "null X/outputs");
  prm.Y = X; prm.lik_w = lik_w; prm.reg = reg; prm.scale = scale; prm.add_prior = add_prior ? 1 : 0;
  prm.loss = loss; prm.sqerr = sqerr; prm.gU = gtheta; prm.gU_stride = gtheta_stride;
  return mlp_dopri5_grad(f, o, T, N, prm, INJ_LIK, scratch, scratch_n, max_rec_steps, (cudaStream_t)stream);
}

extern "C" int bode_mlp_odeint_backward(const bode_mlp_field* f, const bode_grid* g, int32_t method, int32_t grad_mode, int32_t N,
                                        const float* y0, int32_t y0_batched, const float* gout, float* gtheta,
                                        int64_t gtheta_stride, float* gy0, float* scratch, size_t scratch_n, bode_stream_t stream) {
  NpdeKParams prm;
  int st = mlp_fill(prm, f, g, method, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(gout && gtheta, "null gout/gtheta");
  prm.gout = gout; prm.gU = gtheta; prm.gU_stride = gtheta_stride; prm.gy0 = gy0; prm.add_prior = 0;
  return mlp_grad(f, g, method, grad_mode, INJ_GOUT, N, prm, scratch, scratch_n, (cudaStream_t)stream);
}

extern "C" int bode_mlp_sse_grad(const bode_mlp_field* f, const bode_grid* g, int32_t method, int32_t grad_mode, int32_t N,
                                 const float* y0, int32_t y0_batched, const float* X, float lik_w, float reg, float scale,
                                 int32_t add_prior, float* loss, float* sqerr, float* gtheta, int64_t gtheta_stride,
                                 float* scratch, size_t scratch_n, bode_stream_t stream) {
  NpdeKParams prm;
  int st = mlp_fill(prm, f, g, method, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(X && loss && sqerr && gtheta, "null X/outputs");
  prm.Y = X; prm.lik_w = lik_w; prm.reg = reg; prm.scale = scale; prm.add_prior = add_prior ? 1 : 0;
  prm.loss = loss; prm.sqerr = sqerr; prm.gU = gtheta; prm.gU_stride = gtheta_stride;
  return mlp_grad(f, g, method, grad_mode, INJ_LIK, N, prm, scratch, scratch_n, (cudaStream_t)stream);
}
