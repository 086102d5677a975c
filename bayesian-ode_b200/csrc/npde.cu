// Host side of the npde entry points: argument checks, parameter packing, kernel dispatch.
#include "npde_solve.cuh"
#include "dopri5.cuh"
#include <math.h>

namespace bode {
#define BODE_DECL_ROW(MY)                                                                                                   \
  int launch_row_fwd_##MY(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st);           \
  int launch_row_grad_##MY(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st);
BODE_DECL_ROW(8)
BODE_DECL_ROW(12)
BODE_DECL_ROW(16)
int launch_proj_W(const float* AT, const float* U, long long U_stride, int P, int m, float* W, cudaStream_t st);
int launch_proj_back(const float* A, const float* Ksym, const float* U, long long U_stride, float* gU, long long gU_stride, float* loss,
                     float scale, int add_prior, int P, int m, cudaStream_t st);

// one translation unit per grid size keeps the build parallel; see npde_inst.cuh
#define BODE_DECL_SEP(M)                                                                               \
  int launch_sep_fwd_##M(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_sep_grad_##M(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_pair_fwd_##M(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_pair_grad_##M(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_sep_dopri5_##M(const NpdeKParams& prm, const Dopri5Params& dp, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_sep_dopri5_grad_##M(const NpdeKParams& prm, const Dopri5Params& dp, const Dopri5Rec& rec, int inj, dim3 grid, dim3 block, \
                                 size_t smem, cudaStream_t st);
BODE_DECL_SEP(3)
BODE_DECL_SEP(4)
BODE_DECL_SEP(5)
BODE_DECL_SEP(6)

#define BODE_DECL_GEN(J)                                                                                              \
  int launch_gen_fwd_##J(const NpdeKParams& prm, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st);         \
  int launch_gen_grad_##J(const NpdeKParams& prm, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_gen_dopri5_##J(const NpdeKParams& prm, const Dopri5Params& dp, dim3 grid, dim3 block, size_t smem, cudaStream_t st); \
  int launch_gen_dopri5_grad_##J(const NpdeKParams& prm, const Dopri5Params& dp, const Dopri5Rec& rec, int inj, dim3 grid, dim3 block, \
                                 size_t smem, cudaStream_t st);
BODE_DECL_GEN(1)
BODE_DECL_GEN(2)
BODE_DECL_GEN(4)
BODE_DECL_GEN(8)

static bool use_sep(const bode_npde_field* f) {
  return f->grid_mx == f->grid_my && f->grid_mx >= 3 && f->grid_mx <= 6 && f->grid_mx * f->grid_my == f->m;
}
// SVGD score-tile fusion (svgd_tiles.cuh): while armed, the fused likelihood closure of the component-split kernels also writes
// vsign * (gU | glogsn) of every particle into the interaction's operand tiles; `written` counts such launches.
static struct { float* VH; float* VC; float sign; int P; int d; int written; } g_score = {nullptr, nullptr, 0.f, 0, 0, 0};
void npde_arm_score_tiles(float* VH, float* VC, float sign, int P, int d) {
  g_score.VH = VH; g_score.VC = VC; g_score.sign = sign; g_score.P = P; g_score.d = d; g_score.written = 0;
}
int npde_disarm_score_tiles() {
  const int w = g_score.written;
  g_score.VH = g_score.VC = nullptr;
  g_score.written = 0;
  return w;
}

static int gen_jpl(int m) { return m <= 32 ? 1 : (m <= 64 ? 2 : (m <= 128 ? 4 : 8)); }
// Tensor grids beyond the one-thread separable kernels (7x7 .. 16x16): row-sliced separable field, 16 lanes per pair (npde_row.cuh)
static int g_row_kernel = 1;
static bool use_row(const bode_npde_field* f, int N) {
  return g_row_kernel && !use_sep(f) && f->grid_mx >= 7 && f->grid_mx <= 16 && f->grid_my >= 7 && f->grid_my <= 16 &&
         f->grid_mx * f->grid_my == f->m && N <= 10;          // RowField::MAX_THREADS = 160
}

static int stages_of(int method) { return method == BODE_RK4 ? 4 : (method == BODE_MIDPOINT ? 2 : 1); }

static int fill_common(NpdeKParams& prm, const bode_npde_field* f, const bode_grid* g, int method, int N,
                       const float* y0, int y0_batched) {
  BODE_REQUIRE(f && g, "null field/grid");
  BODE_REQUIRE(f->P > 0 && f->m > 0 && N > 0, "P, m, N must be positive (P=%d m=%d N=%d)", f->P, f->m, N);
  BODE_REQUIRE(g->S >= 0 && g->T >= 1, "bad grid S=%d T=%d", g->S, g->T);
  BODE_REQUIRE(method >= BODE_EULER && method <= BODE_DOPRI5, "unknown method %d", method);
  BODE_REQUIRE(f->U && f->A && y0, "null U/A/y0");
  BODE_REQUIRE(method == BODE_DOPRI5 || g->S == 0 || (g->dt && g->obs_ptr), "null dt/obs_ptr");
  BODE_REQUIRE(f->ell[0] > 0 && f->ell[1] > 0, "ell must be positive");
  memset(&prm, 0, sizeof(prm));
  prm.P = f->P; prm.N = N; prm.S = g->S; prm.T = g->T; prm.m = f->m;
  prm.y0_stride = y0_batched ? 2 * N : 0;
  prm.sign = g->sign; prm.scale = 1.f;
  const double LOG2E = 1.4426950408889634074, LN2 = 0.69314718055994530942;
  const double c0 = sqrt(0.5 * LOG2E) / f->ell[0], c1 = sqrt(0.5 * LOG2E) / f->ell[1];
  prm.c0 = (float)c0; prm.c1 = (float)c1;
  prm.k0 = (float)(2.0 * LN2 * c0); prm.k1 = (float)(2.0 * LN2 * c1);
  prm.U = f->U; prm.A = f->A; prm.AT = f->AT; prm.Ksym = f->Ksym; prm.y0 = y0; prm.dt = g->dt; prm.obs_ptr = g->obs_ptr;
  prm.adj_dt = g->adj_dt; prm.adj_ptr = g->adj_ptr; prm.Z = f->Z;
  BODE_REQUIRE(f->U_stride >= 2 * f->m && (f->U_stride % 2) == 0, "U_stride=%lld must be even and >= 2m", (long long)f->U_stride);
  prm.U_stride = f->U_stride;
  return BODE_OK;
}

static int fill_sep(NpdeKParams& prm, const bode_npde_field* f) {
  BODE_REQUIRE(f->grid_mx == f->grid_my && f->grid_mx >= 3 && f->grid_mx <= 6 && f->grid_mx * f->grid_my == f->m,
               "separable npde kernel supports square inducing grids 3x3..6x6 (got %dx%d, m=%d)", f->grid_mx, f->grid_my, f->m);
  const double LOG2E = 1.4426950408889634074;
  const double c0 = sqrt(0.5 * LOG2E) / f->ell[0], c1 = sqrt(0.5 * LOG2E) / f->ell[1];
  for (int a = 0; a < f->grid_mx; ++a) prm.gxs[a] = (float)(c0 * f->gx[a]);
  for (int b = 0; b < f->grid_my; ++b) prm.gys[b] = (float)(c1 * f->gy[b]);
  return BODE_OK;
}

// CTA shape: ppc particles x N trajectories x G lanes, one warp unless N*G > 32.
static int plan(NpdeKParams& prm, int G, int max_threads, dim3* grid, dim3* block) {
  const int per_particle = prm.N * G;
  BODE_REQUIRE(per_particle <= max_threads, "N*G=%d exceeds %d threads per CTA", per_particle, max_threads);
  int ppc = 32 / per_particle;
  if (ppc < 1) ppc = 1;
  prm.ppc = ppc;
  const int threads = ((ppc * per_particle + 31) / 32) * 32;
  *block = dim3(threads);
  *grid = dim3((prm.P + ppc - 1) / ppc);
  return BODE_OK;
}

// Row-sliced field: 16 lanes per pair; two particles per CTA when one particle would leave half a warp idle.
static int plan_row(NpdeKParams& prm, const bode_npde_field* f, dim3* grid, dim3* block) {
  const double LOG2E = 1.4426950408889634074;
  const double c0 = sqrt(0.5 * LOG2E) / f->ell[0], c1 = sqrt(0.5 * LOG2E) / f->ell[1];
  for (int a = 0; a < f->grid_mx; ++a) prm.gxs[a] = (float)(c0 * f->gx[a]);
  for (int b = 0; b < f->grid_my; ++b) prm.gys[b] = (float)(c1 * f->gy[b]);
  prm.gmx = f->grid_mx;
  prm.gmy = f->grid_my;
  int ppc = 1;
  if (prm.N == 1) ppc = 2;
  else if ((prm.N & 1) && 2 * prm.N * 16 <= 160) ppc = 2;
  prm.ppc = ppc;
  *block = dim3(((ppc * prm.N * 16 + 31) / 32) * 32);
  *grid = dim3((prm.P + ppc - 1) / ppc);
  return BODE_OK;
}
static int row_my(const bode_npde_field* f) { return f->grid_my <= 8 ? 8 : (f->grid_my <= 12 ? 12 : 16); }
static int dispatch_row_fwd(const NpdeKParams& prm, int my, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (my) {
    case 8: return launch_row_fwd_8(prm, method, grid, block, smem, st);
    case 12: return launch_row_fwd_12(prm, method, grid, block, smem, st);
    default: return launch_row_fwd_16(prm, method, grid, block, smem, st);
  }
}
static int dispatch_row_grad(const NpdeKParams& prm, int my, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (my) {
    case 8: return launch_row_grad_8(prm, method, inj, adj, grid, block, smem, st);
    case 12: return launch_row_grad_12(prm, method, inj, adj, grid, block, smem, st);
    default: return launch_row_grad_16(prm, method, inj, adj, grid, block, smem, st);
  }
}

// Component-split kernels (npde_pair.cuh): ppc particles x N trajectories x 2 lanes per CTA.  ppc is sized so that the
// grid is about one CTA per SM (persistent-style sizing, every SM gets the same number of warps) and capped by the
// 320-thread launch bound and the per-particle shared-memory footprint.
static size_t stage_floats(const NpdeKParams& prm);
static const int PAIR_MAX_THREADS = 320;   // kernel choice threshold; a CTA may hold pair_max_threads(M) threads
static int pair_max_threads(int grid_m) { return grid_m <= 5 ? 384 : 320; }   // == PairField<M>::MAX_THREADS
static int g_cta_limit = 0;                // 0 = every SM; > 0: at most this many CTAs for the fused pair kernels
static int sm_count();
static int g_lanes_per_pair = 0;   // 0 = automatic, 1 = one thread per pair (npde_sep.cuh), 2 = component-split lanes (npde_pair.cuh)
static bool use_pair(const bode_npde_field* f, int N) {
  if (!(use_sep(f) && 2 * N <= PAIR_MAX_THREADS)) return false;
  if (g_lanes_per_pair != 0) return g_lanes_per_pair == 2;
  // automatic: two lanes per pair while the pairs alone cannot fill the machine (measured on B200, 5x5 grid, rk4: 98 vs 108 us
  // at P = 4096, 0.68 vs 0.50 ms at P = 32768); beyond ~2 resident CTAs of pairs per SM the one-thread-per-pair kernel wins
  return (long long)f->P * N <= (long long)sm_count() * PAIR_MAX_THREADS;
}

static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

static int plan_pair(NpdeKParams& prm, int grid_m, bool grad, dim3* grid, dim3* block, size_t* smem) {
  const int per_particle = 2 * prm.N;
  const int sms = (g_cta_limit > 0 && g_cta_limit < sm_count()) ? g_cta_limit : sm_count();
  int ppc = (prm.P + sms - 1) / sms;
  const int cap = pair_max_threads(grid_m) / per_particle;
  if (ppc > cap) ppc = cap;
  if (ppc < 1) ppc = 1;
  size_t floats = 0;
  for (;; --ppc) {
    prm.ppc = ppc;
    if (grad) {
      prm.stage_off = (int)(((size_t)ppc * 2 * prm.m * (2 + prm.N) + (size_t)ppc * prm.N * 2 + 3) & ~(size_t)3);
      prm.a_off = (int)(((size_t)prm.stage_off + stage_floats(prm) + 3) & ~(size_t)3);
    } else {
      prm.a_off = ppc * 2 * prm.m * 2;
    }
    floats = (size_t)prm.a_off + 2 * (size_t)prm.m * prm.m;       // + A | Ksym
    if (floats * sizeof(float) <= 160 * 1024 || ppc == 1) break;
  }
  BODE_REQUIRE(floats * sizeof(float) <= 200 * 1024, "solver grid too long to stage in shared memory (S=%d)", prm.S);
  *smem = floats * sizeof(float);
  *block = dim3(((ppc * per_particle + 31) / 32) * 32);
  *grid = dim3((prm.P + ppc - 1) / ppc);
  return BODE_OK;
}

static int dispatch_pair_fwd(const NpdeKParams& prm, int M, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (M) {
    case 3: return launch_pair_fwd_3(prm, method, grid, block, smem, st);
    case 4: return launch_pair_fwd_4(prm, method, grid, block, smem, st);
    case 5: return launch_pair_fwd_5(prm, method, grid, block, smem, st);
    case 6: return launch_pair_fwd_6(prm, method, grid, block, smem, st);
  }
  set_error("no separable kernel for M=%d", M);
  return BODE_ERR_UNSUPPORTED;
}

static int dispatch_pair_grad(const NpdeKParams& prm, int M, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st) {
  switch (M) {
    case 3: return launch_pair_grad_3(prm, method, inj, adj, grid, block, smem, st);
    case 4: return launch_pair_grad_4(prm, method, inj, adj, grid, block, smem, st);
    case 5: return launch_pair_grad_5(prm, method, inj, adj, grid, block, smem, st);
    case 6: return launch_pair_grad_6(prm, method, inj, adj, grid, block, smem, st);
  }
  set_error("no separable kernel for M=%d", M);
  return BODE_ERR_UNSUPPORTED;
}

static int dispatch_fwd(const NpdeKParams& prm, int M, int method, dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  switch (M) {
    case 3: return launch_sep_fwd_3(prm, method, grid, block, smem, st);
    case 4: return launch_sep_fwd_4(prm, method, grid, block, smem, st);
    case 5: return launch_sep_fwd_5(prm, method, grid, block, smem, st);
    case 6: return launch_sep_fwd_6(prm, method, grid, block, smem, st);
  }
  set_error("no separable kernel for M=%d", M);
  return BODE_ERR_UNSUPPORTED;
}

static int dispatch_grad(const NpdeKParams& prm, int M, int method, int inj, int adj, dim3 grid, dim3 block, size_t smem,
                         cudaStream_t st) {
  switch (M) {
    case 3: return launch_sep_grad_3(prm, method, inj, adj, grid, block, smem, st);
    case 4: return launch_sep_grad_4(prm, method, inj, adj, grid, block, smem, st);
    case 5: return launch_sep_grad_5(prm, method, inj, adj, grid, block, smem, st);
    case 6: return launch_sep_grad_6(prm, method, inj, adj, grid, block, smem, st);
  }
  set_error("no separable kernel for M=%d", M);
  return BODE_ERR_UNSUPPORTED;
}

// staged dt[S] | obs_ptr[S+1] | Y[N,T,2] after the field's own shared-memory region
static size_t stage_floats(const NpdeKParams& prm) { return (size_t)((2 * prm.S + 2) & ~1) + 2 * (size_t)prm.N * prm.T + 2; }

static size_t scratch_floats(long long P, long long N, int S, int T, int method, int grad_mode) {
  const long long npairs = P * N;
  const long long slots = grad_mode == BODE_GRAD_ADJOINT ? (long long)T : (long long)S * stages_of(method);
  return (size_t)(2 * npairs * (slots > 0 ? slots : 1));
}

static int run_grad(const bode_npde_field* f, const bode_grid* g, int method, int grad_mode, int inj, int N,
                    const float* y0, int y0_batched, NpdeKParams& prm, float* scratch, size_t scratch_n, cudaStream_t st) {
  BODE_REQUIRE(grad_mode == BODE_GRAD_DISCRETE || grad_mode == BODE_GRAD_ADJOINT, "unknown grad_mode %d", grad_mode);
  BODE_REQUIRE(grad_mode != BODE_GRAD_ADJOINT || g->T == 1 || (g->adj_dt && g->adj_ptr), "ADJOINT needs adj_dt/adj_ptr");
  const size_t need = scratch_floats(f->P, N, g->S, g->T, method, grad_mode);
  BODE_REQUIRE(scratch && scratch_n >= need, "scratch too small: have %zu floats, need %zu", scratch_n, need);
  BODE_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "scratch must be 8-byte aligned");
  prm.ck = reinterpret_cast<float2*>(scratch);
  prm.npairs = (long long)f->P * N;
  dim3 grid, block;
  if (use_pair(f, N)) {
    int st_ = fill_sep(prm, f);
    if (st_ != BODE_OK) return st_;
    size_t smem = 0;
    st_ = plan_pair(prm, f->grid_mx, true, &grid, &block, &smem);
    if (st_ != BODE_OK) return st_;
    if (g_score.VH && inj == INJ_LIK && f->P == g_score.P && 2 * f->m + 2 == g_score.d && prm.glogsn && prm.gU) {
      prm.VH = g_score.VH; prm.VC = g_score.VC; prm.vsign = g_score.sign;
      ++g_score.written;
    }
    return dispatch_pair_grad(prm, f->grid_mx, method, inj, grad_mode, grid, block, smem, st);
  }
  if (use_sep(f)) {
    int st_ = fill_sep(prm, f);
    if (st_ != BODE_OK) return st_;
    st_ = plan(prm, 1, 256, &grid, &block);
    if (st_ != BODE_OK) return st_;
    prm.stage_off = (int)(((size_t)prm.ppc * 2 * prm.m * (2 + N) + (size_t)prm.ppc * N * 2 + 3) & ~(size_t)3);
    const size_t smem = sizeof(float) * ((size_t)prm.stage_off + stage_floats(prm));
    BODE_REQUIRE(smem <= 48 * 1024, "solver grid too long to stage in shared memory (S=%d)", prm.S);
    return dispatch_grad(prm, f->grid_mx, method, inj, grad_mode, grid, block, smem, st);
  }
  // general inducing locations (or a grid outside 3x3..6x6): lane-sliced kernel, one warp per (particle, trajectory)
  BODE_REQUIRE(f->Z && f->m <= 256, "general-Z npde kernel needs Z and m <= 256 (got m=%d)", f->m);
  bool row = use_row(f, N);
  int st_ = row ? plan_row(prm, f, &grid, &block) : plan(prm, 32, 256, &grid, &block);
  if (st_ != BODE_OK) return st_;
  prm.stage_off = (int)(((size_t)prm.ppc * 2 * prm.m * (2 + N) + (size_t)prm.ppc * N * 2 + 3) & ~(size_t)3);
  size_t smem = sizeof(float) * ((size_t)prm.stage_off + stage_floats(prm));
  if (row && smem > 48 * 1024) {          // two particles per CTA do not fit next to a long solver grid: the general-Z kernel
    row = false;
    st_ = plan(prm, 32, 256, &grid, &block);
    if (st_ != BODE_OK) return st_;
    prm.stage_off = (int)(((size_t)prm.ppc * 2 * prm.m * (2 + N) + (size_t)prm.ppc * N * 2 + 3) & ~(size_t)3);
    smem = sizeof(float) * ((size_t)prm.stage_off + stage_floats(prm));
  }
  BODE_REQUIRE(smem <= 48 * 1024, "solver grid too long to stage in shared memory (S=%d)", prm.S);
  // Large inducing sets: the projections W = A U and gU = A^T gW + Ksym U as panel GEMMs over all particles (npde_proj.cu) around
  // the solve -- three launches instead of one -- when the caller's scratch holds the P x 2m projected values behind the checkpoints
  // (bode_npde_scratch_floats_m); the prior then enters the loss in proj_back_kernel.
  const size_t w_off = (need + 3) & ~(size_t)3;
  const bool split = f->m >= 64 && f->AT && scratch_n >= w_off + (size_t)f->P * 2 * f->m;
  const int add_prior = prm.add_prior;
  if (split) {
    float* Wpre = scratch + w_off;
    st_ = launch_proj_W(f->AT, f->U, f->U_stride, f->P, f->m, Wpre, st);
    if (st_ != BODE_OK) return st_;
    prm.Wpre = Wpre;
    prm.split = 1;
    prm.add_prior = 0;
  }
  if (row) st_ = dispatch_row_grad(prm, row_my(f), method, inj, grad_mode, grid, block, smem, st);
  else switch (gen_jpl(f->m)) {
    case 1: st_ = launch_gen_grad_1(prm, method, inj, grad_mode, grid, block, smem, st); break;
    case 2: st_ = launch_gen_grad_2(prm, method, inj, grad_mode, grid, block, smem, st); break;
    case 4: st_ = launch_gen_grad_4(prm, method, inj, grad_mode, grid, block, smem, st); break;
    default: st_ = launch_gen_grad_8(prm, method, inj, grad_mode, grid, block, smem, st); break;
  }
  if (st_ != BODE_OK || !split) return st_;
  return launch_proj_back(f->A, f->Ksym, f->U, f->U_stride, prm.gU, prm.gU_stride, inj == INJ_LIK ? prm.loss : nullptr, prm.scale, add_prior,
                          f->P, f->m, st);
}

}  // namespace bode

using namespace bode;

/* Kernel choice for square 3x3..6x6 grids: 0 automatic, 1 one thread per (particle, trajectory) pair, 2 two lanes per pair. */
extern "C" int bode_npde_set_lanes_per_pair(int32_t lanes) {
  const int old = g_lanes_per_pair;
  g_lanes_per_pair = (lanes == 1 || lanes == 2) ? lanes : 0;
  return old;
}

/* Leave SMs free for kernels that run beside the fused solve on another stream (the SVGD Gram pass): the pair kernels then use
 * at most max_ctas CTAs (one per SM), packing up to 12 warps into each.  0 restores one CTA per SM.  Returns the old value. */
extern "C" int bode_npde_set_cta_limit(int32_t max_ctas) {
  const int old = g_cta_limit;
  g_cta_limit = max_ctas > 0 ? max_ctas : 0;
  return old;
}

extern "C" int bode_npde_set_row_kernel(int32_t on) {
  const int old = g_row_kernel;
  g_row_kernel = on ? 1 : 0;
  return old;
}

extern "C" size_t bode_npde_scratch_floats(int32_t P, int32_t N, int32_t S, int32_t T, int32_t method, int32_t grad_mode) {
  return scratch_floats(P, N, S, T, method, grad_mode);
}
/* ... plus room for the projected values W = A U of all particles when m >= 64: the gradient entry points then run the
 * projections as panel GEMMs around the solve (npde_proj.cu) instead of inside it. */
extern "C" size_t bode_npde_scratch_floats_m(int32_t P, int32_t N, int32_t S, int32_t T, int32_t method, int32_t grad_mode, int32_t m) {
  const size_t base = (scratch_floats(P, N, S, T, method, grad_mode) + 3) & ~(size_t)3;
  return base + (m >= 64 ? (size_t)P * 2 * (size_t)m : 0);
}

extern "C" int bode_npde_odeint(const bode_npde_field* f, const bode_grid* g, int32_t method, int32_t N, const float* y0,
                                int32_t y0_batched, float* sol, bode_stream_t stream) {
  NpdeKParams prm;
  int st = fill_common(prm, f, g, method, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(sol, "null sol");
  prm.sol = sol;
  dim3 grid, block;
  if (use_pair(f, N)) {
    st = fill_sep(prm, f);
    if (st != BODE_OK) return st;
    size_t smem = 0;
    st = plan_pair(prm, f->grid_mx, false, &grid, &block, &smem);
    if (st != BODE_OK) return st;
    return dispatch_pair_fwd(prm, f->grid_mx, method, grid, block, smem, (cudaStream_t)stream);
  }
  if (use_sep(f)) {
    st = fill_sep(prm, f);
    if (st != BODE_OK) return st;
    st = plan(prm, 1, 256, &grid, &block);
    if (st != BODE_OK) return st;
    const size_t smem = sizeof(float) * (size_t)prm.ppc * 2 * prm.m * 2;
    return dispatch_fwd(prm, f->grid_mx, method, grid, block, smem, (cudaStream_t)stream);
  }
  BODE_REQUIRE(f->Z && f->m <= 256, "general-Z npde kernel needs Z and m <= 256 (got m=%d)", f->m);
  const bool row = use_row(f, N);
  st = row ? plan_row(prm, f, &grid, &block) : plan(prm, 32, 256, &grid, &block);
  if (st != BODE_OK) return st;
  const size_t smem = sizeof(float) * (size_t)prm.ppc * 2 * prm.m * 2;
  if (row) return dispatch_row_fwd(prm, row_my(f), method, grid, block, smem, (cudaStream_t)stream);
  switch (gen_jpl(f->m)) {
    case 1: return launch_gen_fwd_1(prm, method, grid, block, smem, (cudaStream_t)stream);
    case 2: return launch_gen_fwd_2(prm, method, grid, block, smem, (cudaStream_t)stream);
    case 4: return launch_gen_fwd_4(prm, method, grid, block, smem, (cudaStream_t)stream);
    default: return launch_gen_fwd_8(prm, method, grid, block, smem, (cudaStream_t)stream);
  }
}

int fill_dopri5(Dopri5Params& dp, const bode_dopri5_opts* o, int N) {
  BODE_REQUIRE(o && o->t, "null dopri5 options / t");
  BODE_REQUIRE(o->rtol > 0 && o->atol >= 0 && o->safety > 0 && o->ifactor > 0 && o->dfactor > 0, "bad dopri5 tolerances/factors");
  dp.t = o->t; dp.rtol = (float)o->rtol; dp.atol = (float)o->atol; dp.user_first_step = o->user_first_step;
  dp.safety = o->safety; dp.ifactor = o->ifactor; dp.dfactor = o->dfactor;
  dp.max_num_steps = o->max_num_steps > 0 ? o->max_num_steps : 2147483647;
  dp.stats = o->stats;
  BODE_REQUIRE(o->controller == 0 || o->controller == 1, "dopri5 controller must be 0 (per pair) or 1 (per particle, pooled)");
  dp.pool = o->controller;
  BODE_REQUIRE(o->n_groups >= 0 && o->n_groups <= 4, "dopri5: at most 4 state tensors in a tuple (got %d)", o->n_groups);
  BODE_REQUIRE(o->n_groups <= 1 || o->controller == 1, "dopri5: tuple states need the pooled controller");
  dp.ngroups = o->n_groups;
  for (int g = 0; g < 4; ++g) dp.gend[g] = o->group_end[g];
  if (o->n_groups > 1) {
    for (int g = 0; g < o->n_groups; ++g)
      BODE_REQUIRE(o->group_end[g] > (g ? o->group_end[g - 1] : 0), "dopri5: group_end must be strictly increasing");
    BODE_REQUIRE(o->group_end[o->n_groups - 1] == N, "dopri5: the state tensors of the tuple must cover the N=%d trajectories", N);
  }
  return BODE_OK;
}

extern "C" int bode_npde_dopri5(const bode_npde_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                                const float* y0, int32_t y0_batched, float* sol, bode_stream_t stream) {
  bode_grid g = {};
  g.S = 0; g.T = T; g.sign = sign;
  NpdeKParams prm;
  int st = fill_common(prm, f, &g, BODE_DOPRI5, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(sol, "null sol");
  prm.sol = sol;
  Dopri5Params dp;
  st = fill_dopri5(dp, o, N);
  if (st != BODE_OK) return st;
  dim3 grid, block;
  if (!use_sep(f)) {
    BODE_REQUIRE(f->Z && f->m <= 256, "general-Z npde kernel needs Z and m <= 256 (got m=%d)", f->m);
    st = plan(prm, 32, 256, &grid, &block);
    if (st != BODE_OK) return st;
    const size_t smemg = sizeof(float) * (size_t)prm.ppc * 2 * prm.m * 2;
    switch (gen_jpl(f->m)) {
      case 1: return launch_gen_dopri5_1(prm, dp, grid, block, smemg, (cudaStream_t)stream);
      case 2: return launch_gen_dopri5_2(prm, dp, grid, block, smemg, (cudaStream_t)stream);
      case 4: return launch_gen_dopri5_4(prm, dp, grid, block, smemg, (cudaStream_t)stream);
      default: return launch_gen_dopri5_8(prm, dp, grid, block, smemg, (cudaStream_t)stream);
    }
  }
  st = fill_sep(prm, f);
  if (st != BODE_OK) return st;
  st = plan(prm, 1, 256, &grid, &block);
  if (st != BODE_OK) return st;
  const size_t smem = sizeof(float) * (size_t)prm.ppc * 2 * prm.m * 2;
  switch (f->grid_mx) {
    case 3: return launch_sep_dopri5_3(prm, dp, grid, block, smem, (cudaStream_t)stream);
    case 4: return launch_sep_dopri5_4(prm, dp, grid, block, smem, (cudaStream_t)stream);
    case 5: return launch_sep_dopri5_5(prm, dp, grid, block, smem, (cudaStream_t)stream);
    default: return launch_sep_dopri5_6(prm, dp, grid, block, smem, (cudaStream_t)stream);
  }
}

int carve_dopri5_rec(bode::Dopri5Rec& rec, float* scratch, size_t scratch_n, long long npairs, int T, int max_rec) {
  BODE_REQUIRE(max_rec >= 1, "max_rec_steps must be >= 1");
  const size_t need = bode_dopri5_scratch_floats((int32_t)npairs, 1, T, max_rec);
  BODE_REQUIRE(scratch && scratch_n >= need, "dopri5 scratch too small: have %zu floats, need %zu", scratch_n, need);
  BODE_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "dopri5 scratch must be 16-byte aligned");
  rec.outs = reinterpret_cast<float4*>(scratch);
  rec.steps = reinterpret_cast<float2*>(scratch + 4 * (size_t)T * npairs);
  rec.max_rec = max_rec;
  return BODE_OK;
}

extern "C" size_t bode_dopri5_scratch_floats(int32_t P, int32_t N, int32_t T, int32_t max_rec_steps) {
  const size_t npairs = (size_t)P * N;
  return 4 * (size_t)T * npairs + 14 * (size_t)max_rec_steps * npairs;
}

static int npde_dopri5_grad(const bode_npde_field* f, const bode_dopri5_opts* o, int T, float sign, int N, const float* y0, int y0_batched,
                            NpdeKParams& prm, int inj, float* scratch, size_t scratch_n, int max_rec, cudaStream_t st) {
  Dopri5Params dp;
  int e = fill_dopri5(dp, o, N);
  if (e != BODE_OK) return e;
  Dopri5Rec rec;
  e = carve_dopri5_rec(rec, scratch, scratch_n, (long long)f->P * N, T, max_rec);
  if (e != BODE_OK) return e;
  prm.npairs = (long long)f->P * N;
  dim3 grid, block;
  if (use_sep(f)) {
    e = fill_sep(prm, f);
    if (e != BODE_OK) return e;
    e = plan(prm, 1, 256, &grid, &block);
    if (e != BODE_OK) return e;
    const size_t smem = sizeof(float) * ((size_t)prm.ppc * 2 * prm.m * (2 + N) + (size_t)prm.ppc * N * 2);
    switch (f->grid_mx) {
      case 3: return launch_sep_dopri5_grad_3(prm, dp, rec, inj, grid, block, smem, st);
      case 4: return launch_sep_dopri5_grad_4(prm, dp, rec, inj, grid, block, smem, st);
      case 5: return launch_sep_dopri5_grad_5(prm, dp, rec, inj, grid, block, smem, st);
      default: return launch_sep_dopri5_grad_6(prm, dp, rec, inj, grid, block, smem, st);
    }
  }
  BODE_REQUIRE(f->Z && f->m <= 256, "general-Z npde kernel needs Z and m <= 256 (got m=%d)", f->m);
  e = plan(prm, 32, 256, &grid, &block);
  if (e != BODE_OK) return e;
  const size_t smem = sizeof(float) * ((size_t)prm.ppc * 2 * prm.m * (2 + N) + (size_t)prm.ppc * N * 2);
  switch (gen_jpl(f->m)) {
    case 1: return launch_gen_dopri5_grad_1(prm, dp, rec, inj, grid, block, smem, st);
    case 2: return launch_gen_dopri5_grad_2(prm, dp, rec, inj, grid, block, smem, st);
    case 4: return launch_gen_dopri5_grad_4(prm, dp, rec, inj, grid, block, smem, st);
    default: return launch_gen_dopri5_grad_8(prm, dp, rec, inj, grid, block, smem, st);
  }
}

extern "C" int bode_npde_dopri5_backward(const bode_npde_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                                         const float* y0, int32_t y0_batched, const float* gout, float* gU, int64_t gU_stride,
                                         float* gy0, float* scratch, size_t scratch_n, int32_t max_rec_steps, bode_stream_t stream) {
  bode_grid g = {};
  g.S = 0; g.T = T; g.sign = sign;
  NpdeKParams prm;
  int st = fill_common(prm, f, &g, BODE_DOPRI5, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(gout && gU && gU_stride >= 2 * f->m, "null gout/gU or bad stride");
  prm.gout = gout; prm.gU = gU; prm.gU_stride = gU_stride; prm.gy0 = gy0; prm.add_prior = 0;
  return npde_dopri5_grad(f, o, T, sign, N, y0, y0_batched, prm, INJ_GOUT, scratch, scratch_n, max_rec_steps, (cudaStream_t)stream);
}

extern "C" int bode_npde_dopri5_nlp_grad(const bode_npde_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                                         const float* y0, int32_t y0_batched, const float* Y, const float* logsn, int64_t logsn_stride,
                                         float scale, int32_t add_prior, float* loss, float* sqerr, float* gU, int64_t gU_stride,
                                         float* glogsn, int64_t glogsn_stride, float* scratch, size_t scratch_n, int32_t max_rec_steps,
                                         bode_stream_t stream) {
  bode_grid g = {};
  g.S = 0; g.T = T; g.sign = sign;
  NpdeKParams prm;
  int st = fill_common(prm, f, &g, BODE_DOPRI5, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(Y && logsn && loss && sqerr && gU && glogsn, "null Y/logsn/outputs");
  BODE_REQUIRE(!add_prior || f->Ksym, "add_prior needs Ksym");
  BODE_REQUIRE(gU_stride >= 2 * f->m && logsn_stride >= 2 && glogsn_stride >= 2 && (logsn_stride % 2) == 0, "bad strides");
  prm.Y = Y; prm.logsn = logsn; prm.scale = scale; prm.add_prior = add_prior ? 1 : 0;
  prm.loss = loss; prm.sqerr = sqerr; prm.gU = gU; prm.glogsn = glogsn;
  prm.logsn_stride = logsn_stride; prm.gU_stride = gU_stride; prm.glogsn_stride = glogsn_stride;
  return npde_dopri5_grad(f, o, T, sign, N, y0, y0_batched, prm, INJ_LIK, scratch, scratch_n, max_rec_steps, (cudaStream_t)stream);
}

extern "C" int bode_npde_odeint_backward(const bode_npde_field* f, const bode_grid* g, int32_t method, int32_t grad_mode,
                                         int32_t N, const float* y0, int32_t y0_batched, const float* gout, float* gU,
                                         int64_t gU_stride, float* gy0, float* scratch, size_t scratch_n, bode_stream_t stream) {
  NpdeKParams prm;
  int st = fill_common(prm, f, g, method, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(gout && gU, "null gout/gU");
  BODE_REQUIRE(gU_stride >= 2 * f->m, "gU_stride too small");
  prm.gout = gout; prm.gU = gU; prm.gU_stride = gU_stride; prm.gy0 = gy0; prm.add_prior = 0;
  return run_grad(f, g, method, grad_mode, INJ_GOUT, N, y0, y0_batched, prm, scratch, scratch_n, (cudaStream_t)stream);
}

extern "C" int bode_npde_nlp_grad(const bode_npde_field* f, const bode_grid* g, int32_t method, int32_t grad_mode, int32_t N,
                                  const float* y0, int32_t y0_batched, const float* Y, const float* logsn, int64_t logsn_stride,
                                  float scale, int32_t add_prior, float* loss, float* sqerr, float* gU, int64_t gU_stride,
                                  float* glogsn, int64_t glogsn_stride, float* scratch,
                                  size_t scratch_n, bode_stream_t stream) {
  NpdeKParams prm;
  int st = fill_common(prm, f, g, method, N, y0, y0_batched);
  if (st != BODE_OK) return st;
  BODE_REQUIRE(Y && logsn && loss && sqerr && gU && glogsn, "null Y/logsn/outputs");
  BODE_REQUIRE(!add_prior || f->Ksym, "add_prior needs Ksym");
  prm.Y = Y; prm.logsn = logsn; prm.scale = scale; prm.add_prior = add_prior ? 1 : 0;
  BODE_REQUIRE(gU_stride >= 2 * f->m && logsn_stride >= 2 && glogsn_stride >= 2 && (logsn_stride % 2) == 0, "bad strides");
  BODE_REQUIRE((reinterpret_cast<uintptr_t>(logsn) & 7) == 0, "logsn must be 8-byte aligned");
  prm.loss = loss; prm.sqerr = sqerr; prm.gU = gU; prm.glogsn = glogsn;
  prm.logsn_stride = logsn_stride; prm.gU_stride = gU_stride; prm.glogsn_stride = glogsn_stride;
  return run_grad(f, g, method, grad_mode, INJ_LIK, N, y0, y0_batched, prm, scratch, scratch_n, (cudaStream_t)stream);
}
