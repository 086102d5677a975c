/*
 * bode_b200.h -- C ABI of the B200-native bayesian-ode hot path (libbode_b200.so).
 *
 * The reference (jaivardhankapoor/bayesian-ode) is pure Python and has no FFI; the
 * entry points below are what a binding for its hot path would call.  Each one names
 * the reference interface it replaces (file:line relative to the reference checkout).
 * All pointers are DEVICE pointers (fp32 unless stated) except where marked HOST; all
 * calls are asynchronous on `stream` (a cudaStream_t) and return 0 on success or a
 * negative bode_status; bode_last_error() gives the message of the last failure on the
 * calling thread.  No CPU fallback exists: without a CUDA device every compute entry
 * point fails with BODE_ERR_CUDA.
 *
 * Layouts (row-major, leading index slowest):
 *   U      [P, m, 2]   whitened inducing values per particle      (gp.py:59  KernelRegression.U)
 *   logsn  [P, 2]      log noise std per particle                 (gp.py:60)
 *          U/logsn and their gradients carry an explicit per-particle stride (in floats) so that they can
 *          be column blocks of one resident theta[P, d] / grad[P, d] buffer (d = 2m+2): the samplers and
 *          the SVGD interaction then work on the flat buffers without any packing copies.
 *   Z      [m, 2]      inducing locations, shared                 (gp.py:62)
 *   A      [m, m]      sf^2 * Kzz^-1 L, shared                    (gp.py:67,70-71; sf^2 folded in)
 *   Ksym   [m, m]      (Kzz^-1 + Kzz^-T)/2, shared                (prior term gp.py:350)
 *   y0     [P, N, 2] or [N, 2] (y0_batched = 0)                   (odeint.py:20 y0)
 *   sol    [T, P, N, 2]  time-major like torchdiffeq              (solvers.py:99)
 *   Y      [N, T, 2]   observations                               (gp.py:320,345)
 */
#ifndef BODE_B200_H
#define BODE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* bode_stream_t; /* cudaStream_t */

enum bode_status {
  BODE_OK = 0,
  BODE_ERR_ARG = -1,         /* bad argument / unsupported shape */
  BODE_ERR_CUDA = -2,        /* CUDA runtime error (incl. no device) */
  BODE_ERR_UNSUPPORTED = -3
};

/* torchdiffeq SOLVERS registry entries on the hot path (odeint.py:8-17) */
enum bode_method { BODE_EULER = 0, BODE_MIDPOINT = 1, BODE_RK4 = 2 /* 3/8 rule, fixed_grid.py:29 */, BODE_DOPRI5 = 3 };

/* which gradient (SURVEY.md hard part 1):
 *   DISCRETE = autograd through odeint (exact reverse of the discretisation)
 *   ADJOINT  = odeint_adjoint, adjoint.py:23-102 (per-interval restart, same method) */
enum bode_grad_mode { BODE_GRAD_DISCRETE = 0, BODE_GRAD_ADJOINT = 1 };

int bode_version(void);
const char* bode_last_error(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int bode_device_sm_count(void);

/* Measurement aid: register-resident chains for the non-tensor fp32 FMA peak (kind 0: 32*iters flop per thread)
 * and the MUFU.EX2 peak (kind 1: 8*iters ex2 per thread); `ctas` CTAs of 256 threads.  Timed by the caller. */
int bode_peak_kernel(int32_t kind, int32_t ctas, int32_t iters, float* scratch, bode_stream_t stream);
/* Measurement aid: writes the GPU nanosecond timer (%globaltimer) to *slot (device memory) when `stream` reaches this point; a
 * kernel node, so it can be captured in a CUDA graph where event records cannot be read back (tools/step_timeline.py). */
int bode_stamp(unsigned long long* slot, bode_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * npde field description: KernelRegression (gp.py:56-71) for P particles sharing Z/sf/ell.
 * If the inducing points form a tensor grid (gp.py:315-318 always builds one) set
 * grid_mx, grid_my > 0 and pass the axis coordinates: Z[a*grid_my+b] = (gx[a], gy[b]);
 * the separable kernel is then used.  Otherwise set both to 0 and the general-Z kernel runs.
 * ------------------------------------------------------------------------------------- */
typedef struct bode_npde_field {
  int32_t P;            /* particles / chains */
  int32_t m;            /* inducing points */
  int32_t grid_mx, grid_my;
  double  gx[32];       /* HOST values, axis coordinates (separable case) */
  double  gy[32];
  double  ell[2];       /* length-scales (gp.py:49-54 broadcasts a scalar) */
  const float* Z;       /* [m,2] */
  const float* A;       /* [m,m] */
  const float* Ksym;    /* [m,m]  (may be NULL when no prior term is requested) */
  const float* U;       /* [P,m,2], particle p at U + p*U_stride */
  int64_t U_stride;     /* floats between consecutive particles (>= 2m) */
  const float* AT;      /* optional [m,m]: A transposed.  The projection W = A U_p (gp.py:70, hoisted to once per solve) reads A
                           with the output index fastest; with AT those loads are coalesced (m = 256: 16 cache lines -> 1 per warp
                           load).  NULL: A itself is read. */
} bode_npde_field;

/* Solver grid, precomputed on the host exactly as FixedGridODESolver.integrate does
 * (solvers.py:79-97) in the state dtype:
 *   dt[s]      = grid[s+1]-grid[s], s < S
 *   obs_ptr[s] .. obs_ptr[s+1]  = output indices emitted after step s (output 0 is y0);
 *                 every emitted value is the END-of-step state (solvers.py:93-96 quirk)
 *   sign       = -1 when t was decreasing (misc.py:184-187), else +1
 * For BODE_GRAD_ADJOINT the inner reverse solves (adjoint.py:81-84) use
 *   adj_dt[adj_ptr[i-1] .. adj_ptr[i])  = step sizes of the solve from t[i] to t[i-1], i=1..T-1 */
typedef struct bode_grid {
  int32_t S, T;
  float   sign;
  const float*   dt;       /* [S] device */
  const int32_t* obs_ptr;  /* [S+1] device */
  const float*   adj_dt;   /* device, may be NULL unless ADJOINT */
  const int32_t* adj_ptr;  /* [T] device, may be NULL unless ADJOINT */
} bode_grid;

/* Adaptive Dormand-Prince 5(4), Dopri5Solver (dopri5.py:58-122), ONE step-size controller per (particle, trajectory)
 * pair -- each pair reproduces the control flow of its own reference odeint call (the notebook integrates one row per
 * call).  t is float64 like the reference's controller (solvers.py:28) and already increasing (pass sign = -1 and the
 * negated times for decreasing t, misc.py:184-187).  Defaults of odeint / Dopri5Solver: rtol 1e-7, atol 1e-9, safety .9,
 * ifactor 10, dfactor .2, max_num_steps 2^31-1 per output time; user_first_step != 0 reproduces dopri5.py:81-82 (0.01).
 * stats (optional, device int32 [P*N][3]) receives accepted steps, rejected steps and status bits
 * (1 = max_num_steps exceeded, 2 = dt underflow, 4 = non-finite state; dopri5.py:89,100-102 assert on these). */
typedef struct bode_dopri5_opts {
  const double* t;      /* [T] device */
  double rtol, atol, safety, ifactor, dfactor;
  int32_t max_num_steps, user_first_step;
  int32_t* stats;
  /* Controller granularity.  0: one step-size controller per (particle, trajectory) pair == one reference odeint call per
   * trajectory row (nn.ipynb cell 10).  1: one controller per particle with the error ratio (misc.py:146-157) and the
   * initial-step norms (misc.py:116-143) pooled over its N trajectories x 2 components == one reference odeint call with
   * y0 [N, 2] (gp.py:346, 452).  stats then repeats the particle's counts for each of its pairs. */
  int32_t controller;
  /* Tuple states (misc.py:175-182) with controller = 1: the N trajectories of a particle are the concatenation of n_groups state
   * tensors, group g = trajectories [group_end[g-1], group_end[g]).  The error is pooled per tensor and combined by max, as
   * torchdiffeq does for tuples (dopri5.py:108-109, misc.py:125-141, 161).  n_groups <= 1: a single tensor.  At most 4. */
  int32_t n_groups;
  int32_t group_end[4];
} bode_dopri5_opts;

int bode_npde_dopri5(const bode_npde_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                     const float* y0, int32_t y0_batched, float* sol, bode_stream_t stream);

/* dopri5 gradients: discrete adjoint of the ACCEPTED steps and of the dense-output evaluation with the accepted step sizes
 * frozen (equals odeint_adjoint(dopri5) to ~1e-5 at tight tolerance; autograd through the reference's controller is not a
 * usable gradient).  The solve is redone inside the call (one launch = adaptive forward + record + reverse sweep).
 * scratch: bode_dopri5_scratch_floats(P, N, T, max_rec_steps) floats, 16-byte aligned; a pair that accepts more than
 * max_rec_steps steps sets status bit 8 in stats and contributes no gradient. */
size_t bode_dopri5_scratch_floats(int32_t P, int32_t N, int32_t T, int32_t max_rec_steps);
int bode_npde_dopri5_backward(const bode_npde_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                              const float* y0, int32_t y0_batched, const float* gout, float* gU, int64_t gU_stride,
                              float* gy0, float* scratch, size_t scratch_floats, int32_t max_rec_steps, bode_stream_t stream);
int bode_npde_dopri5_nlp_grad(const bode_npde_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                              const float* y0, int32_t y0_batched, const float* Y, const float* logsn, int64_t logsn_stride,
                              float scale, int32_t add_prior, float* loss, float* sqerr, float* gU, int64_t gU_stride,
                              float* glogsn, int64_t glogsn_stride, float* scratch, size_t scratch_floats, int32_t max_rec_steps,
                              bode_stream_t stream);

/* scratch floats needed by the gradient entry points for (P particles, N trajectories) */
/* Kernel choice for square 3x3..6x6 inducing grids (fixed-step solvers): 0 = automatic, 1 = one thread per (particle,
 * trajectory) pair, 2 = two component-split lanes per pair.  Returns the previous setting. */
int bode_npde_set_lanes_per_pair(int32_t lanes);
/* The component-split kernels normally spread the particles over every SM (one CTA each).  max_ctas > 0 packs them into at most
 * that many CTAs (up to 12 warps each), leaving the other SMs to kernels that run beside the solve on a second stream -- the
 * position-only half of the SVGD interaction (bode_svgd_sqdist_staged + median), whose CTAs cannot share an SM with the solver's.
 * The solve is latency-bound at three warps per scheduler either way.  0 = every SM.  Returns the previous setting. */
int bode_npde_set_cta_limit(int32_t max_ctas);
/* Tensor grids of 7x7 .. 16x16 inducing points run the row-sliced separable kernels (csrc/npde_row.cuh: Mx + My exponentials per
 * evaluation, 16 lanes per pair) in the fixed-step entry points; 0 sends them through the general-Z kernels instead (what
 * non-grid Z always uses) -- the parity tests compare the two.  Returns the previous setting. */
int bode_npde_set_row_kernel(int32_t on);
size_t bode_npde_scratch_floats(int32_t P, int32_t N, int32_t S, int32_t T, int32_t method, int32_t grad_mode);
/* The same plus P x 2m floats for m >= 64 inducing points: with that much scratch (and bode_npde_field.AT set) the gradient entry
 * points run the projections W = A U and gU = A^T gW + Ksym U (gp.py:70, 350) as panel GEMMs over all particles around the solve
 * instead of per-particle loops inside it (three launches; BASELINE config 5, m = 256). */
size_t bode_npde_scratch_floats_m(int32_t P, int32_t N, int32_t S, int32_t T, int32_t method, int32_t grad_mode, int32_t m);

/* odeint(func=KernelRegression, y0, t, method in {euler,midpoint,rk4}) forward
 * replaces torchdiffeq/_impl/odeint.py:20-76 + solvers.py:79-99 + fixed_grid.py for the npde field.
 * sol [T,P,N,2]. */
int bode_npde_odeint(const bode_npde_field* f, const bode_grid* g, int32_t method,
                     int32_t N, const float* y0, int32_t y0_batched,
                     float* sol, bode_stream_t stream);

/* backward of the above for an arbitrary downstream loss: given gout = dL/dsol [T,P,N,2]
 * returns gU [P,m,2] (dL/dU) and optionally gy0 [P,N,2] (may be NULL).
 * DISCRETE replaces autograd-through-odeint (gradient_tests.py:19-37);
 * ADJOINT replaces OdeintAdjointMethod.backward (adjoint.py:23-102). */
int bode_npde_odeint_backward(const bode_npde_field* f, const bode_grid* g, int32_t method, int32_t grad_mode,
                              int32_t N, const float* y0, int32_t y0_batched,
                              const float* gout, float* gU, int64_t gU_stride, float* gy0,
                              float* scratch, size_t scratch_floats, bode_stream_t stream);

/* fused posterior closure + gradient: loss_closure (gp.py:342-353) followed by loss.backward()
 * for every particle in ONE launch:
 *   loss[p]  = sum (Y-x)^2/(2 e^{2 logsn}) + N*T*sum_d logsn_d + tr(U^T Kzz^-1 U)/2   (x scale)
 *   sqerr[p] = sum (Y-x)^2                     (closure(add_prior=False); not scaled)
 *   gU, glogsn = gradients of loss (x scale).  scale = 1/N for pSGLD (langevin.py:528).
 * add_prior=0 drops the prior term from loss and gU (likelihood terms stay). */
int bode_npde_nlp_grad(const bode_npde_field* f, const bode_grid* g, int32_t method, int32_t grad_mode,
                       int32_t N, const float* y0, int32_t y0_batched,
                       const float* Y, const float* logsn, int64_t logsn_stride, float scale, int32_t add_prior,
                       float* loss, float* sqerr, float* gU, int64_t gU_stride, float* glogsn, int64_t glogsn_stride,
                       float* scratch, size_t scratch_floats, bode_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused SG-MCMC parameter updates on flat buffers of n floats (theta[P*d] etc., 16-byte aligned).
 * xi (and xi_resample) are optional INJECTED standard-normal draws for bit-parity runs; when NULL the
 * kernel draws from counter-based Philox4x32-10 keyed on (seed, step).  *status (device int, may be
 * NULL) gets bit 0 set when a parameter is non-finite on entry (langevin.py:184-185 -> ValueError).
 * ------------------------------------------------------------------------------------- */

/* ---------------------------------------------------------------------------------------
 * MLP neural-ODE field (notebooks/jai/nn.ipynb cell 4: Linear(2,H)-ELU-Linear(H,H)-ELU-Linear(H,2)), one weight set
 * per particle.  theta[p] = [W1 (H x 2) | b1 (H) | W2 (H x H) | b2 (H) | W3 (2 x H) | b3 (2)]  (parameters() order),
 * d = H^2 + 6H + 2; built for H = 20 (notebook) and H = 64 (BASELINE config 4); N <= 8 trajectories.
 * Scratch size: bode_npde_scratch_floats(P, N, S, T, method, grad_mode).
 * ------------------------------------------------------------------------------------- */
typedef struct bode_mlp_field {
  int32_t P, H;
  const float* theta;     /* [P, d], particle p at theta + p*theta_stride */
  int64_t theta_stride;
} bode_mlp_field;

/* odeint(net, x0, t, method) forward for every particle and trajectory row (nn.ipynb cell 10 loops rows) */
int bode_mlp_odeint(const bode_mlp_field* f, const bode_grid* g, int32_t method, int32_t N, const float* y0,
                    int32_t y0_batched, float* sol, bode_stream_t stream);
/* H = 64: the 64 x 64 hidden layer runs on the tensor cores (mma.sync tf32, 3x split: fp32-level accuracy) whenever the
 * trajectories of a particle advance in lock-step -- fixed-grid solvers, dopri5 with controller = 1 -- and 4 <= N <= 8.
 * on = 0 forces the FP32-pipe kernels everywhere.  Returns the previous setting. */
int bode_mlp_set_tensor_cores(int32_t on);
/* odeint(net, x0, t, method='dopri5') forward, per-pair controller (see bode_dopri5_opts) */
int bode_mlp_dopri5(const bode_mlp_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                    const float* y0, int32_t y0_batched, float* sol, bode_stream_t stream);
int bode_mlp_dopri5_backward(const bode_mlp_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                             const float* y0, int32_t y0_batched, const float* gout, float* gtheta, int64_t gtheta_stride,
                             float* gy0, float* scratch, size_t scratch_floats, int32_t max_rec_steps, bode_stream_t stream);
int bode_mlp_dopri5_sse_grad(const bode_mlp_field* f, const bode_dopri5_opts* o, int32_t T, float sign, int32_t N,
                             const float* y0, int32_t y0_batched, const float* X, float lik_w, float reg, float scale,
                             int32_t add_prior, float* loss, float* sqerr, float* gtheta, int64_t gtheta_stride,
                             float* scratch, size_t scratch_floats, int32_t max_rec_steps, bode_stream_t stream);
/* its backward for an arbitrary dL/dsol [T,P,N,2]: gtheta [P,d] (+ optional gy0 [P,N,2]) */
int bode_mlp_odeint_backward(const bode_mlp_field* f, const bode_grid* g, int32_t method, int32_t grad_mode, int32_t N,
                             const float* y0, int32_t y0_batched, const float* gout, float* gtheta,
                             int64_t gtheta_stride, float* gy0, float* scratch, size_t scratch_floats, bode_stream_t stream);
/* fused bayesian_closure (nn.ipynb cell 10) + backward:  loss = lik_w * sum (X - x)^2 + reg * sum theta^2  (x scale);
 * sqerr = sum (X - x)^2;  X is [N,T,2]. */
int bode_mlp_sse_grad(const bode_mlp_field* f, const bode_grid* g, int32_t method, int32_t grad_mode, int32_t N,
                      const float* y0, int32_t y0_batched, const float* X, float lik_w, float reg, float scale,
                      int32_t add_prior, float* loss, float* sqerr, float* gtheta, int64_t gtheta_stride,
                      float* scratch, size_t scratch_floats, bode_stream_t stream);

/* Optional DEVICE-resident control block: when non-NULL its fields override the scalar arguments of the same name,
 * so a captured CUDA graph can be replayed while the lr schedule / step counter / phase flags change. */
typedef struct bode_sampler_ctl {
  float    lr;        /* langevin.py:205-210 get_lr(t); for bode_axpy: alpha */
  uint32_t step;      /* Philox stream index */
  int32_t  burn_in;   /* aSGHMC */
  int32_t  resample;  /* aSGHMC */
  uint32_t next_iter; /* global iteration index the next bode_sampler_schedule call will publish */
  uint32_t pad[3];
} bode_sampler_ctl;

/* Device-side schedule: publishes iteration it = ctl->next_iter into the control block and advances it, so a captured
 * graph replays with NO host input:   step = it ;  lr = kind==1 ? lr0 / (t0 + alpha it)^gamma (langevin.py:205-210) : lr0 ;
 * burn_in = it < burn_in_iters ;  resample = !burn_in && resample_every > 0 && (it+1) % resample_every == 0 (hamiltonian.py:81-83). */
int bode_sampler_schedule(bode_sampler_ctl* ctl, int32_t kind, double lr0, double gamma, double t0, double alpha,
                          uint32_t burn_in_iters, uint32_t resample_every, bode_stream_t stream);

/* SGLD.step, samplers/langevin.py:173-202:  p <- p - lr (g + xi / sqrt(lr/2)) */
int bode_sgld_step(float* p, const float* g, const float* xi, int64_t n, float lr, int32_t add_noise,
                   uint64_t seed, uint32_t step, int32_t* status, const bode_sampler_ctl* ctl, bode_stream_t stream);

/* pSGLD.step, samplers/langevin.py:457-500:  V <- a V + (1-a) g^2 ; G = 1/(lambda + sqrt V) ;
 * p <- p - lr (G g + sqrt(G) xi / sqrt(lr/2)) */
int bode_psgld_step(float* p, const float* g, float* V, const float* xi, int64_t n, float lr, float alpha,
                    float lambda, int32_t add_noise, uint64_t seed, uint32_t step, int32_t* status,
                    const bode_sampler_ctl* ctl, bode_stream_t stream);

/* aSGHMC.step, samplers/hamiltonian.py:38-99 (state init tau=gbar=vhat=1, mom=0 is the caller's, :55-60).
 * burn_in: adapt tau/gbar/vhat (:73-77, tau_inv from the OLD tau :70); resample: momentum <- N(0, min(1/minv, 10))
 * from xi_resample (:81-83; the `iteration % k == 0` test is evaluated by the host). */
int bode_asghmc_step(float* p, const float* g, float* tau, float* gbar, float* vhat, float* mom, const float* xi,
                     const float* xi_resample, int64_t n, float lr, float mom_decay, float lambda, int32_t burn_in,
                     int32_t resample, int32_t add_noise, uint64_t seed, uint32_t step, int32_t* status,
                     const bode_sampler_ctl* ctl, bode_stream_t stream);

/* HAMCMC, samplers/langevin.py:619-1107 (L-BFGS-preconditioned Langevin), one CTA per chain, bug-compatible.
 * State buffers (caller-owned, zero-initialised, sizes from bode_hamcmc_floats; meta is int32 [P][4]):
 *   hist_theta/hist_grad [P][2M-1][d] (M = memory+1, :645), pair_s/pair_y [P][M-1][d], work [P][4(M-1)+2][d].
 * metric_step = 0: step_without_metric (:941-964), add_params stores (theta_new, grad) and builds the start-up pairs when the
 *   2M-1 window fills (:902-939);  metric_step = 1: step (:966-1000) = _compute_vector_prod (:717-860) + _update_metric_vars
 *   (:862-900).  The warm-up schedule (first 2M-1+100 burn-in iterations, add_params from 100; :1068-1069) is the host's. */
size_t bode_hamcmc_floats(int32_t P, int32_t d, int32_t memory, int32_t which);
int bode_hamcmc_step(int32_t P, int32_t d, int32_t memory, float* hist_theta, float* hist_grad, float* pair_s,
                     float* pair_y, float* work, int32_t* meta, float* theta, int64_t ld_theta, const float* grad,
                     int64_t ld_grad, const float* xi, float lr, float H_gamma, float trust_reg, int32_t metric_step,
                     int32_t add_params, int32_t add_noise, uint64_t seed, uint32_t step, int32_t* status,
                     bode_stream_t stream);

/* HAMCMC2 / HAMCMC3 / HAMCMC4, samplers/langevin.py:1109-1470: the variants that form (s, y) from contiguous samples (same
 * _compute_vector_prod, different window bookkeeping and base point; variant = 2, 3 or 4).  State buffers (caller-owned,
 * zero-initialised, sizes from bode_hamcmc_contig_floats; meta int32 [P][4]): hist_theta/hist_grad [P][M][d] (M = memory+1),
 * pair_s/pair_y [P][M-1][d], work [P][4(M-1)+2][d].  metric_step = 0: step_without_metric (:1180-1203), update_metric stores
 * (theta_new, grad) and forms the pairs when the M-entry window fills (the first M iterations of sample(), :1254);
 * metric_step = 1: step (:1205-1238, :1364-1397).
 * A metric step while the window is not full yet changes nothing and sets bit 2 of *status (bit 1: non-finite parameter). */
size_t bode_hamcmc_contig_floats(int32_t P, int32_t d, int32_t memory, int32_t which);
int bode_hamcmc_contig_step(int32_t variant, int32_t P, int32_t d, int32_t memory, float* hist_theta, float* hist_grad,
                            float* pair_s, float* pair_y, float* work, int32_t* meta, float* theta, int64_t ld_theta,
                            const float* grad, int64_t ld_grad, const float* xi, float lr, float H_gamma, float trust_reg,
                            int32_t metric_step, int32_t update_metric, int32_t add_noise, uint64_t seed, uint32_t step,
                            int32_t* status, bode_stream_t stream);

/* MALA.accept_or_reject, samplers/langevin.py:57-95, for P chains of d parameters (rows of theta).  The proposal itself is
 * bode_sgld_step (langevin.py:27-54 is the SGLD update).  Per chain:
 *   log_alpha = loss_prev - loss_new - |theta_prev - theta + lr grad_new|^2 / (4 lr) + |theta - theta_prev + lr grad_prev|^2 / (4 lr)
 *   accepted  = isfinite(log_alpha) && log_u < log_alpha           (log_u = log of a U(0,1) draw; NULL: Philox (seed, step))
 * theta_prev == NULL reproduces the reference as it runs: its saved state is a view of the parameter that the proposal then
 * updates in place (:45, :60), so both proposal terms see theta_prev == theta and a rejection restores nothing.  With
 * theta_prev != NULL the textbook ratio is used and, if restore != 0, rejected chains are copied back into theta. */
int bode_mala_accept(const float* theta_prev, int64_t ld_prev, float* theta, int64_t ld_theta, const float* grad_prev,
                     int64_t ld_gprev, const float* grad_new, int64_t ld_gnew, const float* loss_prev, const float* loss_new,
                     const float* log_u, int32_t P, int32_t d, float lr, int32_t restore, uint64_t seed, uint32_t step,
                     float* log_alpha, int32_t* accepted, bode_stream_t stream);

/* p <- p + alpha x   (the SVGD particle update: the optimiser wrapped by stein.py:37-106 descends -phi) */
int bode_axpy(float* p, const float* x, float alpha, int64_t n, int32_t* status, const bode_sampler_ctl* ctl,
              bode_stream_t stream);

/* out[i] ~ N(0,1) from the same Philox stream the samplers use (testing / momentum initialisation) */
int bode_fill_normal(float* out, int64_t n, uint64_t seed, uint32_t step, bode_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * SVGD interaction, samplers/stein.py:12-34 (RBFKernel with median heuristic) and :75-86 (phi), for the
 * LOCAL rows of a particle-sharded job against all gathered particles (columns).  Call order per step:
 *   bode_svgd_sqdist -> for pass in 0..2: bode_svgd_hist_pass, [all-reduce *hist_out over ranks], bode_svgd_select_digit
 *   -> bode_svgd_gamma -> bode_svgd_phi.       Everything stays on `stream`; no host synchronisation.
 * The median is the exact order statistic of the fp32 d2 values (mean of the two middle ones for an even count,
 * like np.median, stein.py:25-26), chosen by an integer radix select.
 * ------------------------------------------------------------------------------------- */
size_t bode_svgd_workspace_bytes(int32_t n_rows, int32_t n_cols, int32_t d);
int bode_svgd_sqdist(const float* Xrows, int64_t ld_rows, int32_t n_rows, const float* Xcols, int64_t ld_cols,
                     int32_t n_cols, int32_t d, int32_t row_offset, uint64_t total_entries, void* workspace,
                     size_t workspace_bytes, void** hist_out, bode_stream_t stream);
/* row_offset = index of local row 0 among the columns (rank * P_local): d2 is forced to 0 where row + row_offset == column, like
 * cdist(x, x); -1 (any negative value) when rows and columns are unrelated sets: nothing is forced.
 * bode_svgd_set_tensor_cores(1) (default) runs the two contractions as 3xTF32 tcgen05.mma (d <= 56); 0 selects the FP32-pipe
 * kernels.  Returns the previous setting. */
int bode_svgd_set_tensor_cores(int32_t on);
/* Persistent selection state (svgd_state.cuh).  Consecutive SVGD steps move the median of d2 by far less than 0.2 %, so
 * bode_svgd_sqdist also counts, in its Gram epilogue, the entries below a window of +-16384 ulps around the previous
 * call's median and histograms the raw fp32 bit patterns inside it (one counter per representable float: still an exact
 * order-statistic selection, np.median semantics of stein.py:25-26).  bode_svgd_window_select reads the median off that
 * table when both middle ranks fall inside the window; the three radix passes below then return immediately.  A miss
 * (first call, large step) falls back to them transparently.  Multi-rank callers all-reduce (sum) the table returned by
 * bode_svgd_window_table between bode_svgd_sqdist and bode_svgd_window_select.  bode_svgd_workspace_init zeroes the state
 * once after the workspace is allocated.  Call order per step:
 *   bode_svgd_sqdist -> [all-reduce table] -> bode_svgd_window_select -> radix passes (no-ops on a hit) -> bode_svgd_gamma -> bode_svgd_phi */
int bode_svgd_workspace_init(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, size_t workspace_bytes, bode_stream_t stream);
int bode_svgd_window_table(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, void** table_out, size_t* count_out);
int bode_svgd_window_select(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, bode_stream_t stream);
/* Disarm the median window (next selection = the 3-pass radix select over the stored d2).  For measuring that path and for
 * callers that replace the particles wholesale; every rank of a peer-mapped job must call it. */
int bode_svgd_window_disarm(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, bode_stream_t stream);
/* single-rank: the three radix passes fused into one cooperative launch (a no-op after a window hit); with med_gamma != NULL it
 * also writes med_gamma[0] = median, med_gamma[1] = gamma (median heuristic, n = n_total) and arms the next window, i.e. it
 * replaces the bode_svgd_gamma call */
int bode_svgd_radix_fallback(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, int32_t n_total, float* med_gamma,
                             bode_stream_t stream);
int bode_svgd_hist_pass(int32_t pass, int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, bode_stream_t stream);
int bode_svgd_select_digit(int32_t pass, int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, bode_stream_t stream);
int bode_svgd_gamma(int32_t n_total, float sigma, int32_t n_rows, int32_t n_cols, int32_t d, void* workspace,
                    float* med_gamma, bode_stream_t stream);
/* Scols = score_sign * (grad log p): pass the loss gradient as is with score_sign = -1. */
int bode_svgd_phi(const float* Xrows, int64_t ld_rows, int32_t n_rows, const float* Xcols, int64_t ld_xc,
                  const float* Scols, int64_t ld_sc, float score_sign, int32_t n_cols, int32_t d, int32_t n_total,
                  const float* med_gamma, void* workspace, float* phi, int64_t ld_phi, float* theta,
                  int64_t ld_theta, float step, bode_stream_t stream);

/* Split forms of bode_svgd_sqdist / bode_svgd_phi for callers that overlap the operand preparation with other work on a second
 * stream (the reference computes everything serially inside RBFKernel.forward / SVGD.phi, stein.py:18-34, 75-86):
 *   bode_svgd_sqdist_staged(BODE_SVGD_PREPARE)  column means, selection-state reset, pre-split centred operands -- needs only the
 *                                               particle positions, so it can run beside the ODE solve that produces the scores;
 *   bode_svgd_phi_staged(BODE_SVGD_PREPARE)     the V = [-G | X - mu | 1] operand -- needs positions and scores but neither d2 nor
 *                                               gamma, so it can run beside the Gram kernel and the median selection;
 *   ..._staged(BODE_SVGD_COMPUTE)               the rest.  PREPARE | COMPUTE equals the plain call.  Where
 * bode_svgd_staged_supported(n_cols, d) returns 0 (shapes outside the pipelined tensor-core kernels) PREPARE alone does nothing
 * and COMPUTE alone does all the work, so callers need no second code path. */
#define BODE_SVGD_PREPARE 1
#define BODE_SVGD_COMPUTE 2
#define BODE_SVGD_PREPARE_POSITIONS 4   /* phi_staged only: the [X - mu | 1] columns of V (no scores needed), see bode_svgd_arm_score_tiles */
int bode_svgd_staged_supported(int32_t n_cols, int32_t d);
/* Layout of the d2 block at the start of the workspace (internal to sqdist -> median -> phi; exposed for tests and for
 * RBFKernel.forward, stein.py:22-32, which returns the kernel matrix).  0: row-major [n_rows][n_cols].  1: tiles
 * [n_rows / 128][n_cols / 32][128][32] -- chosen by the pipelined tensor-core kernels when n_rows and n_cols are multiples of 128, so that
 * every 16 KB tile the K@V pass consumes is contiguous in HBM.  The layout must not change between a bode_svgd_sqdist and the
 * bode_svgd_phi that consumes it (i.e. no bode_svgd_set_tensor_cores in between). */
int bode_svgd_d2_tiled(int32_t n_rows, int32_t n_cols, int32_t d);
/* CTA granularity of the Gram kernel: column_splits CTAs per 128-row block (0 = automatic: one wave over all SMs).  A finer
 * split shortens the tail when the Gram pass shares the GPU with the fused solve.  Returns the previous setting. */
int bode_svgd_set_gram_split(int32_t column_splits);
/* Upper bound on the CTAs of bode_svgd_radix_fallback's cooperative launch (0 = one per SM slot; returns the previous bound).  A
 * cooperative grid starts only when all its CTAs can be resident, so beside a kernel that fills most SMs a full-size grid waits for
 * that kernel to end, even when the launch is the no-op it is after a window hit. */
int bode_svgd_set_select_ctas(int32_t n);
int bode_svgd_sqdist_staged(int32_t stages, const float* Xrows, int64_t ld_rows, int32_t n_rows, const float* Xcols, int64_t ld_cols,
                            int32_t n_cols, int32_t d, int32_t row_offset, uint64_t total_entries, void* workspace,
                            size_t workspace_bytes, void** hist_out, bode_stream_t stream);
int bode_svgd_phi_staged(int32_t stages, const float* Xrows, int64_t ld_rows, int32_t n_rows, const float* Xcols, int64_t ld_xc,
                         const float* Scols, int64_t ld_sc, float score_sign, int32_t n_cols, int32_t d, int32_t n_total,
                         const float* med_gamma, void* workspace, float* phi, int64_t ld_phi, float* theta,
                         int64_t ld_theta, float step, bode_stream_t stream);
/* Score-tile fusion (single rank).  stein.py:75-86 consumes score = -grad loss of every particle, the LAST thing the closure's fused
 * solve produces; between bode_svgd_arm_score_tiles and bode_svgd_disarm_score_tiles every bode_npde_nlp_grad launch of the
 * component-split kernels (3x3 .. 6x6 grids) over exactly n_cols particles with 2m + 2 == d parameters also writes
 * score_sign * gradient into this workspace's phi operand tiles, so that the interaction needs only
 * bode_svgd_phi_staged(BODE_SVGD_PREPARE_POSITIONS) -- any time after the operands' PREPARE, Scols may be NULL -- and
 * bode_svgd_phi_staged(BODE_SVGD_COMPUTE): no operand launch between the solve and phi.  arm returns 1 when armed, 0 when the
 * shapes are outside the pipelined tensor-core path or the workspace has peers (nothing changes then); disarm returns how many
 * closure launches wrote the tiles since arm (0: run BODE_SVGD_PREPARE as usual).  Process-wide state, like bode_npde_set_cta_limit. */
int bode_svgd_arm_score_tiles(int32_t n_rows, int32_t n_cols, int32_t d, void* workspace, float score_sign);
int bode_svgd_disarm_score_tiles(void);

/* ---------------------------------------------------------------------------------------
 * Peer-mapped SVGD workspaces (ranks of one NVLink / NVSwitch node).  The reference is a single-process program; this is what
 * replaces the "tiny histogram all-reduce" of SURVEY.md 8(e): every rank allocates its workspace with bode_peer_alloc, exports
 * it (64-byte CUDA IPC handle, exchanged by the host layer), imports the others and registers the addresses with
 * bode_svgd_set_peers.  bode_svgd_window_select / bode_svgd_radix_fallback then read the peers' window tables and radix
 * histograms over NVLink and meet at flag barriers inside the kernels: the exact distributed median costs no collective launch,
 * and the call order of a rank is the single-rank one (bode_svgd_sqdist -> bode_svgd_window_select -> bode_svgd_radix_fallback
 * -> bode_svgd_phi).  Every rank must issue the same sequence of these calls.
 * ------------------------------------------------------------------------------------- */
int bode_peer_alloc(size_t bytes, void** out);
int bode_peer_free(void* ptr);
int bode_peer_export(void* ptr, void* handle64);
int bode_peer_import(const void* handle64, void** out);
int bode_peer_release(void* imported);
int bode_svgd_set_peers(void* workspace, int32_t n_rows, int32_t n_cols, int32_t d, void* const* bases, int32_t rank, int32_t world);
/* The flag barriers give up after ~2^24 polls instead of hanging when a peer never arrives; *timed_out = 1 if that has happened on
 * this workspace since bode_svgd_workspace_init (the results of that step are invalid).  Blocking device-to-host copy: synchronise
 * the streams that use the workspace first. */
int bode_svgd_peer_status(void* workspace, int32_t n_rows, int32_t n_cols, int32_t d, int32_t* timed_out);
/* The data-path exchange itself over peer memory (the all-gather of particle positions and scores, which the
 * reference, a single-process program, does not have): every rank pushes its rows[n_rows, d] into the gather buffer of every
 * rank's workspace between two flag barriers, in ONE launch and without a collective.  which = 0 positions, 1 scores (separate
 * buffers and barrier flags, so the two gathers may run on different streams).  *gathered_out = the [n_cols, d] rank-major
 * buffer in THIS rank's workspace, complete when the launch retires.  Needs bode_svgd_set_peers and n_cols == world * n_rows;
 * every rank issues the same sequence of gathers per buffer. */
int bode_svgd_peer_gather(int32_t which, const float* rows, int64_t ld, int32_t n_rows, int32_t n_cols, int32_t d, void* workspace,
                          float** gathered_out, bode_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BODE_B200_H */
