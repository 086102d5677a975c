#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native bayesian-ode hot path.

Metric (BASELINE.json): particle*RK-steps/sec, forward + gradient, sampler update and SVGD interaction inside the timed
region.  One particle*RK-step = all N trajectories of one particle advanced by one RK step and differentiated (for adaptive
dopri5: one ATTEMPTED step, six RHS evaluations).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5|c1] [--scaling weak|strong] [--impl b200|reference]

Workloads (sizes are per GPU under the default weak scaling):
  c3 (default)  Van der Pol npde, 5x5 inducing grid, SVGD, 4096 particles, N=5, T=40 -> 39 rk4 (3/8) steps
  c2            Van der Pol npde, pSGLD (loss / N), 1024 independent chains, T=101 -> 100 rk4 steps
  c4            2-64-64-2 ELU MLP neural ODE, aSGHMC, 8192 chains, adaptive dopri5 (--tol tight: 1e-7/1e-9, loose: 1e-5/1e-7)
  c5            Van der Pol npde, 16x16 inducing grid (d = 514), HAMCMC memory 5, 2048 chains, 39 rk4 steps
  c1            1 chain SGLD (the reference's own CPU-runnable case; parity-sized)
One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE).  Independent chains (c1, c2, c4, c5) shard with no data-path
communication at all; SVGD's one exchange is the all-gather of positions and scores (push kernels over NVLink peer memory; NCCL
when the workspaces cannot be peer-mapped).

--impl reference times the UNMODIFIED reference (baseline/_ref, staged by baseline/make_ref.py) on the host cores, see
baseline/ref_arm.py; when baseline/_ref is absent it falls back to the torch port of the reference's execution model
(oracle/ref_torch.py, "kind": "port").
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c3": dict(field="npde", sampler="svgd", P=4096, M=5, T=40, method="rk4",
               desc="VDP npde SVGD, 5x5 grid, 4096 particles/GPU, N=5, T=40 (39 rk4 3/8 steps)"),
    "c2": dict(field="npde", sampler="psgld", P=1024, M=5, T=101, method="rk4",
               desc="VDP npde pSGLD, 5x5 grid, 1024 chains/GPU, N=5, T=101 (100 rk4 3/8 steps)"),
    "c1": dict(field="npde", sampler="sgld", P=1, M=5, T=40, method="rk4",
               desc="VDP npde SGLD, 5x5 grid, 1 chain, N=5, T=40 (39 rk4 3/8 steps)"),
    "c4": dict(field="mlp", sampler="asghmc", P=8192, H=64, T=40, method="dopri5",
               desc="neural ODE 2-64-64-2 ELU MLP, aSGHMC, 8192 chains/GPU, N=5, T=40, adaptive dopri5 batched-step (one controller per chain, "
                    "error pooled over its 5 trajectories = one reference odeint call with y0 [5, 2])"),
    "c5": dict(field="npde", sampler="hamcmc", P=2048, M=16, T=40, method="rk4", ell=0.35,
               desc="VDP npde HAMCMC (memory 5), 16x16 grid (d=514), 2048 chains/GPU, N=5, T=40 (39 rk4 3/8 steps)"),
}
STAGES = 4
FLOP_PER_EVAL = 23.0               # per (RK stage, trajectory, inducing point), fwd 11 + adjoint 12
MLP_FLOP_PER_STAGE_TRAJ = 26112.0  # h = 64, fwd 8704 + adjoint 2 x
SAMPLER_BYTES = {"sgld": 12, "psgld": 20, "asghmc": 44, "hamcmc": None}      # per parameter and step
TOLS = {"tight": (1e-7, 1e-9), "loose": (1e-5, 1e-7)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 1000 for c1-c3, 20 for c4/c5; reference arm: 4)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--particles", type=int, default=None, help="override particles per GPU")
    ap.add_argument("--tol", default="tight", choices=sorted(TOLS), help="c4: dopri5 tolerances")
    ap.add_argument("--init-scale", type=float, default=0.3,
                    help="c4: scale of the U(-0.5, 0.5) weight draw of nn.ipynb cell 4.  The notebook's net has H = 20; the same draw at H = 64 "
                         "has sqrt(64/20) = 1.8x the layer gain, its trajectories grow like e^(5t) and aSGHMC at the notebook's lr = 1e-2 turns the "
                         "chains non-finite within a few iterations (in the reference as well).  0.3 ~ sqrt(20/64) / 2 keeps the ensemble bounded.")
    ap.add_argument("--no-mlp-tc", action="store_true", help="c4: FP32-pipe MLP kernels instead of the tensor-core ones (comparison runs)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="c3, N>1: exchange of positions / scores")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the strong-scaling sub-record")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the multi-GPU parity check")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--side-sms", type=int, default=40, help="SVGD: SMs kept free of the fused solve for the position-only side chain")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 4 if a.impl == "reference" else (20 if a.workload in ("c4", "c5") else 1000)
    return a


def config_of(args, wl, world):
    """The workload description both arms print (same keys, so the driver's same_config check can compare them)."""
    P_gpu = args.particles or wl["P"]
    if args.scaling == "strong":
        P_gpu = max(1, P_gpu // world)
    cfg = {"workload": wl["desc"], "particles_per_gpu": P_gpu, "total_particles": P_gpu * world, "trajectories": 5,
           "sampler": wl["sampler"], "solver": wl["method"], "scaling": args.scaling}
    if wl["field"] == "npde":
        cfg.update(inducing_grid="%dx%d" % (wl["M"], wl["M"]), rk_steps=wl["T"] - 1)
    else:
        cfg.update(hidden=wl["H"], rtol=TOLS[args.tol][0], atol=TOLS[args.tol][1], init_scale=args.init_scale,
                   rk_steps="attempted dopri5 steps, counted")
    return cfg


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_baseline(args, wl, P_total, steps, warmup=1):
    """The reference's CPU path on this box's host cores.  baseline/_ref present: the unmodified reference (baseline/ref_arm.py,
    kind "reference"); otherwise the torch float64 port of its execution model (oracle/ref_torch.py, kind "port", c1-c3 only)."""
    from baseline import ref_arm
    if ref_arm.available():
        over = None
        if wl["field"] == "mlp":
            over = dict(rtol=TOLS[args.tol][0], atol=TOLS[args.tol][1], init_scale=args.init_scale)
        r = ref_arm.run(args.workload, P_total, steps=steps, warmup=warmup, wl_override=over)
        r["unit"] = "particle*RK-steps/s"
        return r
    if wl["field"] != "npde" or wl["sampler"] == "hamcmc":
        return dict(value=None, unit="particle*RK-steps/s", cores=os.cpu_count(), kind="port",
                    sample="unavailable: baseline/_ref is not staged and the port covers c1-c3 only (run baseline/make_ref.py)")
    import multiprocessing as mp
    from oracle import ref_torch
    C = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(C) as pool:
        per_step = pool.map(ref_torch.time_chain, [(i, wl["T"], wl["M"], 2, steps) for i in range(C)])
    rk_steps = wl["T"] - 1
    chain_rate = sum(rk_steps / s for s in per_step)
    sample = "%d processes x (2 warm-up + %d timed) SGLD steps of one chain each, fp64, 1 thread per process" % (C, steps)
    value = chain_rate
    if wl["sampler"] == "svgd":
        t_phi, _ = ref_torch.phi_numpy_seconds(n=min(P_total, 4096), d=2 * wl["M"] ** 2 + 2)
        t_solve = P_total * rk_steps / chain_rate
        value = P_total * rk_steps / (t_solve + t_phi * (P_total / min(P_total, 4096)) ** 2)
        sample += "; + one numpy phi over %d particles (%.3f s)" % (min(P_total, 4096), t_phi)
    return dict(value=value, unit="particle*RK-steps/s", cores=C, kind="port", sample=sample,
                per_process_ms_per_step=round(1e3 * statistics.mean(per_step), 2), wall_s=round(time.perf_counter() - t0, 1))


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm may use every host thread (the per-chain workers pin
    # themselves to one thread each, the phi block runs on all of them).  Must happen before numpy / torch are imported.
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    world = args.gpus
    cfg = config_of(args, wl, world)
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    cb = cpu_baseline(args, wl, cfg["total_particles"], steps, warmup=warm)
    line = {
        "impl": "reference", "metric": "particle*RK-steps/sec (fwd+grad)", "value": cb["value"], "unit": cb["unit"],
        "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * cb["seconds_per_whole_step"] if "seconds_per_whole_step" in cb else cb.get("per_process_ms_per_step"),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ helpers (GPU arm)
def checked(smp):
    """smp.check(): ANY report fails the run -- NaN/Inf parameters (ValueError, langevin.py:184-185), a peer flag barrier of the
    multi-GPU exchange that gave up on a rank (the gathered particles of that step were stale), a CUDA error.  A throughput
    measured over invalid particles is not a measurement."""
    smp.check()


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, n = [], [], set(), 0
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            n += 1
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hot = sorted(sm)[len(sm) // 2:]                      # upper half = samples under load
        return {"sm_mhz": statistics.median(hot), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": n}


def event_ms(torch, fn, iters, flush=None):
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return ts


def graph_seg_ms(torch, fn, reps=20, iters=7, flush=None, graph=True):
    """Device time of one call of `fn` (a short sequence of kernel launches): `reps` back-to-back calls are captured in ONE
    CUDA graph and the replay is timed with events, so neither Python/ctypes launch overhead nor the graph-launch latency
    (~8 us per replay, measured with an empty graph) is attributed to the kernels.  `flush` runs before each replay."""
    fn()
    torch.cuda.synchronize()
    if not graph:
        return statistics.median(event_ms(torch, fn, iters, flush))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts) / reps


def measure_peaks(torch, bode):
    lib = bode._lib.load()
    sms = lib.bode_device_sm_count()
    scratch = torch.zeros(16, device="cuda")
    out = {}
    for kind, name, per_thread in ((0, "fp32_fma_tflops", 32.0), (1, "mufu_ex2_gops", 8.0)):
        iters, ctas = 4096, sms * 8
        fn = lambda: bode._lib.check(lib.bode_peak_kernel(kind, ctas, iters, bode._lib.ptr(scratch), bode._lib.stream_ptr()))
        fn()
        ms = min(event_ms(torch, fn, 5))
        out[name] = ctas * 256 * iters * per_thread / (ms * 1e-3) / (1e12 if kind == 0 else 1e9)
    return out


def measured_peaks_file():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ one configured job on this rank
class Job:
    """Field + posterior + sampler of one workload on this rank, and `step()` = one sampler iteration through the public API."""

    def __init__(self, args, wl, P_gpu, rank, world, torch, bode):
        from bayesian_ode_b200 import problems
        from bayesian_ode_b200.samplers import HAMCMC, SGLD, SVGD, aSGHMC, pSGLD
        self.wl, self.P, self.rank, self.world, self.torch, self.bode = wl, P_gpu, rank, world, torch, bode
        T = wl["T"]
        data = problems.make_dataset("VDP", seed=0, N=5, R=3.0, T=T, t_end=7.0, noise=0.1)
        self.N = data["N"]
        self.x0_host = data["x0"].float().pin_memory()
        gen = torch.Generator().manual_seed(1234 + rank)
        self.sched = None
        self.dopri5 = wl["method"] == "dopri5"
        if wl["field"] == "npde":
            M = wl["M"]
            ell = wl.get("ell", 0.75)
            Z = problems.inducing_grid(data["Y"], M)
            U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, ell)
            U = U0[None] + 0.1 * torch.randn(P_gpu, M * M, 2, generator=gen, dtype=torch.float64)       # gp.py:321 init scale
            # 16x16: cond(Kzz) ~ 1e14 at ell = 0.75 (SURVEY.md hard part 3) -> ell = 0.35 and triangular solves for the constants
            self.field = bode.NPDEField(U, Z, 1.0, ell, 0.1, **({"stable_solve": True} if M > 6 else {}))
            self.obs_host = torch.from_numpy(data["Y"]).float().pin_memory()
            self.post = bode.NPDEPosterior(self.field, data["x0"], data["t"], torch.from_numpy(data["Y"]), method="rk4", grad_mode="discrete")
            self.S = T - 1
            params = [self.field.U, self.field.logsn]
        else:
            self.field = bode.MLPField(P_gpu, hidden_size=wl["H"], generator=gen)
            if args.init_scale != 1.0:
                with torch.no_grad():
                    self.field.theta.mul_(args.init_scale)
            rtol, atol = TOLS[args.tol]
            self.obs_host = torch.from_numpy(data["X"]).float().pin_memory()
            self.post = bode.MLPPosterior(self.field, data["x0"], data["t"], torch.from_numpy(data["X"]), method="dopri5", rtol=rtol,
                                          atol=atol, reg=0.5, options=dict(controller="batch"))     # BASELINE config 4: "dopri5 batched-step"
            self.post.check_status = False                 # no host sync per step; the solver status is checked after the run
            self.S = None
            params = list(self.field.parameters())
        self.field.bind_flat_grads()
        s = wl["sampler"]
        if s == "svgd":
            self.smp = SVGD(params, lr=1e-4, gather_comm=args.gather, side_sms=args.side_sms)
        elif s == "psgld":
            self.smp = pSGLD(params, lr0=5e-3, lr_gamma=0.51, lr_t0=100, lr_alpha=0.1, lambda_=1e-8, alpha=0.99, N=self.N, seed=7 + rank)
            self.post.scale = 1.0 / self.N
            self.sched = dict(kind=1, lr0=5e-3, gamma=0.51, t0=100, alpha=0.1)
        elif s == "sgld":
            self.smp = SGLD(params, lr0=1e-4, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03, seed=7 + rank)
            self.sched = dict(kind=1, lr0=1e-4, gamma=0.51, t0=100, alpha=0.03)
        elif s == "asghmc":
            self.smp = aSGHMC(params, lr=1e-2, mom_decay=5e-2, lambda_=1e-5, seed=7 + rank)             # nn.ipynb cell 11
        else:
            # gp.ipynb cell 23's schedule shape with a step size the 16x16 posterior tolerates (curvature ~1e7)
            self.smp = HAMCMC(params, memory=5, lr0=1e-7, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3, H_gamma=1.0, trust_reg=1.0, seed=7 + rank)
        self.smp.check_finite = "deferred"
        self.it = 0
        self.graphable = s in ("svgd", "psgld", "sgld")

    def step(self):
        s = self.wl["sampler"]
        if s == "svgd":
            self.smp.prefetch()              # position-only half of the interaction (operands, Gram, median) beside the ODE kernel
        self.post.loss_and_grad_()
        if s == "svgd":
            self.smp.phi(update_lr=self.smp.param_groups[0]["lr"])
        elif s in ("psgld", "sgld"):
            self.smp.schedule_(**self.sched)
            self.smp.step(use_ctl=True)
        elif s == "asghmc":
            self.smp.step(lr=1e-2, burn_in=True)
        else:
            M_ = self.smp.memory
            lr = self.smp.get_lr(self.it)
            if self.it < 2 * M_ - 1:
                self.smp.step_without_metric(lr=lr, add_params=True)         # fill the 2M-1 window (langevin.py:1068-1069)
            else:
                self.smp.step(lr=lr)
        self.it += 1

    def units_per_step(self):
        """particle*RK-steps of ONE step on this rank (dopri5: attempted steps of the last launch, summed over pairs / N)."""
        if not self.dopri5:
            return float(self.P * self.S)
        st = self.bode.last_dopri5_stats()
        return float((st[..., 0] + st[..., 1]).sum().item()) / self.N

    def final_checks(self):
        if self.wl["sampler"] == "hamcmc":
            # the reference's HAMCMC diverges on part of the chains of this posterior (langevin.py:846 quirk; the float64 oracle does the
            # same, tests/test_config_sizes_gpu.py): non-finite chains are reported, not an error of the run
            self.smp._status.zero_()
        else:
            checked(self.smp)
        if self.dopri5:
            st = self.bode.last_dopri5_stats()
            bad = int((st[..., 2] != 0).sum().item())
            assert bad == 0, "dopri5 status flags set on %d (particle, trajectory) pairs" % bad
        if self.wl["sampler"] != "hamcmc":
            assert bool(self.torch.isfinite(self.post.loss).all()), "non-finite loss at the end of the run"

    def finite_fraction(self):
        return float(self.torch.isfinite(self.field.theta).all(dim=1).float().mean().item())


def timed_region(job, args, torch, dist, world, rank, local, flush, want_clocks):
    """Warm-up (same policy at every N), graph capture where the step is capturable, K event-timed steps.  Returns a dict."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):                                    # eager warm-up: allocates every buffer
        job.step()
    torch.cuda.synchronize()
    use_graph = job.graphable and not args.no_graph
    run = job.step
    if use_graph:
        try:
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                job.step()
            run = gph.replay
        except Exception as e:                                   # e.g. NCCL capture unsupported: fall back to eager launches
            print("cuda graph capture failed (%s); running eagerly" % str(e).splitlines()[0], file=sys.stderr)
            use_graph, run = False, job.step
            torch.cuda.synchronize()
    if world > 1:
        flag = torch.tensor([int(use_graph)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and use_graph:
            use_graph, run = False, job.step
    clocks = ClockSampler(local) if want_clocks else None
    if clocks is not None:
        clocks.start()
    # warm-up, identical at every N: W >= 3 steps, then as many more as fill 0.6 s (so nvidia-smi delivers samples under THIS load
    # and every world size starts the timed region equally warm); the count is rank 0's, broadcast, because the ranks of an SVGD job
    # meet inside the step and must run the same number of steps
    n_pre = max(args.warmup, 3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_pre):
        flush()
        run()
    barrier()
    per = (time.perf_counter() - t0) / n_pre
    extra = torch.tensor([max(0, int(math.ceil((0.6 - per * n_pre) / max(per, 1e-6))))], device="cuda")
    if world > 1:
        dist.broadcast(extra, src=0)
    for _ in range(int(extra.item())):
        flush()
        run()
    n_pre += int(extra.item())
    barrier()
    t_wall = time.perf_counter()
    units = 0.0
    step_ms = []
    for _ in range(args.steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        b.synchronize()
        step_ms.append(a.elapsed_time(b))
        if job.dopri5:
            units += job.units_per_step()
    if not job.dopri5:
        units = job.units_per_step() * args.steps
    barrier()
    t_wall = time.perf_counter() - t_wall
    clk = clocks.stop() if clocks is not None else None
    if clk is not None:
        clk["window"] = "sampled every 100 ms from %d warm-up steps of the same %s through the %d timed steps" % (
            n_pre, "graph" if use_graph else "launch sequence", args.steps)
    total_ms = sum(step_ms)
    tt = torch.tensor([total_ms, units], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        total_ms, units = float(mx[0].item()), float(tt[1].item())
    job.final_checks()
    return dict(total_ms=total_ms, units=units, ms_per_step=total_ms / args.steps, value=units / (total_ms * 1e-3), n_pre=n_pre,
                use_graph=use_graph, run=run, clocks=clk, t_wall=t_wall, barrier=barrier)


def svgd_parity(job, torch, dist, world, rank):
    """Multi-GPU correctness the driver can see: ONE extra interaction (no update) on the sharded job, then rank 0 -- holding every
    rank's particles and scores -- recomputes phi for 32 of its rows in float64 torch (checker code: cdist form, np.median
    semantics, stein.py:18-34, 75-86) and the exact median of the KERNEL's own d2 (all ranks' blocks gathered) by selection."""
    smp, post, field = job.smp, job.post, job.field
    smp.prefetch()
    post.loss_and_grad_()
    phi = smp.phi().clone()                                           # no update: theta is what the interaction saw
    X, G = field.theta.detach().clone(), field.theta_grad.detach().clone()
    nl, nt = smp.P_local, smp.n_total
    Xs = [torch.empty_like(X) for _ in range(world)]
    Gs = [torch.empty_like(G) for _ in range(world)]
    dist.all_gather(Xs, X)
    dist.all_gather(Gs, G)
    med = smp._ws.med_gamma.clone()
    meds = [torch.empty_like(med) for _ in range(world)]
    dist.all_gather(meds, med)
    d2_local = smp._ws.d2(nl, nt).contiguous()
    out = None
    gather_d2 = world * nl * nt * 4 <= 40e9                           # the kernel's d2 blocks: <= 4.3 GB at 8 x 4096 particles
    d2_all = [torch.empty_like(d2_local) for _ in range(world)] if (rank == 0 and gather_d2) else None
    if gather_d2:
        dist.gather(d2_local, d2_all, dst=0)
    if rank == 0:
        Xa, Ga = torch.cat(Xs).double(), torch.cat(Gs).double()
        rows = torch.linspace(0, nl - 1, 32, device=X.device).long()
        sq = (Xa * Xa).sum(1)
        # exact median over all n^2 squared distances (float64 Gram form, compared as float32 values), in row blocks
        blocks = []
        for i in range(0, nt, 2048):
            blk = (sq[i:i + 2048, None] + sq[None] - 2.0 * Xa[i:i + 2048] @ Xa.T).clamp_min_(0)
            blocks.append(blk.float().reshape(-1))
        flat = torch.cat(blocks)
        del blocks
        n2 = flat.numel()
        lo = torch.kthvalue(flat, (n2 - 1) // 2 + 1).values.double()
        hi = torch.kthvalue(flat, n2 // 2 + 1).values.double()
        med64 = float(0.5 * (lo + hi))
        del flat
        gamma = 1.0 / (1e-8 + 2.0 * (med64 / (2.0 * math.log(nt + 1.0))))
        Xr = Xa[rows]
        d2r = (sq[rows, None] + sq[None] - 2.0 * Xr @ Xa.T).clamp_min_(0)
        K = torch.exp(-gamma * d2r)
        ref = (K @ (-Ga) + 2.0 * gamma * (K.sum(1, keepdim=True) * Xr - K @ Xa)) / nt
        err = float(((phi[rows].double() - ref).abs().max() / ref.abs().max()).item())
        same = all(torch.equal(m, meds[0]) for m in meds)
        med_rel = abs(float(med[0].item()) - med64) / med64
        bit_exact = None
        if gather_d2:
            kd2 = torch.cat([b.reshape(-1) for b in d2_all])
            del d2_all
            klo = torch.kthvalue(kd2, (kd2.numel() - 1) // 2 + 1).values
            khi = torch.kthvalue(kd2, kd2.numel() // 2 + 1).values
            bit_exact = bool((0.5 * (klo + khi)).item() == med[0].item())
            del kd2
        out = dict(phi_rows_checked=32, phi_max_rel_err=err, phi_tol=2e-5, median_rel_err_vs_fp64=med_rel,
                   median_identical_on_all_ranks=bool(same), median_bit_exact_vs_selection_on_kernel_d2=bit_exact,
                   ok=bool(err < 2e-5 and same and med_rel < 1e-5 and bit_exact is not False))
    torch.cuda.synchronize()
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    import bayesian_ode_b200 as bode

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: everything libraries print (NCCL prints its version there) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world), file=sys.stderr)

    cfg = config_of(args, wl, world)
    P_gpu, P_total = cfg["particles_per_gpu"], cfg["total_particles"]
    if args.no_mlp_tc:
        bode._lib.load().bode_mlp_set_tensor_cores(0)
    job = Job(args, wl, P_gpu, rank, world, torch, bode)
    smp, post, field = job.smp, job.post, job.field
    N = job.N
    sampler = wl["sampler"]
    peaks = measure_peaks(torch, bode)
    mp_file, peak_src = measured_peaks_file()
    hbm_peak = mp_file.get("hbm_gbs") or 6650.0
    lib = bode._lib.load()

    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")              # 256 MiB > 126 MB L2

    def flush():
        flush_buf.zero_()

    tr = timed_region(job, args, torch, dist, world, rank, local, flush, want_clocks=(rank == 0))
    run, barrier, use_graph = tr["run"], tr["barrier"], tr["use_graph"]
    ms_per_step, value = tr["ms_per_step"], tr["value"]
    n_pre_main, clk_main, t_wall_main = tr["n_pre"], tr["clocks"], round(tr["t_wall"], 3)
    finite_frac = job.finite_fraction()

    # ---- per-kernel timing pass for the roofline objects: each C-ABI call sequence is captured 20x in its own CUDA graph
    # (graph_seg_ms), so the figures are device time without launch gaps; inputs stay L2-resident between repetitions.
    kernels = []
    if wl["field"] == "npde":
        M = wl["M"]
        ode_all_ms = graph_seg_ms(torch, lambda: post.loss_and_grad_())              # one CTA on every SM
        ode_ms, ode_cfg = ode_all_ms, "one CTA per SM"
        if sampler == "svgd" and smp.overlap == "gram" and smp.side_sms > 0:
            # the configuration the timed step launches: packed into fewer CTAs so that the Gram pass can run beside it
            old = lib.bode_npde_set_cta_limit(lib.bode_device_sm_count() - smp.side_sms)
            ode_ms = graph_seg_ms(torch, lambda: post.loss_and_grad_())
            lib.bode_npde_set_cta_limit(old)
            ode_cfg = "as launched in the step: at most %d CTAs, timed alone" % (lib.bode_device_sm_count() - smp.side_sms)
        ode_flop = FLOP_PER_EVAL * STAGES * N * M * M * job.S * P_gpu
        kname = ("npde_pair_grad_kernel / npde_grad_kernel<Sep> (fused rk4 solve + closure + discrete adjoint; %s)" % ode_cfg) if M <= 6 else \
            "npde_grad_kernel<RowField> (row-sliced separable field, 16 lanes per (particle, trajectory) pair, m = %d)" % (M * M)
        extra = {}
        if M * M >= 64:
            # the closure is three launches here: W = A U and gU = A^T gW + Ksym U as panel GEMMs around the solve (csrc/npde_proj.cu);
            # ms covers all three.  SURVEY.md 8(d) counts the per-solve projections separately from the 92 N m FLOP per
            # particle-RK-step the roofline is graded on: `achieved` uses the latter only, the projections' 3 x 2 m^2 x 2 FLOP per
            # particle are reported beside it
            proj_flop = 3 * 2 * (M * M) ** 2 * 2 * P_gpu
            kname = "proj_W_kernel + " + kname + " + proj_back_kernel (one closure call, three launches)"
            extra = dict(projection_flop_not_counted=proj_flop)
        kernels.append(dict(name=kname, ms=ode_ms, bound="fp32", achieved=ode_flop / (ode_ms * 1e-3) / 1e12, peak=peaks["fp32_fma_tflops"],
                            unit="TFLOP/s", ms_one_cta_per_sm=ode_all_ms, flop_per_launch=ode_flop, **extra))
    else:
        post.loss_and_grad_()
        torch.cuda.synchronize()
        ode_ms = statistics.median(event_ms(torch, lambda: post.loss_and_grad_(), 5))
        att_pairs = float((bode.last_dopri5_stats()[..., :2]).sum().item())
        ode_flop = MLP_FLOP_PER_STAGE_TRAJ * 6.0 * att_pairs
        kernels.append(dict(name=("dopri5_grad_kernel<MlpField<64>> (FP32 pipe; " if args.no_mlp_tc else
                                  "dopri5_grad_kernel<MlpTcField> (64x64 layer on mma.sync tf32 3x split; ") +
                            "adaptive solve, pooled controller + record + frozen-step reverse sweep)",
                            ms=ode_ms, bound="fp32", achieved=ode_flop / (ode_ms * 1e-3) / 1e12, peak=peaks["fp32_fma_tflops"], unit="TFLOP/s",
                            flop_per_launch=ode_flop, attempted_steps_per_pair=att_pairs / (P_gpu * N)))
    if sampler in ("psgld", "sgld", "asghmc", "hamcmc"):
        d = field.d
        if sampler == "hamcmc":
            Kp = float(smp.n_pairs().float().mean().item())
            bpp = (2.0 * Kp + 6.0) * 4.0
            upd = lambda: smp.step(lr=smp.get_lr(job.it))
            note = "hamcmc_kernel (one CTA per chain; mean %.2f curvature pairs per chain -> (2K+6) x 4 B per parameter)" % Kp
        elif sampler == "asghmc":
            bpp = 44.0
            upd = lambda: smp.step(lr=1e-2, burn_in=True)
            note = "sampler_kernel<ASGHMC> burn-in update (44 B per parameter)"
        else:
            bpp = float(SAMPLER_BYTES[sampler])
            upd = lambda: smp.step(use_ctl=True)
            note = "sampler_kernel<%s> (%d B per parameter)" % (sampler.upper(), int(bpp))
        theta_keep = field.theta.detach().clone()
        big_state = bpp * P_gpu * d > 100e6                  # working set beyond L2: flush not needed, every replay streams from HBM
        upd_ms = graph_seg_ms(torch, upd, reps=20, flush=None, graph=job.graphable)
        with torch.no_grad():
            field.theta.copy_(theta_keep)                     # the repeated update walked the chains away: restore
        smp._status.zero_()
        kernels.append(dict(name=note, ms=upd_ms, bound="hbm", achieved=bpp * P_gpu * d / (upd_ms * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s",
                            bytes_per_launch=bpp * P_gpu * d,
                            l2_note="state %.0f MB %s the 126 MB L2" % (bpp * P_gpu * d / 1e6, "exceeds" if big_state else "fits in")))
    median_fallback_ms = None
    if sampler == "svgd":
        import ctypes as C
        d = field.d
        ws = smp._ws
        X = field.theta
        nl, nt = P_gpu, P_total
        Xall = smp._gather_positions(X) if world > 1 else X
        Gall = smp._gather(1, field.theta_grad) if world > 1 else field.theta_grad
        Xloc = Xall[rank * nl:(rank + 1) * nl] if world > 1 else X
        xr, xrs = bode._lib.rows(X, d); xc, xcs = bode._lib.rows(Xall, d); sc, scs = bode._lib.rows(Gall, d)
        sq_fn = lambda: ws.sqdist(Xloc, nl, Xall, nt, d, nt * nt, row_offset=rank * nl)
        phi_fn = lambda: bode._lib.check(lib.bode_svgd_phi(xr, xrs, nl, xc, xcs, sc, scs, -1.0, nt, d, nt, bode._lib.ptr(ws.med_gamma),
                                                           C.c_void_p(ws.base.data_ptr()), bode._lib.ptr(smp.phi_buf), d, None, 0, 0.0,
                                                           bode._lib.stream_ptr()))
        big = nl * nt * 4 > 100e6                          # the d2 block no longer fits in L2: it streams from HBM in every repetition
        sq_ms = graph_seg_ms(torch, sq_fn)
        med_ms = None
        if world == 1:
            def sqmed_fn():
                sq_fn()
                ws.median(nl, nt, d, nt, None, group=None)
            sqmed_ms = graph_seg_ms(torch, sqmed_fn)
            sqmed_fn()                                   # leave the selection state consistent (the repeated sqdist filled the window table)
            med_ms = max(sqmed_ms - sq_ms, 1e-4)
        phi_ms = graph_seg_ms(torch, phi_fn)
        tc = bool(lib.bode_svgd_set_tensor_cores(1)) and d <= 55
        lib.bode_svgd_set_tensor_cores(int(tc))
        tpeak = (mp_file.get("bf16_tflops") or 1590.0) / 2.0          # tf32 dense = half the bf16 rate; no measured tf32 figure exists
        shape = "%d x %d block" % (nl, nt)
        kernels.append(dict(name=("svgd sqdist: prep_x + gram2 (3xTF32 tcgen05, d2 store + median window count), %s" % shape) if tc else "svgd sqdist_kernel",
                            ms=sq_ms, bound="tensor" if tc else "fp32", achieved=2.0 * nl * nt * d / (sq_ms * 1e-3) / 1e12,
                            peak=tpeak if tc else peaks["fp32_fma_tflops"], unit="TFLOP/s", d2_bytes=nl * nt * 4, d2_in_l2=not big))
        if med_ms is not None:
            kernels.append(dict(name="svgd exact median: window_select + radix fallback (no-op after a window hit) + gamma", ms=med_ms, bound="hbm",
                                achieved=(2 * 16384 + 2) * 8 * 2 / (med_ms * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s"))
        kernels.append(dict(name=("svgd phi: prep_v + phi2 K@[S|X|1] (tcgen05: TF32 hi.hi + one bf16 product for both correction terms, TMA d2 tiles) + cluster combine, %s" % shape) if tc else "svgd phi_partial+combine (K@[S|X])",
                            ms=phi_ms, bound="tensor" if tc else "fp32", achieved=4.0 * nl * nt * d / (phi_ms * 1e-3) / 1e12,
                            peak=tpeak if tc else peaks["fp32_fma_tflops"], unit="TFLOP/s"))
        # one step whose median window MISSES (what the first step of a run, or a step that moves the median by > 0.2 %, pays):
        # the window is disarmed on every rank, then the same captured step is replayed once
        barrier()
        miss = []
        for _ in range(3):
            bode._lib.check(lib.bode_svgd_window_disarm(nl, nt, d, C.c_void_p(ws.base.data_ptr()), bode._lib.stream_ptr()))
            barrier()
            miss.append(event_ms(torch, run, 1, flush=flush)[0])
        median_fallback_ms = statistics.median(miss)
        for _ in range(2):
            run()                                         # re-arm the window
        barrier()
    for k in kernels:
        k["frac"] = k["achieved"] / k["peak"]
    dom = max(kernels, key=lambda k: k["ms"])
    # DRAM traffic of the dominant kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of the committed
    # `ncu --set full` capture (profiles/traffic_r*.json, written by tools/ncu_summary.py); not measurable from here
    traffic = None
    keys = {"npde_pair": "npde_pair_grad_kernel", "RowField": "npde_grad_kernel<RowField", "dopri5": "dopri5_grad_kernel", "sqdist": "gram2_kernel",
            "svgd phi": "phi2_kernel", "sampler_kernel": "sampler_kernel", "hamcmc": "hamcmc_kernel"}
    for tname in ("traffic_r02.json", "traffic_r01_final.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if traffic is None and os.path.exists(tpath) and world == 1:
            tb = json.load(open(tpath)).get("bytes_per_launch", {})
            for frag, key in keys.items():
                if frag in dom["name"]:
                    for name, val in tb.items():
                        if name.startswith(key) and traffic is None:
                            traffic = {"bytes_per_launch": val, "source": "profiles/%s (ncu --set full)" % tname}
    roofline = dict(bound=dom["bound"], achieved=dom["achieved"], peak=dom["peak"], unit=dom["unit"], frac=dom["frac"], traffic=traffic,
                    kernel=dom["name"], kernel_ms=dom["ms"], share_of_step=dom["ms"] / ms_per_step,
                    peak_source=("fp32 FMA-chain microbenchmark run by this bench (MEASURED_PEAKS.json has no fp32 figure)"
                                 if dom["bound"] == "fp32" else (("half of %s bf16_tflops (tf32 dense rate); algorithmic flops, "
                                                                  "the kernels issue 2 - 3 MMA passes per product" % peak_src) if dom["bound"] == "tensor" else peak_src)))

    # ---- the same steps back to back (no L2 flush, no host wait in between): what a sampling loop sustains, launch gaps included
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb = min(args.steps, 200) if not job.dopri5 else min(args.steps, 5)
    a.record()
    for _ in range(nb):
        run()
    b.record()
    barrier()
    b2b_ms = a.elapsed_time(b) / nb

    # ---- end to end through the public API with HOST buffers.  Every step copies ITS observations from pinned host memory (H2D),
    # runs, and copies its per-particle loss back (D2H), written the way a sampling loop with a data loader is written so that the
    # GPU never waits for a copy engine or for the host:
    #   * the observations of step i+1 travel on a copy stream into the second of two device buffers while step i runs (the
    #     posterior is pointed at buffer i % 2: two captured graphs where the step replays as a CUDA graph);
    #   * the loss of step i is parked in a device staging buffer by the step and copied to one of two pinned host buffers on the
    #     copy stream; the host READS it one step later (loss_ready event), the last one inside the timed region.
    cs = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    res_x0, res_Y = post.x0, post.Y
    stage_in = [(torch.empty_like(post.x0), torch.empty_like(post.Y)) for _ in range(2)]
    stage_loss = [torch.empty_like(post.loss) for _ in range(2)]
    loss_host = [torch.empty(P_gpu, dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    step_done = [torch.cuda.Event() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def e2e_body(k):
        post.x0, post.Y = stage_in[k]
        job.step()
        stage_loss[k].copy_(post.loss)

    e2e_graphs = None
    if use_graph:
        try:
            gs = []
            for k in range(2):
                for t_, h_ in zip(stage_in[k], (job.x0_host, job.obs_host)):
                    t_.copy_(h_)
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2):
                    e2e_body(k)
                gs.append(g2)
            e2e_graphs = gs
        except Exception as e:
            print("e2e graph capture failed (%s); the e2e steps launch eagerly" % str(e).splitlines()[0], file=sys.stderr)
            torch.cuda.synchronize()
    if world > 1:
        flag = torch.tensor([int(e2e_graphs is not None)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            e2e_graphs = None
    host_sum = [0.0]

    def issue_h2d(i):                                      # observations of step i -> device buffer i % 2, on the copy stream
        k = i & 1
        cs.wait_event(step_done[k])                        # step i-2 has finished reading that buffer
        with torch.cuda.stream(cs):
            stage_in[k][0].copy_(job.x0_host, non_blocking=True)
            stage_in[k][1].copy_(job.obs_host, non_blocking=True)
            h2d_done[k].record(cs)

    def consume(i):                                        # the host-side read of step i's result
        loss_ready[i & 1].synchronize()
        host_sum[0] += float(loss_host[i & 1][0])

    def e2e_step(i, last):
        k = i & 1
        if not last:
            issue_h2d(i + 1)
        main.wait_event(h2d_done[k])
        if e2e_graphs is not None:
            e2e_graphs[k].replay()
        else:
            e2e_body(k)
        step_done[k].record(main)
        cs.wait_event(step_done[k])
        with torch.cuda.stream(cs):
            loss_host[k].copy_(stage_loss[k], non_blocking=True)
            loss_ready[k].record(cs)
        if i > 0:
            consume(i - 1)

    def e2e_run(n):
        for k in range(2):
            step_done[k].record(main)
        issue_h2d(0)
        for i in range(n):
            e2e_step(i, i == n - 1)
        consume(n - 1)                                     # the last result has reached the host
        main.wait_stream(cs)

    e2e_run(3)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_units = 0.0
    a.record()
    if job.dopri5:
        for k in range(2):
            step_done[k].record(main)
        issue_h2d(0)
        for i in range(args.steps):
            e2e_step(i, i == args.steps - 1)
            e2e_units += job.units_per_step()
        consume(args.steps - 1)
        main.wait_stream(cs)
    else:
        e2e_run(args.steps)
    b.record()
    barrier()
    post.x0, post.Y = res_x0, res_Y
    if not job.dopri5:
        e2e_units = job.units_per_step() * args.steps
    e2e_ms = a.elapsed_time(b)
    tt = torch.tensor([e2e_ms, e2e_units], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        e2e_ms, e2e_units = float(mx[0].item()), float(tt[1].item())
    e2e = dict(value=e2e_units / (e2e_ms * 1e-3), unit="particle*RK-steps/s",
               h2d_bytes_per_step=int(job.x0_host.numel() * 4 + job.obs_host.numel() * 4), d2h_bytes_per_step=int(P_gpu * 4),
               ms_per_step=e2e_ms / args.steps, cuda_graph=e2e_graphs is not None,
               pipeline="observations of step i+1 copied (H2D, copy stream, second device buffer) while step i runs; loss of step i "
                        "copied D2H on the copy stream and read on the host one step later; first H2D and last read inside the timed region; "
                        "no L2 flush between e2e steps (value / ms_per_step are the flushed per-step figures)")
    job.final_checks()

    # ---- N > 1: correctness of the sharded interaction, and BASELINE config 3 as written (4096 particles in total)
    parity = None
    strong = None
    if world > 1 and sampler == "svgd":
        if not args.no_parity:
            parity = svgd_parity(job, torch, dist, world, rank)
            job.final_checks()
        if args.scaling == "weak" and not args.no_strong:
            P_s = max(1, (args.particles or wl["P"]) // world)
            del smp, post, field, run, tr
            job = None
            torch.cuda.empty_cache()
            sjob = Job(args, wl, P_s, rank, world, torch, bode)
            sargs = argparse.Namespace(**vars(args))
            sargs.steps = min(args.steps, 300)
            st = timed_region(sjob, sargs, torch, dist, world, rank, local, flush, want_clocks=False)
            strong = dict(total_particles=P_s * world, particles_per_gpu=P_s, ms_per_step=st["ms_per_step"], value=st["value"],
                          unit="particle*RK-steps/s", steps=sargs.steps, warmup=st["n_pre"],
                          note="BASELINE config 3 as written: the same 4096 particles split over the ranks (strong scaling)")
            if not args.no_parity:
                sp = svgd_parity(sjob, torch, dist, world, rank)
                if sp is not None:
                    strong["parity"] = sp
            sjob.final_checks()
    elif world > 1:
        parity = dict(ok=True, note="independent chains: no cross-rank quantity exists (each rank's chains equal a single-GPU run of "
                                    "the same seeds; tests/test_multigpu.py, tests/test_dist_cpu.py)")

    if rank != 0:
        _finish(world, dist)
        return
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args, wl, P_total, args.cpu_steps)
    cfg.update({"grad": "frozen-step discrete adjoint of the accepted dopri5 steps (== odeint_adjoint to solver tolerance)"
                if wl["method"] == "dopri5" else "discrete adjoint (== autograd through odeint)",
                "parallelism": "particles sharded, dp%d" % world, "cuda_graph": bool(use_graph),
                "l2": "256 MiB memset between timed steps (outside the event brackets)"})
    if sampler == "svgd":
        cfg["streams"] = "SVGD operands + Gram + exact median on a side stream beside the fused solve"
        cfg["exchange"] = "none (one GPU)" if world == 1 else (
            ("positions + scores: push kernels over NVLink peer memory (flag barriers, no collective)" if args.gather == "p2p" else
             "positions + scores: NCCL all-gather") + "; exact median via peer reads")
    else:
        cfg["exchange"] = "none (independent chains)"
    launches = {"svgd": 7 + (2 if world > 1 else 0), "psgld": 3, "sgld": 3, "asghmc": 2, "hamcmc": 2}[sampler]
    if wl["field"] == "npde" and wl["M"] ** 2 >= 64:
        launches += 2                                        # proj_W_kernel, proj_back_kernel around the solve
    line = {
        "metric": "particle*RK-steps/sec (fwd+grad)", "value": value, "unit": "particle*RK-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": n_pre_main, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "roofline": roofline, "kernels": kernels, "peaks": peaks,
        "cpu_baseline": cb, "e2e": e2e, "gpu_launches": launches * args.steps, "clocks": clk_main,
        "wall_s_timed_region": t_wall_main, "back_to_back_ms_per_step": b2b_ms, "check": "ok", "finite_chain_fraction": finite_frac,
    }
    if median_fallback_ms is not None:
        line["median_fallback_ms"] = median_fallback_ms
        line["median_fallback_note"] = ("one step with the median window disarmed (3-pass radix select over the stored d2 instead of the "
                                        "window table); the timed steps all hit the window")
    if parity is not None:
        line["parity"] = parity
    if strong is not None:
        line["strong"] = strong
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    _finish(world, dist)


def _finish(world, dist):
    """Multi-rank teardown: final barrier, then destroy the process group; a rank whose teardown blocks (a captured graph that
    contains NCCL kernels was seen to keep the communicator busy) leaves after 20 s without it."""
    if world > 1:
        import threading
        import torch
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        t = threading.Timer(20.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        try:
            dist.destroy_process_group()
        finally:
            t.cancel()


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
