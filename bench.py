#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native bayesian-ode hot path.

Metric (BASELINE.json): particle*RK-steps/sec, forward + gradient, sampler update and SVGD interaction inside the timed
region.  One particle*RK-step = all N trajectories of one particle advanced by one RK step and differentiated.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1] [--scaling weak|strong] [--impl b200|reference]

Workloads:
  c3 (default)  Van der Pol npde, 5x5 inducing grid, SVGD, 4096 particles per GPU, N=5, T=40 -> 39 rk4 (3/8) steps
  c2            Van der Pol npde, pSGLD, 1024 independent chains per GPU, T=101 -> 100 rk4 steps
  c1            1 chain SGLD (the reference's own CPU-runnable case; parity-sized)
One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); the only data-path exchange is the SVGD all-gather of positions and scores
(push kernels over NVLink peer memory; NCCL when the workspaces cannot be peer-mapped).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c3": dict(sampler="svgd", P=4096, M=5, T=40, desc="VDP npde SVGD, 5x5 grid, 4096 particles/GPU, N=5, T=40 (39 rk4 3/8 steps)"),
    "c2": dict(sampler="psgld", P=1024, M=5, T=101, desc="VDP npde pSGLD, 5x5 grid, 1024 chains/GPU, N=5, T=101 (100 rk4 3/8 steps)"),
    "c1": dict(sampler="sgld", P=1, M=5, T=40, desc="VDP npde SGLD, 5x5 grid, 1 chain, N=5, T=40 (39 rk4 3/8 steps)"),
}
STAGES = 4
FLOP_PER_EVAL = 23.0          # per (RK stage, trajectory, inducing point), fwd 11 + adjoint 12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--particles", type=int, default=None, help="override particles per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=12)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_baseline(wl, cpu_steps, P_total):
    """The reference's CPU path (torch float64 port of its execution model, oracle/ref_torch.py) on this box's host
    cores: one chain per process, one thread each, C = os.cpu_count() processes."""
    import multiprocessing as mp
    from oracle import ref_torch
    C = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(C) as pool:
        per_step = pool.map(ref_torch.time_chain, [(i, wl["T"], wl["M"], 2, cpu_steps) for i in range(C)])
    rk_steps = wl["T"] - 1
    chain_rate = sum(rk_steps / s for s in per_step)            # particle*RK-steps/s, all cores busy
    sample = "%d processes x (2 warm-up + %d timed) SGLD steps of one chain each, fp64, 1 thread per process" % (C, cpu_steps)
    value = chain_rate
    if wl["sampler"] == "svgd":
        # the reference's SVGD.step is a stub: per-particle solve+grad as above plus a numpy/BLAS phi over all particles
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
    # themselves to one thread each, the numpy/BLAS phi runs on all of them).  Must happen before numpy / torch are imported.
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    P_total = (args.particles or wl["P"]) * (args.gpus if args.scaling == "weak" else 1)
    t0 = time.perf_counter()
    steps = max(args.steps, 1)
    cb = cpu_baseline(wl, min(steps, 20), P_total)
    line = {
        "impl": "reference", "metric": "particle*RK-steps/sec (fwd+grad)", "value": cb["value"], "unit": cb["unit"],
        "n_gpus": args.gpus, "steps": min(steps, 20), "warmup": 2, "ms_per_step": cb["per_process_ms_per_step"],
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "total_particles": P_total},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ helpers (GPU arm)
def checked(smp, notes):
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


def graph_seg_ms(torch, fn, reps=20, iters=7, flush=None):
    """Device time of one call of `fn` (a short sequence of kernel launches): `reps` back-to-back calls are captured in ONE
    CUDA graph and the replay is timed with events, so neither Python/ctypes launch overhead nor the graph-launch latency
    (~8 us per replay, measured with an empty graph) is attributed to the kernels.  `flush` runs before each replay."""
    fn()
    torch.cuda.synchronize()
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


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    from bayesian_ode_b200.samplers import SGLD, SVGD, pSGLD

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

    P_gpu = args.particles or wl["P"]
    if args.scaling == "strong":
        P_gpu = max(1, P_gpu // world)
    P_total = P_gpu * world
    M, T = wl["M"], wl["T"]
    S = T - 1
    data = problems.make_dataset("VDP", seed=0, N=5, R=3.0, T=T, t_end=7.0, noise=0.1)
    N = data["N"]
    Z = problems.inducing_grid(data["Y"], M)
    U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
    gen = torch.Generator().manual_seed(1234 + rank)
    U = U0[None] + 0.1 * torch.randn(P_gpu, M * M, 2, generator=gen, dtype=torch.float64)       # gp.py:321 init scale
    field = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
    x0_host = data["x0"].float().pin_memory()
    Y_host = torch.from_numpy(data["Y"]).float().pin_memory()
    post = bode.NPDEPosterior(field, data["x0"], data["t"], torch.from_numpy(data["Y"]), method="rk4", grad_mode="discrete")
    field.bind_flat_grads()
    params = [field.U, field.logsn]
    if wl["sampler"] == "svgd":
        smp = SVGD(params, lr=1e-4)
        sched = None
    elif wl["sampler"] == "psgld":
        smp = pSGLD(params, lr0=5e-3, lr_gamma=0.51, lr_t0=100, lr_alpha=0.1, lambda_=1e-8, alpha=0.99, N=N, seed=7 + rank)
        post.scale = 1.0 / N
        sched = dict(kind=1, lr0=5e-3, gamma=0.51, t0=100, alpha=0.1)
    else:
        smp = SGLD(params, lr0=1e-4, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03, seed=7 + rank)
        sched = dict(kind=1, lr0=1e-4, gamma=0.51, t0=100, alpha=0.03)
    smp.check_finite = "deferred"

    def step():
        if wl["sampler"] == "svgd":
            smp.prefetch()                  # position-only half of the interaction (operands, Gram, median) beside the ODE kernel
        post.loss_and_grad_()
        if wl["sampler"] == "svgd":
            smp.phi(update_lr=smp.param_groups[0]["lr"])
        else:
            smp.schedule_(**sched)
            smp.step(use_ctl=True)

    launches_per_step = 7 if wl["sampler"] == "svgd" else 3
    if wl["sampler"] == "svgd" and world > 1 and smp.gather_comm == "p2p":
        launches_per_step += 2              # the two push-kernel all-gathers (positions, scores); with "nccl" they are NCCL's kernels
    peaks = measure_peaks(torch, bode)

    # eager warm-up (allocates every buffer), then capture one step in a CUDA graph (single-GPU path)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    use_graph = not args.no_graph
    run = step
    if use_graph:
        try:
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                step()
            run = gph.replay
        except Exception as e:                                   # e.g. NCCL capture unsupported: fall back to eager launches
            print("cuda graph capture failed (%s); running eagerly" % str(e).splitlines()[0], file=sys.stderr)
            use_graph = False
            run = step
            torch.cuda.synchronize()
    if world > 1:
        flag = torch.tensor([int(use_graph)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and use_graph:
            use_graph, run = False, step

    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")              # 256 MiB > 126 MB L2

    def flush():
        flush_buf.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t_pre = time.perf_counter()
    n_pre = 0
    while n_pre < max(args.warmup, 3) or (rank == 0 and world == 1 and time.perf_counter() - t_pre < 0.6):
        flush()                                       # warm-up; on one GPU it also runs until nvidia-smi delivers samples under this load
        run()
        n_pre += 1
    barrier()
    t_wall = time.perf_counter()
    step_ms = event_ms(torch, run, args.steps, flush=flush)
    barrier()
    t_wall = time.perf_counter() - t_wall
    clk = clocks.stop() if rank == 0 else None
    if clk is not None:
        clk["window"] = "sampled every 100 ms from %d warm-up steps of the same graph through the %d timed steps" % (n_pre, args.steps)
    total_ms = sum(step_ms)
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    check_notes = []
    checked(smp, check_notes)
    ms_per_step = total_ms / args.steps
    value = P_total * S * args.steps / (total_ms * 1e-3)

    # ---- per-kernel timing pass for the roofline objects: each C-ABI call sequence is captured 20x in its own CUDA graph
    # (graph_seg_ms), so the figures are device time without launch gaps; inputs stay L2-resident between repetitions.
    lib = bode._lib.load()
    ode_all_ms = graph_seg_ms(torch, lambda: post.loss_and_grad_())              # one CTA on every SM
    ode_ms, ode_cfg = ode_all_ms, "one CTA per SM"
    if wl["sampler"] == "svgd" and smp.overlap == "gram" and smp.side_sms > 0:
        # the configuration the timed step launches: packed into fewer CTAs so that the Gram pass can run beside it
        old = lib.bode_npde_set_cta_limit(lib.bode_device_sm_count() - smp.side_sms)
        ode_ms = graph_seg_ms(torch, lambda: post.loss_and_grad_())
        lib.bode_npde_set_cta_limit(old)
        ode_cfg = "as launched in the step: at most %d CTAs, timed alone" % (lib.bode_device_sm_count() - smp.side_sms)
    ode_flop = FLOP_PER_EVAL * STAGES * N * M * M * S * P_gpu
    kernels = [dict(name="npde_pair_grad_kernel (fused rk4 solve + closure + discrete adjoint, 2 lanes per pair, FFMA2; %s)" % ode_cfg, ms=ode_ms,
                    bound="fp32", achieved=ode_flop / (ode_ms * 1e-3) / 1e12, peak=peaks["fp32_fma_tflops"], unit="TFLOP/s",
                    ms_one_cta_per_sm=ode_all_ms)]
    if wl["sampler"] == "svgd" and world == 1:
        import ctypes as C
        d = field.d
        ws = smp._ws
        X = field.theta
        nl, nt = P_gpu, P_total
        Xall, Gall = X, field.theta_grad
        xr, xrs = bode._lib.rows(X, d); xc, xcs = bode._lib.rows(Xall, d); sc, scs = bode._lib.rows(Gall, d)
        sq_fn = lambda: ws.sqdist(X, nl, Xall, nt, d, nt * nt, row_offset=rank * nl)
        def sqmed_fn():
            sq_fn()
            ws.median(nl, nt, d, nt, None, group=None)
        phi_fn = lambda: bode._lib.check(lib.bode_svgd_phi(xr, xrs, nl, xc, xcs, sc, scs, -1.0, nt, d, nt, bode._lib.ptr(ws.med_gamma),
                                                           C.c_void_p(ws.base.data_ptr()), bode._lib.ptr(smp.phi_buf), d, None, 0, 0.0,
                                                           bode._lib.stream_ptr()))
        sqmed_ms = graph_seg_ms(torch, sqmed_fn)
        sq_ms = graph_seg_ms(torch, sq_fn)
        sqmed_fn()                                   # leave the selection state consistent (the repeated sqdist filled the window table)
        med_ms = max(sqmed_ms - sq_ms, 1e-4)
        phi_ms = graph_seg_ms(torch, phi_fn)
        tc = bool(lib.bode_svgd_set_tensor_cores(1)) and d <= 55
        lib.bode_svgd_set_tensor_cores(int(tc))
        tpeak = None
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
            tpeak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops")
        tpeak = (tpeak or 1590.0) / 2.0          # tf32 dense = half the bf16 rate; no measured tf32 figure exists
        kernels.append(dict(name="svgd sqdist: prep_x + gram2 (3xTF32 tcgen05, d2 store + median window count)" if tc else "svgd sqdist_kernel",
                            ms=sq_ms, bound="tensor" if tc else "fp32", achieved=2.0 * nl * nt * d / (sq_ms * 1e-3) / 1e12,
                            peak=tpeak if tc else peaks["fp32_fma_tflops"], unit="TFLOP/s"))
        kernels.append(dict(name="svgd exact median: window_select + radix fallback (no-op after a window hit) + gamma", ms=med_ms, bound="hbm",
                            achieved=(2 * 16384 + 2) * 8 * 2 / (med_ms * 1e-3) / 1e9, peak=None, unit="GB/s"))
        kernels.append(dict(name="svgd phi: prep_v + phi2 K@[S|X|1] (3xTF32 tcgen05, TMA d2 tiles) + combine" if tc else "svgd phi_partial+combine (K@[S|X])",
                            ms=phi_ms, bound="tensor" if tc else "fp32", achieved=4.0 * nl * nt * d / (phi_ms * 1e-3) / 1e12,
                            peak=tpeak if tc else peaks["fp32_fma_tflops"], unit="TFLOP/s"))
    hbm_peak = None
    mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_src = "fallback (B200_PROFILING.md)"
    if os.path.exists(mp_path):
        hbm_peak = json.load(open(mp_path)).get("hbm_gbs")
        peak_src = "MEASURED_PEAKS.json"
    hbm_peak = hbm_peak or 6650.0
    for k in kernels:
        if k["bound"] == "hbm":
            k["peak"] = hbm_peak
        k["frac"] = k["achieved"] / k["peak"]
    dom = max(kernels, key=lambda k: k["ms"])
    # DRAM traffic of the dominant kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of the committed
    # `ncu --set full` capture (profiles/traffic_r01_final.json, written by tools/ncu_summary.py); not measurable from here
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01_final.json")
    if os.path.exists(tpath):
        tb = json.load(open(tpath)).get("bytes_per_launch", {})
        key = "npde_pair_grad_kernel" if dom["bound"] == "fp32" else ("gram2_kernel" if "sqdist" in dom["name"] else "phi2_kernel")
        for name, val in tb.items():
            if name.startswith(key):
                traffic = {"bytes_per_launch": val, "source": "profiles/traffic_r01_final.json (ncu --set full, P=4096 c3 launch)"}
    roofline = dict(bound=dom["bound"], achieved=dom["achieved"], peak=dom["peak"], unit=dom["unit"], frac=dom["frac"], traffic=traffic,
                    kernel=dom["name"], kernel_ms=dom["ms"],
                    peak_source=("fp32 FMA-chain microbenchmark run by this bench (MEASURED_PEAKS.json has no fp32 figure)"
                                 if dom["bound"] == "fp32" else ("half of MEASURED_PEAKS.json bf16_tflops (tf32 dense rate); algorithmic flops, "
                                                                 "the kernel issues 3 MMAs per product" if dom["bound"] == "tensor" else peak_src)))

    # ---- end to end through the public API with HOST buffers: H2D of this step's observations, step, D2H of the loss
    loss_host = torch.empty(P_gpu, dtype=torch.float32).pin_memory()
    def e2e_step():
        post.set_data(x0=x0_host, Y=Y_host)
        run()
        loss_host.copy_(post.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(3):
        e2e_step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        e2e_step()
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    if world > 1:
        tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e = dict(value=P_total * S * args.steps / (e2e_ms * 1e-3), unit="particle*RK-steps/s",
               h2d_bytes_per_step=int(x0_host.numel() * 4 + Y_host.numel() * 4), d2h_bytes_per_step=int(P_gpu * 4),
               ms_per_step=e2e_ms / args.steps)
    checked(smp, check_notes)
    assert bool(torch.isfinite(loss_host).all()), "non-finite loss at the end of the run"

    if rank != 0:
        _finish(world, dist)
        return
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(wl, args.cpu_steps, P_total)
    line = {
        "metric": "particle*RK-steps/sec (fwd+grad)", "value": value, "unit": "particle*RK-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": n_pre, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "particles_per_gpu": P_gpu, "total_particles": P_total, "rk_steps": S, "trajectories": N,
                   "inducing_grid": "%dx%d" % (M, M), "sampler": wl["sampler"], "grad": "discrete adjoint (== autograd through odeint)",
                   "parallelism": "particles sharded, dp%d" % world, "cuda_graph": bool(use_graph),
                   "streams": ("SVGD operands + Gram + exact median on a side stream beside the fused solve (solve packed into %d CTAs, %d SMs left "
                               "to the Gram pass)" % (148 - smp.side_sms, smp.side_sms)) if wl["sampler"] == "svgd" and smp.overlap == "gram" else "single",
                   "exchange": ("none (one GPU)" if world == 1 else
                                "positions + scores gathered by %s; exact median via %s" % (
                                    "push kernels over NVLink peer memory (flag barriers, no collective)" if smp.gather_comm == "p2p" else "NCCL all-gather",
                                    "peer reads + flag barriers" if smp.median_comm == "p2p" else "NCCL all-reduce")) if wl["sampler"] == "svgd" else "none (independent chains)",
                   "l2": "256 MiB memset between timed steps (outside the event brackets); working set < L2"},
        "roofline": roofline, "kernels": kernels, "peaks": peaks,
        "cpu_baseline": cb, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clk,
        "wall_s_timed_region": round(t_wall, 3), "check": sorted(set(check_notes)) or "ok",
    }
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    _finish(world, dist)


def _finish(world, dist):
    """Multi-rank teardown.  A captured graph that contains NCCL kernels keeps the communicator busy and
    destroy_process_group() was seen to block on it, so after a final barrier the ranks leave without it."""
    if world > 1:
        import torch
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
