"""Time one SVGD sampler step (c3 shape) captured as a CUDA graph with the side-stream forks switched on and off."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200 import problems, _lib
from bayesian_ode_b200.samplers import SVGD

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
data = problems.make_dataset("VDP", seed=0, N=5, R=3.0, T=40, t_end=7.0, noise=0.1)
Z = problems.inducing_grid(data["Y"], 5)
U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
U = U0[None] + 0.1 * torch.randn(P, 25, 2, generator=torch.Generator().manual_seed(1234), dtype=torch.float64)
flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")


def timed(fn, iters=300):
    for _ in range(20):
        flush_buf.zero_(); fn()
    torch.cuda.synchronize()
    tot = 0.0
    evs = []
    for _ in range(iters):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3


def build(prefetch, overlap, side_sms=40):
    f = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]), method="rk4", grad_mode="discrete")
    f.bind_flat_grads()
    smp = SVGD([f.U, f.logsn], lr=1e-4, overlap=overlap, side_sms=side_sms)

    def step():
        if prefetch:
            smp.prefetch()
        post.loss_and_grad_()
        smp.phi(update_lr=1e-4)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    return g.replay, (f, post, smp)


lib = _lib.load()
f0 = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
post0 = bode.NPDEPosterior(f0, data["x0"], data["t"], torch.from_numpy(data["Y"]), method="rk4", grad_mode="discrete")
for lim in (0, 128, 118, 108, 100):
    lib.bode_npde_set_cta_limit(lim)
    post0.loss_and_grad_()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        post0.loss_and_grad_()
    print("ODE alone, cta limit %3d: %.1f us" % (lim, timed(g.replay)))
lib.bode_npde_set_cta_limit(0)
for js in (0, 8, 16):
    lib.bode_svgd_set_gram_split(js)
    for name, pf, ov, ss in (("serial", False, False, 0), ("gram side_sms=32", True, "gram", 32), ("gram side_sms=40", True, "gram", 40),
                             ("gram side_sms=52", True, "gram", 52)):
        fn, keep = build(pf, ov, ss)
        print("gram split %2d  %-22s %.1f us/step" % (js, name, timed(fn)))
