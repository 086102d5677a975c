import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bayesian_ode_b200 as bode
from bayesian_ode_b200 import problems
for M, ell in ((16, 0.35),):
    for T in (2, 11, 40):
        data = problems.make_dataset("VDP", seed=0, N=5, R=3.0, T=T, t_end=7.0 * (T - 1) / 39.0, noise=0.1)
        Z = problems.inducing_grid(data["Y"], M)
        P = 2048 if M == 16 else 4096
        U = 0.1 * torch.randn(P, M * M, 2, dtype=torch.float64)
        f = bode.NPDEField(U, Z, 1.0, ell, 0.1, **({"stable_solve": True} if M > 6 else {}))
        post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
        for _ in range(3): post.loss_and_grad_()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record()
        for i in range(10):
            post.loss_and_grad_(); ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
        print("M=%d P=%d steps=%d  %.4f ms" % (M, P, T - 1, ts[5]), flush=True)
