"""Developer probe: accuracy of the SVGD interaction (bode_svgd_* through SVGD.phi) against a float64 evaluation of stein.py:75-86
as the particle count grows, split by term (driving term K S / n, repulsion 2 gamma (rowsum x_i - K X) / n)."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200 import problems
from bayesian_ode_b200.samplers import SVGD

data = problems.make_dataset("VDP", seed=0, N=5, R=3.0, T=40, t_end=7.0, noise=0.1)
Z = problems.inducing_grid(data["Y"], 5)
U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
for P in [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384, 32768]:
    gen = torch.Generator().manual_seed(1)
    U = U0[None] + 0.1 * torch.randn(P, 25, 2, generator=gen, dtype=torch.float64)
    f = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
    f.bind_flat_grads()
    smp = SVGD([f.U, f.logsn], lr=1e-4)
    post.loss_and_grad_()
    G0 = f.theta_grad.detach().clone()
    X = f.theta.detach().double()
    sq = (X * X).sum(1)
    rows = torch.linspace(0, P - 1, 32, device=X.device).long()
    for name, scale in (("full", 1.0), ("repulsion only (scores = 0)", 0.0)):
        with torch.no_grad():
            f.theta_grad.copy_(G0 * scale)
        phi = smp.phi().clone().double()
        med = float(smp._ws.med_gamma[0]); gamma = float(smp._ws.med_gamma[1])
        Xr = X[rows]
        d2 = (sq[rows, None] + sq[None] - 2.0 * Xr @ X.T).clamp_min_(0)
        K = torch.exp(-gamma * d2)                      # the kernel's own gamma: isolates the contraction error
        S = -(G0 * scale).double()
        drive = K @ S / P
        rep = 2.0 * gamma * (K.sum(1, keepdim=True) * Xr - K @ X) / P
        ref = drive + rep
        err = float((phi[rows] - ref).abs().max() / ref.abs().max())
        print("P=%6d %-28s max rel err %.2e   |drive| %.2e |rep| %.2e  rowsum(K) mean %.1f  gamma %.3e" % (
            P, name, err, float(drive.abs().max()), float(rep.abs().max()), float(K.sum(1).mean()), gamma), flush=True)
    del smp, post, f
    torch.cuda.empty_cache()
