"""Smallest case that touches every new kernel once (for compute-sanitizer --tool memcheck on the GPU box)."""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200 import problems
from bayesian_ode_b200.samplers import SVGD, MALA
data = problems.make_dataset(seed=0)
Z = problems.inducing_grid(data["Y"], 5)
U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
for P in (200, 384):                      # ragged (200 % 128 != 0) and full tiles
    U = U0[None] + 0.1 * torch.randn(P, 25, 2, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    f = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
    f.bind_flat_grads()
    smp = SVGD([f.U, f.logsn], lr=1e-4)
    for _ in range(3):                    # first call: radix fallback; later calls: window hit
        post.loss_and_grad_()
        smp.phi(update_lr=1e-4)
    torch.cuda.synchronize()
    print("P=%d ok, median/gamma" % P, smp._ws.med_gamma.tolist())
m = MALA([f.U, f.logsn], lr=1e-6, exact=True)
m.sample(post, num_samples=2, burn_in=1)
torch.cuda.synchronize()
print("mala ok")
