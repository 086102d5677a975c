"""c5 closure (16 x 16 general-Z grid, 2048 particles, 39 rk4 steps): a few calls, for an ncu launch list."""
import sys, torch
sys.path.insert(0, "/root/repo")
import bayesian_ode_b200 as bode
from bayesian_ode_b200 import problems
M, ell, T = 16, 0.35, int(sys.argv[1]) if len(sys.argv) > 1 else 40
data = problems.make_dataset("VDP", seed=0, N=5, R=3.0, T=T, t_end=7.0 * (T - 1) / 39.0, noise=0.1)
Z = problems.inducing_grid(data["Y"], M)
U = 0.1 * torch.randn(2048, M * M, 2, dtype=torch.float64)
f = bode.NPDEField(U, Z, 1.0, ell, 0.1, stable_solve=True)
post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
for _ in range(4):
    post.loss_and_grad_()
torch.cuda.synchronize()
print("ok")
