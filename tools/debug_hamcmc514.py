"""Debug aid: HAMCMC at d = 514 on 2048 chains -- which chains go non-finite, when, and does the float64 oracle agree?"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bayesian_ode_b200 as bode
from bayesian_ode_b200.samplers import HAMCMC
from oracle import npde, samplers as osamp
g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "npde_m5.npz")))
lr0 = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-7
scaleU = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
M, ell, P, memory = 16, 0.35, 2048, 5
Z = npde.inducing_grid(g["Y"], M)
rng = np.random.default_rng(33)
U = scaleU * rng.standard_normal((P, M * M, 2))
f = bode.NPDEField(torch.from_numpy(U), torch.from_numpy(Z), 1.0, ell, 0.1, stable_solve=True)
post = bode.NPDEPosterior(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["Y"]))
f.bind_flat_grads()
smp = HAMCMC([f.U, f.logsn], memory=memory, lr0=lr0, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3, H_gamma=1.0, trust_reg=1.0)
smp.check_finite = "deferred"
Mm = memory + 1
hist = []
th0 = f.theta.detach().cpu().numpy().copy()
for it in range(2 * Mm - 1 + 9):
    loss, gU, gl = post.loss_and_grad_()
    grad = f.theta_grad.detach().cpu().numpy().copy()
    xi = rng.standard_normal((P, 514))
    lr = smp.get_lr(it)
    metric = it >= 2 * Mm - 1
    if metric: smp.step(lr=lr, noise=xi)
    else: smp.step_without_metric(lr=lr, add_params=True, noise=xi)
    th = f.theta.detach().cpu().numpy().copy()
    bad = ~np.isfinite(th).all(1)
    print("it %2d metric=%d lr=%.3e |g|max=%.3e loss[0]=%.4e nonfinite chains=%d pairs(min/max)=%s" % (
        it, metric, lr, np.nanmax(np.abs(grad)), float(loss[0]), int(bad.sum()),
        (int(smp.n_pairs().min()), int(smp.n_pairs().max()))), flush=True)
    hist.append((grad, xi, lr, metric, th, bad))
firstbad = None
for it, h in enumerate(hist):
    if h[5].any():
        firstbad = (it, int(np.nonzero(h[5])[0][0])); break
print("first bad:", firstbad)
chains = [0] + ([firstbad[1]] if firstbad else [])
for c in chains:
    orc = osamp.HAMCMC(memory=memory, H_gamma=1.0, trust_reg=1.0)
    th_o = th0[c].astype(np.float64)
    for it, (grad, xi, lr, metric, th, bad) in enumerate(hist):
        gk = grad[c].astype(np.float64)
        if metric:
            sy = [float(s @ y) for s, y in zip(orc.s, orc.y)]
            ss = [float(s @ s) for s in orc.s]
            with np.errstate(all="ignore"):
                th_o = orc.step(gk, lr, xi[c])
        else:
            sy, ss = [], []
            th_o = orc.step_without_metric(th_o, gk, lr, xi[c], add_params=True)
        err = np.abs(th[c] - th_o).max() / np.abs(th_o).max()
        print("chain %d it %2d err=%.2e finite(gpu,oracle)=(%d,%d) pairs=%d sy=%s ss=%s" % (
            c, it, err, np.isfinite(th[c]).all(), np.isfinite(th_o).all(), len(orc.s), ["%.2e" % v for v in sy], ["%.2e" % v for v in ss]))
