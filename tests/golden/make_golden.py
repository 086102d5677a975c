#!/usr/bin/env python
"""Generate golden fixtures from the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py
Writes small .npz files next to this script.  Every array is produced by the
reference's own code (torchdiffeq/, samplers/, scripts/vanderpol/gp.py); the only
restated piece is ``loss_closure`` (a nested function, gp.py:342-353, copied
semantics) and the seeded data generation of notebooks/jai/gp.ipynb cell 3.
"""
import os
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

warnings.filterwarnings("ignore")
for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.ticker", "seaborn"]:
    sys.modules[m] = MagicMock()
REF = os.environ.get("BODE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import torch  # noqa: E402
import scipy.stats as ss  # noqa: E402
import torchdiffeq  # noqa: E402
from scripts.vanderpol import gp  # noqa: E402

torch.set_default_dtype(torch.float64)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


# --------------------------------------------------------------------- VDP data
def make_data():
    np.random.seed(0)
    torch.manual_seed(0)
    N, R = 5, 3
    x0 = torch.from_numpy(2 * R * ss.uniform.rvs(size=[N, 2]) - R)
    t = torch.linspace(0., 7., 40, dtype=torch.float32)          # notebook default dtype
    with torch.no_grad():
        X = torchdiffeq.odeint(gp.VDP(), x0, t, method="rk4").numpy()
    X = np.transpose(X, [1, 0, 2])
    Y = X + ss.norm.rvs(size=X.shape) * 0.1
    return dict(N=N, R=R, noise=0.1, x0=x0, t=t, X=X, Y=Y)


def make_model(data, M, sf=1.0, ell=0.75):
    """gp.py:315-333 verbatim semantics."""
    Y, t = data["Y"], data["t"]
    D = 2
    xv = np.linspace(np.min([np.min(Y_[:, 0]) for Y_ in Y]), np.max([np.max(Y_[:, 0]) for Y_ in Y]), M)
    yv = np.linspace(np.min([np.min(Y_[:, 1]) for Y_ in Y]), np.max([np.max(Y_[:, 1]) for Y_ in Y]), M)
    xv, yv = np.meshgrid(xv, yv)
    Z = np.array([xv.T.flatten(), yv.T.flatten()]).T
    Zt = torch.from_numpy(Z)
    Yt = torch.from_numpy(Y)
    F_ = (Yt[:, 1:, :] - Yt[:, :-1, :]) / (t[1] - t[0])
    F_ = F_.contiguous().view(-1, D)
    Z_ = Yt[:, :-1, :].contiguous().view(-1, D)
    Kxz = gp.K(Zt, Z_, sf, ell)
    Kzzinv = (gp.K(Z_, Z_, sf, ell) + 0.2 * torch.eye(Z_.shape[0])).inverse()
    U0 = torch.mm(torch.mm(Kxz, Kzzinv), F_)
    Linv = torch.linalg.cholesky(gp.K(Zt, Zt, sf, ell)).inverse()
    U0 = torch.mm(Linv, U0)
    return Zt, Yt, U0


def closure_for(kreg, odeint, x0, t, Yt, method, options=None, D=2):
    """gp.py:342-353."""
    Kzzinv = kreg.Kzzinv

    def loss_closure(add_prior=True):
        xode = odeint(kreg, x0, t, method=method, options=options).permute([1, 0, 2])
        if add_prior:
            loss = torch.sum((Yt - xode) ** 2 / (2 * torch.exp(kreg.logsn) ** 2))
            loss += torch.numel(Yt) * torch.sum(kreg.logsn) / D
            loss += torch.sum(torch.diag(torch.mm(kreg.U.t(), torch.mm(Kzzinv, kreg.U)))) / 2
        else:
            loss = torch.sum((Yt - xode) ** 2)
        return loss
    return loss_closure


def gen_npde(data, M, P, tag, methods=("euler", "midpoint", "rk4")):
    torch.cholesky = torch.linalg.cholesky                       # removed alias (gp.py:66)
    Zt, Yt, U0 = make_model(data, M)
    x0, t = data["x0"], data["t"]
    g = torch.Generator().manual_seed(1234 + M)
    Us = U0[None] + 0.1 * torch.randn(P, M * M, 2, generator=g)
    logsns = np.log(0.1) + 0.05 * torch.randn(P, 2, generator=g)
    out = dict(Z=Zt, Y=Yt, x0=x0, t=t, U0=U0, U=Us, logsn=logsns, sf=1.0, ell=0.75)
    for method in methods:
        sols, losses, sqerrs, gUd, gLd, gUa, gLa = [], [], [], [], [], [], []
        for p in range(P):
            kreg = gp.KernelRegression(Us[p].clone(), Zt, 1.0, 0.75, 0.1)
            kreg.logsn.data.copy_(logsns[p])
            if p == 0:
                out.update(Kzz=kreg.Kzz, Kzzinv=kreg.Kzzinv, KzzinvL=kreg.KzzinvL)
            # discrete adjoint: autograd through odeint
            cl = closure_for(kreg, torchdiffeq.odeint, x0, t, Yt, method)
            loss = cl()
            loss.backward()
            gUd.append(kreg.U.grad.clone()); gLd.append(kreg.logsn.grad.clone())
            losses.append(loss.detach())
            with torch.no_grad():
                sqerrs.append(cl(add_prior=False))
                sols.append(torchdiffeq.odeint(kreg, x0, t, method=method))
            # continuous adjoint: what gp.py:26 runs
            kreg.zero_grad()
            cl = closure_for(kreg, torchdiffeq.odeint_adjoint, x0, t, Yt, method)
            cl().backward()
            gUa.append(kreg.U.grad.clone()); gLa.append(kreg.logsn.grad.clone())
        out.update({
            f"{method}_sol": torch.stack(sols, 1),                  # [T,P,N,D]
            f"{method}_loss": torch.stack(losses), f"{method}_sqerr": torch.stack(sqerrs),
            f"{method}_gU_discrete": torch.stack(gUd), f"{method}_glogsn_discrete": torch.stack(gLd),
            f"{method}_gU_adjoint": torch.stack(gUa), f"{method}_glogsn_adjoint": torch.stack(gLa),
        })
    save(tag, **out)
    return out


def gen_grid_options(data):
    """step_size grids + the end-of-step 'interpolation' quirk + reversed time."""
    torch.cholesky = torch.linalg.cholesky
    Zt, Yt, U0 = make_model(data, 5)
    kreg = gp.KernelRegression(U0.clone(), Zt, 1.0, 0.75, 0.1)
    x0 = data["x0"]
    out = dict(U=U0, Z=Zt, x0=x0)
    cases = {
        "a": (torch.tensor([0.0, 0.25, 0.7, 1.0]), 0.5),
        "b": (torch.linspace(0., 2., 9), 0.125),
        "c": (torch.tensor([0.0, 0.3, 0.31, 1.7, 2.0]), 0.25),
    }
    for name, (t, h) in cases.items():
        for method in ("euler", "midpoint", "rk4"):
            with torch.no_grad():
                sol = torchdiffeq.odeint(kreg, x0, t, method=method, options={"step_size": h})
            out[f"{name}_{method}_sol"] = sol
        out[f"{name}_t"] = t
        out[f"{name}_h"] = h
        # gradient of sum(sol * w) wrt (U, x0) through the step-size grid
        w = torch.randn(len(t), *x0.shape, generator=torch.Generator().manual_seed(7))
        x0g = x0.clone().requires_grad_(True)
        kreg.zero_grad()
        sol = torchdiffeq.odeint(kreg, x0g, t, method="rk4", options={"step_size": h})
        (sol * w).sum().backward()
        out[f"{name}_w"] = w
        out[f"{name}_rk4_gU"] = kreg.U.grad.clone()
        out[f"{name}_rk4_gx0"] = x0g.grad.clone()
    # reversed time
    t = torch.linspace(2., 0., 9)
    with torch.no_grad():
        out["rev_t"] = t
        out["rev_rk4_sol"] = torchdiffeq.odeint(kreg, x0, t, method="rk4")
    save("grid_options", **out)


def _replay_noise(seed, shapes):
    """The reference draws Normal(0, std).sample() / torch.normal(0, std) per parameter in order; under the same
    seed these are std * torch.randn(shape) bit for bit -- verified below by reproducing the updates."""
    torch.manual_seed(seed)
    return [torch.randn(sh) for sh in shapes]


def gen_sampler_steps(data):
    """SGLD / pSGLD / aSGHMC: per-step (params, grads, noise) -> (params', state') sequences from the reference."""
    from samplers.langevin import SGLD, pSGLD
    from samplers.hamiltonian import aSGHMC
    from oracle import samplers as osamp
    torch.cholesky = torch.linalg.cholesky
    Zt, Yt, U0 = make_model(data, 5)
    x0, t = data["x0"], data["t"]
    out = {}

    gg = torch.Generator().manual_seed(4242)

    def fresh():
        kreg = gp.KernelRegression(U0.clone(), Zt, 1.0, 0.75, 0.1)
        return kreg, None

    def record(name, sampler, kreg, cl, nsteps, step_fn, shapes_fn):
        """Synthetic (seeded) gradients stand in for closure().backward(): the update rules under test do not care
        where the gradient came from, and real npde gradients make aSGHMC's momentum resampling diverge."""
        rec = {k: [] for k in ("U", "logsn", "gU", "glogsn", "U_new", "logsn_new", "lr")}
        noises = []
        for i in range(nsteps):
            kreg.U.grad = 50.0 * torch.randn(25, 2, generator=gg)
            kreg.logsn.grad = 5.0 * torch.randn(2, generator=gg)
            rec["U"].append(kreg.U.data.clone()); rec["logsn"].append(kreg.logsn.data.clone())
            rec["gU"].append(kreg.U.grad.clone()); rec["glogsn"].append(kreg.logsn.grad.clone())
            torch.manual_seed(1000 + i)
            lr = step_fn(i)
            rec["lr"].append(torch.tensor(float(lr)))
            rec["U_new"].append(kreg.U.data.clone()); rec["logsn_new"].append(kreg.logsn.data.clone())
            noises.append(_replay_noise(1000 + i, shapes_fn(i)))
        for k, v in rec.items():
            out[f"{name}_{k}"] = torch.stack(v)
        return noises

    # ---- SGLD (gen_configs.py: lr0=1e-4, gamma .51, t0 100, alpha .03)
    kreg, cl = fresh()
    smp = SGLD([kreg.U, kreg.logsn], lr0=1e-4, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03)
    def sgld_step(i):
        lr = smp.get_lr(i); smp.step(lr=lr); return lr
    noises = record("sgld", smp, kreg, cl, 4, sgld_step, lambda i: [(25, 2), (2,)])
    out["sgld_xiU"] = torch.stack([n[0] for n in noises]); out["sgld_xilogsn"] = torch.stack([n[1] for n in noises])
    for i in range(4):   # pin the noise-replay assumption
        for nm, k in (("U", 0), ("logsn", 1)):
            ref = out[f"sgld_{nm}_new"][i].numpy()
            mine = osamp.sgld_step(out[f"sgld_{nm}"][i].numpy(), out[f"sgld_g{nm}"][i].numpy(), float(out["sgld_lr"][i]), noises[i][k].numpy())
            assert np.abs(ref - mine).max() < 1e-13, "noise replay mismatch (sgld)"

    # ---- pSGLD (gp.py:372-373: alpha .99, lambda 1e-8, N=5)
    kreg, cl = fresh()
    smp2 = pSGLD([kreg.U, kreg.logsn], lr0=5e-3, lr_gamma=0.51, lr_t0=100, lr_alpha=0.1, lambda_=1e-8, alpha=0.99, N=5)
    def psgld_step(i):
        lr = smp2.get_lr(i); smp2.step(lr=lr); return lr
    noises = record("psgld", smp2, kreg, cl, 4, psgld_step, lambda i: [(25, 2), (2,)])
    out["psgld_xiU"] = torch.stack([n[0] for n in noises]); out["psgld_xilogsn"] = torch.stack([n[1] for n in noises])
    out["psgld_VU_final"] = smp2.state[kreg.U]["V"].clone(); out["psgld_Vlogsn_final"] = smp2.state[kreg.logsn]["V"].clone()
    V = {"U": np.zeros((25, 2)), "logsn": np.zeros(2)}
    for i in range(4):
        for nm, k in (("U", 0), ("logsn", 1)):
            mine, V[nm] = osamp.psgld_step(out[f"psgld_{nm}"][i].numpy(), out[f"psgld_g{nm}"][i].numpy(), V[nm],
                                           float(out["psgld_lr"][i]), 0.99, 1e-8, noises[i][k].numpy())
            assert np.abs(out[f"psgld_{nm}_new"][i].numpy() - mine).max() < 1e-12, "noise replay mismatch (psgld)"
    assert np.abs(V["U"] - out["psgld_VU_final"].numpy()).max() < 1e-12

    # ---- aSGHMC: 3 burn-in steps, then 4 sampling steps with resample_mom_every=2 (iteration 4 and 6 resample)
    kreg, cl = fresh()
    smp3 = aSGHMC([kreg.U, kreg.logsn], lr=1e-2, mom_decay=5e-2, lambda_=1e-5)
    burn, k_res = 3, 2
    def asghmc_step(i):
        smp3.step(lr=1e-2, burn_in=i < burn, resample_mom_every=k_res); return 1e-2
    def shapes(i):
        res = (i >= burn) and ((i + 1) % k_res == 0)
        return ([(25, 2)] * (2 if res else 1)) + ([(2,)] * (2 if res else 1))
    noises = record("asghmc", smp3, kreg, cl, 7, asghmc_step, shapes)
    xiU, xiL, xrU, xrL = [], [], [], []
    for i, n in enumerate(noises):
        res = len(n) == 4
        xrU.append(n[0] if res else torch.zeros(25, 2)); xiU.append(n[1] if res else n[0])
        xrL.append(n[2] if res else torch.zeros(2)); xiL.append(n[3] if res else n[1])
    out.update(asghmc_xiU=torch.stack(xiU), asghmc_xilogsn=torch.stack(xiL), asghmc_xrU=torch.stack(xrU),
               asghmc_xrlogsn=torch.stack(xrL), asghmc_burn=burn, asghmc_resample_every=k_res)
    for nm, pt in (("U", kreg.U), ("logsn", kreg.logsn)):
        for k in ("tau", "g", "v_hat", "momentum"):
            out[f"asghmc_{k}_{nm}_final"] = smp3.state[pt][k].clone()
    st = {"U": osamp.asghmc_init(np.zeros((25, 2))), "logsn": osamp.asghmc_init(np.zeros(2))}
    for i in range(7):
        for nm in ("U", "logsn"):
            mine, st[nm] = osamp.asghmc_step(out[f"asghmc_{nm}"][i].numpy(), out[f"asghmc_g{nm}"][i].numpy(), st[nm], 1e-2, 5e-2, 1e-5,
                                             i < burn, k_res, out[f"asghmc_xi{nm}"][i].numpy(), out[f"asghmc_xr{nm}"][i].numpy())
            assert np.abs(out[f"asghmc_{nm}_new"][i].numpy() - mine).max() < 1e-12, ("noise replay mismatch (asghmc)", i, nm)
    save("sampler_steps", **out)


def gen_svgd(data):
    """RBFKernel (stein.py:12-34) run unmodified; phi (stein.py:75-86) restated with the reference kernel + autograd
    because the reference's SVGD.phi/step reference undefined names."""
    from samplers.stein import RBFKernel
    from oracle import samplers as osamp
    Zt, Yt, U0 = make_model(data, 5)
    g = torch.Generator().manual_seed(99)
    out = {}
    for n in (64, 257):
        X = torch.cat([U0.reshape(1, -1) + 0.1 * torch.randn(n, 50, generator=g),
                       np.log(0.1) + 0.05 * torch.randn(n, 2, generator=g)], 1)
        S = torch.randn(n, 52, generator=g) * 3.0                      # stand-in scores
        Kmod = RBFKernel()
        Xr = X.clone().requires_grad_(True)
        K_XX = Kmod(Xr, Xr.detach())
        grad_K = -torch.autograd.grad(K_XX.sum(), Xr)[0]
        phi = (K_XX.detach().matmul(S) + grad_K) / X.size(0)
        d2 = (torch.cdist(X, X) ** 2).numpy()
        med = np.median(d2)
        out.update({f"n{n}_X": X, f"n{n}_S": S, f"n{n}_K": K_XX.detach(), f"n{n}_phi": phi, f"n{n}_median": med})
        Ko, gam = osamp.rbf_kernel(X.numpy(), X.numpy())
        assert np.abs(Ko - K_XX.detach().numpy()).max() < 1e-10
        assert np.abs(osamp.svgd_phi(X.numpy(), S.numpy()) - phi.numpy()).max() < 1e-10 * np.abs(phi.numpy()).max() + 1e-12
        Kf = RBFKernel(sigma=0.7)(X, X)
        out[f"n{n}_K_sigma07"] = Kf
    save("svgd", **out)


class NN(torch.nn.Module):
    """notebooks/jai/nn.ipynb cell 4, verbatim structure."""

    def __init__(self, input_size, hidden_size=10):
        super(NN, self).__init__()
        nn = torch.nn
        self.layers = nn.Sequential(
            nn.Linear(input_size, hidden_size), nn.ELU(),
            nn.Linear(int(hidden_size * 1.0), int(hidden_size * 1.0)), nn.ELU(),
            nn.Linear(hidden_size, input_size))

    def forward(self, t, x):
        size = x.size()
        x = x.view(-1)
        x = self.layers(x)
        x = x.view(size)
        return x


def gen_mlp(data):
    """MLP field: per-row odeint (nn.ipynb cell 10 loops over trajectory rows), rk4, autograd and adjoint gradients."""
    x0, t = data["x0"], data["t"]
    Xt = torch.from_numpy(data["X"])
    out = dict(x0=x0, t=t, X=Xt)
    for H, P in ((20, 3), (64, 2)):
        torch.manual_seed(100 + H)
        thetas, sols, losses, sqs, gd, ga = [], [], [], [], [], []
        for p in range(P):
            net = NN(2, H)
            for m_ in net.modules():
                if isinstance(m_, torch.nn.Linear):
                    torch.nn.init.uniform_(m_.weight, a=-0.5, b=0.5)          # nn.ipynb cell 4 init_normal
            params = list(net.parameters())
            thetas.append(torch.cat([q.detach().reshape(-1) for q in params]))
            for mode, odeint in (("d", torchdiffeq.odeint), ("a", torchdiffeq.odeint_adjoint)):
                net.zero_grad()
                loss = 0
                rows = []
                for r in range(x0.size(0)):
                    xode = odeint(net, x0[r], t, method="rk4")
                    rows.append(xode.detach())
                    loss = loss + torch.sum((Xt[r] - xode) ** 2)
                sq = loss.detach().clone()
                loss = loss + 0.5 * sum([torch.sum(q ** 2) for q in params])
                loss.backward()
                g = torch.cat([q.grad.reshape(-1) for q in params])
                (gd if mode == "d" else ga).append(g)
            sols.append(torch.stack(rows, 1))               # [T,N,2]
            losses.append(loss.detach()); sqs.append(sq)
        out.update({f"h{H}_theta": torch.stack(thetas), f"h{H}_sol": torch.stack(sols, 1), f"h{H}_loss": torch.stack(losses),
                    f"h{H}_sqerr": torch.stack(sqs), f"h{H}_g_discrete": torch.stack(gd), f"h{H}_g_adjoint": torch.stack(ga)})
    save("mlp", **out)


class _Counted(torch.nn.Module):
    def __init__(self, inner):
        super().__init__()
        self.inner, self.nfe = inner, 0

    def forward(self, t, y):
        self.nfe += 1
        return self.inner(t, y)


def gen_dopri5(data):
    """Adaptive dopri5 through the reference, one odeint call per trajectory row (per-row controller)."""
    torch.cholesky = torch.linalg.cholesky
    Zt, Yt, U0 = make_model(data, 5)
    x0, t = data["x0"], data["t"]
    kreg = gp.KernelRegression(U0.clone(), Zt, 1.0, 0.75, 0.1)
    torch.manual_seed(120)
    net = NN(2, 20)
    for m_ in net.modules():
        if isinstance(m_, torch.nn.Linear):
            torch.nn.init.uniform_(m_.weight, a=-0.5, b=0.5)
    out = dict(x0=x0, t=t, U=U0, Z=Zt, theta=torch.cat([q.detach().reshape(-1) for q in net.parameters()]))
    cases = {"default": dict(), "loose": dict(rtol=1e-5, atol=1e-7), "firststep": dict(rtol=1e-5, atol=1e-7, options=dict(first_step=0.5))}
    for name, kw in cases.items():
        for fname, fobj, mk in (("npde", kreg, lambda r: x0[r:r + 1]), ("mlp", net, lambda r: x0[r])):
            sols, nfes = [], []
            for r in range(x0.size(0)):
                cf = _Counted(fobj)
                with torch.no_grad():
                    sol = torchdiffeq.odeint(cf, mk(r), t, method="dopri5" if "options" in kw else None, **kw)
                sols.append(sol.reshape(len(t), 2)); nfes.append(cf.nfe)
            out[f"{name}_{fname}_sol"] = torch.stack(sols, 1)          # [T,N,2]
            out[f"{name}_{fname}_nfe"] = np.array(nfes)
    # gradients: the reference's adjoint through dopri5 (per-row solves, rtol 1e-7 / atol 1e-9 = odeint defaults)
    Xt = torch.from_numpy(data["X"])
    for q in net.parameters():
        q.grad = None
    loss = 0
    for r in range(x0.size(0)):
        xode = torchdiffeq.odeint_adjoint(net, x0[r], t, rtol=1e-7, atol=1e-9, method="dopri5")
        loss = loss + torch.sum((Xt[r] - xode) ** 2)
    out["mlp_sqerr"] = loss.detach().clone()
    loss = loss + 0.5 * sum([torch.sum(q ** 2) for q in net.parameters()])
    loss.backward()
    out["mlp_loss"] = loss.detach()
    out["mlp_grad_adjoint"] = torch.cat([q.grad.reshape(-1) for q in net.parameters()])
    out["X"] = Xt
    kreg.zero_grad()
    loss = 0
    for r in range(x0.size(0)):
        xode = torchdiffeq.odeint_adjoint(kreg, x0[r:r + 1], t, rtol=1e-7, atol=1e-9, method="dopri5")[:, 0]
        loss = loss + torch.sum((Yt[r] - xode) ** 2 / (2 * torch.exp(kreg.logsn) ** 2))
    loss = loss + torch.numel(Yt) * torch.sum(kreg.logsn) / 2
    loss = loss + torch.sum(torch.diag(torch.mm(kreg.U.t(), torch.mm(kreg.Kzzinv, kreg.U)))) / 2
    loss.backward()
    out["npde_loss"] = loss.detach()
    out["npde_gU_adjoint"] = kreg.U.grad.clone()
    out["npde_glogsn_adjoint"] = kreg.logsn.grad.clone()
    out["Y"] = Yt
    trev = torch.linspace(3., 0., 7)
    sols = []
    for r in range(x0.size(0)):
        with torch.no_grad():
            sols.append(torchdiffeq.odeint(kreg, x0[r:r + 1], trev, rtol=1e-5, atol=1e-7).reshape(7, 2))
    out["rev_t"] = trev
    out["rev_npde_sol"] = torch.stack(sols, 1)
    save("dopri5", **out)


def gen_hamcmc():
    """HAMCMC (langevin.py:619-1107) on a convex quadratic with two parameter tensors; noise replayed per step."""
    import io, contextlib
    from samplers.langevin import HAMCMC
    from oracle import samplers as osamp
    g = torch.Generator().manual_seed(5)
    d = 10
    Q = torch.randn(d, d, generator=g)
    Aq = Q @ Q.t() / d + 0.5 * torch.eye(d)
    a = torch.nn.Parameter(torch.randn(3, 2, generator=g))
    b = torch.nn.Parameter(torch.randn(4, generator=g))

    def closure(add_prior=True):
        th = torch.cat([a.reshape(-1), b.reshape(-1)])
        return 0.5 * th @ (Aq @ th)

    memory = 2
    smp = HAMCMC([a, b], memory=memory, lr0=5e-2, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3, H_gamma=1.0, trust_reg=1.0)
    M = memory + 1
    nsteps = 100 + 2 * M - 1 + 8
    thetas, grads, xis, lrs = [torch.cat([a.detach().reshape(-1), b.detach().reshape(-1)]).clone()], [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(nsteps):
            smp.zero_grad()
            smp.loss = closure()
            smp.loss.backward()
            grads.append(torch.cat([a.grad.reshape(-1), b.grad.reshape(-1)]).clone())
            lr = smp.get_lr(i)
            torch.manual_seed(7000 + i)
            if i < 2 * M - 1 + 100:
                smp.step_without_metric(lr=lr, add_noise=True, add_params=(i >= 100))
            else:
                smp.step(lr=lr, add_noise=True)
            torch.manual_seed(7000 + i)
            xis.append(torch.randn(d))
            lrs.append(lr)
            thetas.append(torch.cat([a.detach().reshape(-1), b.detach().reshape(-1)]).clone())
    out = dict(A=Aq, theta=torch.stack(thetas), grad=torch.stack(grads), xi=torch.stack(xis), lr=np.array(lrs), memory=memory,
               n_pairs=len(smp.state["memory"]["param_diff"]))
    # pin the oracle restatement (and the noise replay) right here
    orc = osamp.HAMCMC(memory=memory, H_gamma=1.0, trust_reg=1.0)
    for i in range(nsteps):
        th, gr, xi = out["theta"][i].numpy(), out["grad"][i].numpy(), out["xi"][i].numpy()
        if i < 2 * M - 1 + 100:
            new = orc.step_without_metric(th, gr, lrs[i], xi, add_params=(i >= 100))
        else:
            new = orc.step(gr, lrs[i], xi)
        assert np.abs(new - out["theta"][i + 1].numpy()).max() < 1e-10, ("hamcmc oracle mismatch", i)
    save("hamcmc", **out)


def gen_mala():
    """MALA (langevin.py:13-149) on a quadratic negative log posterior over two parameter tensors: per step the state before,
    the gradients, the replayed noise and log(u), the proposal, and the reference's own accept flag.  The generator also pins
    the aliased-state quirk: the reference's decisions equal the oracle's ``aliased=True`` ratio and the parameters are never
    restored after a rejection."""
    from samplers.langevin import MALA
    from oracle import samplers as osamp
    gen = torch.Generator().manual_seed(77)
    cA = 0.5 + 4.0 * torch.rand(6, 2, generator=gen)
    cB = 0.5 + 4.0 * torch.rand(3, generator=gen)
    A = torch.nn.Parameter(torch.randn(6, 2, generator=gen))
    B = torch.nn.Parameter(torch.randn(3, generator=gen))
    lr = 0.15

    def closure(add_prior=True):
        return 0.5 * (cA * A * A).sum() + 0.5 * (cB * B * B).sum() + 0.1 * (A.sum() * B.sum())

    smp = MALA([A, B], lr=lr)
    rec = {k: [] for k in ("A", "B", "gA", "gB", "xiA", "xiB", "logu", "A_new", "B_new", "gA_new", "gB_new", "loss", "loss_new", "accepted",
                           "A_after", "B_after")}
    n_rej = 0
    for i in range(12):
        smp.zero_grad()
        smp.loss = closure()
        smp.loss.backward()
        rec["A"].append(A.data.clone()); rec["B"].append(B.data.clone())
        rec["gA"].append(A.grad.clone()); rec["gB"].append(B.grad.clone())
        rec["loss"].append(smp.loss.detach().clone())
        torch.manual_seed(500 + i)
        smp.step()
        rec["A_new"].append(A.data.clone()); rec["B_new"].append(B.data.clone())
        params, acc = smp.accept_or_reject(closure)
        rec["accepted"].append(torch.tensor(int(acc)))
        rec["A_after"].append(A.data.clone()); rec["B_after"].append(B.data.clone())
        rec["gA_new"].append(A.grad.clone()); rec["gB_new"].append(B.grad.clone())
        rec["loss_new"].append(closure().detach().clone())
        torch.manual_seed(500 + i)
        xa, xb = torch.randn(6, 2), torch.randn(3)
        u = torch.rand(1)
        rec["xiA"].append(xa); rec["xiB"].append(xb); rec["logu"].append(torch.log(u)[0])
        # pin: proposal == SGLD update with the replayed noise; nothing restored; decision == aliased ratio
        assert np.abs(osamp.sgld_step(rec["A"][-1].numpy(), rec["gA"][-1].numpy(), lr, xa.numpy()) - rec["A_new"][-1].numpy()).max() < 1e-13
        assert torch.equal(rec["A_after"][-1], rec["A_new"][-1]) and torch.equal(rec["B_after"][-1], rec["B_new"][-1])
        th0 = np.concatenate([rec["A"][-1].numpy().ravel(), rec["B"][-1].numpy().ravel()])[None]
        th1 = np.concatenate([rec["A_new"][-1].numpy().ravel(), rec["B_new"][-1].numpy().ravel()])[None]
        g0 = np.concatenate([rec["gA"][-1].numpy().ravel(), rec["gB"][-1].numpy().ravel()])[None]
        g1 = np.concatenate([rec["gA_new"][-1].numpy().ravel(), rec["gB_new"][-1].numpy().ravel()])[None]
        la = osamp.mala_log_alpha(th0, th1, g0, g1, np.array([float(rec["loss"][-1])]), np.array([float(rec["loss_new"][-1])]), lr, aliased=True)
        assert bool(osamp.mala_accept(la, np.array([float(rec["logu"][-1])]))[0]) == bool(acc), ("aliased-state ratio does not explain the reference", i)
        n_rej += int(not acc)
    assert 0 < n_rej < 12, "want both accepted and rejected steps in the fixture (got %d rejections)" % n_rej
    save("mala_steps", lr=lr, cA=cA, cB=cB, **{k: torch.stack(v) for k, v in rec.items()})


def gen_cyclical(data):
    """cSGLD (langevin.py:1600-1724) and acSGHMC (hamiltonian.py:167-326): cosine step-size cycles with the noise gated by
    r > beta.  Step sequences from the reference with seeded synthetic gradients and replayed noise, as in gen_sampler_steps."""
    from samplers.langevin import cSGLD
    from samplers.hamiltonian import acSGHMC
    from oracle import samplers as osamp
    torch.cholesky = torch.linalg.cholesky
    Zt, Yt, U0 = make_model(data, 5)
    gg = torch.Generator().manual_seed(777)
    out = {}
    num_iters, M, beta = 12, 3, 0.25            # cycle length (12 + 3) // 3 = 5: r in {0.8, 0, 0.2, 0.4, 0.6, ...}
    out.update(num_iters=num_iters, M=M, beta=beta)

    def run(name, smp, kreg, step_fn, draws_fn):
        rec = {k: [] for k in ("U", "logsn", "gU", "glogsn", "U_new", "logsn_new", "lr", "r")}
        noises = []
        for i in range(num_iters):
            kreg.U.grad = 50.0 * torch.randn(25, 2, generator=gg)
            kreg.logsn.grad = 5.0 * torch.randn(2, generator=gg)
            rec["U"].append(kreg.U.data.clone()); rec["logsn"].append(kreg.logsn.data.clone())
            rec["gU"].append(kreg.U.grad.clone()); rec["glogsn"].append(kreg.logsn.grad.clone())
            torch.manual_seed(2000 + i)
            lr = smp.get_lr(i)
            step_fn(i, lr)
            rec["lr"].append(torch.tensor(float(lr))); rec["r"].append(torch.tensor(float(smp._r(i))))
            rec["U_new"].append(kreg.U.data.clone()); rec["logsn_new"].append(kreg.logsn.data.clone())
            noises.append(_replay_noise(2000 + i, draws_fn(i, float(smp._r(i)))))
        for k, v in rec.items():
            out[f"{name}_{k}"] = torch.stack(v)
        return noises

    # ---- cSGLD
    kreg = gp.KernelRegression(U0.clone(), Zt, 1.0, 0.75, 0.1)
    smp = cSGLD([kreg.U, kreg.logsn], lr0=2e-4, M=M, beta=beta)
    smp.num_iters = num_iters                                  # set by sample() (langevin.py:1671)
    noises = run("csgld", smp, kreg, lambda i, lr: smp.step(iter_num=i, lr=lr),
                 lambda i, r: [(25, 2), (2,)] if r > beta else [])
    out["csgld_xiU"] = torch.stack([n[0] if n else torch.zeros(25, 2) for n in noises])
    out["csgld_xilogsn"] = torch.stack([n[1] if n else torch.zeros(2) for n in noises])
    gated = 0
    for i in range(num_iters):
        r = osamp.cyclical_r(i, num_iters, M)
        assert r == float(out["csgld_r"][i]) and osamp.cyclical_lr(i, 2e-4, num_iters, M) == float(out["csgld_lr"][i])
        gated += int(not r > beta)
        for nm, k in (("U", 0), ("logsn", 1)):
            xi = noises[i][k].numpy() if r > beta else None
            mine = osamp.sgld_step(out[f"csgld_{nm}"][i].numpy(), out[f"csgld_g{nm}"][i].numpy(), float(out["csgld_lr"][i]), xi)
            assert np.abs(out[f"csgld_{nm}_new"][i].numpy() - mine).max() < 1e-13, ("noise replay mismatch (csgld)", i)
    assert 0 < gated < num_iters

    # ---- acSGHMC: 4 burn-in iterations, then sampling with momentum resampling every 3rd iteration
    kreg = gp.KernelRegression(U0.clone(), Zt, 1.0, 0.75, 0.1)
    smp2 = acSGHMC([kreg.U, kreg.logsn], lr0=1e-2, M=M, beta=beta, mom_decay=5e-2, lambda_=1e-5)
    smp2.num_iters = num_iters
    burn, k_res = 4, 3
    out.update(acsghmc_burn=burn, acsghmc_resample_every=k_res)

    def draws(i, r):
        res = (i >= burn) and ((i + 1) % k_res == 0)
        per = (1 if res else 0) + (1 if r > beta else 0)
        return [(25, 2)] * per + [(2,)] * per
    noises = run("acsghmc", smp2, kreg,
                 lambda i, lr: smp2.step(lr=lr, iter_num=i, burn_in=i < burn, resample_mom_every=k_res), draws)
    xiU, xiL, xrU, xrL = [], [], [], []
    for i, n in enumerate(noises):
        r = float(out["acsghmc_r"][i])
        res = (i >= burn) and ((i + 1) % k_res == 0)
        per = len(n) // 2
        nU, nL = list(n[:per]), list(n[per:])
        xrU.append(nU.pop(0) if res else torch.zeros(25, 2)); xrL.append(nL.pop(0) if res else torch.zeros(2))
        xiU.append(nU.pop(0) if r > beta else torch.zeros(25, 2)); xiL.append(nL.pop(0) if r > beta else torch.zeros(2))
    out.update(acsghmc_xiU=torch.stack(xiU), acsghmc_xilogsn=torch.stack(xiL), acsghmc_xrU=torch.stack(xrU),
               acsghmc_xrlogsn=torch.stack(xrL))
    for nm, pt in (("U", kreg.U), ("logsn", kreg.logsn)):
        for k in ("tau", "g", "v_hat", "momentum"):
            out[f"acsghmc_{k}_{nm}_final"] = smp2.state[pt][k].clone()
    st = {"U": osamp.asghmc_init(np.zeros((25, 2))), "logsn": osamp.asghmc_init(np.zeros(2))}
    for i in range(num_iters):
        r = float(out["acsghmc_r"][i])
        for nm in ("U", "logsn"):
            xi = out[f"acsghmc_xi{nm}"][i].numpy() if r > beta else None
            mine, st[nm] = osamp.asghmc_step(out[f"acsghmc_{nm}"][i].numpy(), out[f"acsghmc_g{nm}"][i].numpy(), st[nm],
                                             float(out["acsghmc_lr"][i]), 5e-2, 1e-5, i < burn, k_res, xi,
                                             out[f"acsghmc_xr{nm}"][i].numpy())
            assert np.abs(out[f"acsghmc_{nm}_new"][i].numpy() - mine).max() < 1e-12, ("noise replay mismatch (acsghmc)", i, nm)
    save("cyclical_steps", **out)


def gen_predictive(data):
    """Posterior-predictive ensemble exactly as gp.py:440-464 computes it: every chain entry re-integrated with
    ``odeint_adjoint(kreg, x0_, t_)`` (default dopri5), then np.mean / np.std over the chain."""
    torch.cholesky = torch.linalg.cholesky
    Zt, Yt, U0 = make_model(data, 5)
    kreg = gp.KernelRegression(U0.clone(), Zt, 1.0, 0.75, 0.1)
    gen = torch.Generator().manual_seed(99)
    chain = [([[(U0 + 0.05 * torch.randn(25, 2, generator=gen)).numpy(), np.log(0.1) * np.ones(2)]], True) for _ in range(6)]
    rs = np.random.RandomState(5)
    x0_ = torch.from_numpy(2 * 3 * rs.uniform(size=[4, 2]) - 3)
    t_ = torch.linspace(0., 14., 30)
    out = dict(U=np.stack([c[0][0][0] for c in chain]), x0=x0_, t=t_)
    for name, kw in (("dopri5", {}), ("rk4", dict(method="rk4"))):
        xode_gp = []
        for i in range(len(chain)):
            kreg.U.data = torch.from_numpy(chain[i][0][0][0])
            with torch.no_grad():
                xode_gp.append(np.transpose(gp.odeint(kreg, x0_, t_, **kw).detach().numpy(), [1, 0, 2]))
        out[f"{name}_traj"] = np.stack(xode_gp)                        # [S, N, T, 2]
        mean, std = np.zeros_like(xode_gp[0]), np.zeros_like(xode_gp[0])
        for i in range(mean.shape[0]):                                  # gp.py:460-464
            for c in range(2):
                mean[i, :, c] = np.mean([xode_gp[j][i, :, c] for j in range(len(xode_gp))], axis=0)
                std[i, :, c] = np.std([xode_gp[j][i, :, c] for j in range(len(xode_gp))], axis=0)
        out[f"{name}_mean"], out[f"{name}_std"] = mean, std
    save("predictive", **out)


def gen_hamcmc_contiguous():
    """HAMCMC2 / HAMCMC3 / HAMCMC4 (langevin.py:1109-1470) on the convex quadratic of gen_hamcmc; noise replayed per step.
    M warm-up steps (sample(): i < self.memory, :1254) then metric steps, driven step by step like ``sample`` does."""
    import io, contextlib, warnings
    from samplers import langevin
    from oracle import samplers as osamp
    out = {}
    for variant in (2, 3, 4):
        g = torch.Generator().manual_seed(5)
        d = 10
        Q = torch.randn(d, d, generator=g)
        Aq = Q @ Q.t() / d + 0.5 * torch.eye(d)
        a = torch.nn.Parameter(torch.randn(3, 2, generator=g))
        b = torch.nn.Parameter(torch.randn(4, generator=g))

        def closure(add_prior=True):
            th = torch.cat([a.reshape(-1), b.reshape(-1)])
            return 0.5 * th @ (Aq @ th)

        memory = 3
        cls = getattr(langevin, "HAMCMC%d" % variant)
        with contextlib.redirect_stdout(io.StringIO()):
            smp = cls([a, b], memory=memory, lr0=2e-2, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3, H_gamma=1.0, trust_reg=1.0)
        M = memory + 1
        nsteps = M + 10
        thetas, grads, xis, lrs = [torch.cat([a.detach().reshape(-1), b.detach().reshape(-1)]).clone()], [], [], []
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(nsteps):
                smp.zero_grad()
                smp.loss = closure()
                smp.loss.backward()
                grads.append(torch.cat([a.grad.reshape(-1), b.grad.reshape(-1)]).clone())
                lr = smp.get_lr(i)
                torch.manual_seed(9000 + 100 * variant + i)
                if i < M:
                    smp.step_without_metric(lr=lr, add_noise=True)
                else:
                    smp.step(lr=lr, add_noise=True)
                torch.manual_seed(9000 + 100 * variant + i)
                xis.append(torch.randn(d))
                lrs.append(lr)
                thetas.append(torch.cat([a.detach().reshape(-1), b.detach().reshape(-1)]).clone())
        th_all, gr_all, xi_all = torch.stack(thetas), torch.stack(grads), torch.stack(xis)
        assert bool(torch.isfinite(th_all).all())
        # pin the oracle restatement (and the noise replay) right here
        orc = osamp.HAMCMCContiguous(variant, memory=memory, H_gamma=1.0, trust_reg=1.0)
        for i in range(nsteps):
            th, gr, xi = th_all[i].numpy(), gr_all[i].numpy(), xi_all[i].numpy()
            new = orc.step_without_metric(th, gr, lrs[i], xi) if i < M else orc.step(gr, lrs[i], xi)
            err = np.abs(new - th_all[i + 1].numpy()).max() / max(1.0, np.abs(th_all[i + 1].numpy()).max())
            assert err < 1e-10, ("hamcmc%d oracle mismatch" % variant, i, err)
        out.update({"theta%d" % variant: th_all, "grad%d" % variant: gr_all, "xi%d" % variant: xi_all, "lr%d" % variant: np.array(lrs)})
        out["A"] = Aq
        out["memory"] = memory
    save("hamcmc_contiguous", **out)


def gen_dopri5_batched(data):
    """Adaptive dopri5 through the reference with the WHOLE y0 [N, 2] in one odeint call: one controller, error ratio and
    initial-step norms pooled over all N x 2 elements (misc.py:146-157, 116-143) -- what gp.py:346 / 452
    run.  The accept / reject sequence is recorded by wrapping (not changing) Dopri5Solver._adaptive_dopri5_step.  The MLP is the
    notebook's layers applied row-wise (nn.ipynb's NN.forward flattens its input, so it only takes one row per call)."""
    from torchdiffeq._impl import dopri5 as ref_d5
    torch.cholesky = torch.linalg.cholesky
    Zt, Yt, U0 = make_model(data, 5)
    x0, t = data["x0"], data["t"]
    kreg = gp.KernelRegression(U0.clone(), Zt, 1.0, 0.75, 0.1)
    torch.manual_seed(120)
    net = NN(2, 20)
    for m_ in net.modules():
        if isinstance(m_, torch.nn.Linear):
            torch.nn.init.uniform_(m_.weight, a=-0.5, b=0.5)

    class RowWise(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, t, x):
            return self.inner.layers(x)

    accepts = []
    orig = ref_d5.Dopri5Solver._adaptive_dopri5_step

    def recording(self, rk_state):
        new = orig(self, rk_state)
        accepts.append((bool(new.t1 > rk_state.t1), float(rk_state.dt)))
        return new
    ref_d5.Dopri5Solver._adaptive_dopri5_step = recording
    out = dict(x0=x0, t=t, U=U0, Z=Zt, theta=torch.cat([q.detach().reshape(-1) for q in net.parameters()]))
    cases = {"default": dict(), "loose": dict(rtol=1e-5, atol=1e-7), "firststep": dict(rtol=1e-5, atol=1e-7, options=dict(first_step=0.5))}
    try:
        for name, kw in cases.items():
            for fname, fobj in (("npde", kreg), ("mlp", RowWise(net))):
                cf = _Counted(fobj)
                del accepts[:]
                with torch.no_grad():
                    sol = torchdiffeq.odeint(cf, x0, t, method="dopri5" if "options" in kw else None, **kw)
                out[f"{name}_{fname}_sol"] = sol                                   # [T,N,2]
                out[f"{name}_{fname}_nfe"] = np.array(cf.nfe)
                out[f"{name}_{fname}_accept"] = np.array([a for a, _ in accepts], dtype=np.int8)
                out[f"{name}_{fname}_dt"] = np.array([d for _, d in accepts])
                assert cf.nfe == (1 if "options" in kw else 2) + 6 * len(accepts)
        # closure gradients through the reference's adjoint, one batched call (rtol 1e-7 / atol 1e-9)
        kreg.zero_grad()
        xode = torchdiffeq.odeint_adjoint(kreg, x0, t, rtol=1e-7, atol=1e-9, method="dopri5").permute(1, 0, 2)
        loss = torch.sum((Yt - xode) ** 2 / (2 * torch.exp(kreg.logsn) ** 2)) + torch.numel(Yt) * torch.sum(kreg.logsn) / 2
        loss = loss + torch.sum(torch.diag(torch.mm(kreg.U.t(), torch.mm(kreg.Kzzinv, kreg.U)))) / 2
        loss.backward()
        out.update(npde_loss=loss.detach(), npde_gU_adjoint=kreg.U.grad.clone(), npde_glogsn_adjoint=kreg.logsn.grad.clone(), Y=Yt)
        rw = RowWise(net)
        Xt = torch.from_numpy(data["X"])
        for q in net.parameters():
            q.grad = None
        xode = torchdiffeq.odeint_adjoint(rw, x0, t, rtol=1e-7, atol=1e-9, method="dopri5").permute(1, 0, 2)
        sq = torch.sum((Xt - xode) ** 2)
        loss = sq + 0.5 * sum([torch.sum(q ** 2) for q in net.parameters()])
        loss.backward()
        out.update(mlp_sqerr=sq.detach(), mlp_loss=loss.detach(), mlp_grad_adjoint=torch.cat([q.grad.reshape(-1) for q in net.parameters()]), X=Xt)
    finally:
        ref_d5.Dopri5Solver._adaptive_dopri5_step = orig
    # pin the oracle's pooled controller (oracle/dopri5.py takes any y0 shape and pools like misc.py:146-157) right here
    from oracle import dopri5 as od5, npde as onpde
    fo = onpde.NPDEField(U0.numpy()[None], Zt.numpy(), 1.0, 0.75)
    so, st = od5.odeint_dopri5(lambda y: fo.f(y[None])[0], x0.numpy(), t.numpy().astype(np.float64), rtol=1e-5, atol=1e-7)
    na = int(out["loose_npde_accept"].sum())
    assert (st["accepted"], st["rejected"]) == (na, len(out["loose_npde_accept"]) - na), (st, na)
    assert np.abs(so - out["loose_npde_sol"].numpy()).max() < 1e-9
    save("dopri5_batched", **out)


def gen_tuple_and_time(data):
    """(1) Tuple states (misc.py:175-182, api_tests.py:19-38): odeint(tuple_f, (y0a, y0b), t) with tuple_f applying the npde field to
    each element, rk4 and dopri5 (one controller for the tuple: per-tensor error means, max over the tuple).
    (2) dL/dt through odeint_adjoint (adjoint.py:68-76, 99-100; gradient_tests.py:19-37 gradchecks (y0, t)): t.grad of a weighted
    sum of the solution, npde field (rk4 and dopri5) and the row-wise MLP (rk4)."""
    from torchdiffeq._impl import dopri5 as ref_d5
    torch.cholesky = torch.linalg.cholesky
    Zt, Yt, U0 = make_model(data, 5)
    x0, t = data["x0"], data["t"].to(torch.float64)
    kreg = gp.KernelRegression(U0.clone(), Zt, 1.0, 0.75, 0.1)
    out = dict(x0=x0, t=t, U=U0, Z=Zt)
    tuple_f = lambda tt, y: (kreg(tt, y[0]), kreg(tt, y[1]))
    with torch.no_grad():
        ra, rb = torchdiffeq.odeint(tuple_f, (x0[:3], x0[3:]), t, method="rk4")
        out.update(tuple_rk4_a=ra, tuple_rk4_b=rb)
        accepts = []
        orig = ref_d5.Dopri5Solver._adaptive_dopri5_step

        def recording(self, rk_state):
            new = orig(self, rk_state)
            accepts.append(bool(new.t1 > rk_state.t1))
            return new
        ref_d5.Dopri5Solver._adaptive_dopri5_step = recording
        try:
            da, db = torchdiffeq.odeint(tuple_f, (x0[:3], x0[3:]), t, rtol=1e-5, atol=1e-7)
        finally:
            ref_d5.Dopri5Solver._adaptive_dopri5_step = orig
        out.update(tuple_dopri5_a=da, tuple_dopri5_b=db, tuple_dopri5_accept=np.array(accepts, dtype=np.int8))
        # the same rows integrated as ONE tensor take different steps (mean over all rows instead of max of per-tensor means)
        del accepts[:]
        ref_d5.Dopri5Solver._adaptive_dopri5_step = recording
        try:
            torchdiffeq.odeint(kreg, x0, t, rtol=1e-5, atol=1e-7)
        finally:
            ref_d5.Dopri5Solver._adaptive_dopri5_step = orig
        out["single_dopri5_accept"] = np.array(accepts, dtype=np.int8)
    g = torch.Generator().manual_seed(31)
    w = torch.randn(len(t), 5, 2, generator=g)
    out["w"] = w
    for name, kw in (("rk4", dict(method="rk4")), ("dopri5", dict(rtol=1e-7, atol=1e-9, method="dopri5"))):
        tr = t.clone().requires_grad_(True)
        kreg.zero_grad()
        sol = torchdiffeq.odeint_adjoint(kreg, x0, tr, **kw)
        (sol * w).sum().backward()
        out[f"npde_{name}_gt"] = tr.grad.clone()
        out[f"npde_{name}_gU"] = kreg.U.grad.clone()
    torch.manual_seed(120)
    net = NN(2, 20)
    for m_ in net.modules():
        if isinstance(m_, torch.nn.Linear):
            torch.nn.init.uniform_(m_.weight, a=-0.5, b=0.5)

    class RowWise(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, tt, x):
            return self.inner.layers(x)
    tr = t.clone().requires_grad_(True)
    sol = torchdiffeq.odeint_adjoint(RowWise(net), x0, tr, method="rk4")
    (sol * w).sum().backward()
    out.update(theta=torch.cat([q.detach().reshape(-1) for q in net.parameters()]), mlp_rk4_gt=tr.grad.clone())
    save("tuple_time", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "hamcmc_contiguous":
        gen_hamcmc_contiguous()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dopri5_batched":
        gen_dopri5_batched(make_data())
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tuple_time":
        gen_tuple_and_time(make_data())
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "predictive":
        gen_predictive(make_data())
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mala":
        gen_mala()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cyclical":
        gen_cyclical(make_data())
        sys.exit(0)
    data = make_data()
    save("vdp_data", x0=data["x0"], t=data["t"], X=data["X"], Y=data["Y"])
    gen_npde(data, 5, 4, "npde_m5")
    gen_npde(data, 3, 2, "npde_m3")
    gen_grid_options(data)
    gen_sampler_steps(data)
    gen_svgd(data)
    gen_mlp(data)
    gen_dopri5(data)
    gen_dopri5_batched(data)
    gen_tuple_and_time(data)
    gen_hamcmc()
    gen_hamcmc_contiguous()
    gen_mala()
    gen_cyclical(data)
    gen_predictive(data)
