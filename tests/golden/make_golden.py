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


if __name__ == "__main__":
    data = make_data()
    save("vdp_data", x0=data["x0"], t=data["t"], X=data["X"], Y=data["Y"])
    gen_npde(data, 5, 4, "npde_m5")
    gen_npde(data, 3, 2, "npde_m3", methods=("rk4",))
    gen_grid_options(data)
