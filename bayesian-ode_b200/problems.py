"""Experiment inputs of the reference's Van der Pol driver, rebuilt for the batched samplers.

  * toy ODEs                    scripts/vanderpol/gp.py:28-38 (VDP / FHN / LV)
  * synthetic data              notebooks/jai/gp.ipynb cell 3 (N=5, R=3, x0 ~ U(-R,R)^2, t = linspace(0,7,40), noise .1)
  * inducing grid               gp.py:315-318 (M x M over the data bounding box, ordering xv.T.flatten())
  * gradient-matching init      gp.py:323-331 (whitened U0)
One-off float64 host work (a few ms), exactly as the reference does it before sampling starts; none of it is on the
per-step path.
"""
import numpy as np
import torch

from .fields import rbf_kernel


def vdp(x):
    return torch.cat([x[:, 1:2], 1 * (1 - x[:, 0:1] ** 2) * x[:, 1:2] - x[:, 0:1]], 1)


def fhn(x):
    return torch.cat([3 * (x[:, 0:1] - x[:, 0:1] ** 3 / 3. + x[:, 1:2]), (0.2 - 3 * x[:, 0:1] - 0.2 * x[:, 1:2]) / 3.], 1)


def lv(x):
    return torch.cat([1.5 * x[:, 0:1] - x[:, 0:1] * x[:, 1:2], -3 * x[:, 1:2] + x[:, 0:1] * x[:, 1:2]], 1)


ODES = {"VDP": vdp, "FHN": fhn, "LV": lv}


def _rk4_38(f, x0, t):
    """Data synthesis only: the 3/8-rule step torchdiffeq's 'rk4' takes (rk_common.py:72-78), float64 on the host."""
    t = t.to(x0.dtype)
    out, y = [x0], x0
    for i in range(t.numel() - 1):
        dt = t[i + 1] - t[i]
        k1 = f(y)
        k2 = f(y + dt * k1 / 3)
        k3 = f(y + dt * (k1 / -3 + k2))
        k4 = f(y + dt * (k1 - k2 + k3))
        y = y + (k1 + 3 * k2 + 3 * k3 + k4) * (dt / 8)
        out.append(y)
    return torch.stack(out)


def make_dataset(ode="VDP", seed=0, N=5, R=3.0, T=40, t_end=7.0, noise=0.1):
    """dict(N, R, noise, x0 [N,2], t [T] float32, X [N,T,2], Y [N,T,2], ODE) -- the pickle layout gp.py:310 unpacks."""
    rng = np.random.default_rng(seed)
    x0 = torch.from_numpy(2 * R * rng.random((N, 2)) - R)
    t = torch.linspace(0., t_end, T, dtype=torch.float32)
    X = _rk4_38(ODES[ode], x0, t).permute(1, 0, 2).contiguous()
    Y = X + noise * torch.from_numpy(rng.standard_normal(tuple(X.shape)))
    return dict(N=N, R=R, noise=noise, x0=x0, t=t, X=X.numpy(), Y=Y.numpy(), ODE=ode)


def inducing_grid(Y, M):
    Y = np.asarray(Y)
    xv = np.linspace(np.min(Y[..., 0]), np.max(Y[..., 0]), M)
    yv = np.linspace(np.min(Y[..., 1]), np.max(Y[..., 1]), M)
    xv, yv = np.meshgrid(xv, yv)
    return torch.from_numpy(np.array([xv.T.flatten(), yv.T.flatten()]).T)


def gradient_matching_init(Y, t, Z, sf, ell):
    Yt = torch.as_tensor(Y, dtype=torch.float64)
    D = Yt.shape[-1]
    t = torch.as_tensor(t)
    F_ = ((Yt[:, 1:, :] - Yt[:, :-1, :]) / (t[1] - t[0]).double()).contiguous().view(-1, D)
    Z_ = Yt[:, :-1, :].contiguous().view(-1, D)
    Z = torch.as_tensor(Z, dtype=torch.float64)
    Kxz = rbf_kernel(Z, Z_, sf, ell)
    Kinv = (rbf_kernel(Z_, Z_, sf, ell) + 0.2 * torch.eye(Z_.shape[0], dtype=torch.float64)).inverse()
    U0 = torch.mm(torch.mm(Kxz, Kinv), F_)
    Linv = torch.linalg.cholesky(rbf_kernel(Z, Z, sf, ell)).inverse()
    return torch.mm(Linv, U0)
