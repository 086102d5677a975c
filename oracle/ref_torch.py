"""Oracle (test infrastructure): torch float64 port of the reference's EXECUTION MODEL, used as the CPU baseline.

The numpy oracle in this package is vectorised and would flatter the CPU; the reference itself runs one chain at a
time through ~100 tiny ATen ops per RK step with the autograd tape on.  This module restates that
path in the same style -- tuple-of-tensors state, Python loop over the grid, 3/8-rule stages, KernelRegression with
the [N,m]@[m,m]@[m,D] product on every RHS call, loss_closure, loss.backward(), SGLD step with host RNG -- so that its
wall-clock is what the reference's would be on the same cores.  bench.py times it ("kind": "port").
  solver loop   torchdiffeq/_impl/solvers.py:79-99, fixed_grid.py:26-33, rk_common.py:72-78
  field         scripts/vanderpol/gp.py:41-71
  closure       scripts/vanderpol/gp.py:342-353
  SGLD step     samplers/langevin.py:173-202
"""
import time

import numpy as np
import torch


def _sq_dist(X1, X2, ell):
    X1 = X1 / ell
    X1s = torch.sum(X1 ** 2, dim=1).view([-1, 1])
    X2 = X2 / ell
    X2s = torch.sum(X2 ** 2, dim=1).view([1, -1])
    return -2 * torch.mm(X1, torch.t(X2)) + X1s + X2s


def _K(X1, X2, sf, ell):
    return sf ** 2 * torch.exp(-_sq_dist(X1, X2, ell) / 2)


class KReg(torch.nn.Module):
    def __init__(self, U0, Zt, sf, ell, noise):
        super().__init__()
        self.U = torch.nn.Parameter(U0.clone(), requires_grad=True)
        self.logsn = torch.nn.Parameter(torch.zeros(2, dtype=U0.dtype) + np.log(noise), requires_grad=True)
        self.sf, self.ell, self.Z = sf, ell, Zt
        self.Kzz = _K(Zt, Zt, sf, ell)
        self.Kzzinv = self.Kzz.inverse()
        self.L = torch.linalg.cholesky(self.Kzz)
        self.KzzinvL = torch.mm(self.Kzzinv, self.L)

    def forward(self, t, X):
        T = torch.mm(_K(X, self.Z, self.sf, self.ell), self.KzzinvL)
        return torch.mm(T, self.U)


def _rk4_alt(func, t, dt, y):
    k1 = func(t, y)
    k2 = func(t + dt / 3, tuple(y_ + dt * k1_ / 3 for y_, k1_ in zip(y, k1)))
    k3 = func(t + dt * 2 / 3, tuple(y_ + dt * (k1_ / -3 + k2_) for y_, k1_, k2_ in zip(y, k1, k2)))
    k4 = func(t + dt, tuple(y_ + dt * (k1_ - k2_ + k3_) for y_, k1_, k2_, k3_ in zip(y, k1, k2, k3)))
    return tuple((k1_ + 3 * k2_ + 3 * k3_ + k4_) * (dt / 8) for k1_, k2_, k3_, k4_ in zip(k1, k2, k3, k4))


def odeint_rk4(func, y0, t):
    f = lambda t_, y_: (func(t_, y_[0]),)
    y = (y0,)
    t = t.type_as(y0)
    sol = [y]
    for t0, t1 in zip(t[:-1], t[1:]):
        dy = _rk4_alt(f, t0, t1 - t0, y)
        y = tuple(y_ + dy_ for y_, dy_ in zip(y, dy))
        sol.append(y)
    return torch.stack([s[0] for s in sol])


def make_chain(data, M=5, sf=1.0, ell=0.75, seed=0):
    from . import npde
    Z = npde.inducing_grid(data["Y"], M)
    U0 = npde.gradient_matching_init(data["Y"], np.asarray(data["t"], dtype=np.float64), Z, sf, ell)
    rng = np.random.default_rng(seed)
    U0 = U0 + 0.1 * rng.standard_normal(U0.shape)
    kreg = KReg(torch.from_numpy(U0), torch.from_numpy(Z), sf, ell, 0.1)
    x0 = torch.from_numpy(np.asarray(data["x0"]))
    t = torch.from_numpy(np.asarray(data["t"]))
    Yt = torch.from_numpy(np.asarray(data["Y"]))

    def closure():
        xode = odeint_rk4(kreg, x0, t).permute([1, 0, 2])
        loss = torch.sum((Yt - xode) ** 2 / (2 * torch.exp(kreg.logsn) ** 2))
        loss = loss + torch.numel(Yt) * torch.sum(kreg.logsn) / 2
        loss = loss + torch.sum(torch.diag(torch.mm(kreg.U.t(), torch.mm(kreg.Kzzinv, kreg.U)))) / 2
        return loss
    return kreg, closure


def sgld_step(params, lr):
    for p in params:
        noise = torch.distributions.Normal(torch.zeros(p.shape, dtype=p.dtype),
                                           torch.ones(p.shape, dtype=p.dtype) / np.sqrt(0.5 * lr)).sample()
        p.data.add_(p.grad.data + noise, alpha=-lr)


def time_chain(args):
    """Worker: one chain, one thread.  Returns seconds per sampler step (zero_grad -> closure -> backward -> step)."""
    seed, T, M, warm, steps = args
    torch.set_num_threads(1)
    from . import npde
    data = npde.make_vdp_data(seed=0, T=T)
    kreg, closure = make_chain(data, M=M, seed=seed)
    params = [kreg.U, kreg.logsn]
    ts = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        for p in params:
            p.grad = None
        loss = closure()
        loss.backward()
        sgld_step(params, 1e-5)
        if i >= warm:
            ts.append(time.perf_counter() - t0)
    return float(np.mean(ts))


def phi_numpy_seconds(n=4096, d=52, reps=1):
    """RBFKernel + phi (stein.py:18-34, 75-86) for n particles with numpy/BLAS on all host threads."""
    from . import samplers
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d)) * 0.1
    S = rng.standard_normal((n, d))
    t0 = time.perf_counter()
    for _ in range(reps):
        sq = (X * X).sum(1)
        d2 = np.maximum(sq[:, None] + sq[None, :] - 2 * X @ X.T, 0)          # cdist's matmul path
        gamma = samplers.median_bandwidth(d2, n)
        K = np.exp(-gamma * d2)
        phi = (K @ S + 2 * gamma * (K.sum(1)[:, None] * X - K @ X)) / n
    return (time.perf_counter() - t0) / reps, float(phi[0, 0])
