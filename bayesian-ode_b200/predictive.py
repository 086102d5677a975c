"""Posterior-predictive ensembles: integrate every stored sample of a chain over a new horizon in ONE batched solve.

The reference re-integrates the samples one by one (scripts/vanderpol/gp.py:449-455: ``kreg.U.data = chain[i]...;
odeint(kreg, x0_, t_)`` in a Python loop, default method dopri5 through ``odeint_adjoint``, gp.py:26) and then takes
``np.mean`` / ``np.std`` over the chain per trajectory and component (gp.py:457-464).  Here the stored samples
[S, P, d] become the particle axis of a forward-only field and the same kernels that serve the sampler step integrate
all of them at once; mean and (population, ddof = 0) standard deviation are accumulated on the device in float64.
"""
import numpy as np
import torch

from . import _lib
from .fields import NPDEField
from .odeint import odeint_adjoint


def chain_theta(samples, m):
    """The chain as a [E, 2m (+2)] device/host tensor of flattened U (and logsn) rows, E = stored samples x chains.
    Accepts a ChainStore, a tensor [S, P, d] / [E, d] / [E, m, 2], or the reference's list format
    ``[([[U, logsn]], accepted), ...]`` (langevin.py:243-245; cyclical entries with ``None`` parameters are skipped)."""
    dev = getattr(samples, "device_tensor", None)
    if dev is not None and dev() is not None and len(getattr(samples, "_host", [])) == 0:
        th = dev()
        return th.reshape(-1, th.shape[-1])
    if torch.is_tensor(samples):
        if samples.dim() == 3 and samples.shape[-1] == 2 and samples.shape[-2] == m:
            return samples.reshape(samples.shape[0], -1)
        return samples.reshape(-1, samples.shape[-1])
    rows = []
    for entry in samples:
        params = entry[0][0]
        if params[0] is None:
            continue
        U = torch.as_tensor(np.asarray(params[0]))
        rows.append(U.reshape(-1, 2 * m) if U.dim() == 3 else U.reshape(1, 2 * m))
    if not rows:
        raise ValueError("the chain holds no parameter samples")
    return torch.cat(rows, 0)


def _member_field(field, U):
    return NPDEField(U, field.Z, field.sf, field.ell, 1.0, device=field.theta.device, stable_solve=getattr(field, "_stable", False))


def ensemble_trajectories(field, samples, x0, t, method=None, rtol=1e-6, atol=1e-12, options=None, chunk=32768):
    """Trajectories of every ensemble member from the initial values ``x0`` [N, 2] at the times ``t``: a generator of
    ``[T, E_chunk, N, 2]`` device tensors (members in chain order, ``chunk`` members per solve).  ``method=None`` is the
    reference's call (gp.py:452: adaptive dopri5 with odeint_adjoint's tolerances); pass ``method='rk4'`` for the
    solver the sampler itself used."""
    th = chain_theta(samples, field.m)
    E = th.shape[0]
    for lo in range(0, E, chunk):
        U = th[lo:lo + chunk, :2 * field.m].reshape(-1, field.m, 2)
        member = _member_field(field, U)
        with torch.no_grad():
            yield odeint_adjoint(member, x0, t, rtol=rtol, atol=atol, method=method, options=options)


def posterior_predictive(field, samples, x0, t, method=None, rtol=1e-6, atol=1e-12, options=None, chunk=32768,
                         return_trajectories=False):
    """gp.py:449-464: ``(mean, std)`` of the ensemble, each ``[N, T, 2]`` (the reference's transposed layout) as float64
    NumPy arrays; ``std`` is ``np.std`` (ddof = 0).  With ``return_trajectories`` the third value is the ensemble itself,
    ``[E, N, T, 2]`` float32 on the device."""
    s1 = s2 = None
    n = 0
    keep = []
    for sol in ensemble_trajectories(field, samples, x0, t, method, rtol, atol, options, chunk):
        x = sol.to(torch.float64)
        s1 = x.sum(1) if s1 is None else s1 + x.sum(1)
        s2 = (x * x).sum(1) if s2 is None else s2 + (x * x).sum(1)
        n += sol.shape[1]
        if return_trajectories:
            keep.append(sol.permute(1, 2, 0, 3))
    mean = s1 / n
    var = (s2 / n - mean * mean).clamp_min(0.0)
    mean = mean.permute(1, 0, 2).contiguous().cpu().numpy()
    std = var.sqrt().permute(1, 0, 2).contiguous().cpu().numpy()
    if return_trajectories:
        return mean, std, torch.cat(keep, 0)
    return mean, std
