"""Host-side solver-grid logic of the fixed-grid integrators.

Mirrors FixedGridODESolver (reference torchdiffeq/_impl/solvers.py:36-108) and the input
normalisation of misc.py:173-195 -- computed in the STATE dtype on the host, bit for bit what
the reference computes, and shipped to the device once per distinct (t, options):

  grid      t itself, or arange(niters)*step_size + t[0] with the last point clamped (solvers.py:55-68)
  dt[s]     grid[s+1]-grid[s] in the state dtype (solvers.py:91)
  obs_ptr   outputs emitted after step s: ``while j < len(t) and t1 >= t[j]`` (solvers.py:95); every
            emitted value is the end-of-step state (the interpolation at :96 sees y0 == y1)
  sign      -1 for decreasing t (misc.py:184-187)
  adj_*     step sizes of the per-interval reverse solves of odeint_adjoint (adjoint.py:81-84)
"""
import warnings
from collections import OrderedDict

import torch


class SolverGrid:
    __slots__ = ("S", "T", "sign", "grid", "t", "dt", "obs_ptr", "adj_dt", "adj_ptr", "dt_dev", "obs_ptr_dev",
                 "adj_dt_dev", "adj_ptr_dev", "device")


def _decreasing(t):
    return bool((t[1:] < t[:-1]).all()) if t.numel() > 1 else False


def _grid_from_step_size(t, step_size):
    start_time, end_time = t[0], t[-1]
    niters = torch.ceil((end_time - start_time) / step_size + 1).item()
    t_infer = torch.arange(0, niters).to(t) * step_size + start_time
    if t_infer[-1] > t[-1]:
        t_infer[-1] = t[-1]
    return t_infer


def _one_grid(t, dtype, step_size, grid_constructor, func, y0):
    """(t_signed, grid, sign) for one call of the reference's odeint on times ``t``."""
    sign = 1.0
    if _decreasing(t):
        t = -t
        sign = -1.0
    assert t.numel() < 2 or bool((t[1:] > t[:-1]).all()), "t must be strictly increasing or decrasing"
    t = t.to(dtype)
    if step_size is not None and grid_constructor is None:
        grid = _grid_from_step_size(t, step_size)
    elif grid_constructor is None:
        grid = t
    else:
        if step_size is not None:
            raise ValueError("step_size and grid_constructor are exclusive arguments.")
        grid = grid_constructor(func, y0, t)
    assert grid[0] == t[0] and grid[-1] == t[-1]
    return t, grid.to(dtype), sign


def _output_map(t, grid):
    S = grid.numel() - 1
    tl, gl = t.tolist(), grid.tolist()
    ptr = [0] * (S + 1)
    j = 1
    for s in range(S):
        ptr[s] = j
        t1 = gl[s + 1]
        while j < len(tl) and t1 >= tl[j]:
            j += 1
    ptr[S] = j
    assert j == len(tl)
    return ptr


def build(t, dtype, device, step_size=None, grid_constructor=None, with_adjoint=False, func=None, y0=None):
    t_cpu = t.detach().to("cpu")
    tt, grid, sign = _one_grid(t_cpu, dtype, step_size, grid_constructor, func, y0)
    g = SolverGrid()
    g.t, g.grid, g.sign = tt, grid, sign
    g.S, g.T = grid.numel() - 1, tt.numel()
    g.dt = (grid[1:] - grid[:-1]).contiguous()
    g.obs_ptr = torch.tensor(_output_map(tt, grid), dtype=torch.int32)
    g.device = device
    g.dt_dev = g.dt.to(device) if g.S > 0 else torch.zeros(1, dtype=dtype, device=device)
    g.obs_ptr_dev = g.obs_ptr.to(device)
    g.adj_dt = g.adj_ptr = g.adj_dt_dev = g.adj_ptr_dev = None
    if with_adjoint and g.T > 1:
        # adjoint.py:81-84: odeint(aug, aug_y0, tensor([t[i], t[i-1]]), method=method, options=options)
        dts, ptr = [], [0]
        for i in range(1, g.T):
            _, gi, _ = _one_grid(torch.stack([t_cpu[i], t_cpu[i - 1]]), dtype, step_size, grid_constructor, func, y0)
            dts.append(gi[1:] - gi[:-1])
            ptr.append(ptr[-1] + gi.numel() - 1)
        g.adj_dt = torch.cat(dts).contiguous()
        g.adj_ptr = torch.tensor(ptr, dtype=torch.int32)
        g.adj_dt_dev = g.adj_dt.to(device)
        g.adj_ptr_dev = g.adj_ptr.to(device)
    return g


_CACHE = OrderedDict()
_CACHE_MAX = 64


def cached(t, dtype, device, step_size=None, grid_constructor=None, with_adjoint=False, func=None, y0=None):
    """Grids are a pure function of (t contents, options); key on the tensor identity + version so a
    sampler loop that reuses ``t`` pays the host logic and the H2D copies once."""
    if grid_constructor is not None:
        return build(t, dtype, device, step_size, grid_constructor, with_adjoint, func, y0)
    key = (t.data_ptr(), t._version, tuple(t.shape), t.dtype, str(t.device), dtype, str(device),
           None if step_size is None else float(step_size), bool(with_adjoint))
    g = _CACHE.get(key)
    if g is None:
        g = build(t, dtype, device, step_size, None, with_adjoint)
        _CACHE[key] = (g, t)          # keep ``t`` alive so data_ptr cannot be recycled
        if len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)
        return g
    _CACHE.move_to_end(key)
    return g[0]


def split_options(solver_name, options, allowed=("step_size", "grid_constructor")):
    """FixedGridODESolver.__init__ (solvers.py:38-53): rtol/atol are dropped silently, anything else
    unexpected only warns (misc.py:79-81)."""
    options = dict(options or {})
    out = {k: options.pop(k, None) for k in allowed}
    options.pop("rtol", None)
    options.pop("atol", None)
    if len(options) > 0:
        warnings.warn("{}: Unexpected arguments {}".format(solver_name, options))
    return out
