"""``odeint`` / ``odeint_adjoint`` with the reference's signatures, backed by the fused CUDA kernels.

Replaces torchdiffeq/_impl/odeint.py:20-76 (``odeint``) and torchdiffeq/_impl/adjoint.py:105-133
(``odeint_adjoint``) for field modules the library recognises (``NPDEField``; ``MLPField``).
There is deliberately no generic-callable path and no CPU path: an unrecognised ``func`` raises.

Gradients: ``odeint`` back-propagates the exact discrete adjoint of the solver (what autograd
through the reference's ``odeint`` yields); ``odeint_adjoint`` reproduces the reference's continuous
adjoint re-solved per observation interval with the same method (adjoint.py:57-95).
"""
import torch

from . import _grid, _lib
from .fields import MLPField, NPDEField

FIXED_GRID_METHODS = ("euler", "midpoint", "rk4")
# methods of the reference registry (odeint.py:8-17) that are outside this hot path
_UNBUILT = ("explicit_adams", "fixed_adams", "adams", "tsit5")
SOLVERS = {"euler": "Euler", "midpoint": "Midpoint", "rk4": "RK4", "dopri5": "Dopri5Solver"}

_SCRATCH = {}


def _scratch(device, nfloats):
    """Grow-only per-device scratch for solver checkpoints (kept resident between sampler steps)."""
    buf = _SCRATCH.get(device)
    if buf is None or buf.numel() < nfloats:
        buf = torch.empty(max(int(nfloats), 1), dtype=torch.float32, device=device)
        _SCRATCH[device] = buf
    return buf


def _grid_struct(g, with_adjoint):
    gs = _lib.GridStruct()
    gs.S, gs.T, gs.sign = g.S, g.T, g.sign
    gs.dt = g.dt_dev.data_ptr()
    gs.obs_ptr = g.obs_ptr_dev.data_ptr()
    if with_adjoint and g.adj_dt_dev is not None:
        gs.adj_dt = g.adj_dt_dev.data_ptr()
        gs.adj_ptr = g.adj_ptr_dev.data_ptr()
    return gs


def _check_inputs(func, y0, t):
    """misc.py:173-195 (tuple-isation and dtype checks; time reversal lives in _grid)."""
    tensor_input = False
    if torch.is_tensor(y0):
        tensor_input = True
        y0 = (y0,)
    assert isinstance(y0, tuple), "y0 must be either a torch.Tensor or a tuple"
    for y0_ in y0:
        assert torch.is_tensor(y0_), "each element must be a torch.Tensor but received {}".format(type(y0_))
    for y0_ in y0:
        if not torch.is_floating_point(y0_):
            raise TypeError("`y0` must be a floating point Tensor but is a {}".format(y0_.type()))
    if not torch.is_floating_point(t):
        raise TypeError("`t` must be a floating point Tensor but is a {}".format(t.type()))
    if len(y0) != 1:
        raise NotImplementedError("the fused fields integrate a single state tensor; tuple states of length > 1 are not built")
    return tensor_input, y0[0]


def _norm_y0(field, y0):
    """y0 [N,2] (shared by all particles) or [P,N,2] -> (contiguous fp32 device tensor, batched flag, N)."""
    if y0.shape[-1] != 2 or y0.dim() not in (2, 3):
        raise ValueError("y0 must be [N, 2] or [P, N, 2] for the 2-D npde field")
    if y0.dim() == 3 and y0.shape[0] != field.P:
        raise ValueError("y0 leading dimension must equal the number of particles")
    _lib.require_cuda()
    y = y0.to(device=field.theta.device, dtype=torch.float32).contiguous()
    return y, y0.dim() == 3, int(y0.shape[-2])


class _NpdeOdeint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, U, field, g, method, grad_mode, batched, N):
        lib = _lib.load()
        sol = torch.empty((g.T, field.P, N, 2), dtype=torch.float32, device=U.device)
        fs = field.c_struct(U.detach())
        gs = _grid_struct(g, False)
        _lib.check(lib.bode_npde_odeint(fs, gs, method, N, _lib.ptr(y0), int(batched), _lib.ptr(sol), _lib.stream_ptr()))
        ctx.save_for_backward(y0, U)
        ctx.misc = (field, g, method, grad_mode, batched, N)
        return sol

    @staticmethod
    def backward(ctx, gout):
        y0, U = ctx.saved_tensors
        field, g, method, grad_mode, batched, N = ctx.misc
        lib = _lib.load()
        gout = gout.to(torch.float32).contiguous()
        gU = torch.empty((field.P, field.m, 2), dtype=torch.float32, device=U.device)
        gy0 = torch.empty((field.P, N, 2), dtype=torch.float32, device=U.device)
        nsc = lib.bode_npde_scratch_floats(field.P, N, g.S, g.T, method, grad_mode)
        sc = _scratch(U.device, nsc)
        fs = field.c_struct(U.detach())
        gs = _grid_struct(g, grad_mode == _lib.GRAD_ADJOINT)
        _lib.check(lib.bode_npde_odeint_backward(fs, gs, method, grad_mode, N, _lib.ptr(y0), int(batched), _lib.ptr(gout),
                                                 _lib.ptr(gU), 2 * field.m, _lib.ptr(gy0), _lib.ptr(sc), sc.numel(),
                                                 _lib.stream_ptr()))
        if not batched:
            gy0 = gy0.sum(0)
        return gy0, gU, None, None, None, None, None, None


class _MlpOdeint(torch.autograd.Function):
    """odeint for MLPField; the six parameter views enter as inputs so autograd routes d/dtheta back to each."""

    @staticmethod
    def forward(ctx, y0, field, g, method, grad_mode, batched, N, *params):
        lib = _lib.load()
        sol = torch.empty((g.T, field.P, N, 2), dtype=torch.float32, device=y0.device)
        fs = field.c_struct()
        gs = _grid_struct(g, False)
        _lib.check(lib.bode_mlp_odeint(fs, gs, method, N, _lib.ptr(y0), int(batched), _lib.ptr(sol), _lib.stream_ptr()))
        ctx.save_for_backward(y0)
        ctx.misc = (field, g, method, grad_mode, batched, N)
        return sol

    @staticmethod
    def backward(ctx, gout):
        (y0,) = ctx.saved_tensors
        field, g, method, grad_mode, batched, N = ctx.misc
        lib = _lib.load()
        gout = gout.to(torch.float32).contiguous()
        gth = torch.empty((field.P, field.d), dtype=torch.float32, device=y0.device)
        gy0 = torch.empty((field.P, N, 2), dtype=torch.float32, device=y0.device)
        nsc = lib.bode_npde_scratch_floats(field.P, N, g.S, g.T, method, grad_mode)
        sc = _scratch(y0.device, nsc)
        fs = field.c_struct()
        gs = _grid_struct(g, grad_mode == _lib.GRAD_ADJOINT)
        _lib.check(lib.bode_mlp_odeint_backward(fs, gs, method, grad_mode, N, _lib.ptr(y0), int(batched), _lib.ptr(gout),
                                                _lib.ptr(gth), field.d, _lib.ptr(gy0), _lib.ptr(sc), sc.numel(), _lib.stream_ptr()))
        if not batched:
            gy0 = gy0.sum(0)
        grads = tuple(gth[:, o:o + n].view((field.P,) + shp) for (o, n, shp) in field._blocks().values())
        return (gy0, None, None, None, None, None, None) + grads


_last_dopri5_stats = None


def last_dopri5_stats():
    """[P, N, 3] int32 (accepted steps, rejected steps, status bits) of the most recent dopri5 call (diagnostics)."""
    return _last_dopri5_stats


DOPRI5_MAX_REC_STEPS = 256      # accepted steps recorded per (particle, trajectory) pair for the gradient


CONTROLLERS = {"pair": 0, "batch": 1}


def dopri5_setup(func, y0, t, rtol, atol, options, controller="batch"):
    """Normalise the dopri5 call (dopri5.py:60-75 options, misc.py:184-187 time reversal) into the C-ABI structures.

    ``options['controller']`` (an extension; every other unknown key warns like misc.py:79-81) picks what ONE reference call is:
      "batch" (default of ``odeint`` / ``odeint_adjoint``)  the whole ``y0 [N, 2]`` of a particle is one call -- one step-size
               controller per particle, error ratio and initial-step norms pooled over all N x 2 elements (misc.py:146-157);
      "pair"   one call per trajectory row, as nn.ipynb cell 10 integrates (``MLPPosterior``'s default)."""
    import warnings
    options = dict(options or {})
    controller = options.pop("controller", controller)
    if controller not in CONTROLLERS:
        raise ValueError("options['controller'] must be 'batch' or 'pair'")
    known = {k: options.pop(k) for k in ("first_step", "safety", "ifactor", "dfactor", "max_num_steps") if k in options}
    if len(options) > 0:
        warnings.warn("Dopri5Solver: Unexpected arguments {}".format(options))          # misc.py:79-81
    y0c, batched, N = _norm_y0(func, y0)
    dev = y0c.device
    t64 = t.detach().to("cpu", torch.float64)
    sign = 1.0
    if t64.numel() > 1 and bool((t64[1:] < t64[:-1]).all()):
        t64, sign = -t64, -1.0
    assert t64.numel() < 2 or bool((t64[1:] > t64[:-1]).all()), "t must be strictly increasing or decrasing"
    tdev = t64.to(dev)
    stats = torch.zeros((func.P, N, 3), dtype=torch.int32, device=dev)
    o = _lib.Dopri5Opts()
    o.t = tdev.data_ptr()
    o.rtol, o.atol = float(rtol), float(atol)
    o.safety, o.ifactor, o.dfactor = float(known.get("safety", 0.9)), float(known.get("ifactor", 10.0)), float(known.get("dfactor", 0.2))
    o.max_num_steps = int(min(known.get("max_num_steps", 2 ** 31 - 1), 2 ** 31 - 1))
    o.user_first_step = int(known.get("first_step") is not None)
    o.stats = stats.data_ptr()
    o.controller = CONTROLLERS[controller]
    return dict(o=o, tdev=tdev, T=int(t64.numel()), sign=sign, y0=y0c, batched=batched, N=N, stats=stats)


def dopri5_check(stats, sync=True):
    """dopri5.py:89,100-102 assert on these conditions; bit 8 = more accepted steps than the gradient record holds."""
    global _last_dopri5_stats
    _last_dopri5_stats = stats
    if not sync:
        return
    flags = 0
    for b in (1, 2, 4, 8):
        if bool(((stats[..., 2] & b) != 0).any()):
            flags |= b
    assert not (flags & 1), "max_num_steps exceeded"
    assert not (flags & 2), "underflow in dt"
    assert not (flags & 4), "non-finite values in state `y`"
    assert not (flags & 8), "more than DOPRI5_MAX_REC_STEPS accepted steps: raise bayesian_ode_b200.odeint.DOPRI5_MAX_REC_STEPS"


class _Dopri5Odeint(torch.autograd.Function):
    """Adaptive forward; backward = discrete adjoint of the accepted steps with frozen step sizes (one fused launch that
    redoes the solve, records it and sweeps back)."""

    @staticmethod
    def forward(ctx, y0c, func, cfg, *params):
        lib = _lib.load()
        T, N = cfg["T"], cfg["N"]
        sol = torch.empty((T, func.P, N, 2), dtype=torch.float32, device=y0c.device)
        if isinstance(func, NPDEField):
            st = lib.bode_npde_dopri5(func.c_struct(func.U.detach()), cfg["o"], T, cfg["sign"], N, _lib.ptr(y0c), int(cfg["batched"]),
                                      _lib.ptr(sol), _lib.stream_ptr())
        else:
            st = lib.bode_mlp_dopri5(func.c_struct(), cfg["o"], T, cfg["sign"], N, _lib.ptr(y0c), int(cfg["batched"]), _lib.ptr(sol),
                                     _lib.stream_ptr())
        _lib.check(st)
        dopri5_check(cfg["stats"])
        ctx.save_for_backward(y0c)
        ctx.misc = (func, cfg)
        return sol

    @staticmethod
    def backward(ctx, gout):
        (y0c,) = ctx.saved_tensors
        func, cfg = ctx.misc
        lib = _lib.load()
        T, N = cfg["T"], cfg["N"]
        gout = gout.to(torch.float32).contiguous()
        gy0 = torch.empty((func.P, N, 2), dtype=torch.float32, device=y0c.device)
        nsc = lib.bode_dopri5_scratch_floats(func.P, N, T, DOPRI5_MAX_REC_STEPS)
        sc = _scratch(y0c.device, nsc)
        if isinstance(func, NPDEField):
            g = torch.empty((func.P, func.m, 2), dtype=torch.float32, device=y0c.device)
            st = lib.bode_npde_dopri5_backward(func.c_struct(func.U.detach()), cfg["o"], T, cfg["sign"], N, _lib.ptr(y0c), int(cfg["batched"]),
                                               _lib.ptr(gout), _lib.ptr(g), 2 * func.m, _lib.ptr(gy0), _lib.ptr(sc), sc.numel(),
                                               DOPRI5_MAX_REC_STEPS, _lib.stream_ptr())
            grads = (g,)
        else:
            g = torch.empty((func.P, func.d), dtype=torch.float32, device=y0c.device)
            st = lib.bode_mlp_dopri5_backward(func.c_struct(), cfg["o"], T, cfg["sign"], N, _lib.ptr(y0c), int(cfg["batched"]), _lib.ptr(gout),
                                              _lib.ptr(g), func.d, _lib.ptr(gy0), _lib.ptr(sc), sc.numel(), DOPRI5_MAX_REC_STEPS,
                                              _lib.stream_ptr())
            grads = tuple(g[:, o:o + n].view((func.P,) + shp) for (o, n, shp) in func._blocks().values())
        _lib.check(st)
        dopri5_check(cfg["stats"])
        if not cfg["batched"]:
            gy0 = gy0.sum(0)
        return (gy0, None, None) + grads


def _dopri5(func, y0, t, rtol, atol, options, tensor_input):
    """Dopri5Solver (dopri5.py:58-122); controller granularity per ``options['controller']`` (see dopri5_setup)."""
    cfg = dopri5_setup(func, y0, t, rtol, atol, options)
    params = [func.U] if isinstance(func, NPDEField) else [getattr(func, k) for k in func._blocks()]
    sol = _Dopri5Odeint.apply(cfg["y0"], func, cfg, *params)
    if isinstance(func, NPDEField) and not func.batched:
        sol = sol[:, 0]
    return sol if tensor_input else (sol,)


def _odeint_impl(func, y0, t, rtol, atol, method, options, grad_mode, who):
    tensor_input, y0 = _check_inputs(func, y0, t)
    if options is None:
        options = {}
    elif method is None:
        raise ValueError("cannot supply `options` without specifying `method`")
    if method is None:
        method = "dopri5"
    if method in _UNBUILT:
        raise NotImplementedError("method '{}' of the reference registry is outside the B200 hot path".format(method))
    solver_name = SOLVERS[method]          # KeyError for unknown methods, like odeint.py:71
    if t.requires_grad:
        raise NotImplementedError("gradients with respect to `t` are not built")
    if isinstance(func, NPDEField):
        if method in FIXED_GRID_METHODS:
            opts = _grid.split_options(solver_name, options)
            if opts["step_size"] is not None and opts["grid_constructor"] is not None:
                raise ValueError("step_size and grid_constructor are exclusive arguments.")
            y0c, batched, N = _norm_y0(func, y0)
            g = _grid.cached(t, torch.float32, func.U.device, opts["step_size"], opts["grid_constructor"],
                             with_adjoint=grad_mode == _lib.GRAD_ADJOINT, func=func, y0=(y0c,))
            sol = _NpdeOdeint.apply(y0c, func.U, func, g, _lib.METHODS[method], grad_mode, batched, N)
            if not func.batched:
                sol = sol[:, 0]
            return sol if tensor_input else (sol,)
        return _dopri5(func, y0, t, rtol, atol, options, tensor_input)
    if isinstance(func, MLPField):
        if method in FIXED_GRID_METHODS:
            opts = _grid.split_options(solver_name, options)
            if opts["step_size"] is not None and opts["grid_constructor"] is not None:
                raise ValueError("step_size and grid_constructor are exclusive arguments.")
            y0c, batched, N = _norm_y0(func, y0)
            g = _grid.cached(t, torch.float32, func.theta.device, opts["step_size"], opts["grid_constructor"],
                             with_adjoint=grad_mode == _lib.GRAD_ADJOINT, func=func, y0=(y0c,))
            params = [getattr(func, k) for k in func._blocks()]
            sol = _MlpOdeint.apply(y0c, func, g, _lib.METHODS[method], grad_mode, batched, N, *params)
            return sol if tensor_input else (sol,)
        return _dopri5(func, y0, t, rtol, atol, options, tensor_input)
    raise TypeError(
        "{}: `func` must be a field module of bayesian_ode_b200 (NPDEField / MLPField); got {}. The B200 build "
        "has no generic-callable or CPU path.".format(who, type(func).__name__))


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    """Same contract as torchdiffeq.odeint (odeint.py:20-76); output is time-major ``[T, (P,) N, 2]``."""
    return _odeint_impl(func, y0, t, rtol, atol, method, options, _lib.GRAD_DISCRETE, "odeint")


def odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None):
    """Same contract as torchdiffeq.odeint_adjoint (adjoint.py:105-133)."""
    if not isinstance(func, torch.nn.Module):
        raise ValueError("func is required to be an instance of nn.Module.")
    return _odeint_impl(func, y0, t, rtol, atol, method, options, _lib.GRAD_ADJOINT, "odeint_adjoint")
