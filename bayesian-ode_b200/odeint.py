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
    return tensor_input, y0


class TupleField(torch.nn.Module):
    """``func`` for TUPLE states (misc.py:175-182; api_tests.py:19-38 ``tuple_f = lambda t, y: (f(t, y[0]), f(t, y[1]))``): applies one
    fused field to every element of the state tuple.  ``odeint(TupleField(f), (y0a, y0b), t)`` integrates the concatenated
    trajectories in ONE launch and returns a tuple; with dopri5 the single step-size controller pools the error per state tensor
    and takes the maximum over the tuple, exactly as torchdiffeq does (dopri5.py:108-109, misc.py:125-141, 161)."""

    def __init__(self, field):
        super().__init__()
        if not isinstance(field, (NPDEField, MLPField)):
            raise TypeError("TupleField wraps an NPDEField or MLPField")
        self.field = field

    def forward(self, t, y):
        return tuple(self.field(t, y_) for y_ in y)


def _field_eval(func, sol):
    """f(t_i, y_i) for every stored solution point with the field's torch evaluation (sol [T, N, 2] or [T, P, N, 2])."""
    T = sol.shape[0]
    with torch.no_grad():
        if sol.dim() == 3:                                   # single un-batched particle
            return func(None, sol.reshape(-1, 2)).reshape(sol.shape)
        X = sol.permute(1, 0, 2, 3).reshape(sol.shape[1], T * sol.shape[2], 2)
        return func(None, X).reshape(sol.shape[1], T, sol.shape[2], 2).permute(1, 0, 2, 3)


class _TimeGrad(torch.autograd.Function):
    """dL/dt of ``odeint_adjoint`` (adjoint.py:68-76, 99-100) for the autonomous fused fields: moving observation time t_i moves the
    observed state along f, dL/dt_i = f(t_i, y_i) . dL/dy_i for i >= 1, and the start time collects the negative sum (the
    ``adj_time`` component of the augmented state is only changed by these terms, since df/dt = 0)."""

    @staticmethod
    def forward(ctx, sol, t, func):
        ctx.func = func
        ctx.save_for_backward(sol.detach(), t)
        return sol.view_as(sol)

    @staticmethod
    def backward(ctx, gout):
        sol, t = ctx.saved_tensors
        f = _field_eval(ctx.func, sol)
        per_t = (f * gout.to(f.dtype)).reshape(sol.shape[0], -1).sum(1)
        gt = per_t.clone()
        gt[0] = -per_t[1:].sum()
        return gout, gt.to(device=t.device, dtype=t.dtype), None


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
        nsc = lib.bode_npde_scratch_floats_m(field.P, N, g.S, g.T, method, grad_mode, field.m)
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
        # the parameters AS THEY ARE NOW: the sampler kernels update theta through raw pointers (no autograd version bump), so a
        # backward() after an in-place update would otherwise differentiate a different solve
        theta0 = field.theta.detach().clone() if any(ctx.needs_input_grad) else None
        ctx.save_for_backward(y0)
        ctx.misc = (field, g, method, grad_mode, batched, N, theta0)
        return sol

    @staticmethod
    def backward(ctx, gout):
        (y0,) = ctx.saved_tensors
        field, g, method, grad_mode, batched, N, theta0 = ctx.misc
        lib = _lib.load()
        gout = gout.to(torch.float32).contiguous()
        gth = torch.empty((field.P, field.d), dtype=torch.float32, device=y0.device)
        gy0 = torch.empty((field.P, N, 2), dtype=torch.float32, device=y0.device)
        nsc = lib.bode_npde_scratch_floats(field.P, N, g.S, g.T, method, grad_mode)
        sc = _scratch(y0.device, nsc)
        fs = field.c_struct(theta0)
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


def dopri5_setup(func, y0, t, rtol, atol, options, controller="batch", groups=None):
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
    if groups is not None and len(groups) > 1:                  # tuple state: per-tensor error pooling, max over the tuple
        if controller != "batch":
            raise ValueError("a tuple state is one odeint call: options['controller'] must be 'batch'")
        if len(groups) > 4:
            raise NotImplementedError("tuple states of more than 4 tensors are not built")
        o.n_groups = len(groups)
        end = 0
        for g, n in enumerate(groups):
            end += int(n)
            o.group_end[g] = end
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
        # forward-time parameters for the backward re-solve (see _MlpOdeint.forward)
        p0 = None
        if any(ctx.needs_input_grad):
            p0 = (func.U if isinstance(func, NPDEField) else func.theta).detach().clone()
        ctx.misc = (func, cfg, p0)
        return sol

    @staticmethod
    def backward(ctx, gout):
        (y0c,) = ctx.saved_tensors
        func, cfg, p0 = ctx.misc
        lib = _lib.load()
        T, N = cfg["T"], cfg["N"]
        gout = gout.to(torch.float32).contiguous()
        gy0 = torch.empty((func.P, N, 2), dtype=torch.float32, device=y0c.device)
        nsc = lib.bode_dopri5_scratch_floats(func.P, N, T, DOPRI5_MAX_REC_STEPS)
        sc = _scratch(y0c.device, nsc)
        if isinstance(func, NPDEField):
            g = torch.empty((func.P, func.m, 2), dtype=torch.float32, device=y0c.device)
            st = lib.bode_npde_dopri5_backward(func.c_struct(p0), cfg["o"], T, cfg["sign"], N, _lib.ptr(y0c), int(cfg["batched"]),
                                               _lib.ptr(gout), _lib.ptr(g), 2 * func.m, _lib.ptr(gy0), _lib.ptr(sc), sc.numel(),
                                               DOPRI5_MAX_REC_STEPS, _lib.stream_ptr())
            grads = (g,)
        else:
            g = torch.empty((func.P, func.d), dtype=torch.float32, device=y0c.device)
            st = lib.bode_mlp_dopri5_backward(func.c_struct(p0), cfg["o"], T, cfg["sign"], N, _lib.ptr(y0c), int(cfg["batched"]), _lib.ptr(gout),
                                              _lib.ptr(g), func.d, _lib.ptr(gy0), _lib.ptr(sc), sc.numel(), DOPRI5_MAX_REC_STEPS,
                                              _lib.stream_ptr())
            grads = tuple(g[:, o:o + n].view((func.P,) + shp) for (o, n, shp) in func._blocks().values())
        _lib.check(st)
        dopri5_check(cfg["stats"])
        if not cfg["batched"]:
            gy0 = gy0.sum(0)
        return (gy0, None, None) + grads


def _dopri5(func, y0, t, rtol, atol, options, tensor_input, groups=None):
    """Dopri5Solver (dopri5.py:58-122); controller granularity per ``options['controller']`` (see dopri5_setup)."""
    cfg = dopri5_setup(func, y0, t, rtol, atol, options, groups=groups)
    params = [func.U] if isinstance(func, NPDEField) else [getattr(func, k) for k in func._blocks()]
    sol = _Dopri5Odeint.apply(cfg["y0"], func, cfg, *params)
    if isinstance(func, NPDEField) and not func.batched:
        sol = sol[:, 0]
    return sol if tensor_input else (sol,)


def _odeint_impl(func, y0, t, rtol, atol, method, options, grad_mode, who):
    tensor_input, y0s = _check_inputs(func, y0, t)
    if options is None:
        options = {}
    elif method is None:
        raise ValueError("cannot supply `options` without specifying `method`")
    if method is None:
        method = "dopri5"
    if method in _UNBUILT:
        raise NotImplementedError("method '{}' of the reference registry is outside the B200 hot path".format(method))
    solver_name = SOLVERS[method]          # KeyError for unknown methods, like odeint.py:71
    want_gt = bool(t.requires_grad)
    if want_gt and grad_mode != _lib.GRAD_ADJOINT:
        raise NotImplementedError("dL/dt is built for odeint_adjoint (adjoint.py:68-76); plain odeint differentiates the state and the "
                                  "parameters only -- detach `t` or use odeint_adjoint")
    t_in, t = t, t.detach()
    # ---- tuple states: one fused field applied to every element (TupleField); the elements travel as one concatenated batch
    groups = None
    field = func
    if isinstance(func, TupleField):
        field = func.field
        if tensor_input:
            raise TypeError("TupleField expects a tuple state; pass the field itself for a tensor state")
    elif len(y0s) > 1:
        raise TypeError("{}: a tuple state needs `func = TupleField(field)` (the reference's func returns a tuple there, "
                        "misc.py:175-182); got {}".format(who, type(func).__name__))
    if len(y0s) > 1:
        if any(y_.dim() != y0s[0].dim() or y_.shape[:-2] != y0s[0].shape[:-2] for y_ in y0s):
            raise ValueError("the elements of a tuple state must agree in all but the trajectory dimension")
        groups = [int(y_.shape[-2]) for y_ in y0s]
        y0 = torch.cat(list(y0s), dim=-2)
    else:
        y0 = y0s[0]

    def finish(sol):
        if want_gt:
            sol = _TimeGrad.apply(sol, t_in, field)
        if tensor_input:
            return sol
        if groups is None:
            return (sol,)
        return tuple(torch.split(sol, groups, dim=-2))

    if isinstance(field, NPDEField):
        if method in FIXED_GRID_METHODS:
            opts = _grid.split_options(solver_name, options)
            if opts["step_size"] is not None and opts["grid_constructor"] is not None:
                raise ValueError("step_size and grid_constructor are exclusive arguments.")
            y0c, batched, N = _norm_y0(field, y0)
            g = _grid.cached(t, torch.float32, field.U.device, opts["step_size"], opts["grid_constructor"],
                             with_adjoint=grad_mode == _lib.GRAD_ADJOINT, func=field, y0=(y0c,))
            sol = _NpdeOdeint.apply(y0c, field.U, field, g, _lib.METHODS[method], grad_mode, batched, N)
            if not field.batched:
                sol = sol[:, 0]
            return finish(sol)
        return finish(_dopri5(field, y0, t, rtol, atol, options, True, groups=groups))
    if isinstance(field, MLPField):
        if method in FIXED_GRID_METHODS:
            opts = _grid.split_options(solver_name, options)
            if opts["step_size"] is not None and opts["grid_constructor"] is not None:
                raise ValueError("step_size and grid_constructor are exclusive arguments.")
            y0c, batched, N = _norm_y0(field, y0)
            g = _grid.cached(t, torch.float32, field.theta.device, opts["step_size"], opts["grid_constructor"],
                             with_adjoint=grad_mode == _lib.GRAD_ADJOINT, func=field, y0=(y0c,))
            params = [getattr(field, k) for k in field._blocks()]
            sol = _MlpOdeint.apply(y0c, field, g, _lib.METHODS[method], grad_mode, batched, N, *params)
            return finish(sol)
        return finish(_dopri5(field, y0, t, rtol, atol, options, True, groups=groups))
    raise TypeError(
        "{}: `func` must be a field module of bayesian_ode_b200 (NPDEField / MLPField, or TupleField around one); got {}. The B200 "
        "build has no generic-callable or CPU path.".format(who, type(func).__name__))


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    """Same contract as torchdiffeq.odeint (odeint.py:20-76); output is time-major ``[T, (P,) N, 2]``."""
    return _odeint_impl(func, y0, t, rtol, atol, method, options, _lib.GRAD_DISCRETE, "odeint")


def odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None):
    """Same contract as torchdiffeq.odeint_adjoint (adjoint.py:105-133)."""
    if not isinstance(func, torch.nn.Module):
        raise ValueError("func is required to be an instance of nn.Module.")
    return _odeint_impl(func, y0, t, rtol, atol, method, options, _lib.GRAD_ADJOINT, "odeint_adjoint")
