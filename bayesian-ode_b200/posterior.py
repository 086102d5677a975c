"""The posterior closure of the npde model as ONE fused launch per sampler step.

``NPDEPosterior`` is the particle-batched counterpart of ``loss_closure`` in
scripts/vanderpol/gp.py:342-353:

    loss = sum (Y - x)^2 / (2 exp(logsn)^2) + numel(Y) * sum(logsn) / D + tr(U^T Kzz^-1 U) / 2
    closure(add_prior=False) = sum (Y - x)^2

It is a callable with the closure protocol the reference samplers expect (``closure()`` returns a loss
whose ``backward()`` fills ``.grad``; ``closure(add_prior=False)`` returns the squared error), and it
also exposes ``loss_and_grad_()`` -- the graph-capturable fast path the batched samplers call.
"""
import torch

from . import _grid, _lib
from .fields import MLPField, NPDEField
import sys as _sys

from .odeint import _grid_struct, _norm_y0, _scratch, odeint

_om = _sys.modules[__package__ + ".odeint"]      # the module (the package attribute `odeint` is the function)


class _FusedNLP(torch.autograd.Function):
    """loss[P] with precomputed gradients; backward scales them by the incoming per-particle cotangent."""

    @staticmethod
    def forward(ctx, U, logsn, post):
        loss, gU, gl = post._launch(U.detach(), logsn.detach())
        ctx.save_for_backward(gU, gl)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        gU, gl = ctx.saved_tensors
        return gU * gloss.view(-1, 1, 1), gl * gloss.view(-1, 1), None


class NPDEPosterior:
    def __init__(self, field, x0, t, Y, method="rk4", options=None, grad_mode="discrete", scale=1.0, rtol=1e-7, atol=1e-9):
        """field: NPDEField; x0 [N,2] or [P,N,2]; t [T]; Y [N,T,2] (gp.py:320); method/options as odeint.
        grad_mode "discrete" (autograd-through-odeint semantics) or "adjoint" (odeint_adjoint, what gp.py:26
        runs).  ``scale`` multiplies loss and gradients (pSGLD passes 1/N, langevin.py:528)."""
        if not isinstance(field, NPDEField):
            raise TypeError("NPDEPosterior needs an NPDEField")
        if method not in _lib.METHODS and method != "dopri5":
            raise NotImplementedError("NPDEPosterior is built for euler/midpoint/rk4 and dopri5")
        self.field = field
        self.method = method
        self.options = dict(options or {})
        self.grad_mode = {"discrete": _lib.GRAD_DISCRETE, "adjoint": _lib.GRAD_ADJOINT}[grad_mode]
        self.scale = float(scale)
        self.rtol, self.atol = rtol, atol
        dev = field.U.device
        self.x0, self.x0_batched, self.N = _norm_y0(field, torch.as_tensor(x0))
        self.t = torch.as_tensor(t)
        self.Y = torch.as_tensor(Y).to(dev, torch.float32).contiguous()
        if self.Y.shape != (self.N, self.t.numel(), 2):
            raise ValueError("Y must be [N, T, 2]")
        if method == "dopri5":
            self.grid = None
            self.options.setdefault("controller", "batch")           # gp.py:346: ONE odeint call with y0 [N, 2]
            self._d5 = _om.dopri5_setup(field, self.x0, self.t, rtol, atol, self.options)
        else:
            opts = _grid.split_options("NPDEPosterior", self.options)
            self.grid = _grid.cached(self.t, torch.float32, dev, opts["step_size"], opts["grid_constructor"],
                                     with_adjoint=self.grad_mode == _lib.GRAD_ADJOINT, func=field, y0=(self.x0,))
        P = field.P
        self.loss = torch.empty(P, dtype=torch.float32, device=dev)
        self.sqerr = torch.empty(P, dtype=torch.float32, device=dev)
        # gradients land directly in the column blocks of the field's flat grad buffer
        self.gtheta = field.theta_grad
        self.gU = self.gtheta[:, :2 * field.m].view(P, field.m, 2)
        self.glogsn = self.gtheta[:, 2 * field.m:]
        self.add_prior = True
        self.check_status = True      # dopri5: raise on solver status flags right after the launch (costs a sync)

    # ------------------------------------------------------------------------------------------------
    def set_data(self, x0=None, Y=None):
        """Overwrite the resident observations in place (a new minibatch; keeps captured graphs valid)."""
        if x0 is not None:
            self.x0.copy_(x0, non_blocking=True)
        if Y is not None:
            self.Y.copy_(Y, non_blocking=True)

    def _launch(self, U, logsn, out=None):
        lib = _lib.load()
        f = self.field
        if out is not None:
            loss, sqerr, gU, gl = out
        else:
            gflat = torch.empty_like(self.gtheta)
            loss, sqerr = torch.empty_like(self.loss), self.sqerr
            gU, gl = gflat[:, :2 * f.m].view(f.P, f.m, 2), gflat[:, 2 * f.m:]
        fs = f.c_struct(U)
        lp, ls = _lib.rows(logsn, 2)
        gUp, gUs = _lib.rows(gU, 2 * f.m)
        glp, gls = _lib.rows(gl, 2)
        if self.method == "dopri5":
            c = self._d5
            nsc = lib.bode_dopri5_scratch_floats(f.P, self.N, c["T"], _om.DOPRI5_MAX_REC_STEPS)
            sc = _scratch(U.device, nsc)
            _lib.check(lib.bode_npde_dopri5_nlp_grad(
                fs, c["o"], c["T"], c["sign"], self.N, _lib.ptr(self.x0), int(self.x0_batched), _lib.ptr(self.Y), lp, ls, self.scale,
                int(self.add_prior), _lib.ptr(loss), _lib.ptr(sqerr), gUp, gUs, glp, gls, _lib.ptr(sc), sc.numel(),
                _om.DOPRI5_MAX_REC_STEPS, _lib.stream_ptr()))
            _om.dopri5_check(c["stats"], sync=self.check_status)
            return loss, gU, gl
        nsc = lib.bode_npde_scratch_floats_m(f.P, self.N, self.grid.S, self.grid.T, _lib.METHODS[self.method], self.grad_mode, f.m)
        sc = _scratch(U.device, nsc)
        gs = _grid_struct(self.grid, self.grad_mode == _lib.GRAD_ADJOINT)
        _lib.check(lib.bode_npde_nlp_grad(
            fs, gs, _lib.METHODS[self.method], self.grad_mode, self.N, _lib.ptr(self.x0), int(self.x0_batched),
            _lib.ptr(self.Y), lp, ls, self.scale, int(self.add_prior),
            _lib.ptr(loss), _lib.ptr(sqerr), gUp, gUs, glp, gls, _lib.ptr(sc), sc.numel(), _lib.stream_ptr()))
        return loss, gU, gl

    def loss_and_grad_(self):
        """One fused launch: fills self.loss/self.sqerr/self.gU/self.glogsn in place (no autograd graph,
        no allocation -> CUDA-graph capturable).  Returns (loss, gU, glogsn)."""
        return self._launch(self.field.U.data, self.field.logsn.data, out=(self.loss, self.sqerr, self.gU, self.glogsn))

    def __call__(self, add_prior=True):
        f = self.field
        if not add_prior:
            with torch.no_grad():
                sol = odeint(f, self.x0, self.t, rtol=self.rtol, atol=self.atol, method=self.method, options=self.options or None)
                sol = sol if f.batched else sol[:, None]
                r2 = (self.Y[None] - sol.permute(1, 2, 0, 3)) ** 2
                out = r2.sum(dim=(1, 2, 3))
            return out if f.batched else out[0]
        loss = _FusedNLP.apply(f.U, f.logsn, self)
        return loss if f.batched else loss[0]


class _FusedSSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, post, *params):
        loss, gth = post._launch()
        ctx.save_for_backward(gth)
        ctx.post = post
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (gth,) = ctx.saved_tensors
        f = ctx.post.field
        g = gth * gloss.view(-1, 1)
        return (None,) + tuple(g[:, o:o + n].view((f.P,) + shp) for (o, n, shp) in f._blocks().values())


class MLPPosterior:
    """``bayesian_closure`` of notebooks/jai/nn.ipynb cell 10 for P chains as one fused launch:
    loss_p = lik_w * sum_rows sum (X[row] - odeint(net_p, x0[row], t))^2 + reg * sum theta_p^2   (reg = 0.5 in the notebook);
    ``closure(add_prior=False)`` returns the squared error."""

    def __init__(self, field, x0, t, X, method="rk4", options=None, grad_mode="discrete", reg=0.5, lik_w=1.0, scale=1.0,
                 rtol=1e-7, atol=1e-9):
        if not isinstance(field, MLPField):
            raise TypeError("MLPPosterior needs an MLPField")
        if method not in _lib.METHODS and method != "dopri5":
            raise NotImplementedError("MLPPosterior is built for euler/midpoint/rk4 and dopri5")
        self.rtol, self.atol = rtol, atol
        self.check_status = True
        self.field, self.method, self.options = field, method, dict(options or {})
        self.grad_mode = {"discrete": _lib.GRAD_DISCRETE, "adjoint": _lib.GRAD_ADJOINT}[grad_mode]
        self.reg, self.lik_w, self.scale = float(reg), float(lik_w), float(scale)
        dev = field.theta.device
        self.x0, self.x0_batched, self.N = _norm_y0(field, torch.as_tensor(x0))
        self.t = torch.as_tensor(t)
        self.Y = torch.as_tensor(X).to(dev, torch.float32).contiguous()
        if self.Y.shape != (self.N, self.t.numel(), 2):
            raise ValueError("X must be [N, T, 2]")
        if method == "dopri5":
            self.grid = None
            self.options.setdefault("controller", "pair")            # nn.ipynb cell 10: one odeint call per trajectory row
            self._d5 = _om.dopri5_setup(field, self.x0, self.t, rtol, atol, self.options)
        else:
            opts = _grid.split_options("MLPPosterior", self.options)
            self.grid = _grid.cached(self.t, torch.float32, dev, opts["step_size"], opts["grid_constructor"],
                                     with_adjoint=self.grad_mode == _lib.GRAD_ADJOINT, func=field, y0=(self.x0,))
        self.loss = torch.empty(field.P, dtype=torch.float32, device=dev)
        self.sqerr = torch.empty(field.P, dtype=torch.float32, device=dev)
        self.gtheta = field.theta_grad
        self.add_prior = True

    def set_data(self, x0=None, Y=None):
        if x0 is not None:
            self.x0.copy_(x0, non_blocking=True)
        if Y is not None:
            self.Y.copy_(Y, non_blocking=True)

    def _launch(self, out=None):
        lib = _lib.load()
        f = self.field
        loss, gth = out if out is not None else (torch.empty_like(self.loss), torch.empty_like(self.gtheta))
        if self.method == "dopri5":
            c = self._d5
            nsc = lib.bode_dopri5_scratch_floats(f.P, self.N, c["T"], _om.DOPRI5_MAX_REC_STEPS)
            sc = _scratch(f.theta.device, nsc)
            _lib.check(lib.bode_mlp_dopri5_sse_grad(f.c_struct(), c["o"], c["T"], c["sign"], self.N, _lib.ptr(self.x0), int(self.x0_batched),
                                                    _lib.ptr(self.Y), self.lik_w, self.reg, self.scale, int(self.add_prior), _lib.ptr(loss),
                                                    _lib.ptr(self.sqerr), _lib.ptr(gth), f.d, _lib.ptr(sc), sc.numel(),
                                                    _om.DOPRI5_MAX_REC_STEPS, _lib.stream_ptr()))
            _om.dopri5_check(c["stats"], sync=self.check_status)
            return loss, gth
        m = _lib.METHODS[self.method]
        nsc = lib.bode_npde_scratch_floats(f.P, self.N, self.grid.S, self.grid.T, m, self.grad_mode)
        sc = _scratch(f.theta.device, nsc)
        gs = _grid_struct(self.grid, self.grad_mode == _lib.GRAD_ADJOINT)
        _lib.check(lib.bode_mlp_sse_grad(f.c_struct(), gs, m, self.grad_mode, self.N, _lib.ptr(self.x0), int(self.x0_batched),
                                         _lib.ptr(self.Y), self.lik_w, self.reg, self.scale, int(self.add_prior), _lib.ptr(loss),
                                         _lib.ptr(self.sqerr), _lib.ptr(gth), f.d, _lib.ptr(sc), sc.numel(), _lib.stream_ptr()))
        return loss, gth

    def loss_and_grad_(self):
        loss, g = self._launch(out=(self.loss, self.gtheta))
        return loss, g, None

    def __call__(self, add_prior=True):
        f = self.field
        if not add_prior:
            with torch.no_grad():
                sol = odeint(f, self.x0, self.t, rtol=self.rtol, atol=self.atol, method=self.method, options=self.options or None)
                return ((self.Y[None] - sol.permute(1, 2, 0, 3)) ** 2).sum(dim=(1, 2, 3))
        return _FusedSSE.apply(self, *[getattr(f, k) for k in f._blocks()])
