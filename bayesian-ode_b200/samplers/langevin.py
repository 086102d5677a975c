"""Langevin-family samplers (reference samplers/langevin.py) as fused CUDA updates over particle-batched chains.

SGLD   langevin.py:151-258      pSGLD  langevin.py:422-567
Constructor kwargs, ``step`` / ``get_lr`` / ``sample`` signatures and the lr schedule follow the reference; the extra
``noise=`` argument of ``step`` injects standard-normal draws (one tensor per parameter, param_groups order, or one
flat [P, d] tensor) for bit-parity runs.
"""
import numpy as np
import torch

from .. import _lib
from .sampler import Sampler


def _flat_noise(noise, sampler, launch_index, like):
    """Injected xi for the launch covering ``like``: a list (one tensor per parameter) or a single tensor."""
    if noise is None:
        return None
    if torch.is_tensor(noise):
        xi = noise
    else:
        if sampler._grad_flat() is not None:
            xi = torch.cat([n.reshape(n.shape[0], -1) if n.dim() > 1 else n.reshape(1, -1) for n in noise], dim=1)
        else:
            xi = noise[launch_index]
    xi = xi.to(device=like.device, dtype=torch.float32).contiguous()
    if xi.numel() != like.numel():
        raise ValueError("injected noise has %d elements, parameter has %d" % (xi.numel(), like.numel()))
    return xi


class _LangevinBase(Sampler):
    def get_lr(self, t):
        """langevin.py:205-210: lr0 / (t0 + alpha t)^gamma, t = global iteration index incl. burn-in."""
        g = self.param_groups[0]
        return g["lr0"] / np.power(g["lr_t0"] + g["lr_alpha"] * t, g["lr_gamma"])

    def _loss_scale(self):
        return 1.0

    def sample(self, closure, num_samples=1000, burn_in=100, print_iters=False, print_loss=False, arr_closure=None,
               clipping=False, thinning=1):
        """langevin.py:213-258 / 510-567.  ``closure()`` returns the (per-chain) negative log posterior and accepts
        ``add_prior=False``.  When the closure exposes ``loss_and_grad_`` (NPDEPosterior) each iteration is one fused
        solve+gradient launch plus one fused update launch, and the chain stays on the device."""
        chain = self.samples
        fused = hasattr(closure, "loss_and_grad_") and self._flat is not None
        scale = self._loss_scale()
        if fused:
            closure.scale = scale
            if self._grad_flat() is None and hasattr(closure.field, "bind_flat_grads"):
                closure.field.bind_flat_grads()
            chain.reserve((num_samples + thinning - 1) // thinning, self._flat, self._plist)
        log = arr_closure is not None or (print_iters and print_loss)
        for i in range(burn_in + num_samples):
            if fused:
                self.loss = closure.loss_and_grad_()[0]
            else:
                self.zero_grad()
                self.loss = closure()
                if scale != 1.0:
                    self.loss = self.loss * scale          # langevin.py:528
                self._backward(self.loss)
            self.step(lr=self.get_lr(i))
            if i >= burn_in and (i - burn_in) % thinning == 0:
                self._record(chain)
            if log:
                sq_err_loss = closure(add_prior=False)      # langevin.py:226 (logging-only second solve)
                if arr_closure is not None:
                    arr_closure(self.loss, sq_err_loss)
                if print_iters:
                    tag = "Burn-in" if i < burn_in else "Sample"
                    k = i + 1 if i < burn_in else i - burn_in + 1
                    print("{} iter {:04d} | loss {:.06f}".format(tag, k, float(sq_err_loss.sum())))
            elif print_iters:
                print(("Burn-in iter {:04d}" if i < burn_in else "Sample iter {:04d}").format(
                    i + 1 if i < burn_in else i - burn_in + 1))
        return chain


class SGLD(_LangevinBase):
    """langevin.py:151-258.  SGLD(params, lr0=, lr_gamma=, lr_t0=, lr_alpha=, add_noise=True)."""

    def __init__(self, params, **kwargs):
        defaults = kwargs
        if "add_noise" not in defaults:
            defaults["add_noise"] = True
        super().__init__(params, defaults)
        self.logp = None

    def step(self, lr=None, noise=None, use_ctl=False):
        lib = _lib.load()
        for group in self.param_groups:
            if lr:
                group["lr"] = lr
        group = self.param_groups[0]
        ctl = self.ctl() if use_ctl else None
        for k, (p, g) in enumerate(self._tensors_for_launch()):
            xi = _flat_noise(noise, self, k, p)
            _lib.check(lib.bode_sgld_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(xi), p.numel(), float(group["lr"]),
                                          int(bool(group["add_noise"])), self.seed + k, self._step_index,
                                          _lib.ptr(self._status), _lib.ptr(ctl), _lib.stream_ptr()))
        self._after_step()


class pSGLD(_LangevinBase):
    """langevin.py:422-567.  pSGLD(params, lr0=, ..., alpha=0.99, lambda_=1e-5, N=1): RMSprop-preconditioned SGLD
    (no Gamma term).  ``sample`` divides the loss by N (langevin.py:528)."""

    def __init__(self, params, **kwargs):
        defaults = kwargs
        defaults.setdefault("add_noise", True)
        defaults.setdefault("lr", 1e-5)
        defaults.setdefault("alpha", 0.99)
        defaults.setdefault("lambda_", 1e-5)
        defaults.setdefault("N", 1)
        super().__init__(params, defaults)
        self.logp = None
        self._V = {}

    def _loss_scale(self):
        return 1.0 / float(self.param_groups[0]["N"])

    def step(self, lr=None, clipping=False, noise=None, use_ctl=False):
        lib = _lib.load()
        for group in self.param_groups:
            if lr:
                group["lr"] = lr
        group = self.param_groups[0]
        ctl = self.ctl() if use_ctl else None
        for k, (p, g) in enumerate(self._tensors_for_launch()):
            key = (p.data_ptr(), p.numel())
            if key not in self._V:
                self._V[key] = torch.zeros_like(p)          # langevin.py:474-475
            xi = _flat_noise(noise, self, k, p)
            _lib.check(lib.bode_psgld_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(self._V[key]), _lib.ptr(xi), p.numel(),
                                           float(group["lr"]), float(group["alpha"]), float(group["lambda_"]),
                                           int(bool(group["add_noise"])), self.seed + k, self._step_index,
                                           _lib.ptr(self._status), _lib.ptr(ctl), _lib.stream_ptr()))
        self._after_step()
