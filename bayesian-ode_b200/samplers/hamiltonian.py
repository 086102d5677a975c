"""Adaptive SGHMC (reference samplers/hamiltonian.py:11-164; Springenberg et al. 2016) as one fused update launch."""
import torch

from .. import _lib
from .langevin import _CyclicalSchedule, _flat_noise
from .sampler import Sampler


class aSGHMC(Sampler):
    """hamiltonian.py:11-164.  aSGHMC(params, lr=1e-5, mom_decay=5e-2, lambda_=1e-5, add_noise=True)."""

    def __init__(self, params, **kwargs):
        defaults = kwargs
        defaults.setdefault("add_noise", True)
        defaults.setdefault("lr", 1e-5)
        defaults.setdefault("mom_decay", 5e-2)
        defaults.setdefault("lambda_", 1e-5)
        super().__init__(params, defaults)
        self.loss = None
        self._st = {}

    def _state_for(self, p):
        key = (p.data_ptr(), p.numel())
        st = self._st.get(key)
        if st is None:                                       # hamiltonian.py:55-60
            st = {"iteration": 0, "tau": torch.ones_like(p), "g": torch.ones_like(p), "v_hat": torch.ones_like(p),
                  "momentum": torch.zeros_like(p)}
            self._st[key] = st
        return st

    def step(self, lr, burn_in=False, resample_mom_every=50, noise=None, noise_resample=None, use_ctl=False, _add_noise=None):
        lib = _lib.load()
        group = self.param_groups[0]
        add_noise = bool(group["add_noise"]) if _add_noise is None else bool(_add_noise)
        for k, (p, g) in enumerate(self._tensors_for_launch()):
            st = self._state_for(p)
            st["iteration"] += 1
            # hamiltonian.py:81-83 -- integer selection logic stays on the host, bit-exact
            resample = (not burn_in) and resample_mom_every is not None and st["iteration"] % resample_mom_every == 0
            xi = _flat_noise(noise, self, k, p) if add_noise else None
            xr = _flat_noise(noise_resample, self, k, p) if resample else None
            ctl = self.ctl() if use_ctl else None
            _lib.check(lib.bode_asghmc_step(
                _lib.ptr(p), _lib.ptr(g), _lib.ptr(st["tau"]), _lib.ptr(st["g"]), _lib.ptr(st["v_hat"]), _lib.ptr(st["momentum"]),
                _lib.ptr(xi), _lib.ptr(xr), p.numel(), float(lr), float(group["mom_decay"]), float(group["lambda_"]),
                int(bool(burn_in)), int(bool(resample)), int(add_noise), self.seed + k, self._step_index,
                _lib.ptr(self._status), _lib.ptr(ctl), _lib.stream_ptr()))
        self._after_step()

    def get_lr(self):
        if "iter" not in self.__dict__:
            self.iter = 0
        self.iter += 1
        return self.param_groups[0]["lr"]

    def sample(self, closure, num_samples=1000, burn_in=100, print_iters=True, print_loss=False, lr_scheduler=None,
               resample_mom_every=None, arr_closure=None, thinning=1):
        """hamiltonian.py:107-164."""
        chain = self.samples
        if lr_scheduler is None:
            lr_scheduler = self.get_lr
        fused = hasattr(closure, "loss_and_grad_") and self._flat is not None
        if fused:
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
                self._backward(self.loss)
            self.step(lr=lr_scheduler(), burn_in=i < burn_in, resample_mom_every=resample_mom_every)
            if i >= burn_in and (i - burn_in) % thinning == 0:
                self._record(chain)
            if log:
                sq_err_loss = closure(add_prior=False)
                if arr_closure is not None:
                    arr_closure(self.loss, sq_err_loss)
                if print_iters:
                    print("{} iter {:04d} | loss {:.06f}".format("Burn-in" if i < burn_in else "Sample",
                                                                 i + 1 if i < burn_in else i - burn_in + 1,
                                                                 float(sq_err_loss.sum())))
        return chain


class acSGHMC(_CyclicalSchedule, aSGHMC):
    """hamiltonian.py:167-326: adaptive SGHMC on the cosine cycle schedule.  acSGHMC(params, lr0=0.01, M=5, beta=0.25,
    mom_decay=5e-2, lambda_=1e-5, add_noise=True): the aSGHMC update (same fused launch, stale ``tau_inv`` included) with
    the momentum noise switched by ``r(iter_num) > beta and add_noise`` (hamiltonian.py:250-254)."""

    def __init__(self, params, **kwargs):
        kwargs.setdefault("lr0", 0.01)
        kwargs.setdefault("M", 5)
        kwargs.setdefault("beta", 0.25)
        super().__init__(params, **kwargs)
        self.param_groups[0].pop("lr", None)            # the reference's acSGHMC has no fixed ``lr`` default

    def step(self, lr, iter_num, burn_in=False, resample_mom_every=50, noise=None, noise_resample=None):
        noisy = self._sampling_phase(iter_num) and bool(self.param_groups[0]["add_noise"])
        aSGHMC.step(self, lr, burn_in=burn_in, resample_mom_every=resample_mom_every, noise=noise,
                    noise_resample=noise_resample, _add_noise=noisy)

    def get_lr(self, t):
        return _CyclicalSchedule.get_lr(self, t)

    def sample(self, closure, num_samples=1000, burn_in=100, print_iters=True, print_loss=False, resample_mom_every=None,
               arr_closure=None):
        return self._cyclical_sample(closure, num_samples, burn_in, print_iters, print_loss, arr_closure,
                                     lambda i: self.step(lr=self.get_lr(i), iter_num=i, burn_in=i < burn_in,
                                                         resample_mom_every=resample_mom_every))
