"""Langevin-family samplers (reference samplers/langevin.py) as fused CUDA updates over particle-batched chains.

SGLD   langevin.py:151-258      pSGLD  langevin.py:422-567      MALA  langevin.py:13-149      cSGLD  langevin.py:1600-1724
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
            # with use_ctl the kernel takes the step size from the device schedule (schedule_()): no host value is needed
            _lib.check(lib.bode_sgld_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(xi), p.numel(),
                                          float(group.get("lr", 0.0) if use_ctl else group["lr"]),
                                          int(bool(group["add_noise"])), self.seed + k, self._step_index,
                                          _lib.ptr(self._status), _lib.ptr(ctl), _lib.stream_ptr()))
        self._after_step()


class _CyclicalSchedule:
    """Cosine step-size cycles shared by cSGLD (langevin.py:1659-1667) and acSGHMC (hamiltonian.py:259-267):
    ``r(t) = ((t - 1) mod L) / L`` with ``L = (num_iters + M) // M``, ``lr(t) = lr0 / 2 (cos(pi r) + 1)``; noise is
    injected only in the sampling part of a cycle, ``r > beta``.  ``num_iters`` is set by ``sample()``; the modulo is
    Python's (t = 0 gives r = (L - 1) / L), integer selection logic kept on the host."""

    def _r(self, t):
        M = self.param_groups[0]["M"]
        L = (self.num_iters + M) // M
        return ((t - 1) % L) / L

    def get_lr(self, t):
        return self.param_groups[0]["lr0"] / 2.0 * (np.cos(np.pi * self._r(t)) + 1)

    def _sampling_phase(self, t):
        return self._r(t) > self.param_groups[0]["beta"]

    def _cyclical_sample(self, closure, num_samples, burn_in, print_iters, print_loss, arr_closure, step_fn):
        """The reference's loop (langevin.py:1669-1724 / hamiltonian.py:269-326): every sampling iteration appends an
        entry; outside the sampling part of a cycle the entry is ``([[None, ...]], True)``."""
        chain = self.samples
        self.num_iters = num_samples + burn_in
        fused = hasattr(closure, "loss_and_grad_") and self._flat is not None
        if fused:
            if self._grad_flat() is None and hasattr(closure.field, "bind_flat_grads"):
                closure.field.bind_flat_grads()
            keep = sum(1 for i in range(burn_in, self.num_iters) if self._sampling_phase(i))
            chain.reserve(max(keep, 1), self._flat, self._plist)
        log = arr_closure is not None or (print_iters and print_loss)
        for i in range(self.num_iters):
            if fused:
                self.loss = closure.loss_and_grad_()[0]
            else:
                self.zero_grad()
                self.loss = closure()
                self._backward(self.loss)
            step_fn(i)
            if i >= burn_in:
                if self._sampling_phase(i):
                    self._record(chain)
                else:
                    chain.push_none(len(self._plist))
            if log:
                sq_err_loss = closure(add_prior=False)
                if arr_closure is not None:
                    arr_closure(self.loss, sq_err_loss)
            if print_iters:
                tag, k = ("Burn-in", i + 1) if i < burn_in else ("Sample", i - burn_in + 1)
                if print_loss:
                    print("{} iter {:04d} | loss {:.06f}".format(tag, k, float(sq_err_loss.sum())))
                else:
                    print("{} iter {:04d}".format(tag, k))
        return chain


class cSGLD(_CyclicalSchedule, Sampler):
    """langevin.py:1600-1724: cyclical SGLD.  cSGLD(params, lr0=0.01, M=5, beta=0.25, add_noise=True); the update is the
    fused SGLD launch with the noise term switched by ``r(iter_num) > beta`` (the reference ignores ``add_noise`` here,
    langevin.py:1648-1649, and so does this class)."""

    def __init__(self, params, **kwargs):
        defaults = kwargs
        defaults.setdefault("add_noise", True)
        defaults.setdefault("lr0", 0.01)
        defaults.setdefault("M", 5)
        defaults.setdefault("beta", 0.25)
        super().__init__(params, defaults)
        self.logp = None

    def step(self, iter_num, lr=None, noise=None):
        lib = _lib.load()
        for group in self.param_groups:
            if lr:
                group["lr"] = lr
        group = self.param_groups[0]
        noisy = self._sampling_phase(iter_num)
        for k, (p, g) in enumerate(self._tensors_for_launch()):
            xi = _flat_noise(noise, self, k, p) if noisy else None
            _lib.check(lib.bode_sgld_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(xi), p.numel(), float(group["lr"]), int(noisy),
                                          self.seed + k, self._step_index, _lib.ptr(self._status), None, _lib.stream_ptr()))
        self._after_step()

    def sample(self, closure, num_samples=1000, burn_in=100, print_iters=False, print_loss=False, arr_closure=None):
        return self._cyclical_sample(closure, num_samples, burn_in, print_iters, print_loss, arr_closure,
                                     lambda i: self.step(lr=self.get_lr(i), iter_num=i))


class MALA(Sampler):
    """langevin.py:13-149.  MALA(params, lr=, add_noise=True): Langevin proposal (the SGLD update, :27-54) followed by
    ``accept_or_reject(closure)`` (:57-95), batched over the chain axis: every chain gets its own log-ratio and uniform draw.

    ``exact=False`` (default) reproduces the reference as it runs: its saved state ``self.state[p]['data'] = p.data`` (:45) is a
    view of the parameter the proposal then updates in place, so both proposal terms of the ratio see theta_prev == theta_new
    and a rejection restores nothing (the accept flags are still recorded in the chain).  ``exact=True`` keeps a copy of the
    state before the proposal, uses the textbook ratio and restores rejected chains.
    ``step(noise=)`` / ``accept_or_reject(closure, log_u=)`` inject the draws for bit-parity runs.
    """

    def __init__(self, params, exact=False, **kwargs):
        defaults = kwargs
        if "add_noise" not in defaults:
            defaults["add_noise"] = True
        super().__init__(params, defaults)
        if self._flat is None:
            raise _lib.BodeError("MALA needs parameters laid out as column blocks of one theta[P, d] buffer")
        self.exact = bool(exact)
        self.logp = None
        self.loss = None
        P, d = self._flat.shape
        dev = self._flat.device
        self._prev = torch.empty_like(self._flat) if self.exact else None
        self._gprev = torch.empty_like(self._flat)
        self.log_alpha = torch.zeros(P, dtype=torch.float32, device=dev)
        self.accepted = torch.ones(P, dtype=torch.int32, device=dev)

    def step(self, noise=None):
        """Proposal step (langevin.py:27-54): theta <- theta - lr (g + xi / sqrt(lr / 2)); remembers state and gradient."""
        lib = _lib.load()
        group = self.param_groups[0]
        (p, g), = self._tensors_for_launch()
        self._gprev.copy_(g)                                   # state['grad'] (:46)
        if self.exact:
            self._prev.copy_(p)
        xi = _flat_noise(noise, self, 0, p)
        _lib.check(lib.bode_sgld_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(xi), p.numel(), float(group["lr"]),
                                      int(bool(group["add_noise"])), self.seed, self._step_index, _lib.ptr(self._status), None,
                                      _lib.stream_ptr()))
        self._after_step()

    def accept_or_reject(self, closure, log_u=None):
        """langevin.py:57-95.  Re-evaluates loss and gradient at the proposal, then one fused accept / reject launch.
        Returns ``(params, accepted)`` like the reference, ``accepted`` being the per-chain int32 flags (device tensor)."""
        lib = _lib.load()
        group = self.param_groups[0]
        if group["add_noise"]:
            if hasattr(closure, "loss_and_grad_"):
                new_loss = closure.loss_and_grad_()[0]
            else:
                self.zero_grad()
                new_loss = closure()
                self._backward(new_loss)
            (p, g), = self._tensors_for_launch()
            P, d = p.shape
            lp = self.loss.detach().to(torch.float32).reshape(-1).expand(P).contiguous()
            ln = new_loss.detach().to(torch.float32).reshape(-1).expand(P).contiguous()
            lu = None if log_u is None else log_u.to(device=p.device, dtype=torch.float32).reshape(-1).expand(P).contiguous()
            prev = self._prev if self.exact else None
            _lib.check(lib.bode_mala_accept(_lib.ptr(prev), p.stride(0), _lib.ptr(p), p.stride(0), _lib.ptr(self._gprev),
                                            self._gprev.stride(0), _lib.ptr(g), g.stride(0), _lib.ptr(lp), _lib.ptr(ln), _lib.ptr(lu),
                                            P, d, float(group["lr"]), int(self.exact), self.seed + 0x9E37, self._step_index,
                                            _lib.ptr(self.log_alpha), _lib.ptr(self.accepted), _lib.stream_ptr()))
        else:
            self.accepted.fill_(1)
        params = [[q.detach().clone() for q in g_["params"]] for g_ in self.param_groups]
        return params, self.accepted

    def sample(self, closure, num_samples=1000, burn_in=100, print_loss=False, print_iters=False, arr_closure=None):
        """langevin.py:98-149: loss + backward, proposal, accept / reject, record ``[params, accepted]``."""
        chain = self.samples
        fused = hasattr(closure, "loss_and_grad_")
        if fused and self._grad_flat() is None and hasattr(closure.field, "bind_flat_grads"):
            closure.field.bind_flat_grads()
        for i in range(burn_in + num_samples):
            if fused:
                self.loss = closure.loss_and_grad_()[0].clone()
            else:
                self.zero_grad()
                self.loss = closure()
                self._backward(self.loss)
            self.step()
            params, acc = self.accept_or_reject(closure)
            if i >= burn_in:
                chain.append((params, acc.clone()))
            if arr_closure is not None or (print_iters and print_loss):
                sq_err_loss = closure(add_prior=False)
                if arr_closure is not None:
                    arr_closure(self.loss, sq_err_loss)
            if print_iters:
                print(("Burn-in iter {:04d} | accepted={}" if i < burn_in else "Sample iter {:04d} | accepted={}").format(
                    i + 1 if i < burn_in else i - burn_in + 1, int(acc.sum())))
        return chain


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
                                           float(group.get("lr", 0.0) if use_ctl else group["lr"]), float(group["alpha"]), float(group["lambda_"]),
                                           int(bool(group["add_noise"])), self.seed + k, self._step_index,
                                           _lib.ptr(self._status), _lib.ptr(ctl), _lib.stream_ptr()))
        self._after_step()


class HAMCMC(_LangevinBase):
    """langevin.py:619-1107: L-BFGS-preconditioned Langevin dynamics (marked DEBUG in the reference and reproduced as is).
    HAMCMC(params, memory=5, lr0=, lr_gamma=, lr_t0=, lr_alpha=, trust_reg=1.0, H_gamma=1.0, add_noise=True).
    Every chain (row of the flat theta[P, d] buffer, or the concatenation of the parameter tensors for a single chain)
    keeps its own history window and (s, y) pairs on the device; one CTA per chain applies the product-form BFGS."""

    def __init__(self, params, memory=5, **kwargs):
        defaults = kwargs
        defaults.setdefault("add_noise", True)
        super().__init__(params, defaults)
        g0 = self.param_groups[0]
        g0.setdefault("trust_reg", 1e0)
        g0.setdefault("H_gamma", 1e0)
        self.loss = None
        self.memory = memory + 1                       # langevin.py:645
        self._user_memory = int(memory)
        self._hs = None

    # -- flat view of the chain state ---------------------------------------------------------------------
    def _chain_buffers(self):
        """(theta [P,d], grad [P,d], scatter_back) -- the flat buffers, or packed copies for free-standing tensors
        (one chain, parameters_to_vector order, langevin.py:944)."""
        gf = self._grad_flat()
        if self._flat is not None and gf is not None:
            return self._flat, gf, None
        if self._flat is not None:
            self._tensors_for_launch()
            return self._flat, self._grad_flat(), None
        th = torch.cat([p.data.reshape(-1) for p in self._plist])[None].contiguous()
        gr = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self._plist])[None].contiguous()

        def scatter():
            o = 0
            for p in self._plist:
                n = p.numel()
                p.data.copy_(th[0, o:o + n].view_as(p))
                o += n
        return th, gr, scatter

    def _state(self, P, d, dev):
        if self._hs is None:
            lib = _lib.load()
            z = lambda w: torch.zeros(lib.bode_hamcmc_floats(P, d, self._user_memory, w), dtype=torch.float32, device=dev)
            self._hs = dict(ht=z(0), hg=z(0), ps=z(1), py=z(1), wk=z(2), meta=torch.zeros(P, 4, dtype=torch.int32, device=dev))
        return self._hs

    def _launch(self, lr, metric, add_params, add_noise, noise):
        lib = _lib.load()
        th, gr, scatter = self._chain_buffers()
        P, d = th.shape
        hs = self._state(P, d, th.device)
        g0 = self.param_groups[0]
        xi = None
        if noise is not None:
            xi = torch.as_tensor(noise).to(th.device, torch.float32).reshape(P, d).contiguous()
        _lib.check(lib.bode_hamcmc_step(P, d, self._user_memory, _lib.ptr(hs["ht"]), _lib.ptr(hs["hg"]), _lib.ptr(hs["ps"]),
                                        _lib.ptr(hs["py"]), _lib.ptr(hs["wk"]), _lib.ptr(hs["meta"]), _lib.ptr(th), d, _lib.ptr(gr), d,
                                        _lib.ptr(xi), float(lr), float(g0["H_gamma"]), float(g0["trust_reg"]), int(metric),
                                        int(add_params), int(add_noise), self.seed, self._step_index, _lib.ptr(self._status),
                                        _lib.stream_ptr()))
        if scatter is not None:
            scatter()
        self._after_step()

    def step_without_metric(self, lr, update_metric=True, add_noise=True, add_params=False, noise=None):
        self._launch(lr, False, add_params and update_metric, add_noise, noise)

    def step(self, lr, use_old_lbfgs=False, add_noise=True, noise=None):
        if use_old_lbfgs:
            raise NotImplementedError("the dense-BFGS + Cholesky variant (langevin.py:669-715) is not built")
        self._launch(lr, True, False, add_noise, noise)

    def n_pairs(self):
        return None if self._hs is None else self._hs["meta"][:, 2].clone()

    def sample(self, closure, num_samples=1000, burn_in=100, print_iters=True, print_loss=False, use_metric=True,
               use_old_lbfgs=False, add_noise=True, thinning=1):
        """langevin.py:1057-1107; returns (chain, logp_array) like the reference."""
        chain = self.samples
        logp_array = []
        fused = hasattr(closure, "loss_and_grad_") and self._flat is not None
        if fused:
            if self._grad_flat() is None and hasattr(closure.field, "bind_flat_grads"):
                closure.field.bind_flat_grads()
            chain.reserve((num_samples + thinning - 1) // thinning, self._flat, self._plist)
        for i in range(burn_in + num_samples):
            if fused:
                self.loss = closure.loss_and_grad_()[0]
            else:
                self.zero_grad()
                self.loss = closure()
                self._backward(self.loss)
            lr = self.get_lr(i)
            if i < burn_in and i < self.memory * 2 - 1 + 100:                 # langevin.py:1068-1069
                self.step_without_metric(lr=lr, add_noise=add_noise, add_params=(i >= 100))
            elif use_metric:
                self.step(lr=lr, use_old_lbfgs=use_old_lbfgs, add_noise=add_noise)
            else:
                self.step_without_metric(lr=lr, update_metric=use_metric, add_noise=add_noise)
            logp_array.append(-self.loss.detach())
            if i >= burn_in and (i - burn_in) % thinning == 0:
                self._record(chain)
            if print_iters:
                tag, k = ("Burn-in", i + 1) if i < burn_in else ("Sample", i - burn_in + 1)
                if print_loss:
                    print("{} iter {:04d} | loss {:.06f}".format(tag, k, float(closure(add_prior=False).sum())))
                else:
                    print("{} iter {:04d}".format(tag, k))
        return chain, logp_array


class _HAMCMCContiguous(HAMCMC):
    """langevin.py:1109-1470: HAMCMC2 / HAMCMC3 / HAMCMC4, the variants that build the L-BFGS pairs from contiguous samples
    (``bode_hamcmc_contig_step``; bookkeeping restated in oracle/samplers.py::HAMCMCContiguous, which is pinned to reference runs of
    all three).  Same constructor and stepping API as the reference classes: the first ``self.memory`` iterations of ``sample``
    are plain Langevin steps that fill the window (:1254), every later one is a metric step.  A metric step on a window that is
    not full yet (``step()`` called directly, or ``burn_in < memory + 1``) raises RuntimeError and leaves the parameters unchanged
    (the reference fails there with an IndexError)."""

    _variant = None

    def _state(self, P, d, dev):
        if self._hs is None:
            lib = _lib.load()
            z = lambda w: torch.zeros(lib.bode_hamcmc_contig_floats(P, d, self._user_memory, w), dtype=torch.float32, device=dev)
            self._hs = dict(ht=z(0), hg=z(0), ps=z(1), py=z(1), wk=z(2), meta=torch.zeros(P, 4, dtype=torch.int32, device=dev))
        return self._hs

    def _launch(self, lr, metric, add_params, add_noise, noise):
        lib = _lib.load()
        th, gr, scatter = self._chain_buffers()
        P, d = th.shape
        hs = self._state(P, d, th.device)
        g0 = self.param_groups[0]
        xi = None
        if noise is not None:
            xi = torch.as_tensor(noise).to(th.device, torch.float32).reshape(P, d).contiguous()
        _lib.check(lib.bode_hamcmc_contig_step(int(self._variant), P, d, self._user_memory, _lib.ptr(hs["ht"]), _lib.ptr(hs["hg"]),
                                               _lib.ptr(hs["ps"]), _lib.ptr(hs["py"]), _lib.ptr(hs["wk"]), _lib.ptr(hs["meta"]),
                                               _lib.ptr(th), d, _lib.ptr(gr), d, _lib.ptr(xi), float(lr), float(g0["H_gamma"]),
                                               float(g0["trust_reg"]), int(metric), int(add_params), int(add_noise), self.seed,
                                               self._step_index, _lib.ptr(self._status), _lib.stream_ptr()))
        if scatter is not None:
            scatter()
        self._after_step()

    def step_without_metric(self, lr, update_metric=True, add_noise=True, noise=None):
        self._launch(lr, False, update_metric, add_noise, noise)

    def step(self, lr, use_old_lbfgs=False, add_noise=True, noise=None):
        if use_old_lbfgs:
            raise NotImplementedError("the dense-BFGS + Cholesky variant (langevin.py:669-715) is not built")
        self._launch(lr, True, False, add_noise, noise)

    def sample(self, closure, num_samples=1000, burn_in=100, print_loss=False, use_metric=True, use_old_lbfgs=False, add_noise=True,
               print_iters=True, thinning=1):
        """langevin.py:1243-1290; returns (chain, logp_array) like the reference."""
        chain = self.samples
        logp_array = []
        fused = hasattr(closure, "loss_and_grad_") and self._flat is not None
        if fused:
            if self._grad_flat() is None and hasattr(closure.field, "bind_flat_grads"):
                closure.field.bind_flat_grads()
            chain.reserve((num_samples + thinning - 1) // thinning, self._flat, self._plist)
        for i in range(burn_in + num_samples):
            if fused:
                self.loss = closure.loss_and_grad_()[0]
            else:
                self.zero_grad()
                self.loss = closure()
                self._backward(self.loss)
            lr = self.get_lr(i)
            if i < burn_in and i < self.memory:                                  # langevin.py:1254
                self.step_without_metric(lr=lr, add_noise=add_noise)
            elif use_metric:
                self.step(lr=lr, use_old_lbfgs=use_old_lbfgs, add_noise=add_noise)
            else:
                self.step_without_metric(lr=lr, update_metric=use_metric, add_noise=add_noise)
            logp_array.append(-self.loss.detach())
            if i >= burn_in and (i - burn_in) % thinning == 0:
                self._record(chain)
            if print_iters:
                tag, k = ("Burn-in", i + 1) if i < burn_in else ("Sample", i - burn_in + 1)
                if print_loss:
                    print("{} iter {:04d} | loss {:.06f}".format(tag, k, float(closure(add_prior=False).sum())))
                else:
                    print("{} iter {:04d}".format(tag, k))
        return chain, logp_array


class HAMCMC2(_HAMCMCContiguous):
    """langevin.py:1109-1290: theta_t = theta_{t-M} - lr H(theta_{t-M+1:t-1}) grad + noise (base point = the oldest stored sample)."""
    _variant = 2


class HAMCMC3(_HAMCMCContiguous):
    """langevin.py:1292-1400: base point = the newest stored sample, the pair that joins the window lags one sample behind."""
    _variant = 3


class HAMCMC4(_HAMCMCContiguous):
    """langevin.py:1402-1470: like HAMCMC3 with the newest pair included (M-1 pairs)."""
    _variant = 4
