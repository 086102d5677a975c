"""Sampler base class and the flat-buffer plumbing shared by all B200 samplers.

Mirrors samplers/sampler.py:9-21 (``Sampler(torch.optim.Optimizer)`` with a ``samples`` chain) and adds what
particle-batched chains need: detection of parameters that are column blocks of one resident theta[P, d]
buffer (NPDEField / MLPField lay them out that way) so an update is ONE fused launch over the flat buffer,
an on-device chain store, and deferred (asynchronous) NaN reporting.
"""
import numpy as np
import torch
from torch.optim.optimizer import Optimizer

from .. import _lib



def _default_seed():
    import torch.distributed as dist
    draw = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return (draw ^ (rank * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF


def _flat_base(tensors):
    """If every tensor is a dense column block [P, ...] of one row-major [P, d] buffer and together they tile its
    columns, return that buffer as a [P, d] view; else None."""
    ts = list(tensors)
    if not ts or any((not t.is_cuda) or t.dtype != torch.float32 or t.dim() < 2 for t in ts):
        return None
    st = ts[0].untyped_storage().data_ptr()
    P, d = ts[0].shape[0], ts[0].stride(0)
    blocks = []
    for t in ts:
        if t.untyped_storage().data_ptr() != st or t.shape[0] != P or (P > 1 and t.stride(0) != d):
            return None
        inner = t[0].numel()
        if not t[0].is_contiguous():
            return None
        blocks.append((t.storage_offset(), inner))
    blocks.sort()
    off0 = blocks[0][0]
    pos = off0
    for o, n in blocks:
        if o != pos:
            return None
        pos += n
    if P == 1:
        d = pos - off0
    if pos - off0 != d:
        return None
    return torch.as_strided(ts[0], (P, d), (d, 1), off0)


class ChainStore:
    """List-like view of an on-device chain [num_samples, P, d]; entries materialise lazily in the reference's
    format ``([[param arrays ...]], True)`` (langevin.py:243-245) so no per-sample D2H copy or sync happens.
    Entries keep their order of arrival whether they were appended the reference way (host), pushed from the flat
    device buffer, or recorded as the cyclical samplers' ``[[None, ...]]`` placeholders (langevin.py:1703-1706)."""

    def __init__(self):
        self._host = []
        self._dev = None
        self._count = 0
        self._split = None
        self._order = []          # ("h", index into _host) | ("d", device slot) | ("n", number of parameters)

    def reserve(self, n, flat, params):
        self._dev = torch.empty((n,) + tuple(flat.shape), dtype=flat.dtype, device=flat.device)
        self._count = 0
        self._order = [e for e in self._order if e[0] != "d"]
        offs, pos = [], 0
        for p in params:
            inner = p[0].numel()
            offs.append((pos, inner, tuple(p.shape)))
            pos += inner
        self._split = offs

    def push_flat(self, flat):
        self._dev[self._count].copy_(flat, non_blocking=True)
        self._order.append(("d", self._count))
        self._count += 1

    def push_none(self, nparams):
        self._order.append(("n", nparams))

    def append(self, item):
        self._order.append(("h", len(self._host)))
        self._host.append(item)

    def device_tensor(self):
        return None if self._dev is None else self._dev[:self._count]

    def __len__(self):
        return len(self._order)

    def _entry(self, i):
        kind, k = self._order[i]
        if kind == "h":
            return self._host[k]
        if kind == "n":
            return ([[None] * k], True)
        row = self._dev[k].cpu().numpy()
        return ([[row[:, o:o + n].reshape(shape) for (o, n, shape) in self._split]], True)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._entry(j) for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self._entry(i)

    def __iter__(self):
        return (self._entry(i) for i in range(len(self)))


class Sampler(Optimizer):
    """samplers/sampler.py:9-21.  ``self.samples`` holds the chain as ``[(params, accepted), ...]``."""

    def __init__(self, params, defaults):
        self.samples = ChainStore()
        super().__init__(params, defaults)
        self._plist = [p for g in self.param_groups for p in g["params"]]
        _lib.require_cuda(*self._plist)
        self._flat = _flat_base([p.data for p in self._plist])
        self._flat_grad = None
        self._status = torch.zeros(1, dtype=torch.int32, device=self._plist[0].device)
        self._ctl_dev = None
        self._step_index = 0
        # Philox key of the in-kernel noise.  ``seed=`` pins it (reproducible runs); by default it is a fresh draw from torch's
        # global generator -- so ``torch.manual_seed`` governs it like it governs the reference's ``Normal(...).sample()`` draws
        # (langevin.py:194-199) and successive samplers differ -- mixed with the distributed rank, so that the shards of a
        # multi-GPU job never replay each other's stream.
        seed = defaults.get("seed") if isinstance(defaults, dict) else None
        self.seed = int(seed) if seed is not None else _default_seed()
        self.check_finite = "sync"       # "sync": raise inside step() like the reference; "deferred": raise at check()

    # ---- flat buffers ------------------------------------------------------------------------------------
    def _grad_flat(self):
        """The gradient as one [P, d] buffer matching ``self._flat`` (bound lazily: grads may be created by autograd)."""
        if self._flat is None:
            return None
        grads = [p.grad for p in self._plist]
        if any(g is None for g in grads):
            return None
        gf = _flat_base(grads)
        if gf is None or gf.shape != self._flat.shape:
            return None
        return gf

    def bind_flat_grads(self):
        """Allocate one flat gradient buffer and make every ``p.grad`` a column-block view of it."""
        if self._flat is None:
            return None
        gf = self._grad_flat()
        if gf is not None:
            return gf
        gf = torch.zeros_like(self._flat)
        base_off = self._flat.storage_offset()
        d = self._flat.shape[1]
        for p in self._plist:
            o = p.data.storage_offset() - base_off
            n = p[0].numel()
            p.grad = gf[:, o:o + n].view(p.shape)
        return gf

    def zero_grad(self, set_to_none=False):
        # in place, so flat-bound gradient views survive (the reference's torch version zeroed in place too)
        with torch.no_grad():
            for p in self._plist:
                if p.grad is not None:
                    p.grad.zero_()

    def _tensors_for_launch(self):
        """[(param, grad)] pairs a fused kernel can run on: the flat pair, or each contiguous tensor."""
        gf = self._grad_flat()
        if gf is not None:
            return [(self._flat, gf)]
        if self._flat is not None and all(p.grad is not None for p in self._plist):
            # gradients were produced by plain autograd as separate tensors: pack them once into a flat buffer and
            # rebind p.grad to its column blocks so later backward() calls accumulate in place
            grads = [p.grad for p in self._plist]
            for p in self._plist:
                p.grad = None
            gf = self.bind_flat_grads()
            with torch.no_grad():
                for p, g in zip(self._plist, grads):
                    p.grad.copy_(g)
            return [(self._flat, gf)]
        out = []
        for p in self._plist:
            if p.grad is None:
                continue
            if not p.data.is_contiguous() or not p.grad.is_contiguous():
                raise _lib.BodeError("sampler parameters must be contiguous or column blocks of one flat buffer")
            out.append((p.data, p.grad))
        return out

    # ---- NaN reporting (langevin.py:184-185) -------------------------------------------------------------
    def check(self):
        st = int(self._status.item())
        if st != 0:
            self._status.zero_()
            if st & 2:
                raise RuntimeError("metric step before the history window is full: run `memory + 1` step_without_metric() "
                                   "iterations first (sample() does, langevin.py:1254); parameters were left unchanged")
            raise ValueError("Encountered NaN/Inf in parameter")

    def _after_step(self):
        self._step_index += 1
        if self.check_finite == "sync":
            self.check()

    # ---- device control block for graph replay -----------------------------------------------------------
    def ctl(self):
        """Device control block (bode_sampler_ctl, 8 x 32 bit).  ``schedule_()`` advances it on the device."""
        if self._ctl_dev is None:
            self._ctl_dev = torch.zeros(8, dtype=torch.int32, device=self._status.device)
        return self._ctl_dev

    def schedule_(self, kind=0, lr0=0.0, gamma=0.0, t0=0.0, alpha=0.0, burn_in_iters=0, resample_every=0):
        """Enqueue the device-side schedule for the next iteration (graph-capturable; see bode_sampler_schedule)."""
        _lib.check(_lib.load().bode_sampler_schedule(_lib.ptr(self.ctl()), int(kind), float(lr0), float(gamma), float(t0),
                                                     float(alpha), int(burn_in_iters), int(resample_every), _lib.stream_ptr()))

    def step(self):
        raise NotImplementedError

    def sample(self):
        raise NotImplementedError

    # ---- helpers shared by the sample() loops ----------------------------------------------------------
    @staticmethod
    def _backward(loss):
        if loss.dim() == 0:
            loss.backward()
        else:
            loss.backward(torch.ones_like(loss))      # independent chains: d(sum_p loss_p)/d theta_p

    def _record(self, chain):
        if self._flat is not None and chain._dev is not None:
            chain.push_flat(self._flat)
        else:
            params = [[p.clone().detach().data.cpu().numpy() for p in group["params"]] for group in self.param_groups]
            chain.append((params, True))
