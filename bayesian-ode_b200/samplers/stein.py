"""Stein variational gradient descent over particle-sharded chains.

The reference's ``samplers/stein.py`` holds a usable kernel (``RBFKernel`` :12-34, median heuristic) and the canonical
``phi`` formula (:75-86) but an unrunnable ``SVGD.step`` (:94-106).  This module completes it:

    phi_i = (1/n) [ sum_j K_ij s_j + grad-K term ],  s = grad log p = -grad loss,  K = exp(-gamma ||x_i - x_j||^2)
    theta_i <- theta_i + lr * phi_i         (the wrapped optimiser of the reference descends -phi, lr default 1e-4, :40-41)

Particles are the rows of one resident theta[P_local, d] buffer per GPU.  With ``torch.distributed`` initialised the
only data-path collective is one all-gather of [theta | score] per step (NCCL over NVLink); the median bandwidth is an
exact distributed radix select whose three tiny histogram all-reduces ride the same communicator.
"""
import ctypes as C
import math

import torch

from .. import _lib
from .sampler import Sampler


class RBFKernel(torch.nn.Module):
    """stein.py:12-34.  ``forward(X, Y)`` returns K_XY = exp(-gamma ||x - y||^2); sigma=None -> median heuristic
    h = median(d2) / (2 ln(n + 1)), sigma = sqrt(h), gamma = 1 / (1e-8 + 2 sigma^2)."""

    def __init__(self, sigma=None):
        super().__init__()
        self.sigma = sigma
        self.last_median = None
        self.last_gamma = None

    def forward(self, X, Y):
        _lib.require_cuda(X, Y)
        lib = _lib.load()
        X = X.detach().to(torch.float32).contiguous()
        Y = Y.detach().to(torch.float32).contiguous()
        n, m, d = X.shape[0], Y.shape[0], X.shape[1]
        ws = _Workspace(n, m, d, X.device)
        ws.sqdist(X, n, Y, m, d, n * m, row_offset=0 if (n == m and (X.data_ptr() == Y.data_ptr() or torch.equal(X, Y))) else -1)
        ws.median(n, m, d, X.shape[0], self.sigma)
        d2 = ws.d2(n, m)
        mg = ws.med_gamma
        self.last_median, self.last_gamma = mg[0], mg[1]
        return torch.exp(-mg[1] * d2)


def radix_select_protocol(hist_pass, allreduce, select_digit, passes=3):
    """The exact-median protocol every rank runs in lock-step: for each radix pass, histogram the local block, SUM the
    histograms over ranks (the only communication: 2 x 2048 counters), then pick the digit -- so all ranks extend the
    same prefix and end with the same order statistics.  ``allreduce`` is None on a single rank."""
    for ps in range(passes):
        hist_pass(ps)
        if allreduce is not None:
            allreduce()
        select_digit(ps)


def shard_range(n, rank, world):
    """Contiguous block of particles / chains owned by ``rank``."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def untile_d2(flat, nr, nc):
    """[nr/128][nc/32][128][32] stage tiles (the layout the pipelined kernels keep d2 in when nr and nc are multiples of 128,
    ``bode_svgd_d2_tiled``; element (i, j) sits in tile (i // 128, j // 32) at (i % 128, j % 32)) -> the [nr, nc] matrix."""
    return flat.view(nr // 128, nc // 32, 128, 32).permute(0, 2, 1, 3).reshape(nr, nc)


class _RawCuda:
    """Zero-copy torch view of a raw device allocation (``torch.as_tensor`` reads ``__cuda_array_interface__``)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class _Workspace:
    """SVGD workspace (d2 block, operands, median state).  With ``peers=(rank, world)`` the block is a cudaMalloc allocation
    mapped into every other rank of the node (CUDA IPC handles exchanged through torch.distributed), so that the exact
    distributed median reads the peers' window tables / radix histograms over NVLink instead of launching collectives
    (``bode_svgd_set_peers``); ``self.p2p`` tells whether that mapping is active on ALL ranks."""

    def __init__(self, nr, nc, d, device, peers=None):
        lib = _lib.load()
        nbytes = lib.bode_svgd_workspace_bytes(nr, nc, d)
        self._raw = None
        self._imported = []
        self.p2p = False
        if peers is not None:
            p = C.c_void_p()
            _lib.check(lib.bode_peer_alloc(nbytes, C.byref(p)))
            self._raw = p.value
            self.base = torch.as_tensor(_RawCuda(self._raw, nbytes), device=device)
            self.buf = self.base
        else:
            self.buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            off = (-self.buf.data_ptr()) % 256
            self.base = self.buf[off:off + nbytes]
        self.nbytes = nbytes
        self.med_gamma = torch.zeros(2, dtype=torch.float32, device=device)
        self.hist_ptr = C.c_void_p()
        self._hist = None
        self._gath = [None, None]
        self._dims = (nr, nc, d)
        _lib.check(lib.bode_svgd_workspace_init(nr, nc, d, C.c_void_p(self.base.data_ptr()), nbytes, _lib.stream_ptr()))
        tp, tn = C.c_void_p(), C.c_size_t()
        _lib.check(lib.bode_svgd_window_table(nr, nc, d, C.c_void_p(self.base.data_ptr()), C.byref(tp), C.byref(tn)))
        off = tp.value - self.base.data_ptr()
        self._wtable = self.base[off:off + tn.value * 8].view(torch.int64)      # median window counters (summed over ranks)
        if peers is not None:
            self._map_peers(*peers)

    def _map_peers(self, rank, world):
        """Exchange IPC handles and register the peer addresses; falls back (on every rank) to the collective protocol when
        any rank could not map its peers."""
        lib = _lib.load()
        dist = torch.distributed
        nr, nc, d = self._dims
        ok = 1
        bases = (C.c_void_p * world)()
        try:
            h = (C.c_ubyte * 64)()
            _lib.check(lib.bode_peer_export(C.c_void_p(self._raw), h))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(h))
            for q in range(world):
                if q == rank:
                    bases[q] = self._raw
                else:
                    p = C.c_void_p()
                    hq = (C.c_ubyte * 64).from_buffer_copy(handles[q])
                    _lib.check(lib.bode_peer_import(hq, C.byref(p)))
                    self._imported.append(p.value)
                    bases[q] = p.value
        except Exception:                                        # noqa: BLE001 -- any failure means "no peer access here"
            ok = 0
        flag = torch.tensor([ok], device=self.base.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            _lib.check(lib.bode_svgd_set_peers(C.c_void_p(self._raw), nr, nc, d, bases, rank, world))
            self.p2p = True
        torch.cuda.synchronize()
        dist.barrier()                                           # every rank's workspace is initialised and mapped

    def close(self):
        lib = _lib.load()
        if self.p2p:
            nr, nc, d = self._dims
            lib.bode_svgd_set_peers(C.c_void_p(self._raw), nr, nc, d, None, 0, 1)
            self.p2p = False
        for p in self._imported:
            lib.bode_peer_release(C.c_void_p(p))
        self._imported = []
        if self._raw is not None:
            torch.cuda.synchronize()
            lib.bode_peer_free(C.c_void_p(self._raw))
            self._raw = None

    def __del__(self):
        try:
            self.close()
        except Exception:                                        # noqa: BLE001 -- interpreter shutdown
            pass

    def sqdist(self, Xr, nr, Xc, nc, d, total, row_offset=-1, stages=_lib.SVGD_PREPARE | _lib.SVGD_COMPUTE):
        lib = _lib.load()
        rp, rs = _lib.rows(Xr, d)
        cp, cs = _lib.rows(Xc, d)
        _lib.check(lib.bode_svgd_sqdist_staged(int(stages), rp, rs, nr, cp, cs, nc, d, int(row_offset), int(total),
                                               C.c_void_p(self.base.data_ptr()), self.nbytes, C.byref(self.hist_ptr), _lib.stream_ptr()))
        if self._hist is None:
            off = self.hist_ptr.value - self.base.data_ptr()
            self._hist = self.base[off:off + 2 * 2048 * 8].view(torch.int64)

    def median(self, nr, nc, d, n_total, sigma=None, group=None):
        lib = _lib.load()
        w = C.c_void_p(self.base.data_ptr())
        if sigma is None and self.p2p:
            # peer-mapped workspaces: the single-rank call order; the kernels sum the peers' tables / histograms themselves
            _lib.check(lib.bode_svgd_window_select(nr, nc, d, w, _lib.stream_ptr()))
            _lib.check(lib.bode_svgd_radix_fallback(nr, nc, d, w, n_total, _lib.ptr(self.med_gamma), _lib.stream_ptr()))
            return
        if sigma is None:
            # fast path: the window around the previous call's median usually holds both middle ranks (one all-reduce)
            if group is not None:
                torch.distributed.all_reduce(self._wtable, group=group if group is not True else None)
            _lib.check(lib.bode_svgd_window_select(nr, nc, d, w, _lib.stream_ptr()))
            if group is None:
                # single rank: the radix passes are one cooperative launch that returns at once after a window hit
                _lib.check(lib.bode_svgd_radix_fallback(nr, nc, d, w, n_total, _lib.ptr(self.med_gamma), _lib.stream_ptr()))
                return                                   # gamma and the next window were written by the same launch
            else:
                radix_select_protocol(
                    lambda ps: _lib.check(lib.bode_svgd_hist_pass(ps, nr, nc, d, w, _lib.stream_ptr())),
                    lambda: torch.distributed.all_reduce(self._hist, group=group if group is not True else None),
                    lambda ps: _lib.check(lib.bode_svgd_select_digit(ps, nr, nc, d, w, _lib.stream_ptr())))
        _lib.check(lib.bode_svgd_gamma(n_total, float(sigma or 0.0), nr, nc, d, w, _lib.ptr(self.med_gamma), _lib.stream_ptr()))

    def peer_gather(self, which, X):
        """All-gather of this rank's rows over peer memory (``bode_svgd_peer_gather``: one push kernel between two flag barriers,
        no collective).  which = 0 positions, 1 scores.  Returns the [n_cols, d] buffer inside the workspace."""
        lib = _lib.load()
        nr, nc, d = self._dims
        xp, xs = _lib.rows(X, d)
        out = C.c_void_p()
        _lib.check(lib.bode_svgd_peer_gather(int(which), xp, xs, nr, nc, d, C.c_void_p(self.base.data_ptr()), C.byref(out),
                                             _lib.stream_ptr()))
        if self._gath[which] is None:
            off = out.value - self.base.data_ptr()
            self._gath[which] = self.base[off:off + nc * d * 4].view(torch.float32).view(nc, d)
        return self._gath[which]

    def d2(self, nr, nc):
        """The squared-distance block as a [nr, nc] matrix (a view when the workspace holds it row-major, a gathered copy when
        it is stored as [nr/128][nc/32][128][32] tiles, ``bode_svgd_d2_tiled``)."""
        flat = self.base[:nr * nc * 4].view(torch.float32)
        if _lib.load().bode_svgd_d2_tiled(nr, nc, self._dims[2]):
            return untile_d2(flat, nr, nc)
        return flat.view(nr, nc)


class SVGD(Sampler):
    """SVGD(params, kernel=RBFKernel(), lr=1e-4) over the particle axis (dim 0) of the parameters.

    ``step(lr=None)`` consumes the gradients of the per-particle negative log posterior left in ``p.grad`` (score =
    -grad), all-gathers [theta | score] across ranks if torch.distributed is initialised, and applies theta += lr*phi.

    Stream overlap.  Half of the interaction depends only on the particle POSITIONS, not on the scores: the centred,
    pre-split Gram operands, the Gram kernel (d2) and the exact median / bandwidth.  ``prefetch()`` forks that half onto a
    side stream so it runs beside the ODE solve that produces the scores (``sample()`` calls it before the closure; call
    it yourself before ``closure.loss_and_grad_()`` in a hand-written loop) and ``phi()`` joins it:
      ``overlap="gram"`` (default)  operands + Gram + median on the side stream.  The Gram CTAs cannot share an SM with the
                                    solver's, so ``side_sms`` SMs (default 40) are kept free of the fused solve while it is
                                    prefetched (``bode_npde_set_cta_limit``); the solve is latency-bound at three warps per
                                    scheduler either way;
      ``overlap="operands"``        only the operand preparation (forking the small ``[-G | X - mu | 1]`` operand kernel
                                    beside the Gram kernel as well was measured and bought nothing);
      ``overlap=False``             everything in call order on one stream.
    Every fork rejoins the calling stream, so the sequence is capturable in one CUDA graph, and all three orders produce
    bit-identical particles.  Nothing may write the particles between ``prefetch()`` and ``phi()``.
    """

    def __init__(self, params, optimizer=None, kernel=None, num_particles=None, particle_init_fn=None, overlap="gram", side_sms=40,
                 median_comm="p2p", gather_comm="p2p", **kwargs):
        defaults = kwargs
        if "lr" not in defaults:
            defaults["lr"] = 1e-4                        # stein.py:40-41
        super().__init__(params, defaults)
        if self._flat is None:
            raise _lib.BodeError("SVGD needs parameters laid out as column blocks of one theta[P, d] buffer")
        self.kernel = kernel if kernel is not None else RBFKernel()
        self.loss = None
        self.P_local, self.d = self._flat.shape
        dist = torch.distributed
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n_total = self.P_local * self.world
        dev = self._flat.device
        if self.world > 1:
            # the exchange buffers and the Gram blocks assume equal shards (shard_range() hands out uneven ones when P % world != 0)
            sizes = torch.tensor([self.P_local, -self.P_local], device=dev)
            dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
            if int(sizes[0]) != self.P_local or int(-sizes[1]) != self.P_local:
                raise _lib.BodeError("SVGD: every rank must hold the same number of particles (got %d here, %d..%d over the ranks)"
                                     % (self.P_local, int(-sizes[1]), int(sizes[0])))
        # several ranks: "p2p" maps the workspaces into each other (NVLink peer reads + flag barriers inside the median kernels);
        # "nccl" keeps the collective protocol (one table all-reduce + three histogram all-reduces per step)
        if median_comm not in ("p2p", "nccl"):
            raise ValueError("median_comm must be 'p2p' or 'nccl'")
        use_p2p = (self.world > 1 and self.world <= 8 and median_comm == "p2p" and getattr(self.kernel, "sigma", None) is None
                   and bool(_lib.load().bode_svgd_staged_supported(self.n_total, self.d)))
        self._ws = _Workspace(self.P_local, self.n_total, self.d, dev, peers=(self.rank, self.world) if use_p2p else None)
        self.median_comm = "p2p" if self._ws.p2p else ("nccl" if self.world > 1 else "local")
        # the data-path exchange: "p2p" = push kernels over the peer-mapped workspaces (needs the mapping above), "nccl" = all-gather
        if gather_comm not in ("p2p", "nccl"):
            raise ValueError("gather_comm must be 'p2p' or 'nccl'")
        self.gather_comm = "p2p" if (self._ws.p2p and gather_comm == "p2p") else ("nccl" if self.world > 1 else "local")
        self.phi_buf = torch.empty_like(self._flat)
        self._Xall = None
        if self.world > 1 and self.gather_comm != "p2p":
            self._gath = torch.empty(2, self.n_total, self.d, dtype=torch.float32, device=dev)
        if overlap is True:
            overlap = "operands"
        if overlap not in (False, None, "operands", "gram"):
            raise ValueError("overlap must be False, 'operands' or 'gram'")
        self.overlap = overlap if (overlap and _lib.load().bode_svgd_staged_supported(self.n_total, self.d)) else False
        self.side_sms = int(side_sms)
        self.fuse_scores = True          # let the closure kernel write the score half of the phi operand (single rank, overlap="gram")
        self._scores_armed = False
        self._side2_used = False
        self.last_scores_fused = False
        self._side = torch.cuda.Stream(device=dev) if self.overlap else None
        self._side2 = torch.cuda.Stream(device=dev) if self.overlap else None     # position half of the phi operand, beside the Gram pass
        self._prefetched = False
        self._saved_cta_limit = None

    def _gather(self, which, X):
        """The data-path exchange: rows of every rank -> [n_total, d] (which = 0 positions, 1 scores)."""
        if self.world == 1:
            return X
        if self.gather_comm == "p2p":
            return self._ws.peer_gather(which, X)
        torch.distributed.all_gather_into_tensor(self._gath[which], X if X.is_contiguous() else X.contiguous())
        return self._gath[which]

    def _gather_positions(self, X):
        return self._gather(0, X)

    def prefetch(self):
        """Fork the position-only part of the interaction (all-gather of the positions, pre-split Gram operands and, with
        ``overlap="gram"``, the Gram kernel and the median selection) onto the side stream; the next ``phi()`` joins it.
        A no-op with ``overlap=False``."""
        if not self.overlap or self._prefetched:
            return
        lib = _lib.load()
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        nl, nt, d = self.P_local, self.n_total, self.d
        with torch.cuda.stream(self._side):
            Xall = self._Xall = self._gather_positions(self._flat)
            # the local rows as a block of the gathered columns: the Gram kernel then reuses the column operands for them
            Xloc = Xall[self.rank * nl:(self.rank + 1) * nl] if self.world > 1 else self._flat
            self._ws.sqdist(Xloc, nl, Xall, nt, d, nt * nt, row_offset=self.rank * nl, stages=_lib.SVGD_PREPARE)
            if self.overlap == "gram" and self.world == 1 and self.fuse_scores:
                # The position half of the phi operand on a second side stream (it needs the column means just computed, and runs
                # beside the Gram pass -- behind it, it would lengthen the chain phi() waits for); the score half is written by the
                # closure's own kernel while armed (bode_svgd_arm_score_tiles): no operand launch between the solve and phi.
                self._side2.wait_stream(self._side)
                self._side2_used = True
                with torch.cuda.stream(self._side2):
                    xr, xrs = _lib.rows(self._flat, d)
                    _lib.check(lib.bode_svgd_phi_staged(_lib.SVGD_PREPARE_POSITIONS, xr, xrs, nl, xr, xrs, None, 0, -1.0, nt, d, nt,
                                                        _lib.ptr(self._ws.med_gamma), C.c_void_p(self._ws.base.data_ptr()), None, d,
                                                        None, 0, 0.0, _lib.stream_ptr()))
                self._scores_armed = lib.bode_svgd_arm_score_tiles(nl, nt, d, C.c_void_p(self._ws.base.data_ptr()), -1.0) == 1
            if self.overlap == "gram":
                # the pass shares the GPU with the solve: its grid is sized for the side_sms SMs it gets (full waves of CTAs with
                # as many column tiles each as that allows; a grid sized for the whole GPU leaves a long tail there)
                old_split = lib.bode_svgd_set_gram_split(self._side_gram_split(nl, nt))
                self._ws.sqdist(Xloc, nl, Xall, nt, d, nt * nt, row_offset=self.rank * nl, stages=_lib.SVGD_COMPUTE)
                lib.bode_svgd_set_gram_split(old_split)
                if self.world == 1 or self._ws.p2p:
                    # the cooperative fallback launch (a no-op after a window hit) must fit the SMs the solve leaves free, or it
                    # waits for the solve to end
                    # (single rank only: with several ranks the Gram pass outlasts the solve and the SMs are free by then)
                    old_ctas = lib.bode_svgd_set_select_ctas(self.side_sms) if (self.side_sms > 0 and self.world == 1) else None
                    self._ws.median(nl, nt, d, nt, getattr(self.kernel, "sigma", None), group=None)
                    if old_ctas is not None:
                        lib.bode_svgd_set_select_ctas(old_ctas)

                # collective protocol (median_comm="nccl"): the median's all-reduces are issued by phi() AFTER the all-gather
                # of the scores -- one communicator executes collectives in issue order, and the scores are ready long
                # before the Gram pass ends
        if self.overlap == "gram" and self.side_sms > 0:
            sms = lib.bode_device_sm_count()
            if sms > self.side_sms:
                self._saved_cta_limit = lib.bode_npde_set_cta_limit(sms - self.side_sms)   # until phi() restores it
        self._prefetched = self.overlap

    def _side_gram_split(self, nr, nc):
        """Column chunks per row block for a Gram pass that has ``side_sms`` SMs to itself (one CTA per SM): the cost model of
        csrc/svgd_tc2.cu (waves x (tiles per CTA + one tile of prologue)) by enumeration.  4096 x 4096 on 40 SMs: 5 chunks of
        7 tiles = 160 CTAs = 4 full waves (measured: c3 step 142.7 us with 16 chunks, 137.5 us with 4 - 8)."""
        if self.world > 1:
            # several ranks: the rectangular pass outlasts the solve and spreads over the SMs the solve frees -- fine CTAs (16
            # chunks) follow that change of width best (2 GPUs: 199.2 us per step against 200.9 with the enumeration)
            return 16
        sms = max(1, self.side_sms)
        nrb, nct = (nr + 127) // 128, (nc + 127) // 128
        best, best_cost = 1, None
        for tp in range(1, nct + 1):
            chunks = (nct + tp - 1) // tp
            if chunks > 16:
                continue
            cost = ((chunks * nrb + sms - 1) // sms) * (tp + 1)
            if best_cost is None or cost < best_cost:
                best, best_cost = chunks, cost
        return best

    def _restore_cta_limit(self):
        """Undo prefetch()'s process-wide CTA cap of the fused solve (phi() does; so does an aborted step)."""
        if self._saved_cta_limit is not None:
            _lib.load().bode_npde_set_cta_limit(self._saved_cta_limit)
            self._saved_cta_limit = None

    def _join_side2(self, cur):
        if self._side2_used:
            cur.wait_stream(self._side2)
            self._side2_used = False

    def _disarm_scores(self):
        """End the closure kernels' writes into the phi operand tiles; how many launches wrote them since prefetch()."""
        n = 0
        if self._scores_armed:
            n = _lib.load().bode_svgd_disarm_score_tiles()
            self._scores_armed = False
        return n

    def cancel_prefetch(self):
        """Drop a prefetch() whose phi() will not follow (the closure raised): joins the side stream, restores the CTA cap."""
        self._restore_cta_limit()
        self._disarm_scores()
        if self._prefetched:
            torch.cuda.current_stream().wait_stream(self._side)
            self._join_side2(torch.cuda.current_stream())
            self._prefetched = False

    def phi(self, X=None, grad=None, update_lr=None):
        """stein.py:75-86 for the local rows.  Returns phi [P_local, d]; with ``update_lr`` the update is fused."""
        lib = _lib.load()
        X = self._flat if X is None else X
        G = self._grad_flat() if grad is None else grad
        if G is None:
            raise _lib.BodeError("SVGD.phi: gradients missing (call closure + backward, or pass grad=)")
        d, nl, nt = self.d, self.P_local, self.n_total
        ws = self._ws
        cur = torch.cuda.current_stream()
        self._restore_cta_limit()
        scores_in_tiles = self._disarm_scores() >= 1 and X is self._flat and grad is None
        self.last_scores_fused = scores_in_tiles
        prefetched = self._prefetched if X is self._flat else False
        if self._prefetched and prefetched != "gram":
            cur.wait_stream(self._side)                     # join: operands (and the gathered positions) are ready
            self._join_side2(cur)
        if self._prefetched and not prefetched:
            cur.wait_stream(self._side)                     # prefetched for other positions: drop it
            self._join_side2(cur)
        self._prefetched = False
        # the one data-path exchange: all-gather of particle positions and loss gradients (NCCL over NVLink)
        Xall = self._Xall if (prefetched and self.world > 1) else self._gather_positions(X)
        Gall = self._gather(1, G)
        Xloc = Xall[self.rank * nl:(self.rank + 1) * nl] if self.world > 1 else X
        xr, xrs = _lib.rows(X, d)
        xc, xcs = _lib.rows(Xall, d)
        sc, scs = _lib.rows(Gall, d)
        th = (xr, xrs) if update_lr is not None else (None, 0)

        def phi_stage(stages):
            _lib.check(lib.bode_svgd_phi_staged(int(stages), xr, xrs, nl, xc, xcs, sc, scs, -1.0, nt, d, nt, _lib.ptr(ws.med_gamma),
                                                C.c_void_p(ws.base.data_ptr()), _lib.ptr(self.phi_buf), d, th[0], th[1],
                                                float(update_lr or 0.0), _lib.stream_ptr()))
        both = _lib.SVGD_PREPARE | _lib.SVGD_COMPUTE
        if prefetched == "gram":
            if not scores_in_tiles:
                phi_stage(_lib.SVGD_PREPARE)                # the V operand needs the scores: here, while the side stream finishes
            cur.wait_stream(self._side)                     # join: d2 (single rank: also median and gamma) are ready
            self._join_side2(cur)
            if self.world > 1 and not ws.p2p:
                ws.median(nl, nt, d, nt, getattr(self.kernel, "sigma", None), group=True)
            phi_stage(_lib.SVGD_COMPUTE)
        elif prefetched:
            ws.sqdist(Xloc, nl, Xall, nt, d, nt * nt, row_offset=self.rank * nl, stages=_lib.SVGD_COMPUTE)
            ws.median(nl, nt, d, nt, getattr(self.kernel, "sigma", None), group=True if self.world > 1 else None)
            phi_stage(both)
        else:
            ws.sqdist(Xloc, nl, Xall, nt, d, nt * nt, row_offset=self.rank * nl, stages=both)
            ws.median(nl, nt, d, nt, getattr(self.kernel, "sigma", None), group=True if self.world > 1 else None)
            phi_stage(both)
        return self.phi_buf

    def step(self, lr=None, closure=None):
        if closure is not None:
            self.zero_grad()
            self.loss = closure()
            self._backward(self.loss)
        group = self.param_groups[0]
        if lr:
            group["lr"] = lr
        self.phi(update_lr=group["lr"])
        self._after_step()                                   # "sync": check() now, like langevin.py:184-185; "deferred": at check()
        return self.loss

    def check(self):
        """Non-finite particles (the fused phi + update launch carries no status word, so the particles themselves are tested:
        one NaN/Inf poisons every row of the next interaction), plus: did a flag barrier of the peer-memory exchange give up on a
        peer?  Called after every step() unless ``check_finite == "deferred"``, and at the end of sample()."""
        super().check()
        if not bool(torch.isfinite(self._flat).all()):
            raise ValueError("Encountered NaN/Inf in parameter")
        if self._ws.p2p:
            torch.cuda.synchronize()
            t = C.c_int32(0)
            _lib.check(_lib.load().bode_svgd_peer_status(C.c_void_p(self._ws.base.data_ptr()), self.P_local, self.n_total, self.d,
                                                         C.byref(t)))
            if t.value:
                raise _lib.BodeError("SVGD: a peer flag barrier timed out (a rank never arrived); the particles are invalid")

    def get_particles(self):
        return self._flat

    def sample(self, closure, num_samples=1000, burn_in=100, print_iters=False, print_loss=False, arr_closure=None, thinning=1):
        chain = self.samples
        fused = hasattr(closure, "loss_and_grad_")
        if fused and self._grad_flat() is None and hasattr(closure.field, "bind_flat_grads"):
            closure.field.bind_flat_grads()
        chain.reserve((num_samples + thinning - 1) // thinning, self._flat, self._plist)
        mode, self.check_finite = self.check_finite, "deferred"      # one check at the end instead of a device sync per step
        try:
            for i in range(burn_in + num_samples):
                self.prefetch()                              # position-only operands beside the solve
                if fused:
                    self.loss = closure.loss_and_grad_()[0]
                else:
                    self.zero_grad()
                    self.loss = closure()
                    self._backward(self.loss)
                self.step()
                if i >= burn_in and (i - burn_in) % thinning == 0:
                    self._record(chain)
                if arr_closure is not None:
                    arr_closure(self.loss, closure(add_prior=False))
                if mode == "sync" and (i + 1) % 64 == 0:
                    self.check()
        finally:
            self.check_finite = mode
            self.cancel_prefetch()
        self.check()                                         # NaN/Inf particles or a peer barrier that gave up: raise, do not return
        return chain
