"""`bench.py --impl reference`: the UNMODIFIED reference (baseline/_ref, staged by baseline/make_ref.py) timed on the host cores.

What is run is the reference's own code -- ``torchdiffeq.odeint`` / ``odeint_adjoint``, ``scripts.vanderpol.gp.KernelRegression``,
``samplers.langevin.{SGLD,pSGLD,HAMCMC}``, ``samplers.hamiltonian.aSGHMC``, ``samplers.stein.RBFKernel`` -- driven exactly as
``gp.py:290-391`` / ``nn.ipynb`` cells 10-11 drive it: one chain per process, ``torch.set_num_threads(1)``, float64, each timed
iteration = ``zero_grad -> closure -> backward -> sampler.step`` (the reference's extra logging solve, langevin.py:226, and its
prints are excluded).  Restated here because the reference keeps them inside functions / notebooks: the
``loss_closure`` of gp.py:342-353 (12 lines), the data synthesis of gp.ipynb cell 3, the ``NN`` module of nn.ipynb cell 4 and
``SVGD.phi`` of stein.py:75-86 (the reference's own ``step`` is a stub that references undefined names).

A "step" of this arm is a BOUNDED SAMPLE of the workload (the whole workload would take minutes per step): every worker process
advances ONE chain by one sampler iteration; for SVGD a block of ``rows`` particles additionally evaluates phi against all P
particles through the reference's RBFKernel.  The whole-job rate is then
    P * rk_steps / (ceil(P / C) * t_chain + P / rows * t_phi_block)
with C = worker processes = host cores.  Both ``odeint`` (autograd = discrete adjoint, like the B200 kernel) and
``odeint_adjoint`` (what gp.py:26 imports) are timed; the headline uses ``odeint`` (the faster one: conservative for the ratio).
"""
import math
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "torchdiffeq", "__init__.py")) and os.path.isfile(
        os.path.join(REF_DIR, "scripts", "vanderpol", "gp.py"))


def _import_reference():
    from unittest.mock import MagicMock
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.ticker", "seaborn"):
        sys.modules.setdefault(m, MagicMock())
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torch
    if not hasattr(torch, "cholesky") or "deprecat" in (torch.cholesky.__doc__ or "").lower():
        torch.cholesky = torch.linalg.cholesky                        # gp.py:66 uses the removed alias
    torch.set_default_dtype(torch.float64)                            # gp.py:314
    import torchdiffeq
    from scripts.vanderpol import gp
    from samplers import hamiltonian, langevin, stein
    return torch, torchdiffeq, gp, langevin, hamiltonian, stein


def _data(torch, torchdiffeq, gp, T, seed=0):
    """notebooks/jai/gp.ipynb cell 3 (seeded): N=5, R=3, x0 = 2R U - R, t = linspace(0, 7, T), Y = X + 0.1 N(0,1)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    N, R = 5, 3.0
    x0 = torch.from_numpy(2 * R * rng.random((N, 2)) - R)
    t = torch.linspace(0., 7., T, dtype=torch.float32).to(torch.float64)
    with torch.no_grad():
        X = torchdiffeq.odeint(gp.VDP(), x0, t, method="rk4").permute(1, 0, 2)
    Y = X + 0.1 * torch.from_numpy(rng.standard_normal(tuple(X.shape)))
    return N, x0, t, X, Y


def _npde_model(torch, gp, Y, t, M, ell, jitter_seed):
    """gp.py:315-333: inducing grid, gradient-matched whitened U0; + the per-chain jitter of the benchmark init (gp.py:321 scale)."""
    import numpy as np
    Yn = Y.numpy()
    xv = np.linspace(Yn[..., 0].min(), Yn[..., 0].max(), M)
    yv = np.linspace(Yn[..., 1].min(), Yn[..., 1].max(), M)
    xv, yv = np.meshgrid(xv, yv)
    Z = torch.from_numpy(np.array([xv.T.flatten(), yv.T.flatten()]).T)
    D = 2
    F_ = ((Y[:, 1:, :] - Y[:, :-1, :]) / (t[1] - t[0])).contiguous().view(-1, D)
    Z_ = Y[:, :-1, :].contiguous().view(-1, D)
    Kxz = gp.K(Z, Z_, 1.0, ell)
    Kinv = (gp.K(Z_, Z_, 1.0, ell) + 0.2 * torch.eye(Z_.shape[0])).inverse()
    U0 = torch.mm(torch.mm(Kxz, Kinv), F_)
    U0 = torch.mm(torch.linalg.cholesky(gp.K(Z, Z, 1.0, ell)).inverse(), U0)
    g = torch.Generator().manual_seed(1000 + jitter_seed)
    U0 = U0 + 0.1 * torch.randn(M * M, 2, generator=g)
    return gp.KernelRegression(U0, Z, 1.0, ell, 0.1)


def _npde_closure(torch, kreg, odeint, x0, t, Y, scale_by=None):
    """gp.py:342-353 (nested in run_sampler there)."""
    Kzzinv = kreg.Kzzinv

    def loss_closure(add_prior=True):
        xode = odeint(kreg, x0, t, method="rk4").permute([1, 0, 2])
        if add_prior:
            loss = torch.sum((Y - xode) ** 2 / (2 * torch.exp(kreg.logsn) ** 2))
            loss += torch.numel(Y) * torch.sum(kreg.logsn) / 2
            loss += torch.sum(torch.diag(torch.mm(kreg.U.t(), torch.mm(Kzzinv, kreg.U)))) / 2
        else:
            loss = torch.sum((Y - xode) ** 2)
        return loss
    return loss_closure


def _time_loop(step_fn, warmup, steps):
    for _ in range(warmup):
        step_fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step_fn()
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts)


def chain_worker(args):
    """One process = one chain.  Returns {variant: seconds per sampler iteration} (+ attempted dopri5 steps for c4)."""
    idx, wl, warmup, steps = args
    torch, torchdiffeq, gp, langevin, hamiltonian, stein = _import_reference()
    import contextlib
    import io
    import warnings
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)
    torch.manual_seed(100 + idx)
    out = {}
    kind = wl["kind"]
    if kind in ("npde_sgld", "npde_psgld", "npde_svgd", "npde_hamcmc"):
        N, x0, t, X, Y = _data(torch, torchdiffeq, gp, wl["T"])
        for variant, odeint in (("odeint", torchdiffeq.odeint), ("odeint_adjoint", torchdiffeq.odeint_adjoint)):
            kreg = _npde_model(torch, gp, Y, t, wl["M"], wl.get("ell", 0.75), idx)
            params = [kreg.U, kreg.logsn]                                                   # gp.py:337
            closure = _npde_closure(torch, kreg, odeint, x0, t, Y)
            it = [0]
            with contextlib.redirect_stdout(io.StringIO()):
                if kind == "npde_psgld":
                    smp = langevin.pSGLD(params, lr0=5e-3, lr_gamma=0.51, lr_t0=100, lr_alpha=0.1, lambda_=1e-8, alpha=0.99, N=N)
                elif kind == "npde_hamcmc":
                    smp = langevin.HAMCMC(params, memory=5, lr0=1e-7, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3, H_gamma=1.0, trust_reg=1.0)
                else:
                    smp = langevin.SGLD(params, lr0=1e-4, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03)

            def one():
                smp.zero_grad()
                loss = closure()
                if kind == "npde_psgld":
                    loss = loss / N                                                         # langevin.py:528
                smp.loss = loss
                loss.backward()
                lr = smp.get_lr(it[0])
                with contextlib.redirect_stdout(io.StringIO()):
                    if kind == "npde_svgd":
                        pass                                                                # the particle update is the phi block's
                    elif kind == "npde_hamcmc":
                        M_ = smp.memory
                        if it[0] < 2 * M_ - 1:
                            smp.step_without_metric(lr=lr, add_noise=True, add_params=True)
                        else:
                            smp.step(lr=lr, add_noise=True)
                    else:
                        smp.step(lr=lr)
                it[0] += 1
            w = warmup if kind != "npde_hamcmc" else max(warmup, 2 * 6 - 1)                  # time METRIC steps (window full)
            out[variant] = _time_loop(one, w, steps)
    elif kind == "mlp_asghmc":
        N, x0, t, X, Y = _data(torch, torchdiffeq, gp, wl["T"])
        H = wl["H"]

        class NN(torch.nn.Module):                                                          # nn.ipynb cell 4
            def __init__(self, input_size, hidden_size=10):
                super().__init__()
                nn = torch.nn
                self.layers = nn.Sequential(nn.Linear(input_size, hidden_size), nn.ELU(), nn.Linear(hidden_size, hidden_size), nn.ELU(),
                                            nn.Linear(hidden_size, input_size))

            def forward(self, t, x):                                                         # the notebook flattens one row: x [2];
                return self.layers(x)                                                        # row-wise on y0 [N, 2] = the batched call

        class Counted(torch.nn.Module):
            def __init__(self, inner):
                super().__init__()
                self.inner, self.nfe = inner, 0

            def forward(self, t, x):
                self.nfe += 1
                return self.inner(t, x)
        for variant, odeint in (("odeint_adjoint", torchdiffeq.odeint_adjoint),):            # autograd through the dopri5 controller is
            net = NN(2, H)                                                                   # not a usable gradient (SURVEY hard part 6)
            for m_ in net.modules():
                if isinstance(m_, torch.nn.Linear):
                    torch.nn.init.uniform_(m_.weight, a=-0.5, b=0.5)                          # nn.ipynb cell 4
            with torch.no_grad():
                for q in net.parameters():
                    q.mul_(wl["init_scale"])                                                 # same scaling as the GPU arm (bench.py --init-scale)
            cnet = Counted(net)
            params = list(net.parameters())
            smp = hamiltonian.aSGHMC(params, lr=1e-2, mom_decay=5e-2, lambda_=1e-5)
            fwd_nfe = [0]

            def one():
                smp.zero_grad()
                loss = 0
                cnet.nfe = 0
                # BASELINE config 4 "dopri5 batched-step": ONE odeint call for the chain's N trajectories (one controller, error
                # pooled over all N x 2 elements, misc.py:146-157); nn.ipynb cell 10 loops over the rows instead
                xode = odeint(cnet, x0, t, rtol=wl["rtol"], atol=wl["atol"], method="dopri5").permute(1, 0, 2)
                loss = loss + torch.sum((X - xode) ** 2)
                fwd_nfe[0] = cnet.nfe
                loss = loss + 0.5 * sum(torch.sum(q ** 2) for q in params)
                loss.backward()
                smp.step(lr=1e-2, burn_in=True)
            out[variant] = _time_loop(one, warmup, steps)
            out["attempted_steps_per_solve"] = (fwd_nfe[0] - 2) / 6.0                        # forward solve: 2 + 6 * attempted steps
    else:
        raise ValueError(kind)
    return out


def phi_block_seconds(P, d, rows, threads):
    """stein.py:18-34 + 75-86: K = RBFKernel()(X_rows, X_all) with the median heuristic (np.median on the host, :25-26), then
    phi = (K S + grad_K) / n with grad_K by autograd exactly as SVGD.phi forms it.  All host threads (ATen intra-op)."""
    torch, torchdiffeq, gp, langevin, hamiltonian, stein = _import_reference()
    from torch import autograd
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(3)
    Xall = torch.randn(P, d, generator=g)
    S = torch.randn(P, d, generator=g)
    kern = stein.RBFKernel()

    def block():
        X = Xall[:rows].detach().requires_grad_(True)
        K = kern(X, Xall.detach())
        grad_K = -autograd.grad(K.sum(), X)[0]
        return (K.detach().matmul(S) + grad_K) / P
    block()
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        block()
    return (time.perf_counter() - t0) / n


WORKLOADS = {
    "c1": dict(kind="npde_sgld", M=5, T=40),
    "c2": dict(kind="npde_psgld", M=5, T=101),
    "c3": dict(kind="npde_svgd", M=5, T=40),
    "c4": dict(kind="mlp_asghmc", H=64, T=40, rtol=1e-5, atol=1e-7, init_scale=0.3),
    "c5": dict(kind="npde_hamcmc", M=16, T=40, ell=0.35),
}


def run(workload, P_total, steps, warmup, cores=None, wl_override=None):
    """Returns the measurement dict of the reference arm for `workload` (see module docstring for the sampling)."""
    import multiprocessing as mp
    wl = dict(WORKLOADS[workload])
    wl.update(wl_override or {})
    C = cores or os.cpu_count() or 1
    t_wall = time.perf_counter()
    ctx = mp.get_context("spawn")
    with ctx.Pool(C) as pool:
        res = pool.map(chain_worker, [(i, wl, warmup, steps) for i in range(C)])
    variants = [k for k in res[0] if k.startswith("odeint")]
    per = {v: sum(r[v] for r in res) / len(res) for v in variants}
    best = min(per, key=per.get)
    out = dict(cores=C, kind="reference", per_process_s_per_iter=per, variant=best,
               sample="%d processes x 1 chain x (%d warm-up + %d timed) sampler iterations, fp64, 1 thread each" % (C, warmup, steps))
    if wl["kind"] == "mlp_asghmc":
        att = sum(r["attempted_steps_per_solve"] for r in res) / len(res)
        out["attempted_steps_per_solve"] = att
        rk_steps = att
    else:
        rk_steps = wl["T"] - 1
    t_solve_all = math.ceil(P_total / C) * per[best]            # chains are not divisible: P < C still takes one chain's time
    t_phi_all = 0.0
    if wl["kind"] == "npde_svgd":
        rows = min(P_total, 256)
        d = 2 * wl["M"] ** 2 + 2
        tb = phi_block_seconds(P_total, d, rows, C)
        t_phi_all = tb * (P_total / rows)
        out["phi_block"] = dict(rows=rows, cols=P_total, seconds=tb)
        out["sample"] += "; + RBFKernel/phi on a %d x %d block (%.3f s, %d threads), scaled by P/rows" % (rows, P_total, tb, C)
    out["rk_steps"] = rk_steps
    out["value"] = P_total * rk_steps / (t_solve_all + t_phi_all)
    out["seconds_per_whole_step"] = t_solve_all + t_phi_all
    out["wall_s"] = round(time.perf_counter() - t_wall, 1)
    return out
