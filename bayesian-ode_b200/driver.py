"""Experiment-driver compatibility: run the reference's JSON configurations on the B200 samplers.

Mirrors scripts/vanderpol/gp.py:290-391 (``run_sampler``) and :504-564 (``worker`` / the ``__main__`` loader) without the
plotting: same hyper-parameter JSON (``{"output", "data": {"pickle_file"}, "configs": [...]}``, gen_configs.py / json/10000.json),
same data pickle layout (dict ``N, R, noise, x0, t, X, Y, ODE``, gp.py:310), same inducing grid, gradient-matching init,
posterior closure, sampler constructors per ``config['method']``, loss arrays and ``chain[chain_start::thinning]``.
What differs by design: ``chains`` > 1 runs that many independent chains (or SVGD particles) of the same configuration in the
same launches, and the chain stays on the device until it is read.
"""
import json
import os
import pickle

import numpy as np
import torch

from . import problems
from .fields import NPDEField
from .posterior import NPDEPosterior
from . import samplers as S

DATA_KEYS = ("N", "R", "noise", "x0", "t", "X", "Y", "ODE")


def load_hyperparameters(json_dir, job_id):
    """gp.py:544: ``json.load(open(os.path.join(args.json_dir, str(args.id) + '.json')))``."""
    with open(os.path.join(json_dir, str(job_id) + ".json")) as f:
        hyp = json.load(f)
    for cfg in hyp["configs"]:
        cfg.setdefault("id", str(job_id))                                   # gp.py:558-559
    return hyp


def load_data(pickle_file):
    """gp.py:548 -- the pickle holds a dict with the keys of ``DATA_KEYS`` in that order (gp.py:310 unpacks ``.values()``)."""
    with open(pickle_file, "rb") as f:
        data = pickle.load(f)
    missing = [k for k in DATA_KEYS if k not in data]
    if missing:
        raise KeyError("data pickle lacks %s" % missing)
    return data


def save_data(data, pickle_file):
    with open(pickle_file, "wb") as f:
        pickle.dump({k: data[k] for k in DATA_KEYS}, f)


def _make_sampler(config, params, N):
    m = config["method"]
    if m == "MALA":                                                          # gp.py:363-366
        return S.MALA(params, lr=config["lr"], add_noise=True), {}
    if m == "SGLD":                                                          # gp.py:367-370
        return S.SGLD(params, lr0=config["lr0"], lr_gamma=config["lr_gamma"], lr_t0=config["lr_t0"], lr_alpha=config["lr_alpha"]), {}
    if m == "pSGLD":                                                         # gp.py:371-375
        return S.pSGLD(params, lr0=config["lr0"], lr_gamma=config["lr_gamma"], lr_t0=config["lr_t0"], lr_alpha=config["lr_alpha"],
                       lambda_=config["lambda_"], alpha=config["psgld_alpha"], N=N), {}
    if m == "aSGHMC":                                                        # gp.py:376-378
        return S.aSGHMC(params, lr=config["lr"], add_noise=True), {}
    # samplers of the reference's samplers/ package that its driver does not dispatch (constructor kwargs from the config)
    if m == "cSGLD":
        return S.cSGLD(params, lr0=config["lr0"], M=config.get("cycles", 5), beta=config.get("beta", 0.25)), {}
    if m == "acSGHMC":
        return S.acSGHMC(params, lr0=config["lr0"], M=config.get("cycles", 5), beta=config.get("beta", 0.25)), {}
    if m == "HAMCMC":
        return S.HAMCMC(params, memory=config.get("memory", 5), lr0=config["lr0"], lr_gamma=config["lr_gamma"], lr_t0=config["lr_t0"],
                        lr_alpha=config["lr_alpha"]), {"print_iters": False}
    if m == "SVGD":
        return S.SVGD(params, lr=config.get("lr", 1e-4)), {}
    raise KeyError("unknown sampler method %r" % m)


def run_sampler(config, data, output=None, chains=1, jitter=0.0, seed=0, device=None):
    """gp.py:290-391.  Returns a dict with ``chain`` (all recorded samples), ``chain_`` (``chain[chain_start::thinning]``, gp.py:381),
    ``total_loss_arr`` / ``sq_err_loss_arr`` (one entry per iteration; scalars for one chain like the reference's ``.item()``,
    arrays of length ``chains`` otherwise), and the ``field`` / ``posterior`` / ``sampler`` objects.  With ``output`` the two loss
    pickles are written where the reference writes them (gp.py:384-387)."""
    M, sf, ell = config["M"], config["sf"], config["ell"]
    N, noise, x0, t, Y = data["N"], data["noise"], data["x0"], data["t"], data["Y"]
    if "noise" in config:                                                    # gp.py:311-312
        noise = config["noise"]
    Z = problems.inducing_grid(Y, M)                                          # gp.py:315-318
    U0 = problems.gradient_matching_init(Y, t, Z, sf, ell)                    # gp.py:323-331
    if chains > 1:
        gen = torch.Generator().manual_seed(seed)
        U0 = U0[None] + jitter * torch.randn(chains, M * M, 2, generator=gen, dtype=torch.float64)
    field = NPDEField(U0, Z, sf, ell, noise, device=device)
    post = NPDEPosterior(field, torch.as_tensor(x0), torch.as_tensor(t), torch.as_tensor(np.asarray(Y)), method="rk4",
                         grad_mode=config.get("grad_mode", "adjoint"))        # gp.py:26 imports odeint_adjoint as odeint
    params = [field.U, field.logsn]                                           # gp.py:337
    total_loss_arr, sq_err_loss_arr = [], []

    def arr_closure(total_loss, sq_err_loss):                                 # gp.py:355-357
        tl, se = total_loss.detach().reshape(-1), sq_err_loss.detach().reshape(-1)
        total_loss_arr.append(tl.item() if tl.numel() == 1 else tl.cpu().numpy())
        sq_err_loss_arr.append(se.item() if se.numel() == 1 else se.cpu().numpy())

    sampler, extra = _make_sampler(config, params, N)
    sampler.seed = int(seed)                                                  # the run's seed keys the in-kernel noise too
    kwargs = dict(burn_in=config["burn_in"], num_samples=config["num_samples"], print_iters=False)
    kwargs.update(extra)
    if config["method"] != "HAMCMC":
        kwargs["arr_closure"] = arr_closure
    chain = sampler.sample(post, **kwargs)
    if isinstance(chain, tuple):                                              # HAMCMC returns (chain, logp_array)
        chain, logp = chain
        total_loss_arr = [(-l).reshape(-1).cpu().numpy() if l.numel() > 1 else float(-l) for l in logp]
    chain_ = chain[config.get("chain_start", 0)::config.get("thinning", 1)]  # gp.py:381
    if output is not None:
        out_dir = os.path.join(output, config["method"], str(config.get("id", 0)) + config.get("dir_name", ""))   # gp.py:292-296
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, str(config.get("id", 0)) + ".json"), "w") as f:
            json.dump(config, f)
        with open(os.path.join(out_dir, "total_loss_arr.pickle"), "wb") as f:
            pickle.dump(total_loss_arr, f)
        with open(os.path.join(out_dir, "sq_err_loss_arr.pickle"), "wb") as f:
            pickle.dump(sq_err_loss_arr, f)
    return dict(chain=chain, chain_=chain_, total_loss_arr=total_loss_arr, sq_err_loss_arr=sq_err_loss_arr, field=field,
                posterior=post, sampler=sampler)


def worker(config, data, output, **kw):
    """gp.py:504-524: dispatch on ``config['inf_type']``; the optimiser baselines (Adam / L-BFGS) are outside the hot path."""
    output_ = os.path.join(output, data["ODE"])
    if config["inf_type"] == "optim":
        raise NotImplementedError("inf_type 'optim' (Adam / FullBatchLBFGS point estimates, gp.py:74-287) is outside the B200 hot path")
    return run_sampler(config, data, os.path.join(output_, "samplers"), **kw)
