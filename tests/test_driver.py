"""Driver compatibility: the reference's hyper-parameter JSON and data-pickle formats (gp.py:529-564,
gp.py:310) and ``run_sampler`` (gp.py:290-391)."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

# scripts/vanderpol/json/10000.json of the reference, verbatim
REF_JSON = ('{"output": "exp/vanderpol/gp/", "data": {"pickle_file": "data/vdp.pickle"}, "configs": [{"inf_type": "optim", "method": "Adam", '
            '"M": 6, "sf": 1, "ell": 0.75, "lr": 0.005, "num_iters": 50}, {"inf_type": "samplers", "method": "aSGHMC", "M": 6, "sf": 1, '
            '"ell": 0.75, "lr": 0.005, "burn_in": 50, "num_samples": 100, "chain_start": 10, "thinning": 2}]}')


def test_reference_json_and_pickle_formats_round_trip(tmp_path):
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    (tmp_path / "10000.json").write_text(REF_JSON)
    hyp = bode.driver.load_hyperparameters(str(tmp_path), 10000)
    assert [c["method"] for c in hyp["configs"]] == ["Adam", "aSGHMC"] and all(c["id"] == "10000" for c in hyp["configs"])
    data = problems.make_dataset("VDP", seed=0)
    bode.driver.save_data(data, str(tmp_path / "vdp.pickle"))
    raw = pickle.load(open(tmp_path / "vdp.pickle", "rb"))
    assert tuple(raw.keys()) == ("N", "R", "noise", "x0", "t", "X", "Y", "ODE")      # gp.py:310 unpacks .values() in this order
    back = bode.driver.load_data(str(tmp_path / "vdp.pickle"))
    assert back["ODE"] == "VDP" and np.array_equal(back["Y"], data["Y"])
    with pytest.raises(NotImplementedError):
        bode.driver.worker(hyp["configs"][0], data, str(tmp_path))                     # optimiser baselines: out of scope


@pytest.mark.gpu
def test_run_sampler_on_reference_config(tmp_path):
    """The aSGHMC entry of the reference's own JSON (6x6 grid), shortened: chain bookkeeping, loss arrays and output files."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    cfg = json.loads(REF_JSON)["configs"][1]
    cfg.update(burn_in=5, num_samples=12, chain_start=2, thinning=3, id="10000", lr=1e-4)
    data = problems.make_dataset("VDP", seed=0)
    out = bode.driver.worker(cfg, data, str(tmp_path))
    assert len(out["chain"]) == 12 and len(out["chain_"]) == len(range(2, 12, 3))
    assert len(out["total_loss_arr"]) == 17 and all(np.isfinite(v) for v in out["total_loss_arr"])
    params, acc = out["chain_"][0]
    assert params[0][0].shape == (1, 36, 2) and acc is True
    d = os.path.join(str(tmp_path), "VDP", "samplers", "aSGHMC", "10000")
    assert sorted(os.listdir(d)) == ["10000.json", "sq_err_loss_arr.pickle", "total_loss_arr.pickle"]
    assert pickle.load(open(os.path.join(d, "sq_err_loss_arr.pickle"), "rb")) == out["sq_err_loss_arr"]


@pytest.mark.gpu
@pytest.mark.parametrize("method,extra", [("SGLD", dict(lr0=1e-4, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03)),
                                          ("pSGLD", dict(lr0=5e-3, lr_gamma=0.51, lr_t0=100, lr_alpha=0.1, lambda_=1e-8, psgld_alpha=0.99)),
                                          ("MALA", dict(lr=1e-6)), ("SVGD", dict(lr=1e-4))])
def test_run_sampler_batched_chains(method, extra):
    """gen_configs.py's sampler grid (gp.py:363-378) on 64 chains at once: the first SGLD losses equal the single-chain run's."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    cfg = dict(inf_type="samplers", method=method, M=5, sf=1, ell=0.75, burn_in=2, num_samples=4, chain_start=0, thinning=1, **extra)
    data = problems.make_dataset("VDP", seed=0)
    out = bode.driver.run_sampler(cfg, data, chains=64, jitter=0.0)
    assert len(out["chain"]) == 4 and np.asarray(out["total_loss_arr"]).shape == (6, 64)
    one = bode.driver.run_sampler(cfg, data)
    assert abs(out["total_loss_arr"][0][0] - one["total_loss_arr"][0]) <= 1e-5 * abs(one["total_loss_arr"][0])
