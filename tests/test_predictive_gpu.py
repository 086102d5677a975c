"""GPU: posterior-predictive ensembles vs the reference's serial re-integration loop (gp.py:440-464)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def _field(g):
    import bayesian_ode_b200 as bode
    d = load_golden("npde_m5")
    return bode.NPDEField(torch.from_numpy(g["U"][0]), torch.from_numpy(d["Z"]), 1.0, 0.75, 0.1)


@pytest.mark.parametrize("name,method,tol", [("rk4", "rk4", 2e-5), ("dopri5", None, 5e-4)])
def test_posterior_predictive_matches_reference_loop(name, method, tol):
    """One batched solve over all stored samples == the reference's per-sample odeint calls; mean / np.std(ddof=0) per
    trajectory and component.  rk4: same fixed grid, fp32 vs fp64 (2e-5).  dopri5: the reference runs one controller for the
    whole [N, 2] batch at rtol 1e-6, this build one per trajectory -- both solve to tolerance, compared at 5e-4."""
    import bayesian_ode_b200 as bode
    g = load_golden("predictive")
    f = _field(g)
    chain = [([[g["U"][i], np.log(0.1) * np.ones(2)]], True) for i in range(g["U"].shape[0])]      # reference chain format
    mean, std, traj = bode.posterior_predictive(f, chain, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), method=method,
                                                return_trajectories=True)
    assert traj.shape == (6, 4, 30, 2)
    assert relerr(traj.cpu().numpy(), g[f"{name}_traj"]) < tol
    assert relerr(mean, g[f"{name}_mean"]) < tol
    assert np.abs(std - g[f"{name}_std"]).max() < tol * np.abs(g[f"{name}_mean"]).max()


def test_posterior_predictive_from_device_chain_store():
    """The on-device ChainStore of a sampler run feeds the ensemble without a host round trip, chunked or not."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    from bayesian_ode_b200.samplers import SGLD
    data = problems.make_dataset(seed=0)
    Z = problems.inducing_grid(data["Y"], 5)
    U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
    f = bode.NPDEField(U0[None].repeat(16, 1, 1), Z, 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
    smp = SGLD([f.U, f.logsn], lr0=1e-5, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03, seed=3)
    chain = smp.sample(post, num_samples=5, burn_in=2)
    x0 = torch.tensor([[1.0, 0.5], [-1.5, 1.0]])
    t = torch.linspace(0., 10., 25)
    m1, s1 = bode.posterior_predictive(f, chain, x0, t, method="rk4")
    m2, s2 = bode.posterior_predictive(f, chain, x0, t, method="rk4", chunk=32)
    assert m1.shape == (2, 25, 2) and np.isfinite(m1).all() and (s1 > 0).any()
    assert np.allclose(m1, m2, rtol=0, atol=1e-12) and np.allclose(s1, s2, rtol=0, atol=1e-9)
    # same numbers from the materialised reference-format entries
    m3, s3 = bode.posterior_predictive(f, list(chain), x0, t, method="rk4")
    assert np.allclose(m1, m3, rtol=0, atol=1e-12)
