"""GPU: MLP neural-ODE field (nn.ipynb cell 4/10) through the C ABI vs the reference goldens."""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def _field(g, H):
    import bayesian_ode_b200 as bode
    th = g[f"h{H}_theta"]
    return bode.MLPField(th.shape[0], hidden_size=H, theta=torch.from_numpy(th))


@pytest.mark.parametrize("H", [20, 64])
def test_mlp_odeint_rk4_matches_reference(H):
    import bayesian_ode_b200 as bode
    g = load_golden("mlp")
    f = _field(g, H)
    with torch.no_grad():
        sol = bode.odeint(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), method="rk4")
    assert sol.shape == g[f"h{H}_sol"].shape
    assert relerr(sol.cpu().numpy(), g[f"h{H}_sol"]) < 1e-5


@pytest.mark.parametrize("H", [20, 64])
@pytest.mark.parametrize("mode", ["discrete", "adjoint"])
def test_mlp_fused_closure_matches_reference(H, mode):
    import bayesian_ode_b200 as bode
    g = load_golden("mlp")
    f = _field(g, H)
    post = bode.MLPPosterior(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["X"]), grad_mode=mode, reg=0.5)
    loss, gth, _ = post.loss_and_grad_()
    torch.cuda.synchronize()
    assert relerr(loss.cpu().numpy(), g[f"h{H}_loss"]) < 1e-5
    assert relerr(post.sqerr.cpu().numpy(), g[f"h{H}_sqerr"]) < 1e-5
    ref = g[f"h{H}_g_{mode}"]
    err = np.abs(gth.cpu().numpy() - ref).max(axis=1) / np.abs(ref).max(axis=1)       # per chain
    assert err.max() < 1e-4, err


def test_mlp_autograd_protocol_and_sampler():
    """closure() -> backward() -> p.grad for the six parameter views; aSGHMC (nn.ipynb cell 11) runs on the flat buffer."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200.samplers import aSGHMC
    g = load_golden("mlp")
    f = _field(g, 20)
    x0, t, X = (torch.from_numpy(g[k]) for k in ("x0", "t", "X"))
    params = list(f.parameters())
    assert [tuple(p.shape) for p in params] == [(3, 20, 2), (3, 20), (3, 20, 20), (3, 20), (3, 2, 20), (3, 2)]
    # generic loss through odeint + torch ops
    xode = bode.odeint(f, x0, t, method="rk4").permute(1, 2, 0, 3)
    loss = ((X.cuda().float()[None] - xode) ** 2).sum() + 0.5 * sum((p ** 2).sum() for p in params)
    loss.backward()
    gflat = torch.cat([p.grad.reshape(3, -1) for p in params], 1)
    assert relerr(gflat.cpu().numpy(), g["h20_g_discrete"]) < 1e-4
    for p in params:
        p.grad = None
    post = bode.MLPPosterior(f, x0, t, X)
    smp = aSGHMC(params, lr=1e-4, mom_decay=5e-2)
    assert smp._flat is not None and smp._flat.shape == (3, 522)
    th0 = f.theta.clone()
    chain = smp.sample(post, num_samples=2, burn_in=2, print_iters=False)
    assert len(chain) == 2 and not torch.equal(th0, f.theta) and bool(torch.isfinite(f.theta).all())
