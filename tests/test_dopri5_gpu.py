"""GPU: adaptive dopri5 through the C ABI vs the reference goldens and the float32 oracle -- per-pair controller (one reference call
per trajectory row, tests/golden/dopri5.npz) and pooled controller (one reference call with y0 [N, 2], dopri5_batched.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def _fields(g):
    import bayesian_ode_b200 as bode
    fn = bode.NPDEField(torch.from_numpy(g["U"]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    fm = bode.MLPField(1, hidden_size=20, theta=torch.from_numpy(g["theta"])[None])
    return fn, fm


@pytest.mark.parametrize("case,kw", [("default", {}), ("loose", dict(rtol=1e-5, atol=1e-7)),
                                     ("firststep", dict(rtol=1e-5, atol=1e-7, method="dopri5", options=dict(first_step=0.5)))])
def test_dopri5_matches_reference(case, kw):
    """fp32 state vs the float64 reference: trajectories to 1e-5 relative (the north_star bar) at the reference's own
    tolerances; the number of attempted steps stays within 25% of the reference's (the controller sees fp32 rounding,
    and one flipped borderline accept changes every later step size)."""
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5")
    fn, fm = _fields(g)
    x0, t = torch.from_numpy(g["x0"]), torch.from_numpy(g["t"])
    kw = dict(kw, method="dopri5", options=dict(kw.get("options", {}), controller="pair"))   # the fixture: one call per row
    for fname, f in (("npde", fn), ("mlp", fm)):
        sol = bode.odeint(f, x0, t, **kw)
        sol = sol if sol.dim() == 3 else sol[:, 0]
        tol = 1e-5 if case == "default" else 1e-4      # the solver itself only controls the error to rtol
        assert relerr(sol.detach().cpu().numpy(), g[f"{case}_{fname}_sol"]) < tol, (fname, case)
        st = bode.last_dopri5_stats().cpu().numpy().reshape(-1, 3)
        attempted_ref = (g[f"{case}_{fname}_nfe"] - (1 if case == "firststep" else 2)) / 6.0
        assert np.all(st[:, 2] == 0)
        assert np.all(np.abs(st[:, 0] + st[:, 1] - attempted_ref) <= np.maximum(3, 0.25 * attempted_ref)), (st, attempted_ref)


def test_dopri5_control_flow_matches_float32_oracle():
    """Same dtype on both sides (float32 state, float64 t/dt): accepted/rejected counts agree with the NumPy oracle for
    most rows -- a borderline accept can flip on the last ulp because the device field uses FMA contraction and
    MUFU.EX2 where NumPy does not (allowed: 2 rows of 5) -- and the trajectories agree to the solver tolerance."""
    import bayesian_ode_b200 as bode
    from oracle import dopri5, npde
    g = load_golden("dopri5")
    fn, _ = _fields(g)
    fo = npde.NPDEField(g["U"][None], g["Z"], 1.0, 0.75)
    sol = bode.odeint(fn, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), rtol=1e-5, atol=1e-7, method="dopri5",
                      options=dict(controller="pair"))
    st = bode.last_dopri5_stats().cpu().numpy().reshape(-1, 3)
    same = 0
    for r in range(5):
        f32 = lambda y: fo.f(y[None, None].astype(np.float64))[0, 0]
        so, so_st = dopri5.odeint_dopri5(f32, g["x0"][r].astype(np.float32), g["t"], rtol=1e-5, atol=1e-7)
        same += int(st[r, 0] == so_st["accepted"] and st[r, 1] == so_st["rejected"])
        assert relerr(sol[:, r].detach().cpu().numpy(), so) < 1e-4
    assert same >= 3, st


def test_dopri5_reversed_time_and_errors():
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5")
    fn, _ = _fields(g)
    x0 = torch.from_numpy(g["x0"])
    sol = bode.odeint(fn, x0, torch.from_numpy(g["rev_t"]), rtol=1e-5, atol=1e-7, method="dopri5", options=dict(controller="pair"))
    assert relerr(sol.detach().cpu().numpy(), g["rev_npde_sol"]) < 1e-4
    with pytest.raises(AssertionError):                      # dopri5.py:89 max_num_steps
        bode.odeint(fn, x0, torch.from_numpy(g["t"]), method="dopri5", options=dict(max_num_steps=2))
    with pytest.warns(UserWarning):                          # misc.py:79-81
        bode.odeint(fn, x0, torch.from_numpy(g["t"]), rtol=1e-4, atol=1e-6, method="dopri5", options=dict(bogus=1))


def test_dopri5_config4_sizes_run():
    """BASELINE config 4 shape: 2-64-64-2 MLP, many chains, per-pair adaptive stepping, finite output and sane step counts."""
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5")
    f = bode.MLPField(512, hidden_size=64, generator=torch.Generator().manual_seed(0))
    sol = bode.odeint(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), rtol=1e-5, atol=1e-7, method="dopri5",
                      options=dict(controller="pair"))
    assert sol.shape == (40, 512, 5, 2) and bool(torch.isfinite(sol.detach()).all())
    st = bode.last_dopri5_stats()
    assert int(st[..., 2].max()) == 0 and int(st[..., 0].min()) >= 5


def test_dopri5_zero_field_takes_the_ratio_zero_branch():
    """U = 0 -> f = 0 exactly on both sides: err = 0, ratio == 0 -> dt *= ifactor (misc.py:162-163), h0 = 1e-6 branch of
    the initial step (misc.py:119-120); arithmetic is exact so the step counts must equal the oracle's bit for bit."""
    import bayesian_ode_b200 as bode
    from oracle import dopri5
    g = load_golden("dopri5")
    fn = bode.NPDEField(torch.zeros(25, 2), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    sol = bode.odeint(fn, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), method="dopri5", options=dict(controller="pair"))
    st = bode.last_dopri5_stats().cpu().numpy().reshape(-1, 3)
    so, so_st = dopri5.odeint_dopri5(lambda y: np.zeros_like(y), g["x0"][0].astype(np.float32), g["t"])
    assert np.all(st[:, 0] == so_st["accepted"]) and np.all(st[:, 1] == so_st["rejected"])
    assert np.array_equal(sol[:, 0].detach().cpu().numpy(), so)


def test_dopri5_gradient_mlp_matches_reference_adjoint():
    """bayesian_closure through dopri5 (rtol 1e-7 / atol 1e-9): loss and gradient vs the reference's odeint_adjoint(dopri5).
    The reference's own adjoint-vs-backprop cross-check bars are 3e-4 .. 2e-3 (gradient_tests.py:114-116)."""
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5")
    f = bode.MLPField(1, hidden_size=20, theta=torch.from_numpy(g["theta"])[None])
    post = bode.MLPPosterior(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["X"]), method="dopri5", reg=0.5)
    loss, gth, _ = post.loss_and_grad_()
    assert relerr(loss.cpu().numpy(), g["mlp_loss"][None]) < 1e-4
    assert relerr(post.sqerr.cpu().numpy(), g["mlp_sqerr"][None]) < 1e-4
    assert relerr(gth.cpu().numpy()[0], g["mlp_grad_adjoint"]) < 2e-3


def test_dopri5_gradient_npde_matches_reference_adjoint():
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5")
    f = bode.NPDEField(torch.from_numpy(g["U"]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["Y"]), method="dopri5",
                              options=dict(controller="pair"))          # the fixture sums one odeint_adjoint call per row
    loss, gU, gl = post.loss_and_grad_()
    assert relerr(loss.cpu().numpy(), g["npde_loss"][None]) < 1e-4
    assert relerr(gU.cpu().numpy()[0], g["npde_gU_adjoint"]) < 2e-3
    assert relerr(gl.cpu().numpy()[0], g["npde_glogsn_adjoint"]) < 1e-3


def test_dopri5_autograd_matches_frozen_step_oracle():
    """odeint(method='dopri5') + a torch loss + backward(): the CUDA reverse sweep vs the float64 oracle of the same
    definition (accepted steps frozen), one trajectory row at a time."""
    import bayesian_ode_b200 as bode
    from oracle import dopri5, mlp
    g = load_golden("dopri5")
    f = bode.MLPField(1, hidden_size=20, theta=torch.from_numpy(g["theta"])[None])
    rng = np.random.default_rng(4)
    w = rng.standard_normal((40, 1, 5, 2))
    x0 = torch.from_numpy(g["x0"]).cuda().float().requires_grad_(True)
    sol = bode.odeint(f, x0, torch.from_numpy(g["t"]), rtol=1e-6, atol=1e-8, method="dopri5", options=dict(controller="pair"))
    (sol * torch.from_numpy(w).cuda().float()).sum().backward()
    got = torch.cat([p.grad.reshape(1, -1) for p in f.parameters()], 1)[0].cpu().numpy()
    fm = mlp.MLPField(g["theta"][None], 20)

    class F:
        def f(self, y): return fm.f(y[None, None])[0, 0]
        def vjp(self, y, a):
            j, gg = fm.vjp(y[None, None], a[None, None])
            return j[0, 0], gg[0]
        def zero_grad(self): return np.zeros(522)
        add_grad = staticmethod(lambda a, b: a + b)
    want, wx0 = np.zeros(522), []
    for r in range(5):
        _, gy0, gth = dopri5.solve_and_grad(F(), g["x0"][r], g["t"].astype(np.float64), lambda s, r=r: w[:, 0, r], rtol=1e-6, atol=1e-8)
        want += gth
        wx0.append(gy0)
    assert relerr(got, want) < 2e-3
    assert relerr(x0.grad.cpu().numpy(), np.stack(wx0)) < 2e-3


def test_dopri5_config4_step_runs_with_asghmc():
    """BASELINE config 4 in miniature: 2-64-64-2 MLP chains, dopri5 fwd+grad, aSGHMC update on the flat buffer."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200.samplers import aSGHMC
    g = load_golden("dopri5")
    f = bode.MLPField(64, hidden_size=64, generator=torch.Generator().manual_seed(0))
    post = bode.MLPPosterior(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["X"]), method="dopri5",
                             rtol=1e-5, atol=1e-7)
    smp = aSGHMC(list(f.parameters()), lr=1e-4, mom_decay=5e-2)
    th0 = f.theta.clone()
    smp.sample(post, num_samples=1, burn_in=2, print_iters=False)
    assert bool(torch.isfinite(f.theta).all()) and not torch.equal(th0, f.theta)
    st = bode.last_dopri5_stats()
    assert int(st[..., 2].max()) == 0


# ---------------------------------------------------------------------------------------------------- pooled controller
@pytest.mark.parametrize("case,kw", [("default", {}), ("loose", dict(rtol=1e-5, atol=1e-7)),
                                     ("firststep", dict(rtol=1e-5, atol=1e-7, method="dopri5", options=dict(first_step=0.5)))])
def test_dopri5_batched_controller_matches_reference_call(case, kw):
    """odeint(f, y0[N, 2], t) as the reference runs it: ONE controller, error pooled over all N x 2 elements (misc.py:146-157;
    SURVEY.md A.8 quirk 4) -- the default of bode.odeint.  Fixture: the reference's own batched call with its accept / reject
    sequence recorded (tests/golden/dopri5_batched.npz)."""
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5_batched")
    fn = bode.NPDEField(torch.from_numpy(g["U"]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    fm = bode.MLPField(1, hidden_size=20, theta=torch.from_numpy(g["theta"])[None])
    x0, t = torch.from_numpy(g["x0"]), torch.from_numpy(g["t"])
    from oracle import dopri5 as od5, mlp as omlp, npde as onpde
    fo, mo = onpde.NPDEField(g["U"][None], g["Z"], 1.0, 0.75), omlp.MLPField(g["theta"][None], 20)
    okw = {k: v for k, v in kw.items() if k in ("rtol", "atol")}
    if "options" in kw:
        okw["first_step"] = kw["options"]["first_step"]
    for fname, f, of in (("npde", fn, lambda y: fo.f(y[None].astype(np.float64))[0]), ("mlp", fm, lambda y: mo.f(y[None].astype(np.float64))[0])):
        sol = bode.odeint(f, x0, t, **kw)                        # method=None -> dopri5 (odeint.py:68-69), controller "batch"
        sol = sol if sol.dim() == 3 else sol[:, 0]
        st = bode.last_dopri5_stats().cpu().numpy().reshape(-1, 3)
        assert np.all(st == st[0]) and st[0, 2] == 0            # every pair of the particle reports the particle's controller
        acc = g[f"{case}_{fname}_accept"]
        ref = (int(acc.sum()), int(len(acc) - acc.sum()))
        got = (int(st[0, 0]), int(st[0, 1]))
        err = relerr(sol.detach().cpu().numpy(), g[f"{case}_{fname}_sol"])
        print("dopri5 batched %s/%s: accepted,rejected = %s (reference %s), trajectory err %.2e" % (case, fname, got, ref, err))
        # Selection logic: against the oracle in the SAME arithmetic (float32 state, float64 t / dt; its float64 run reproduces the
        # reference's sequence exactly, test_oracle_golden).  At these tolerances the fp32 error estimate sits at rounding-noise
        # level, so single decisions differ from the float64 reference; the kernel must follow the float32 oracle, give or take one
        # borderline decision (FMA contraction and MUFU.EX2 differ from NumPy in the last ulp).
        _, so = od5.odeint_dopri5(of, g["x0"].astype(np.float32), g["t"], **okw)
        print("   float32 oracle: accepted,rejected = (%d, %d)" % (so["accepted"], so["rejected"]))
        assert abs(got[0] - so["accepted"]) <= 1 and abs(got[1] - so["rejected"]) <= 1, (case, fname, got, so)
        assert abs(sum(got) - sum(ref)) <= 3, (case, fname, got, ref)
        # trajectories: the solver's own bar -- 1e-5 at the default tolerances (rtol 1e-7), 10 rtol otherwise
        assert err < (1e-5 if case == "default" else 1e-4), (case, fname, err)


def test_dopri5_batched_and_pair_controllers_differ_and_validate():
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5_batched")
    fn = bode.NPDEField(torch.from_numpy(g["U"]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    x0, t = torch.from_numpy(g["x0"]), torch.from_numpy(g["t"])
    bode.odeint(fn, x0, t, rtol=1e-5, atol=1e-7, method="dopri5", options=dict(controller="pair"))
    sp = bode.last_dopri5_stats().cpu().numpy().reshape(-1, 3)
    bode.odeint(fn, x0, t, rtol=1e-5, atol=1e-7, method="dopri5", options=dict(controller="batch"))
    sb = bode.last_dopri5_stats().cpu().numpy().reshape(-1, 3)
    assert not np.all(sp == sp[0]) and np.all(sb == sb[0])       # rows step independently vs in lock-step
    with pytest.raises(ValueError):
        bode.odeint(fn, x0, t, method="dopri5", options=dict(controller="row"))


def test_dopri5_batched_gradients_match_reference_adjoint():
    """The closures through ONE batched odeint_adjoint(dopri5) call of the reference (rtol 1e-7 / atol 1e-9): loss and gradient.
    NPDEPosterior defaults to the pooled controller (gp.py:346 integrates y0 [N, 2] in one call); the MLP closure is asked for it."""
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5_batched")
    f = bode.NPDEField(torch.from_numpy(g["U"]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["Y"]), method="dopri5")
    loss, gU, gl = post.loss_and_grad_()
    st = bode.last_dopri5_stats().cpu().numpy().reshape(-1, 3)
    assert np.all(st == st[0])
    assert relerr(loss.cpu().numpy(), g["npde_loss"][None]) < 1e-4
    assert relerr(gU.cpu().numpy()[0], g["npde_gU_adjoint"]) < 2e-3
    assert relerr(gl.cpu().numpy()[0], g["npde_glogsn_adjoint"]) < 1e-3
    fm = bode.MLPField(1, hidden_size=20, theta=torch.from_numpy(g["theta"])[None])
    pm = bode.MLPPosterior(fm, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["X"]), method="dopri5", reg=0.5,
                           options=dict(controller="batch"))
    loss, gth, _ = pm.loss_and_grad_()
    assert relerr(loss.cpu().numpy(), g["mlp_loss"][None]) < 1e-4
    assert relerr(pm.sqerr.cpu().numpy(), g["mlp_sqerr"][None]) < 1e-4
    assert relerr(gth.cpu().numpy()[0], g["mlp_grad_adjoint"]) < 2e-3


def test_dopri5_batched_many_particles_lockstep_groups():
    """P particles x N rows: every particle keeps its own pooled controller (particles of one warp / CTA take different numbers of
    attempts); particle p of the batch must reproduce a single-particle launch of its own parameters bit for bit."""
    import bayesian_ode_b200 as bode
    g = load_golden("dopri5_batched")
    rng = np.random.default_rng(17)
    P = 13
    U = g["U"][None] + 0.2 * rng.standard_normal((P, 25, 2))
    x0, t = torch.from_numpy(g["x0"]), torch.from_numpy(g["t"])
    f = bode.NPDEField(torch.from_numpy(U), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    with torch.no_grad():
        sol = bode.odeint(f, x0, t, rtol=1e-5, atol=1e-7)
    st = bode.last_dopri5_stats().cpu().numpy().reshape(P, 5, 3)
    assert np.all(st == st[:, :1]) and len({tuple(r) for r in st[:, 0, :2]}) > 1
    for p in (0, 6, 12):
        f1 = bode.NPDEField(torch.from_numpy(U[p:p + 1]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
        with torch.no_grad():
            s1 = bode.odeint(f1, x0, t, rtol=1e-5, atol=1e-7)
        assert torch.equal(s1.reshape(40, 5, 2), sol[:, p])
    th = torch.from_numpy(g["theta"])[None] * (1.0 + 0.1 * torch.from_numpy(rng.standard_normal((P, 522))))
    fm = bode.MLPField(P, hidden_size=20, theta=th)
    with torch.no_grad():
        solm = bode.odeint(fm, x0, t, rtol=1e-5, atol=1e-7)
    stm = bode.last_dopri5_stats().cpu().numpy().reshape(P, 5, 3)
    assert np.all(stm == stm[:, :1]) and int(stm[..., 2].max()) == 0 and bool(torch.isfinite(solm).all())
