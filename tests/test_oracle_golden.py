"""The oracle restatement vs. fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import npde, solvers


@pytest.mark.parametrize("tag", ["npde_m5", "npde_m3"])
def test_precompute_matches_reference(tag):
    g = load_golden(tag)
    pre = npde.precompute(g["Z"], float(g["sf"]), float(g["ell"]))
    assert relerr(pre["Kzz"], g["Kzz"]) < 1e-14
    assert relerr(pre["Kzzinv"], g["Kzzinv"]) < 1e-10
    assert relerr(pre["KzzinvL"], g["KzzinvL"]) < 1e-10


def test_inducing_grid_and_init_match_reference():
    g = load_golden("npde_m5")
    Z = npde.inducing_grid(g["Y"], 5)
    assert np.array_equal(Z, g["Z"])
    U0 = npde.gradient_matching_init(g["Y"], g["t"].astype(np.float64), Z, 1.0, 0.75)
    assert relerr(U0, g["U0"]) < 1e-9


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
@pytest.mark.parametrize("mode", ["discrete", "adjoint"])
def test_closure_and_gradients_match_reference(method, mode):
    g = load_golden("npde_m5")
    loss, gU, gl, sol = npde.nlp_grad(g["U"], g["logsn"], g["Z"], 1.0, 0.75, g["x0"], g["t"], g["Y"],
                                      method=method, grad_mode=mode)
    assert relerr(sol, g[f"{method}_sol"]) < 1e-12
    assert relerr(loss, g[f"{method}_loss"]) < 1e-12
    assert relerr(gU, g[f"{method}_gU_{mode}"]) < 1e-10
    assert relerr(gl, g[f"{method}_glogsn_{mode}"]) < 1e-12
    sq = npde.nlp(g["U"], g["logsn"], sol, g["Y"], None, add_prior=False)
    assert relerr(sq, g[f"{method}_sqerr"]) < 1e-12


def test_m3_rk4():
    g = load_golden("npde_m3")
    for mode in ("discrete", "adjoint"):
        loss, gU, gl, sol = npde.nlp_grad(g["U"], g["logsn"], g["Z"], 1.0, 0.75, g["x0"], g["t"], g["Y"],
                                          grad_mode=mode)
        assert relerr(sol, g["rk4_sol"]) < 1e-12
        assert relerr(gU, g[f"rk4_gU_{mode}"]) < 1e-10


def test_discrete_and_adjoint_gradients_differ():
    """SURVEY.md hard part 1: the two gradient definitions are different functions."""
    g = load_golden("npde_m5")
    assert relerr(g["rk4_gU_discrete"], g["rk4_gU_adjoint"]) > 1e-6


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_step_size_grid_and_end_of_step_quirk(case):
    g = load_golden("grid_options")
    field = npde.NPDEField(g["U"][None], g["Z"], 1.0, 0.75)
    y0 = g["x0"][None]
    t, h = g[f"{case}_t"], float(g[f"{case}_h"])
    for method in ("euler", "midpoint", "rk4"):
        sol = solvers.odeint_fixed(field, y0, t, method, step_size=h)[:, 0]
        assert relerr(sol, g[f"{case}_{method}_sol"]) < 1e-12
    w = g[f"{case}_w"][:, None]
    gy0, gU = solvers.odeint_fixed_backward(field, y0, t, w, "rk4", step_size=h)
    assert relerr(gU[0], g[f"{case}_rk4_gU"]) < 1e-10
    assert relerr(gy0[0], g[f"{case}_rk4_gx0"]) < 1e-10


def test_reversed_time():
    g = load_golden("grid_options")
    field = npde.NPDEField(g["U"][None], g["Z"], 1.0, 0.75)
    sol = solvers.odeint_fixed(field, g["x0"][None], g["rev_t"], "rk4")[:, 0]
    assert relerr(sol, g["rev_rk4_sol"]) < 1e-12


def test_output_map_is_end_of_step():
    # SURVEY.md A.1 probe: dy/dt = 2t is not autonomous; use the published index logic instead:
    t = np.array([0.0, 0.25, 0.7, 1.0])
    t_, grid, sign = solvers.build_grid(t, np.float64, 0.5)
    assert np.array_equal(grid, [0.0, 0.5, 1.0])
    assert np.array_equal(solvers.output_map(t_, grid), [1, 2, 4])


@pytest.mark.parametrize("H", [20, 64])
@pytest.mark.parametrize("mode", ["discrete", "adjoint"])
def test_mlp_field_closure_and_gradients_match_reference(H, mode):
    from oracle import mlp
    g = load_golden("mlp")
    loss, gr, sq, sol = mlp.sse_grad(g[f"h{H}_theta"], H, g["x0"], g["t"], g["X"], grad_mode=mode)
    assert relerr(sol, g[f"h{H}_sol"]) < 1e-12
    assert relerr(loss, g[f"h{H}_loss"]) < 1e-12
    assert relerr(sq, g[f"h{H}_sqerr"]) < 1e-12
    assert relerr(gr, g[f"h{H}_g_{mode}"]) < 1e-10


def test_torch_port_used_as_cpu_baseline_matches_reference():
    """oracle/ref_torch.py (the timed CPU baseline) computes the same trajectories as the reference."""
    import torch
    from oracle import ref_torch
    g = load_golden("npde_m5")
    kreg = ref_torch.KReg(torch.from_numpy(g["U"][0]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    xode = ref_torch.odeint_rk4(kreg, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]))
    assert relerr(xode.detach().numpy(), g["rk4_sol"][:, 0]) < 1e-12


@pytest.mark.parametrize("case,kw", [("default", {}), ("loose", dict(rtol=1e-5, atol=1e-7)),
                                     ("firststep", dict(rtol=1e-5, atol=1e-7, first_step=0.5))])
def test_dopri5_oracle_matches_reference_trajectories_and_nfe(case, kw):
    """Per-row adaptive solves: identical number of function evaluations (accept/reject + controller logic) and
    trajectories to 1e-8 for the npde and the MLP field; the first_step quirk (dopri5.py:81-82) included."""
    from oracle import dopri5, mlp
    g = load_golden("dopri5")
    fn = npde.NPDEField(g["U"][None], g["Z"], 1.0, 0.75)
    fm = mlp.MLPField(g["theta"][None], 20)
    for fname, fld in (("npde", fn), ("mlp", fm)):
        for r in range(5):
            sol, st = dopri5.odeint_dopri5(lambda y: fld.f(y[None, None])[0, 0], g["x0"][r], g["t"], **kw)
            assert st["nfe"] == g[f"{case}_{fname}_nfe"][r]
            assert relerr(sol, g[f"{case}_{fname}_sol"][:, r]) < 1e-8


def test_dopri5_oracle_reversed_time():
    from oracle import dopri5
    g = load_golden("dopri5")
    fn = npde.NPDEField(g["U"][None], g["Z"], 1.0, 0.75)
    for r in range(5):
        sol, _ = dopri5.odeint_dopri5(lambda y: fn.f(y[None, None])[0, 0], g["x0"][r], g["rev_t"], rtol=1e-5, atol=1e-7)
        assert relerr(sol, g["rev_npde_sol"][:, r]) < 1e-8


def test_oracle_dopri5_pooled_controller_matches_reference_batched_call():
    """oracle/dopri5.py with y0 [N, 2] pools the error ratio and the initial-step norms over the whole tensor like misc.py:116-157;
    fixture: the reference's batched odeint call with its accept / reject sequence recorded (dopri5_batched.npz)."""
    from oracle import dopri5, mlp, npde
    g = load_golden("dopri5_batched")
    fo = npde.NPDEField(g["U"][None], g["Z"], 1.0, 0.75)
    fm = mlp.MLPField(g["theta"][None], 20)
    t = g["t"].astype(np.float64)
    for case, kw in (("default", {}), ("loose", dict(rtol=1e-5, atol=1e-7)), ("firststep", dict(rtol=1e-5, atol=1e-7, first_step=0.5))):
        for fname, f in (("npde", lambda y: fo.f(y[None])[0]), ("mlp", lambda y: fm.f(y[None])[0])):
            sol, st = dopri5.odeint_dopri5(f, g["x0"], t, **kw)
            acc = g[f"{case}_{fname}_accept"]
            assert (st["accepted"], st["rejected"]) == (int(acc.sum()), int(len(acc) - acc.sum())), (case, fname)
            assert st["nfe"] == int(g[f"{case}_{fname}_nfe"])
            assert relerr(sol, g[f"{case}_{fname}_sol"]) < 1e-9, (case, fname)


def test_oracle_dopri5_tuple_state_matches_reference():
    """Tuple states: one controller, error pooled per state tensor, max over the tuple (dopri5.py:108-109, misc.py:125-141, 161);
    fixture: the reference's odeint(tuple_f, (y0[:3], y0[3:]), t) with its accept / reject sequence (tuple_time.npz)."""
    from oracle import dopri5, npde
    g = load_golden("tuple_time")
    fo = npde.NPDEField(g["U"][None], g["Z"], 1.0, 0.75)
    f = lambda y: fo.f(y[None])[0]
    sol, st = dopri5.odeint_dopri5(f, g["x0"], g["t"], rtol=1e-5, atol=1e-7, groups=[3, 2])
    acc = g["tuple_dopri5_accept"]
    assert (st["accepted"], st["rejected"]) == (int(acc.sum()), int(len(acc) - acc.sum()))
    assert relerr(sol[:, :3], g["tuple_dopri5_a"]) < 1e-9 and relerr(sol[:, 3:], g["tuple_dopri5_b"]) < 1e-9
    _, st1 = dopri5.odeint_dopri5(f, g["x0"], g["t"], rtol=1e-5, atol=1e-7)
    acc1 = g["single_dopri5_accept"]
    assert (st1["accepted"], st1["rejected"]) == (int(acc1.sum()), int(len(acc1) - acc1.sum())) and len(acc1) != len(acc)
