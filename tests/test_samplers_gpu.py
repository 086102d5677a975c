"""GPU: fused sampler updates and the SVGD interaction (through the C ABI) vs the reference goldens / the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu
TOL = 2e-6          # fp32 elementwise update vs the float64 reference


def _single_chain_field():
    import bayesian_ode_b200 as bode
    g = load_golden("npde_m5")
    return bode.NPDEField(torch.from_numpy(g["U0"]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)


def _set(f, U, logsn):
    f.U.data.copy_(torch.from_numpy(U).view_as(f.U))
    f.logsn.data.copy_(torch.from_numpy(logsn).view_as(f.logsn))


def _setgrad(f, gU, gl):
    f.bind_flat_grads()
    f.U.grad.copy_(torch.from_numpy(gU).view_as(f.U))
    f.logsn.grad.copy_(torch.from_numpy(gl).view_as(f.logsn))


def _noise(g, key, i):
    return [torch.from_numpy(g[f"{key}U"][i])[None], torch.from_numpy(g[f"{key}logsn"][i])[None]]


def test_sgld_matches_reference_steps():
    from bayesian_ode_b200.samplers import SGLD
    g = load_golden("sampler_steps")
    f = _single_chain_field()
    smp = SGLD([f.U, f.logsn], lr0=1e-4, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03)
    assert smp._flat is not None and smp._flat.shape == (1, 52)
    _set(f, g["sgld_U"][0], g["sgld_logsn"][0])
    for i in range(4):
        _setgrad(f, g["sgld_gU"][i], g["sgld_glogsn"][i])
        lr = smp.get_lr(i)
        assert lr == float(g["sgld_lr"][i])                      # schedule is host float64: bit-exact
        smp.step(lr=lr, noise=_noise(g, "sgld_xi", i))
        assert relerr(f.U.data.cpu().numpy()[0], g["sgld_U_new"][i]) < TOL
        assert relerr(f.logsn.data.cpu().numpy()[0], g["sgld_logsn_new"][i]) < TOL


def test_psgld_matches_reference_steps():
    from bayesian_ode_b200.samplers import pSGLD
    g = load_golden("sampler_steps")
    f = _single_chain_field()
    smp = pSGLD([f.U, f.logsn], lr0=5e-3, lr_gamma=0.51, lr_t0=100, lr_alpha=0.1, lambda_=1e-8, alpha=0.99, N=5)
    _set(f, g["psgld_U"][0], g["psgld_logsn"][0])
    for i in range(4):
        _setgrad(f, g["psgld_gU"][i], g["psgld_glogsn"][i])
        smp.step(lr=smp.get_lr(i), noise=_noise(g, "psgld_xi", i))
        assert relerr(f.U.data.cpu().numpy()[0], g["psgld_U_new"][i]) < 1e-5
        assert relerr(f.logsn.data.cpu().numpy()[0], g["psgld_logsn_new"][i]) < 1e-5
    V = list(smp._V.values())[0].cpu().numpy()[0]
    assert relerr(V[:50].reshape(25, 2), g["psgld_VU_final"]) < 1e-5


def test_asghmc_matches_reference_steps():
    from bayesian_ode_b200.samplers import aSGHMC
    g = load_golden("sampler_steps")
    burn, k = int(g["asghmc_burn"]), int(g["asghmc_resample_every"])
    f = _single_chain_field()
    smp = aSGHMC([f.U, f.logsn], lr=1e-2, mom_decay=5e-2, lambda_=1e-5)
    _set(f, g["asghmc_U"][0], g["asghmc_logsn"][0])
    for i in range(7):
        _setgrad(f, g["asghmc_gU"][i], g["asghmc_glogsn"][i])
        smp.step(lr=1e-2, burn_in=i < burn, resample_mom_every=k, noise=_noise(g, "asghmc_xi", i),
                 noise_resample=_noise(g, "asghmc_xr", i))
        assert relerr(f.U.data.cpu().numpy()[0], g["asghmc_U_new"][i]) < 1e-5, i
        assert relerr(f.logsn.data.cpu().numpy()[0], g["asghmc_logsn_new"][i]) < 1e-5, i
    st = list(smp._st.values())[0]
    for key in ("tau", "g", "v_hat", "momentum"):
        assert relerr(st[key].cpu().numpy()[0][:50].reshape(25, 2), g[f"asghmc_{key}_U_final"]) < 1e-5, key


def test_csgld_matches_reference_steps_incl_noise_gating():
    """langevin.py:1600-1724: cosine schedule (host float64, bit-exact) and noise only where r > beta."""
    from bayesian_ode_b200.samplers import cSGLD
    g = load_golden("cyclical_steps")
    n_it, M, beta = int(g["num_iters"]), int(g["M"]), float(g["beta"])
    f = _single_chain_field()
    smp = cSGLD([f.U, f.logsn], lr0=2e-4, M=M, beta=beta)
    smp.num_iters = n_it
    _set(f, g["csgld_U"][0], g["csgld_logsn"][0])
    for i in range(n_it):
        _setgrad(f, g["csgld_gU"][i], g["csgld_glogsn"][i])
        assert smp._r(i) == float(g["csgld_r"][i]) and smp.get_lr(i) == float(g["csgld_lr"][i])
        # the injected draws are ignored outside the sampling part of a cycle, exactly like the reference draws none
        smp.step(iter_num=i, lr=smp.get_lr(i), noise=[torch.ones(1, 25, 2), torch.ones(1, 2)] if not smp._sampling_phase(i)
                 else _noise(g, "csgld_xi", i))
        assert relerr(f.U.data.cpu().numpy()[0], g["csgld_U_new"][i]) < TOL, i
        assert relerr(f.logsn.data.cpu().numpy()[0], g["csgld_logsn_new"][i]) < TOL, i


def test_acsghmc_matches_reference_steps():
    """hamiltonian.py:167-326: aSGHMC update on the cosine schedule, momentum noise gated by r > beta, resampling."""
    from bayesian_ode_b200.samplers import acSGHMC
    g = load_golden("cyclical_steps")
    n_it, M, beta = int(g["num_iters"]), int(g["M"]), float(g["beta"])
    burn, k = int(g["acsghmc_burn"]), int(g["acsghmc_resample_every"])
    f = _single_chain_field()
    smp = acSGHMC([f.U, f.logsn], lr0=1e-2, M=M, beta=beta, mom_decay=5e-2, lambda_=1e-5)
    smp.num_iters = n_it
    _set(f, g["acsghmc_U"][0], g["acsghmc_logsn"][0])
    for i in range(n_it):
        _setgrad(f, g["acsghmc_gU"][i], g["acsghmc_glogsn"][i])
        lr = smp.get_lr(i)
        assert lr == float(g["acsghmc_lr"][i])
        smp.step(lr=lr, iter_num=i, burn_in=i < burn, resample_mom_every=k, noise=_noise(g, "acsghmc_xi", i),
                 noise_resample=_noise(g, "acsghmc_xr", i))
        assert relerr(f.U.data.cpu().numpy()[0], g["acsghmc_U_new"][i]) < 1e-5, i
        assert relerr(f.logsn.data.cpu().numpy()[0], g["acsghmc_logsn_new"][i]) < 1e-5, i
    st = list(smp._st.values())[0]
    for key in ("tau", "g", "v_hat", "momentum"):
        assert relerr(st[key].cpu().numpy()[0][:50].reshape(25, 2), g[f"acsghmc_{key}_U_final"]) < 1e-5, key


def test_cyclical_sample_loop_records_none_outside_sampling_phase():
    """langevin.py:1697-1706: every sampling iteration appends an entry; params are None where r <= beta."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    from bayesian_ode_b200.samplers import cSGLD
    data = problems.make_dataset(seed=0)
    Z = problems.inducing_grid(data["Y"], 5)
    U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
    f = bode.NPDEField(U0[None].repeat(8, 1, 1), Z, 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
    smp = cSGLD([f.U, f.logsn], lr0=1e-6, M=3, beta=0.25)
    chain = smp.sample(post, num_samples=9, burn_in=3)
    assert len(chain) == 9
    for i in range(9):
        params, acc = chain[i]
        if smp._sampling_phase(i + 3):
            assert params[0][0].shape == (8, 25, 2) and params[0][1].shape == (8, 2) and np.isfinite(params[0][0]).all()
        else:
            assert params == [[None, None]]
        assert acc is True


def test_non_flat_tensors_and_scalar_tail():
    """Parameters that are NOT column blocks of one buffer take one launch per tensor; logsn (2 elements) exercises
    the scalar tail of the vectorised kernel."""
    from bayesian_ode_b200.samplers import SGLD
    from oracle import samplers as osamp
    rng = np.random.default_rng(3)
    a = torch.nn.Parameter(torch.from_numpy(rng.standard_normal((7, 3))).float().cuda())
    b = torch.nn.Parameter(torch.from_numpy(rng.standard_normal(2)).float().cuda())
    smp = SGLD([a, b], lr0=1e-2, lr_gamma=0.5, lr_t0=1, lr_alpha=1)
    assert smp._flat is None
    ga, gb = rng.standard_normal((7, 3)), rng.standard_normal(2)
    xa, xb = rng.standard_normal((7, 3)), rng.standard_normal(2)
    a.grad, b.grad = torch.from_numpy(ga).float().cuda(), torch.from_numpy(gb).float().cuda()
    a0, b0 = a.data.cpu().numpy().astype(np.float64), b.data.cpu().numpy().astype(np.float64)
    smp.step(lr=0.01, noise=[torch.from_numpy(xa), torch.from_numpy(xb)])
    assert relerr(a.data.cpu().numpy(), osamp.sgld_step(a0, ga, 0.01, xa)) < TOL
    assert relerr(b.data.cpu().numpy(), osamp.sgld_step(b0, gb, 0.01, xb)) < TOL


def test_nan_parameter_raises_like_reference():
    from bayesian_ode_b200.samplers import SGLD
    f = _single_chain_field()
    smp = SGLD([f.U, f.logsn], lr0=1e-4, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03)
    f.bind_flat_grads()
    f.U.data[0, 3, 1] = float("nan")
    with pytest.raises(ValueError):                    # langevin.py:184-185
        smp.step(lr=1e-4)


def test_philox_normals_are_normal_and_counter_based():
    import bayesian_ode_b200 as bode
    lib = bode._lib.load()
    n = 1 << 20
    out = torch.empty(n, device="cuda")
    st = bode._lib.stream_ptr()
    bode._lib.check(lib.bode_fill_normal(bode._lib.ptr(out), n, 1234, 0, st))
    x = out.double().cpu().numpy()
    assert abs(x.mean()) < 5e-3 and abs(x.var() - 1) < 5e-3
    assert abs((x ** 4).mean() - 3) < 0.05 and abs((x ** 3).mean()) < 0.02
    out2 = torch.empty(n, device="cuda")
    bode._lib.check(lib.bode_fill_normal(bode._lib.ptr(out2), n, 1234, 0, st))
    assert torch.equal(out, out2)                       # same (seed, step) -> same stream
    bode._lib.check(lib.bode_fill_normal(bode._lib.ptr(out2), n, 1234, 1, st))
    assert abs(float((out * out2).mean())) < 5e-3       # different step -> independent stream
    # in-kernel noise in SGLD: after one step with g = 0 the displacement is sqrt(2 lr) * N(0,1)
    p = torch.zeros(n, device="cuda"); g = torch.zeros(n, device="cuda")
    bode._lib.check(lib.bode_sgld_step(bode._lib.ptr(p), bode._lib.ptr(g), None, n, 1e-2, 1, 7, 0, None, None, st))
    assert abs(float(p.var()) / (2 * 1e-2) - 1) < 1e-2


@pytest.mark.parametrize("n", [64, 257])
def test_rbf_kernel_matches_reference(n):
    from bayesian_ode_b200.samplers import RBFKernel
    g = load_golden("svgd")
    X = torch.from_numpy(g[f"n{n}_X"]).float().cuda()
    ker = RBFKernel()
    K = ker(X, X)
    assert relerr(K.cpu().numpy(), g[f"n{n}_K"]) < 1e-5
    assert relerr(float(ker.last_median), g[f"n{n}_median"]) < 1e-5
    Kf = RBFKernel(sigma=0.7)(X, X)
    assert relerr(Kf.cpu().numpy(), g[f"n{n}_K_sigma07"]) < 1e-5


@pytest.mark.parametrize("n", [64, 257])
def test_svgd_phi_matches_reference(n):
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200.samplers import SVGD
    g = load_golden("svgd")
    X, S = g[f"n{n}_X"], g[f"n{n}_S"]
    f = bode.NPDEField(torch.from_numpy(X[:, :50].reshape(n, 25, 2)), torch.from_numpy(load_golden("npde_m5")["Z"]), 1.0, 0.75, 0.1)
    f.logsn.data.copy_(torch.from_numpy(X[:, 50:]))
    smp = SVGD([f.U, f.logsn], lr=1e-4)
    f.bind_flat_grads().copy_(torch.from_numpy(-S))               # grad of the loss = -score
    phi = smp.phi().clone()
    assert relerr(phi.cpu().numpy(), g[f"n{n}_phi"]) < 1e-4
    th0 = f.theta.clone()
    smp.step()                                                    # theta += lr * phi  (descends -phi)
    # the fused update is theta + lr*phi in fp32 (one FMA): compare against the same expression, not the difference
    want = th0.double() + 1e-4 * phi.double()
    assert float((f.theta.double() - want).abs().max()) <= 2.4e-7 * float(th0.abs().max())


def test_median_selection_is_bit_exact_at_full_size():
    """P = 4096 (BASELINE config 3): the selected order statistics equal np.median of the very d2 matrix the kernel
    produced, bit for bit (even count -> mean of the two middle values); phi rows agree with the float64 oracle."""
    from bayesian_ode_b200.samplers.stein import _Workspace
    from oracle import samplers as osamp
    g = load_golden("npde_m5")
    rng = np.random.default_rng(5)
    n, d = 4096, 52
    X = np.concatenate([g["U0"].reshape(1, -1) + 0.1 * rng.standard_normal((n, 50)),
                        np.log(0.1) + 0.05 * rng.standard_normal((n, 2))], 1)
    S = 3.0 * rng.standard_normal((n, d))
    Xd = torch.from_numpy(X).float().cuda()
    ws = _Workspace(n, n, d, Xd.device)
    ws.sqdist(Xd, n, Xd, n, d, n * n, row_offset=0)
    ws.median(n, n, d, n)
    d2 = ws.d2(n, n).cpu().numpy()
    med = float(ws.med_gamma[0])
    assert np.float32(med) == np.median(d2)                       # bit-exact selection
    assert abs(med - np.median(osamp.sq_dists(X[:512], X[:512]))) / med < 0.05    # sanity vs float64 on a subsample
    assert np.allclose(d2, d2.T, rtol=0, atol=2e-6) and np.all(np.diag(d2) == 0)      # tensor-core tiles: symmetric to rounding
    gam64 = 1.0 / (1e-8 + 2 * (np.median(d2.astype(np.float64)) / (2 * np.log(n + 1))))
    assert abs(float(ws.med_gamma[1]) - gam64) / gam64 < 1e-6
    # odd count (n*n odd): single middle element
    n2 = 255
    X2 = Xd[:n2].contiguous()
    ws2 = _Workspace(n2, n2, d, Xd.device)
    ws2.sqdist(X2, n2, X2, n2, d, n2 * n2, row_offset=0)
    ws2.median(n2, n2, d, n2)
    assert np.float32(float(ws2.med_gamma[0])) == np.median(ws2.d2(n2, n2).cpu().numpy())
    # phi on a row subset vs the oracle using the kernel's own gamma
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200.samplers import SVGD
    f = bode.NPDEField(torch.from_numpy(X[:, :50].reshape(n, 25, 2)), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    f.logsn.data.copy_(torch.from_numpy(X[:, 50:]))
    smp = SVGD([f.U, f.logsn])
    f.bind_flat_grads().copy_(torch.from_numpy(-S))
    phi = smp.phi().cpu().numpy()
    rows = np.concatenate([np.arange(4), rng.choice(n, 28, replace=False)])
    Xf = f.theta.cpu().numpy().astype(np.float64)
    ref = osamp.svgd_phi(Xf, S.astype(np.float32).astype(np.float64), rows=rows, gamma=float(smp._ws.med_gamma[1]))
    assert relerr(phi[rows], ref) < 1e-4


def test_median_window_path_is_bit_exact_and_falls_back():
    """The warm-workspace path (window table filled by the Gram epilogue, svgd_state.cuh) must return the same order
    statistics, bit for bit, as np.median of the kernel's own d2 -- when the median drifts a little (window hit), when it
    jumps (window miss -> radix fallback), for an odd and an even entry count, and for ragged tile edges (n % 128 != 0)."""
    from bayesian_ode_b200 import _lib
    from bayesian_ode_b200.samplers.stein import _Workspace
    lib = _lib.load()
    rng = np.random.default_rng(11)
    for n, side in ((300, 0), (1024, 0), (1024, 40)):             # n*n even; 300 % 128 != 0 exercises the ragged tiles
        # side > 0: the launches of the overlapped step -- 4-CTA selection cluster, cooperative fallback grid of `side` CTAs
        d = 52
        X0 = torch.from_numpy((rng.standard_normal((n, d)) * 0.3 + 1.5).astype(np.float32)).cuda()
        ws = _Workspace(n, n, d, X0.device)
        scales = [1.0, 1.0 + 2e-5, 1.0 - 3e-5, 1.7, 1.7 + 1e-5, 0.2]    # small drifts hit the window, 1.7x / 0.2x jumps miss it
        old = lib.bode_svgd_set_select_ctas(side)
        try:
            for it, sc in enumerate(scales):
                X = (X0 * sc).contiguous()
                ws.sqdist(X, n, X, n, d, n * n, row_offset=0)
                ws.median(n, n, d, n)
                d2 = ws.d2(n, n).cpu().numpy()
                assert np.float32(float(ws.med_gamma[0])) == np.median(d2), (n, side, it)
        finally:
            lib.bode_svgd_set_select_ctas(old)
    # rectangular block (rows are a prefix of the columns, as on rank 0 of a sharded job)
    n = 255
    X0 = torch.from_numpy((rng.standard_normal((n, 52)) * 0.3 + 1.5).astype(np.float32)).cuda()
    Xp = torch.zeros(256, 52, device="cuda"); Xp[:n] = X0        # rows padded so the column count stays a multiple of 4
    ws = _Workspace(n, 256, 52, X0.device)
    for sc in (1.0, 1.0 + 1e-5):
        X = (Xp * sc).contiguous()
        ws.sqdist(X[:n], n, X, 256, 52, n * 256, row_offset=0)
        ws.median(n, 256, 52, 256)
        assert np.float32(float(ws.med_gamma[0])) == np.median(ws.d2(n, 256).cpu().numpy())


def test_tiled_d2_layout_rectangular_blocks():
    """Blocks whose edges are whole 128 x 128 Gram tiles are stored as [rows/128][cols/32][128][32] tiles (bode_svgd_d2_tiled),
    the others row-major; the accessor must return the same matrix as the float64 oracle either way -- for rows at an offset among
    the columns (rank > 0 of a sharded job: the diagonal of cdist(x, x) is forced to zero at row_offset), for column counts that
    are not a multiple of the Gram tile, and for unrelated row / column sets (row_offset = -1: NO entry is forced to zero)."""
    from bayesian_ode_b200 import _lib
    from bayesian_ode_b200.samplers.stein import _Workspace
    from oracle import samplers as osamp
    rng = np.random.default_rng(23)
    d = 52
    cases = ((128, 160, 0, 0), (128, 160, 32, 0), (256, 512, 256, 1), (128, 96, -1, 0), (128, 256, 128, 1), (256, 384, 0, 1),
             (128, 256, -1, 1))
    for nr, nc, off, tiled in cases:
        Xc = (rng.standard_normal((nc, d)) * 0.3 + 1.5).astype(np.float32)
        Xr = Xc[off:off + nr] if off >= 0 else (rng.standard_normal((nr, d)) * 0.3 + 1.5).astype(np.float32)
        Xcd = torch.from_numpy(Xc).cuda()
        Xrd = Xcd[off:off + nr] if off >= 0 else torch.from_numpy(Xr).cuda()
        ws = _Workspace(nr, nc, d, Xcd.device)
        assert _lib.load().bode_svgd_d2_tiled(nr, nc, d) == tiled, (nr, nc)
        ref = osamp.sq_dists(Xr.astype(np.float64), Xc.astype(np.float64))
        for _ in range(2):                                        # second call: armed window
            ws.sqdist(Xrd, nr, Xcd, nc, d, nr * nc, row_offset=off)
            ws.median(nr, nc, d, nc)
            d2 = ws.d2(nr, nc).cpu().numpy()
            assert np.abs(d2 - ref).max() < 2e-5 * ref.max(), (nr, nc, off)
            if off >= 0:
                assert np.all(d2[np.arange(nr), off + np.arange(nr)] == 0)
            assert np.float32(float(ws.med_gamma[0])) == np.median(d2), (nr, nc, off)


def test_svgd_stream_overlap_is_bit_identical_and_graph_capturable():
    """prefetch()/phi() fork the position-only operands and the V operand onto a side stream (same kernels, same inputs):
    the particles must come out bit-identical to the serial order, eagerly and when the step is replayed from one CUDA graph."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    from bayesian_ode_b200.samplers import SVGD
    data = problems.make_dataset(seed=0)
    Z = problems.inducing_grid(data["Y"], 5)
    U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
    U = U0[None] + 0.1 * torch.randn(384, 25, 2, generator=torch.Generator().manual_seed(5), dtype=torch.float64)

    def run(overlap, graph, fuse=True):
        f = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
        post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
        f.bind_flat_grads()
        smp = SVGD([f.U, f.logsn], lr=1e-4, overlap=overlap)
        smp.fuse_scores = fuse
        assert smp.overlap == overlap
        lib = bode._lib.load()

        def step():
            smp.prefetch()
            post.loss_and_grad_()
            smp.phi(update_lr=1e-4)
        step()
        torch.cuda.synchronize()
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            step = g.replay
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        assert lib.bode_npde_set_cta_limit(0) == 0              # phi() restored the solver's SM budget
        assert lib.bode_svgd_disarm_score_tiles() == 0          # ... and ended the closure kernel's writes into the phi operand
        # "gram": the closure kernel itself wrote the score half of the phi operand (bode_svgd_arm_score_tiles), no launch for it
        assert smp.last_scores_fused == (overlap == "gram" and fuse)
        return f.theta.clone(), smp._ws.med_gamma.clone()

    th0, mg0 = run(False, False)
    for overlap, graph, fuse in (("operands", False, True), ("operands", True, True), ("gram", False, True), ("gram", True, True),
                                 ("gram", False, False)):
        th, mg = run(overlap, graph, fuse)
        assert torch.equal(mg, mg0), (overlap, graph)
        # "gram" packs the solver into fewer, larger CTAs: the per-particle reduction over trajectories is unchanged, so the
        # particles still agree bit for bit
        assert torch.equal(th, th0), (overlap, graph)


def test_sample_loop_fused_and_protocol_paths_agree():
    """SGLD.sample() drives NPDEPosterior through the fused path; the autograd closure protocol gives the same chain
    when the same noise stream is used (in-kernel Philox is a pure function of (seed, step, element))."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200.samplers import SGLD
    g = load_golden("npde_m5")
    chains = []
    for fused in (True, False):
        f = bode.NPDEField(torch.from_numpy(g["U"]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
        post = bode.NPDEPosterior(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["Y"]))
        smp = SGLD([f.U, f.logsn], lr0=1e-4, lr_gamma=0.51, lr_t0=100, lr_alpha=0.03, seed=11)
        closure = post if fused else (lambda add_prior=True, post=post: post(add_prior))
        chain = smp.sample(closure, num_samples=3, burn_in=2)
        assert len(chain) == 3
        params, accepted = chain[-1]
        assert accepted is True and params[0][0].shape == (4, 25, 2) and params[0][1].shape == (4, 2)
        chains.append(np.concatenate([params[0][0].reshape(4, -1), params[0][1]], 1))
    assert relerr(chains[0], chains[1]) < 1e-6


def test_hamcmc_matches_reference_run():
    """HAMCMC on the reference's own 113-step run (two free-standing parameter tensors, injected noise, gradients of the
    same quadratic computed by autograd on the device): warm-up, history fill, start-up pairs, 8 metric steps."""
    from bayesian_ode_b200.samplers import HAMCMC
    g = load_golden("hamcmc")
    memory = int(g["memory"])
    M = memory + 1
    A = torch.from_numpy(g["A"]).float().cuda()
    th0 = torch.from_numpy(g["theta"][0]).float().cuda()
    a = torch.nn.Parameter(th0[:6].reshape(3, 2).clone())
    b = torch.nn.Parameter(th0[6:].clone())
    smp = HAMCMC([a, b], memory=memory, lr0=5e-2, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3, H_gamma=1.0, trust_reg=1.0)
    for i in range(g["grad"].shape[0]):
        smp.zero_grad()
        th = torch.cat([a.reshape(-1), b.reshape(-1)])
        (0.5 * th @ (A @ th)).backward()
        lr = smp.get_lr(i)
        assert lr == float(g["lr"][i])
        if i < 2 * M - 1 + 100:
            smp.step_without_metric(lr=lr, add_params=(i >= 100), noise=g["xi"][i])
        else:
            smp.step(lr=lr, noise=g["xi"][i])
        got = torch.cat([a.detach().reshape(-1), b.detach().reshape(-1)]).cpu().numpy()
        assert relerr(got, g["theta"][i + 1]) < (2e-5 if i < 105 else 5e-4), i
    assert int(smp.n_pairs()[0]) == int(g["n_pairs"])


def test_hamcmc_contiguous_variants_match_reference_runs():
    """HAMCMC2 / HAMCMC3 / HAMCMC4 on the reference's own runs (tests/golden/hamcmc_contiguous.npz: M warm-up + 10 metric steps
    each, injected noise, gradients of the same quadratic by autograd on the device).
    Tolerance per step: the fixture's runs are ill-conditioned in places (HAMCMC2's last step multiplies |theta| by 370, HAMCMC4
    grows to 1.6e6 -- the reference diverges on this quadratic), so the bar is tied to the step's MEASURED conditioning: the float64
    oracle re-run on the float32-rounded inputs (theta_0, A, xi) deviates from the reference by dev_i through input rounding alone
    (4e-8 on a well-conditioned step, 9e-4 on HAMCMC2's last one); the fp32 kernel must stay within max(2e-5, 40 dev_i)."""
    from bayesian_ode_b200 import samplers
    from oracle import samplers as osamp
    g = load_golden("hamcmc_contiguous")
    memory = int(g["memory"])
    M = memory + 1
    A = torch.from_numpy(g["A"]).float().cuda()
    A32 = g["A"].astype(np.float32).astype(np.float64)
    for variant in (2, 3, 4):
        th_ref, xi, lrs = g["theta%d" % variant], g["xi%d" % variant], g["lr%d" % variant]
        orc = osamp.HAMCMCContiguous(variant, memory=memory, H_gamma=1.0, trust_reg=1.0)
        th_o = th_ref[0].astype(np.float32).astype(np.float64)
        th0 = torch.from_numpy(th_ref[0]).float().cuda()
        a = torch.nn.Parameter(th0[:6].reshape(3, 2).clone())
        b = torch.nn.Parameter(th0[6:].clone())
        smp = getattr(samplers, "HAMCMC%d" % variant)([a, b], memory=memory, lr0=2e-2, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3,
                                                       H_gamma=1.0, trust_reg=1.0)
        for i in range(xi.shape[0]):
            smp.zero_grad()
            th = torch.cat([a.reshape(-1), b.reshape(-1)])
            (0.5 * th @ (A @ th)).backward()
            lr = smp.get_lr(i)
            assert lr == float(lrs[i])
            x64 = xi[i].astype(np.float32).astype(np.float64)
            if i < M:
                smp.step_without_metric(lr=lr, noise=xi[i])
                th_o = orc.step_without_metric(th_o, A32 @ th_o, float(np.float32(lr)), x64)
            else:
                smp.step(lr=lr, noise=xi[i])
                th_o = orc.step(A32 @ th_o, float(np.float32(lr)), x64)
            dev = relerr(th_o, th_ref[i + 1])
            got = torch.cat([a.detach().reshape(-1), b.detach().reshape(-1)]).cpu().numpy()
            assert relerr(got, th_ref[i + 1]) < max(2e-5, 40.0 * dev), (variant, i, dev)
        assert int(smp.n_pairs()[0]) == (M - 1 if variant == 4 else M - 2)


def test_hamcmc_contiguous_metric_step_on_partial_window_raises():
    """langevin.py:1205-1238 indexes the pair lists of a window that does not exist yet; the kernel leaves theta alone and reports."""
    from bayesian_ode_b200.samplers import HAMCMC4
    a = torch.nn.Parameter(torch.randn(6, device="cuda"))
    a.grad = torch.randn(6, device="cuda")
    smp = HAMCMC4([a], memory=3, lr0=1e-2, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3)
    smp.step_without_metric(lr=1e-2)
    before = a.detach().clone()
    with pytest.raises(RuntimeError, match="history window"):
        smp.step(lr=1e-2)
    assert torch.equal(a.detach(), before)


@pytest.mark.parametrize("d", [9, 52, 100, 300, 514, 900])
def test_hamcmc_register_sliced_kernel_equals_generic_kernel(d, monkeypatch):
    """hamcmc.cu has two kernels: the register-sliced one (d <= 1024, product-form vectors in shared memory) and the generic one
    over the global work buffer (BODE_HAMCMC_GENERIC=1 forces it).  Same element ownership where both run 128 threads, float64 dot
    products in both: 24 warm-up + metric steps from the same state and the same injected noise agree to 1e-6 of |theta| (bit-identical
    for d > 64) and build the same number of curvature pairs.  d spans every instantiation."""
    from bayesian_ode_b200.samplers import HAMCMC
    P, memory = 33, 3
    gen = torch.Generator().manual_seed(d)
    A = torch.randn(P, d, d, generator=gen) / d ** 0.5
    A = (A @ A.transpose(1, 2) + 0.5 * torch.eye(d)).cuda()
    th0 = torch.randn(P, d, generator=gen).cuda()
    xi = torch.randn(24, P, d, generator=gen)
    out = {}
    for generic in (0, 1):
        if generic:
            monkeypatch.setenv("BODE_HAMCMC_GENERIC", "1")
        else:
            monkeypatch.delenv("BODE_HAMCMC_GENERIC", raising=False)
        th = torch.nn.Parameter(th0.clone())
        smp = HAMCMC([th], memory=memory, lr0=2e-3, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3, H_gamma=1.0, trust_reg=1.0)
        smp.check_finite = "deferred"                      # the reference's metric step can blow a chain up; both kernels must agree on it
        traj = []
        for i in range(24):
            th.grad = torch.einsum("pij,pj->pi", A, th.detach())
            lr = smp.get_lr(i)
            if i < 2 * (memory + 1) - 1:
                smp.step_without_metric(lr=lr, add_params=True, noise=xi[i])
            else:
                smp.step(lr=lr, noise=xi[i])
            traj.append(th.detach().clone())
        out[generic] = (torch.stack(traj), smp.n_pairs().clone())
    a, b = out[0][0], out[1][0]
    fin = torch.isfinite(a).all(dim=(0, 2))
    assert torch.equal(fin, torch.isfinite(b).all(dim=(0, 2))) and int(fin.sum()) >= P // 2
    assert torch.equal(out[0][1], out[1][1]) and int(out[0][1].max()) >= 1
    if d > 64:
        assert torch.equal(a[:, fin], b[:, fin])
    else:
        assert float((a[:, fin] - b[:, fin]).abs().max()) <= 1e-5 * float(b[:, fin].abs().max())


@pytest.mark.parametrize("variant", [2, 3, 4])
@pytest.mark.parametrize("d", [20, 130, 514])
def test_hamcmc_contiguous_register_sliced_kernel_equals_generic_kernel(variant, d, monkeypatch):
    """hamcmc_contig.cu: register-sliced kernel vs the generic one (BODE_HAMCMC_GENERIC=1) for HAMCMC2 / 3 / 4 -- window fill, the
    contiguous pairs, ten metric steps, rings wrapping around; same state, same injected noise."""
    from bayesian_ode_b200 import samplers
    P, memory = 19, 3
    M = memory + 1
    gen = torch.Generator().manual_seed(10 * d + variant)
    A = torch.randn(P, d, d, generator=gen) / d ** 0.5
    A = (A @ A.transpose(1, 2) + 0.5 * torch.eye(d)).cuda()
    th0 = torch.randn(P, d, generator=gen).cuda()
    xi = torch.randn(M + 10, P, d, generator=gen)
    out = {}
    for generic in (0, 1):
        if generic:
            monkeypatch.setenv("BODE_HAMCMC_GENERIC", "1")
        else:
            monkeypatch.delenv("BODE_HAMCMC_GENERIC", raising=False)
        th = torch.nn.Parameter(th0.clone())
        smp = getattr(samplers, "HAMCMC%d" % variant)([th], memory=memory, lr0=2e-3, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3,
                                                       H_gamma=1.0, trust_reg=1.0)
        smp.check_finite = "deferred"
        traj = []
        for i in range(M + 10):
            th.grad = torch.einsum("pij,pj->pi", A, th.detach())
            lr = smp.get_lr(i)
            if i < M:
                smp.step_without_metric(lr=lr, noise=xi[i])
            else:
                smp.step(lr=lr, noise=xi[i])
            traj.append(th.detach().clone())
        out[generic] = (torch.stack(traj), smp.n_pairs().clone())
    a, b = out[0][0], out[1][0]
    fin = torch.isfinite(a).all(dim=(0, 2))
    assert torch.equal(fin, torch.isfinite(b).all(dim=(0, 2))) and int(fin.sum()) >= P // 2
    assert torch.equal(out[0][1], out[1][1])
    if d > 64:
        assert torch.equal(a[:, fin], b[:, fin])
    else:
        assert float((a[:, fin] - b[:, fin]).abs().max()) <= 1e-5 * float(b[:, fin].abs().max())


def test_hamcmc_batched_chains_on_npde():
    """P chains on the flat theta buffer: sample() drives warm-up -> metric steps through the fused closure."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200.samplers import HAMCMC
    g = load_golden("npde_m5")
    f = bode.NPDEField(torch.from_numpy(g["U"]), torch.from_numpy(g["Z"]), 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, torch.from_numpy(g["x0"]), torch.from_numpy(g["t"]), torch.from_numpy(g["Y"]))
    smp = HAMCMC([f.U, f.logsn], memory=2, lr0=1e-6, lr_gamma=0.55, lr_t0=100, lr_alpha=0.3)
    chain, logp = smp.sample(post, num_samples=3, burn_in=108, print_iters=False)
    assert len(chain) == 3 and len(logp) == 111
    assert bool(torch.isfinite(f.theta).all())
    assert smp._hs["meta"][:, 0].tolist() == [5, 5, 5, 5]


def test_mala_matches_reference_run():
    """samplers/langevin.py:13-149 against the fixture recorded from the reference: the batched MALA class replays the
    reference chain (proposal noise and log u injected) and must reproduce every accept flag (selection logic: bit exact) and
    the proposals (fp32 of an fp64 reference); the kernel's log-ratio matches the float64 oracle in both modes."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200.samplers import MALA
    from oracle import samplers as osamp
    g = load_golden("mala_steps")
    lr = float(g["lr"])
    cA, cB = torch.from_numpy(g["cA"]).float().cuda(), torch.from_numpy(g["cB"]).float().cuda()
    P = 3                                                    # three identical chains: the batch axis must not mix them
    theta = torch.zeros(P, 15, device="cuda")
    A = torch.nn.Parameter(theta[:, :12].view(P, 6, 2))
    B = torch.nn.Parameter(theta[:, 12:].view(P, 3))

    def closure(add_prior=True):
        return 0.5 * (cA * A * A).sum((1, 2)) + 0.5 * (cB * B * B).sum(1) + 0.1 * (A.sum((1, 2)) * B.sum(1))

    smp = MALA([A, B], lr=lr)
    for i in range(len(g["accepted"])):
        with torch.no_grad():                                # every step starts from the reference's own state
            A.copy_(torch.from_numpy(g["A"][i]).float().cuda().expand(P, 6, 2))
            B.copy_(torch.from_numpy(g["B"][i]).float().cuda().expand(P, 3))
        smp.zero_grad()
        smp.loss = closure()
        smp._backward(smp.loss)
        assert relerr(A.grad[1].cpu().numpy(), g["gA"][i]) < 1e-5
        xi = torch.from_numpy(np.concatenate([g["xiA"][i].ravel(), g["xiB"][i].ravel()])).float().cuda().expand(P, 15).contiguous()
        smp.step(noise=xi)
        assert relerr(A.detach()[2].cpu().numpy(), g["A_new"][i]) < 2e-6 and relerr(B.detach()[0].cpu().numpy(), g["B_new"][i]) < 2e-6
        params, acc = smp.accept_or_reject(closure, log_u=torch.full((P,), float(g["logu"][i])))
        assert acc.cpu().tolist() == [int(g["accepted"][i])] * P, i
        assert relerr(A.detach()[0].cpu().numpy(), g["A_after"][i]) < 2e-6             # nothing restored (reference quirk)
        flat = lambda a, b: np.concatenate([g[a][i].ravel(), g[b][i].ravel()])[None]
        la = osamp.mala_log_alpha(flat("A", "B"), flat("A_new", "B_new"), flat("gA", "gB"), flat("gA_new", "gB_new"),
                                  np.array([g["loss"][i]]), np.array([g["loss_new"][i]]), lr, aliased=True)
        assert abs(float(smp.log_alpha[0]) - float(la[0])) < 2e-4 * max(1.0, abs(float(la[0])))

    # textbook mode: ratio vs the oracle (aliased=False), rejected chains restored, accepted ones kept
    lib = bode._lib.load()
    rng = np.random.default_rng(3)
    P, d = 257, 52
    th0, th1 = rng.standard_normal((P, d)), None
    g0, g1 = rng.standard_normal((P, d)), rng.standard_normal((P, d))
    th1 = th0 - 0.05 * g0 + np.sqrt(0.1) * rng.standard_normal((P, d))
    l0, l1 = rng.standard_normal(P) * 3, rng.standard_normal(P) * 3
    lu = np.log(rng.uniform(size=P))
    lu[:3] = [-np.inf, 0.0, -1e-30]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
    t0, t1, G0, G1, L0, L1, LU = dev(th0), dev(th1), dev(g0), dev(g1), dev(l0), dev(l1), dev(lu)
    la_out = torch.empty(P, device="cuda"); acc = torch.empty(P, dtype=torch.int32, device="cuda")
    t1_in = t1.clone()
    bode._lib.check(lib.bode_mala_accept(bode._lib.ptr(t0), d, bode._lib.ptr(t1), d, bode._lib.ptr(G0), d, bode._lib.ptr(G1), d,
                                         bode._lib.ptr(L0), bode._lib.ptr(L1), bode._lib.ptr(LU), P, d, 0.05, 1, 0, 0,
                                         bode._lib.ptr(la_out), bode._lib.ptr(acc), bode._lib.stream_ptr()))
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    la = osamp.mala_log_alpha(f32(th0), f32(th1), f32(g0), f32(g1), f32(l0), f32(l1), float(np.float32(0.05)), aliased=False)
    assert np.abs(la_out.cpu().numpy() - la).max() < 1e-4 * np.abs(la).max()
    want = osamp.mala_accept(la, f32(lu))
    clear = np.abs(la - f32(lu)) > 1e-3 * np.maximum(1.0, np.abs(la))               # away from fp32 ties the decisions are identical
    assert np.array_equal(acc.cpu().numpy().astype(bool)[clear], want[clear]) and clear.sum() > 200
    a = acc.cpu().numpy().astype(bool)
    assert torch.equal(t1[torch.from_numpy(a).cuda()], t1_in[torch.from_numpy(a).cuda()])
    assert torch.equal(t1[torch.from_numpy(~a).cuda()], t0[torch.from_numpy(~a).cuda()])
    assert 20 < a.sum() < P - 20


def test_mala_sample_loop_on_npde_posterior():
    """MALA.sample with the fused NPDEPosterior closure (langevin.py:98-149 protocol): chain entries are [params, accepted],
    exact=True restores rejected chains (theta of a rejected chain equals its value before the proposal)."""
    import bayesian_ode_b200 as bode
    from bayesian_ode_b200 import problems
    from bayesian_ode_b200.samplers import MALA
    data = problems.make_dataset(seed=0)
    Z = problems.inducing_grid(data["Y"], 5)
    U0 = problems.gradient_matching_init(data["Y"], data["t"], Z, 1.0, 0.75)
    P = 16
    U = U0[None] + 0.05 * torch.randn(P, 25, 2, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    f = bode.NPDEField(U, Z, 1.0, 0.75, 0.1)
    post = bode.NPDEPosterior(f, data["x0"], data["t"], torch.from_numpy(data["Y"]))
    f.bind_flat_grads()
    smp = MALA([f.U, f.logsn], lr=2e-5, exact=True, seed=3)
    chain = smp.sample(post, num_samples=6, burn_in=2)
    assert len(chain) == 6
    params, acc = chain[-1]
    assert params[0][0].shape == (P, 25, 2) and acc.shape == (P,) and acc.dtype == torch.int32
    assert bool(torch.isfinite(f.theta).all())
    # one more step by hand: rejected chains are restored exactly, accepted ones keep the proposal
    smp.loss = post.loss_and_grad_()[0].clone()
    before = f.theta.clone()
    smp.step()
    proposal = f.theta.clone()
    _, acc = smp.accept_or_reject(post, log_u=torch.full((P,), -float("inf")))     # every finite ratio accepts: theta stays the proposal
    assert bool(acc.bool().all()) and torch.equal(f.theta, proposal)
    la = smp.log_alpha.clone()
    sign = torch.tensor([1.0, -1.0], device="cuda").repeat(P // 2)                 # even chains: log u above the ratio -> rejected
    _, acc = smp.accept_or_reject(post, log_u=la + sign * (1.0 + 0.01 * la.abs()))
    a = acc.bool()
    assert a.tolist() == [False, True] * (P // 2)
    assert torch.equal(f.theta[a], proposal[a]) and torch.equal(f.theta[~a], before[~a])
    assert bool(a.any()) and bool((~a).any())
