"""CPU: oracle sampler updates / SVGD kernel+phi vs fixtures produced by the reference's samplers/ package."""
import numpy as np

from conftest import load_golden, relerr
from oracle import samplers as osamp


def test_lr_schedule_matches_reference():
    g = load_golden("sampler_steps")
    for i in range(4):
        assert osamp.get_lr(i, 1e-4, 0.51, 100, 0.03) == float(g["sgld_lr"][i])
        assert osamp.get_lr(i, 5e-3, 0.51, 100, 0.1) == float(g["psgld_lr"][i])


def test_sgld_steps():
    g = load_golden("sampler_steps")
    for i in range(4):
        for nm in ("U", "logsn"):
            out = osamp.sgld_step(g[f"sgld_{nm}"][i], g[f"sgld_g{nm}"][i], float(g["sgld_lr"][i]), g[f"sgld_xi{nm}"][i])
            assert relerr(out, g[f"sgld_{nm}_new"][i]) < 1e-13


def test_psgld_steps():
    g = load_golden("sampler_steps")
    V = {"U": np.zeros((25, 2)), "logsn": np.zeros(2)}
    for i in range(4):
        for nm in ("U", "logsn"):
            out, V[nm] = osamp.psgld_step(g[f"psgld_{nm}"][i], g[f"psgld_g{nm}"][i], V[nm], float(g["psgld_lr"][i]), 0.99, 1e-8,
                                          g[f"psgld_xi{nm}"][i])
            assert relerr(out, g[f"psgld_{nm}_new"][i]) < 1e-12
    assert relerr(V["U"], g["psgld_VU_final"]) < 1e-13


def test_asghmc_steps_incl_resample_and_stale_tau_inv():
    g = load_golden("sampler_steps")
    burn, k = int(g["asghmc_burn"]), int(g["asghmc_resample_every"])
    st = {"U": osamp.asghmc_init(np.zeros((25, 2))), "logsn": osamp.asghmc_init(np.zeros(2))}
    for i in range(7):
        for nm in ("U", "logsn"):
            out, st[nm] = osamp.asghmc_step(g[f"asghmc_{nm}"][i], g[f"asghmc_g{nm}"][i], st[nm], 1e-2, 5e-2, 1e-5, i < burn, k,
                                            g[f"asghmc_xi{nm}"][i], g[f"asghmc_xr{nm}"][i])
            assert relerr(out, g[f"asghmc_{nm}_new"][i]) < 1e-12
    for key in ("tau", "g", "v_hat", "momentum"):
        assert relerr(st["U"][key], g[f"asghmc_{key}_U_final"]) < 1e-12


def test_cyclical_schedule_and_gated_steps():
    """cSGLD / acSGHMC (langevin.py:1600-1724, hamiltonian.py:167-326): schedule bit-exact, updates to 1e-12."""
    g = load_golden("cyclical_steps")
    n_it, M, beta = int(g["num_iters"]), int(g["M"]), float(g["beta"])
    burn, k = int(g["acsghmc_burn"]), int(g["acsghmc_resample_every"])
    st = {"U": osamp.asghmc_init(np.zeros((25, 2))), "logsn": osamp.asghmc_init(np.zeros(2))}
    for i in range(n_it):
        r = osamp.cyclical_r(i, n_it, M)
        assert r == float(g["csgld_r"][i]) == float(g["acsghmc_r"][i])
        assert osamp.cyclical_lr(i, 2e-4, n_it, M) == float(g["csgld_lr"][i])
        assert osamp.cyclical_lr(i, 1e-2, n_it, M) == float(g["acsghmc_lr"][i])
        for nm in ("U", "logsn"):
            xi = g[f"csgld_xi{nm}"][i] if r > beta else None
            out = osamp.sgld_step(g[f"csgld_{nm}"][i], g[f"csgld_g{nm}"][i], float(g["csgld_lr"][i]), xi)
            assert relerr(out, g[f"csgld_{nm}_new"][i]) < 1e-13
            xi = g[f"acsghmc_xi{nm}"][i] if r > beta else None
            out, st[nm] = osamp.asghmc_step(g[f"acsghmc_{nm}"][i], g[f"acsghmc_g{nm}"][i], st[nm], float(g["acsghmc_lr"][i]), 5e-2,
                                            1e-5, i < burn, k, xi, g[f"acsghmc_xr{nm}"][i])
            assert relerr(out, g[f"acsghmc_{nm}_new"][i]) < 1e-12
    for key in ("tau", "g", "v_hat", "momentum"):
        assert relerr(st["U"][key], g[f"acsghmc_{key}_U_final"]) < 1e-12


def test_rbf_kernel_and_phi():
    g = load_golden("svgd")
    for n in (64, 257):
        X, S = g[f"n{n}_X"], g[f"n{n}_S"]
        K, gam = osamp.rbf_kernel(X, X)
        assert relerr(K, g[f"n{n}_K"]) < 1e-10
        assert relerr(np.median(osamp.sq_dists(X, X)), g[f"n{n}_median"]) < 1e-10
        assert relerr(osamp.svgd_phi(X, S), g[f"n{n}_phi"]) < 1e-9
        Kf, _ = osamp.rbf_kernel(X, X, sigma=0.7)
        assert relerr(Kf, g[f"n{n}_K_sigma07"]) < 1e-10
        rows = np.arange(5, 20)
        assert relerr(osamp.svgd_phi(X, S, rows=rows), g[f"n{n}_phi"][rows]) < 1e-9


def test_hamcmc_oracle_matches_reference_run():
    """113 reference iterations (100 SGLD warm-up, 5 history-filling, 8 metric steps) incl. the u = scalar + vector quirk."""
    g = load_golden("hamcmc")
    memory = int(g["memory"])
    M = memory + 1
    orc = osamp.HAMCMC(memory=memory, H_gamma=1.0, trust_reg=1.0)
    for i in range(g["grad"].shape[0]):
        if i < 2 * M - 1 + 100:
            new = orc.step_without_metric(g["theta"][i], g["grad"][i], float(g["lr"][i]), g["xi"][i], add_params=(i >= 100))
        else:
            new = orc.step(g["grad"][i], float(g["lr"][i]), g["xi"][i])
        assert relerr(new, g["theta"][i + 1]) < 1e-10, i
    assert len(orc.s) == int(g["n_pairs"])


def test_hamcmc_contiguous_oracle_matches_reference_runs():
    """HAMCMC2 / HAMCMC3 / HAMCMC4 (langevin.py:1109-1470): M warm-up steps, then 10 metric steps of each variant recorded from the
    reference; the three differ only in which stored samples form the (s, y) pairs and in the base point of the update."""
    g = load_golden("hamcmc_contiguous")
    memory = int(g["memory"])
    M = memory + 1
    finals = {}
    for variant in (2, 3, 4):
        orc = osamp.HAMCMCContiguous(variant, memory=memory, H_gamma=1.0, trust_reg=1.0)
        th, gr, xi, lr = (g["%s%d" % (k, variant)] for k in ("theta", "grad", "xi", "lr"))
        for i in range(gr.shape[0]):
            new = orc.step_without_metric(th[i], gr[i], float(lr[i]), xi[i]) if i < M else orc.step(gr[i], float(lr[i]), xi[i])
            assert relerr(new, th[i + 1]) < 1e-10, (variant, i)
        assert len(orc.params) == M and len(orc.s) == (M - 1 if variant == 4 else M - 2)      # the reference's own asserts
        finals[variant] = new
    # the variants are genuinely different samplers, and none of them is the wrong restatement of another
    g2 = osamp.HAMCMCContiguous(3, memory=memory)
    th, gr, xi, lr = (g["%s2" % k] for k in ("theta", "grad", "xi", "lr"))
    bad = 0.0
    for i in range(gr.shape[0]):
        new = g2.step_without_metric(th[i], gr[i], float(lr[i]), xi[i]) if i < M else g2.step(gr[i], float(lr[i]), xi[i])
        bad = max(bad, relerr(new, th[i + 1]))
    assert bad > 1e-6


def test_hamcmc_contig_kernel_ring_bookkeeping_mirror():
    """csrc/hamcmc_contig.cu keeps the M-entry history and the pair list as RINGS (head / pair_head indices in `meta`) instead of
    the reference's Python lists.  This is a line-by-line float64 NumPy mirror of the kernel's two modes -- same ring arithmetic,
    same order of operations -- run on the reference's own HAMCMC2 / HAMCMC3 / HAMCMC4 runs: it pins the index logic of the kernel
    (which entry is the base point, which two entries form the joining pair, which slots are overwritten) on the CPU."""
    g = load_golden("hamcmc_contiguous")
    memory = int(g["memory"]); M = memory + 1; d = 10
    H_gamma, trust_reg = 1.0, 1.0

    def run(variant):
        th_ref, gr, xi, lrs = g["theta%d"%variant], g["grad%d"%variant], g["xi%d"%variant], g["lr%d"%variant]
        ht = np.zeros((M, d)); hg = np.zeros((M, d)); ps = np.zeros((M-1, d)); py = np.zeros((M-1, d))
        U = np.zeros((M-1, d)); V = np.zeros((M-1, d)); Pp = np.zeros((M-1, d)); Q = np.zeros((M-1, d))
        meta = [0, 0, 0, 0]
        th = th_ref[0].copy()
        worst = 0.0
        for it in range(gr.shape[0]):
            lr = float(lrs[it]); gg = gr[it]; nscale = 1.0/np.sqrt(0.5*lr); noise = xi[it]*nscale
            n_hist0, head, K0, phead = meta
            if it < M:   # mode 0
                n_hist, K = n_hist0, K0
                store = n_hist < M
                t = th + (-lr)*gg; t = t + (-lr)*noise; th = t.copy()
                if store:
                    ht[n_hist] = t; hg[n_hist] = gg; n_hist += 1
                if store and n_hist == M:
                    first = 1 if variant == 2 else 0
                    K = M-1 if variant == 4 else M-2
                    for i in range(K):
                        s = ht[first+i+1] - ht[first+i]; y = hg[first+i+1] - hg[first+i] + trust_reg*s
                        ps[i] = s; py[i] = y
                meta = [n_hist, 0, K, 0]
            else:        # mode 1
                K = K0
                newest = (head + M - 1) % M; prev = (head + M - 2) % M
                B0 = 1.0/H_gamma; C0 = np.sqrt(B0); S0 = 1.0/np.sqrt(B0)
                base = ht[head if variant == 2 else newest].copy()
                nu = 0
                for i in range(K):
                    s = ps[(phead+i) % K]; y = py[(phead+i) % K]
                    sy = s @ y
                    if sy < 0: continue
                    if nu == 0: z = B0*s
                    else:
                        z = s.copy()
                        for j in range(nu-1, -1, -1):
                            c = z @ V[j]; z = z - c*U[j]
                        z = z*C0*C0
                        for j in range(nu):
                            c = z @ U[j]; z = z - c*V[j]
                    sBs = s @ z
                    cq = np.sqrt(sy/sBs); cu = np.sqrt(sBs/sy)
                    Q[nu] = cq*z - y; Pp[nu] = s/sy; U[nu] = cu + z; V[nu] = s/sBs
                    nu += 1
                z = gg.copy(); z2 = S0*noise
                if nu == 0: z = z/B0
                else:
                    for j in range(nu-1, -1, -1):
                        c = z @ Q[j]; z = z - c*Pp[j]
                    z = z*S0*S0
                    for j in range(nu):
                        c = z @ Pp[j]; z = z - c*Q[j]
                for j in range(nu):
                    c = z2 @ Pp[j]; z2 = z2 - c*Q[j]
                tn, gn, tp, gp = ht[newest].copy(), hg[newest].copy(), ht[prev].copy(), hg[prev].copy()
                t = base + (-lr)*z; t = t + (-lr)*z2
                if variant == 3: s = tn - tp; y = gn - gp + trust_reg*s
                else: s = t - tn; y = gg - gn + trust_reg*s
                th = t.copy()
                if K > 0: ps[phead] = s; py[phead] = y
                ht[head] = t; hg[head] = gg
                meta = [n_hist0, (head+1) % M, K, (phead+1) % K if K > 0 else 0]
            err = np.abs(th - th_ref[it+1]).max() / max(1.0, np.abs(th_ref[it+1]).max())
            worst = max(worst, err)
        return worst, meta

    for variant in (2, 3, 4):
        worst, meta = run(variant)
        assert worst < 1e-10, (variant, worst)
        assert meta[0] == M and meta[2] == (M - 1 if variant == 4 else M - 2)


def test_mala_oracle_explains_reference_decisions():
    """langevin.py:57-95: the oracle's aliased-state ratio reproduces every accept/reject decision of the reference run in the
    fixture; the textbook ratio (aliased=False) does not (that is the quirk the fixture pins)."""
    g = load_golden("mala_steps")
    lr = float(g["lr"])
    flat = lambda a, b, i: np.concatenate([g[a][i].ravel(), g[b][i].ravel()])[None]
    agree_aliased, agree_exact = 0, 0
    for i in range(len(g["accepted"])):
        args = (flat("A", "B", i), flat("A_new", "B_new", i), flat("gA", "gB", i), flat("gA_new", "gB_new", i),
                np.array([g["loss"][i]]), np.array([g["loss_new"][i]]), lr)
        lu = np.array([g["logu"][i]])
        agree_aliased += bool(osamp.mala_accept(osamp.mala_log_alpha(*args, aliased=True), lu)[0]) == bool(g["accepted"][i])
        agree_exact += bool(osamp.mala_accept(osamp.mala_log_alpha(*args, aliased=False), lu)[0]) == bool(g["accepted"][i])
        # the proposal is the SGLD update with the replayed noise, and a rejection restores nothing
        assert np.abs(osamp.sgld_step(g["A"][i], g["gA"][i], lr, g["xiA"][i]) - g["A_new"][i]).max() < 1e-13
        assert np.array_equal(g["A_after"][i], g["A_new"][i]) and np.array_equal(g["B_after"][i], g["B_new"][i])
    assert agree_aliased == len(g["accepted"])
    assert agree_exact < len(g["accepted"])
