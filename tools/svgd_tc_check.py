#!/usr/bin/env python
"""Developer check of the tcgen05 SVGD path against the FP32-pipe path (run under `timeout` on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200.samplers.stein import _Workspace
import ctypes as C

lib = bode._lib.load()
rng = np.random.default_rng(0)
for n in (128, 257, 4096):
    d = 52
    X = torch.from_numpy((rng.standard_normal((n, d)) * 0.3 + 1.5).astype(np.float32)).cuda()
    G = torch.from_numpy((rng.standard_normal((n, d)) * 3).astype(np.float32)).cuda()
    res = {}
    for tc in (0, 1):
        lib.bode_svgd_set_tensor_cores(tc)
        ws = _Workspace(n, n, d, X.device)
        ws.sqdist(X, n, X, n, d, n * n, row_offset=0)
        ws.median(n, n, d, n)
        phi = torch.empty(n, d, device="cuda")
        xr, xs = bode._lib.rows(X, d); gr, gs = bode._lib.rows(G, d)
        bode._lib.check(lib.bode_svgd_phi(xr, xs, n, xr, xs, gr, gs, -1.0, n, d, n, bode._lib.ptr(ws.med_gamma), C.c_void_p(ws.base.data_ptr()),
                                          bode._lib.ptr(phi), d, None, 0, 0.0, bode._lib.stream_ptr()))
        torch.cuda.synchronize()
        res[tc] = (ws.d2(n, n).clone(), ws.med_gamma.clone(), phi.clone())
    d2e = float((res[0][0] - res[1][0]).abs().max() / res[0][0].abs().max())
    phe = float((res[0][2] - res[1][2]).abs().max() / res[0][2].abs().max())
    print("n=%d  d2 relerr %.2e  median %s vs %s  phi relerr %.2e" % (n, d2e, res[0][1].tolist(), res[1][1].tolist(), phe), flush=True)
    X64 = X.double().cpu().numpy()
    ref = ((X64[:, None, :] - X64[None, :, :]) ** 2).sum(-1) if n <= 512 else None
    if ref is not None:
        print("   vs float64: fp32-pipe %.2e  tensor %.2e" % (np.abs(res[0][0].cpu().numpy() - ref).max() / ref.max(), np.abs(res[1][0].cpu().numpy() - ref).max() / ref.max()))

# ---- median window (svgd_state.cuh): a second call on the SAME workspace must hit the window and return bit-identical
# order statistics to a fresh workspace (three radix passes)
lib.bode_svgd_set_tensor_cores(1)
for n in (512, 4096):
    d = 52
    X = torch.from_numpy((rng.standard_normal((n, d)) * 0.3 + 1.5).astype(np.float32)).cuda()
    ws = _Workspace(n, n, d, X.device)
    meds = []
    for it in range(3):
        Xi = X + 1e-4 * it * torch.randn_like(X)
        ws.sqdist(Xi, n, Xi, n, d, n * n, row_offset=0)
        ws.median(n, n, d, n)
        fresh = _Workspace(n, n, d, X.device)
        fresh.sqdist(Xi, n, Xi, n, d, n * n, row_offset=0)
        fresh.median(n, n, d, n)
        torch.cuda.synchronize()
        st_off = 0
        same = torch.equal(ws.med_gamma, fresh.med_gamma)
        ref = float(np.median(ws.d2(n, n).cpu().numpy()))
        meds.append((ws.med_gamma[0].item(), fresh.med_gamma[0].item(), ref, same))
    print("n=%d window path (warm workspace, fresh workspace, np.median, identical):" % n, meds, flush=True)
    # timing of the whole interaction on a warm workspace
    G = torch.randn_like(X)
    phi = torch.empty(n, d, device="cuda")
    xr, xs = bode._lib.rows(X, d); gr, gs = bode._lib.rows(G, d)
    def step():
        ws.sqdist(X, n, X, n, d, n * n, row_offset=0)
        ws.median(n, n, d, n)
        bode._lib.check(lib.bode_svgd_phi(xr, xs, n, xr, xs, gr, gs, -1.0, n, d, n, bode._lib.ptr(ws.med_gamma), C.c_void_p(ws.base.data_ptr()),
                                          bode._lib.ptr(phi), d, None, 0, 0.0, bode._lib.stream_ptr()))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("n=%d interaction (sqdist+median+phi), eager launches: %.1f us/step" % (n, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
