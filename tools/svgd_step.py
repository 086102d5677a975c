#!/usr/bin/env python
"""Developer driver: a few SVGD interaction steps (sqdist + median + phi) at n particles, for ncu / timing."""
import argparse, os, sys
import ctypes as C
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200.samplers.stein import _Workspace

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--d", type=int, default=52)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
lib = bode._lib.load()
rng = np.random.default_rng(0)
n, d = a.n, a.d
X = torch.from_numpy((rng.standard_normal((n, d)) * 0.3 + 1.5).astype(np.float32)).cuda()
G = torch.from_numpy((rng.standard_normal((n, d)) * 3).astype(np.float32)).cuda()
ws = _Workspace(n, n, d, X.device)
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
if a.graph:
    # per-segment timing: each C-ABI call sequence captured in its own graph and replayed back to back (warm L2, like a real step)
    segs = {"sqdist (colmean+prep+gram)": lambda: ws.sqdist(X, n, X, n, d, n * n, row_offset=0),
            "median (window+fallback+gamma)": lambda: ws.median(n, n, d, n),
            "phi (prep_v+phi+combine)": lambda: bode._lib.check(lib.bode_svgd_phi(xr, xs, n, xr, xs, gr, gs, -1.0, n, d, n, bode._lib.ptr(ws.med_gamma),
                                        C.c_void_p(ws.base.data_ptr()), bode._lib.ptr(phi), d, None, 0, 0.0, bode._lib.stream_ptr()))}
    graphs = {}
    for k, f in segs.items():
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            f()
        graphs[k] = g
    tot = {k: [] for k in segs}
    for it in range(a.iters):
        for k, g in graphs.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); e1.synchronize()
            tot[k].append(e0.elapsed_time(e1) * 1e3)
    for k, v in tot.items():
        v.sort()
        print("  %-34s median %.1f us" % (k, v[len(v) // 2]))
run = step
if a.graph:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    run = g.replay
e = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
e[0].record()
for i in range(a.iters):
    run()
    e[i + 1].record()
torch.cuda.synchronize()
ts = sorted(e[i].elapsed_time(e[i + 1]) for i in range(a.iters))
print("n=%d d=%d %s: median %.1f us, min %.1f us per interaction step" % (n, d, "graph" if a.graph else "eager", ts[len(ts) // 2] * 1e3, ts[0] * 1e3))
