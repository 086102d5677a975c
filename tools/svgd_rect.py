#!/usr/bin/env python
"""Developer driver: ONE rank's share of a G-GPU SVGD interaction on a single GPU -- nr local rows against nc = G * nr gathered
columns (row block `rank`) -- timed segment by segment (each C-ABI call sequence in its own CUDA graph, replayed back to back).
Tells where a weak-scaling step spends its time without paying for a G-GPU box."""
import argparse, os, sys
import ctypes as C
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200 import _lib
from bayesian_ode_b200.samplers.stein import _Workspace

ap = argparse.ArgumentParser()
ap.add_argument("--nr", type=int, default=4096)
ap.add_argument("--world", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--d", type=int, default=52)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
lib = _lib.load()
rng = np.random.default_rng(0)
d, nr = a.d, a.nr
for world in a.world:
    nc = nr * world
    X = torch.from_numpy((rng.standard_normal((nc, d)) * 0.3 + 1.5).astype(np.float32)).cuda()
    G = torch.from_numpy((rng.standard_normal((nc, d)) * 3).astype(np.float32)).cuda()
    Xr = X[:nr]
    ws = _Workspace(nr, nc, d, X.device)
    phi = torch.empty(nr, d, device="cuda")
    xr, xs = _lib.rows(Xr, d); xc, xcs = _lib.rows(X, d); gc, gcs = _lib.rows(G, d)
    total = nr * nc                       # the block's own median (a real run counts all ranks' blocks)

    def phi_stage(st):
        _lib.check(lib.bode_svgd_phi_staged(int(st), xr, xs, nr, xc, xcs, gc, gcs, -1.0, nc, d, nc, _lib.ptr(ws.med_gamma),
                                            C.c_void_p(ws.base.data_ptr()), _lib.ptr(phi), d, None, 0, 0.0, _lib.stream_ptr()))
    segs = {
        "prep_x (column operands)": lambda: ws.sqdist(Xr, nr, X, nc, d, total, row_offset=0, stages=_lib.SVGD_PREPARE),
        "gram2": lambda: ws.sqdist(Xr, nr, X, nc, d, total, row_offset=0, stages=_lib.SVGD_COMPUTE),
        "median": lambda: ws.median(nr, nc, d, nc),
        "prep_v": lambda: phi_stage(_lib.SVGD_PREPARE),
        "phi2": lambda: phi_stage(_lib.SVGD_COMPUTE),
    }
    for _ in range(3):
        for f in segs.values():
            f()
    torch.cuda.synchronize()
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
    s = 0.0
    print("nr=%d nc=%d (world %d) d=%d  d2 block %.0f MB" % (nr, nc, world, d, nr * nc * 4 / 1e6))
    for k, v in tot.items():
        v.sort()
        s += v[len(v) // 2]
        print("  %-28s median %7.1f us" % (k, v[len(v) // 2]))
    print("  %-28s        %7.1f us" % ("sum", s))
    del ws, graphs
