import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200.samplers.stein import _Workspace
lib = bode._lib.load()
n, d = 4096, 52
rng = np.random.default_rng(0)
X = torch.from_numpy((rng.standard_normal((n, d)) * 0.3 + 1.5).astype(np.float32)).cuda()
G = torch.from_numpy((rng.standard_normal((n, d)) * 3).astype(np.float32)).cuda()
ws = _Workspace(n, n, d, X.device); phi = torch.empty(n, d, device="cuda")
xr, xs = bode._lib.rows(X, d); gr, gs = bode._lib.rows(G, d)
tr = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
def step():
    ws.sqdist(X, n, X, n, d, n * n, row_offset=0); ws.median(n, n, d, n)
    bode._lib.check(lib.bode_svgd_phi(xr, xs, n, xr, xs, gr, gs, -1.0, n, d, n, bode._lib.ptr(ws.med_gamma), C.c_void_p(ws.base.data_ptr()), bode._lib.ptr(phi), d, None, 0, 0.0, bode._lib.stream_ptr()))
for _ in range(3): step()
lib.bode_svgd_debug_trace.argtypes = [C.c_void_p]
lib.bode_svgd_debug_trace(C.c_void_p(tr.data_ptr()))
step(); torch.cuda.synchronize()
t = tr.cpu().numpy().reshape(64, 8)[:32]
t0 = t[0, 0]
print("stage  barR_done  comp  barM_done  sttm  loop_top | mma:barK_done  barV_done  committed")
for i in range(32):
    print(i, *(int(x - t0) for x in t[i]))
print("per-stage period (worker sttm):", np.diff(t[:, 3]).astype(int))
print("mma issue time:", (t[:, 7] - t[:, 6]).astype(int), " mma wait barK (from prev commit):", (t[1:, 5] - t[:-1, 7]).astype(int), " barV wait:", (t[:, 6] - t[:, 5]).astype(int))
print("worker barR wait:", (t[:, 0] - t[:, 4]).astype(int))
print("worker: barR->comp", (t[:, 1] - t[:, 0]).astype(int), " barM wait", (t[:, 2] - t[:, 1]).astype(int), " sttm", (t[:, 3] - t[:, 2]).astype(int))
