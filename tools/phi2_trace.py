"""Developer tool: clock64 timeline of one phi2 CTA (needs a build with the trace hooks; see git history)."""
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
tr = torch.zeros(512, dtype=torch.int64, device="cuda")
def step():
    ws.sqdist(X, n, X, n, d, n * n, row_offset=0); ws.median(n, n, d, n)
    bode._lib.check(lib.bode_svgd_phi(xr, xs, n, xr, xs, gr, gs, -1.0, n, d, n, bode._lib.ptr(ws.med_gamma), C.c_void_p(ws.base.data_ptr()), bode._lib.ptr(phi), d, None, 0, 0.0, bode._lib.stream_ptr()))
for _ in range(3): step()
lib.bode_svgd_debug_trace.argtypes = [C.c_void_p]
lib.bode_svgd_debug_trace(C.c_void_p(tr.data_ptr()))
step(); torch.cuda.synchronize()
a = tr.cpu().numpy()
t = a[:256].reshape(32, 8)
t0 = a[511]
print("kernel start -> end:", int(a[510] - t0), "cycles;  first stage top at", int(t[0, 0] - t0), " last commit at", int(t[31, 4] - t0))
print("worker: top->barR", (t[:, 1] - t[:, 0]).astype(int))
print("worker: compute", (t[:, 2] - t[:, 1]).astype(int))
print("worker: barM wait", (t[:, 3] - t[:, 2]).astype(int))
print("worker period", np.diff(t[:, 0]).astype(int))
print("mma: 8 MMAs", (t[:, 6] - t[:, 5]).astype(int))
print("mma: next-stage wait", (t[:, 7] - t[:, 6]).astype(int))
print("mma: 4 MMAs + commit", (t[:, 4] - t[:, 7]).astype(int))
print("mma period", np.diff(t[:, 4]).astype(int))
