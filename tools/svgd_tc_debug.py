import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_ode_b200 as bode
from bayesian_ode_b200.samplers.stein import _Workspace
import ctypes as C
lib = bode._lib.load()
rng = np.random.default_rng(0)
n, d = 256, 52
X = torch.from_numpy((rng.standard_normal((n, d)) * 0.3 + 1.5).astype(np.float32)).cuda()
G = torch.from_numpy((rng.standard_normal((n, d)) * 3).astype(np.float32)).cuda()
out = {}
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
    off = ((n * n * 4 + 255) // 256) * 256
    part = ws.base[off:off + n * 105 * 4].view(torch.float32).view(n, 105).clone()      # split 0
    out[tc] = (phi.clone(), part, ws.d2(n, n).clone(), ws.med_gamma.clone())
gam = float(out[0][3][1])
K = torch.exp(-gam * out[1][2].double())
mu = X.double().mean(0)
V = torch.cat([-G.double(), X.double() - mu, torch.ones(n, 1, device="cuda", dtype=torch.float64)], 1)
ref = K @ V
print("tc part[0,:6]   ", out[1][1][0, :6].tolist())
print("ref (all j)[0,:6]", ref[0, :6].tolist())
print("tc part[0,100:105]", out[1][1][0, 100:105].tolist(), "ref", ref[0, 100:105].tolist())
print("tc part[5,52:56]", out[1][1][5, 52:56].tolist(), "ref", ref[5, 52:56].tolist())
print("phi tc", out[1][0][0, :4].tolist(), "phi fp32", out[0][0][0, :4].tolist())
