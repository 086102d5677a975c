"""Vector-field modules recognised by ``odeint`` and dispatched to the fused CUDA kernels.

``NPDEField`` is the particle-batched counterpart of the reference's ``KernelRegression``
(scripts/vanderpol/gp.py:56-71): f(x) = K(x, Z) . Kzz^-1 L . U  with an RBF kernel
(gp.py:41-54).  Z, sf, ell are shared by all particles; ``U`` ([P, m, 2] or [m, 2]) and
``logsn`` ([P, 2] or [2]) are the sampled parameters (gp.py:337).
"""
import math

import numpy as np
import torch

from . import _lib


def _sq_dist(X1, X2, ell):
    # gp.py:49-54 (expanded form kept for the float64 host precompute)
    X1 = X1 / ell
    X1s = torch.sum(X1 ** 2, dim=-1).unsqueeze(-1)
    X2 = X2 / ell
    X2s = torch.sum(X2 ** 2, dim=-1).unsqueeze(-2)
    return -2 * X1 @ X2.transpose(-1, -2) + X1s + X2s


def rbf_kernel(X1, X2, sf, ell):
    """gp.py:41-43."""
    return sf ** 2 * torch.exp(-_sq_dist(X1, X2, ell) / 2)


def _detect_grid(Z):
    """Return (gx, gy) when Z[a*My+b] == (gx[a], gy[b]) exactly (the layout gp.py:315-318 builds)."""
    Z = np.asarray(Z)
    m = Z.shape[0]
    gy = []
    for b in range(m):
        if b > 0 and Z[b, 0] != Z[0, 0]:
            break
        gy.append(Z[b, 1])
    My = len(gy)
    if My == 0 or m % My != 0:
        return None
    Mx = m // My
    gx = Z[::My, 0]
    gy = np.asarray(gy)
    ok = np.array_equal(Z[:, 0], np.repeat(gx, My)) and np.array_equal(Z[:, 1], np.tile(gy, Mx))
    if not ok or len(set(gx.tolist())) != Mx or len(set(gy.tolist())) != My:
        return None
    return gx, gy


class NPDEField(torch.nn.Module):
    """KernelRegression(U0, Zt, sf, ell, noise) for P particles (gp.py:56-71).

    Arguments keep the reference's names and order.  ``U0`` is ``[m, 2]`` (one chain, exactly the
    reference) or ``[P, m, 2]``.  The float64 constants (Kzz, Kzzinv, L, KzzinvL) are computed on the
    host like the reference does (gp.py:64-67, float64 by gp.py:314) and kept as attributes; the
    device copies the kernels read are fp32.
    """

    def __init__(self, U0, Zt, sf, ell, noise, device=None, stable_solve=False):
        super().__init__()
        _lib.require_cuda()
        device = torch.device(device if device is not None else "cuda")
        U0 = torch.as_tensor(U0)
        self.batched = U0.dim() == 3
        self._stable = bool(stable_solve)
        if U0.dim() not in (2, 3) or U0.shape[-1] != 2:
            raise ValueError("U0 must be [m, 2] or [P, m, 2]")
        U0 = U0 if self.batched else U0[None]
        self.P, self.m = int(U0.shape[0]), int(U0.shape[1])
        # One resident buffer theta[P, d] = [U_p (2m) | logsn_p (2)], particle-major; U and logsn are Parameters
        # that VIEW its column blocks, so the sampler updates and the SVGD interaction run on the flat buffer
        # while the reference's ``params = [kreg.U, kreg.logsn]`` (gp.py:337) keeps working.
        self.d = 2 * self.m + 2
        self.theta = torch.empty(self.P, self.d, dtype=torch.float32, device=device)
        self.theta[:, :2 * self.m] = U0.to(device=device, dtype=torch.float32).reshape(self.P, -1)
        self.theta[:, 2 * self.m:] = math.log(noise)
        self.theta_grad = torch.zeros_like(self.theta)
        self.U = torch.nn.Parameter(self.theta[:, :2 * self.m].view(self.P, self.m, 2), requires_grad=True)
        self.logsn = torch.nn.Parameter(self.theta[:, 2 * self.m:], requires_grad=True)
        self.sf = float(sf)
        ell_t = torch.as_tensor(ell, dtype=torch.float64).reshape(-1)
        self.ell = ell
        self._ell2 = (float(ell_t[0]), float(ell_t[-1]))
        Z64 = torch.as_tensor(Zt).detach().to("cpu", torch.float64)
        if Z64.shape != (self.m, 2):
            raise ValueError("Zt must be [m, 2] matching U0")
        self.Z = Z64
        ell64 = ell_t if ell_t.numel() > 1 else float(ell_t[0])
        self.Kzz = rbf_kernel(Z64, Z64, self.sf, ell64)
        self.L = torch.linalg.cholesky(self.Kzz)
        if stable_solve:
            # Kzz^-1 L == L^-T ; triangular solves stay accurate when cond(Kzz) ~ 1e14 (16x16 grids)
            eye = torch.eye(self.m, dtype=torch.float64)
            Linv = torch.linalg.solve_triangular(self.L, eye, upper=False)
            self.KzzinvL = Linv.t().contiguous()
            self.Kzzinv = Linv.t() @ Linv
        else:
            self.Kzzinv = self.Kzz.inverse()               # gp.py:65
            self.KzzinvL = torch.mm(self.Kzzinv, self.L)   # gp.py:67
        self._A = (self.sf ** 2 * self.KzzinvL).to(device, torch.float32).contiguous()
        self._AT = self._A.t().contiguous()                # coalesced reads for the W = A U projection (bode_npde_field.AT)
        self._Ksym = (0.5 * (self.Kzzinv + self.Kzzinv.t())).to(device, torch.float32).contiguous()
        self._Zdev = Z64.to(device, torch.float32).contiguous()
        self.grid_axes = _detect_grid(Z64.numpy())

    # -- plain evaluation f(t, X) with torch ops (plotting / inspection; not used by odeint) -------------
    def forward(self, t, X):
        X = torch.as_tensor(X, device=self.U.device, dtype=torch.float32)
        Kxz = rbf_kernel(X, self._Zdev, 1.0, torch.tensor(self._ell2, device=X.device, dtype=torch.float32))
        W = torch.einsum("jk,pkd->pjd", self._A, self.U)
        if X.dim() == 2:
            out = torch.einsum("nj,pjd->pnd", Kxz, W)
            return out if self.batched else out[0]
        return torch.einsum("pnj,pjd->pnd", Kxz, W)

    # -- C-ABI description -------------------------------------------------------------------------------
    def c_struct(self, U=None):
        fs = _lib.NpdeFieldStruct()
        fs.P, fs.m = self.P, self.m
        if self.grid_axes is not None and len(self.grid_axes[0]) <= 32 and len(self.grid_axes[1]) <= 32:
            gx, gy = self.grid_axes
            fs.grid_mx, fs.grid_my = len(gx), len(gy)
            for i, v in enumerate(gx):
                fs.gx[i] = float(v)
            for i, v in enumerate(gy):
                fs.gy[i] = float(v)
        else:
            fs.grid_mx = fs.grid_my = 0
        fs.ell[0], fs.ell[1] = self._ell2
        fs.Z = self._Zdev.data_ptr()
        fs.A = self._A.data_ptr()
        fs.AT = self._AT.data_ptr()
        fs.Ksym = self._Ksym.data_ptr()
        U = self.U if U is None else U
        p, stride = _lib.rows(U, 2 * self.m)
        fs.U = p.value
        fs.U_stride = stride
        return fs

    def bind_flat_grads(self):
        """Point U.grad / logsn.grad at the column blocks of the flat ``theta_grad`` buffer."""
        self.U.grad = self.theta_grad[:, :2 * self.m].view(self.P, self.m, 2)
        self.logsn.grad = self.theta_grad[:, 2 * self.m:]
        return self.theta_grad


class MLPField(torch.nn.Module):
    """The notebook's ``NN(input_size=2, hidden_size=H)`` (notebooks/jai/nn.ipynb cell 4) for P particles:
    Linear(2,H) - ELU - Linear(H,H) - ELU - Linear(H,2), weights U(-0.5, 0.5) (cell 4 ``init_normal``), biases with
    nn.Linear's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)).  Parameters are views of one flat theta[P, d] buffer in
    ``parameters()`` order: W1 [P,H,2], b1 [P,H], W2 [P,H,H], b2 [P,H], W3 [P,2,H], b3 [P,2]."""

    def __init__(self, P, hidden_size=64, input_size=2, device=None, generator=None, theta=None):
        super().__init__()
        _lib.require_cuda()
        if input_size != 2:
            raise ValueError("the fused MLP field integrates 2-D states")
        if hidden_size not in (20, 64):
            raise NotImplementedError("MLPField kernels are built for hidden_size 20 and 64")
        device = torch.device(device if device is not None else "cuda")
        H = self.H = int(hidden_size)
        self.P = int(P)
        self.batched = True
        self.d = H * H + 6 * H + 2
        self.theta = torch.empty(self.P, self.d, dtype=torch.float32, device=device)
        if theta is not None:
            self.theta.copy_(torch.as_tensor(theta).reshape(self.P, self.d))
        else:
            u = torch.rand(self.P, self.d, generator=generator, dtype=torch.float32)
            lo = torch.empty(self.d)
            shapes = self._blocks()
            for name, (o, n, _), bound in zip(shapes, shapes.values(), (0.5, 2 ** -0.5, 0.5, H ** -0.5, 0.5, H ** -0.5)):
                lo[o:o + n] = bound
            self.theta.copy_((2 * u - 1) * lo)
        self.theta_grad = torch.zeros_like(self.theta)
        for name, (o, n, shp) in self._blocks().items():
            setattr(self, name, torch.nn.Parameter(self.theta[:, o:o + n].view((self.P,) + shp), requires_grad=True))

    def _blocks(self):
        H = self.H
        out, o = {}, 0
        for name, shp in (("W1", (H, 2)), ("b1", (H,)), ("W2", (H, H)), ("b2", (H,)), ("W3", (2, H)), ("b3", (2,))):
            n = 1
            for s_ in shp:
                n *= s_
            out[name] = (o, n, shp)
            o += n
        return out

    def forward(self, t, X):
        X = torch.as_tensor(X, device=self.theta.device, dtype=torch.float32)
        if X.dim() == 2:
            X = X[None].expand(self.P, -1, -1)
        h = torch.nn.functional.elu(torch.einsum("phd,pnd->pnh", self.W1, X) + self.b1[:, None])
        h = torch.nn.functional.elu(torch.einsum("phk,pnk->pnh", self.W2, h) + self.b2[:, None])
        return torch.einsum("pdh,pnh->pnd", self.W3, h) + self.b3[:, None]

    def c_struct(self, theta=None):
        fs = _lib.MlpFieldStruct()
        fs.P, fs.H = self.P, self.H
        th = self.theta if theta is None else theta
        p, stride = _lib.rows(th, self.d)
        fs.theta, fs.theta_stride = p.value, stride
        return fs

    def bind_flat_grads(self):
        for name, (o, n, shp) in self._blocks().items():
            getattr(self, name).grad = self.theta_grad[:, o:o + n].view((self.P,) + shp)
        return self.theta_grad


# the reference's name for the same object (gp.py:56)
KernelRegression = NPDEField
