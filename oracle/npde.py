"""Oracle (test infrastructure): the npde GP vector field and its posterior closure.

NumPy float64 restatement of scripts/vanderpol/gp.py, batched over P particles:
  * sq_dist / K                gp.py:41-54  (expanded ||a||^2+||b||^2-2ab form, kept)
  * KernelRegression           gp.py:56-71  (Kzz, Kzz^-1, L=chol(Kzz), Kzz^-1 L; f = K(X,Z) Kzz^-1 L U)
  * inducing grid + U0 init    gp.py:315-333
  * loss_closure               gp.py:342-353
  * synthetic data             notebooks/jai/gp.ipynb cell 3
"""
import numpy as np
from . import solvers


def sq_dist(X1, X2, ell):
    X1 = X1 / ell
    X1s = np.sum(X1 ** 2, axis=-1)[..., :, None]
    X2 = X2 / ell
    X2s = np.sum(X2 ** 2, axis=-1)[..., None, :]
    return -2.0 * X1 @ np.swapaxes(X2, -1, -2) + X1s + X2s


def K(X1, X2, sf, ell):
    return sf ** 2 * np.exp(-sq_dist(X1, X2, ell) / 2.0)


def precompute(Z, sf, ell):
    """gp.py:64-67 -> dict(Kzz, Kzzinv, L, KzzinvL)."""
    Kzz = K(Z, Z, sf, ell)
    Kzzinv = np.linalg.inv(Kzz)
    L = np.linalg.cholesky(Kzz)
    return dict(Kzz=Kzz, Kzzinv=Kzzinv, L=L, KzzinvL=Kzzinv @ L)


def inducing_grid(Y, M):
    """gp.py:315-318: M x M grid over the data bounding box, ordered xv.T.flatten()."""
    xv = np.linspace(np.min(Y[..., 0]), np.max(Y[..., 0]), M)
    yv = np.linspace(np.min(Y[..., 1]), np.max(Y[..., 1]), M)
    xg, yg = np.meshgrid(xv, yv)
    return np.array([xg.T.flatten(), yg.T.flatten()]).T


def gradient_matching_init(Y, t, Z, sf, ell):
    """gp.py:323-331: whitened gradient-matched U0."""
    D = Y.shape[-1]
    F_ = ((Y[:, 1:, :] - Y[:, :-1, :]) / (t[1] - t[0])).reshape(-1, D)
    Z_ = Y[:, :-1, :].reshape(-1, D)
    Kxz = K(Z, Z_, sf, ell)
    Kinv = np.linalg.inv(K(Z_, Z_, sf, ell) + 0.2 * np.eye(Z_.shape[0]))
    U0 = Kxz @ Kinv @ F_
    Linv = np.linalg.inv(np.linalg.cholesky(K(Z, Z, sf, ell)))
    return Linv @ U0


def vdp(x):
    """gp.py:28-30."""
    return np.concatenate([x[..., 1:2], (1 - x[..., 0:1] ** 2) * x[..., 1:2] - x[..., 0:1]], -1)


class _Closed:
    def __init__(self, fn):
        self.fn = fn

    def f(self, y):
        return self.fn(y)


def make_vdp_data(seed=0, N=5, R=3.0, T=40, t_end=7.0, noise=0.1):
    """gp.ipynb cell 3 with NumPy's Generator (the notebook used unseeded scipy.stats)."""
    rng = np.random.default_rng(seed)
    x0 = 2 * R * rng.random((N, 2)) - R
    t = np.linspace(0.0, t_end, T, dtype=np.float32)
    X = solvers.odeint_fixed(_Closed(vdp), x0[None], t, "rk4")[:, 0]          # [T,N,D]
    X = np.transpose(X, [1, 0, 2])
    Y = X + noise * rng.standard_normal(X.shape)
    return dict(N=N, R=R, noise=noise, x0=x0, t=t, X=X, Y=Y)


class NPDEField:
    """f_p(x) = K(x,Z) . KzzinvL . U_p  for every particle p (gp.py:69-71)."""

    def __init__(self, U, Z, sf, ell, pre=None):
        self.U = np.asarray(U, dtype=np.float64)          # [P,m,D]
        self.Z = np.asarray(Z, dtype=np.float64)
        self.sf, self.ell = float(sf), ell
        self.pre = pre if pre is not None else precompute(self.Z, sf, ell)
        self.W = np.einsum("jk,pkd->pjd", self.pre["KzzinvL"], self.U)   # [P,m,D]

    def f(self, y):                                       # y [P,N,D]
        Kxz = K(y, self.Z, self.sf, self.ell)             # [P,N,m]
        return np.einsum("pnj,pjd->pnd", Kxz, self.W)

    def vjp(self, y, a):
        """J^T a and dU contribution for cotangent a [P,N,D]."""
        Kxz = K(y, self.Z, self.sf, self.ell)             # [P,N,m]
        gW = np.einsum("pnj,pnd->pjd", Kxz, a)
        c = np.einsum("pnd,pjd->pnj", a, self.W) * Kxz    # [P,N,m]
        diff = (self.Z[None, None, :, :] - y[:, :, None, :]) / (np.asarray(self.ell) ** 2)
        jta = np.einsum("pnj,pnjd->pnd", c, diff)
        gU = np.einsum("jk,pjd->pkd", self.pre["KzzinvL"], gW)
        return jta, gU

    def zero_grad(self):
        return np.zeros_like(self.U)

    @staticmethod
    def add_grad(g, h):
        return g + h

    @staticmethod
    def scale_grad(g, c):
        return c * g


def nlp(U, logsn, sol, Y, Kzzinv, add_prior=True):
    """loss_closure gp.py:342-353.  sol [T,P,N,D], Y [N,T,D], logsn [P,D] -> [P]."""
    D = Y.shape[-1]
    xode = np.transpose(sol, [1, 2, 0, 3])                # [P,N,T,D]
    r2 = (Y[None] - xode) ** 2
    if not add_prior:
        return r2.sum(axis=(1, 2, 3))
    loss = (r2 / (2.0 * np.exp(logsn)[:, None, None, :] ** 2)).sum(axis=(1, 2, 3))
    loss = loss + Y.size * logsn.sum(-1) / D
    loss = loss + 0.5 * np.einsum("pjd,jk,pkd->p", U, Kzzinv, U)
    return loss


def nlp_grad(U, logsn, Z, sf, ell, x0, t, Y, method="rk4", step_size=None,
             grad_mode="discrete", scale=1.0, pre=None):
    """Full closure value + gradient wrt (U, logsn) for every particle.

    grad_mode "discrete"  == autograd through ``odeint``        
              "adjoint"   == ``odeint_adjoint`` as gp.py:26 uses
    ``scale`` multiplies loss and gradients (pSGLD divides by N, langevin.py:528).
    Returns (loss[P], gU[P,m,D], glogsn[P,D], sol[T,P,N,D]).
    """
    U = np.asarray(U, dtype=np.float64)
    logsn = np.asarray(logsn, dtype=np.float64)
    P = U.shape[0]
    field = NPDEField(U, Z, sf, ell, pre)
    y0 = np.broadcast_to(np.asarray(x0, dtype=np.float64), (P,) + x0.shape[-2:]).copy()
    sol = solvers.odeint_fixed(field, y0, t, method, step_size)
    Kzzinv = field.pre["Kzzinv"]
    loss = nlp(U, logsn, sol, Y, Kzzinv)
    xode = np.transpose(sol, [1, 2, 0, 3])                # [P,N,T,D]
    r = Y[None] - xode
    e2 = np.exp(2.0 * logsn)[:, None, None, :]
    gsol = np.transpose(-r / e2, [2, 0, 1, 3])            # dL/dsol [T,P,N,D]
    D = Y.shape[-1]
    glogsn = -(r ** 2 / e2).sum(axis=(1, 2)) + Y.size / D
    if grad_mode == "discrete":
        _, gU = solvers.odeint_fixed_backward(field, y0, t, gsol, method, step_size)
    elif grad_mode == "adjoint":
        _, gU = solvers.odeint_adjoint_backward(field, sol, t, gsol, method, step_size)
    else:
        raise ValueError(grad_mode)
    gU = gU + np.einsum("jk,pkd->pjd", 0.5 * (Kzzinv + Kzzinv.T), U)
    return scale * loss, scale * gU, scale * glogsn, sol
