"""Oracle (test infrastructure): the MLP neural-ODE field of notebooks/jai/nn.ipynb cell 4 and its closure (cell 10),
NumPy float64, batched over P particles.  theta[p] = [W1 (Hx2) | b1 | W2 (HxH) | b2 | W3 (2xH) | b3] (parameters() order).
"""
import numpy as np
from . import solvers


def dim(H):
    return H * H + 6 * H + 2


def unpack(theta, H):
    P = theta.shape[0]
    o = 0
    out = []
    for shp in ((H, 2), (H,), (H, H), (H,), (2, H), (2,)):
        n = int(np.prod(shp))
        out.append(theta[:, o:o + n].reshape((P,) + shp))
        o += n
    return out


def elu(z):
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))


def delu(z):
    return np.where(z > 0, 1.0, np.exp(np.minimum(z, 0)))


class MLPField:
    def __init__(self, theta, H):
        self.theta = np.asarray(theta, dtype=np.float64)
        self.H = H
        self.W1, self.b1, self.W2, self.b2, self.W3, self.b3 = unpack(self.theta, H)

    def _hidden(self, y):
        z1 = np.einsum("phd,pnd->pnh", self.W1, y) + self.b1[:, None]
        h1 = elu(z1)
        z2 = np.einsum("phk,pnk->pnh", self.W2, h1) + self.b2[:, None]
        h2 = elu(z2)
        return z1, h1, z2, h2

    def f(self, y):
        _, _, _, h2 = self._hidden(y)
        return np.einsum("pdh,pnh->pnd", self.W3, h2) + self.b3[:, None]

    def vjp(self, y, a):
        z1, h1, z2, h2 = self._hidden(y)
        gW3 = np.einsum("pnd,pnh->pdh", a, h2)
        gb3 = a.sum(1)
        gz2 = np.einsum("pdh,pnd->pnh", self.W3, a) * delu(z2)
        gW2 = np.einsum("pnh,pnk->phk", gz2, h1)
        gb2 = gz2.sum(1)
        gz1 = np.einsum("phk,pnh->pnk", self.W2, gz2) * delu(z1)
        gW1 = np.einsum("pnh,pnd->phd", gz1, y)
        gb1 = gz1.sum(1)
        jta = np.einsum("phd,pnh->pnd", self.W1, gz1)
        P = y.shape[0]
        g = np.concatenate([x.reshape(P, -1) for x in (gW1, gb1, gW2, gb2, gW3, gb3)], 1)
        return jta, g

    def zero_grad(self):
        return np.zeros_like(self.theta)

    @staticmethod
    def add_grad(g, h):
        return g + h

    @staticmethod
    def scale_grad(g, c):
        return c * g


def sse_grad(theta, H, x0, t, X, method="rk4", step_size=None, grad_mode="discrete", lik_w=1.0, reg=0.5, scale=1.0):
    """bayesian_closure (nn.ipynb cell 10): loss = lik_w * sum (X - x)^2 + reg * sum theta^2, and its gradient."""
    theta = np.asarray(theta, dtype=np.float64)
    P = theta.shape[0]
    field = MLPField(theta, H)
    y0 = np.broadcast_to(np.asarray(x0, dtype=np.float64), (P,) + x0.shape[-2:]).copy()
    sol = solvers.odeint_fixed(field, y0, t, method, step_size)           # [T,P,N,2]
    r = X[None] - np.transpose(sol, [1, 2, 0, 3])                         # [P,N,T,2]
    sq = (r ** 2).sum(axis=(1, 2, 3))
    loss = lik_w * sq + reg * (theta ** 2).sum(1)
    gsol = np.transpose(-2.0 * lik_w * r, [2, 0, 1, 3])
    if grad_mode == "discrete":
        _, g = solvers.odeint_fixed_backward(field, y0, t, gsol, method, step_size)
    else:
        _, g = solvers.odeint_adjoint_backward(field, sol, t, gsol, method, step_size)
    g = g + 2.0 * reg * theta
    return scale * loss, scale * g, sq, sol
