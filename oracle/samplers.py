"""Oracle (test infrastructure): NumPy float64 restatement of the reference sampler updates.

  sgld_step     samplers/langevin.py:173-202
  psgld_step    samplers/langevin.py:457-500
  asghmc_step   samplers/hamiltonian.py:38-99
  get_lr        samplers/langevin.py:205-210
  cyclical_r / cyclical_lr   samplers/langevin.py:1659-1667 == hamiltonian.py:259-267 (cSGLD / acSGHMC cosine schedule;
                the update rules are sgld_step / asghmc_step with the noise gated by r > beta, langevin.py:1648-1658,
                hamiltonian.py:250-254)
  mala_*        samplers/langevin.py:27-95   (proposal = sgld_step; acceptance ratio incl. the aliased-state quirk)
  rbf_kernel    samplers/stein.py:18-34  (cdist^2, median heuristic over all n*n entries)
  svgd_phi      samplers/stein.py:75-86  (closed form of the autograd expression)
``xi`` arguments are standard-normal draws (the reference's Normal(0, std).sample() == std * randn bit for bit).
"""
import numpy as np


def get_lr(t, lr0, lr_gamma, lr_t0, lr_alpha):
    return lr0 / np.power(lr_t0 + lr_alpha * t, lr_gamma)


def cyclical_r(t, num_iters, M):
    """Position inside the current cycle, in [0, 1): Python's floor-mod semantics matter at t = 0 (r = (L-1)/L)."""
    L = (num_iters + M) // M
    return ((t - 1) % L) / L


def cyclical_lr(t, lr0, num_iters, M):
    return lr0 / 2.0 * (np.cos(np.pi * cyclical_r(t, num_iters, M)) + 1)


def sgld_step(p, g, lr, xi=None):
    noise = 0.0 if xi is None else xi * (1.0 / np.sqrt(0.5 * lr))
    return p + (-lr) * (g + noise)


def psgld_step(p, g, V, lr, alpha, lambda_, xi=None):
    V = alpha * V + (1 - alpha) * g ** 2
    G = 1.0 / (lambda_ + np.sqrt(V))
    noise = 0.0 if xi is None else xi * (1.0 / np.sqrt(0.5 * lr))
    return p + (-lr) * (G * g + np.sqrt(G) * noise), V


def asghmc_init(p):
    return dict(iteration=0, tau=np.ones_like(p), g=np.ones_like(p), v_hat=np.ones_like(p), momentum=np.zeros_like(p))


def asghmc_step(p, grad, st, lr, mom_decay, lambda_, burn_in, resample_mom_every=50, xi=None, xi_resample=None):
    st = dict(st)
    st["iteration"] += 1
    tau, g, v_hat, mom = st["tau"].copy(), st["g"].copy(), st["v_hat"].copy(), st["momentum"].copy()
    tau_inv = 1.0 / (tau + 1.0)                               # stale: from the old tau (hamiltonian.py:70)
    if burn_in:
        tau = tau + (-tau * (g * g / (v_hat + lambda_)) + 1)
        g = g + (-g * tau_inv + tau_inv * grad)
        v_hat = v_hat + (-v_hat * tau_inv + tau_inv * grad ** 2)
    minv = 1.0 / (np.sqrt(v_hat) + lambda_)
    if (not burn_in) and resample_mom_every is not None and st["iteration"] % resample_mom_every == 0:
        mom = xi_resample * np.minimum(1.0 / minv, 1e1)
    sigma = np.sqrt(np.maximum(2.0 * lr ** 2 * mom_decay * minv - lr ** 4, 1e-16))
    mom = mom + (-(lr ** 2) * minv * grad - mom_decay * mom)
    if xi is not None:
        mom = mom + xi * sigma
    st.update(tau=tau, g=g, v_hat=v_hat, momentum=mom)
    return p + mom, st


def mala_log_alpha(theta_prev, theta_new, g_prev, g_new, loss_prev, loss_new, lr, aliased=True):
    """samplers/langevin.py:57-86: log acceptance ratio of the Langevin proposal, one value per chain (leading axis).
    ``aliased=True`` reproduces the reference as it actually runs: ``self.state[p]['data'] = p.data`` (:45) keeps a VIEW of
    the parameter, which ``p.data.add_`` (:60) then updates in place, so both proposal terms see theta_prev == theta_new;
    ``aliased=False`` is the textbook MALA ratio."""
    tp = theta_new if aliased else theta_prev
    P = theta_new.shape[0]
    rev = ((tp - theta_new + lr * g_new) ** 2).reshape(P, -1).sum(1)
    fwd = ((theta_new - tp + lr * g_prev) ** 2).reshape(P, -1).sum(1)
    return loss_prev - loss_new + (-1.0 / (4 * lr)) * rev - (-1.0 / (4 * lr)) * fwd


def mala_accept(log_alpha, log_u):
    """langevin.py:88: accepted iff log_alpha is finite and log(u) < log_alpha."""
    return np.isfinite(log_alpha) & (log_u < log_alpha)


def sq_dists(X, Y):
    """||x_i - y_j||^2 in difference form (what cdist(X, Y) ** 2 means; exact zeros on the diagonal)."""
    d = X[:, None, :] - Y[None, :, :]
    return np.einsum("ijk,ijk->ij", d, d)


def median_bandwidth(d2, n):
    h = np.median(d2) / (2 * np.log(n + 1))
    sigma = np.sqrt(h)
    return 1.0 / (1e-8 + 2 * sigma ** 2)


def rbf_kernel(X, Y, sigma=None, d2=None):
    d2 = sq_dists(X, Y) if d2 is None else d2
    gamma = median_bandwidth(d2, X.shape[0]) if sigma is None else 1.0 / (1e-8 + 2 * sigma ** 2)
    return np.exp(-gamma * d2), gamma


def svgd_phi(X, score, sigma=None, rows=None, gamma=None):
    """phi_i = (1/n)[sum_j K_ij s_j + 2 gamma sum_j K_ij (x_i - x_j)] for rows i (default all)."""
    n = X.shape[0]
    Xr = X if rows is None else X[rows]
    d2 = sq_dists(Xr, X)
    if gamma is None:
        gamma = median_bandwidth(sq_dists(X, X), n) if sigma is None else 1.0 / (1e-8 + 2 * sigma ** 2)
    K = np.exp(-gamma * d2)
    return (K @ score + 2 * gamma * (K.sum(1)[:, None] * Xr - K @ X)) / n


class HAMCMC:
    """samplers/langevin.py:619-1000 restated for ONE chain with a flat parameter vector (float64 NumPy), bug-compatible:
    history of 2M-1 (theta, grad) entries with M = memory+1 (:645), theta stored AFTER the warm-up update but the gradient
    taken BEFORE it (:954-961), (s, y) pairs i <-> i+M with y += trust_reg*s and the 1e-4 curvature filter at start-up
    (:924-935), base point params[M-1] (:970), product-form BFGS with u = sqrt(sBs/sy) + Bs -- a scalar added to a
    vector -- (:846), pair refresh with the 1e-8 filter (:871-882; a chain that starts with zero pairs never gains one)."""

    def __init__(self, memory=5, H_gamma=1.0, trust_reg=1.0):
        self.M = memory + 1
        self.H_gamma, self.trust_reg = H_gamma, trust_reg
        self.params, self.grads, self.s, self.y = [], [], [], []

    def step_without_metric(self, theta, grad, lr, xi, add_noise=True, add_params=False):
        new = theta + (-lr) * grad
        if add_noise:
            new = new + (-lr) * (xi * (1.0 / np.sqrt(0.5 * lr)))
        if add_params:
            self.params.append(new.copy())
            self.grads.append(grad.copy())
        M = self.M
        if len(self.params) >= 2 * M - 1:
            for i in range(M - 1):
                si = -self.params[i] + self.params[i + M]
                yi = -self.grads[i] + self.grads[i + M] + self.trust_reg * si
                if si @ yi > 1e-4 * (si @ si):
                    self.s.append(si)
                    self.y.append(yi)
        return new

    def vector_prod(self, grad, noise):
        B0 = 1.0 / self.H_gamma
        C0, S0 = np.sqrt(B0), 1.0 / np.sqrt(B0)
        u, v, p, q = [], [], [], []

        def Cz(z):
            z = C0 * z
            for i in range(len(u)):
                z = z - v[i] * (z @ u[i])
            return z

        def CTz(z):
            for i in reversed(range(len(u))):
                z = z - u[i] * (z @ v[i])
            return C0 * z

        def Sz(z):
            z = S0 * z
            for i in range(len(u)):
                z = z - q[i] * (z @ p[i])
            return z

        def STz(z):
            for i in reversed(range(len(u))):
                z = z - p[i] * (z @ q[i])
            return S0 * z

        for s, y in zip(self.s, self.y):
            sy = s @ y
            if sy < 0:
                continue
            Bs = B0 * s if len(u) == 0 else Cz(CTz(s))
            sBs = s @ Bs
            q.append(np.sqrt(sy / sBs) * Bs - y)
            p.append(s / sy)
            u.append(np.sqrt(sBs / sy) + Bs)           # scalar + vector, as in the reference
            v.append(s / sBs)
        Hg = grad / B0 if len(u) == 0 else Sz(STz(grad))
        return Hg, Sz(noise)

    def step(self, grad, lr, xi, add_noise=True):
        M = self.M
        base = self.params[M - 1]
        noise = xi * (1.0 / np.sqrt(0.5 * lr))
        Hg, Sn = self.vector_prod(grad, noise)
        new = base + (-lr) * Hg
        if add_noise:
            new = new + (-lr) * Sn
        self.params.append(new.copy())
        self.grads.append(grad.copy())
        si = -self.params[M - 1] + self.params[2 * M - 1]
        yi = -self.grads[M - 1] + self.grads[2 * M - 1] + self.trust_reg * si
        if si @ yi > 1e-8 * (si @ si):
            self.s.append(si)
            self.y.append(yi)
            self.s.pop(0)
            self.y.pop(0)
        self.params.pop(0)
        self.grads.pop(0)
        return new


class HAMCMCContiguous(HAMCMC):
    """HAMCMC2 / HAMCMC3 / HAMCMC4 (samplers/langevin.py:1109-1470): the variants that form (s, y) from CONTIGUOUS samples.  They
    share ``_compute_vector_prod`` with HAMCMC (``vector_prod`` above) and differ only in the window bookkeeping, restated here
    bug-compatibly for one chain (``variant`` = 2, 3 or 4; M = memory + 1, langevin.py:645):
      warm-up  M plain Langevin steps, each storing (theta AFTER the update, gradient BEFORE it) (:1180-1203, :1151-1178); when the
               M-th lands, the pairs are formed without any curvature filter: variant 2 from entries 1..M-2 -> 2..M-1 (:1173-1178),
               variant 3 from 0..M-3 -> 1..M-2 (:1355-1359), variant 4 from 0..M-2 -> 1..M-1 (:1462-1466);
      step     base point = the OLDEST stored theta for variant 2 (:1209), the NEWEST for 3 and 4 (:1368); after the update the new
               (theta, grad) is appended and ONE pair is added -- between the last two entries for variants 2 and 4 (:1131-1134,
               :1424-1427), between the two entries BEFORE the last for variant 3 (:1315-1318) -- then the oldest entry of every
               list is dropped."""

    def __init__(self, variant, memory=5, H_gamma=1.0, trust_reg=1.0):
        assert variant in (2, 3, 4)
        super().__init__(memory=memory, H_gamma=H_gamma, trust_reg=trust_reg)
        self.variant = variant

    def _pair(self, a, b):
        si = -self.params[a] + self.params[b]
        yi = -self.grads[a] + self.grads[b] + self.trust_reg * si
        self.s.append(si)
        self.y.append(yi)

    def step_without_metric(self, theta, grad, lr, xi, add_noise=True, update_metric=True):
        new = theta + (-lr) * grad
        if add_noise:
            new = new + (-lr) * (xi * (1.0 / np.sqrt(0.5 * lr)))
        if update_metric:
            self.params.append(new.copy())
            self.grads.append(grad.copy())
            M = self.M
            if len(self.params) >= M:
                if self.variant == 2:
                    for j in range(M - 2):
                        self._pair(j + 1, j + 2)
                else:
                    for i in range(M - 2 if self.variant == 3 else M - 1):
                        self._pair(i, i + 1)
        return new

    def step(self, grad, lr, xi, add_noise=True):
        base = self.params[0] if self.variant == 2 else self.params[-1]
        noise = xi * (1.0 / np.sqrt(0.5 * lr))
        Hg, Sn = self.vector_prod(grad, noise)
        new = base + (-lr) * Hg
        if add_noise:
            new = new + (-lr) * Sn
        self.params.append(new.copy())
        self.grads.append(grad.copy())
        if self.variant == 3:
            self._pair(len(self.params) - 3, len(self.params) - 2)
        else:
            self._pair(len(self.params) - 2, len(self.params) - 1)
        for lst in (self.params, self.grads, self.s, self.y):
            lst.pop(0)
        return new
