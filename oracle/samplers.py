"""Oracle (test infrastructure): NumPy float64 restatement of the reference sampler updates.

  sgld_step     samplers/langevin.py:173-202
  psgld_step    samplers/langevin.py:457-500
  asghmc_step   samplers/hamiltonian.py:38-99
  get_lr        samplers/langevin.py:205-210
  rbf_kernel    samplers/stein.py:18-34  (cdist^2, median heuristic over all n*n entries)
  svgd_phi      samplers/stein.py:75-86  (closed form of the autograd expression)
``xi`` arguments are standard-normal draws (the reference's Normal(0, std).sample() == std * randn bit for bit).
"""
import numpy as np


def get_lr(t, lr0, lr_gamma, lr_t0, lr_alpha):
    return lr0 / np.power(lr_t0 + lr_alpha * t, lr_gamma)


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
