"""Oracle (test infrastructure): NumPy restatement of torchdiffeq's adaptive Dopri5Solver for ONE state tensor
(dopri5.py:58-122, rk_common.py:22-61, misc.py:84-170, interp.py:5-65).  ``f`` maps an array to an array; the state
dtype follows ``y0`` (float64 mirrors the reference, float32 mirrors the product); t and dt are float64 like the
reference's controller.  Returns (solution [T, ...], dict(nfe, accepted, rejected))."""
import numpy as np

ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.]
BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
C_ERROR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
           -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
         187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def _norm(x):
    return np.sqrt(np.mean(np.square(x)))


def _split(x, groups):
    """Tuple states (misc.py:175-182): ``groups`` = row counts of the state tensors along axis -2 of the concatenated state."""
    if groups is None:
        return [x]
    out, o = [], 0
    for n in groups:
        out.append(x[..., o:o + n, :])
        o += n
    return out


def _initial_step(f, y0, f0, rtol, atol, dtype, groups=None):
    """misc.py:84-143; for a tuple the norms are taken per state tensor and combined as the reference does (max; h0 from the
    largest d0 / d1 ratio, :128)."""
    scale = atol + np.abs(y0) * rtol
    d0 = [_norm(a) for a in _split(y0 / scale, groups)]
    d1 = [_norm(a) for a in _split(f0 / scale, groups)]
    if max(d0) < 1e-5 or max(d1) < 1e-5:
        h0 = dtype(1e-6)
    elif groups is None:
        h0 = dtype(0.01) * d0[0] / d1[0]
    else:
        h0 = dtype(0.01) * max(a / b for a, b in zip(d0, d1))
    f1 = f(y0 + h0 * f0)
    d2 = [_norm(a) / h0 for a in _split((f1 - f0) / scale, groups)]
    if max(d1) <= 1e-15 and max(d2) <= 1e-15:
        h1 = max(dtype(1e-6), h0 * dtype(1e-3))
    else:
        h1 = (dtype(0.01) / max(d1 + d2)) ** dtype(1. / 5.)
    return min(dtype(100) * h0, h1)


def odeint_dopri5(f, y0, t, rtol=1e-7, atol=1e-9, first_step=None, safety=0.9, ifactor=10.0, dfactor=0.2,
                  max_num_steps=2 ** 31 - 1, groups=None):
    y0 = np.asarray(y0)
    dtype = y0.dtype.type
    t = np.asarray(t, dtype=np.float64)
    sign = 1.0
    if t.size > 1 and np.all(t[1:] < t[:-1]):
        t, sign = -t, -1.0
    assert t.size < 2 or np.all(t[1:] > t[:-1])
    nfe = [0]

    def func(y):
        nfe[0] += 1
        return (sign * np.asarray(f(y))).astype(dtype)

    rtol_, atol_ = dtype(rtol), dtype(atol)
    f0 = func(y0)
    if first_step is None:
        dt = float(_initial_step(func, y0, f0, rtol_, atol_, dtype, groups))
    else:
        dt = 0.01                                              # dopri5.py:81-82
    y, fcur = y0, f0
    t0 = t1 = float(t[0])
    coeff = [y0] * 5
    sol = [y0]
    acc = rej = 0
    for i in range(1, len(t)):
        next_t = float(t[i])
        n_steps = 0
        while next_t > t1:
            assert n_steps < max_num_steps
            ts = t1
            assert ts + dt > ts, "underflow in dt"
            assert np.all(np.isfinite(y))
            h = dtype(dt)
            k = [fcur]
            for beta in BETA:
                yi = y + sum((h * dtype(b)) * kj for b, kj in zip(beta, k) if b != 0)
                k.append(func(yi))
            y1, f1 = yi, k[-1]
            err = sum((h * dtype(c)) * kj for c, kj in zip(C_ERROR, k) if c != 0)
            tol = atol_ + rtol_ * np.maximum(np.abs(y), np.abs(y1))
            ratio = max(np.mean(np.square(a)) for a in _split(err / tol, groups))     # per tensor, then max (misc.py:151-161)
            accept = ratio <= 1
            if accept:
                ymid = y + sum((h * dtype(c)) * kj for c, kj in zip(C_MID, k) if c != 0)
                fa, fb = k[0], k[-1]
                coeff = [-2 * h * fa + 2 * h * fb - 8 * y - 8 * y1 + 16 * ymid,
                         5 * h * fa - 3 * h * fb + 18 * y + 14 * y1 - 32 * ymid,
                         -4 * h * fa + h * fb - 11 * y - 5 * y1 + 16 * ymid,
                         h * fa, y]
                y, fcur = y1, f1
                t0, t1 = ts, ts + dt
                acc += 1
            else:
                t0 = ts
                rej += 1
            if ratio == 0:
                dt = dt * ifactor
            else:
                dfac = 1.0 if ratio < 1 else dfactor
                er = float(np.sqrt(ratio))                     # sqrt in the ratio dtype, then to float64
                factor = max(1.0 / ifactor, min(er ** (1.0 / 5.0) / safety, 1.0 / dfac))
                dt = dt / factor
            n_steps += 1
        x = (dtype(next_t) - dtype(t0)) / (dtype(t1) - dtype(t0))
        xs = [dtype(1), x, x * x, x * x * x, x * x * x * x]
        sol.append(sum(c * xp for c, xp in zip(coeff, reversed(xs))))
    return np.stack(sol), dict(nfe=nfe[0], accepted=acc, rejected=rej)


# ------------------------------------------------------------------ gradient: discrete adjoint with FROZEN step sizes
# The reference's autograd-through-dopri5 also differentiates the step-size controller; the product (and this oracle)
# define the gradient as the exact reverse of the accepted-step recursion plus the dense-output evaluation with the
# accepted step sizes held fixed (SURVEY.md hard part 6), and validate it against odeint_adjoint(dopri5) at tight
# tolerances with the reference's own cross-check bars (neuralode_tests/gradient_tests.py:114-116).
C_SOL = BETA[-1]


def interp_weights(x):
    """out = w_k1*h*k1 + w_k7*h*k7 + w_y0*y0 + w_y1*y1 + w_mid*ymid  (interp.py:21-35 collected by operand)."""
    x2, x3, x4 = x * x, x * x * x, x * x * x * x
    return (-2 * x4 + 5 * x3 - 4 * x2 + x, 2 * x4 - 3 * x3 + x2, -8 * x4 + 18 * x3 - 11 * x2 + 1, -8 * x4 + 14 * x3 - 5 * x2,
            16 * x4 - 32 * x3 + 16 * x2)


def solve_and_grad(field, y0, t, gout_fn, rtol=1e-7, atol=1e-9, first_step=None):
    """field: object with f(y), vjp(y, a) -> (J^T a, g) on arrays shaped like y0 (leading particle/trajectory axes allowed,
    ONE controller for the whole tensor as in the reference).  gout_fn(sol) -> dL/dsol.  Returns (sol, dL/dy0, dL/dtheta)."""
    y0 = np.asarray(y0, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    sign = 1.0
    if t.size > 1 and np.all(t[1:] < t[:-1]):
        t, sign = -t, -1.0
    f = lambda y: sign * field.f(y)
    rec = []            # accepted steps: (h, stage_points[6] (y2..y6, y1), k list)
    out_step, out_x = [None], [None]
    fcur = f(y0)
    dt = 0.01 if first_step is not None else float(_initial_step(f, y0, fcur, rtol, atol, np.float64))
    y = y0
    t0 = t1 = float(t[0])
    coeff = [y0] * 5
    sol = [y0]
    for i in range(1, len(t)):
        next_t = float(t[i])
        while next_t > t1:
            ts, h = t1, dt
            k, pts = [fcur], []
            for beta in BETA:
                yi = y + sum((h * b) * kj for b, kj in zip(beta, k) if b != 0)
                pts.append(yi)
                k.append(f(yi))
            y1 = yi
            err = sum((h * c) * kj for c, kj in zip(C_ERROR, k) if c != 0)
            tol = atol + rtol * np.maximum(np.abs(y), np.abs(y1))
            ratio = np.mean(np.square(err / tol))
            if ratio <= 1:
                ymid = y + sum((h * c) * kj for c, kj in zip(C_MID, k) if c != 0)
                fa, fb = k[0], k[-1]
                coeff = [-2 * h * fa + 2 * h * fb - 8 * y - 8 * y1 + 16 * ymid, 5 * h * fa - 3 * h * fb + 18 * y + 14 * y1 - 32 * ymid,
                         -4 * h * fa + h * fb - 11 * y - 5 * y1 + 16 * ymid, h * fa, y]
                rec.append((h, y, pts))
                y, fcur = y1, k[-1]
                t0, t1 = ts, ts + dt
            else:
                t0 = ts
            if ratio == 0:
                dt = dt * 10.0
            else:
                dfac = 1.0 if ratio < 1 else 0.2
                dt = dt / max(0.1, min(float(np.sqrt(ratio)) ** 0.2 / 0.9, 1.0 / dfac))
        x = (next_t - t0) / (t1 - t0)
        xs = [1.0, x, x * x, x ** 3, x ** 4]
        sol.append(sum(c * xp for c, xp in zip(coeff, reversed(xs))))
        out_step.append(len(rec) - 1)
        out_x.append(x)
    sol = np.stack(sol)
    gout = gout_fn(sol)
    # ---- reverse sweep over accepted steps
    g = field.zero_grad()
    abar = np.zeros_like(y0)           # adjoint of the state at the END of the current step
    k7bar = np.zeros_like(y0)          # adjoint of k7 of the current step (= k1 of the next one, FSAL)
    for kidx in range(len(rec) - 1, -1, -1):
        h, ystart, pts = rec[kidx]
        kbar = [np.zeros_like(y0) for _ in range(7)]
        kbar[6] = k7bar
        y0bar = np.zeros_like(y0)
        y1bar = abar
        for i in range(1, len(t)):
            if out_step[i] == kidx:
                w1, w7, wy0, wy1, wm = interp_weights(out_x[i])
                gi = gout[i]
                kbar[0] = kbar[0] + w1 * h * gi
                kbar[6] = kbar[6] + w7 * h * gi
                y0bar = y0bar + (wy0 + wm) * gi
                y1bar = y1bar + wy1 * gi
                for j, cm in enumerate(C_MID):
                    if cm != 0:
                        kbar[j] = kbar[j] + (h * cm * wm) * gi
        # k7 = f(y1)
        v, gi_ = field.vjp(pts[5], sign * kbar[6])
        g = field.add_grad(g, gi_)
        y1bar = y1bar + v
        y0bar = y0bar + y1bar
        for j, c in enumerate(C_SOL):
            if c != 0:
                kbar[j] = kbar[j] + (h * c) * y1bar
        # stages 6..2: k_i = f(pts[i-2]), pts[i-2] = y0 + h sum_j beta[i-2][j] k_j
        for i in range(5, 0, -1):
            v, gi_ = field.vjp(pts[i - 1], sign * kbar[i])
            g = field.add_grad(g, gi_)
            y0bar = y0bar + v
            for j, b in enumerate(BETA[i - 1]):
                if b != 0:
                    kbar[j] = kbar[j] + (h * b) * v
        abar = y0bar
        k7bar = kbar[0]
        if kidx == 0:
            v, gi_ = field.vjp(ystart, sign * kbar[0])      # k1 of the first step = f(y0)
            g = field.add_grad(g, gi_)
            abar = abar + v
    return sol, abar + gout[0], g
