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


def _initial_step(f, y0, f0, rtol, atol, dtype):
    scale = atol + np.abs(y0) * rtol
    d0, d1 = _norm(y0 / scale), _norm(f0 / scale)
    h0 = dtype(1e-6) if (d0 < 1e-5 or d1 < 1e-5) else dtype(0.01) * d0 / d1
    f1 = f(y0 + h0 * f0)
    d2 = _norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(dtype(1e-6), h0 * dtype(1e-3))
    else:
        h1 = (dtype(0.01) / max(d1, d2)) ** dtype(1. / 5.)
    return min(dtype(100) * h0, h1)


def odeint_dopri5(f, y0, t, rtol=1e-7, atol=1e-9, first_step=None, safety=0.9, ifactor=10.0, dfactor=0.2,
                  max_num_steps=2 ** 31 - 1):
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
        dt = float(_initial_step(func, y0, f0, rtol_, atol_, dtype))
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
            ratio = np.mean(np.square(err / tol))
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
