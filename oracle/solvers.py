"""Oracle (test infrastructure): fixed-grid explicit RK integrators and their gradients.

NumPy restatement of the reference's vendored torchdiffeq, batched over a leading
particle axis.  ``y`` arrays are ``[P, N, D]`` (P particles/chains, N trajectories,
D state dims); every function is dtype-generic (float64 to mirror the reference,
float32 to mirror the product's host-side selection logic bit for bit).

Follows (reference file:line):
  * grid / output selection   torchdiffeq/_impl/solvers.py:36-108
  * euler / midpoint / rk4    torchdiffeq/_impl/fixed_grid.py:5-33,
                              torchdiffeq/_impl/rk_common.py:72-78 (3/8 rule)
  * time reversal             torchdiffeq/_impl/misc.py:184-187
  * continuous adjoint        torchdiffeq/_impl/adjoint.py:23-102
The discrete adjoint is the exact reverse-mode derivative of the forward
recursion (what autograd through ``odeint`` produces).

A *field* is any object with
    f(y)            -> dy/dt                           [P,N,D]
    vjp(y, a)       -> (J(y)^T a  [P,N,D], g)          g = d(a.f)/dtheta, opaque
    zero_grad()     -> g filled with zeros
    add_grad(g, h)  -> g + h
    scale_grad(g,c) -> c * g
"""
import math
import numpy as np

# Butcher tableaus as (beta rows for stages 2..s, b weights).  rk4 is the 3/8 rule
# (fixed_grid.py:29 -> rk_common.py:72-78), NOT the classic tableau.
TABLEAUS = {
    "euler": ([], [1.0]),
    "midpoint": ([[0.5]], [0.0, 1.0]),
    "rk4": ([[1.0 / 3.0], [-1.0 / 3.0, 1.0], [1.0, -1.0, 1.0]],
            [1.0 / 8.0, 3.0 / 8.0, 3.0 / 8.0, 1.0 / 8.0]),
}


# --------------------------------------------------------------------------- grid
def build_grid(t, dtype, step_size=None):
    """solvers.py:55-68,79-84: the solver grid in the state dtype.

    Returns (grid, sign).  Decreasing ``t`` is integrated as ``-t`` with the
    negated field (misc.py:184-187) -> ``sign=-1``.
    """
    t = np.asarray(t)
    sign = 1.0
    if t.size > 1 and np.all(t[1:] < t[:-1]):
        t = -t
        sign = -1.0
    assert t.size < 2 or np.all(t[1:] > t[:-1]), "t must be strictly increasing or decrasing"
    t = t.astype(dtype)
    if step_size is None:
        grid = t.copy()
    else:
        step = np.asarray(step_size, dtype=dtype)
        niters = int(np.ceil((t[-1] - t[0]) / step + dtype(1)))
        grid = (np.arange(0, niters).astype(dtype) * step + t[0]).astype(dtype)
        if grid[-1] > t[-1]:
            grid[-1] = t[-1]
    assert grid[0] == t[0] and grid[-1] == t[-1]
    return t, grid, sign


def output_map(t, grid):
    """solvers.py:88-97: obs_ptr[s]..obs_ptr[s+1] are the output indices emitted
    after grid step s (``while j < len(t) and t1 >= t[j]``); output 0 is y0.
    Every emitted value is the END-of-step state (interpolation quirk: y0 is
    overwritten by y1 before ``_linear_interp``)."""
    S = len(grid) - 1
    ptr = np.zeros(S + 1, dtype=np.int32)
    j = 1
    for s in range(S):
        ptr[s] = j
        t1 = grid[s + 1]
        while j < len(t) and t1 >= t[j]:
            j += 1
    ptr[S] = j
    return ptr


# ------------------------------------------------------------------------ forward
def rk_stages(field, y, dt, method, sign=1.0):
    """One explicit RK step; returns (y_next, stage_points, ks)."""
    betas, b = TABLEAUS[method]
    ks, ys = [], []
    yi = y
    for i in range(len(b)):
        if i > 0:
            acc = 0.0
            for j, bij in enumerate(betas[i - 1]):
                if bij != 0.0:
                    acc = acc + bij * ks[j]
            yi = y + dt * acc
        ys.append(yi)
        ks.append(sign * field.f(yi))
    inc = 0.0
    for bi, k in zip(b, ks):
        if bi != 0.0:
            inc = inc + bi * k
    return y + dt * inc, ys, ks


def odeint_fixed(field, y0, t, method="rk4", step_size=None, return_ckpt=False):
    """odeint(func, y0, t, method in {euler, midpoint, rk4}) -> [T, P, N, D]."""
    dtype = y0.dtype.type
    t_, grid, sign = build_grid(t, dtype, step_size)
    ptr = output_map(t_, grid)
    out = [y0]
    ck = [y0]
    y = y0
    for s in range(len(grid) - 1):
        dt = grid[s + 1] - grid[s]
        y, _, _ = rk_stages(field, y, dt, method, sign)
        ck.append(y)
        for _ in range(ptr[s], ptr[s + 1]):
            out.append(y)
    assert len(out) == len(t_)
    sol = np.stack(out, 0)
    if return_ckpt:
        return sol, (grid, ptr, sign, ck)
    return sol


# ------------------------------------------------------------- discrete adjoint
def rk_step_adjoint(field, y, dt, method, abar, sign=1.0):
    """Reverse-mode derivative of one RK step.  Returns (ybar, g)."""
    betas, b = TABLEAUS[method]
    _, ys, _ = rk_stages(field, y, dt, method, sign)
    s = len(b)
    kbar = [dt * bi * abar for bi in b]
    ybar = abar.copy()
    g = field.zero_grad()
    for i in range(s - 1, -1, -1):
        yi_bar, gi = field.vjp(ys[i], sign * kbar[i])
        g = field.add_grad(g, gi)
        ybar = ybar + yi_bar
        if i > 0:
            for j, bij in enumerate(betas[i - 1]):
                if bij != 0.0:
                    kbar[j] = kbar[j] + dt * bij * yi_bar
    return ybar, g


def odeint_fixed_backward(field, y0, t, grad_out, method="rk4", step_size=None):
    """Discrete adjoint of ``odeint_fixed``: given dL/d(sol) ``[T,P,N,D]`` return
    (dL/dy0, dL/dtheta).  Equals autograd through the reference's ``odeint``
    (neuralode_tests/gradient_tests.py:19-37 gradchecks exactly this)."""
    _, (grid, ptr, sign, ck) = odeint_fixed(field, y0, t, method, step_size, return_ckpt=True)
    S = len(grid) - 1
    a = np.zeros_like(y0)
    g = field.zero_grad()
    for s in range(S - 1, -1, -1):
        for j in range(ptr[s], ptr[s + 1]):
            a = a + grad_out[j]
        dt = grid[s + 1] - grid[s]
        a, gs = rk_step_adjoint(field, ck[s], dt, method, a, sign)
        g = field.add_grad(g, gs)
    a = a + grad_out[0]
    return a, g


# ----------------------------------------------------------- continuous adjoint
def odeint_adjoint_backward(field, sol, t, grad_out, method="rk4", step_size=None):
    """adjoint.py:57-102 restated: for i = T-1 .. 1 restart from the STORED forward
    value ``sol[i]`` and integrate the augmented system (y, a_y, a_theta) from t[i]
    to t[i-1] with the same method/options, then add grad_out[i-1].
    ``a_t`` is not tracked (only needed when ``t.requires_grad``)."""
    dtype = sol.dtype.type
    t = np.asarray(t)
    T = sol.shape[0]
    a = grad_out[-1].copy()
    g = field.zero_grad()
    betas, b = TABLEAUS[method]
    for i in range(T - 1, 0, -1):
        # inner solve on [t_i, t_{i-1}] through the public odeint => own grid/reversal
        tt = np.array([t[i], t[i - 1]])
        t_, grid, sign = build_grid(tt, dtype, step_size)
        y = sol[i]
        for s in range(len(grid) - 1):
            dt = grid[s + 1] - grid[s]
            # augmented dynamics (adjoint.py:32-55) under sign (misc.py:186-187):
            #   F(y, a, gth) = sign * ( f(y), -J^T a, -(df/dth)^T a )
            ky, ka, kg = [], [], []
            for st in range(len(b)):
                if st == 0:
                    yi, ai = y, a
                else:
                    accy, acca = 0.0, 0.0
                    for j, bij in enumerate(betas[st - 1]):
                        if bij != 0.0:
                            accy = accy + bij * ky[j]
                            acca = acca + bij * ka[j]
                    yi, ai = y + dt * accy, a + dt * acca
                fy = field.f(yi)
                jta, gth = field.vjp(yi, ai)
                ky.append(sign * fy)
                ka.append(-sign * jta)
                kg.append(field.scale_grad(gth, -sign))
            incy, inca = 0.0, 0.0
            for bi, k1, k2, k3 in zip(b, ky, ka, kg):
                if bi != 0.0:
                    incy = incy + bi * k1
                    inca = inca + bi * k2
                    g = field.add_grad(g, field.scale_grad(k3, dt * bi))
            y = y + dt * incy
            a = a + dt * inca
        a = a + grad_out[i - 1]
    return a, g
