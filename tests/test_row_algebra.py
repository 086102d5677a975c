"""CPU: the algebra of the row-sliced separable field (csrc/npde_row.cuh) against the oracle's npde field.

The kernel never forms k_j(x) for the m = Mx*My inducing points of a tensor grid; it uses
    k_ab = kx_a ky_b,   q_a = sum_b ky_b W_ab,   r_a = sum_b ky_b dy_b W_ab
    f          = sum_a kx_a q_a
    (J^T a)_0  = -k0 sum_a kx_a dx_a (a . q_a)
    (J^T a)_1  = -k1 sum_a kx_a      (a . r_a)
    gW_ab     += (kx_a a) ky_b
with dx_a = c0 (x0 - gx_a), dy_b = c1 (x1 - gy_b), c_d = sqrt(log2(e) / 2) / ell_d, k_d = 2 ln2 c_d, kx = 2^-(dx^2).
This NumPy restatement of exactly those formulas (same constants as csrc/npde.cu fill_common / plan_row) must reproduce
oracle.npde.NPDEField.f / .vjp on non-square grids -- it pins the index convention Z[a*My + b] = (gx[a], gy[b]) and the signs."""
import numpy as np
import pytest

from oracle import npde


def _row_sliced(W, gx, gy, ell, x, a):
    Mx, My = len(gx), len(gy)
    LOG2E, LN2 = 1.4426950408889634074, 0.69314718055994530942
    c0, c1 = np.sqrt(0.5 * LOG2E) / ell[0], np.sqrt(0.5 * LOG2E) / ell[1]
    k0, k1 = 2.0 * LN2 * c0, 2.0 * LN2 * c1
    Wg = W.reshape(Mx, My, 2)                                  # row a of the lane, columns b
    dx = c0 * x[0] - c0 * gx
    dy = c1 * x[1] - c1 * gy
    kx, ky = np.exp2(-dx * dx), np.exp2(-dy * dy)
    q = np.einsum("b,abd->ad", ky, Wg)
    r = np.einsum("b,abd->ad", ky * dy, Wg)
    f = np.einsum("a,ad->d", kx, q)
    jta = np.array([-k0 * np.sum(kx * dx * (q @ a)), -k1 * np.sum(kx * (r @ a))])
    gW = (kx[:, None, None] * a[None, None, :]) * ky[None, :, None]
    return f, jta, gW.reshape(Mx * My, 2)


@pytest.mark.parametrize("Mx,My", [(7, 9), (16, 16), (12, 8)])
def test_row_sliced_formulas_match_oracle_field(Mx, My):
    rng = np.random.default_rng(Mx * 100 + My)
    gx, gy = np.linspace(-2.5, 2.0, Mx), np.linspace(-3.0, 3.5, My)
    Z = np.stack([np.repeat(gx, My), np.tile(gy, Mx)], 1)
    ell = np.array([0.6, 0.8])
    U = rng.standard_normal((1, Mx * My, 2))
    A = rng.standard_normal((Mx * My, Mx * My)) / (Mx * My) ** 0.5
    pre = dict(KzzinvL=A)                                      # any A: only W = A U enters the field
    fld = npde.NPDEField(U, Z, 1.0, ell, pre=pre)
    W = fld.W[0]
    for _ in range(5):
        x, a = rng.uniform(-3, 3, 2), rng.standard_normal(2)
        f_o = fld.f(x[None, None])[0, 0]
        jta_o, gU_o = fld.vjp(x[None, None], a[None, None])
        f, jta, gW = _row_sliced(W, gx, gy, ell, x, a)
        assert np.allclose(f, f_o, rtol=1e-12, atol=1e-13)
        assert np.allclose(jta, jta_o[0, 0], rtol=1e-11, atol=1e-12)
        assert np.allclose(A.T @ gW, gU_o[0], rtol=1e-11, atol=1e-12)


def test_side_gram_split_enumeration():
    """SVGD._side_gram_split (host logic): column chunks per row block for a Gram pass confined to `side_sms` SMs -- full waves of
    CTAs with as many tiles each as that allows; several ranks keep the fine 16-way split."""
    from types import SimpleNamespace
    from bayesian_ode_b200.samplers.stein import SVGD
    f = SVGD._side_gram_split
    assert f(SimpleNamespace(world=1, side_sms=40), 4096, 4096) == 5          # 5 chunks x 32 row blocks = 160 CTAs = 4 waves of 40
    assert f(SimpleNamespace(world=1, side_sms=32), 4096, 4096) == 1          # one wave of 32 CTAs
    assert f(SimpleNamespace(world=2, side_sms=40), 4096, 8192) == 16
    for side in (8, 24, 40, 64):
        for n in (256, 1000, 4096):
            c = f(SimpleNamespace(world=1, side_sms=side), n, n)
            assert 1 <= c <= 16 and c <= (n + 127) // 128
