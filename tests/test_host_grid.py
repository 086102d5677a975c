"""CPU: the product's host-side grid / output-selection logic is bit-exact against the oracle restatement of
solvers.py:55-97 in the state dtype (float32), including step-size grids, clamping and reversed time."""
import numpy as np
import pytest
import torch

from oracle import solvers


def _load_grid_module():
    import bayesian_ode_b200  # noqa: F401
    from bayesian_ode_b200 import _grid
    return _grid


CASES = [
    (np.linspace(0, 7, 40, dtype=np.float32), None),
    (np.linspace(0, 7, 101), None),
    (np.array([0.0, 0.25, 0.7, 1.0]), 0.5),
    (np.linspace(0.0, 2.0, 9), 0.125),
    (np.array([0.0, 0.3, 0.31, 1.7, 2.0]), 0.25),
    (np.array([0.0, 1.0]), 0.3),                    # last grid point clamped to t[-1]
    (np.linspace(2.0, 0.0, 9), None),               # decreasing
    (np.array([1.0, 0.5, 0.2]), 0.125),
    (np.array([3.0]), None),                        # single time point (odeint_tests.py:121-151)
]


@pytest.mark.parametrize("t,h", CASES)
def test_grid_bit_exact_vs_oracle(t, h):
    _grid = _load_grid_module()
    g = _grid.build(torch.from_numpy(t), torch.float32, "cpu", step_size=h, with_adjoint=True)
    t_o, grid_o, sign_o = solvers.build_grid(t, np.float32, h)
    assert g.sign == sign_o
    assert np.array_equal(g.grid.numpy(), grid_o)
    assert np.array_equal(g.dt.numpy(), grid_o[1:] - grid_o[:-1])
    assert np.array_equal(g.obs_ptr.numpy(), solvers.output_map(t_o, grid_o))
    assert g.S == len(grid_o) - 1 and g.T == len(t)
    if len(t) > 1:
        dts, ptr = [], [0]
        for i in range(1, len(t)):
            _, gi, si = solvers.build_grid(np.array([t[i], t[i - 1]]), np.float32, h)
            assert si == -sign_o
            dts.append(gi[1:] - gi[:-1])
            ptr.append(ptr[-1] + len(gi) - 1)
        assert np.array_equal(g.adj_dt.numpy(), np.concatenate(dts))
        assert np.array_equal(g.adj_ptr.numpy(), ptr)


def test_non_monotone_asserts():
    _grid = _load_grid_module()
    with pytest.raises(AssertionError):
        _grid.build(torch.tensor([0.0, 1.0, 0.5]), torch.float32, "cpu")


def test_grid_must_hit_end_point():
    """solvers.py:83: a grid_constructor whose last point misses t[-1] asserts."""
    _grid = _load_grid_module()
    with pytest.raises(AssertionError):
        _grid.build(torch.tensor([0.0, 1.0]), torch.float32, "cpu", grid_constructor=lambda f, y, t: t * 0.5)


def test_untile_d2_matches_the_documented_tile_layout():
    """bode_svgd_d2_tiled layout (include/bode_b200.h): element (i, j) of the block lives in tile (i // 128, j // 32) at
    (i % 128, j % 32), tiles ordered [row block][column stage]."""
    import numpy as np
    import torch
    from bayesian_ode_b200.samplers.stein import untile_d2
    nr, nc = 256, 384
    want = np.arange(nr * nc, dtype=np.float32).reshape(nr, nc)
    flat = np.empty(nr * nc, dtype=np.float32)
    for i in range(nr):
        for j0 in range(0, nc, 32):
            off = ((i // 128) * (nc // 32) + j0 // 32) * 128 * 32 + (i % 128) * 32
            flat[off:off + 32] = want[i, j0:j0 + 32]
    got = untile_d2(torch.from_numpy(flat), nr, nc).numpy()
    assert np.array_equal(got, want)
