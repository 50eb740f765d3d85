"""CPU checks of the optimizer front end: the RotoSolve reconstruction and its derivative against outputs of the
reference (tests/golden/golden_rotosolve.npz), the sweep on an analytic trigonometric function, and the error
behaviour of the method dispatch (reference optimizers.py:133-157)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from slowquant_b200 import optimizers as opt


@pytest.fixture(scope="module")
def gr():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_rotosolve.npz"))


@pytest.mark.parametrize("R", [1, 2, 4])
def test_reconstruction_matches_reference(gr, R):
    xs = gr["x_vals"]
    e_single = [float(v) for v in gr[f"R{R}_e_single"]]
    e_sa = [np.array(v) for v in gr[f"R{R}_e_sa"]]
    assert np.max(np.abs(opt.reconstructed_f(xs, e_single, R) - gr[f"R{R}_f_single"])) < 1e-12
    assert np.max(np.abs(opt.reconstructed_f_derivative(xs, e_single, R) - gr[f"R{R}_df_single"])) < 1e-11
    assert np.max(np.abs(opt.reconstructed_f(xs, e_sa, R) - gr[f"R{R}_f_sa"])) < 1e-12
    assert np.max(np.abs(opt.reconstructed_f_derivative(xs, e_sa, R) - gr[f"R{R}_df_sa"])) < 1e-11


def test_reconstruction_is_exact_for_trigonometric_polynomials():
    R = 2
    f = lambda x: 0.3 + np.cos(x - 0.4) - 0.7 * np.sin(2 * x + 0.1)  # noqa: E731
    e_vals = opt.get_energy_evals(lambda p: float(f(p[0])), [0.0], 0, R)
    xs = np.linspace(-3, 3, 17)
    assert np.max(np.abs(opt.reconstructed_f(xs, e_vals, R) - f(xs))) < 1e-12
    df = -np.sin(xs - 0.4) - 1.4 * np.cos(2 * xs + 0.1)
    assert np.max(np.abs(opt.reconstructed_f_derivative(xs, e_vals, R) - df)) < 1e-10


def test_rotosolve_minimises_a_separable_function():
    # E(x, y) = cos(x - 1) + 2 cos(y + 0.5) has its minimum -3 at x = 1 - pi (wrapped), y = pi - 0.5
    calls = []

    def f(p):
        calls.append(1)
        return float(np.cos(p[0] - 1.0) + 2 * np.cos(p[1] + 0.5))

    def f_batched(p, shifts, idx):
        q = list(p)
        vals = []
        for s in shifts:
            q[idx] = s
            vals.append(f(q))
        return vals

    for batched in (None, f_batched):
        o = opt.Optimizers(f, "RotoSolve", tol=1e-10, is_silent=True)
        res = o.minimize([0.2, -0.1], extra_options={"R": {"a": 1, "b": 1}, "param_names": ["a", "b"], **({"f_rotosolve_optimized": batched} if batched else {})})
        assert res.success and abs(res.fun + 3.0) < 1e-9
        assert np.all(np.abs(res.x) <= np.pi + 1e-12)
        assert abs(np.cos(res.x[0] - 1.0) + 1.0) < 1e-9 and abs(np.cos(res.x[1] + 0.5) + 1.0) < 1e-9


def test_dispatch_errors():
    o = opt.Optimizers(lambda p: 0.0, "rotosolve")
    with pytest.raises(TypeError):
        o.minimize([0.0])
    with pytest.raises(ValueError):
        o.minimize([0.0], extra_options={"param_names": ["a"]})
    with pytest.raises(ValueError):
        o.minimize([0.0], extra_options={"R": {"a": 1}})
    with pytest.raises(ValueError):
        opt.Optimizers(lambda p: 0.0, "nonsense").minimize([0.0])
    res = opt.Optimizers(lambda p: float((p[0] - 2.0) ** 2), "bfgs", grad=lambda p: np.array([2 * (p[0] - 2.0)]), is_silent=True).minimize([0.0])
    assert abs(res.x[0] - 2.0) < 1e-6
