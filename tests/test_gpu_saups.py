"""State-averaged UPS wave function (reference sa_ups_wavefunction.py) on the CUDA path, at the FIXED (theta, c_mo) of the
reference's two SA-UPS tests (tests/test_unitary_product_state.py:156-242; goldens from make_golden_saups.py): states,
state-averaged RDMs and energy, subspace diagonalisation, excitation energies, oscillator strengths, the analytic
state-averaged gradient, RotoSolve shifted energies; plus one optimisation from scratch against the literals of the
reference test (excitation energies +-1e-6, oscillator strengths +-1e-3)."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

S2 = 2 ** (-1 / 2)
CASES = {
    "h2": (([[1], [S2, -S2], [1]], [["1100"], ["1001", "0110"], ["0011"]]), {"n_layers": 1, "skip_last_singles": True}),
    "h3": (
        ([[1], [S2, -S2], [S2, -S2]], [["110000"], ["100100", "011000"], ["100001", "010010"]]),
        {"n_layers": 2, "skip_last_singles": True},
    ),
}


@pytest.fixture(scope="module")
def gs():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_saups.npz"))


def _wf(g, name):
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.sa_ups_wavefunction import WaveFunctionSAUPS

    pre = name + "_"
    states, options = CASES[name]
    ints = ArrayIntegrals(g[pre + "h_ao"], g[pre + "eri_ao"], int(g[pre + "num_elec"]), dipole=tuple(g[pre + "dipole_ao"]))
    cas = tuple(int(x) for x in g[pre + "cas"])
    return WaveFunctionSAUPS(cas, g[pre + "c_mo"], ints, states, "tUPS", dict(options), include_active_kappa=True)


def _d(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


@pytest.mark.parametrize("name", ["h2", "h3"])
def test_saups_at_reference_parameters(gs, name):
    pre = name + "_"
    WF = _wf(gs, name)
    assert np.array_equal(WF.csf_coeffs, gs[pre + "csf"])
    WF.thetas = gs[pre + "thetas"].tolist()
    assert _d(WF.ci_coeffs, gs[pre + "ci"]) < 1e-12
    assert _d(WF.rdm1, gs[pre + "rdm1"]) < 1e-10
    assert _d(WF.rdm2, gs[pre + "rdm2"]) < 1e-10
    assert abs(WF.sa_energy - float(gs[pre + "sa_energy"])) < 1e-10
    assert _d(WF.energy_states, gs[pre + "energy_states"]) < 1e-10
    assert _d(WF.excitation_energies, gs[pre + "excitation_energies"]) < 1e-10
    assert _d(WF.get_oscillator_strenghts(), gs[pre + "oscillator_strengths"]) < 1e-9
    th = gs[pre + "pert_thetas"].tolist()
    params = [0.0] * len(WF.kappa_idx) + th
    WF._old_opt_parameters = np.zeros(len(params)) + 10**20
    assert abs(WF._calc_energy_optimization(params, True, True) - float(gs[pre + "pert_energy"])) < 1e-10
    assert _d(WF._calc_gradient_optimization(params, True, True), gs[pre + "pert_gradient"]) < 1e-10
    WF._old_opt_parameters = np.zeros(len(params)) + 10**20
    assert _d(WF._calc_energy_optimization(params, True, True, return_all_states=True), gs[pre + "pert_energy_states"]) < 1e-10
    roto = WF._calc_energy_rotosolve_optimization(th, gs[pre + "rs_shifts"].tolist(), int(gs[pre + "rs_idx"]))
    assert _d(roto, gs[pre + "rs_energies"]) < 1e-10


def test_saups_h2_optimisation_reaches_reference_literals(gs):
    """tests/test_unitary_product_state.py:190-198 (started from the reference's converged orbitals)."""
    WF = _wf(gs, "h2")
    with contextlib.redirect_stdout(io.StringIO()):
        WF.run_wf_optimization_1step("BFGS", True)
    assert abs(WF.excitation_energies[0] - 0.974553) < 1e-6
    assert abs(WF.excitation_energies[1] - 1.632364) < 1e-6
    osc = WF.get_oscillator_strenghts()
    assert abs(osc[0] - 0.8706) < 1e-3 and abs(osc[1]) < 1e-3


def test_saups_input_validation(gs):
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.sa_ups_wavefunction import WaveFunctionSAUPS

    g = gs
    ints = ArrayIntegrals(g["h2_h_ao"], g["h2_eri_ao"], 2)
    c = g["h2_c_mo"]
    with pytest.raises(ValueError, match="not normalized"):
        WaveFunctionSAUPS((2, 2), c, ints, ([[0.5]], [["1100"]]), "tUPS", {"n_layers": 1})
    with pytest.raises(ValueError, match="not orthogonal"):
        WaveFunctionSAUPS((2, 2), c, ints, ([[1], [S2, S2]], [["1100"], ["1100", "0011"]]), "tUPS", {"n_layers": 1})
    with pytest.raises(ValueError, match="Mismatch"):
        WaveFunctionSAUPS((2, 2), c, ints, ([[1, 0]], [["1100"]]), "tUPS", {"n_layers": 1})
    with pytest.raises(ValueError, match="Length of determinant"):
        WaveFunctionSAUPS((2, 2), c, ints, ([[1]], [["110"]]), "tUPS", {"n_layers": 1})
    with pytest.raises(ValueError, match="perfect pairing"):
        WaveFunctionSAUPS((2, 2), c, ints, ([[1]], [["1100"]]), "tUPS", {"n_layers": 1, "do_pp": True})
    with pytest.raises(ValueError, match="unknown ansatz"):
        WaveFunctionSAUPS((2, 2), c, ints, ([[1]], [["1100"]]), "fUCCSD")


def test_saups_h3_two_step_optimisation_reaches_reference_literals(gs):
    """tests/test_unitary_product_state.py:201-242 (two-step driver, started from the reference's converged orbitals)."""
    WF = _wf(gs, "h3")
    with contextlib.redirect_stdout(io.StringIO()):
        WF.run_wf_optimization_2step("BFGS", True, is_silent_subiterations=True)
    assert abs(WF.excitation_energies[0] - 0.838466) < 1e-6
    assert abs(WF.excitation_energies[1] - 0.838466) < 1e-6
    osc = WF.get_oscillator_strenghts()
    assert abs(osc[0] - 0.7569) < 1e-3 and abs(osc[1] - 0.7569) < 1e-3
