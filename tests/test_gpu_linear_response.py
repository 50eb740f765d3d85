"""Config 1 (BASELINE.json configs[0]): tUPS energy + naive linear response on the CUDA path, at the FIXED
(theta, c_mo) exported from the reference run (tests/golden/make_golden_config1.py).  Compared element-wise:
ci_coeffs, rdm1, rdm2, energy, the LR matrices A / B / Sigma, excitation energies, excited-state norms and
oscillator strengths; LiH also against the literals of the reference's own test_ups_naivelr
(tests/test_unitary_product_state.py:35-61).  Tolerances: amplitudes 1e-12, energies / RDM / matrix elements 1e-10
(north_star), eigenvalues 1e-8 (generalised eigenproblem of a matrix pair known to 1e-13)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

LIH_ENERGIES = [0.129476, 0.178749, 0.178749, 0.604681, 0.646707, 0.740632, 0.740632, 1.002914,
                2.074822, 2.137193, 2.137193, 2.455191, 2.954372]
LIH_OSC = [0.049920, 0.241184, 0.241184, 0.158045, 0.166539, 0.010379, 0.010379, 0.006256,
           0.062386, 0.128862, 0.128862, 0.046007, 0.003904]


@pytest.fixture(scope="module")
def g1():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_config1.npz"))


def _wavefunction(g, name, options):
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.ups_wavefunction import WaveFunctionUPS

    pre = name + "_"
    ints = ArrayIntegrals(g[pre + "h_ao"], g[pre + "eri_ao"], int(g[pre + "num_elec"]), dipole=tuple(g[pre + "dipole_ao"]))
    cas = tuple(int(x) for x in g[pre + "cas"])
    WF = WaveFunctionUPS(cas, g[pre + "c_mo"], ints, "tUPS", ansatz_options=dict(options), include_active_kappa=True)
    WF.thetas = g[pre + "thetas"].tolist()
    return WF


@pytest.mark.parametrize(
    "name,options", [("lih", {"n_layers": 1, "skip_last_singles": True}), ("h2o", {"n_layers": 3})]
)
def test_tups_energy_and_naive_linear_response(g1, name, options):
    from slowquant_b200 import _lib
    from slowquant_b200.linear_response.naive import LinearResponse

    pre = name + "_"
    WF = _wavefunction(g1, name, options)
    assert np.max(np.abs(WF.h_mo - g1[pre + "h_mo"])) < 1e-11
    assert np.max(np.abs(WF.g_mo - g1[pre + "g_mo"])) < 1e-11
    assert np.max(np.abs(WF.ci_coeffs - g1[pre + "ci"])) < 1e-12
    assert abs(WF.energy_elec - float(g1[pre + "energy"])) < 1e-10
    assert np.max(np.abs(WF.rdm1 - g1[pre + "rdm1"])) < 1e-10
    assert np.max(np.abs(WF.rdm2 - g1[pre + "rdm2"])) < 1e-10
    before = _lib.load().sq_launch_count()
    LR = LinearResponse(WF, excitations="SD")
    assert _lib.load().sq_launch_count() > before, "the LR build launched no CUDA kernels"
    assert [len(LR.G_ops), len(LR.q_ops)] == [int(x) for x in g1[pre + "num_G_q"]]
    for key in ("A", "B", "Sigma", "Delta"):
        assert np.max(np.abs(getattr(LR, key) - g1[pre + key])) < 1e-10, key
    LR.calc_excitation_energies()
    assert np.max(np.abs(LR.excitation_energies - g1[pre + "excitation_energies"])) < 1e-8
    assert np.max(np.abs(LR.get_excited_state_norm() - g1[pre + "norms"])) < 1e-8
    osc = LR.get_oscillator_strength()
    assert np.max(np.abs(osc - g1[pre + "oscillator_strengths"])) < 1e-8
    # transition dipoles are defined up to the sign of each eigenvector
    assert np.max(np.abs(np.abs(LR.get_transition_dipole()) - np.abs(g1[pre + "transition_dipoles"]))) < 1e-7
    if name == "lih":
        assert np.max(np.abs(LR.excitation_energies - np.array(LIH_ENERGIES))) < 1e-4
        assert np.max(np.abs(osc - np.array(LIH_OSC))) < 1e-3
    else:
        # lowest roots of the H2O tUPS(4,4) L=3 naive LR
        assert np.max(np.abs(LR.excitation_energies[:5] - np.array([0.42592102, 0.52387332, 0.58068225, 0.69052476, 0.91356175]))) < 1e-7


def test_naive_lr_refuses_an_unconverged_wave_function(g1):
    """naive.py:89-92: max|<[G, H]>| > 1e-3 raises."""
    from slowquant_b200.linear_response.naive import LinearResponse

    WF = _wavefunction(g1, "h2o", {"n_layers": 3})
    WF.thetas = [t + 0.1 for t in WF.thetas]
    with pytest.raises(ValueError, match="Large Gradient"):
        LinearResponse(WF, excitations="SD")


def test_rotosolve_energies_and_optimisation(g1):
    """_calc_energy_rotosolve_optimization (ups_wavefunction.py:1144-1194) and a 3-sweep RotoSolve run
    (optimizers.py:166-270) on H2O/STO-3G tUPS(4,4) against the reference's outputs (golden_rotosolve.npz)."""
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.ups_wavefunction import WaveFunctionUPS

    gr = np.load(os.path.join(ROOT, "tests", "golden", "golden_rotosolve.npz"))
    g0 = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    ints = ArrayIntegrals(g0["h2o_h_mo"], g0["h2o_g_mo"], num_elec=10)
    eye = np.eye(g0["h2o_h_mo"].shape[0])
    WF = WaveFunctionUPS((4, 4), eye, ints, "tUPS", {"n_layers": 2}, include_active_kappa=True)
    th = g0["tups44_thetas"].tolist()
    evals0 = WF.num_energy_evals
    for idx in (0, 4, len(th) - 1):
        shifts = gr[f"rs_idx{idx}_shifts"].tolist()
        e = WF._calc_energy_rotosolve_optimization(th, shifts, idx)
        assert np.max(np.abs(np.array(e) - gr[f"rs_idx{idx}_energies"])) < 1e-10, idx
    assert WF.num_energy_evals > evals0
    WF1 = WaveFunctionUPS((4, 4), eye, ints, "tUPS", {"n_layers": 2})
    WF1.thetas = gr["opt_start_thetas"].tolist()
    assert abs(WF1.energy_elec - float(gr["opt_start_energy"])) < 1e-10
    WF1.run_wf_optimization_1step("rotosolve", False, maxiter=3)
    # thetas the energy does not depend on are fixed by rounding noise (see make_golden_rotosolve.py): compare the energy
    assert abs(WF1.energy_elec - float(gr["opt_energy"])) < 1e-8
    assert WF1.energy_elec < float(gr["opt_start_energy"]) - 1e-3
    with pytest.raises(ValueError):
        WaveFunctionUPS((4, 4), eye, ints, "tUPS", {"n_layers": 1}, include_active_kappa=True).run_wf_optimization_1step("rotosolve", True)


@pytest.mark.parametrize(
    "variant,tag",
    [
        ("projected", "proj"),
        ("statetransfer", "st"),
        ("selfconsistent", "sc"),
        ("allstatetransfer", "allst"),
        ("allselfconsistent", "allsc"),
        ("allprojected", "allproj"),
        ("projected_statetransfer", "projst"),
    ],
)
@pytest.mark.parametrize("name,options", [("lih", {"n_layers": 1, "skip_last_singles": True}), ("h2o", {"n_layers": 3})])
def test_linear_response_parametrisations(g1, name, options, variant, tag):
    """linear_response/projected.py, statetransfer.py ("U" / "Ud" operator lists run through the fused unitary kernels)
    selfconsistent.py / allstatetransfer.py / allselfconsistent.py (extended CI spaces, do_unsafe operators) and
    allprojected.py / projected_statetransfer.py (folded q^d H q products) against the reference's matrices at the same
    fixed (theta, c_mo) (golden_lr_variants.npz)."""
    import importlib

    gv = np.load(os.path.join(ROOT, "tests", "golden", "golden_lr_variants.npz"))
    mod = importlib.import_module("slowquant_b200.linear_response." + variant)
    WF = _wavefunction(g1, name, options)
    LR = mod.LinearResponse(WF, excitations="SD")
    pre = f"{name}_{tag}_"
    for key in ("A", "B", "Sigma", "Delta"):
        assert np.max(np.abs(getattr(LR, key) - gv[pre + key])) < 1e-10, key
    LR.calc_excitation_energies()
    assert np.max(np.abs(LR.excitation_energies - gv[pre + "excitation_energies"])) < 1e-8
    assert np.max(np.abs(LR.get_excited_state_norm() - gv[pre + "norms"])) < 1e-8
    assert np.max(np.abs(LR.get_oscillator_strength() - gv[pre + "oscillator_strengths"])) < 1e-8


def test_all_statetransfer_reference_literals(g1):
    """The reference's own all-ST test (tests/test_unitary_product_state.py:64-128: LiH/STO-3G, UCCSD-optimised orbitals,
    tUPS(2,2) L=1, excitation energies and oscillator strengths +-1e-3), re-run end to end on the CUDA path: orbital
    optimisation with WaveFunctionUCC, theta optimisation with WaveFunctionUPS, then the all-ST response."""
    import contextlib
    import io

    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.linear_response.allstatetransfer import LinearResponse
    from slowquant_b200.ucc_wavefunction import WaveFunctionUCC
    from slowquant_b200.ups_wavefunction import WaveFunctionUPS

    gl = np.load(os.path.join(ROOT, "tests", "golden", "golden_lih167.npz"))
    ints = ArrayIntegrals(gl["h_ao"], gl["eri_ao"], 4, dipole=tuple(gl["dipole_ao"]))
    with contextlib.redirect_stdout(io.StringIO()):
        WF = WaveFunctionUCC((2, 2), gl["c_mo_rhf"], ints, "SD")
        WF.run_wf_optimization_1step("BFGS", True)
        WF2 = WaveFunctionUPS((2, 2), WF.c_mo, ints, "tUPS", ansatz_options={"n_layers": 1})
        WF2.run_wf_optimization_1step("BFGS", False)
        LR = LinearResponse(WF2, excitations="SD")
        LR.calc_excitation_energies()
        osc = LR.get_oscillator_strength()
    solutions = np.array([0.1851181, 0.24715136, 0.24715136, 0.6230648, 0.85960395, 2.07752209, 2.13720198, 2.13720198, 2.55113802])
    assert np.allclose(LR.excitation_energies, solutions, atol=1e-3)
    ref_osc = [0.06668878, 0.33360367, 0.33360367, 0.30588158, 0.02569977, 0.06690658, 0.13411942, 0.13411942, 0.04689274]
    assert np.max(np.abs(osc - np.array(ref_osc))) < 1e-3


def test_ucc_wavefunction_linear_response(g1):
    """The reference's UCC + LR test (tests/test_unitary_coupled_cluster.py:281-363: H4/STO-3G UCCSD(4,4), naive and
    self-consistent LR) at the reference's converged thetas: the "U"/"Ud" panels go through the matrix-free exponential of
    the non-factorised UCC; matrices element-wise against the reference, spectra also against the literals of its test."""
    import contextlib
    import io

    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.linear_response import naive, selfconsistent
    from slowquant_b200.ucc_wavefunction import WaveFunctionUCC

    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_h4_ucc.npz"))
    ints = ArrayIntegrals(g["h_ao"], g["eri_ao"], 4, dipole=tuple(g["dipole_ao"]))
    WF = WaveFunctionUCC((4, 4), g["c_mo_rhf"], ints, "SD")
    WF.thetas = g["thetas"].tolist()
    assert np.max(np.abs(WF.ci_coeffs - g["ci"])) < 1e-10
    assert abs(WF.energy_elec - float(g["energy"])) < 1e-10
    literals = {
        "naive": [0.162961, 0.418771, 0.550513, 0.585337, 0.600209, 0.602964, 0.680440, 0.705532, 0.805980, 0.843321,
                  0.923462, 1.189881, 1.512350, 1.515402],
        "sc": [0.162962, 0.385979, 0.516725, 0.585337, 0.600210, 0.602570, 0.671853, 0.705532, 0.805981, 0.843321,
               0.923462, 1.189882, 1.512350, 1.515402],
    }
    for tag, mod in (("naive", naive), ("sc", selfconsistent)):
        with contextlib.redirect_stdout(io.StringIO()):
            LR = mod.LinearResponse(WF, excitations="SD")
            LR.calc_excitation_energies()
            osc = LR.get_oscillator_strength()
        for key in ("A", "B", "Sigma", "Delta"):
            assert np.max(np.abs(getattr(LR, key) - g[f"{tag}_{key}"])) < 1e-9, (tag, key)
        assert np.max(np.abs(LR.excitation_energies - g[f"{tag}_excitation_energies"])) < 1e-8
        assert np.max(np.abs(osc - g[f"{tag}_oscillator_strengths"])) < 1e-7
        assert np.max(np.abs(LR.excitation_energies - np.array(literals[tag]))) < 1e-5
