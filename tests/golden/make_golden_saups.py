"""Golden vectors for the state-averaged UPS wave function (reference sa_ups_wavefunction.py), produced by RUNNING THE
REFERENCE in the build container:

    python tests/golden/make_golden_saups.py        ->  tests/golden/golden_saups.npz

The reference's two SA-UPS tests (tests/test_unitary_product_state.py:156-242): H2/STO-3G (2,2) with three states,
one-step BFGS with orbital optimisation; H3+/STO-3G (2,3) with three states, two-step BFGS.  Stored at the converged
(theta, c_mo): AO integrals, ci_coeffs of all states, SA rdm1 / rdm2, SA energy, state energies, excitation energies,
transition dipoles, oscillator strengths; and at perturbed thetas: the SA energy, the analytic theta + kappa gradient
and RotoSolve shifted energies.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
stub = tempfile.mkdtemp(prefix="pyscf_stub_")
os.makedirs(os.path.join(stub, "pyscf", "gto"))
open(os.path.join(stub, "pyscf", "__init__.py"), "w").write("from . import gto\n")
open(os.path.join(stub, "pyscf", "gto", "__init__.py"), "w").write("from . import mole\n")
open(os.path.join(stub, "pyscf", "gto", "mole.py"), "w").write("class Mole:\n    pass\n")
sys.path.insert(0, stub)
sys.path.insert(0, "/root/reference")

import slowquant.SlowQuant as sq  # noqa: E402
from slowquant.unitary_coupled_cluster.sa_ups_wavefunction import WaveFunctionSAUPS  # noqa: E402

s2 = 2 ** (-1 / 2)
CASES = {
    "h2": dict(
        geom="""H  0.0  0.0  0.0;
            H  0.0  0.0  0.735;""",
        charge=0, cas=(2, 2), two_step=False, options={"n_layers": 1, "skip_last_singles": True},
        states=([[1], [s2, -s2], [1]], [["1100"], ["1001", "0110"], ["0011"]]),
    ),
    "h3": dict(
        geom="""H   -0.45  -0.3897114317  0.0;
           H   0.45  -0.3897114317  0.0;
           H   0.0  0.3897114317  0.0;""",
        charge=1, cas=(2, 3), two_step=True, options={"n_layers": 2, "skip_last_singles": True},
        states=([[1], [s2, -s2], [s2, -s2]], [["110000"], ["100100", "011000"], ["100001", "010010"]]),
    ),
}
out = {}
for name, c in CASES.items():
    SQobj = sq.SlowQuant()
    SQobj.set_molecule(c["geom"], distance_unit="angstrom", molecular_charge=c["charge"])
    SQobj.set_basis_set("STO-3G")
    SQobj.init_hartree_fock()
    SQobj.hartree_fock.run_restricted_hartree_fock()
    WF = WaveFunctionSAUPS(c["cas"], SQobj.hartree_fock.mo_coeff, SQobj, c["states"], "tUPS", ansatz_options=dict(c["options"]), include_active_kappa=True)
    if c["two_step"]:
        WF.run_wf_optimization_2step("BFGS", True)
    else:
        WF.run_wf_optimization_1step("BFGS", True)
    pre = name + "_"
    out[pre + "cas"] = np.array(c["cas"], dtype=np.int64)
    out[pre + "num_elec"] = np.array(SQobj.molecule.number_electrons, dtype=np.int64)
    out[pre + "h_ao"] = np.array(SQobj.integral.kinetic_energy_matrix + SQobj.integral.nuclear_attraction_matrix)
    out[pre + "eri_ao"] = np.array(SQobj.integral.electron_repulsion_tensor)
    out[pre + "dipole_ao"] = np.array([SQobj.integral.get_multipole_matrix(np.array(v)) for v in ([1, 0, 0], [0, 1, 0], [0, 0, 1])])
    out[pre + "c_mo"] = np.array(WF.c_mo)
    out[pre + "thetas"] = np.array(WF.thetas)
    out[pre + "csf"] = np.array(WF.csf_coeffs)
    out[pre + "ci"] = np.array(WF.ci_coeffs)
    out[pre + "rdm1"] = np.array(WF.rdm1)
    out[pre + "rdm2"] = np.array(WF.rdm2)
    WF._sa_energy = None
    out[pre + "sa_energy"] = np.array(WF.sa_energy)
    out[pre + "energy_states"] = np.array(WF.energy_states)
    out[pre + "excitation_energies"] = np.array(WF.excitation_energies)
    dip = WF.int_gen.electric_dipole
    out[pre + "transition_dipoles"] = np.array([WF.get_transition_property(dip[k]) for k in range(3)])
    out[pre + "oscillator_strengths"] = np.array(WF.get_oscillator_strenghts())
    # perturbed parameters: energy, gradient (kappa then theta), RotoSolve shifted energies
    rng = np.random.default_rng(17)
    th = (np.array(WF.thetas) + rng.uniform(-0.3, 0.3, len(WF.thetas))).tolist()
    params = [0.0] * len(WF.kappa_idx) + th
    out[pre + "pert_thetas"] = np.array(th)
    WF._old_opt_parameters = np.zeros(len(params)) + 10**20
    out[pre + "pert_energy"] = np.array(WF._calc_energy_optimization(params, True, True))
    out[pre + "pert_gradient"] = np.array(WF._calc_gradient_optimization(params, True, True))
    WF._old_opt_parameters = np.zeros(len(params)) + 10**20
    out[pre + "pert_energy_states"] = np.array(WF._calc_energy_optimization(params, True, True, return_all_states=True))
    idx = len(th) - 1
    R = WF.ups_layout.grad_param_R[WF.ups_layout.param_names[idx]]
    shifts = [2 * mu / (2 * R + 1) * np.pi for mu in range(-R, R + 1)]
    out[pre + "rs_idx"] = np.array(idx)
    out[pre + "rs_shifts"] = np.array(shifts)
    out[pre + "rs_energies"] = np.array(WF._calc_energy_rotosolve_optimization(th, shifts, idx))
    print(name, out[pre + "excitation_energies"], out[pre + "oscillator_strengths"])
np.savez_compressed(os.path.join(HERE, "golden_saups.npz"), **out)
