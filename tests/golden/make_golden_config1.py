"""Golden vectors for config 1 (BASELINE.json configs[0]): tUPS energy + naive linear response, produced by
RUNNING THE REFERENCE ITSELF in the build container (pure Python + numba; it does not travel to the GPU box).

    python tests/golden/make_golden_config1.py        ->  tests/golden/golden_config1.npz

Two molecules:
  * "h2o": H2O/STO-3G tUPS(4,4) n_layers=3, orbital-optimised (BFGS, tol 1e-10), naive LR "SD"
    (geometry of reference tests/test_unitary_product_state.py:134-139; the combination BASELINE.md §2 pins);
  * "lih": LiH/STO-3G tUPS(2,2) n_layers=1 skip_last_singles, the reference's own test_ups_naivelr
    (tests/test_unitary_product_state.py:13-61, excitation energies +-1e-4, oscillator strengths +-1e-3).
Everything needed to rebuild the wave function at FIXED (theta, c_mo) is stored: AO integrals, dipole integrals,
the optimised MO coefficients and thetas; and everything to compare: ci_coeffs, rdm1, rdm2, energy, the LR
A / B / Sigma / Delta matrices, excitation energies, excited-state norms, transition dipoles, oscillator strengths.
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
import slowquant.unitary_coupled_cluster.linear_response.naive as naivelr  # noqa: E402
from slowquant.unitary_coupled_cluster.ups_wavefunction import WaveFunctionUPS  # noqa: E402

MOLECULES = {
    "h2o": (
        """O   0.0  0.0           0.1035174918;
        H   0.0  0.7955612117 -0.4640237459;
        H   0.0 -0.7955612117 -0.4640237459;""",
        (4, 4),
        {"n_layers": 3},
        1e-10,
    ),
    "lih": (
        """Li 0.0           0.0  0.0;
           H  1.6717072740  0.0  0.0;""",
        (2, 2),
        {"n_layers": 1, "skip_last_singles": True},
        None,
    ),
}

out = {}
for name, (geom, cas, options, tol) in MOLECULES.items():
    SQobj = sq.SlowQuant()
    SQobj.set_molecule(geom, distance_unit="angstrom")
    SQobj.set_basis_set("STO-3G")
    SQobj.init_hartree_fock()
    SQobj.hartree_fock.run_restricted_hartree_fock()
    WF = WaveFunctionUPS(cas, SQobj.hartree_fock.mo_coeff, SQobj, "tUPS", ansatz_options=dict(options), include_active_kappa=True)
    if tol is None:
        WF.run_wf_optimization_1step("BFGS", True)
    else:
        WF.run_wf_optimization_1step("BFGS", True, tol=tol)
    LR = naivelr.LinearResponse(WF, excitations="SD")
    LR.calc_excitation_energies()
    osc = LR.get_oscillator_strength()
    tdm = LR.get_transition_dipole()
    pre = name + "_"
    out[pre + "cas"] = np.array(cas, dtype=np.int64)
    out[pre + "num_elec"] = np.array(SQobj.molecule.number_electrons, dtype=np.int64)
    out[pre + "h_ao"] = np.array(SQobj.integral.kinetic_energy_matrix + SQobj.integral.nuclear_attraction_matrix)
    out[pre + "eri_ao"] = np.array(SQobj.integral.electron_repulsion_tensor)
    out[pre + "dipole_ao"] = np.array(
        [SQobj.integral.get_multipole_matrix(np.array(v)) for v in ([1, 0, 0], [0, 1, 0], [0, 0, 1])]
    )
    out[pre + "c_mo_rhf"] = np.array(SQobj.hartree_fock.mo_coeff)
    out[pre + "c_mo"] = np.array(WF.c_mo)  # orbital-optimised: C_rhf expm(-kappa)
    out[pre + "thetas"] = np.array(WF.thetas)
    out[pre + "h_mo"] = np.array(WF.h_mo)
    out[pre + "g_mo"] = np.array(WF.g_mo)
    out[pre + "ci"] = np.array(WF.ci_coeffs)
    out[pre + "rdm1"] = np.array(WF.rdm1)
    out[pre + "rdm2"] = np.array(WF.rdm2)
    out[pre + "energy"] = np.array(WF.energy_elec)
    out[pre + "A"] = np.array(LR.A)
    out[pre + "B"] = np.array(LR.B)
    out[pre + "Sigma"] = np.array(LR.Sigma)
    out[pre + "Delta"] = np.array(LR.Delta)
    out[pre + "excitation_energies"] = np.array(LR.excitation_energies)
    out[pre + "norms"] = np.array(LR.get_excited_state_norm())
    out[pre + "transition_dipoles"] = np.array(tdm)
    out[pre + "oscillator_strengths"] = np.array(osc)
    out[pre + "num_G_q"] = np.array([len(LR.G_ops), len(LR.q_ops)], dtype=np.int64)
    print(name, "E =", float(WF.energy_elec), "exc =", LR.excitation_energies[:5])
np.savez_compressed(os.path.join(HERE, "golden_config1.npz"), **out)
