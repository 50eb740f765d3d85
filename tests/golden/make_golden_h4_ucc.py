"""Golden data for the reference's UCC + linear-response test (tests/test_unitary_coupled_cluster.py:281-363: H4/STO-3G
UCCSD(4,4), naive and self-consistent LR), produced by RUNNING THE REFERENCE in the build container:

    python tests/golden/make_golden_h4_ucc.py        ->  tests/golden/golden_h4_ucc.npz

AO integrals and RHF orbitals (to re-run the test end to end), the converged thetas, ci_coeffs, energy, and the A / B / Sigma
matrices, excitation energies and oscillator strengths of both parametrisations at those thetas.
"""
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
import slowquant.unitary_coupled_cluster.linear_response.selfconsistent as sclr  # noqa: E402
from slowquant.unitary_coupled_cluster.ucc_wavefunction import WaveFunctionUCC  # noqa: E402

SQobj = sq.SlowQuant()
SQobj.set_molecule(
    """H  0.0  0.0  0.0;
       H  1.8  0.0  0.0;
       H  0.0  1.5  0.0;
       H  1.8  1.5  0.0;""",
    distance_unit="angstrom",
)
SQobj.set_basis_set("STO-3G")
SQobj.init_hartree_fock()
SQobj.hartree_fock.run_restricted_hartree_fock()
WF = WaveFunctionUCC((4, 4), SQobj.hartree_fock.mo_coeff, SQobj, "SD")
WF.run_wf_optimization_1step("BFGS", False)
out = {
    "h_ao": np.array(SQobj.integral.kinetic_energy_matrix + SQobj.integral.nuclear_attraction_matrix),
    "eri_ao": np.array(SQobj.integral.electron_repulsion_tensor),
    "dipole_ao": np.array([SQobj.integral.get_multipole_matrix(np.array(v)) for v in ([1, 0, 0], [0, 1, 0], [0, 0, 1])]),
    "c_mo_rhf": np.array(SQobj.hartree_fock.mo_coeff),
    "thetas": np.array(WF.thetas),
    "ci": np.array(WF.ci_coeffs),
    "energy": np.array(WF.energy_elec),
}
for tag, mod in (("naive", naivelr), ("sc", sclr)):
    LR = mod.LinearResponse(WF, excitations="SD")
    LR.calc_excitation_energies()
    for key in ("A", "B", "Sigma", "Delta"):
        out[f"{tag}_{key}"] = np.array(getattr(LR, key))
    out[f"{tag}_excitation_energies"] = np.array(LR.excitation_energies)
    out[f"{tag}_oscillator_strengths"] = np.array(LR.get_oscillator_strength())
    print(tag, LR.excitation_energies[:4])
np.savez_compressed(os.path.join(HERE, "golden_h4_ucc.npz"), **out)
