"""Golden 3- and 4-RDMs (reference ups_wavefunction.py:478-754), produced by RUNNING THE REFERENCE in the build container
on H2O/STO-3G tUPS at the `tups44` thetas of golden.npz, CAS(4,4) for rdm3 and -- the n^8 loop is slow -- CAS(2,3) and CAS(4,3) for rdm4:

    python tests/golden/make_golden_rdm34.py        ->  tests/golden/golden_rdm34.npz
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
from slowquant.unitary_coupled_cluster.ups_wavefunction import WaveFunctionUPS  # noqa: E402

SQobj = sq.SlowQuant()
SQobj.set_molecule(
    """O   0.0  0.0           0.1035174918;
    H   0.0  0.7955612117 -0.4640237459;
    H   0.0 -0.7955612117 -0.4640237459;""",
    distance_unit="angstrom",
)
SQobj.set_basis_set("sto-3g")
SQobj.init_hartree_fock()
SQobj.hartree_fock.run_restricted_hartree_fock()
c_mo = SQobj.hartree_fock.mo_coeff
g0 = np.load(os.path.join(HERE, "golden.npz"))
out = {}
WF = WaveFunctionUPS((4, 4), c_mo, SQobj, "tUPS", ansatz_options={"n_layers": 2}, include_active_kappa=True)
WF.thetas = g0["tups44_thetas"].tolist()
out["cas44_rdm3"] = np.array(WF.rdm3)
print("rdm3 (4,4) done", flush=True)
WF2 = WaveFunctionUPS((2, 3), c_mo, SQobj, "tUPS", ansatz_options={"n_layers": 2}, include_active_kappa=True)
th = np.random.default_rng(3).uniform(-1, 1, WF2.ups_layout.n_params)
WF2.thetas = th.tolist()
out["cas23_thetas"] = th
out["cas23_ci"] = np.array(WF2.ci_coeffs)
out["cas23_rdm3"] = np.array(WF2.rdm3)
out["cas23_rdm4"] = np.array(WF2.rdm4)
# four electrons in three orbitals: a non-vanishing 4-RDM (for two electrons rdm3 = rdm4 = 0 tests the cancellations)
WF3 = WaveFunctionUPS((4, 3), c_mo, SQobj, "tUPS", ansatz_options={"n_layers": 2}, include_active_kappa=True)
th3 = np.random.default_rng(4).uniform(-1, 1, WF3.ups_layout.n_params)
WF3.thetas = th3.tolist()
out["cas43_thetas"] = th3
out["cas43_rdm3"] = np.array(WF3.rdm3)
out["cas43_rdm4"] = np.array(WF3.rdm4)
np.savez_compressed(os.path.join(HERE, "golden_rdm34.npz"), **out)
print("done")
