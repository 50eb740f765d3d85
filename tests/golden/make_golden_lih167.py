"""AO integrals of the LiH/STO-3G geometry of the reference's all-ST test (tests/test_unitary_product_state.py:64-80),
from the reference's own integral engine and RHF, so that the test can be re-run end to end on the GPU box:

    python tests/golden/make_golden_lih167.py        ->  tests/golden/golden_lih167.npz
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

SQobj = sq.SlowQuant()
SQobj.set_molecule(
    """Li  0.0  0.0  0.0;
        H 1.67 0.0 0.0;""",
    distance_unit="angstrom",
)
SQobj.set_basis_set("sto-3g")
SQobj.init_hartree_fock()
SQobj.hartree_fock.run_restricted_hartree_fock()
np.savez_compressed(
    os.path.join(HERE, "golden_lih167.npz"),
    h_ao=np.array(SQobj.integral.kinetic_energy_matrix + SQobj.integral.nuclear_attraction_matrix),
    eri_ao=np.array(SQobj.integral.electron_repulsion_tensor),
    dipole_ao=np.array([SQobj.integral.get_multipole_matrix(np.array(v)) for v in ([1, 0, 0], [0, 1, 0], [0, 0, 1])]),
    c_mo_rhf=np.array(SQobj.hartree_fock.mo_coeff),
    nuclear_repulsion=np.array(SQobj.molecule.nuclear_repulsion),
)
print("ok")
