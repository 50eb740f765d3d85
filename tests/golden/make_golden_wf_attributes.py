"""Integer bookkeeping of the reference's WaveFunctionUPS / WaveFunctionUCC constructors (ups_wavefunction.py:39-313,
ucc_wavefunction.py:36-200: orbital partitions, shifted index lists, kappa index lists incl. redundant / active-active ones,
electron counts), produced by RUNNING THE REFERENCE on H2O / STO-3G for several active spaces:

    python tests/golden/make_golden_wf_attributes.py        ->  tests/golden/golden_wf_attributes.json

Every attribute of the reference object that is an int, a bool, or a (nested) list / tuple / array of ints is stored; the test
requires the engine's object to carry the same attribute with the same value.
"""
from __future__ import annotations

import json
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
from slowquant.unitary_coupled_cluster.ucc_wavefunction import WaveFunctionUCC  # noqa: E402
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


def intlike(v):
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, np.ndarray):
        if v.dtype.kind in "iu" and v.size < 5000:
            return v.tolist()
        return None
    if isinstance(v, (list, tuple)):
        if len(v) > 5000:
            return None
        out = [intlike(x) for x in v]
        return None if any(x is None for x in out) else out
    return None


cases = []
SPECS = [
    ("ups", (2, 2), "tUPS", {"n_layers": 1}, False),
    ("ups", (2, 2), "tUPS", {"n_layers": 1}, True),
    ("ups", (4, 4), "tUPS", {"n_layers": 2}, True),
    ("ups", (4, 4), "tUPS", {"n_layers": 1, "do_pp": True}, False),
    ("ups", (4, 3), "fUCCSD", {}, False),
    ("ups", (6, 5), "fUCCSD", {}, True),
    ("ups", (2, 3), "QNP", {"n_layers": 2}, True),
    ("ups", (6, 4), "kSAfUpCCGSD", {"n_layers": 1}, False),
    ("ucc", (2, 2), "SD", None, False),
    ("ucc", (4, 4), "SD", None, True),
    ("ucc", (4, 3), "SDT", None, False),
]
for kind, cas, ansatz, options, iak in SPECS:
    if kind == "ups":
        WF = WaveFunctionUPS(cas, c_mo, SQobj, ansatz, ansatz_options=dict(options), include_active_kappa=iak)
    else:
        WF = WaveFunctionUCC(cas, c_mo, SQobj, ansatz, include_active_kappa=iak)
    attrs = {}
    for k, v in WF.__dict__.items():
        if k.startswith("_"):
            continue
        iv = intlike(v)
        if iv is not None:
            attrs[k] = iv
    cases.append({"kind": kind, "cas": list(cas), "ansatz": ansatz, "options": options, "include_active_kappa": iak, "attributes": attrs})
    print(kind, cas, ansatz, sorted(attrs)[:6], len(attrs))

# state-averaged class on the same integrals: (2,2) with the singlet subspace of the reference's test_saups_h2_3states, and (4,3)
from slowquant.unitary_coupled_cluster.sa_ups_wavefunction import WaveFunctionSAUPS  # noqa: E402

SA_SPECS = [
    ((2, 2), ([[1.0], [2 ** (-1 / 2), -(2 ** (-1 / 2))], [1.0]], [["1100"], ["1001", "0110"], ["0011"]]), "tUPS", {"n_layers": 1}, True),
    ((4, 3), ([[1.0], [2 ** (-1 / 2), -(2 ** (-1 / 2))]], [["111100"], ["111001", "110110"]]), "SAfUCCSD", {}, False),
]
for cas, states, ansatz, options, iak in SA_SPECS:
    WF = WaveFunctionSAUPS(cas, c_mo, SQobj, states, ansatz, ansatz_options=dict(options), include_active_kappa=iak)
    attrs = {}
    for k, v in WF.__dict__.items():
        if k.startswith("_"):
            continue
        iv = intlike(v)
        if iv is not None:
            attrs[k] = iv
    cases.append({"kind": "saups", "cas": list(cas), "ansatz": ansatz, "options": options, "include_active_kappa": iak,
                  "states": [states[0], states[1]], "attributes": attrs})
    print("saups", cas, ansatz, len(attrs))

ig = SQobj.integral
json.dump(
    {
        "c_mo": np.asarray(c_mo).tolist(),
        "h_ao": (np.asarray(ig.kinetic_energy_matrix) + np.asarray(ig.nuclear_attraction_matrix)).tolist(),
        "eri_ao": np.asarray(ig.electron_repulsion_tensor).tolist(),
        "num_elec": int(SQobj.molecule.number_electrons),
        "cases": cases,
    },
    open(os.path.join(HERE, "golden_wf_attributes.json"), "w"),
)
print("wrote golden_wf_attributes.json:", len(cases), "cases")
