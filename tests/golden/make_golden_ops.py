"""Golden operator dictionaries for the linear-response Hamiltonians and one-electron operators
(reference operators.py:532-737), produced by RUNNING THE REFERENCE in the build container:

    python tests/golden/make_golden_ops.py        ->  tests/golden/golden_ops.npz

Each operator is stored as labels[n_strings, 8] (entry = 2*spin_orbital + dagger, -1 padded) + coeffs[n_strings].
Also the folded products the naive LR applies to the state (naive.py:124-194).
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

import slowquant.unitary_coupled_cluster.operators as rops  # noqa: E402


def pack(op, width=12):
    items = sorted(op.operators.items())
    labels = -np.ones((len(items), width), dtype=np.int64)
    coeffs = np.zeros(len(items))
    for k, (label, c) in enumerate(items):
        for m, (idx, dag) in enumerate(label):
            labels[k, m] = 2 * idx + (1 if dag else 0)
        coeffs[k] = c
    return labels, coeffs


out = {}
nI, nA, nV = 2, 2, 2
N = nI + nA + nV
rng = np.random.default_rng(77)
h = rng.normal(size=(N, N))
g = rng.normal(size=(N, N, N, N))
h[0, 3] = 0.0  # exercise the 1e-14 screening
g[1, 2, 3, 4] = 0.0
out["dims"] = np.array([nI, nA, nV], dtype=np.int64)
out["h"] = h
out["g"] = g
H1 = rops.hamiltonian_1i_1a(h, g, nI, nA, nV)
ops = {
    "H_1i_1a": H1,
    "H_2i_2a": rops.hamiltonian_2i_2a(h, g, nI, nA, nV),
    "one_full": rops.one_elec_op_full_space(h, N),
    "one_1i_1a": rops.one_elec_op_1i_1a(h, nI, nA, nV),
    "one_0i_0a": rops.one_elec_op_0i_0a(h, nI, nA),
    # folded products of the q-G blocks: q = G1_sa(1, 4) (inactive -> virtual), G = G2_sa(2,2,3,3,1)
    "Hq_folded": (H1 * rops.G1_sa(1, 4)).get_folded_operator(nI, nA, nV),
    "qdH_folded": (rops.G1_sa(1, 4).dagger * H1).get_folded_operator(nI, nA, nV),
    "HGdq_folded": (H1 * rops.G2_sa(2, 2, 3, 3, 1).dagger * rops.G1_sa(0, 3)).get_folded_operator(nI, nA, nV),
    "qdGdH_folded": (rops.G1_sa(0, 3).dagger * rops.G2_sa(2, 2, 3, 3, 1).dagger * H1).get_folded_operator(nI, nA, nV),
}
for name, op in ops.items():
    out[name + "_labels"], out[name + "_coeffs"] = pack(op)
np.savez_compressed(os.path.join(HERE, "golden_ops.npz"), **out)
print({k: len(v.operators) for k, v in ops.items()})
