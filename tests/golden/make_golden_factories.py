"""Golden operator dictionaries for the operator factories of the reference (operators.py: Epq, Eminuspq, epqrs, G1..G6, G1_sa,
G2_sa incl. every degenerate index combination i == j / a == b of all five cases, commutator, double_commutator, the 0i_0a /
1i_1a / 2i_2a / full-space Hamiltonians and one-electron operators on small random integrals), produced by RUNNING THE REFERENCE:

    python tests/golden/make_golden_factories.py        ->  tests/golden/golden_factories.json.gz
"""
from __future__ import annotations

import gzip
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from slowquant.unitary_coupled_cluster import operators as R  # noqa: E402


def dump(op):
    return [[[[int(i), bool(d)] for i, d in label], float(v)] for label, v in op.operators.items() if abs(v) > 1e-14]


cases = []
for i, j, a, b in itertools.product(range(2), range(2), range(2, 4), range(2, 4)):
    for case in (1, 2, 3, 4, 5):
        for ah in (True, False):
            call = ["G2_sa", [i, j, a, b, case, ah]]
            try:
                cases.append({"call": call, "op": dump(R.G2_sa(i, j, a, b, case, ah))})
            except Exception as e:  # noqa: BLE001
                cases.append({"call": call, "raises": type(e).__name__})
for f, args in [("G1", (1, 6)), ("G2", (0, 3, 5, 6)), ("G3", (0, 1, 2, 5, 6, 7)), ("G4", (0, 1, 2, 3, 4, 5, 6, 7)), ("G5", tuple(range(10))),
                ("G6", tuple(range(12))), ("G1_sa", (0, 3))]:
    for ah in (True, False):
        cases.append({"call": [f, list(args) + [ah]], "op": dump(getattr(R, f)(*args, ah))})
for p, q in itertools.product(range(3), repeat=2):
    cases.append({"call": ["Epq", [p, q]], "op": dump(R.Epq(p, q))})
    cases.append({"call": ["Eminuspq", [p, q]], "op": dump(R.Eminuspq(p, q))})
for p, q, r, s in [(0, 1, 2, 3), (0, 0, 1, 1), (1, 0, 0, 1), (2, 2, 2, 2), (0, 1, 1, 0)]:
    cases.append({"call": ["epqrs", [p, q, r, s]], "op": dump(R.epqrs(p, q, r, s))})
A, B, C = R.Epq(0, 2), R.G2_sa(0, 1, 2, 3, 2, True), R.G1(1, 4, True)
cases.append({"call": ["commutator", "Epq(0,2), G2_sa(0,1,2,3,2,True)"], "op": dump(R.commutator(A, B))})
cases.append({"call": ["double_commutator", "Epq(0,2), G2_sa(0,1,2,3,2,True), G1(1,4,True)"], "op": dump(R.double_commutator(A, B, C))})
rng = np.random.default_rng(3)
N = 4
h = rng.normal(size=(N, N))
g = rng.normal(size=(N, N, N, N))
for nI, nA in [(1, 2), (0, 3), (2, 2), (1, 1)]:
    nV = N - nI - nA
    cases.append({"call": ["hamiltonian_0i_0a", [nI, nA]], "op": dump(R.hamiltonian_0i_0a(h, g, nI, nA))})
    cases.append({"call": ["hamiltonian_1i_1a", [nI, nA, nV]], "op": dump(R.hamiltonian_1i_1a(h, g, nI, nA, nV))})
    cases.append({"call": ["hamiltonian_2i_2a", [nI, nA, nV]], "op": dump(R.hamiltonian_2i_2a(h, g, nI, nA, nV))})
    cases.append({"call": ["one_elec_op_0i_0a", [nI, nA]], "op": dump(R.one_elec_op_0i_0a(h, nI, nA))})
    cases.append({"call": ["one_elec_op_1i_1a", [nI, nA, nV]], "op": dump(R.one_elec_op_1i_1a(h, nI, nA, nV))})
cases.append({"call": ["hamiltonian_full_space", [N]], "op": dump(R.hamiltonian_full_space(h, g, N))})
cases.append({"call": ["one_elec_op_full_space", [N]], "op": dump(R.one_elec_op_full_space(h, N))})
with gzip.open(os.path.join(HERE, "golden_factories.json.gz"), "wt", compresslevel=9) as f:
    json.dump({"h": h.tolist(), "g": g.tolist(), "cases": cases}, f)
print("wrote golden_factories.json.gz:", len(cases), "cases,", sum(len(c.get("op", [])) for c in cases), "terms")
