"""Golden vectors for the reference's per-string kernels (operator_state_algebra.py:33-410: bitcount,
apply_operator_serial / _threaded, apply_operator_SA_serial / _threaded, add_operator_matrix) and for
generate_spin_strings (ci_spaces.py:56-73), produced by RUNNING THE REFERENCE in the build container:

    python tests/golden/make_golden_strings.py        ->  tests/golden/golden_strings.npz

Space CAS(4,5) with 3 alpha / 1 beta electrons (unequal spin counts on purpose).  Every case stores the index arrays
exactly as the reference's propagate_state builds them (osa.py:560-612), the factor, and the kernel's output for a
seeded state accumulated onto a seeded, non-zero tmp_state.  Also idx2det of degenerate spaces (edge_*).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

from slowquant.unitary_coupled_cluster import operator_state_algebra as rosa  # noqa: E402
from slowquant.unitary_coupled_cluster import operators as rops  # noqa: E402
from slowquant.unitary_coupled_cluster.ci_spaces import generate_spin_strings, get_indexing  # noqa: E402

nA, na, nb = 5, 3, 1
ci = get_indexing(0, nA, 0, na, nb)
N = len(ci.idx2det)
rng = np.random.default_rng(77)
state = rng.normal(size=N)
states = rng.normal(size=(3, N))
tmp0 = rng.normal(size=N)
tmps0 = rng.normal(size=(3, N))
parity_check = np.zeros(2 * nA + 1, dtype=np.int64)
num = 0
for i in range(2 * nA - 1, -1, -1):
    num += 2**i
    parity_check[2 * nA - i] = num

out = {"space": np.array([0, nA, 0, na, nb]), "idx2det": np.array(ci.idx2det, dtype=np.int64), "state": state, "states": states,
       "tmp0": tmp0, "tmps0": tmps0, "parity_check": parity_check}
out["bitcount_in"] = np.array([0, 1, 2, 3, 255, 1023, 2**40 + 5, 2**62 - 1], dtype=np.int64)
out["bitcount_out"] = np.array([rosa.bitcount(int(x)) for x in out["bitcount_in"]], dtype=np.int64)
for n, k in ((5, 3), (5, 1), (4, 0), (3, 3)):
    out[f"strings_{n}_{k}"] = np.array(list(generate_spin_strings(n, k)), dtype=np.int64).reshape(-1, n)

# a mix of strings: one- and two-body, number operators (creator == annihilator), alpha/beta mixed, a triple
ops = rops.Epq(1, 3) * 0.7 + rops.Epq(2, 2) * (-1.3) + rops.epqrs(0, 2, 3, 1) * 0.45 + rops.epqrs(4, 4, 1, 0) * 1.1
ops = ops + rops.G2(0, 3, 5, 8, True) * 0.9 + rops.G3(0, 2, 1, 6, 4, 7, False) * (-0.6)
labels = list(ops.operators.keys())
out["n_cases"] = np.array(len(labels))
for c, label in enumerate(labels):
    factor = ops.operators[label]
    anni = [o[0] for o in label if not o[1]]
    create = [o[0] for o in label if o[1]]
    create_screen = [i for i in create if i not in anni]
    anni_screen = [i for i in anni if i not in create]
    a_ser = np.array(anni + create, dtype=np.int64)
    a_thr = np.array(create + anni, dtype=np.int64)
    i64 = lambda x: np.array(x, dtype=np.int64)  # noqa: E731
    pre = f"c{c}_"
    out[pre + "factor"] = np.array(factor)
    out[pre + "a_serial"], out[pre + "create_screen"], out[pre + "anni_idx"] = a_ser, i64(create_screen), i64(anni)
    out[pre + "a_threaded"], out[pre + "create_idx"], out[pre + "anni_screen"] = a_thr, i64(create), i64(anni_screen)
    out[pre + "serial"] = rosa.apply_operator_serial(
        state, a_ser, i64(create_screen), i64(anni), nA, parity_check, ci.idx2det, ci.det2idx, False, tmp0.copy(), factor)
    out[pre + "threaded"] = rosa.apply_operator_threaded(
        state, a_thr, i64(create), i64(anni_screen), nA, parity_check, ci.idx2det, ci.det2idx, False, tmp0.copy(), factor)
    out[pre + "sa_serial"] = rosa.apply_operator_SA_serial(
        states, a_ser, i64(create_screen), i64(anni), nA, parity_check, ci.idx2det, ci.det2idx, False, tmps0.copy(), factor)
    out[pre + "sa_threaded"] = rosa.apply_operator_SA_threaded(
        states, a_thr, i64(create), i64(anni_screen), nA, parity_check, ci.idx2det, ci.det2idx, False, tmps0.copy(), factor)
    out[pre + "matrix"] = rosa.add_operator_matrix(
        np.zeros((N, N)), a_ser, i64(create_screen), i64(anni), nA, parity_check, ci.idx2det, ci.det2idx, False, factor)
    assert np.array_equal(out[pre + "serial"], out[pre + "threaded"])

# degenerate spaces (no electrons of one spin, a full spin string, one orbital, empty space): idx2det bit-exact targets
EDGE = [(3, 0, 2), (3, 2, 0), (3, 3, 3), (1, 1, 0), (1, 0, 0), (4, 4, 1), (6, 0, 0), (5, 0, 5), (2, 1, 1)]
out["edge_spaces"] = np.array(EDGE, dtype=np.int64)
for n, a, b in EDGE:
    out[f"edge_idx2det_{n}_{a}_{b}"] = np.array(get_indexing(0, n, 0, a, b).idx2det, dtype=np.int64)

np.savez_compressed(os.path.join(HERE, "golden_strings.npz"), **out)
print("wrote golden_strings.npz:", len(labels), "strings,", N, "determinants")

# get_determinant_expansion_from_operator_on_HF (osa.py:2979-3033): reference outputs for a set of operators on two spaces
import json  # noqa: E402

hf_cases = []
for nA_, na_, nb_ in ((4, 2, 2), (5, 3, 1)):
    ops_ = {
        "G1": rops.G1(0, 2 * nA_ - 2, True),
        "G2": rops.G2(0, 1, 2 * nA_ - 2, 2 * nA_ - 1, True),
        "G1_sa": rops.G1_sa(0, nA_ - 1, True),
        "G2_sa_1": rops.G2_sa(0, 1, nA_ - 2, nA_ - 1, 1, True),
        "G2_sa_4": rops.G2_sa(0, 1, nA_ - 2, nA_ - 1, 4, False),
        "number": rops.Epq(0, 0) * 1.5 + rops.Epq(nA_ - 1, nA_ - 1) * 0.5 + rops.epqrs(0, 0, 1, 1) * 0.25,
        "mixed": rops.Epq(nA_ - 1, 0) * rops.Epq(nA_ - 2, 1) * 0.7 + rops.Epq(1, 0) * 0.3,
    }
    for name, op in ops_.items():
        c, d = rosa.get_determinant_expansion_from_operator_on_HF(op, nA_, na_, nb_)
        hf_cases.append({"space": [nA_, na_, nb_], "name": name, "operator": [[[list(x) for x in k], v] for k, v in op.operators.items()],
                         "coeffs": [float(x) for x in c], "dets": list(d)})
json.dump(hf_cases, open(os.path.join(HERE, "golden_hf_expansion.json"), "w"))
print("wrote golden_hf_expansion.json:", len(hf_cases), "cases")

# FermionicOperator.get_qiskit_form (fermionic_operator.py:357-377): reference key strings and coefficients
qk = []
for num_orbs_, op in ((4, rops.G2_sa(0, 1, 2, 3, 4, True)), (3, rops.Epq(0, 2) * rops.Epq(1, 1)), (5, rops.G3(0, 2, 1, 6, 4, 7, False) * 0.5 + rops.G1(3, 8, True))):
    qk.append({"num_orbs": num_orbs_, "operator": [[[list(x) for x in k], v] for k, v in op.operators.items()],
               "qiskit_form": [[k, float(v)] for k, v in op.get_qiskit_form(num_orbs_).items()]})
json.dump(qk, open(os.path.join(HERE, "golden_qiskit_form.json"), "w"))
print("wrote golden_qiskit_form.json:", len(qk), "operators")
