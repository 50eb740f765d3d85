"""Golden vectors for extended CI spaces (reference ci_spaces.py:119-341 + the operator-state algebra on them), produced
by RUNNING THE REFERENCE in the build container:

    python tests/golden/make_golden_extended.py        ->  tests/golden/golden_extended.npz

  * idx2det of get_indexing_extended for several (nI, nA, nV, n_alpha, n_beta, order): bit-exact targets;
  * on (1, 2, 2; 1, 1) order 1 and (1, 3, 1; 2, 1) order 2: a seeded state, construct_ups_state (tUPS, with
    space_extension_offset) and its adjoint, propagate_unitary, get_grad_action, generic operators with do_unsafe=True
    (an orbital rotation that leaves the space), expectation values.
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

from slowquant.unitary_coupled_cluster import operators as rops  # noqa: E402
from slowquant.unitary_coupled_cluster.ci_spaces import get_indexing_extended  # noqa: E402
from slowquant.unitary_coupled_cluster.operator_state_algebra import (  # noqa: E402
    construct_ups_state,
    expectation_value,
    get_grad_action,
    propagate_state,
    propagate_unitary,
)
from slowquant.unitary_coupled_cluster.util import UpsStructure  # noqa: E402

out = {}
SPACES = [(1, 2, 1, 1, 1, 1), (1, 2, 2, 1, 1, 1), (2, 2, 1, 1, 1, 2), (1, 3, 1, 2, 1, 2), (0, 3, 2, 1, 2, 2), (2, 2, 0, 1, 1, 2), (1, 2, 2, 1, 1, 2)]
out["spaces"] = np.array(SPACES, dtype=np.int64)
for k, sp in enumerate(SPACES):
    out[f"idx2det_{k}"] = np.array(get_indexing_extended(*sp).idx2det, dtype=np.int64)

for tag, sp, L in (("a", (1, 2, 2, 1, 1, 1), 2), ("b", (1, 3, 1, 2, 1, 2), 1)):
    nI, nA, nV, na, nb, order = sp
    ci = get_indexing_extended(*sp)
    lay = UpsStructure()
    lay.create_tiled(nA, {"n_layers": L, "do_tups": True})
    rng = np.random.default_rng(40 + order)
    th = rng.uniform(-np.pi, np.pi, lay.n_params)
    state = rng.normal(size=len(ci.idx2det))
    state /= np.linalg.norm(state)
    pre = tag + "_"
    out[pre + "space"] = np.array(sp, dtype=np.int64)
    out[pre + "n_layers"] = np.array(L)
    out[pre + "thetas"] = th
    out[pre + "state"] = state
    U = construct_ups_state(state, ci, th.tolist(), lay)
    out[pre + "U_state"] = U
    out[pre + "Ud_state"] = construct_ups_state(state, ci, th.tolist(), lay, dagger=True)
    out[pre + "unitary3"] = propagate_unitary(state, 3, ci, th.tolist(), lay)
    out[pre + "grad0"] = get_grad_action(state, 0, ci, lay)
    out[pre + "grad1"] = get_grad_action(state, 1, ci, lay)
    N = nI + nA + nV
    q = rops.G1_sa(0, N - 1)                       # inactive -> last orbital: leaves the space for most determinants
    E = rops.Epq(nI, nI + 1) * rops.Epq(nI + 1, nI) + 0.5 * rops.Epq(nI, nI)   # stays inside every sector
    out[pre + "q_state"] = propagate_state([q], state, ci, do_unsafe=True)
    out[pre + "qd_q_state"] = propagate_state([q.dagger, q], state, ci, do_unsafe=True)
    out[pre + "E_state"] = propagate_state([E], state, ci)
    out[pre + "U_q_state"] = propagate_state(["U", q], state, ci, th.tolist(), lay, do_unsafe=True)
    out[pre + "expval"] = np.array(expectation_value(state, ["Ud", E, "U"], state, ci, th.tolist(), lay))
    try:
        propagate_state([q], state, ci)
        out[pre + "q_raises"] = np.array(0)
    except Exception as exc:  # numba KeyError
        out[pre + "q_raises"] = np.array(1)
        print("safe mode raised", type(exc).__name__)
np.savez_compressed(os.path.join(HERE, "golden_extended.npz"), **out)
print({k: v.shape for k, v in out.items() if k.startswith("idx2det")})
