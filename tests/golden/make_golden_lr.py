"""Golden vectors for the RDM-only linear-response orbital blocks (reference density_matrix.py:233-563),
produced by RUNNING THE REFERENCE ITSELF on seeded random inputs (the functions are pure index arithmetic on
h, g, rdm1, rdm2).  Build container only:

    python tests/golden/make_golden_lr.py        ->  tests/golden/golden_lr.npz
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

from slowquant.unitary_coupled_cluster.density_matrix import (  # noqa: E402
    get_orbital_gradient_response,
    get_orbital_response_hessian_block,
    get_orbital_response_metric_sigma,
    get_orbital_response_property_gradient,
    get_orbital_response_vector_norm,
)

out = {}
for case, (nI, nA, nV, seed) in enumerate([(2, 3, 2, 11), (0, 4, 3, 12), (3, 2, 0, 13)]):
    rng = np.random.default_rng(seed)
    N = nI + nA + nV
    h = rng.normal(size=(N, N))
    g = 0.3 * rng.normal(size=(N, N, N, N))
    x = rng.normal(size=(N, N))
    rdm1 = rng.normal(size=(nA, nA))
    rdm2 = rng.normal(size=(nA, nA, nA, nA))
    # non-redundant rotations: inactive->active, inactive->virtual, active->virtual (p > q as in the reference)
    kappa = [(p, q) for p in range(N) for q in range(p) if not (p < nI) and not (q >= nI + nA) and not (nI <= q and p < nI + nA)]
    kappa = np.asarray(kappa, dtype=np.int64)
    kappa_dagger = kappa[:, ::-1].copy()
    K = len(kappa)
    n_exc = K + 3
    resp = rng.normal(size=(2 * n_exc, 4))
    pre = f"c{case}_"
    out[pre + "dims"] = np.array([nI, nA, nV, n_exc], dtype=np.int64)
    for name, arr in [("h", h), ("g", g), ("x", x), ("rdm1", rdm1), ("rdm2", rdm2), ("kappa", kappa), ("resp", resp)]:
        out[pre + name] = arr
    out[pre + "grad_response"] = get_orbital_gradient_response(h, g, kappa, nI, nA, rdm1, rdm2)
    out[pre + "metric_sigma"] = get_orbital_response_metric_sigma(kappa, nI, nA, rdm1)
    out[pre + "vector_norm"] = np.array([get_orbital_response_vector_norm(kappa, nI, nA, rdm1, resp, s, n_exc) for s in range(4)])
    out[pre + "property_gradient"] = np.array(
        [get_orbital_response_property_gradient(x, kappa, nI, nA, rdm1, resp, s, n_exc) for s in range(4)]
    )
    out[pre + "hessian_A"] = get_orbital_response_hessian_block(h, g, kappa_dagger, kappa, nI, nA, rdm1, rdm2)
    out[pre + "hessian_B"] = get_orbital_response_hessian_block(h, g, kappa_dagger, kappa_dagger, nI, nA, rdm1, rdm2)
np.savez_compressed(os.path.join(HERE, "golden_lr.npz"), **out)
print("wrote golden_lr.npz with", len(out), "arrays")
