"""CPU checks of the linear-response host side: the LR Hamiltonians / one-electron operators and folded products
against operator dictionaries exported from the reference (tests/golden/golden_ops.npz), the fold-aware sector
split, and the generalised eigenproblem / normalisation of the base class on the reference's own A, B, Sigma
(tests/golden/golden_config1.npz).  No kernels run here."""
import os

import numpy as np
import pytest

from conftest import ROOT
from slowquant_b200 import operators as ops
from slowquant_b200.fermionic_operator import FermionicOperator
from slowquant_b200.linear_response._symbolic import SectorSplit


@pytest.fixture(scope="module")
def gops():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_ops.npz"))


def _unpack(g, name):
    out = {}
    for row, c in zip(g[name + "_labels"], g[name + "_coeffs"]):
        out[tuple((int(x) // 2, bool(x % 2)) for x in row if x >= 0)] = float(c)
    return out


def _assert_same(op: FermionicOperator, ref: dict, tol=1e-13):
    mine = {k: v for k, v in op.operators.items()}
    assert set(mine) == set(ref), (len(mine), len(ref), sorted(set(mine) ^ set(ref))[:3])
    for k, v in ref.items():
        assert abs(mine[k] - v) <= tol, (k, mine[k], v)


def test_lr_hamiltonians_and_one_electron_operators(gops):
    nI, nA, nV = (int(x) for x in gops["dims"])
    h, g = gops["h"], gops["g"]
    N = nI + nA + nV
    _assert_same(ops.hamiltonian_1i_1a(h, g, nI, nA, nV), _unpack(gops, "H_1i_1a"))
    _assert_same(ops.hamiltonian_2i_2a(h, g, nI, nA, nV), _unpack(gops, "H_2i_2a"))
    _assert_same(ops.one_elec_op_full_space(h, N), _unpack(gops, "one_full"))
    _assert_same(ops.one_elec_op_1i_1a(h, nI, nA, nV), _unpack(gops, "one_1i_1a"))
    _assert_same(ops.one_elec_op_0i_0a(h, nI, nA), _unpack(gops, "one_0i_0a"))


def test_folded_products_of_the_orbital_blocks(gops):
    """H q, q^d H, H G^d q, q^d G^d H folded onto the active space (naive.py:125-184): plain product and the
    sector-split product must both equal the reference's folded dictionaries."""
    nI, nA, nV = (int(x) for x in gops["dims"])
    H1 = ops.hamiltonian_1i_1a(gops["h"], gops["g"], nI, nA, nV)
    split = SectorSplit(H1, nI, nA)
    q, q2, Gd = ops.G1_sa(1, 4), ops.G1_sa(0, 3), ops.G2_sa(2, 2, 3, 3, 1).dagger
    cases = {
        "Hq_folded": (H1 * q, split.times(q)),
        "qdH_folded": (q.dagger * H1, split.rtimes(q.dagger)),
        "HGdq_folded": (H1 * Gd * q2, split.times(Gd * q2)),
        "qdGdH_folded": (q2.dagger * Gd * H1, split.rtimes(q2.dagger * Gd)),
    }
    for name, (plain, fast) in cases.items():
        ref = _unpack(gops, name)
        for op in (plain, fast):
            folded = op.get_folded_operator(nI, nA, nV)
            folded = FermionicOperator({k: v for k, v in folded.operators.items() if abs(v) > 1e-15})
            _assert_same(folded, {k: v for k, v in ref.items() if abs(v) > 1e-15}, 1e-12)


def test_eigenproblem_norms_and_table_on_reference_matrices():
    """calc_excitation_energies / get_excited_state_norm (lr_baseclass.py:115-188) on the reference's A, B, Sigma for
    LiH tUPS(2,2) and H2O tUPS(4,4); LiH also against the literals of the reference's own test
    (tests/test_unitary_product_state.py:35-47)."""
    from slowquant_b200.linear_response.lr_baseclass import LinearResponseBaseClass

    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_config1.npz"))
    for name in ("lih", "h2o"):
        LR = object.__new__(LinearResponseBaseClass)
        LR.A, LR.B, LR.Sigma, LR.Delta = (g[f"{name}_{k}"] for k in ("A", "B", "Sigma", "Delta"))
        nG, nq = (int(x) for x in g[f"{name}_num_G_q"])
        LR.G_ops, LR.q_ops = [None] * nG, [None] * nq
        LR.calc_excitation_energies()
        assert np.max(np.abs(LR.excitation_energies - g[f"{name}_excitation_energies"])) < 1e-9
        assert np.max(np.abs(LR.get_excited_state_norm() - g[f"{name}_norms"])) < 1e-8
        with pytest.raises(ValueError):
            LR.get_formatted_oscillator_strength()
        LR.oscillator_strengths = g[f"{name}_oscillator_strengths"]
        assert LR.get_formatted_oscillator_strength().count("\n") == 1 + len(LR.excitation_energies)
        if name == "lih":
            literal = [0.129476, 0.178749, 0.178749, 0.604681, 0.646707, 0.740632, 0.740632, 1.002914,
                       2.074822, 2.137193, 2.137193, 2.455191, 2.954372]
            assert np.max(np.abs(LR.excitation_energies - np.array(literal))) < 1e-4
