"""Extended CI spaces (reference ci_spaces.py:119-259: CAS + singles/doubles into inactive and virtual orbitals) on the
CUDA path, against vectors produced by the reference on the same seeded inputs (make_golden_extended.py): the ansatz
unitary with space_extension_offset, single unitaries, generators, generic operators that leave the space with
do_unsafe=True (silently skipped, osa.py:131-135), KeyError without it, "U"/"Ud" operator lists, build_operator_matrix."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ge():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_extended.npz"))


def _d(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_extended_space_algebra(ge, tag):
    import slowquant_b200.operator_state_algebra as osa
    from slowquant_b200 import operators as mops
    from slowquant_b200.ci_spaces import get_indexing_extended
    from slowquant_b200.util import UpsStructure

    pre = tag + "_"
    sp = tuple(int(x) for x in ge[pre + "space"])
    nI, nA, nV, _, _, _ = sp
    ci = get_indexing_extended(*sp)
    assert np.array_equal(ci.idx2det, ge[f"idx2det_{1 if tag == 'a' else 3}"])
    lay = UpsStructure()
    lay.create_tiled(nA, {"n_layers": int(ge[pre + "n_layers"]), "do_tups": True})
    th = ge[pre + "thetas"].tolist()
    state = ge[pre + "state"]
    keep = state.copy()
    tol = 1e-12
    U = osa.construct_ups_state(state, ci, th, lay)
    assert _d(U, ge[pre + "U_state"]) < tol
    assert _d(osa.construct_ups_state(state, ci, th, lay, dagger=True), ge[pre + "Ud_state"]) < tol
    assert _d(osa.construct_ups_state(U, ci, th, lay, dagger=True), state) < tol
    assert _d(osa.propagate_unitary(state, 3, ci, th, lay), ge[pre + "unitary3"]) < tol
    assert _d(osa.get_grad_action(state, 0, ci, lay), ge[pre + "grad0"]) < tol
    assert _d(osa.get_grad_action(state, 1, ci, lay), ge[pre + "grad1"]) < tol
    N = nI + nA + nV
    q = mops.G1_sa(0, N - 1)
    E = mops.Epq(nI, nI + 1) * mops.Epq(nI + 1, nI) + 0.5 * mops.Epq(nI, nI)
    assert _d(osa.propagate_state([q], state, ci, do_unsafe=True), ge[pre + "q_state"]) < tol
    assert _d(osa.propagate_state([q.dagger, q], state, ci, do_unsafe=True), ge[pre + "qd_q_state"]) < tol
    assert _d(osa.propagate_state([E], state, ci), ge[pre + "E_state"]) < tol
    assert _d(osa.propagate_state(["U", q], state, ci, th, lay, do_unsafe=True), ge[pre + "U_q_state"]) < tol
    assert abs(osa.expectation_value(state, ["Ud", E, "U"], state, ci, th, lay) - float(ge[pre + "expval"])) < tol
    if int(ge[pre + "q_raises"]):
        with pytest.raises(KeyError):
            osa.propagate_state([q], state, ci)
    else:
        osa.propagate_state([q], state, ci)
    assert np.array_equal(state, keep), "inputs must not be modified"
    # device-resident vectors and the dense matrix view
    out_dev = osa.propagate_state([E], torch.from_numpy(state).cuda(), ci)
    assert isinstance(out_dev, torch.Tensor) and _d(out_dev.cpu().numpy(), ge[pre + "E_state"]) < tol
    if tag == "a":
        M = osa.build_operator_matrix(E, ci)
        assert _d(M @ state, ge[pre + "E_state"]) < tol
        Mq = osa.build_operator_matrix(q, ci, do_unsafe=True)
        assert _d(Mq @ state, ge[pre + "q_state"]) < tol
