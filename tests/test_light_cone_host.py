"""Host-side check of the light-cone split of a circuit (operator_state_algebra._light_cone_plan): for a reference determinant the
head of the circuit is run in the orbital window it can reach, embedded into the full space, and the tail is run there -- with the
oracle doing the arithmetic on both routes, the vectors must agree.  (The GPU test runs the same split through the kernels.)"""
import numpy as np
import pytest

from oracle import sq_oracle as orc


def _tups(n, L, qnp=False):
    types, idx = [], []
    for _ in range(L):
        for start in (0, 1):
            for p in range(start, n - 1, 2):
                if not qnp:
                    types.append("sa_single"); idx.append((p, p + 1))
                types.append("double"); idx.append((2 * p, 2 * p + 1, 2 * p + 2, 2 * p + 3))
                types.append("sa_single"); idx.append((p, p + 1))
    return types, idx


@pytest.mark.parametrize("n,ne,L,qnp,frac", [(6, 3, 2, False, 0.3), (8, 4, 3, False, 0.2), (8, 4, 3, True, 0.2), (7, 3, 2, False, 0.5),
                                             (8, 3, 2, False, 0.3)])
def test_light_cone_split_equals_full_circuit(n, ne, L, qnp, frac):
    from slowquant_b200 import operator_state_algebra as osa
    from slowquant_b200.ci_spaces import get_indexing
    from slowquant_b200.util import UpsStructure

    types, idx = _tups(n, L, qnp)
    lay = UpsStructure()
    for t, i in zip(types, idx):
        lay._push(t, i, None)
    th = np.random.default_rng(n * 10 + L).uniform(-np.pi, np.pi, len(types))
    info = get_indexing(0, n, 0, ne, ne, device=-1)
    a = int(info.strings(0)[0])
    plan = osa._light_cone_plan(info, lay, a, a, max_fraction=frac)
    assert plan is not None
    lo, hi = plan["window"]
    assert 0 < hi - lo < n and plan["k0"] > len(plan["head"]) > 1          # identities were dropped, the window is a proper sub-space
    # full route
    sp = orc.get_indexing(0, n, 0, ne, ne)
    hf = np.zeros(sp.num_det)
    hf[0] = 1.0
    want = orc.construct_ups_state(hf, sp, th, types, idx)
    # split route: head in the window space, embedding, tail in the full space
    sub = orc.get_indexing(0, hi - lo, 0, ne - lo, ne - lo)
    s0 = np.zeros(sub.num_det)
    s0[plan["ref"]] = 1.0
    st = plan["struct"]
    head = orc.construct_ups_state(s0, sub, th[plan["head"]], st.excitation_operator_type, st.excitation_indices)
    full = np.zeros(sp.num_det)
    full[plan["embed"]] = head
    got = orc.construct_ups_state(full, sp, th[plan["k0"]:], types[plan["k0"]:], idx[plan["k0"]:])
    assert np.max(np.abs(got - want)) < 1e-13
    assert abs(np.linalg.norm(got) - 1.0) < 1e-12


def test_no_light_cone_when_the_first_operators_span_the_space():
    from slowquant_b200 import operator_state_algebra as osa
    from slowquant_b200.ci_spaces import get_indexing
    from slowquant_b200.util import UpsStructure

    n, ne = 6, 3
    lay = UpsStructure()
    lay._push("sa_single", (0, 5), None)          # couples the lowest occupied with the highest empty orbital at once
    lay._push("sa_single", (2, 3), None)
    info = get_indexing(0, n, 0, ne, ne, device=-1)
    a = int(info.strings(0)[0])
    assert osa._light_cone_plan(info, lay, a, a, max_fraction=0.5) is None
