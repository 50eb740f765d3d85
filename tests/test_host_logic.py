"""CPU checks of the product's host side: symbolic algebra, ansatz layouts, integer tables of the C ABI
(host-only space, device = -1) and the closed-form string action, against the reference goldens and the
oracle.  No kernels run here."""
import ctypes as C
import itertools
import re

import numpy as np
import pytest

from conftest import ROOT, assert_opdict_close, json_to_opdict
from oracle import sq_oracle as orc
from slowquant_b200 import _lib
from slowquant_b200 import operators as ops
from slowquant_b200.fermionic_operator import FermionicOperator
from slowquant_b200.util import UpsStructure


def test_abi_exports_every_declared_symbol():
    header = open(f"{ROOT}/include/sqsv.h").read()
    declared = set(re.findall(r"\b(sq_[a-z0-9_]+)\s*\(", header))
    lib = _lib.load()
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"libsqsv.so does not export {name}"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.sq_version() >= 100


def _host_space(n, na, nb):
    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.sq_space_create(n, na, nb, -1, 0, -1, C.byref(h)))
    return lib, h


def test_idx2det_and_det2idx_bit_exact(golden):
    arrays, _, _ = golden
    for key in arrays.files:
        if not key.startswith("idx2det_"):
            continue
        n, na, nb = (int(x) for x in key.split("_")[1:])
        lib, h = _host_space(n, na, nb)
        ref = arrays[key]
        nd = lib.sq_space_num_det(h)
        assert nd == len(ref)
        out = np.empty(nd, dtype=np.int64)
        _lib.check(lib.sq_space_export_idx2det(h, 0, nd, out.ctypes.data_as(C.POINTER(C.c_int64))))
        assert np.array_equal(out, ref), key
        # det2idx inverts idx2det; determinants outside the space map to -1 (KeyError in the reference)
        idx = np.empty(nd, dtype=np.int64)
        _lib.check(lib.sq_space_det2idx(h, nd, ref.ctypes.data_as(C.POINTER(C.c_int64)), idx.ctypes.data_as(C.POINTER(C.c_int64))))
        assert np.array_equal(idx, np.arange(nd))
        bad = np.array([0, (1 << (2 * n)) - 1, -5, 1 << (2 * n)], dtype=np.int64)
        res = np.empty(4, dtype=np.int64)
        _lib.check(lib.sq_space_det2idx(h, 4, bad.ctypes.data_as(C.POINTER(C.c_int64)), res.ctypes.data_as(C.POINTER(C.c_int64))))
        assert res[2] == -1 and res[3] == -1
        if 0 < na + nb < 2 * n:
            assert res[0] == -1 and res[1] == -1
        lib.sq_space_destroy(h)


def test_degenerate_spaces_bit_exact():
    """Spaces with no electrons of one spin, a full spin string, a single orbital or no electrons at all (one string per
    empty / full spin, ci_spaces.py:56-73 yields exactly one list for k = 0) against the reference's idx2det."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_strings.npz"))
    for n, na, nb in g["edge_spaces"]:
        n, na, nb = int(n), int(na), int(nb)
        lib, h = _host_space(n, na, nb)
        ref = g[f"edge_idx2det_{n}_{na}_{nb}"]
        nd = lib.sq_space_num_det(h)
        assert nd == len(ref), (n, na, nb)
        out = np.empty(nd, dtype=np.int64)
        _lib.check(lib.sq_space_export_idx2det(h, 0, nd, out.ctypes.data_as(C.POINTER(C.c_int64))))
        assert np.array_equal(out, ref), (n, na, nb)
        idx = np.empty(nd, dtype=np.int64)
        _lib.check(lib.sq_space_det2idx(h, nd, ref.ctypes.data_as(C.POINTER(C.c_int64)), idx.ctypes.data_as(C.POINTER(C.c_int64))))
        assert np.array_equal(idx, np.arange(nd))
        lib.sq_space_destroy(h)


def test_ci_info_constructor_takes_the_reference_tables(golden):
    """CI_Info(nI, nA, nV, n_alpha, n_beta, idx2det, det2idx) -- the reference's positional call (ci_spaces.py:12-21): the tables are
    derived by the engine, a supplied idx2det must be this product space's list in get_indexing order."""
    from slowquant_b200.ci_spaces import CI_Info

    arrays, _, _ = golden
    ref = arrays["idx2det_5_2_3"]
    info = CI_Info(0, 5, 0, 2, 3, ref, {int(d): i for i, d in enumerate(ref)}, device=-1)
    assert info.num_det == len(ref) and info.det2idx[int(ref[7])] == 7 and int(ref[3]) in info.det2idx and 0 not in info.det2idx
    with pytest.raises(ValueError):
        CI_Info(0, 5, 0, 2, 3, ref[::-1].copy(), None, device=-1)
    with pytest.raises(ValueError):
        CI_Info(0, 5, 0, 2, 3, ref[:-1].copy(), None, device=-1)


def test_space_argument_errors():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.sq_space_create(0, 0, 0, -1, 0, -1, C.byref(h)) == _lib.SQ_ERR_INVALID
    assert lib.sq_space_create(4, 5, 1, -1, 0, -1, C.byref(h)) == _lib.SQ_ERR_INVALID
    assert lib.sq_space_create(4, 2, 2, -1, 3, 2, C.byref(h)) == _lib.SQ_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(lib.sq_space_create(4, 2, 2, -1, 0, 99, C.byref(h)))


def test_fermionic_operator_algebra(golden):
    _, _, gops = golden
    E = ops.Epq
    built = {
        "Epq_2_0": E(2, 0),
        "Epq_1_1": E(1, 1),
        "Epq_3_1*Epq_1_2": E(3, 1) * E(1, 2),
        "Epq_0_1*Epq_1_0": E(0, 1) * E(1, 0),
        "epqrs_0_1_1_2": ops.epqrs(0, 1, 1, 2),
        "epqrs_2_2_2_2": ops.epqrs(2, 2, 2, 2),
        "G1_1_4_AH": ops.G1(1, 4, True),
        "G2_0_3_4_7_AH": ops.G2(0, 3, 4, 7, True),
        "G2_0_1_6_7_AH": ops.G2(0, 1, 6, 7, True),
        "G2_2_3_4_5_H": ops.G2(2, 3, 4, 5, False),
        "G3_0_1_2_5_6_7_AH": ops.G3(0, 1, 2, 5, 6, 7, True),
        "G4_0_1_2_3_4_5_6_7_AH": ops.G4(0, 1, 2, 3, 4, 5, 6, 7, True),
        "G1_sa_0_2_AH": ops.G1_sa(0, 2, True),
        "G2_sa_0_0_2_2_c1_AH": ops.G2_sa(0, 0, 2, 2, 1, True),
        "G2_sa_0_0_2_3_c2_AH": ops.G2_sa(0, 0, 2, 3, 2, True),
        "G2_sa_0_1_2_2_c3_AH": ops.G2_sa(0, 1, 2, 2, 3, True),
        "G2_sa_0_1_2_3_c4_AH": ops.G2_sa(0, 1, 2, 3, 4, True),
        "G2_sa_0_1_2_3_c5_AH": ops.G2_sa(0, 1, 2, 3, 5, True),
        "commutator_E01_E12": ops.commutator(E(0, 1), E(1, 2)),
    }
    for name, op in built.items():
        assert_opdict_close(op.operators, json_to_opdict(gops[name]))


def test_algebra_agrees_with_oracle_on_random_products():
    rng = np.random.default_rng(3)
    for _ in range(40):
        k = int(rng.integers(2, 7))
        label = tuple((int(rng.integers(0, 6)), bool(rng.integers(0, 2))) for _ in range(k))
        a = FermionicOperator({(): 1.0}) * FermionicOperator({label: 1.5})
        assert_opdict_close(a.operators, orc.normal_order(label, 1.5))
        assert_opdict_close(a.dagger.operators, orc.op_dagger(orc.normal_order(label, 1.5)))


def test_operator_factories_match_reference():
    """221 operators built by the reference's factories (tests/golden/make_golden_factories.py): G2_sa for every index combination
    incl. i == j / a == b of all five cases (anti-Hermitian and plain), G1..G6, G1_sa, Epq, Eminuspq, epqrs, commutator,
    double_commutator, the 0i_0a / 1i_1a / 2i_2a / full-space Hamiltonians and one-electron operators on random integrals --
    identical label sets and coefficients."""
    import gzip
    import json
    import os

    with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_factories.json.gz"), "rt") as f:
        d = json.load(f)
    h, g = np.array(d["h"]), np.array(d["g"])
    A, B, Cc = ops.Epq(0, 2), ops.G2_sa(0, 1, 2, 3, 2, True), ops.G1(1, 4, True)
    assert len(d["cases"]) == 221
    for case in d["cases"]:
        name, args = case["call"]
        try:
            if name == "commutator":
                got = ops.commutator(A, B)
            elif name == "double_commutator":
                got = ops.double_commutator(A, B, Cc)
            elif name.startswith("hamiltonian_"):
                got = getattr(ops, name)(h, g, *args)
            elif name.startswith("one_elec_op_"):
                got = getattr(ops, name)(h, *args)
            else:
                got = getattr(ops, name)(*args)
        except Exception as e:  # noqa: BLE001
            assert case.get("raises") == type(e).__name__, case["call"]
            continue
        assert "raises" not in case, case["call"]
        ref = {tuple((int(i), bool(dg)) for i, dg in label): v for label, v in case["op"]}
        mine = {k: v for k, v in got.operators.items() if abs(v) > 1e-14}
        assert set(mine) == set(ref), case["call"]
        assert max([abs(mine[k] - ref[k]) for k in ref] or [0.0]) < 1e-12, case["call"]


def test_hamiltonian_folding(golden):
    arrays, _, gops = golden
    h, g = arrays["fold_h"], arrays["fold_g"]
    H = ops.hamiltonian_0i_0a(h, g, 1, 3)
    assert isinstance(H, FermionicOperator)
    folded_ref = json_to_opdict(gops["H0i0a_seed7_1_3_folded"])
    assert len(H.operators) == gops["H0i0a_seed7_1_3_unfolded_count"]["n"]
    assert_opdict_close(H.get_folded_operator(1, 3, 1).operators, folded_ref, 1e-12)
    # closed-form folded integrals reproduce the folded strings: E_core + h_eff E + 1/2 g e
    e_core, h_eff, g_act = ops.fold_hamiltonian_0i_0a(h, g, 1, 3)
    rebuilt = FermionicOperator({(): e_core})
    for p, q in itertools.product(range(3), repeat=2):
        rebuilt += float(h_eff[p, q]) * ops.Epq(p, q)
    for p, q, r, s in itertools.product(range(3), repeat=4):
        rebuilt += (0.5 * float(g_act[p, q, r, s])) * ops.epqrs(p, q, r, s)
    assert_opdict_close(rebuilt.operators, folded_ref, 1e-12)


def test_hamiltonian_folding_unsymmetric_integrals():
    rng = np.random.default_rng(9)
    h = rng.normal(size=(5, 5))
    g = rng.normal(size=(5, 5, 5, 5))
    H = ops.hamiltonian_0i_0a(h, g, 2, 2)
    folded = H.get_folded_operator(2, 2, 1).operators
    assert_opdict_close(folded, orc.op_fold(orc.hamiltonian_0i_0a(h, g, 2, 2), 2, 2, 1), 1e-12)
    e_core, h_eff, g_act = ops.fold_hamiltonian_0i_0a(h, g, 2, 2)
    rebuilt = FermionicOperator({(): e_core})
    for p, q in itertools.product(range(2), repeat=2):
        rebuilt += float(h_eff[p, q]) * ops.Epq(p, q)
    for p, q, r, s in itertools.product(range(2), repeat=4):
        rebuilt += (0.5 * float(g_act[p, q, r, s])) * ops.epqrs(p, q, r, s)
    assert_opdict_close(rebuilt.operators, folded, 1e-12)


def _layout_from_meta(m):
    lay = UpsStructure()
    ansatz = m["ansatz"].lower()
    opts = dict(m["options"])
    n = m["num_active_orbs"]
    args = (
        m["active_occ_idx_shifted"],
        m["active_unocc_idx_shifted"],
        m["active_occ_spin_idx_shifted"],
        m["active_unocc_spin_idx_shifted"],
        n,
        opts,
    )
    if ansatz in ("tups", "qnp"):
        lay.create_tiled(n, opts)
    elif ansatz in ("sdsfuccsd",):
        lay.create_SDSfUCC(*args)
    else:
        lay.create_fUCC(*args)
    return lay


@pytest.mark.parametrize("name", ["tups44", "qnp44", "fuccsd44", "sa44", "tq44", "gsd44", "ksa44", "sds44", "sad65"])
def test_layouts_match_reference(golden, name):
    _, meta, _ = golden
    m = meta[name]
    lay = _layout_from_meta(m)
    assert lay.excitation_operator_type == m["types"]
    assert [list(t) for t in lay.excitation_indices] == m["indices"]
    assert lay.n_params == len(m["types"])
    assert lay.grad_param_R == m["grad_param_R"]


def test_synthetic_layouts_match_reference(golden):
    _, meta, _ = golden
    for name in ("syn_tups_5_23", "syn_tups_6_33", "syn_qnp_6_24"):
        m = meta[name]
        lay = UpsStructure()
        lay.create_tiled(m["n"], dict(m["options"]))
        assert lay.excitation_operator_type == m["types"]
        assert [list(t) for t in lay.excitation_indices] == m["indices"]
    for name, opts in (
        ("syn_gsd_6_33", {"n_layers": 1, "SAGS": True, "GpD": True, "S": True}),
        ("syn_q56_6_33", {"n_layers": 1, "Q": True, "5": True, "6": True}),
    ):
        m = meta[name]
        lay = UpsStructure()
        lay.create_fUCC([0, 1, 2], [3, 4, 5], list(range(6)), list(range(6, 12)), 6, opts)
        assert lay.excitation_operator_type == m["types"]
        assert [list(t) for t in lay.excitation_indices] == m["indices"]


def test_layouts_match_reference_over_option_combinations():
    """312 layouts built by the reference (tests/golden/make_golden_layouts.py): create_fUCC with single flags and seeded random
    subsets of its 13 excitation flags on three index sets (one with unequal spin-orbital sets), create_SDSfUCC with every subset of
    {D, pD, GpD}, create_tiled with every combination of do_tups / do_qnp / skip_last_singles for 2..7 orbitals, the UccStructure
    builders -- types, indices, n_params, grad_param_R, and the same exception type where the reference raises."""
    import gzip
    import json
    import os

    from slowquant_b200.util import UccStructure

    with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_layouts.json.gz"), "rt") as f:
        d = json.load(f)
    assert len(d["cases"]) >= 300
    for x in d["cases"]:
        label = (x["kind"], x.get("options", x.get("builders")))
        try:
            if x["kind"] == "ucc":
                n, occ, unocc, occs, unoccs = d["spaces"][x["space"]]
                lay = UccStructure()
                for b in x["builders"]:
                    getattr(lay, "add_" + b)(*((occ, unocc) if b in ("sa_singles", "sa_doubles") else (occs, unoccs)))
            else:
                lay = UpsStructure()
                if x["kind"] == "tiled":
                    lay.create_tiled(x["n"], dict(x["options"]))
                else:
                    n, occ, unocc, occs, unoccs = d["spaces"][x["space"]]
                    getattr(lay, "create_" + x["kind"])(occ, unocc, occs, unoccs, n, dict(x["options"]))
        except Exception as e:  # noqa: BLE001
            assert x.get("raises") == type(e).__name__, label
            continue
        assert "raises" not in x, label
        assert lay.excitation_operator_type == x["types"], label
        assert [[int(i) for i in t] for t in lay.excitation_indices] == x["indices"], label
        assert lay.n_params == x["n_params"], label
        if "grad_param_R" in x:
            assert {str(k): int(v) for k, v in dict(lay.grad_param_R).items()} == x["grad_param_R"], label


def test_layout_option_errors():
    lay = UpsStructure()
    with pytest.raises(ValueError):
        lay.create_tiled(4, {"n_layers": 1})  # no tiled ansatz specified
    with pytest.raises(ValueError):
        lay.create_tiled(4, {"do_tups": True})  # n_layers missing
    with pytest.raises(ValueError):
        lay.create_tiled(4, {"n_layers": 1, "do_tups": True, "bogus": 1})
    with pytest.raises(ValueError):
        lay.create_fUCC([0], [1], [0, 1], [2, 3], 2, {"n_layers": 1})


def _literal_action(n, label, A, B):
    """The reference's sequential bit-flip / popcount loop (operator_state_algebra.py:112-135) on one
    determinant given as occupation masks (bit o = orbital o)."""
    occ = {}
    for o in range(n):
        occ[2 * o] = (A >> o) & 1
        occ[2 * o + 1] = (B >> o) & 1
    anni = [i for i, d in label if not d]
    crea = [i for i, d in label if d]
    screen = [i for i in crea if i not in anni]
    if any(occ[i] == 0 for i in anni) or any(occ[i] == 1 for i in screen):
        return False, A, B, 0
    phase = 0
    for k in anni + crea:
        occ[k] ^= 1
        phase += sum(occ[j] for j in range(k))
    A2 = sum(occ[2 * o] << o for o in range(n))
    B2 = sum(occ[2 * o + 1] << o for o in range(n))
    return True, A2, B2, 1 - 2 * (phase & 1)


def test_closed_form_string_action_equals_literal_loop():
    n = 5
    lib, h = _host_space(n, 2, 3)
    rng = np.random.default_rng(1)
    labels = [
        ((6, True), (2, False)),
        ((7, True), (3, False)),
        ((4, True), (4, False)),
        ((9, True), (8, True), (3, False), (2, False)),
        ((8, True), (5, True), (5, False), (0, False)),
        ((9, True), (6, True), (2, True), (7, False), (4, False), (1, False)),
        ((3, True), (2, False)),  # spin flip: valid action, target leaves the sector
    ]
    for _ in range(20):
        k = int(rng.integers(1, 4))
        cre = sorted(rng.choice(2 * n, size=k, replace=False).tolist(), reverse=True)
        ann = sorted(rng.choice(2 * n, size=k, replace=False).tolist(), reverse=True)
        labels.append(tuple((int(i), True) for i in cre) + tuple((int(i), False) for i in ann))
    valid, tA, tB, sg = C.c_int(), C.c_uint32(), C.c_uint32(), C.c_int()
    for label in labels:
        flat = np.array([2 * i + (1 if d else 0) for i, d in label], dtype=np.int32)
        for A in range(1 << n):
            for B in range(0, 1 << n, 3):
                _lib.check(
                    lib.sq_debug_string_action(
                        h, flat.ctypes.data_as(C.POINTER(C.c_int32)), len(flat), A, B,
                        C.byref(valid), C.byref(tA), C.byref(tB), C.byref(sg),
                    )
                )
                ok, A2, B2, s = _literal_action(n, label, A, B)
                assert bool(valid.value) == ok, (label, A, B)
                if ok:
                    assert (tA.value, tB.value, sg.value) == (A2, B2, s), (label, A, B)
    lib.sq_space_destroy(h)


def test_host_only_space_refuses_kernels():
    lib, h = _host_space(4, 2, 2)
    lay = C.c_void_p()
    codes = np.array([0], dtype=np.int32)
    offs = np.array([0, 2], dtype=np.int32)
    flat = np.array([0, 1], dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    # a host-only layout exists for planning (work lists, exchange plan) ...
    _lib.check(lib.sq_layout_create(h, 1, p(codes), p(offs), p(flat), C.byref(lay)))
    assert lib.sq_layout_num_launches(lay, 0, 1) == 1
    # ... but no kernel may run on it
    th = np.array([0.1])
    with pytest.raises(ValueError):
        _lib.check(lib.sq_ups_apply(h, lay, th.ctypes.data_as(C.POINTER(C.c_double)), 0, 1, 0, C.c_void_p(8), None))
    lib.sq_layout_destroy(lay)
    lib.sq_space_destroy(h)


def _tups_layout(lib, h, n, L):
    from slowquant_b200.util import UpsStructure

    lay = UpsStructure()
    lay.create_tiled(n, {"n_layers": L, "do_tups": True})
    codes = np.array([_lib.EXC_CODES[t] for t in lay.excitation_operator_type], dtype=np.int32)
    offs = np.zeros(len(codes) + 1, dtype=np.int32)
    flat = []
    for k, idx in enumerate(lay.excitation_indices):
        flat.extend(int(x) for x in idx)
        offs[k + 1] = len(flat)
    flat = np.asarray(flat, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    handle = C.c_void_p()
    _lib.check(lib.sq_layout_create(h, len(codes), p(codes), p(offs), p(flat), C.byref(handle)))
    return handle, len(codes)


def test_launch_planner_on_host_only_space():
    """The commutation-aware planner (window sweeps) is pure host logic: every brick is executed exactly once,
    a tUPS circuit needs far fewer sweeps than bricks, and the run-time switch turns the window sweeps off."""
    n, L = 12, 6
    lib, h = _host_space(n, n // 2, n // 2)
    lay, P = _tups_layout(lib, h, n, L)
    stats = (C.c_int64 * 6)()
    try:
        _lib.check(lib.sq_set_option(b"win", b"1"))
        _lib.check(lib.sq_layout_plan_stats(lay, 0, P, stats))
        launches, sweeps, bricks_in, quads, singles, other = (int(x) for x in stats)
        n_bricks = (n - 1) * L
        assert other == 0
        assert bricks_in + 2 * quads + singles == n_bricks          # every brick exactly once
        assert sweeps > 0 and launches == sweeps + quads + singles
        assert launches <= n_bricks // 4                               # >= 4 bricks per launch on average
        # kernel launches = plan launches + the two gauge sweeps around the window sweeps
        assert lib.sq_layout_num_launches(lay, 0, P) == launches + 2
        # amplitudes read+written: every sweep touches the whole vector
        ndet = int(lib.sq_space_num_det(h))
        assert lib.sq_layout_touched_amplitudes(lay, 0, P) >= (sweeps + 2) * ndet
        # a sub-range and a single operator: isolated bricks do not pay gauge sweeps
        assert lib.sq_layout_num_launches(lay, 0, 1) == 1
        _lib.check(lib.sq_set_option(b"win", b"0"))
        _lib.check(lib.sq_layout_plan_stats(lay, 0, P, stats))
        assert int(stats[1]) == 0 and 2 * int(stats[3]) + int(stats[4]) == n_bricks
        # narrower windows, brick cap 4
        _lib.check(lib.sq_set_option(b"win", b"4:3:0,72,2,4,2"))
        _lib.check(lib.sq_layout_plan_stats(lay, 0, P, stats))
        assert int(stats[2]) + 2 * int(stats[3]) + int(stats[4]) == n_bricks
        assert int(stats[1]) > 0 and int(stats[2]) <= 4 * int(stats[1])
        assert lib.sq_set_option(b"nonsense", b"1") == _lib.SQ_ERR_INVALID
    finally:
        _lib.check(lib.sq_set_option(b"win", b"1"))
        lib.sq_layout_destroy(lay)
        lib.sq_space_destroy(h)


def _layout_handle(lib, h, lay):
    codes = np.array([_lib.EXC_CODES[t] for t in lay.excitation_operator_type], dtype=np.int32)
    offs = np.zeros(len(codes) + 1, dtype=np.int32)
    flat = []
    for k, idx in enumerate(lay.excitation_indices):
        flat.extend(int(x) for x in idx)
        offs[k + 1] = len(flat)
    flat = np.asarray(flat, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    handle = C.c_void_p()
    _lib.check(lib.sq_layout_create(h, len(codes), p(codes), p(offs), p(flat), C.byref(handle)))
    return handle


def _spatial_orbitals(exc_type, idx):
    return {int(x) for x in idx} if exc_type.startswith("sa_") else {int(x) // 2 for x in idx}


@pytest.mark.parametrize(
    "n,na,nb,options,win",
    [
        (8, 4, 4, {"n_layers": 5, "do_tups": True}, b"1"),
        (10, 5, 5, {"n_layers": 7, "do_qnp": True}, b"1"),
        (11, 5, 6, {"n_layers": 4, "do_tups": True, "skip_last_singles": True}, b"1"),
        (12, 6, 6, {"n_layers": 9, "do_tups": True}, b"1"),
        (12, 6, 6, {"n_layers": 5, "do_tups": True}, b"4:3:0,72,2,4,2"),
        (9, 4, 5, {"n_layers": 6, "do_tups": True}, b"0"),
    ],
)
def test_launch_plan_is_a_valid_reordering(n, na, nb, options, win):
    """The planner may only move an operator past operators it commutes with.  Export the plan (sq_layout_plan_export) and check,
    for forward and adjoint circuits, full ranges, sub-ranges and with some angles exactly zero: (1) every active operator runs
    exactly once, skipped ones (|theta| < 1e-28, osa.py:998) never; (2) any two operators that share a spatial orbital keep the
    order of the circuit (reversed for the adjoint); (3) launches are contiguous in the exported order."""
    from slowquant_b200.util import UpsStructure

    lib, h = _host_space(n, na, nb)
    ups = UpsStructure()
    ups.create_tiled(n, dict(options))
    # a few generic operators in the middle: they split the circuit into stretches the planner treats separately
    P0 = ups.n_params
    types, idxs = list(ups.excitation_operator_type), [tuple(i) for i in ups.excitation_indices]
    mid = P0 // 2
    types[mid:mid] = ["single", "double"]
    idxs[mid:mid] = [(0, 2 * (n - 1)), (0, 1, 2 * (n - 2), 2 * (n - 1) + 1)]
    ups.excitation_operator_type, ups.excitation_indices, ups.n_params = types, idxs, len(types)
    lay = _layout_handle(lib, h, ups)
    P = len(types)
    orbs = [_spatial_orbitals(t, i) for t, i in zip(types, idxs)]
    rng = np.random.default_rng(n * 100 + P)
    ops_out = np.empty(P, dtype=np.int32)
    launch_out = np.empty(P, dtype=np.int32)
    n_out = C.c_int(0)
    pi32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    try:
        _lib.check(lib.sq_set_option(b"win", win))
        for first, last, zero_some in ((0, P, False), (0, P, True), (3, P - 5, False), (mid - 7, mid + 9, True)):
            th = rng.uniform(-np.pi, np.pi, P)
            if zero_some:
                th[rng.choice(P, size=P // 6, replace=False)] = 0.0
            for dagger in (0, 1):
                _lib.check(
                    lib.sq_layout_plan_export(
                        lay, th.ctypes.data_as(C.POINTER(C.c_double)), first, last, dagger, pi32(ops_out), pi32(launch_out), P, C.byref(n_out)
                    )
                )
                m = n_out.value
                planned = ops_out[:m].tolist()
                active = [k for k in range(first, last) if abs(th[k]) >= 1e-28]
                assert sorted(planned) == active                                   # (1)
                pos = {k: i for i, k in enumerate(planned)}
                for a_i, ka in enumerate(active):                                  # (2)
                    for kb in active[a_i + 1 :]:
                        if orbs[ka] & orbs[kb]:
                            assert (pos[ka] < pos[kb]) == (not dagger), (ka, kb, dagger, types[ka], types[kb])
                la = launch_out[:m]
                assert np.all(np.diff(la) >= 0) and (m == 0 or la[0] == 0)         # (3)
                if win == b"1" and n >= 10 and (first, last) == (0, P) and not zero_some:
                    # the check above is not vacuous: the plan is time-skewed (not the circuit order) and fuses many operators
                    assert planned != (active if not dagger else active[::-1]) and int(la[-1]) + 1 < m // 6
        assert lib.sq_layout_plan_export(lay, None, 0, P, 0, pi32(ops_out), pi32(launch_out), 3, C.byref(n_out)) == _lib.SQ_ERR_INVALID
    finally:
        _lib.check(lib.sq_set_option(b"win", b"1"))
        lib.sq_layout_destroy(lay)
        lib.sq_space_destroy(h)


@pytest.mark.parametrize("n,na,nb,L,qnp", [(6, 3, 3, 4, False), (7, 3, 4, 3, True), (8, 4, 4, 3, False)])
def test_launch_plan_order_gives_the_same_state(n, na, nb, L, qnp):
    """Numerical twin of test_launch_plan_is_a_valid_reordering: applying the operators one by one in the ORDER OF THE PLAN
    (oracle, CPU) gives the state of the circuit order -- forward and adjoint, with some angles zero."""
    from slowquant_b200.util import UpsStructure

    lib, h = _host_space(n, na, nb)
    ups = UpsStructure()
    ups.create_tiled(n, {"n_layers": L, "do_qnp": True} if qnp else {"n_layers": L, "do_tups": True})
    lay = _layout_handle(lib, h, ups)
    P = ups.n_params
    types, idxs = list(ups.excitation_operator_type), [tuple(int(x) for x in i) for i in ups.excitation_indices]
    rng = np.random.default_rng(90 + n)
    th = rng.uniform(-np.pi, np.pi, P)
    th[rng.choice(P, size=P // 8, replace=False)] = 0.0
    sp = orc.get_indexing(0, n, 0, na, nb)
    state = rng.normal(size=sp.num_det)
    state /= np.linalg.norm(state)
    ops_out = np.empty(P, dtype=np.int32)
    launch_out = np.empty(P, dtype=np.int32)
    n_out = C.c_int(0)
    pi32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    try:
        _lib.check(lib.sq_set_option(b"win", b"1"))
        for dagger in (False, True):
            ref = orc.construct_ups_state(state, sp, th, types, idxs, dagger=dagger)
            _lib.check(
                lib.sq_layout_plan_export(lay, th.ctypes.data_as(C.POINTER(C.c_double)), 0, P, int(dagger), pi32(ops_out), pi32(launch_out), P, C.byref(n_out))
            )
            cur = state.copy()
            for k in ops_out[: n_out.value]:
                k = int(k)
                cur = orc.construct_ups_state(cur, sp, th[k : k + 1], types[k : k + 1], idxs[k : k + 1], dagger=dagger)
            assert np.max(np.abs(cur - ref)) < 1e-13
            assert int(launch_out[n_out.value - 1]) + 1 < n_out.value        # the plan fuses operators
    finally:
        lib.sq_layout_destroy(lay)
        lib.sq_space_destroy(h)


def test_determinant_expansion_on_hf_matches_reference():
    """get_determinant_expansion_from_operator_on_HF (osa.py:2979-3033) through the engine's closed-form string action, against
    outputs of the reference (tests/golden/make_golden_strings.py): excitation generators, spin-adapted doubles, number
    operators (creator == annihilator) and products, on CAS(4,4) and CAS(4,5) with 3 alpha / 1 beta electrons."""
    import json
    import os

    from slowquant_b200.operator_state_algebra import get_determinant_expansion_from_operator_on_HF

    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_hf_expansion.json")))
    assert len(cases) == 14
    n_terms = 0
    for case in cases:
        op = FermionicOperator({tuple((int(i), bool(d)) for i, d in label): float(v) for label, v in case["operator"]})
        coeffs, dets = get_determinant_expansion_from_operator_on_HF(op, *case["space"])
        assert dets == case["dets"], case["name"]
        assert np.allclose(coeffs, case["coeffs"], rtol=0, atol=1e-15), case["name"]
        n_terms += len(dets)
    assert n_terms > 20


def test_qiskit_form_matches_reference():
    """FermionicOperator.get_qiskit_form / operator_to_qiskit_key (fermionic_operator.py:7-24, 357-377): key strings in blocked
    spin order and coefficients against reference outputs (string formatting only, no Qiskit)."""
    import json
    import os

    from slowquant_b200.fermionic_operator import operator_to_qiskit_key

    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_qiskit_form.json")))
    for case in cases:
        op = FermionicOperator({tuple((int(i), bool(d)) for i, d in label): float(v) for label, v in case["operator"]})
        got = op.get_qiskit_form(case["num_orbs"])
        assert list(got.keys()) == [k for k, _ in case["qiskit_form"]]
        assert np.allclose(list(got.values()), [v for _, v in case["qiskit_form"]], rtol=0, atol=0)
    assert operator_to_qiskit_key(((3, True), (0, False)), {0: 0, 3: 5}) == "+_5 -_0"


def test_lr_orbital_blocks_match_reference():
    """RDM-only linear-response orbital blocks (reference density_matrix.py:233-563) against outputs of the reference
    itself on seeded random h, g, x, rdm1, rdm2 (tests/golden/make_golden_lr.py), incl. no-inactive / no-virtual spaces."""
    from slowquant_b200 import density_matrix as dm

    gold = np.load(f"{ROOT}/tests/golden/golden_lr.npz")
    for case in range(3):
        pre = f"c{case}_"
        nI, nA, nV, n_exc = (int(x) for x in gold[pre + "dims"])
        h, g, x = gold[pre + "h"], gold[pre + "g"], gold[pre + "x"]
        rdm1, rdm2, kappa, resp = gold[pre + "rdm1"], gold[pre + "rdm2"], gold[pre + "kappa"], gold[pre + "resp"]
        kd = kappa[:, ::-1].copy()
        assert np.max(np.abs(dm.get_orbital_gradient_response(h, g, kappa, nI, nA, rdm1, rdm2) - gold[pre + "grad_response"])) < 1e-11
        assert np.max(np.abs(dm.get_orbital_response_metric_sigma(kappa, nI, nA, rdm1) - gold[pre + "metric_sigma"])) < 1e-13
        for s in range(4):
            assert abs(dm.get_orbital_response_vector_norm(kappa, nI, nA, rdm1, resp, s, n_exc) - gold[pre + "vector_norm"][s]) < 1e-11
            assert abs(dm.get_orbital_response_property_gradient(x, kappa, nI, nA, rdm1, resp, s, n_exc) - gold[pre + "property_gradient"][s]) < 1e-11
        assert np.max(np.abs(dm.get_orbital_response_hessian_block(h, g, kd, kappa, nI, nA, rdm1, rdm2) - gold[pre + "hessian_A"])) < 1e-10
        assert np.max(np.abs(dm.get_orbital_response_hessian_block(h, g, kd, kd, nI, nA, rdm1, rdm2) - gold[pre + "hessian_B"])) < 1e-10


def test_extended_space_tables_bit_exact():
    """get_indexing_extended (ci_spaces.py:119-259): idx2det bit-exact against the reference's lists for seven spaces
    (orders 1 and 2, empty inactive / virtual spaces, unequal spin counts), det2idx, and the embedding into the parent
    product space of all orbitals (host-only space, device = -1)."""
    from slowquant_b200.ci_spaces import extended_idx2det, get_indexing_extended

    g = np.load(f"{ROOT}/tests/golden/golden_extended.npz")
    for k, sp in enumerate(g["spaces"]):
        sp = tuple(int(x) for x in sp)
        ref = g[f"idx2det_{k}"]
        assert np.array_equal(extended_idx2det(*sp), ref), sp
        ci = get_indexing_extended(*sp, device=-1)
        nI, nA, nV, na, nb, _ = sp
        assert (ci.num_inactive_orbs, ci.num_active_orbs, ci.num_virtual_orbs) == (0, nI + nA + nV, 0)
        assert (ci.num_active_elec_alpha, ci.num_active_elec_beta) == (na + nI, nb + nI)
        assert ci.space_extension_offset == nI and ci.num_det == len(ref)
        assert all(ci.det2idx[int(d)] == i for i, d in enumerate(ref))
        assert np.array_equal(ci.parent.idx2det[ci.embedding], ref)      # same determinants, parent numbering
        assert len(set(ci.embedding.tolist())) == len(ref)
    with pytest.raises(ValueError):
        extended_idx2det(1, 2, 1, 1, 1, 3)


def test_table_free_E_records_equal_the_table():
    """sq_set_option("etab", "alu"): the record of E_pq the table-free panel kernels compute from (p, q, spin) equals, field by
    field, the table entry built from the closed-form string action (which test_closed_form_string_action_equals_literal_loop
    pins to the reference's flip / popcount loop), for every (p, q, spin) of several active-space sizes."""
    for n, na, nb in [(2, 1, 1), (5, 3, 1), (8, 4, 4), (16, 8, 8), (20, 10, 10), (26, 2, 2)]:
        lib, h = _host_space(n, na, nb)
        bad = C.c_int(-1)
        _lib.check(lib.sq_debug_etab_closed_form(h, C.byref(bad)))
        assert bad.value == 0, (n, bad.value)
        lib.sq_space_destroy(h)


def test_c_abi_rejects_malformed_arguments_without_crashing():
    """Every entry point returns a status and never throws or dereferences a bad pointer (include/sqsv.h contract): malformed
    index offsets (decreasing, negative, too long), unknown excitation codes, out-of-range indices, NULL pointers, bad ranges,
    bad partitions.  Host-only space, no kernels."""
    lib, h = _host_space(4, 2, 2)
    E = _lib.EXC_CODES
    pi32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    pi64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))  # noqa: E731

    def create(codes, offs, flat):
        codes, offs = np.asarray(codes, dtype=np.int32), np.asarray(offs, dtype=np.int32)
        flat = np.asarray(flat if len(flat) else [0], dtype=np.int32)
        hl = C.c_void_p()
        st = lib.sq_layout_create(h, len(codes), pi32(codes), pi32(offs), pi32(flat), C.byref(hl))
        if st == 0:
            lib.sq_layout_destroy(hl)
        return st

    assert create([E["sa_single"]], [0, 2], [0, 1]) == 0 and create([], [0], []) == 0
    for codes, offs, flat in (
        ([99], [0, 2], [0, 1]), ([E["sa_single"]], [0, 2], [-1, 1]), ([E["sa_single"]], [0, 2], [0, 7]), ([E["sa_single"]], [0, 3], [0, 1, 2]),
        ([E["double"]], [0, 2], [0, 1]), ([E["double"]], [0, 4], [0, 1, 2, 99]), ([E["sa_single"]], [0, 2], [1, 1]),
        ([E["sa_single"], E["sa_single"]], [0, 2, 1], [0, 1, 2, 3]), ([E["sa_single"]], [-2, 0], [0, 1, 2, 3]), ([E["sextuple"]], [0, 14], list(range(14))),
    ):
        assert create(codes, offs, flat) == _lib.SQ_ERR_INVALID, (codes, offs, flat)
    assert lib.sq_layout_create(h, 1, None, None, None, C.byref(C.c_void_p())) == _lib.SQ_ERR_INVALID
    assert lib.sq_layout_create(None, 0, None, None, None, C.byref(C.c_void_p())) == _lib.SQ_ERR_INVALID
    lay, P = _tups_layout(lib, h, 4, 1)
    out6 = (C.c_int64 * 6)()
    assert lib.sq_layout_plan_stats(lay, 0, P, out6) == 0
    for f, l in ((-1, 3), (0, P + 1), (2, 1)):
        assert lib.sq_layout_plan_stats(lay, f, l, out6) == _lib.SQ_ERR_INVALID
    assert lib.sq_layout_plan_stats(None, 0, 1, out6) != 0 and lib.sq_layout_plan_stats(lay, 0, 1, None) != 0
    assert lib.sq_layout_num_launches(lay, 0, P + 5) == -1 and lib.sq_layout_touched_amplitudes(lay, 5, 2) == -1
    ops, cf = np.array([1, 4], dtype=np.int32), np.array([1.0])
    pd = cf.ctypes.data_as(C.POINTER(C.c_double))
    for offs in ([0, 2], [0, -1], [2, 0], [0, 40]):     # host-only space refuses kernels; bad offsets must not be dereferenced either
        o = np.array(offs, dtype=np.int32)
        assert lib.sq_apply_strings(h, 1, pi32(ops), pi32(o), pd, C.c_void_p(8), C.c_void_p(16), 0, 0, None) != 0
    d2, o2 = np.array([5, 6], dtype=np.int64), np.empty(2, dtype=np.int64)
    assert lib.sq_space_det2idx(h, 2, None, pi64(o2)) != 0 and lib.sq_space_det2idx(None, 2, pi64(d2), pi64(o2)) != 0
    assert lib.sq_space_export_idx2det(h, -1, 5, pi64(o2)) != 0 and lib.sq_space_export_idx2det(h, 0, 10**9, pi64(o2)) != 0
    assert lib.sq_space_export_strings(h, 2, None) != 0 and lib.sq_space_export_strings(h, 0, None) != 0
    assert lib.sq_space_num_strings(None, 0) == -1 and lib.sq_space_num_det(None) == -1
    rs = np.zeros(17, dtype=np.int64)
    for args in ((4, 2, 0), (4, 2, 3), (4, 2, 64), (0, 0, 2), (4, 9, 2), (4, -1, 2)):
        assert lib.sq_partition_prefix(*args, pi64(rs)) == _lib.SQ_ERR_INVALID, args
    assert lib.sq_partition_prefix(4, 2, 2, pi64(rs)) == 0 and list(rs[:3]) == [0, 3, 6]
    bad = np.array([0, 5, 3], dtype=np.int64)
    assert lib.sq_space_set_partition(h, 2, 0, pi64(bad)) != 0 and lib.sq_space_set_partition(h, 2, 5, pi64(rs)) != 0
    assert lib.sq_space_set_partition(h, 0, 0, None) != 0
    v, sgn, ta, tb = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_uint32()
    big = np.array([2 * 40 + 1, 0], dtype=np.int32)
    for ops_ptr, n_ops in ((pi32(big), 2), (pi32(big), 99), (None, 2)):
        assert lib.sq_debug_string_action(h, ops_ptr, n_ops, 3, 3, C.byref(v), C.byref(ta), C.byref(tb), C.byref(sgn)) == _lib.SQ_ERR_INVALID
    ex, n_out = np.empty(8, dtype=np.int32), C.c_int()
    assert lib.sq_layout_plan_export(lay, None, 0, P + 4, 0, pi32(ex), pi32(ex), 8, C.byref(n_out)) != 0
    assert lib.sq_layout_plan_export(None, None, 0, 3, 0, pi32(ex), pi32(ex), 8, C.byref(n_out)) != 0
    assert lib.sq_set_option(None, b"1") != 0
    # generator strings of an sa_double operator: bad operator index / kind, NULL pointers, bad offsets, spin orbital out of range
    codes, offs, flat = np.array([E["sa_double_1"], E["sa_single"]], dtype=np.int32), np.array([0, 4, 6], dtype=np.int32), np.array([0, 1, 2, 3, 0, 1], dtype=np.int32)
    lay2 = C.c_void_p()
    assert lib.sq_layout_create(h, 2, pi32(codes), pi32(offs), pi32(flat), C.byref(lay2)) == 0
    gops, goffs = np.array([2 * 4 + 1, 0], dtype=np.int32), np.array([0, 2], dtype=np.int32)
    assert lib.sq_layout_attach_generator(lay2, 0, 1, pi32(gops), pi32(goffs), pd) == 0
    for k, n_str, po, poff, pc in (
        (5, 1, pi32(gops), pi32(goffs), pd), (-1, 1, pi32(gops), pi32(goffs), pd), (1, 1, pi32(gops), pi32(goffs), pd), (0, 1, None, None, None),
        (0, 1, pi32(gops), pi32(np.array([2, 0], dtype=np.int32)), pd), (0, 1, pi32(gops), pi32(np.array([-2, 0], dtype=np.int32)), pd),
        (0, -3, pi32(gops), pi32(goffs), pd), (0, 1, pi32(np.array([2 * 40 + 1, 0], dtype=np.int32)), pi32(goffs), pd),
    ):
        assert lib.sq_layout_attach_generator(lay2, k, n_str, po, poff, pc) == _lib.SQ_ERR_INVALID
    assert lib.sq_layout_attach_generator(None, 0, 1, pi32(gops), pi32(goffs), pd) == _lib.SQ_ERR_INVALID
    assert lib.sq_layout_needs_exchange(lay2, -1, 9) == -1 and lib.sq_layout_needs_exchange(None, 0, 1) == -1
    assert lib.sq_layout_op_stats(lay2, 7, (C.c_int64 * 6)()) != 0 and lib.sq_layout_op_stats(lay2, 0, None) != 0
    assert lib.sq_layout_num_ops(lay2) == 2 and lib.sq_layout_num_ops(None) == -1
    lib.sq_layout_destroy(lay2)
    lib.sq_layout_destroy(lay)
    lib.sq_layout_destroy(None)
    lib.sq_space_destroy(h)
    lib.sq_space_destroy(None)


def test_runtime_switches_are_known():
    """sq_set_option: every switch named in include/sqsv.h is accepted, anything else is an error (host call, no kernel)."""
    lib = _lib.load()
    for name, value in (
        (b"win", b"1"), (b"wingrad", b"0"), (b"pipeline", b"1"), (b"etab", b"smem"), (b"rows", b"0"), (b"rows_cfg", b"1024,0"),
        (b"panel", b"0"), (b"rdm_tri", b"0"), (b"etab", b"alu"), (b"etab", b"smem"),
    ):
        assert lib.sq_set_option(name, value) == 0, name
    assert lib.sq_set_option(b"no-such-switch", b"1") != 0
    header = open(f"{ROOT}/include/sqsv.h").read()
    for name in ("win", "wingrad", "pipeline", "etab", "rows", "rows_cfg", "panel", "rdm_tri"):
        assert f'"{name}"' in header, f"switch {name} is not documented in include/sqsv.h"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys,
    at a small CAS so that it finishes in seconds."""
    import json
    import subprocess
    import sys

    res = subprocess.run(
        [sys.executable, f"{ROOT}/bench.py", "--impl", "reference", "--cas", "8", "--layers", "2", "--steps", "1", "--warmup", "0"],
        capture_output=True, text=True, timeout=300,
    )
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "layers/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["dtype"] == "f64" and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "layers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
