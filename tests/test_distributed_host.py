"""Host-side logic of the alpha-sharded path, exercised with world_size-2/4 gloo process groups on CPU:
partition, per-rank work lists, the exchange plan (identical on every rank) and complete, non-overlapping
coverage of every orbital-pair tile.  No kernels run (host-only spaces, device = -1)."""
import ctypes as C
import os
import socket
from math import comb

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from slowquant_b200 import _lib


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, n: int, na: int, nb: int, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from slowquant_b200.distributed import ShardedSpace
        from slowquant_b200.util import UpsStructure

        sp = ShardedSpace(0, n, 0, na, nb, device=-1, rank=rank, world=world)
        lay = UpsStructure()
        lay.create_tiled(n, {"n_layers": 1, "do_tups": True})
        plan = sp.exchange_plan(lay, 0, lay.n_params, False)
        plan_d = sp.exchange_plan(lay, 0, lay.n_params, True)
        lib = _lib.load()
        handle = __import__("slowquant_b200.operator_state_algebra", fromlist=["x"]).compile_layout(sp.ci_info, lay)
        stats = []
        for k in range(lay.n_params):
            out = np.zeros(6, dtype=np.int64)
            _lib.check(lib.sq_layout_op_stats(handle, k, out.ctypes.data_as(C.POINTER(C.c_int64))))
            stats.append(out.tolist())
        mine = {"rank": rank, "rows": (sp.row_begin, sp.row_end), "plan": plan, "plan_d": plan_d, "stats": stats}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            ret["all"] = gathered
            ret["types"] = list(lay.excitation_operator_type)
            ret["indices"] = [tuple(t) for t in lay.excitation_indices]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,na,nb", [(2, 6, 3, 3), (4, 7, 3, 4), (2, 5, 2, 2)])
def test_sharded_work_lists_cover_every_tile(world, n, na, nb):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, na, nb, ret), nprocs=world, join=True)
    allr = sorted(ret["all"], key=lambda d: d["rank"])
    # partition: contiguous, complete, prefix classes
    assert allr[0]["rows"][0] == 0 and allr[-1]["rows"][1] == comb(n, na)
    for a, b in zip(allr[:-1], allr[1:]):
        assert a["rows"][1] == b["rows"][0]
    # the exchange plan is the same on every rank, covers all operators, in order
    for key in ("plan", "plan_d"):
        for d in allr[1:]:
            assert d[key] == allr[0][key]
    plan = allr[0]["plan"]
    assert plan[0][0] == 0 and plan[-1][1] == len(ret["types"])
    for (f0, l0, _), (f1, l1, _) in zip(plan[:-1], plan[1:]):
        assert l0 == f1
    assert [(f, l) for f, l, _ in allr[0]["plan_d"]] == [(f, l) for f, l, _ in plan[::-1]]
    k_bits = world.bit_length() - 1
    for k, (t, idx) in enumerate(zip(ret["types"], ret["indices"])):
        p = idx[0] if t == "sa_single" else idx[0] // 2
        crosses = p < k_bits           # pair (p, p+1) changes the occupation of the first log2(world) orbitals
        for d in allr:
            assert bool(d["stats"][k][4]) == crosses, (k, t, idx)
        in_exchange_range = any(f <= k < l and x for f, l, x in plan)
        assert in_exchange_range == crosses
        # coverage: every (src,tgt) row pair is worked on either by one rank entirely, or by two ranks that
        # split its columns; every inert row by exactly its owner
        n_pairs = comb(n - 2, na - 1)
        n_inert = comb(n, na) - 2 * n_pairs
        local_pairs = sum(d["stats"][k][1] for d in allr)
        cross_items = sum(d["stats"][k][3] for d in allr)
        assert cross_items % 2 == 0
        assert local_pairs + cross_items // 2 == n_pairs
        assert sum(d["stats"][k][2] for d in allr) == n_inert
        if not crosses:
            assert cross_items == 0
        # touched amplitudes add up to the single-device count
        nbeta = comb(n, nb)
        nsrc_b = comb(n - 2, nb - 1)
        assert sum(d["stats"][k][5] for d in allr) == 2 * n_pairs * nbeta + n_inert * 2 * nsrc_b


def test_partition_prefix_matches_string_order():
    from slowquant_b200.distributed import partition_prefix

    lib = _lib.load()
    for n, na, world in [(6, 3, 2), (7, 3, 4), (8, 4, 8), (5, 1, 4)]:
        starts = partition_prefix(n, na, world)
        h = C.c_void_p()
        _lib.check(lib.sq_space_create(n, na, na, -1, 0, -1, C.byref(h)))
        NA = lib.sq_space_num_strings(h, 0)
        masks = np.empty(NA, dtype=np.uint32)
        _lib.check(lib.sq_space_export_strings(h, 0, masks.ctypes.data_as(C.POINTER(C.c_uint32))))
        k = world.bit_length() - 1
        prefix = masks & ((1 << k) - 1)
        for r in range(world):
            block = prefix[starts[r] : starts[r + 1]]
            assert len(set(block.tolist())) <= 1, "a rank's rows must share the occupation of the first orbitals"
        assert starts[-1] == NA
        lib.sq_space_destroy(h)
    with pytest.raises(ValueError):
        partition_prefix(6, 3, 3)


def test_shift_rule_is_exact_on_trigonometric_polynomials():
    from slowquant_b200.distributed import shift_rule

    rng = np.random.default_rng(5)
    for R in (1, 2, 3, 4):
        l = np.arange(-R, R + 1)
        c = rng.normal(size=2 * R + 1) + 1j * rng.normal(size=2 * R + 1)
        x, w = shift_rule(R)
        assert len(x) == 2 * R
        fprime = sum(wi * np.sum(c * np.exp(1j * l * xi)) for xi, wi in zip(x, w))
        assert abs(fprime - np.sum(1j * l * c)) < 1e-13


@pytest.mark.parametrize("layout", ["tups", "qnp", "generic"])
def test_sharded_gradient_composition_equals_literal_loop(layout):
    """The arithmetic of distributed.energy_and_theta_gradient_sharded -- g_k = 2 sum_mu w_mu <bra|exp(x_mu T_k)|ket> with the
    frequency table _AMPLITUDE_FREQUENCIES, then both vectors advanced by U_k -- on the oracle's CPU primitives, against
    the oracle's restatement of the reference's gradient loop (ups_wavefunction.py:1091-1138, get_grad_action per step)."""
    from oracle import sq_oracle as orc
    from slowquant_b200.distributed import _AMPLITUDE_FREQUENCIES, shift_rule

    rng = np.random.default_rng(11)
    n, na, nb = 4, 2, 2
    sp = orc.get_indexing(0, n, 0, na, nb)
    if layout == "generic":
        types = ["single", "sa_single", "double", "sa_double_1", "double", "single", "sa_single"]
        idx = [(0, 4), (1, 3), (0, 1, 4, 5), (0, 1, 2, 3), (2, 3, 6, 7), (3, 7), (0, 2)]
    else:
        types, idx = orc.tiled_layout(n, 2, do_qnp=(layout == "qnp"))
    P = len(types)
    th = rng.uniform(-np.pi, np.pi, P)
    th[1] = 0.0                                        # a zero angle: gradient without rotation
    h = rng.normal(size=(n, n))
    g = 0.1 * rng.normal(size=(n, n, n, n))
    h, g = h + h.T, g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    csf = rng.normal(size=sp.num_det)
    csf /= np.linalg.norm(csf)
    ref = orc.theta_gradient(csf, th, types, idx, h, g, sp)
    ci = orc.construct_ups_state(csf, sp, th, types, idx)
    bra = orc.propagate_state([orc.hamiltonian_0i_0a(h, g, 0, n)], ci, sp)
    bra = orc.construct_ups_state(bra, sp, th, types, idx, dagger=True)
    ket = csf.copy()
    grad = np.zeros(P)
    for k in range(P):
        xs, ws = shift_rule(_AMPLITUDE_FREQUENCIES[types[k]])
        grad[k] = 2.0 * sum(w * (bra @ orc.construct_ups_state(ket, sp, [x], types[k : k + 1], idx[k : k + 1])) for x, w in zip(xs, ws))
        bra = orc.construct_ups_state(bra, sp, th[k : k + 1], types[k : k + 1], idx[k : k + 1])
        ket = orc.construct_ups_state(ket, sp, th[k : k + 1], types[k : k + 1], idx[k : k + 1])
    assert np.max(np.abs(grad - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_gradient_segments_cover_the_circuit_once():
    """distributed.gradient_segments: the stretches tile [0, P) in order; only local stretches made of brick operators are fused;
    fused_local=False sends everything to the shift rule."""
    from slowquant_b200.distributed import gradient_segments
    from slowquant_b200.util import UpsStructure

    ups = UpsStructure()
    ups.create_tiled(6, {"n_layers": 2, "do_tups": True})
    types, idxs = list(ups.excitation_operator_type), [tuple(i) for i in ups.excitation_indices]
    types[10:10] = ["single"]                       # a generic operator inside a local stretch
    idxs[10:10] = [(4, 10)]
    ups.excitation_operator_type, ups.excitation_indices, ups.n_params = types, idxs, len(types)
    P = len(types)
    # a plan as exchange_plan builds it: maximal ranges, exchange ranges hold one orbital pair
    plan = [(0, 3, True), (3, 16, False), (16, 19, True), (19, P, False)]
    seg = gradient_segments(plan, ups)
    assert seg[0][0] == 0 and seg[-1][1] == P and all(a[1] == b[0] for a, b in zip(seg, seg[1:]))
    assert [s[2] for s in seg] == ["shift", "shift", "shift", "fused"]      # (3, 16) holds the generic single -> not fused
    assert all(s[2] == "shift" for s in gradient_segments(plan, ups, fused_local=False))
    assert [s[2] for s in gradient_segments(plan, ups, peer_gradient=True)] == ["peer", "shift", "peer", "fused"]
    del types[10], idxs[10]
    ups.n_params = len(types)
    seg = gradient_segments([(0, 3, True), (3, 15, False), (15, 18, True), (18, P - 1, False)], ups)
    assert [s[2] for s in seg] == ["shift", "fused", "shift", "fused"]


# ---- re-sharding route: phase planner, row tables, constrained (layout B) spaces ---------------------------------------
def _circuits():
    from oracle import sq_oracle as orc

    out = []
    for n, L, qnp in [(6, 3, False), (8, 4, False), (9, 2, True), (16, 16, False), (20, 2, False)]:
        types, idx = orc.tiled_layout(n, L, do_qnp=qnp)
        out.append((n, list(types), [tuple(int(x) for x in t) for t in idx]))
    # generic operators in between (single / double on spin orbitals, an sa_double), incl. one that neither layout can run
    n = 8
    types, idx = orc.tiled_layout(n, 1)
    types, idx = list(types), [tuple(int(x) for x in t) for t in idx]
    types[5:5] = ["single", "double", "sa_double_1", "double"]
    idx[5:5] = [(4, 8), (2, 5, 10, 13), (2, 3, 4, 5), (0, 1, 14, 15)]
    out.append((n, types, idx))
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
def test_reshard_schedule_is_an_equivalent_reordering(world):
    from slowquant_b200.distributed import operator_orbitals, reshard_schedule

    k = world.bit_length() - 1
    for n, types, idx in _circuits():
        if 2 * k > n:
            continue
        for dagger in (False, True):
            for first, last in [(0, len(types)), (2, len(types) - 3)]:
                phases = reshard_schedule(types, idx, n, world, first, last, dagger)
                flat = [j for _, ops in phases for j in ops]
                want = list(range(first, last))[::-1] if dagger else list(range(first, last))
                assert sorted(flat) == sorted(want)                      # every operator exactly once
                pos = {j: p for p, j in enumerate(flat)}
                orbs = {j: operator_orbitals(types[j], idx[j]) for j in want}
                for a_i, a in enumerate(want):                          # operators that share an orbital keep their order
                    for b in want[a_i + 1:]:
                        if orbs[a] & orbs[b]:
                            assert pos[a] < pos[b], (n, world, a, b)
                low, high = (1 << k) - 1, ((1 << k) - 1) << (n - k)
                for (name, ops), nxt in zip(phases, phases[1:] + [("end", [])]):
                    for j in ops:
                        if name == "A":
                            assert not orbs[j] & low
                        elif name == "B":
                            assert not orbs[j] & high
                        else:
                            assert orbs[j] & low and orbs[j] & high and len(ops) == 1
                    assert name == "X" or nxt[0] != name                 # maximal phases: neighbours differ
        # a 16-layer tUPS circuit needs only a handful of re-shards
    from oracle import sq_oracle as orc

    types, idx = orc.tiled_layout(16, 16)
    assert len(reshard_schedule(types, idx, 16, world)) <= 6


def test_reshard_schedule_order_gives_the_same_state():
    """Executing the phases in order (oracle, full vector) reproduces the circuit's state, forward and adjoint."""
    from oracle import sq_oracle as orc
    from slowquant_b200.distributed import reshard_schedule

    n, na, nb = 6, 3, 3
    types, idx = orc.tiled_layout(n, 3)
    rng = np.random.default_rng(5)
    th = rng.uniform(-np.pi, np.pi, len(types))
    sp = orc.get_indexing(0, n, 0, na, nb)
    st = rng.normal(size=sp.num_det)
    st /= np.linalg.norm(st)
    for world in (2, 4):
        for dagger in (False, True):
            ref = orc.construct_ups_state(st, sp, th, types, idx, dagger=dagger)
            cur = st.copy()
            for _, ops in reshard_schedule(types, idx, n, world, dagger=dagger):
                for j in ops:
                    cur = orc.construct_ups_state(cur, sp, [th[j]], [types[j]], [idx[j]], dagger=dagger)
            assert np.max(np.abs(cur - ref)) < 1e-13


def test_backwards_gradient_loop_over_adjoint_phases_equals_literal_loop():
    """The sharded energy + gradient driver walks the phases of the ADJOINT circuit from (H|psi>, |psi>): g_k = 2 <bra|T_k|ket>, then
    both vectors <- U_k^dagger.  With the oracle executing the phases on full vectors this must give the literal gradient loop
    of ups_wavefunction.py:1091-1138 (which starts from (U^dagger H|psi>, |ref>) and walks forwards) -- T_k commutes with its own
    rotation, and operators that change places inside a phase commute with each other."""
    from oracle import sq_oracle as orc
    from slowquant_b200.distributed import reshard_schedule

    n, na, nb = 6, 3, 3
    types, idx = orc.tiled_layout(n, 2)
    types = list(types) + ["single", "double", "sa_single"]
    idx = list(idx) + [(1, 9), (0, 1, 8, 11), (1, 4)]
    rng = np.random.default_rng(11)
    th = rng.uniform(-1.0, 1.0, len(types))
    th[4] = 0.0
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    sp = orc.get_indexing(0, n, 0, na, nb)
    ref_state = np.zeros(sp.num_det)
    ref_state[0] = 1.0
    want = orc.theta_gradient(ref_state, th, types, idx, h, g, sp)
    H = orc.hamiltonian_0i_0a(h, g, 0, n)
    for world in (2, 4):
        ket = orc.construct_ups_state(ref_state, sp, th, types, idx)
        bra = orc.propagate_state([H], ket, sp)
        grad = np.zeros(len(types))
        for _, ops in reshard_schedule(types, idx, n, world, 0, len(types), True):
            for k in ops:
                grad[k] = 2.0 * float(bra @ orc.get_grad_action(ket, k, sp, types, idx))
                bra = orc.construct_ups_state(bra, sp, [th[k]], [types[k]], [idx[k]], dagger=True)
                ket = orc.construct_ups_state(ket, sp, [th[k]], [types[k]], [idx[k]], dagger=True)
        assert np.max(np.abs(grad - want)) < 1e-12, world
        assert np.max(np.abs(ket - ref_state)) < 1e-12        # the ket is swept back to the reference state


@pytest.mark.parametrize("world,n,na", [(2, 6, 3), (4, 7, 3), (8, 9, 4), (8, 16, 8), (4, 6, 5)])
def test_reshard_tables_are_inverse_permutations(world, n, na):
    """A -> B moves every row to exactly one slot of the rank that owns its last-orbitals pattern, B -> A brings it back."""
    from slowquant_b200.ci_spaces import CI_Info
    from slowquant_b200.distributed import partition_prefix, reshard_tables

    info = CI_Info(0, n, 0, na, min(na, 2), device=-1)
    strs = info.strings(0)
    starts = partition_prefix(n, na, world)
    k = world.bit_length() - 1
    tabs = [reshard_tables(strs, n, world, starts, r) for r in range(world)]
    cmask = ((1 << k) - 1) << (n - k)
    lists_B = []
    for r in range(world):
        b = CI_Info(0, n, 0, na, min(na, 2), device=-1, alpha_constraint=(cmask, r << (n - k)))
        sb = b.strings(0)
        assert np.array_equal(sb, strs[(strs & cmask) == (r << (n - k))])      # constrained list = subset in list order
        lists_B.append(sb)
    # scatter the string masks themselves through the tables
    buf_B = [np.full(len(l), -1, dtype=np.int64) for l in lists_B]
    for r in range(world):
        rank_t, row_t = tabs[r][0]
        rows = strs[int(starts[r]):int(starts[r + 1])]
        assert len(rank_t) == len(rows)
        for m, dr, dl in zip(rows, rank_t, row_t):
            assert buf_B[dr][dl] == -1                                          # written once
            buf_B[dr][dl] = m
    for r in range(world):
        assert np.array_equal(buf_B[r], lists_B[r].astype(np.int64))           # every row landed where layout B expects it
    buf_A = np.full(len(strs), -1, dtype=np.int64)
    for r in range(world):
        rank_t, row_t = tabs[r][1]
        for m, dr, dl in zip(lists_B[r], rank_t, row_t):
            g = int(starts[dr]) + int(dl)
            assert buf_A[g] == -1
            buf_A[g] = m
    assert np.array_equal(buf_A, strs.astype(np.int64))


def test_constrained_space_blocks_the_right_operators():
    """Layout-B space on the host: operators that move an alpha electron on a constrained orbital are blocked, the others keep
    complete partner tables; running a blocked operator is refused."""
    from slowquant_b200 import operator_state_algebra as osa
    from slowquant_b200.ci_spaces import CI_Info
    from slowquant_b200.util import UpsStructure

    lib = _lib.load()
    n, na, nb, k = 7, 3, 4, 2
    cmask = ((1 << k) - 1) << (n - k)
    info = CI_Info(0, n, 0, na, nb, device=-1, alpha_constraint=(cmask, 1 << (n - k)))
    lay = UpsStructure()
    lay.create_tiled(n, {"n_layers": 1, "do_tups": True})
    lay.excitation_operator_type += ["single", "single", "double"]
    lay.excitation_indices += [(0, 4), (2, 12), (1, 3, 9, 11)]      # alpha 0 -> 2 (free), alpha 1 -> 6 (constrained), beta only
    lay.n_params = len(lay.excitation_operator_type)
    handle = osa.compile_layout(info, lay)
    for j, (t, idx) in enumerate(zip(lay.excitation_operator_type, lay.excitation_indices)):
        if t == "sa_single":
            want = max(idx) >= n - k
        elif t == "double" and len(idx) == 4 and idx[0] % 2 == 0 and idx[1] == idx[0] + 1:
            want = max(idx) // 2 >= n - k
        else:
            want = any(x % 2 == 0 and x // 2 >= n - k for x in idx)
        assert lib.sq_layout_op_blocked(handle, j) == int(want), (j, t, idx)
    with pytest.raises(ValueError):
        CI_Info(0, n, 0, na, nb, device=-1, alpha_constraint=(cmask, 1))     # pattern outside the mask
