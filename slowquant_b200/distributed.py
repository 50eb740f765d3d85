"""Alpha-string-sharded CI vectors over the GPUs of one node (one process per GPU).

The reference is single-process; this is the scaling axis the north star asks for: the
N_alpha x N_beta coefficient matrix is split by rows.  Rows are grouped by the occupation of the first
log2(G) orbitals, which in itertools.combinations order are contiguous ranges, so every tUPS brick on an
orbital pair (p, p+1) with p >= log2(G) is purely local.  The remaining bricks pair rows that live on two
GPUs; their tiles are rotated in place through CUDA-IPC peer mappings over NVLink (the two owners split
the columns of each pair), with a stream-ordered NCCL barrier before and after such an operator.  Scalars
(dots, norms) are reduced with ``all_reduce``.
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import numpy as np
import torch
import torch.distributed as dist

from slowquant_b200 import _lib
from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.ci_spaces import CI_Info


import os as _os

_RESHARD_DEFAULT = _os.environ.get("SQ_RESHARD", "1") != "0"   # SQ_RESHARD=0: peer-memory exchange route (A/B comparisons)
_SPINSYM_SHARDED = _os.environ.get("SQ_SPINSYM_SHARDED", "1") != "0"   # 0: always the full sigma build of sharded vectors (A/B)


def partition_prefix(n_orb: int, n_alpha: int, world: int) -> np.ndarray:
    """Row ranges of the prefix-class partition: rank r owns rows [starts[r], starts[r+1])."""
    lib = _lib.load()
    out = np.zeros(world + 1, dtype=np.int64)
    _lib.check(lib.sq_partition_prefix(n_orb, n_alpha, world, out.ctypes.data_as(C.POINTER(C.c_int64))))
    return out


def operator_orbitals(exc_type: str, indices) -> int:
    """Bit mask of the spatial orbitals an ansatz operator acts on (``sa_*``: spatial indices; otherwise spin orbitals)."""
    m = 0
    for x in indices:
        m |= 1 << (int(x) if exc_type.startswith("sa_") else int(x) // 2)
    return m


def reshard_schedule(
    types: Sequence[str], indices: Sequence[Sequence[int]], n_orb: int, world: int, first: int = 0, last: int | None = None,
    dagger: bool = False, start: str = "A",
) -> list[tuple[str, list[int]]]:
    """Phases ``(layout, [operator indices in execution order])`` of operators [first, last) on a vector sharded over
    ``world`` = 2^k ranks.  Layout "A" can run an operator iff it touches none of the first k orbitals, layout "B" iff it
    touches none of the last k; ``("X", [op])`` is an operator neither layout can run (it stays in layout A and exchanges
    tiles over peer memory).  Within a phase an operator may overtake earlier, not yet executed ones only if it shares no
    orbital with any of them (products of an even number of ladder operators on disjoint orbitals commute), so the
    concatenation of the phases is equivalent to the circuit order.  Pure function of the circuit: identical on every rank."""
    n = len(types)
    last = n if last is None else last
    k = max(world - 1, 0).bit_length()
    low = (1 << k) - 1
    high = low << (n_orb - k) if k else 0
    full = (1 << n_orb) - 1
    order = list(range(first, last))
    if dagger:
        order.reverse()
    orbs = {j: operator_orbitals(types[j], indices[j]) for j in order}
    ok = {"A": lambda j: not (orbs[j] & low), "B": lambda j: not (orbs[j] & high)}
    phases: list[tuple[str, list[int]]] = []
    cur = start
    remaining = order
    while remaining:
        sel: list[int] = []
        rest: list[int] = []
        blocked = 0
        for pos, j in enumerate(remaining):
            if blocked == full:
                rest.extend(remaining[pos:])
                break
            if ok[cur](j) and not (orbs[j] & blocked):
                sel.append(j)
            else:
                blocked |= orbs[j]
                rest.append(j)
        if sel:
            phases.append((cur, sel))
        remaining = rest
        if not remaining:
            break
        head = remaining[0]
        other = "B" if cur == "A" else "A"
        if ok[other](head):
            cur = other
        elif not ok[cur](head):   # local in neither layout
            phases.append(("X", [head]))
            remaining = remaining[1:]
            cur = "A"
        # else: the head was only blocked by a skipped operator that shares an orbital -- cannot happen for the first entry
    return phases



def reshard_tables(strings: np.ndarray, n_orb: int, world: int, row_starts: np.ndarray, rank: int):
    """Destination (rank, local row) of every local row for the re-shards A -> B and B -> A of rank ``rank``.

    ``strings``: occupation masks of ALL alpha strings in itertools.combinations order.  Layout A: rank r owns the contiguous
    rows [row_starts[r], row_starts[r+1]).  Layout B: rank r owns the strings whose last log2(world) orbitals carry the bit
    pattern r, in the order they have in the full list (``sq_space_create_constrained``)."""
    k = max(world - 1, 0).bit_length()
    strings = np.asarray(strings, dtype=np.int64)
    pat = (strings >> (n_orb - k)) & (world - 1)                      # owner in layout B
    b_local = np.zeros(len(strings), dtype=np.int64)
    for r in range(world):
        sel = pat == r
        b_local[sel] = np.arange(int(sel.sum()))
    idx = np.arange(len(strings), dtype=np.int64)
    a_owner = np.searchsorted(np.asarray(row_starts[1:], dtype=np.int64), idx, side="right")
    a_local = idx - np.asarray(row_starts, dtype=np.int64)[a_owner]
    mine_a = slice(int(row_starts[rank]), int(row_starts[rank + 1]))
    mine_b = pat == rank
    a2b = (pat[mine_a].astype(np.int32), b_local[mine_a].astype(np.int32))
    b2a = (a_owner[mine_b].astype(np.int32), a_local[mine_b].astype(np.int32))
    return a2b, b2a


class ShardedSpace:
    """CI space whose vector is sharded by alpha string over the ranks of a process group."""

    def __init__(
        self,
        num_inactive_orbs: int,
        num_active_orbs: int,
        num_virtual_orbs: int,
        num_active_elec_alpha: int,
        num_active_elec_beta: int,
        device: int | None = None,
        rank: int | None = None,
        world: int | None = None,
    ) -> None:
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.row_starts = partition_prefix(num_active_orbs, num_active_elec_alpha, self.world)
        rb, re = int(self.row_starts[self.rank]), int(self.row_starts[self.rank + 1])
        self.ci_info = CI_Info(
            num_inactive_orbs,
            num_active_orbs,
            num_virtual_orbs,
            num_active_elec_alpha,
            num_active_elec_beta,
            device=device,
            row_range=(rb, re),
        )
        lib = _lib.load()
        _lib.check(
            lib.sq_space_set_partition(
                self.ci_info._handle, self.world, self.rank, self.row_starts.ctypes.data_as(C.POINTER(C.c_int64))
            )
        )
        self.row_begin, self.row_end = rb, re
        self.local_len = self.ci_info.local_len
        self._barrier_token = None
        self._plans: dict = {}
        # ---- layout B (rows grouped by the occupation of the LAST log2(world) orbitals) and the re-shard tables ----
        self.ci_info_B: CI_Info | None = None
        self.local_len_B = 0
        self._tab_AB = self._tab_BA = None
        k = max(self.world - 1, 0).bit_length()
        n = num_active_orbs
        self.reshard_ok = self.world > 1 and (1 << k) == self.world and 2 * k <= n
        if self.reshard_ok:
            cmask = ((1 << k) - 1) << (n - k)
            self.ci_info_B = CI_Info(
                num_inactive_orbs, num_active_orbs, num_virtual_orbs, num_active_elec_alpha, num_active_elec_beta,
                device=self.ci_info.device, alpha_constraint=(cmask, self.rank << (n - k)),
            )
            self.local_len_B = self.ci_info_B.num_alpha_strings * self.ci_info_B.num_beta_strings
            a2b, b2a = reshard_tables(self.ci_info.strings(0), n, self.world, self.row_starts, self.rank)
            self.reshard_tables_host = (a2b, b2a)
            if self.ci_info.device >= 0:
                dev = torch.device("cuda", self.ci_info.device)
                self._tab_AB = tuple(torch.from_numpy(t).to(dev) for t in a2b)
                self._tab_BA = tuple(torch.from_numpy(t).to(dev) for t in b2a)
            assert len(b2a[0]) == self.ci_info_B.num_alpha_strings

    # ---- shards -------------------------------------------------------------------------------
    def alloc_state(self, zero: bool = True) -> "ShardedState":
        return ShardedState(self, zero)

    def barrier(self) -> None:
        """Stream-ordered device-wide barrier (tiny NCCL all-reduce on the current stream)."""
        if self.world == 1:
            return
        if self._barrier_token is None:
            self._barrier_token = torch.zeros(1, dtype=torch.float32, device=torch.device("cuda", self.ci_info.device))
        dist.all_reduce(self._barrier_token)

    def exchange_plan(self, wf_struct, first: int, last: int, dagger: bool) -> list[tuple[int, int, bool]]:
        """Split operators [first,last) into maximal ranges (f, l, needs_exchange) in execution order."""
        lib = _lib.load()
        lay = osa.compile_layout(self.ci_info, wf_struct)
        flags = [bool(lib.sq_layout_needs_exchange(lay, k, k + 1)) for k in range(first, last)]
        types, idx = wf_struct.excitation_operator_type, wf_struct.excitation_indices
        pair_key = []
        for k in range(first, last):
            t, ind = types[k], idx[k]
            if t == "sa_single":
                pair_key.append((int(ind[0]), int(ind[1])))
            elif t == "double" and len(ind) == 4 and ind[0] % 2 == 0 and ind[1] == ind[0] + 1 and ind[2] % 2 == 0 and ind[3] == ind[2] + 1:
                pair_key.append((int(ind[0]) // 2, int(ind[2]) // 2))
            else:
                pair_key.append(None)
        order = list(range(last - first))
        if dagger:
            order.reverse()
        plan: list[tuple[int, int, bool]] = []
        cur: list[int] = []
        for j in order:
            # operators on the same orbital pair stay together (they fuse into one launch);
            # an exchange operator never shares a range with a different pair
            if cur and (flags[j] != flags[cur[-1]] or (flags[j] and pair_key[j] != pair_key[cur[-1]])):
                plan.append((first + min(cur), first + max(cur) + 1, flags[cur[0]]))
                cur = []
            cur.append(j)
        if cur:
            plan.append((first + min(cur), first + max(cur) + 1, flags[cur[0]]))
        return plan


class ShardedState:
    """One rank's shard of a CI vector plus the peer mappings of all other shards."""

    def __init__(self, space: ShardedSpace, zero: bool = True) -> None:
        lib = _lib.load()
        self.space = space
        dev = space.ci_info.device
        ptr = C.c_void_p()
        _lib.check(lib.sq_dist_alloc(dev, space.local_len, C.byref(ptr)))
        self._ptr = ptr
        n = space.local_len

        class _Iface:
            __cuda_array_interface__ = {"shape": (max(n, 1),), "typestr": "<f8", "data": (ptr.value, False), "version": 2}

        self._iface = _Iface()
        self.local = torch.as_tensor(self._iface, device=torch.device("cuda", dev))[:n]
        if zero:
            self.local.zero_()
        self._opened: list[C.c_void_p] = []
        self._peer_ptrs = self._share(ptr)
        # layout-B buffer of the re-sharding driver: allocated (collectively) the first time a circuit needs it
        self._ptr_B = None
        self._peer_ptrs_B = None
        self.local_B = None

    def _share(self, ptr: C.c_void_p):
        """Peer pointers of the buffers `ptr` of all ranks (collective: CUDA-IPC handles are all-gathered)."""
        lib = _lib.load()
        space = self.space
        dev = space.ci_info.device
        peers = (C.c_void_p * space.world)()
        peers[space.rank] = ptr.value
        if space.world > 1:
            handle = C.create_string_buffer(64)
            _lib.check(lib.sq_ipc_export(ptr, handle))
            gathered: list = [None] * space.world
            dist.all_gather_object(gathered, bytes(handle.raw))
            for r in range(space.world):
                if r == space.rank:
                    continue
                p = C.c_void_p()
                _lib.check(lib.sq_ipc_import(dev, gathered[r], C.byref(p)))
                peers[r] = p.value
                self._opened.append(p)
            torch.cuda.synchronize()
            dist.barrier()
        return peers

    def ensure_layout_B(self) -> None:
        """Collective: allocate and peer-map the layout-B buffer (rows grouped by the last log2(world) orbitals)."""
        if self._ptr_B is not None:
            return
        lib = _lib.load()
        sp = self.space
        dev = sp.ci_info.device
        ptr = C.c_void_p()
        _lib.check(lib.sq_dist_alloc(dev, sp.local_len_B, C.byref(ptr)))
        self._ptr_B = ptr
        n = sp.local_len_B

        class _Iface:
            __cuda_array_interface__ = {"shape": (max(n, 1),), "typestr": "<f8", "data": (ptr.value, False), "version": 2}

        self._iface_B = _Iface()
        self.local_B = torch.as_tensor(self._iface_B, device=torch.device("cuda", dev))[:n]
        self._peer_ptrs_B = self._share(ptr)

    def close(self) -> None:
        lib = _lib.load()
        if self.space.world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        for p in self._opened:
            lib.sq_ipc_close(p)
        self._opened = []
        if self._ptr_B:
            self.local_B = None
            lib.sq_dist_free(self._ptr_B)
            self._ptr_B = None
        if self._ptr:
            self.local = None
            lib.sq_dist_free(self._ptr)
            self._ptr = None

    # convenience: fill the local shard from a full-length host vector / gather the full vector
    def set_from_full(self, full: np.ndarray) -> None:
        sp = self.space
        nb = sp.ci_info.num_beta_strings
        self.local.copy_(torch.from_numpy(np.ascontiguousarray(full[sp.row_begin * nb : sp.row_end * nb])))

    def set_determinant(self, index: int) -> None:
        sp = self.space
        nb = sp.ci_info.num_beta_strings
        self.local.zero_()
        lo, hi = sp.row_begin * nb, sp.row_end * nb
        if lo <= index < hi:
            self.local[index - lo] = 1.0


def _reshard(state: ShardedState, to_B: bool) -> None:
    """All-to-all between the two row layouts: every rank writes its rows into the new owners' buffers (``sq_reshard_rows``),
    then a device-wide barrier (all rows have landed; nobody reads the old buffer any more when it is written next time)."""
    lib = _lib.load()
    sp = state.space
    nb = sp.ci_info.num_beta_strings
    if to_B:
        tab, src, n_rows, dst = sp._tab_AB, state.local, sp.row_end - sp.row_begin, state._peer_ptrs_B
    else:
        tab, src, n_rows, dst = sp._tab_BA, state.local_B, sp.ci_info_B.num_alpha_strings, state._peer_ptrs
    _lib.check(
        lib.sq_reshard_rows(sp.ci_info.device, n_rows, nb, osa._ptr(src), osa._ptr(tab[0]), osa._ptr(tab[1]), dst, sp.world,
                            osa._stream())
    )
    sp.barrier()


def construct_ups_state_sharded(
    state: ShardedState, thetas: Sequence[float], ups_struct, dagger: bool = False, first: int = 0, last: int | None = None,
    reshard: bool | None = None,
) -> None:
    """In place: state <- U state (or U^dagger state) on the sharded vector (osa.py:963-1412 semantics).

    Default route (``reshard=True`` wherever the partition allows it): the phases of :func:`reshard_schedule`, each one a fused
    launch sequence on the local shard in the row layout that makes its operators local, with one all-to-all re-shard between
    phases; the vector stays in the window kernel's sign-free gauge from the first window sweep to the end of the call.
    ``reshard=False``: the vector stays in layout A and operators whose row pairs span two GPUs rotate their tiles in place
    over peer memory, one barrier-fenced operator range at a time."""
    lib = _lib.load()
    sp = state.space
    n_ops = len(ups_struct.excitation_operator_type)
    last = n_ops if last is None else last
    lay = osa.compile_layout(sp.ci_info, ups_struct)
    th = osa._thetas_array(thetas, n_ops)
    thp = th.ctypes.data_as(C.POINTER(C.c_double))
    if reshard is None:
        reshard = sp.reshard_ok and _RESHARD_DEFAULT
    if reshard and sp.world > 1 and sp.reshard_ok:
        key = ("reshard", id(ups_struct), n_ops, first, last, bool(dagger))
        phases = sp._plans.get(key)
        if phases is None:
            phases = reshard_schedule(
                ups_struct.excitation_operator_type, ups_struct.excitation_indices, sp.ci_info.num_active_orbs, sp.world, first, last,
                dagger,
            )
            phases = [(name, np.asarray(ops, dtype=np.int32)) for name, ops in phases]
            sp._plans[key] = phases
        if any(name == "B" for name, _ in phases):
            state.ensure_layout_B()
            lay_B = osa.compile_layout(sp.ci_info_B, ups_struct)
        PI = C.POINTER(C.c_int32)
        where, gauge = "A", False

        def leave_gauge():
            nonlocal gauge
            if gauge:
                info, handle, buf, nloc = (
                    (sp.ci_info, lay, state.local, sp.local_len) if where == "A" else (sp.ci_info_B, lay_B, state.local_B, sp.local_len_B)
                )
                if nloc:
                    _lib.check(lib.sq_ups_apply_list(info._handle, handle, thp, 0, None, 0, 1, osa._ptr(buf), osa._stream()))
                gauge = False

        for name, ops in phases:
            if name == "X":
                # local in neither layout: cross-device tiles are rotated in place over peer memory (layout A, reference gauge)
                leave_gauge()
                if where == "B":
                    _reshard(state, to_B=False)
                    where = "A"
                k = int(ops[0])
                sp.barrier()
                _lib.check(lib.sq_ups_apply_dist(sp.ci_info._handle, lay, thp, k, k + 1, 1 if dagger else 0, state._peer_ptrs, osa._stream()))
                sp.barrier()
                continue
            if name != where:
                _reshard(state, to_B=(name == "B"))
                where = name
            info, handle, buf, nloc = (
                (sp.ci_info, lay, state.local, sp.local_len) if where == "A" else (sp.ci_info_B, lay_B, state.local_B, sp.local_len_B)
            )
            # every phase ends in the sign-free gauge on EVERY rank (the ranks' launch plans may differ, the gauge of the rows
            # they exchange must not); it is left once, at the end of the call
            if nloc:
                _lib.check(
                    lib.sq_ups_apply_list(info._handle, handle, thp, len(ops), ops.ctypes.data_as(PI), 1 if dagger else 0,
                                          (1 if gauge else 0) | 2, osa._ptr(buf), osa._stream())
                )
            gauge = True
        leave_gauge()
        if where == "B":
            _reshard(state, to_B=False)
        return
    key = (id(ups_struct), n_ops, first, last, bool(dagger))
    plan = sp._plans.get(key)
    if plan is None:
        plan = sp.exchange_plan(ups_struct, first, last, dagger)
        sp._plans[key] = plan
    # a device-wide barrier separates an exchange range from its neighbours on both sides: before it every
    # rank must have finished writing its own rows, after it the remote writes must have landed
    prev_exchange = False
    for f, l, exchange in plan:
        if exchange or prev_exchange:
            sp.barrier()
        _lib.check(
            lib.sq_ups_apply_dist(sp.ci_info._handle, lay, thp, f, l, 1 if dagger else 0, state._peer_ptrs, osa._stream())
        )
        prev_exchange = exchange
    if prev_exchange:
        sp.barrier()


def dot_sharded(a: ShardedState, b: ShardedState) -> float:
    """<a|b> over all shards."""
    sp = a.space
    local = osa._dot(a.local, b.local, sp.ci_info) if sp.local_len else 0.0
    if sp.world == 1:
        return local
    t = torch.tensor([local], dtype=torch.float64, device=a.local.device)
    dist.all_reduce(t)
    return float(t.item())


def _measure_spin_flip_symmetry(state: "ShardedState") -> float:
    """lambda = +-1 if the sharded vector satisfies c[B,A] = lambda (-1)^popc(A & B) c[A,B] to 1e-12 of its largest amplitude (measured
    on every rank's rows against the mirrored elements in the owners' shards, MAX over the ranks), else 0.  Every shard must be
    complete (barrier) before the call."""
    lib = _lib.load()
    sp = state.space
    res = np.zeros(3)
    _lib.check(lib.sq_spinsym_measure_dist(sp.ci_info._handle, state._peer_ptrs, res.ctypes.data_as(C.POINTER(C.c_double)), osa._stream()))
    if sp.world > 1:
        t = torch.from_numpy(res).to(state.local.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res = t.cpu().numpy()
    if res[0] > 0.0 and res[1] <= 1e-12 * res[0]:
        return 1.0
    if res[0] > 0.0 and res[2] <= 1e-12 * res[0]:
        return -1.0
    return 0.0


def rdm12_sharded(bra: ShardedState, ket: ShardedState, want_rdm2: bool = True) -> tuple[np.ndarray, np.ndarray | None]:
    """Active-space (transition) 1-/2-RDMs of alpha-sharded vectors: every rank contracts the determinants of its own rows
    (alpha partners that live on another GPU are read in place through the peer mappings, ``sq_rdm12_dist``), the n^2 + n^4
    partial sums are added with one all-reduce."""
    lib = _lib.load()
    sp = ket.space
    n = sp.ci_info.num_active_orbs
    d1 = np.zeros((n, n), dtype=np.float64)
    d2 = np.zeros((n, n, n, n), dtype=np.float64) if want_rdm2 else None
    PD = C.POINTER(C.c_double)
    torch.cuda.synchronize()
    sp.barrier()        # every shard is complete before anybody reads it remotely
    lam = 0.0
    info = sp.ci_info
    if _SPINSYM_SHARDED and bra is ket and want_rdm2 and info.num_active_elec_alpha == info.num_active_elec_beta:
        lam = _measure_spin_flip_symmetry(ket)      # +-1: the panels hold the kept half of every rank's rows only (weighted)
    _lib.check(
        lib.sq_rdm12_dist_sym(
            sp.ci_info._handle, bra._peer_ptrs, ket._peer_ptrs, lam, d1.ctypes.data_as(PD),
            d2.ctypes.data_as(PD) if want_rdm2 else None, osa._stream(),
        )
    )
    sp.barrier()        # ... and nobody changes a shard while a neighbour may still be reading it
    if sp.world > 1:
        dev = ket.local.device
        parts = [d1] if d2 is None else [d1, d2]
        flat = torch.from_numpy(np.concatenate([p.ravel() for p in parts])).to(dev)
        dist.all_reduce(flat)
        flat = flat.cpu().numpy()
        d1 = flat[: n * n].reshape(n, n)
        if d2 is not None:
            d2 = flat[n * n :].reshape(n, n, n, n)
    return d1, d2


def energy_sharded(state: ShardedState, h_act: np.ndarray, g_act: np.ndarray, e_core: float = 0.0) -> float:
    r""":math:`\langle\Psi|H|\Psi\rangle = E_\text{core} + \sum h_{pq}\Gamma^1_{pq} + \tfrac12\sum g_{pqrs}\Gamma^2_{pqrs}` of a sharded vector
    from its RDMs (the RDM route of ups_wavefunction.py:1041-1050 / density_matrix.py:139-178 in the active space)."""
    d1, d2 = rdm12_sharded(state, state)
    return float(e_core + np.sum(np.asarray(h_act) * d1) + 0.5 * np.sum(np.asarray(g_act) * d2))


def sigma_sharded(state: ShardedState, h_act: np.ndarray, g_act: np.ndarray, e_core: float = 0.0, out: ShardedState | None = None) -> ShardedState:
    r""":math:`H|\Psi\rangle` of an alpha-sharded vector as a new sharded vector (the string route of
    ups_wavefunction.py:770-784 / :1091-1112 in the active space; ``h_act`` / ``g_act`` are the folded active integrals of
    ``operators.fold_hamiltonian_0i_0a``).  Every rank treats the determinants of its rows as sources (``sq_sigma_dist``):
    alpha partners on other GPUs are read over NVLink, images in other GPUs' rows are added there with system-scope
    atomics; no transpose, no collective besides the two barriers.  ``sq_sigma_dist`` refuses device pairs without native
    peer-to-peer atomics (not NVLink-connected).  Parity: tests/test_gpu_distributed.py::test_sharded_sigma_matches_single_gpu."""
    lib = _lib.load()
    sp = state.space
    n = sp.ci_info.num_active_orbs
    h = np.ascontiguousarray(h_act, dtype=np.float64).reshape(n, n)
    g = np.ascontiguousarray(g_act, dtype=np.float64).reshape(n, n, n, n)
    if out is None:
        out = sp.alloc_state(zero=False)
    if out is state:
        raise ValueError("sigma_sharded: out must not be the input state")
    PD = C.POINTER(C.c_double)
    # Spin-flip symmetric input (c[B,A] = lambda (-1)^popc(A & B) c[A,B]: every tUPS state on a closed-shell reference) needs half
    # of the build: MEASURED here (max over the ranks), never assumed; anything else takes the full build.
    lam = 0.0
    info = sp.ci_info
    real_orbital = bool(np.allclose(g, g.transpose(1, 0, 2, 3), rtol=0.0, atol=1e-14 * max(1.0, float(np.max(np.abs(g))))) and
                        np.allclose(g, g.transpose(0, 1, 3, 2), rtol=0.0, atol=1e-14 * max(1.0, float(np.max(np.abs(g))))) and
                        np.allclose(h, h.T, rtol=0.0, atol=1e-14 * max(1.0, float(np.max(np.abs(h))))))
    if _SPINSYM_SHARDED and real_orbital and info.num_active_elec_alpha == info.num_active_elec_beta:
        torch.cuda.synchronize()
        sp.barrier()    # every in shard is complete before the remote reads of the measurement
        lam = _measure_spin_flip_symmetry(state)
    if lam != 0.0:
        # S' = sum over the kept half of the sources (all targets), then sigma = S' + lambda U S' in place (doubles the e_core term)
        torch.mul(state.local, 0.5 * float(e_core), out=out.local)
        torch.cuda.synchronize()
        sp.barrier()
        used = C.c_int(0)
        _lib.check(lib.sq_sigma_dist_sym(info._handle, h.ctypes.data_as(PD), g.ctypes.data_as(PD), state._peer_ptrs, out._peer_ptrs,
                                         lam, C.byref(used), osa._stream()))
        torch.cuda.synchronize()
        sp.barrier()        # all contributions of the kept sources have landed on every rank
        _lib.check(lib.sq_spinsym_mirror_dist(info._handle, out._peer_ptrs, lam, osa._stream()))
        torch.cuda.synchronize()
        sp.barrier()        # symmetrised
        return out
    torch.mul(state.local, float(e_core), out=out.local)
    torch.cuda.synchronize()
    sp.barrier()        # every out shard is initialised and every in shard complete before remote reads / atomics start
    _lib.check(lib.sq_sigma_dist(sp.ci_info._handle, h.ctypes.data_as(PD), g.ctypes.data_as(PD), state._peer_ptrs, out._peer_ptrs, osa._stream()))
    torch.cuda.synchronize()
    sp.barrier()        # all remote contributions have landed
    return out


def energy_sharded_sigma(state: ShardedState, h_act: np.ndarray, g_act: np.ndarray, e_core: float = 0.0) -> float:
    r""":math:`\langle\Psi|H|\Psi\rangle` of a sharded vector through :func:`sigma_sharded` and one all-reduced dot."""
    sig = sigma_sharded(state, h_act, g_act, e_core)
    try:
        return dot_sharded(state, sig)
    finally:
        sig.close()


# ---- theta gradient of a sharded vector ---------------------------------------------------------------------------------
# The single-GPU path fuses g_k = 2 <bra|T_k|ket> with the two rotations in one kernel (sq_ups_grad_sweep).  For vectors that
# only exist sharded, the same reference loop (ups_wavefunction.py:1114-1138) is composed from the sharded primitives:
# <bra|T_k|ket> is the derivative at x = 0 of f(x) = <bra|exp(x T_k)|ket>, a trigonometric polynomial whose frequencies are
# the (integer) |eigenvalues| of the anti-Hermitian generator, so it is EXACTLY a weighted sum of 2R rotated overlaps
# (general parameter-shift rule for equidistant frequencies).  Correct first, not fast: 2R single-operator sweeps + dots per
# parameter; a peer-memory variant of the fused gradient kernel is the next step.
_AMPLITUDE_FREQUENCIES = {
    # largest |eigenvalue| of T_k (spectra verified against the reference's generators)
    "sa_single": 2,        # T_alpha + T_beta, commuting, each with spectrum {0, +-i}  (osa.py:1003-1042)
    "single": 1, "double": 1, "triple": 1, "quadruple": 1, "quintuple": 1, "sextuple": 1,   # G - G^dagger: {0, +-i}
    "sa_double_1": 1,
}


def shift_rule(R: int) -> tuple[np.ndarray, np.ndarray]:
    """Points x_mu and weights w_mu with f'(0) = sum_mu w_mu f(x_mu) for every f(x) = sum_{|l| <= R} c_l exp(i l x)."""
    mu = np.arange(1, 2 * R + 1)
    x = (2 * mu - 1) * np.pi / (2 * R)
    w = (-1.0) ** (mu - 1) / (4 * R * np.sin(x / 2) ** 2)
    return x, w


def gradient_segments(
    plan: list[tuple[int, int, bool]], ups_struct, fused_local: bool = True, peer_gradient: bool = False
) -> list[tuple[int, int, str]]:
    """Split the exchange plan of a circuit into (first, last, mode) stretches for the sharded theta gradient.  Modes:
    ``"fused"`` -- every row pair is local and every operator is a brick operator (sa_single, or a pair double
    (2p, 2p+1, 2q, 2q+1)): the fused single-GPU tile gradient kernels run on the shards; ``"peer"`` -- an exchange stretch of brick
    operators, fused kernel on peer memory (``sq_ups_grad_sweep_dist``; only with ``peer_gradient=True``); ``"shift"`` -- everything
    else, differentiated with the shift rule one operator at a time."""
    types, idx = ups_struct.excitation_operator_type, ups_struct.excitation_indices

    def brick_operator(k: int) -> bool:
        t, ind = types[k], idx[k]
        if t == "sa_single":
            return True
        return t == "double" and len(ind) == 4 and ind[0] % 2 == 0 and ind[1] == ind[0] + 1 and ind[2] % 2 == 0 and ind[3] == ind[2] + 1

    out: list[tuple[int, int, str]] = []
    for f, l, exchange in plan:
        bricks = all(brick_operator(k) for k in range(f, l))
        if bricks and not exchange and fused_local:
            mode = "fused"
        elif bricks and exchange and peer_gradient:
            mode = "peer"
        else:
            mode = "shift"
        out.append((f, l, mode))
    return out


_BACKWARD_SWEEP = _os.environ.get("SQ_BACKWARD_SWEEP", "1") != "0"   # sharded theta gradient without the adjoint pass (A/B switch)


def _energy_and_theta_gradient_resharded(
    sp: "ShardedSpace", reference, th: np.ndarray, ups_struct, h_act: np.ndarray, g_act: np.ndarray, e_core: float,
    timings: dict | None = None,
) -> tuple[float, np.ndarray]:
    """Energy and theta gradient with every circuit traversal (U, U^dagger, the gradient loop of ups_wavefunction.py:1114-1138) as
    local phases between re-shards."""
    lib = _lib.load()
    types, indices = ups_struct.excitation_operator_type, ups_struct.excitation_indices
    P = len(types)
    thp = th.ctypes.data_as(C.POINTER(C.c_double))
    PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    import time as _time

    def _mark(name, t0):
        if timings is not None:
            torch.cuda.synchronize()
            timings[name] = timings.get(name, 0.0) + _time.perf_counter() - t0
        return _time.perf_counter()

    t0 = _time.perf_counter()
    ket = sp.alloc_state(zero=False)
    _load_reference(ket, reference)
    construct_ups_state_sharded(ket, th, ups_struct, reshard=True)
    t0 = _mark("state_s", t0)
    bra = sigma_sharded(ket, h_act, g_act, e_core)
    grad = np.zeros(P)
    try:
        energy = dot_sharded(ket, bra)
        t0 = _mark("sigma_s", t0)
        # The gradient loop runs BACKWARDS through the circuit from (H|psi>, |psi>) -- the phases of the adjoint circuit, every
        # operator differentiated and then undone on both vectors (sq_ups_grad_sweep_list_rev): the numbers of
        # ups_wavefunction.py:1114-1138 without the adjoint pass U^dagger H|psi> (one circuit traversal less; 6.4 of 68.5 s at
        # CAS(20,20) on 8 GPUs).  A circuit with an operator that is local in neither layout keeps the forward route.
        phases = reshard_schedule(types, indices, sp.ci_info.num_active_orbs, sp.world, 0, P, True)
        backwards = _BACKWARD_SWEEP and not any(name == "X" for name, _ in phases)
        sweep = lib.sq_ups_grad_sweep_list_rev if backwards else lib.sq_ups_grad_sweep_list
        if not backwards:
            construct_ups_state_sharded(bra, th, ups_struct, dagger=True, reshard=True)
            _load_reference(ket, reference)
            t0 = _mark("adjoint_s", t0)
            phases = reshard_schedule(types, indices, sp.ci_info.num_active_orbs, sp.world)
        lay = osa.compile_layout(sp.ci_info, ups_struct)
        lay_B = None
        if any(name == "B" for name, _ in phases):
            bra.ensure_layout_B()
            ket.ensure_layout_B()
            lay_B = osa.compile_layout(sp.ci_info_B, ups_struct)
        where = "A"
        for name, ops in phases:
            arr = np.asarray(ops, dtype=np.int32)
            if name == "X":
                # local in neither layout: one operator through the fused gradient kernel on peer memory (layout A)
                if where == "B":
                    _reshard(bra, to_B=False)
                    _reshard(ket, to_B=False)
                    where = "A"
                k = int(arr[0])
                part = np.zeros(1)
                torch.cuda.synchronize()
                sp.barrier()
                _lib.check(lib.sq_ups_grad_sweep_dist(sp.ci_info._handle, lay, thp, k, k + 1, bra._peer_ptrs, ket._peer_ptrs,
                                                      part.ctypes.data_as(PD), osa._stream()))
                torch.cuda.synchronize()
                sp.barrier()
                grad[k] = part[0]
                continue
            if name != where:
                _reshard(bra, to_B=(name == "B"))
                _reshard(ket, to_B=(name == "B"))
                where = name
            info, handle, b_buf, k_buf, nloc = (
                (sp.ci_info, lay, bra.local, ket.local, sp.local_len) if where == "A"
                else (sp.ci_info_B, lay_B, bra.local_B, ket.local_B, sp.local_len_B)
            )
            if nloc:
                part = np.zeros(len(arr))
                _lib.check(sweep(info._handle, handle, thp, len(arr), arr.ctypes.data_as(PI), osa._ptr(b_buf),
                                 osa._ptr(k_buf), part.ctypes.data_as(PD), osa._stream()))
                grad[arr] = part
        if sp.world > 1:   # every rank holds the partial sums over its rows
            t = torch.from_numpy(grad).to(ket.local.device)
            dist.all_reduce(t)
            grad = t.cpu().numpy()
        _mark("gradient_sweep_s", t0)
    finally:
        bra.close()
        ket.close()
    return energy, grad


def _load_reference(dst: "ShardedState", reference) -> None:
    """dst <- reference: a sharded vector, or the index of a determinant (e.g. 0 = Hartree-Fock; saves one vector of memory)."""
    if isinstance(reference, ShardedState):
        dst.local.copy_(reference.local)
    else:
        dst.set_determinant(int(reference))


def energy_and_theta_gradient_sharded(
    reference, thetas: Sequence[float], ups_struct, h_act: np.ndarray, g_act: np.ndarray, e_core: float = 0.0,
    fused_local: bool = True,
    peer_gradient: bool = False,
    reshard: bool | None = None,
    space: "ShardedSpace | None" = None,
    timings: dict | None = None,
) -> tuple[float, np.ndarray]:
    r"""Energy and :math:`\partial E/\partial\theta_k` of :math:`U(\theta)|\text{reference}\rangle` on an alpha-sharded
    vector (the theta part of ``_calc_gradient_optimization``, ups_wavefunction.py:1091-1138).  ``reference`` is not modified; it may
    also be the index of a determinant together with ``space`` (no reference vector is held: CAS(20,20) on 8 GPUs needs it).

    Default route (``reshard``): state construction, adjoint and the gradient loop all run as the local phases of
    :func:`reshard_schedule` -- bra and ket are re-sharded together between phases and every phase is one fused
    ``sq_ups_grad_sweep_list`` on the shards (commuting operators may change places, which leaves every <bra|T_k|ket> unchanged);
    ONE all-reduce of the gradient at the end.  ``reshard=False`` keeps both vectors in layout A:

    Stretches of bricks whose row pairs are all local (on G GPUs: every pair (p, p+1) with p >= log2 G) go through the fused
    single-GPU gradient kernels on the shards plus one all-reduce per stretch; operators that pair rows of two GPUs use the
    shift rule (``fused_local=False``: shift rule everywhere), or, with ``peer_gradient=True``, the fused gradient kernel on peer
    memory (``sq_ups_grad_sweep_dist``).  All three routes are compared with the fused single-GPU call in
    tests/dist_sigma_worker.py; the shift-rule arithmetic is also checked on the CPU against the oracle's literal gradient
    loop (tests/test_distributed_host.py)."""
    sp = reference.space if isinstance(reference, ShardedState) else space
    if sp is None:
        raise ValueError("a determinant index as reference needs space=")
    types = list(ups_struct.excitation_operator_type)
    P = len(types)
    th = osa._thetas_array(thetas, P)
    if reshard is None:
        reshard = sp.reshard_ok and _RESHARD_DEFAULT and fused_local and not peer_gradient
    if reshard and sp.world > 1 and sp.reshard_ok:
        return _energy_and_theta_gradient_resharded(sp, reference, th, ups_struct, h_act, g_act, e_core, timings)
    for t in types:
        if t not in _AMPLITUDE_FREQUENCIES:
            raise NotImplementedError(f"theta gradient of sharded vectors: no shift rule for operator type {t}")
    ket = sp.alloc_state(zero=False)
    _load_reference(ket, reference)
    construct_ups_state_sharded(ket, th, ups_struct)
    bra = sigma_sharded(ket, h_act, g_act, e_core)
    energy = dot_sharded(ket, bra)
    construct_ups_state_sharded(bra, th, ups_struct, dagger=True)
    _load_reference(ket, reference)
    tmp = sp.alloc_state(zero=False)
    grad = np.zeros(P)
    probe = np.zeros(P)
    lib = _lib.load()
    lay = osa.compile_layout(sp.ci_info, ups_struct)
    PD = C.POINTER(C.c_double)

    try:
        for f, l, mode in gradient_segments(sp.exchange_plan(ups_struct, 0, P, False), ups_struct, fused_local, peer_gradient):
            if mode == "peer":
                # exchange bricks through the fused gradient kernel on peer memory (both owners of a cross-device row pair take
                # half of its columns); barrier before (all rows complete) and after (remote writes have landed)
                part = np.zeros(l - f, dtype=np.float64)
                torch.cuda.synchronize()
                sp.barrier()
                _lib.check(lib.sq_ups_grad_sweep_dist(sp.ci_info._handle, lay, th.ctypes.data_as(PD), f, l, bra._peer_ptrs,
                                                      ket._peer_ptrs, part.ctypes.data_as(PD), osa._stream()))
                torch.cuda.synchronize()
                sp.barrier()
                if sp.world > 1:
                    t = torch.from_numpy(part).to(ket.local.device)
                    dist.all_reduce(t)
                    part = t.cpu().numpy()
                grad[f:l] = part
                continue
            if mode == "fused":
                # every row pair of these operators is local: the fused single-GPU sweep (g_k and both rotations in one kernel
                # per brick, sq_ups_grad_sweep) runs on the shards; <bra|T_k|ket> is a sum over rows -> one all-reduce
                part = np.zeros(l - f, dtype=np.float64)
                if sp.local_len:
                    _lib.check(lib.sq_ups_grad_sweep(sp.ci_info._handle, lay, th.ctypes.data_as(PD), f, l, osa._ptr(bra.local),
                                                     osa._ptr(ket.local), part.ctypes.data_as(PD), osa._stream()))
                if sp.world > 1:
                    t = torch.from_numpy(part).to(ket.local.device)
                    dist.all_reduce(t)
                    part = t.cpu().numpy()
                grad[f:l] = part
                continue
            for k in range(f, l):   # operators that pair rows of two GPUs (or that the tile kernels do not take): shift rule
                xs, ws = shift_rule(_AMPLITUDE_FREQUENCIES[types[k]])
                acc = 0.0
                for x, w in zip(xs, ws):
                    tmp.local.copy_(ket.local)
                    probe[k] = x
                    construct_ups_state_sharded(tmp, probe, ups_struct, first=k, last=k + 1)
                    acc += w * dot_sharded(bra, tmp)
                probe[k] = 0.0
                grad[k] = 2.0 * acc
                construct_ups_state_sharded(bra, th, ups_struct, first=k, last=k + 1)
                construct_ups_state_sharded(ket, th, ups_struct, first=k, last=k + 1)
    finally:
        tmp.close()
        bra.close()
        ket.close()
    return energy, grad
