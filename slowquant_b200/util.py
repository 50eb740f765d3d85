"""Ansatz layouts: ordered lists (excitation type, indices) that the engine compiles into kernel launches.

Same classes, method names, option names and ordering as the reference's
slowquant/unitary_coupled_cluster/util.py (iterators :5-544, UccStructure :547-639, UpsStructure :642-1073);
the orderings are restated with itertools instead of nested index loops.
"""
from __future__ import annotations

import itertools
from collections.abc import Sequence
from typing import Any


def _spin_balanced(creators: Sequence[int], annihilators: Sequence[int]) -> bool:
    """Excitation keeps N_alpha and N_beta (even spin-orbital index = alpha)."""
    return sum(1 for x in creators if x % 2 == 0) == sum(1 for x in annihilators if x % 2 == 0)


# ---- spin-adapted iterators (util.py:5-110) -------------------------------------------------------
def iterate_t1_sa(active_occ_idx: Sequence[int], active_unocc_idx: Sequence[int]):
    for i in active_occ_idx:
        for a in active_unocc_idx:
            yield a, i, 2 ** (-1 / 2)


def _sa_double_cases(i: int, j: int, a: int, b: int):
    fac = 1.0
    if a == b:
        fac *= 2.0
    if i == j:
        fac *= 2.0
    fac = 1 / 2 * (fac) ** (-1 / 2)
    if i == j and a == b:
        yield a, i, b, j, fac, 1
    elif i == j:
        yield a, i, b, j, fac, 2
    elif a == b:
        yield a, i, b, j, fac, 3
    else:
        yield a, i, b, j, fac, 4
        yield a, i, b, j, 1 / (2 * 3 ** (1 / 2)), 5


def iterate_t2_sa(active_occ_idx: Sequence[int], active_unocc_idx: Sequence[int]):
    for i, j in itertools.combinations_with_replacement(active_occ_idx, 2):
        for a, b in itertools.combinations_with_replacement(active_unocc_idx, 2):
            yield from _sa_double_cases(i, j, a, b)


def iterate_t1_sa_generalized(num_orbs: int):
    for i, a in itertools.combinations(range(num_orbs), 2):
        yield a, i, 2 ** (-1 / 2)


def iterate_t2_sa_generalized(num_orbs: int):
    for i in range(num_orbs):
        for j in range(i, num_orbs):
            for a in range(max(i, j) + 1, num_orbs):
                for b in range(a, num_orbs):
                    yield from _sa_double_cases(i, j, a, b)


# ---- spin-orbital iterators (util.py:113-544) ---------------------------------------------------------
def _iterate_tn(n: int, occ: Sequence[int], unocc: Sequence[int]):
    """All spin-conserving n-fold excitations; virtual tuples vary slowest (util.py:171-510)."""
    for virt in itertools.combinations(unocc, n):
        for holes in itertools.combinations(occ, n):
            if _spin_balanced(virt, holes):
                yield tuple(x for pair in zip(virt, holes) for x in pair)


def iterate_t1(active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]):
    yield from _iterate_tn(1, active_occ_spin_idx, active_unocc_spin_idx)


def iterate_t2(active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]):
    yield from _iterate_tn(2, active_occ_spin_idx, active_unocc_spin_idx)


def iterate_t3(active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]):
    yield from _iterate_tn(3, active_occ_spin_idx, active_unocc_spin_idx)


def iterate_t4(active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]):
    yield from _iterate_tn(4, active_occ_spin_idx, active_unocc_spin_idx)


def iterate_t5(active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]):
    yield from _iterate_tn(5, active_occ_spin_idx, active_unocc_spin_idx)


def iterate_t6(active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]):
    yield from _iterate_tn(6, active_occ_spin_idx, active_unocc_spin_idx)


def iterate_t1_generalized(num_spin_orbs: int):
    for i, a in itertools.combinations(range(num_spin_orbs), 2):
        if _spin_balanced((a,), (i,)):
            yield a, i


def iterate_t2_generalized(num_spin_orbs: int):
    for i in range(num_spin_orbs):
        for j in range(i, num_spin_orbs):
            for a in range(max(i, j) + 1, num_spin_orbs):
                for b in range(a, num_spin_orbs):
                    if _spin_balanced((a, b), (i, j)):
                        yield a, i, b, j


def iterate_pair_t2(active_occ_idx: Sequence[int], active_unocc_idx: Sequence[int]):
    for i in active_occ_idx:
        for a in active_unocc_idx:
            yield 2 * a, 2 * i, 2 * a + 1, 2 * i + 1


def iterate_pair_t2_generalized(num_orbs: int):
    for i, a in itertools.combinations(range(num_orbs), 2):
        yield 2 * a, 2 * i, 2 * a + 1, 2 * i + 1


# ---- UCC layout (util.py:547-639) ---------------------------------------------------------------
class UccStructure:
    __slots__ = ("excitation_indices", "excitation_operator_type", "n_params", "_sq_cache")

    def __init__(self) -> None:
        self.excitation_indices: list[tuple[int, ...]] = []
        self.excitation_operator_type: list[str] = []
        self.n_params = 0
        self._sq_cache: dict = {}

    def _push(self, kind: str, idx: tuple[int, ...]) -> None:
        self.excitation_operator_type.append(kind)
        self.excitation_indices.append(idx)
        self.n_params += 1

    def add_sa_singles(self, active_occ_idx: Sequence[int], active_unocc_idx: Sequence[int]) -> None:
        for a, i, _ in iterate_t1_sa(active_occ_idx, active_unocc_idx):
            self._push("sa_single", (i, a))

    def add_sa_doubles(self, active_occ_idx: Sequence[int], active_unocc_idx: Sequence[int]) -> None:
        for a, i, b, j, _, op_case in iterate_t2_sa(active_occ_idx, active_unocc_idx):
            self._push(f"sa_double_{op_case}", (i, j, a, b))

    def add_triples(self, active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]) -> None:
        for a, i, b, j, c, k in iterate_t3(active_occ_spin_idx, active_unocc_spin_idx):
            self._push("triple", (i, j, k, a, b, c))

    def add_quadruples(self, active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]) -> None:
        for a, i, b, j, c, k, d, l in iterate_t4(active_occ_spin_idx, active_unocc_spin_idx):
            self._push("quadruple", (i, j, k, l, a, b, c, d))

    def add_quintuples(self, active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]) -> None:
        for a, i, b, j, c, k, d, l, e, m in iterate_t5(active_occ_spin_idx, active_unocc_spin_idx):
            self._push("quintuple", (i, j, k, l, m, a, b, c, d, e))

    def add_sextuples(self, active_occ_spin_idx: Sequence[int], active_unocc_spin_idx: Sequence[int]) -> None:
        for a, i, b, j, c, k, d, l, e, m, f, n in iterate_t6(active_occ_spin_idx, active_unocc_spin_idx):
            self._push("sextuple", (i, j, k, l, m, n, a, b, c, d, e, f))


# ---- UPS layout (util.py:642-1073) --------------------------------------------------------------
class UpsStructure:
    __slots__ = (
        "excitation_indices",
        "excitation_operator_type",
        "grad_param_R",
        "n_params",
        "param_names",
        "_sq_cache",
    )

    def __init__(self) -> None:
        self.excitation_indices: list[tuple[int, ...]] = []
        self.excitation_operator_type: list[str] = []
        self.n_params: int = 0
        self.grad_param_R: dict[str, int] = {}
        self.param_names: list[str] = []
        self._sq_cache: dict = {}

    def _push(self, kind: str, idx: tuple[int, ...], R: int | None) -> None:
        name = f"p{self.n_params:09d}"
        self.excitation_operator_type.append(kind)
        self.excitation_indices.append(idx)
        if R is not None:
            self.grad_param_R[name] = R
        self.param_names.append(name)
        self.n_params += 1

    @staticmethod
    def _check_options(name: str, ansatz_options: dict[str, Any], valid: tuple[str, ...]) -> None:
        for option in ansatz_options:
            if option not in valid:
                raise ValueError(f"Got unknown option for {name}, {option}. Valid options are: {valid}")
        if "n_layers" not in ansatz_options.keys():
            raise ValueError(f"{name} require the option 'n_layers'")

    def create_tiled(self, num_active_orbs: int, ansatz_options: dict[str, Any]) -> None:
        """tUPS / QNP brick-wall ansatz (util.py:653-745)."""
        self._check_options("tUPS", ansatz_options, ("n_layers", "do_qnp", "skip_last_singles", "do_tups", "do_pp"))
        n_layers = ansatz_options["n_layers"]
        do_tups = bool(ansatz_options.get("do_tups", False))
        do_qnp = bool(ansatz_options.get("do_qnp", False))
        if not (do_tups or do_qnp):
            raise ValueError("No tiled ansatz specified.")
        if do_tups and do_qnp:
            raise ValueError("More than one tiled ansatz specfied.")
        skip_last_singles = ansatz_options.get("skip_last_singles", False)
        for layer in range(n_layers):
            last = layer + 1 == n_layers
            for start in (0, 1):  # the two columns of the brick wall
                for p in range(start, num_active_orbs - 1, 2):
                    if do_tups:  # QNP has no leading single
                        self._push("sa_single", (p, p + 1), 4)
                    self._push("double", (2 * p, 2 * p + 1, 2 * p + 2, 2 * p + 3), 2)
                    if last and skip_last_singles and (start == 1 or num_active_orbs == 2):
                        continue
                    self._push("sa_single", (p, p + 1), 4)

    def create_fUCC(
        self,
        occ_idx: list[int],
        unocc_idx: list[int],
        occ_spin_idx: list[int],
        unocc_spin_idx: list[int],
        num_orbs: int,
        ansatz_options: dict[str, Any],
    ) -> None:
        """Factorised UCC (util.py:747-953); option names and operator order as in the reference."""
        valid = ("n_layers", "S", "D", "SAGS", "pD", "GpD", "SAS", "T", "Q", "5", "6", "SAD", "GS", "GD")
        self._check_options("fUCC", ansatz_options, valid)
        flags = {k: bool(ansatz_options.get(k, False)) for k in valid if k != "n_layers"}
        if not any(flags.values()):
            raise ValueError("fUCC requires some excitations got none.")
        for _ in range(ansatz_options["n_layers"]):
            if flags["S"]:
                for a, i in iterate_t1(occ_spin_idx, unocc_spin_idx):
                    self._push("single", (i, a), 2)
            if flags["GS"]:
                for a, i in iterate_t1_generalized(2 * num_orbs):
                    self._push("single", (i, a), 2)
            if flags["SAS"]:
                for a, i, _ in iterate_t1_sa(occ_idx, unocc_idx):
                    self._push("sa_single", (i, a), 4)
            if flags["SAGS"]:
                for a, i, _ in iterate_t1_sa_generalized(num_orbs):
                    self._push("sa_single", (i, a), 4)
            if flags["D"]:
                for a, i, b, j in iterate_t2(occ_spin_idx, unocc_spin_idx):
                    self._push("double", (i, j, a, b), 2)
            if flags["GD"]:
                for a, i, b, j in iterate_t2_generalized(2 * num_orbs):
                    self._push("double", (i, j, a, b), 2)
            if flags["pD"]:
                for a, i, b, j in iterate_pair_t2(occ_idx, unocc_idx):
                    self._push("double", (i, j, a, b), 2)
            if flags["GpD"]:
                for a, i, b, j in iterate_pair_t2_generalized(num_orbs):
                    self._push("double", (i, j, a, b), 2)
            if flags["T"]:
                for a, i, b, j, c, k in iterate_t3(occ_spin_idx, unocc_spin_idx):
                    self._push("triple", (i, j, k, a, b, c), 2)
            if flags["Q"]:
                for a, i, b, j, c, k, d, l in iterate_t4(occ_spin_idx, unocc_spin_idx):
                    self._push("quadruple", (i, j, k, l, a, b, c, d), 2)
            if flags["5"]:
                for a, i, b, j, c, k, d, l, e, m in iterate_t5(occ_spin_idx, unocc_spin_idx):
                    self._push("quintuple", (i, j, k, l, m, a, b, c, d, e), 2)
            if flags["6"]:
                for a, i, b, j, c, k, d, l, e, m, f, n in iterate_t6(occ_spin_idx, unocc_spin_idx):
                    self._push("sextuple", (i, j, k, l, m, n, a, b, c, d, e, f), 2)
            if flags["SAD"]:
                for a, i, b, j, _, op_case in iterate_t2_sa(occ_idx, unocc_idx):
                    # RotoSolve is not defined for SA doubles: no grad_param_R entry (util.py:947-950)
                    self._push(f"sa_double_{op_case}", (i, j, a, b), None)

    def create_SDSfUCC(
        self,
        occ_idx: list[int],
        unocc_idx: list[int],
        occ_spin_idx: list[int],
        unocc_spin_idx: list[int],
        num_orbs: int,
        ansatz_options: dict[str, Any],
    ) -> None:
        """Single-double-single ordered fUCC (util.py:955-1073)."""
        self._check_options("SDSfUCC", ansatz_options, ("n_layers", "D", "pD", "GpD"))
        do_D = bool(ansatz_options.get("D", False))
        do_pD = bool(ansatz_options.get("pD", False))
        do_GpD = bool(ansatz_options.get("GpD", False))
        if not (do_D or do_pD or do_GpD):
            raise ValueError("SDSfUCC requires some excitations got none.")
        for _ in range(ansatz_options["n_layers"]):
            if do_D:
                for a, i, b, j in iterate_t2(occ_spin_idx, unocc_spin_idx):
                    same = i % 2 == a % 2
                    self._push("single", (i, a) if same else (i, b), 2)
                    self._push("double", (i, j, a, b), 2)
                    self._push("single", (j, b) if same else (j, a), 2)
            if do_pD:
                for a, i, b, j in iterate_pair_t2(occ_idx, unocc_idx):
                    # the reference labels this leading spin-adapted single "double" with a 2-tuple
                    # (util.py:1040-1043); that entry cannot be applied there either.  Kept verbatim.
                    self._push("double", (i // 2, a // 2), 4)
                    self._push("double", (i, j, a, b), 2)
                    self._push("sa_single", (i // 2, a // 2), 4)
            if do_GpD:
                for a, i, b, j in iterate_pair_t2_generalized(num_orbs):
                    self._push("sa_single", (i // 2, a // 2), 4)
                    self._push("double", (i, j, a, b), 2)
                    self._push("sa_single", (i // 2, a // 2), 4)
