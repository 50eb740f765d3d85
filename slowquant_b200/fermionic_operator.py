"""Host-side symbolic fermionic operator algebra (input format of the state-vector engine).

Mirrors the interface of the reference's ``FermionicOperator``
(slowquant/unitary_coupled_cluster/fermionic_operator.py:107-488): an operator is a dictionary

    { ((spin_orbital_index, is_creation), ...) : coefficient }

whose keys are *normal ordered* ladder strings: all creation operators first, then all annihilation
operators, each block sorted by descending spin-orbital index (fermionic_operator.py:27-40).  The
normal form of an operator is unique, so any correct reduction produces the same dictionary as the
reference (up to insertion order); the reduction below is a Wick-style insertion, not the reference's
bubble passes.  This stays on the host: it is tiny symbolic work that produces the strings the CUDA
kernels consume.
"""
from __future__ import annotations

from collections.abc import Iterable

Ladder = tuple[int, bool]
Label = tuple[Ladder, ...]

_DROP = 10**-14  # threshold below which a coefficient created by cancellation is removed (fermionic_operator.py:100)


def operator_to_qiskit_key(operator_string: Label, remapping: dict[int, int]) -> str:
    """``"+_i -_j ..."`` key of one ladder string with remapped spin-orbital indices (fermionic_operator.py:7-24)."""
    return " ".join(("+_" if dagger else "-_") + str(remapping[idx]) for idx, dagger in operator_string)


def _canonical_before(x: Ladder, y: Ladder) -> bool:
    """True when x may stand directly left of y in a normal-ordered string (x != y assumed)."""
    if x[1] != y[1]:
        return x[1]  # creator left of annihilator
    return x[0] > y[0]  # descending index inside a block


def _insert(sorted_ops: list[Ladder], op: Ladder, factor: float, out: list[tuple[list[Ladder], float]]) -> None:
    """Multiply the normal-ordered string `sorted_ops` from the right by `op` and normal order the result.

    Moves `op` leftwards by anticommutation; every time it passes an annihilator/creator pair on the same
    orbital a contraction term (the string without the pair) is emitted.  Results are appended to `out`.
    """
    pos = len(sorted_ops)
    sign = 1.0
    while pos > 0:
        left = sorted_ops[pos - 1]
        if left == op:
            return  # a a = 0 and a+ a+ = 0
        if _canonical_before(left, op):
            break
        if (not left[1]) and op[1] and left[0] == op[0]:
            # a_p a+_p = 1 - a+_p a_p : contraction term keeps everything except this pair
            rest = sorted_ops[: pos - 1] + sorted_ops[pos:]
            out.append((rest, factor * sign))
        sign = -sign
        pos -= 1
    # a repeated operator further left would also kill the string; it can only sit directly at pos-1
    out.append((sorted_ops[:pos] + [op] + sorted_ops[pos:], factor * sign))


def normal_order(label: Iterable[Ladder], factor: float) -> dict[Label, float]:
    """Normal order one ladder string; returns {normal-ordered label: coefficient}."""
    current: list[tuple[list[Ladder], float]] = [([], factor)]
    for op in label:
        nxt: list[tuple[list[Ladder], float]] = []
        for ops, fac in current:
            _insert(ops, op, fac, nxt)
        current = nxt
    result: dict[Label, float] = {}
    for ops, fac in current:
        key = tuple(ops)
        if key in result:
            result[key] += fac
            if abs(result[key]) < _DROP:
                del result[key]
        else:
            result[key] = fac
    return result


def do_extended_normal_ordering(fermistring: "FermionicOperator") -> dict[Label, float]:
    """Normal order every string of an operator (same contract as fermionic_operator.py:27-104)."""
    result: dict[Label, float] = {}
    for label, fac in fermistring.operators.items():
        for key, val in normal_order(label, fac).items():
            if key in result:
                result[key] += val
                if abs(result[key]) < _DROP:
                    del result[key]
            else:
                result[key] = val
    return result


class FermionicOperator:
    """Sum of normal-ordered ladder strings with real coefficients."""

    __slots__ = ("operators",)

    def __init__(self, annihilation_operator: dict[Label, float]) -> None:
        if not isinstance(annihilation_operator, dict):
            raise ValueError(f"Could not assign operator of {type(annihilation_operator)}.")
        self.operators = annihilation_operator

    # ---- linear structure -------------------------------------------------------------------
    @staticmethod
    def _accumulate(target: dict[Label, float], source: dict[Label, float], scale: float) -> None:
        for key, val in source.items():
            if key in target:
                target[key] += scale * val
                if abs(target[key]) < _DROP:
                    del target[key]
            else:
                target[key] = scale * val

    def __add__(self, other: "FermionicOperator") -> "FermionicOperator":
        ops = dict(self.operators)
        self._accumulate(ops, other.operators, 1.0)
        return FermionicOperator(ops)

    def __iadd__(self, other: "FermionicOperator") -> "FermionicOperator":
        self._accumulate(self.operators, other.operators, 1.0)
        return self

    def __sub__(self, other: "FermionicOperator") -> "FermionicOperator":
        ops = dict(self.operators)
        self._accumulate(ops, other.operators, -1.0)
        return FermionicOperator(ops)

    def __isub__(self, other: "FermionicOperator") -> "FermionicOperator":
        self._accumulate(self.operators, other.operators, -1.0)
        return self

    def __neg__(self) -> "FermionicOperator":
        return FermionicOperator({k: -v for k, v in self.operators.items()})

    # ---- products ---------------------------------------------------------------------------
    def _product(self, other: "FermionicOperator") -> dict[Label, float]:
        # (self * other): strings of self stand to the LEFT of strings of other
        result: dict[Label, float] = {}
        for right, fr in other.operators.items():
            for left, fl in self.operators.items():
                self._accumulate(result, normal_order(left + right, fl * fr), 1.0)
        return result

    def __mul__(self, other: "FermionicOperator | float | int") -> "FermionicOperator":
        if type(other) in (float, int):
            return FermionicOperator({k: v * other for k, v in self.operators.items()})
        if isinstance(other, FermionicOperator):
            return FermionicOperator(self._product(other))
        raise TypeError(f"Got unknown type of fermistring: {type(other)}")

    def __imul__(self, other: "FermionicOperator | float | int") -> "FermionicOperator":
        if type(other) in (float, int):
            for k in self.operators:
                self.operators[k] *= other
            return self
        if isinstance(other, FermionicOperator):
            self.operators = self._product(other)
            return self
        raise TypeError(f"Got unknown type of fermistring: {type(other)}")

    def __rmul__(self, number: float) -> "FermionicOperator":
        return FermionicOperator({k: v * number for k, v in self.operators.items()})

    @property
    def dagger(self) -> "FermionicOperator":
        """Hermitian conjugate, normal ordered (fermionic_operator.py:302-321)."""
        flipped: dict[Label, float] = {}
        for label, fac in self.operators.items():
            flipped[tuple((idx, not dag) for idx, dag in reversed(label))] = fac
        return FermionicOperator(do_extended_normal_ordering(FermionicOperator(flipped)))

    # ---- inspection -------------------------------------------------------------------------
    @property
    def operator_count(self) -> dict[int, int]:
        count: dict[int, int] = {}
        for label in self.operators:
            count[len(label)] = count.get(len(label), 0) + 1
        return count

    @property
    def operators_readable(self) -> dict[str, float]:
        return {
            "".join(("c" if dag else "a") + str(idx) for idx, dag in label): fac
            for label, fac in self.operators.items()
        }

    # ---- folding onto the active space --------------------------------------------------------
    def get_qiskit_form(self, num_orbs: int) -> dict[str, float]:
        """Operator as ``{"+_i -_j": coefficient}`` in blocked spin order -- interleaved index 2p + s maps to p + s * num_orbs
        (all alpha, then all beta; fermionic_operator.py:357-377).  String formatting only; no Qiskit import."""
        remapping = {2 * p + spin: p + spin * num_orbs for p in range(num_orbs) for spin in (0, 1)}
        return {operator_to_qiskit_key(label, remapping): factor for label, factor in self.operators.items()}

    def get_folded_operator(
        self, num_inactive_orbs: int, num_active_orbs: int, num_virtual_orbs: int
    ) -> "FermionicOperator":
        r"""Fold inactive / virtual indices away (contract of fermionic_operator.py:379-471).

        For a state |I> (x) |A> (x) |vac_V>: a string contributes only if it has no virtual index and
        its inactive creators equal its inactive annihilators (as ordered lists).  The surviving
        active part is re-indexed from 0 and picks up
          * (-1) if the numbers of inactive creators and of active creators are both odd
            (moving the inactive annihilator block through the active creator block), and
          * (-1)^{k(k-1)/2}... expressed in the reference as one factor -1 per even position of the
            inactive annihilator block (reversal of that block), k = number of inactive annihilators.
        """
        n_in = 2 * num_inactive_orbs
        n_act_end = n_in + 2 * num_active_orbs
        folded: dict[Label, float] = {}
        for label, fac in self.operators.items():
            inactive_c: list[int] = []
            inactive_a: list[int] = []
            active_c: list[Ladder] = []
            active_a: list[Ladder] = []
            has_virtual = False
            for idx, dag in label:
                if idx < n_in:
                    (inactive_c if dag else inactive_a).append(idx)
                elif idx < n_act_end:
                    (active_c if dag else active_a).append((idx - n_in, dag))
                else:
                    has_virtual = True
                    break
            if has_virtual or inactive_c != inactive_a:
                continue
            sign = 1.0
            if len(inactive_c) % 2 == 1 and len(active_c) % 2 == 1:
                sign = -sign
            if (len(inactive_a) // 2) % 2 == 1:
                sign = -sign
            key = tuple(active_c + active_a)
            if key in folded:
                folded[key] += sign * fac
            else:
                folded[key] = sign * fac
        return FermionicOperator(folded)

    def get_info(self) -> tuple[list[list[int]], list[list[int]], list[float]]:
        """Creation indices, annihilation indices and coefficient of every string."""
        creation, annihilation, coefficients = [], [], []
        for label, fac in self.operators.items():
            creation.append([idx for idx, dag in label if dag])
            annihilation.append([idx for idx, dag in label if not dag])
            coefficients.append(fac)
        return annihilation, creation, coefficients
