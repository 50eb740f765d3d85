"""Operator factories (host side).  Same names and argument meaning as the reference's
slowquant/unitary_coupled_cluster/operators.py; the strings they produce are the input format of the
CUDA kernels.  ``hamiltonian_0i_0a`` returns a lazy operator that carries the integrals, so that the
engine can take (e_core, h_eff, g_act) directly instead of 2n^2 + 2C(n,2)^2 + n^4 strings.
"""
from __future__ import annotations

import numpy as np

from slowquant_b200.fermionic_operator import FermionicOperator


def a_op(spinless_idx: int, spin: str, dagger: bool) -> FermionicOperator:
    """Annihilation/creation operator on spatial orbital `spinless_idx` (operators.py:8-24)."""
    if spin not in ("alpha", "beta"):
        raise ValueError(f'spin must be "alpha" or "beta" got {spin}')
    return FermionicOperator({((2 * spinless_idx + (spin == "beta"), dagger),): 1})


def a_op_spin(spin_idx: int, dagger: bool) -> FermionicOperator:
    """Annihilation/creation operator on spin orbital `spin_idx` (operators.py:27-37)."""
    return FermionicOperator({((spin_idx, dagger),): 1})


def Epq(p: int, q: int) -> FermionicOperator:
    r"""Singlet one-electron excitation operator :math:`E_{pq}=\sum_\sigma a^\dagger_{p\sigma}a_{q\sigma}` (operators.py:40-58)."""
    return FermionicOperator(
        {
            ((2 * p, True), (2 * q, False)): 1,
            ((2 * p + 1, True), (2 * q + 1, False)): 1,
        }
    )


def epqrs(p: int, q: int, r: int, s: int) -> FermionicOperator:
    r""":math:`e_{pqrs}=E_{pq}E_{rs}-\delta_{qr}E_{ps}` (operators.py:61-80)."""
    op = Epq(p, q) * Epq(r, s)
    if q == r:
        op -= Epq(p, s)
    return op


def Eminuspq(p: int, q: int) -> FermionicOperator:
    r""":math:`E^-_{pq}=E_{pq}-E_{qp}` (operators.py:83-98)."""
    return Epq(p, q) - Epq(q, p)


def commutator(A: FermionicOperator, B: FermionicOperator) -> FermionicOperator:
    """[A, B] (operators.py:101-114)."""
    return A * B - B * A


def double_commutator(
    A: FermionicOperator, B: FermionicOperator, C: FermionicOperator, do_symmetrized: bool = False
) -> FermionicOperator:
    """[A, [B, C]] or its symmetrised form (operators.py:117-142)."""
    if do_symmetrized:
        return A * B * C + C * B * A - 1 / 2 * (A * C * B + B * C * A + C * A * B + B * A * C)
    return A * B * C - A * C * B - B * C * A + C * B * A


def _excitation(occ: tuple[int, ...], unocc: tuple[int, ...], return_anti_hermitian: bool) -> FermionicOperator:
    """a+_{a} a+_{b} ... a_{k} a_{j} a_{i} for occ=(i,j,k,..), unocc=(a,b,..) spin-orbital indices."""
    label = tuple((a, True) for a in unocc) + tuple((i, False) for i in reversed(occ))
    op = FermionicOperator({(): 1.0}) * FermionicOperator({label: 1})
    if return_anti_hermitian:
        op -= op.dagger
    return op


def G1(i: int, a: int, return_anti_hermitian: bool = False) -> FermionicOperator:
    """One-electron excitation a+_a a_i on spin orbitals (operators.py:145-163)."""
    return _excitation((i,), (a,), return_anti_hermitian)


def G2(i: int, j: int, a: int, b: int, return_anti_hermitian: bool = False) -> FermionicOperator:
    """Two-electron excitation a+_a a+_b a_j a_i (operators.py:166-188)."""
    return _excitation((i, j), (a, b), return_anti_hermitian)


def G3(i: int, j: int, k: int, a: int, b: int, c: int, return_anti_hermitian: bool = False) -> FermionicOperator:
    """Three-electron excitation (operators.py:191-219)."""
    return _excitation((i, j, k), (a, b, c), return_anti_hermitian)


def G4(
    i: int, j: int, k: int, l: int, a: int, b: int, c: int, d: int, return_anti_hermitian: bool = False
) -> FermionicOperator:
    """Four-electron excitation (operators.py:222-254)."""
    return _excitation((i, j, k, l), (a, b, c, d), return_anti_hermitian)


def G5(
    i: int, j: int, k: int, l: int, m: int, a: int, b: int, c: int, d: int, e: int,
    return_anti_hermitian: bool = False,
) -> FermionicOperator:
    """Five-electron excitation (operators.py:257-303)."""
    return _excitation((i, j, k, l, m), (a, b, c, d, e), return_anti_hermitian)


def G6(
    i: int, j: int, k: int, l: int, m: int, n: int, a: int, b: int, c: int, d: int, e: int, f: int,
    return_anti_hermitian: bool = False,
) -> FermionicOperator:
    """Six-electron excitation (operators.py:306-359)."""
    return _excitation((i, j, k, l, m, n), (a, b, c, d, e, f), return_anti_hermitian)


def G1_sa(i: int, a: int, return_anti_hermitian: bool = False) -> FermionicOperator:
    r"""Spin-adapted single :math:`E_{ai}/\sqrt2` on spatial orbitals (operators.py:362-379)."""
    op = 2 ** (-1 / 2) * Epq(a, i)
    if return_anti_hermitian:
        op -= op.dagger
    return op


def G2_sa(i: int, j: int, a: int, b: int, case: int, return_anti_hermitian: bool = False) -> FermionicOperator:
    """Spin-adapted doubles, cases 1-5 (operators.py:382-443)."""
    if case in (1, 2, 3, 4):
        fac = 1
        if a == b:
            fac *= 2
        if i == j:
            fac *= 2
        op = 1 / 2 * (fac) ** (-1 / 2) * (Epq(a, i) * Epq(b, j) + Epq(a, j) * Epq(b, i))
    elif case == 5:
        op = 1 / (2 * 3 ** (1 / 2)) * (Epq(a, i) * Epq(b, j) - Epq(a, j) * Epq(b, i))
    else:
        raise ValueError("Got unknown case for spin-adapted doubles, {case}")
    if return_anti_hermitian:
        op -= op.dagger
    return op


def hamiltonian_full_space(h_mo: np.ndarray, g_mo: np.ndarray, num_orbs: int) -> FermionicOperator:
    """Full-space electronic Hamiltonian as explicit strings (operators.py:446-473)."""
    H = FermionicOperator({})
    for p in range(num_orbs):
        for q in range(num_orbs):
            if abs(h_mo[p, q]) < 10**-14:
                continue
            H += float(h_mo[p, q]) * Epq(p, q)
    for p in range(num_orbs):
        for q in range(num_orbs):
            for r in range(num_orbs):
                for s in range(num_orbs):
                    if abs(g_mo[p, q, r, s]) < 10**-14:
                        continue
                    H += (1 / 2 * float(g_mo[p, q, r, s])) * epqrs(p, q, r, s)
    return H


def fold_hamiltonian_0i_0a(
    h_mo: np.ndarray, g_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int
) -> tuple[float, np.ndarray, np.ndarray]:
    r"""Closed form of ``hamiltonian_0i_0a(...).get_folded_operator(...)`` (operators.py:476-529 folded by
    fermionic_operator.py:379-471):

    .. math::
        H_A = E_\text{core} + \sum_{vw} h^\text{eff}_{vw}E_{vw} + \tfrac12\sum_{vwxy} g_{vwxy} e_{vwxy}

    with exactly the integral elements the reference touches (it does not assume permutational symmetry of g).
    """
    nI, nA = num_inactive_orbs, num_active_orbs
    act = slice(nI, nI + nA)
    h = np.where(np.abs(h_mo) > 10**-14, h_mo, 0.0)
    g = np.where(np.abs(g_mo) > 10**-14, g_mo, 0.0)
    e_core = 0.0
    for i in range(nI):
        e_core += 2 * h[i, i]
        for j in range(nI):
            e_core += 2 * g[i, i, j, j]
            if i != j:
                e_core -= g[j, i, i, j]
            else:
                e_core -= g[i, i, i, i]
    h_eff = np.array(h[act, act], dtype=np.float64)
    for i in range(nI):
        h_eff += g[i, i, act, act] + g[act, act, i, i] - 0.5 * g[act, i, i, act]
        h_eff -= 0.5 * g[i, act, act, i].T
    g_act = np.ascontiguousarray(g[act, act, act, act], dtype=np.float64)
    return float(e_core), np.ascontiguousarray(h_eff), g_act


class ActiveSpaceHamiltonian(FermionicOperator):
    """``hamiltonian_0i_0a`` result: behaves as a FermionicOperator (strings built on first access of
    ``.operators``) and additionally carries the integrals for the dedicated sigma kernel."""

    __slots__ = ("_ops", "h_mo", "g_mo", "num_inactive_orbs", "num_active_orbs")

    def __init__(self, h_mo: np.ndarray, g_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int) -> None:
        self._ops = None
        self.h_mo = h_mo
        self.g_mo = g_mo
        self.num_inactive_orbs = num_inactive_orbs
        self.num_active_orbs = num_active_orbs

    @property
    def operators(self):  # type: ignore[override]
        if self._ops is None:
            self._ops = _hamiltonian_0i_0a_strings(
                self.h_mo, self.g_mo, self.num_inactive_orbs, self.num_active_orbs
            ).operators
        return self._ops

    @operators.setter
    def operators(self, value) -> None:
        self._ops = value

    def folded_integrals(self) -> tuple[float, np.ndarray, np.ndarray]:
        return fold_hamiltonian_0i_0a(self.h_mo, self.g_mo, self.num_inactive_orbs, self.num_active_orbs)


def _hamiltonian_0i_0a_strings(
    h_mo: np.ndarray, g_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int
) -> FermionicOperator:
    """Explicit-string energy Hamiltonian; term list of operators.py:476-529."""
    nI, nA = num_inactive_orbs, num_active_orbs
    H = FermionicOperator({})
    act = range(nI, nI + nA)
    for i in range(nI):
        if abs(h_mo[i, i]) > 10**-14:
            H += float(h_mo[i, i]) * Epq(i, i)
    for p in act:
        for q in act:
            if abs(h_mo[p, q]) > 10**-14:
                H += float(h_mo[p, q]) * Epq(p, q)
    for i in range(nI):
        for j in range(nI):
            if abs(g_mo[i, i, j, j]) > 10**-14:
                H += (1 / 2 * float(g_mo[i, i, j, j])) * epqrs(i, i, j, j)
            if i != j and abs(g_mo[j, i, i, j]) > 10**-14:
                H += (1 / 2 * float(g_mo[j, i, i, j])) * epqrs(j, i, i, j)
    for i in range(nI):
        for p in act:
            for q in act:
                for (w, x, y, z) in ((i, i, p, q), (p, q, i, i), (p, i, i, q), (i, p, q, i)):
                    if abs(g_mo[w, x, y, z]) > 10**-14:
                        H += (1 / 2 * float(g_mo[w, x, y, z])) * epqrs(w, x, y, z)
    for p in act:
        for q in act:
            for r in act:
                for s in act:
                    if abs(g_mo[p, q, r, s]) > 10**-14:
                        H += (1 / 2 * float(g_mo[p, q, r, s])) * epqrs(p, q, r, s)
    return H


def hamiltonian_0i_0a(
    h_mo: np.ndarray, g_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int
) -> FermionicOperator:
    """Energy Hamiltonian (no inactive/virtual excitations); same call as operators.py:476-529."""
    return ActiveSpaceHamiltonian(h_mo, g_mo, num_inactive_orbs, num_active_orbs)


def one_elec_op_0i_0a(ints_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int) -> FermionicOperator:
    """One-electron operator restricted to inactive diagonal + active block (operators.py:687-708)."""
    op = FermionicOperator({})
    for i in range(num_inactive_orbs):
        if abs(ints_mo[i, i]) > 10**-14:
            op += float(ints_mo[i, i]) * Epq(i, i)
    for p in range(num_inactive_orbs, num_inactive_orbs + num_active_orbs):
        for q in range(num_inactive_orbs, num_inactive_orbs + num_active_orbs):
            if abs(ints_mo[p, q]) > 10**-14:
                op += float(ints_mo[p, q]) * Epq(p, q)
    return op


def _inactive_virtual_changes(p: int, q: int, r: int, s: int, num_inactive_orbs: int, virtual_start: int) -> tuple[int, int]:
    """(net inactive changes, virtual indices) of e_pqrs, counted as operators.py:562-591: every inactive index
    counts once, and every create/annihilate pair on the same inactive orbital ((p,q), (r,s), (p,s), (q,r)) gives two back."""
    idx = (p, q, r, s)
    n_virt = sum(1 for x in idx if x >= virtual_start)
    n_inact = sum(1 for x in idx if x < num_inactive_orbs)
    for x, y in ((p, q), (r, s), (p, s), (q, r)):
        if x == y and x < num_inactive_orbs:
            n_inact -= 2
    return n_inact, n_virt


def _hamiltonian_ni_na(
    h_mo: np.ndarray, g_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int, num_virtual_orbs: int, level: int
) -> FermionicOperator:
    num_orbs = num_inactive_orbs + num_active_orbs + num_virtual_orbs
    virtual_start = num_inactive_orbs + num_active_orbs
    H = FermionicOperator({})
    for p in range(num_orbs):
        for q in range(num_orbs):
            if level == 1:
                if p >= virtual_start and q >= virtual_start:
                    continue
                if p < num_inactive_orbs and q < num_inactive_orbs and p != q:
                    continue
            if abs(h_mo[p, q]) > 10**-14:
                H += float(h_mo[p, q]) * Epq(p, q)
    for p in range(num_orbs):
        for q in range(num_orbs):
            for r in range(num_orbs):
                for s in range(num_orbs):
                    if abs(g_mo[p, q, r, s]) <= 10**-14:
                        continue
                    n_inact, n_virt = _inactive_virtual_changes(p, q, r, s, num_inactive_orbs, virtual_start)
                    if n_virt > level or n_inact > level:
                        continue
                    H += (1 / 2 * float(g_mo[p, q, r, s])) * epqrs(p, q, r, s)
    return H


def hamiltonian_1i_1a(
    h_mo: np.ndarray, g_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int, num_virtual_orbs: int
) -> FermionicOperator:
    """Hamiltonian terms with at most one inactive and one virtual change (operators.py:532-598): the operator
    that multiplies ONE orbital-rotation generator in the linear-response blocks."""
    return _hamiltonian_ni_na(h_mo, g_mo, num_inactive_orbs, num_active_orbs, num_virtual_orbs, 1)


def hamiltonian_2i_2a(
    h_mo: np.ndarray, g_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int, num_virtual_orbs: int
) -> FermionicOperator:
    """Hamiltonian terms with at most two inactive and two virtual changes (operators.py:601-663)."""
    return _hamiltonian_ni_na(h_mo, g_mo, num_inactive_orbs, num_active_orbs, num_virtual_orbs, 2)


def one_elec_op_full_space(ints_mo: np.ndarray, num_orbs: int) -> FermionicOperator:
    r""":math:`\sum_{pq} o_{pq} E_{pq}` over all orbitals (operators.py:666-684)."""
    op = FermionicOperator({})
    for p in range(num_orbs):
        for q in range(num_orbs):
            if abs(ints_mo[p, q]) > 10**-14:
                op += float(ints_mo[p, q]) * Epq(p, q)
    return op


def one_elec_op_1i_1a(
    ints_mo: np.ndarray, num_inactive_orbs: int, num_active_orbs: int, num_virtual_orbs: int
) -> FermionicOperator:
    """One-electron operator with at most one inactive/virtual change (operators.py:711-737)."""
    num_orbs = num_inactive_orbs + num_active_orbs + num_virtual_orbs
    virtual_start = num_inactive_orbs + num_active_orbs
    op = FermionicOperator({})
    for p in range(num_orbs):
        for q in range(num_orbs):
            if p >= virtual_start and q >= virtual_start:
                continue
            if p < num_inactive_orbs and q < num_inactive_orbs and p != q:
                continue
            if abs(ints_mo[p, q]) > 10**-14:
                op += float(ints_mo[p, q]) * Epq(p, q)
    return op
