"""Host-side helper for the orbital-rotation blocks of linear response.

The reference multiplies the full ``hamiltonian_1i_1a`` with every (G, q) pair and folds the product onto the
active space afterwards (naive.py:143-150, 177-184); almost every string of that product is discarded by the fold.
A string survives ``get_folded_operator`` only if it has no virtual index and its inactive creators equal its
inactive annihilators -- and normal ordering never changes, per spin orbital, the number of creators minus the number
of annihilators.  So a term ``h * t`` can survive only if that net count of ``h`` is the negative of the one of ``t`` on
every inactive / virtual spin orbital.  ``SectorSplit`` indexes the Hamiltonian strings by this signature once; a
product then touches only the matching strings.  The folded result is identical to folding the full product.
"""
from __future__ import annotations

from slowquant_b200.fermionic_operator import FermionicOperator


def _signature(label, n_in: int, n_act_end: int) -> tuple:
    net: dict[int, int] = {}
    for idx, dag in label:
        if idx < n_in or idx >= n_act_end:
            net[idx] = net.get(idx, 0) + (1 if dag else -1)
    return tuple(sorted((i, c) for i, c in net.items() if c != 0))


def _negated(sig: tuple) -> tuple:
    return tuple((i, -c) for i, c in sig)


class SectorSplit:
    """``op`` split by the net inactive / virtual ladder content of its strings."""

    def __init__(self, op: FermionicOperator, num_inactive_orbs: int, num_active_orbs: int) -> None:
        self.n_in = 2 * num_inactive_orbs
        self.n_act_end = self.n_in + 2 * num_active_orbs
        groups: dict[tuple, dict] = {}
        for label, fac in op.operators.items():
            groups.setdefault(_signature(label, self.n_in, self.n_act_end), {})[label] = fac
        self.parts = {sig: FermionicOperator(d) for sig, d in groups.items()}

    def _split(self, small: FermionicOperator) -> dict[tuple, FermionicOperator]:
        groups: dict[tuple, dict] = {}
        for label, fac in small.operators.items():
            groups.setdefault(_signature(label, self.n_in, self.n_act_end), {})[label] = fac
        return {sig: FermionicOperator(d) for sig, d in groups.items()}

    def times(self, small: FermionicOperator) -> FermionicOperator:
        """Fold-surviving part of ``op * small``."""
        out = FermionicOperator({})
        for sig, piece in self._split(small).items():
            part = self.parts.get(_negated(sig))
            if part is not None:
                out += part * piece
        return out

    def rtimes(self, small: FermionicOperator) -> FermionicOperator:
        """Fold-surviving part of ``small * op``."""
        out = FermionicOperator({})
        for sig, piece in self._split(small).items():
            part = self.parts.get(_negated(sig))
            if part is not None:
                out += piece * part
        return out

    def sandwich(self, left: FermionicOperator, right: FermionicOperator) -> FermionicOperator:
        """Fold-surviving part of ``left * op * right``."""
        out = FermionicOperator({})
        for sl, pl in self._split(left).items():
            for sr, pr in self._split(right).items():
                net: dict[int, int] = {}
                for i, c in sl + sr:
                    net[i] = net.get(i, 0) + c
                need = tuple(sorted((i, -c) for i, c in net.items() if c != 0))
                part = self.parts.get(need)
                if part is not None:
                    out += pl * part * pr
        return out


def projected_orbital_blocks(lr, H_2i_2a: FermionicOperator, psi, ci_info) -> None:
    """q-q blocks of the projected parametrisations (allprojected.py:86-112, projected_statetransfer.py:82-104):
    A = <0|q_I^d H q_J|0> - E <0|q_I^d q_J|0>, Sigma = <0|q_I^d q_J|0>, B = 0, every expectation value over a product
    folded onto the active space (only the strings of ``hamiltonian_2i_2a`` that can survive the fold are multiplied)."""
    import torch

    from slowquant_b200.linear_response import _panels as pn

    wf = lr.wf
    nq = len(lr.q_ops)
    if nq == 0:
        return
    H2 = SectorSplit(H_2i_2a, wf.num_inactive_orbs, wf.num_active_orbs)
    E = wf.energy_elec
    tmp = torch.empty_like(psi)
    lr.A[:nq, :nq] = 0.0
    lr.B[:nq, :nq] = 0.0
    lr.Sigma[:nq, :nq] = 0.0
    q_dag = [q.dagger for q in lr.q_ops]
    for j, qJ in enumerate(lr.q_ops):
        for i in range(j, nq):
            pn.apply_into(H2.sandwich(q_dag[i], qJ), psi, tmp, ci_info)
            hqq = float(torch.dot(psi, tmp))
            pn.apply_into(q_dag[i] * qJ, psi, tmp, ci_info)
            qq = float(torch.dot(psi, tmp))
            lr.A[i, j] = lr.A[j, i] = hqq - qq * E
            lr.Sigma[i, j] = lr.Sigma[j, i] = qq
