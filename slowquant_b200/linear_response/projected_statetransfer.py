"""Projected state-transfer linear response (same class name and results as the reference's
slowquant/unitary_coupled_cluster/linear_response/projected_statetransfer.py).

Active-space excitations are state-transfer operators (G-G block A = GC X^T - E 1 and Sigma = 1 from the panels of
statetransfer.py), the orbital rotations are projected (q-q blocks as in allprojected.py), and the coupling block is
A_Gq = GC (U^d H q|0>)^T with B = 0 (projected_statetransfer.py:82-137).
"""
from __future__ import annotations

from slowquant_b200.linear_response._symbolic import projected_orbital_blocks
from slowquant_b200.linear_response.statetransfer import LinearResponse as _StateTransfer
from slowquant_b200.operators import hamiltonian_2i_2a
from slowquant_b200.ucc_wavefunction import WaveFunctionUCC
from slowquant_b200.ups_wavefunction import WaveFunctionUPS


class LinearResponse(_StateTransfer):
    def __init__(self, wave_function: WaveFunctionUCC | WaveFunctionUPS, excitations: str) -> None:
        super().__init__(wave_function, excitations)
        wf = self.wf
        nq = len(self.q_ops)
        H_2i_2a = hamiltonian_2i_2a(wf.h_mo, wf.g_mo, wf.num_inactive_orbs, wf.num_active_orbs, wf.num_virtual_orbs)
        projected_orbital_blocks(self, H_2i_2a, self._psi, wf.ci_info)
        self.B[nq:, :nq] = 0.0
        self.B[:nq, nq:] = 0.0
