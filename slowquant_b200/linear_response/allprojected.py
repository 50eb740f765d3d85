"""All-projected linear response (same class name and results as the reference's
slowquant/unitary_coupled_cluster/linear_response/allprojected.py).

The G-G and G-q blocks of A and the G-G blocks of B and Sigma are those of the projected parametrisation (device panels,
projected.py); the orbital rotations are projected as well, so the q-q blocks are expectation values of folded
products with ``hamiltonian_2i_2a`` instead of the RDM formulas, and the G-q block of B vanishes (allprojected.py:86-125).
"""
from __future__ import annotations

import numpy as np
import torch

from slowquant_b200.integral_manager import one_electron_integral_transform
from slowquant_b200.linear_response import _panels as pn
from slowquant_b200.linear_response._symbolic import projected_orbital_blocks
from slowquant_b200.linear_response.projected import LinearResponse as _Projected
from slowquant_b200.operators import hamiltonian_2i_2a, one_elec_op_0i_0a
from slowquant_b200.ucc_wavefunction import WaveFunctionUCC
from slowquant_b200.ups_wavefunction import WaveFunctionUPS


class LinearResponse(_Projected):
    def __init__(self, wave_function: WaveFunctionUCC | WaveFunctionUPS, excitations: str) -> None:
        super().__init__(wave_function, excitations)
        wf = self.wf
        nq = len(self.q_ops)
        H_2i_2a = hamiltonian_2i_2a(wf.h_mo, wf.g_mo, wf.num_inactive_orbs, wf.num_active_orbs, wf.num_virtual_orbs)
        projected_orbital_blocks(self, H_2i_2a, self._psi, wf.ci_info)
        self.B[nq:, :nq] = 0.0
        self.B[:nq, nq:] = 0.0

    def get_transition_dipole(self) -> np.ndarray:
        """allprojected.py:292-486: Z_i (<0|G_i^d|0><0|mu|0> - <0|G_i^d mu|0>) - Y_i (<0|G_i|0><0|mu|0> - <0|mu G_i|0>)."""
        wf = self.wf
        ci_info = wf.ci_info
        nI, nA = wf.num_inactive_orbs, wf.num_active_orbs
        number_excitations = len(self.excitation_energies)
        dev = self._psi.device
        Z = torch.from_numpy(np.ascontiguousarray(self.Z_G_normed.T)).to(dev)
        Y = torch.from_numpy(np.ascontiguousarray(self.Y_G_normed.T)).to(dev)
        g0 = self._Gk @ self._psi
        dipole_integrals = wf.int_gen.electric_dipole
        transition_dipoles = np.zeros((number_excitations, 3))
        for axis in range(3):
            mu = one_electron_integral_transform(wf.c_mo, dipole_integrals[axis])
            mu_op = one_elec_op_0i_0a(mu, nI, nA)
            mu_ket = pn.apply(mu_op, self._psi, ci_info)
            mud_ket = pn.apply(mu_op.dagger, self._psi, ci_info)
            exp_mu = torch.dot(self._psi, mu_ket)
            active = (Z @ (g0 * exp_mu - self._Gk @ mu_ket) - Y @ (g0 * exp_mu - self._Gk @ mud_ket)).cpu().numpy()
            for s in range(number_excitations):
                transition_dipoles[s, axis] = pn.orbital_property_part(self, mu, s, number_excitations) + active[s]
        return transition_dipoles
