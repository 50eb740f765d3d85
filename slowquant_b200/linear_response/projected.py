"""Projected linear response (same class name and results as the reference's
slowquant/unitary_coupled_cluster/linear_response/projected.py), built from device panels.

With the panel Gk[I] = G_I|0>, its image HG = H Gk and the vectors g0 = Gk psi, gh = Gk (H psi), hg = HG psi, the
G-G blocks of projected.py:166-290 are

    A = Gk HG^T + E (g0 g0^T - Gk Gk^T) - 1/2 (g0 hg^T + gh g0^T)
    B = 1/2 (gh g0^T + g0 gh^T) - E g0 g0^T
    Sigma = Gk Gk^T - g0 g0^T

(lower triangle mirrored as the reference's i >= j loop does); the q-G blocks are Gk (H q|0>)^T and -1/2 Gk (q^d H|0>)^T.
"""
from __future__ import annotations

import numpy as np
import torch

from slowquant_b200.integral_manager import one_electron_integral_transform
from slowquant_b200.linear_response import _panels as pn
from slowquant_b200.linear_response._symbolic import SectorSplit
from slowquant_b200.linear_response.lr_baseclass import LinearResponseBaseClass
from slowquant_b200.operators import one_elec_op_0i_0a
from slowquant_b200.ucc_wavefunction import WaveFunctionUCC
from slowquant_b200.ups_wavefunction import WaveFunctionUPS


class LinearResponse(LinearResponseBaseClass):
    def __init__(self, wave_function: WaveFunctionUCC | WaveFunctionUPS, excitations: str) -> None:
        super().__init__(wave_function, excitations)
        wf = self.wf
        ci_info = wf.ci_info
        nq, nG = len(self.q_ops), len(self.G_ops)
        nI, nA = wf.num_inactive_orbs, wf.num_active_orbs
        print("Gs", nG)
        print("qs", nq)
        pn.orbital_blocks(self)
        E = wf.energy_elec
        psi = pn.state_on_device(wf.ci_coeffs, ci_info)
        H0 = pn.apply(self.H_0i_0a, psi, ci_info)
        Gk = pn.panel_from_operators(self.G_ops, psi, ci_info)
        self._Gk, self._psi = Gk, psi
        g0 = Gk @ psi                                                   # <0|G_I^d|0> = <G_I 0|0>
        gh = Gk @ H0                                                    # <0|G_I^d H|0>
        # <0|[H - E, G]|0>-type conditions of projected.py:54-80
        grad = (gh - E * g0).cpu().numpy()
        pn.check_active_gradient(np.concatenate([grad, grad]))
        if nq != 0 and nG != 0:
            H1 = SectorSplit(self.H_1i_1a, nI, nA)
            Hq = pn.panel_from_operators([H1.times(q) for q in self.q_ops], psi, ci_info)
            qdH = pn.panel_from_operators([H1.rtimes(q.dagger) for q in self.q_ops], psi, ci_info)
            A_Gq = pn.gram(Gk, Hq).cpu().numpy()
            B_Gq = -0.5 * pn.gram(Gk, qdH).cpu().numpy()
            self.A[nq:, :nq], self.A[:nq, nq:] = A_Gq, A_Gq.T
            self.B[nq:, :nq], self.B[:nq, nq:] = B_Gq, B_Gq.T
        if nG != 0:
            HG = pn.panel_from_rows(self.H_0i_0a, Gk, ci_info)
            hg = HG @ psi                                               # <0|H G_J|0>
            GG = pn.gram(Gk, Gk)
            g0g0 = torch.outer(g0, g0)
            A_GG = pn.gram(Gk, HG) + E * (g0g0 - GG) - 0.5 * (torch.outer(g0, hg) + torch.outer(gh, g0))
            B_GG = 0.5 * (torch.outer(gh, g0) + torch.outer(g0, gh)) - E * g0g0
            self.A[nq:, nq:] = pn.mirror_lower(A_GG)
            self.B[nq:, nq:] = pn.mirror_lower(B_GG)
            self.Sigma[nq:, nq:] = pn.mirror_lower(GG - g0g0)

    def get_transition_dipole(self) -> np.ndarray:
        """projected.py:292-396: sum_i (Z_i - Y_i) (<0|G_i|0><0|mu|0> - <0|G_i^d mu|0>) + orbital part."""
        wf = self.wf
        ci_info = wf.ci_info
        nI, nA = wf.num_inactive_orbs, wf.num_active_orbs
        number_excitations = len(self.excitation_energies)
        dev = self._psi.device
        ZmY = torch.from_numpy(np.ascontiguousarray((self.Z_G_normed - self.Y_G_normed).T)).to(dev)   # [n_exc, N_G]
        g0 = self._Gk @ self._psi
        dipole_integrals = wf.int_gen.electric_dipole
        transition_dipoles = np.zeros((number_excitations, 3))
        for axis in range(3):
            mu = one_electron_integral_transform(wf.c_mo, dipole_integrals[axis])
            mu_ket = pn.apply(one_elec_op_0i_0a(mu, nI, nA), self._psi, ci_info)
            exp_mu = torch.dot(self._psi, mu_ket)
            active = (ZmY @ (g0 * exp_mu - self._Gk @ mu_ket)).cpu().numpy()
            for s in range(number_excitations):
                transition_dipoles[s, axis] = pn.orbital_property_part(self, mu, s, number_excitations) + active[s]
        return transition_dipoles
