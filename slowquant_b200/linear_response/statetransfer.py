"""State-transfer linear response (same class name and results as the reference's
slowquant/unitary_coupled_cluster/linear_response/statetransfer.py), built from device panels.

The excitation operators act on the reference (CSF) state and are carried to the correlated state by the ansatz
unitary.  With GC[I] = G_I|CSF> and X[J] = U^d H U G_J|CSF> -- every row of X costs two passes of the fused
unitary kernels and one sigma build -- the G-G block of statetransfer.py:141-162 is  A = GC X^T - E 1  (lower triangle
mirrored), Sigma = 1, B = 0; the q-G blocks are GC (U^d H q|0>)^T and -1/2 GC (U^d q^d H|0>)^T (:118-140).
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
        psi = pn.state_on_device(wf.ci_coeffs, ci_info)
        csf = pn.state_on_device(wf.csf_coeffs, ci_info)
        UdH0 = pn.apply(self.H_0i_0a, psi, ci_info)                     # U^d H|0>
        pn.apply_unitary_rows(UdH0, self.index_info, dagger=True)
        GC = pn.panel_from_operators(self.G_ops, csf, ci_info)          # G_I|CSF>
        self._GC, self._psi = GC, psi
        if nG != 0:
            gh = (GC @ UdH0).cpu().numpy()                              # <CSF|G_I^d U^d H|0>
            pn.check_active_gradient(np.concatenate([-gh, gh]))
        if nq != 0 and nG != 0:
            H1 = SectorSplit(self.H_1i_1a, nI, nA)
            UdHq = pn.panel_from_operators([H1.times(q) for q in self.q_ops], psi, ci_info)
            UdqdH = pn.panel_from_operators([H1.rtimes(q.dagger) for q in self.q_ops], psi, ci_info)
            pn.apply_unitary_rows(UdHq, self.index_info, dagger=True)
            pn.apply_unitary_rows(UdqdH, self.index_info, dagger=True)
            A_Gq = pn.gram(GC, UdHq).cpu().numpy()
            B_Gq = -0.5 * pn.gram(GC, UdqdH).cpu().numpy()
            self.A[nq:, :nq], self.A[:nq, nq:] = A_Gq, A_Gq.T
            self.B[nq:, :nq], self.B[:nq, nq:] = B_Gq, B_Gq.T
        if nG != 0:
            X = GC.clone()
            pn.apply_unitary_rows(X, self.index_info, dagger=False)     # U G_J|CSF>
            X = pn.panel_from_rows(self.H_0i_0a, X, ci_info)            # H U G_J|CSF>
            pn.apply_unitary_rows(X, self.index_info, dagger=True)      # U^d H U G_J|CSF>
            A_GG = pn.gram(GC, X) - wf.energy_elec * torch.eye(nG, dtype=torch.float64, device=GC.device)
            self.A[nq:, nq:] = pn.mirror_lower(A_GG)
            self.Sigma[nq:, nq:] = np.eye(nG)

    def get_transition_dipole(self) -> np.ndarray:
        """statetransfer.py:164-279: -Z_i <0|mu U G_i|CSF> + Y_i <CSF|G_i^d U^d mu|0> + orbital part."""
        wf = self.wf
        ci_info = wf.ci_info
        nI, nA = wf.num_inactive_orbs, wf.num_active_orbs
        number_excitations = len(self.excitation_energies)
        dev = self._psi.device
        Z = torch.from_numpy(np.ascontiguousarray(self.Z_G_normed.T)).to(dev)
        Y = torch.from_numpy(np.ascontiguousarray(self.Y_G_normed.T)).to(dev)
        dipole_integrals = wf.int_gen.electric_dipole
        n_states = len(self.normed_response_vectors[0])
        transition_dipoles = np.zeros((n_states, 3))
        for axis in range(3):
            mu = one_electron_integral_transform(wf.c_mo, dipole_integrals[axis])
            mu_op = one_elec_op_0i_0a(mu, nI, nA)
            Ud_mud = pn.apply(mu_op.dagger, self._psi, ci_info)
            Ud_mu = pn.apply(mu_op, self._psi, ci_info)
            pn.apply_unitary_rows(Ud_mud, self.index_info, dagger=True)
            pn.apply_unitary_rows(Ud_mu, self.index_info, dagger=True)
            active = (-Z @ (self._GC @ Ud_mud) + Y @ (self._GC @ Ud_mu)).cpu().numpy()
            for s in range(n_states):
                transition_dipoles[s, axis] = pn.orbital_property_part(self, mu, s, number_excitations) + active[s]
        return transition_dipoles
