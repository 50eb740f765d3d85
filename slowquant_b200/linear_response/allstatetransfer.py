"""All-state-transfer linear response (same class name and results as the reference's
slowquant/unitary_coupled_cluster/linear_response/allstatetransfer.py), built from panels in the extended CI space.

Both operator pools -- orbital rotations q = 2^{-1/2} E_ai over the Hartree-Fock-like pairs and active-space excitations G
-- act on the reference (CSF) state and are carried by the ansatz unitary, so everything is one panel
O[k] = O_k|CSF> and its image X[k] = U^d H U O_k|CSF> (H = hamiltonian_2i_2a for the q rows, which leave the CAS;
hamiltonian_0i_0a for the G rows); allstatetransfer.py:126-182 is then

    A = O X^T - E 1   (lower triangle mirrored, G-q block taken from the q images),   Sigma = 1,   B = 0.
"""
from __future__ import annotations

import numpy as np
import torch

from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.ci_spaces import get_indexing_extended
from slowquant_b200.integral_manager import one_electron_integral_transform
from slowquant_b200.linear_response import _panels as pn
from slowquant_b200.linear_response.lr_baseclass import LinearResponseBaseClass
from slowquant_b200.operators import Epq, hamiltonian_2i_2a, one_elec_op_0i_0a, one_elec_op_1i_1a
from slowquant_b200.ucc_wavefunction import WaveFunctionUCC
from slowquant_b200.ups_wavefunction import WaveFunctionUPS


class LinearResponse(LinearResponseBaseClass):
    def __init__(self, wave_function: WaveFunctionUCC | WaveFunctionUPS, excitations: str) -> None:
        super().__init__(wave_function, excitations)
        wf = self.wf
        nI, nA, nV = wf.num_inactive_orbs, wf.num_active_orbs, wf.num_virtual_orbs
        ci_ext = get_indexing_extended(nI, nA, nV, wf.num_active_elec_alpha, wf.num_active_elec_beta, 1, device=wf.ci_info.device)
        layout = wf.ucc_layout if isinstance(wf, WaveFunctionUCC) else wf.ups_layout
        self.index_info_extended = ext = (ci_ext, wf.thetas, layout)
        dev = torch.device("cuda", ci_ext.device)
        hf_det = int("1" * wf.int_gen.num_elec + "0" * (wf.num_spin_orbs - wf.int_gen.num_elec), 2)
        csf = torch.zeros(ci_ext.num_det, dtype=torch.float64, device=dev)
        csf[ci_ext.det2idx[hf_det]] = 1.0
        self.csf_coeffs = csf.cpu().numpy()
        psi = osa.propagate_state(["U"], csf, *ext)
        self.ci_coeffs = psi.cpu().numpy()
        self._csf, self._psi = csf, psi
        # the orbital pool of this parametrisation (allstatetransfer.py:72-75) replaces the base-class one
        self.q_ops = [2 ** (-1 / 2) * Epq(int(a), int(i)) for i, a in wf.kappa_hf_like_idx]
        nq, nG = len(self.q_ops), len(self.G_ops)
        num_parameters = nq + nG
        self.A = np.zeros((num_parameters, num_parameters))
        self.B = np.zeros((num_parameters, num_parameters))
        self.Sigma = np.zeros((num_parameters, num_parameters))
        self.Delta = np.zeros((num_parameters, num_parameters))
        H_2i_2a = hamiltonian_2i_2a(wf.h_mo, wf.g_mo, nI, nA, nV)
        print("Gs", nG)
        print("qs", nq)
        print("WARNING!")
        print("Gradient working equations not implemented for state transfer q operators")

        def panel(ops_lists, src, unsafe):
            rows = [osa.propagate_state(ops, src, *ext, do_unsafe=unsafe) for ops in ops_lists]
            return torch.stack(rows) if rows else torch.zeros((0, ci_ext.num_det), dtype=torch.float64, device=dev)

        UdH0 = osa.propagate_state(["Ud", self.H_0i_0a], psi, *ext)
        QC = panel([[q] for q in self.q_ops], csf, True)                  # q_I|CSF>
        GC = panel([[G] for G in self.G_ops], csf, False)                 # G_I|CSF>
        self._QC, self._GC = QC, GC
        if nG != 0:
            gh = (GC @ UdH0).cpu().numpy()
            pn.check_active_gradient(np.concatenate([-gh, gh]))
        E = wf.energy_elec
        if nq != 0:
            Xq = panel([["Ud", H_2i_2a, "U", q] for q in self.q_ops], csf, True)   # U^d H U q_J|CSF>
            self.A[:nq, :nq] = pn.mirror_lower(pn.gram(QC, Xq) - E * torch.eye(nq, dtype=torch.float64, device=dev))
            self.Sigma[:nq, :nq] = np.eye(nq)
            if nG != 0:
                A_Gq = pn.gram(GC, Xq).cpu().numpy()
                self.A[nq:, :nq], self.A[:nq, nq:] = A_Gq, A_Gq.T
        if nG != 0:
            XG = panel([["Ud", self.H_0i_0a, "U", G] for G in self.G_ops], csf, False)
            self.A[nq:, nq:] = pn.mirror_lower(pn.gram(GC, XG) - E * torch.eye(nG, dtype=torch.float64, device=dev))
            self.Sigma[nq:, nq:] = np.eye(nG)

    def get_transition_dipole(self) -> np.ndarray:
        """allstatetransfer.py:184-308: -Z_k <0|mu U O_k|CSF> + Y_k <CSF|O_k^d U^d mu|0> over both pools (mu restricted
        to one inactive/virtual change for the q rows, to the active space for the G rows)."""
        wf = self.wf
        ext = self.index_info_extended
        nI, nA, nV = wf.num_inactive_orbs, wf.num_active_orbs, wf.num_virtual_orbs
        dev = self._psi.device

        def amp(M):
            return torch.from_numpy(np.ascontiguousarray(M.T)).to(dev)

        Zq, Yq, ZG, YG = amp(self.Z_q_normed), amp(self.Y_q_normed), amp(self.Z_G_normed), amp(self.Y_G_normed)
        dipole_integrals = wf.int_gen.electric_dipole
        n_states = len(self.normed_response_vectors[0])
        transition_dipoles = np.zeros((n_states, 3))
        for axis in range(3):
            mu = one_electron_integral_transform(wf.c_mo, dipole_integrals[axis])
            mu_G = one_elec_op_0i_0a(mu, nI, nA)
            mu_q = one_elec_op_1i_1a(mu, nI, nA, nV)
            total = torch.zeros(n_states, dtype=torch.float64, device=dev)
            for Z, Y, O, mu_op, unsafe in ((Zq, Yq, self._QC, mu_q, True), (ZG, YG, self._GC, mu_G, False)):
                if O.shape[0] == 0:
                    continue
                left = osa.propagate_state(["Ud", mu_op.dagger], self._psi, *ext, do_unsafe=unsafe)   # <0|mu U ...
                right = osa.propagate_state(["Ud", mu_op], self._psi, *ext, do_unsafe=unsafe)         # ... U^d mu|0>
                total += -Z @ (O @ left) + Y @ (O @ right)
            transition_dipoles[:, axis] = total.cpu().numpy()
        return transition_dipoles
