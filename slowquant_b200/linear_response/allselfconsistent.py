"""All-self-consistent linear response (same class name and results as the reference's
slowquant/unitary_coupled_cluster/linear_response/allselfconsistent.py), built from panels in the extended CI space.

Orbital rotations q = 2^{-1/2} E_ai (Hartree-Fock-like pairs) and active-space excitations G both act on the reference
(CSF) state.  With O = q or G, OC[k] = O_k|CSF>, and the images of H|0> under U^d (H2 = hamiltonian_2i_2a, H1 =
hamiltonian_1i_1a, H0 = hamiltonian_0i_0a; operators that leave the space are applied with do_unsafe semantics), the
blocks of allselfconsistent.py:127-262 are Gram matrices:

    A_qq = QC Xq^T - 1/2 (QC qUdH2^T + qUdH2 QC^T),   B_qq = -QC qdUdH2^T,      Xq[J] = U^d H2 U q_J|CSF>
    A_Gq = GC (U^d H1 U q|CSF>)^T,                    B_Gq = -1/2 (GC (q^d U^d H1|0>)^T + (G^d U^d H2|0>) QC^T)
    A_GG = GC XG^T - 1/2 (GC GUdH0^T + GUdH0 GC^T),   B_GG = -GC GdUdH0^T,      XG[J] = U^d H0 U G_J|CSF>
    Sigma = 1
"""
from __future__ import annotations

import numpy as np
import torch

from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.ci_spaces import get_indexing_extended
from slowquant_b200.linear_response import _panels as pn
from slowquant_b200.linear_response.allstatetransfer import LinearResponse as _AllStateTransfer
from slowquant_b200.linear_response.lr_baseclass import LinearResponseBaseClass
from slowquant_b200.operators import Epq, hamiltonian_2i_2a
from slowquant_b200.ucc_wavefunction import WaveFunctionUCC
from slowquant_b200.ups_wavefunction import WaveFunctionUPS


class LinearResponse(LinearResponseBaseClass):
    def __init__(self, wave_function: WaveFunctionUCC | WaveFunctionUPS, excitations: str) -> None:
        super().__init__(wave_function, excitations)
        wf = self.wf
        nI, nA, nV = wf.num_inactive_orbs, wf.num_active_orbs, wf.num_virtual_orbs
        ci_ext = get_indexing_extended(nI, nA, nV, wf.num_active_elec_alpha, wf.num_active_elec_beta, 2, device=wf.ci_info.device)  # order 2 (allselfconsistent.py:45-52)
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
        self.q_ops = [2 ** (-1 / 2) * Epq(int(a), int(i)) for i, a in wf.kappa_hf_like_idx]
        nq, nG = len(self.q_ops), len(self.G_ops)
        num_parameters = nq + nG
        self.A = np.zeros((num_parameters, num_parameters))
        self.B = np.zeros((num_parameters, num_parameters))
        self.Sigma = np.zeros((num_parameters, num_parameters))
        self.Delta = np.zeros((num_parameters, num_parameters))
        H2, H1, H0 = hamiltonian_2i_2a(wf.h_mo, wf.g_mo, nI, nA, nV), self.H_1i_1a, self.H_0i_0a
        print("Gs", nG)
        print("qs", nq)
        print("WARNING!")
        print("Gradient working equations not implemented for state transfer q operators")

        def panel(ops_lists, src, unsafe):
            rows = [osa.propagate_state(ops, src, *ext, do_unsafe=unsafe) for ops in ops_lists]
            return torch.stack(rows) if rows else torch.zeros((0, ci_ext.num_det), dtype=torch.float64, device=dev)

        G_dag = [G.dagger for G in self.G_ops]
        q_dag = [q.dagger for q in self.q_ops]
        UdH0 = osa.propagate_state(["Ud", H0], psi, *ext)                                     # U^d H0|0>
        UdH2 = osa.propagate_state(["Ud", H2], psi, *ext, do_unsafe=True)                     # U^d H2|0>
        QC = panel([[q] for q in self.q_ops], csf, True)
        GC = panel([[G] for G in self.G_ops], csf, False)
        self._QC, self._GC = QC, GC
        if nG != 0:
            gh = (GC @ UdH0).cpu().numpy()
            pn.check_active_gradient(np.concatenate([-gh, gh]))
        g = pn.gram
        if nq != 0:
            Xq = panel([["Ud", H2, "U", q] for q in self.q_ops], csf, True)
            qUdH2 = panel([[q] for q in self.q_ops], UdH2, True)
            qdUdH2 = panel([[qd] for qd in q_dag], UdH2, True)
            self.A[:nq, :nq] = pn.mirror_lower(g(QC, Xq) - 0.5 * (g(QC, qUdH2) + g(qUdH2, QC)))
            self.B[:nq, :nq] = pn.mirror_lower(-g(QC, qdUdH2))
            self.Sigma[:nq, :nq] = np.eye(nq)
        if nq != 0 and nG != 0:
            UdH1Uq = panel([["Ud", H1, "U", q] for q in self.q_ops], csf, True)
            qdUdH1 = panel([[qd, "Ud", H1] for qd in q_dag], psi, True)
            GdUdH2 = panel([[Gd] for Gd in G_dag], UdH2, True)
            A_Gq = g(GC, UdH1Uq).cpu().numpy()
            B_Gq = (-0.5 * (g(GC, qdUdH1) + g(GdUdH2, QC))).cpu().numpy()
            self.A[nq:, :nq], self.A[:nq, nq:] = A_Gq, A_Gq.T
            self.B[nq:, :nq], self.B[:nq, nq:] = B_Gq, B_Gq.T
        if nG != 0:
            XG = panel([["Ud", H0, "U", G] for G in self.G_ops], csf, False)
            GUdH0 = panel([[G] for G in self.G_ops], UdH0, False)
            GdUdH0 = panel([[Gd] for Gd in G_dag], UdH0, False)
            self.A[nq:, nq:] = pn.mirror_lower(g(GC, XG) - 0.5 * (g(GC, GUdH0) + g(GUdH0, GC)))
            self.B[nq:, nq:] = pn.mirror_lower(-g(GC, GdUdH0))
            self.Sigma[nq:, nq:] = np.eye(nG)

    # the transition dipole has the working equations of the all-state-transfer parametrisation (allselfconsistent.py:264-406)
    get_transition_dipole = _AllStateTransfer.get_transition_dipole
