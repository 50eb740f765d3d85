"""Naive linear response (same class name, constructor and results as the reference's
slowquant/unitary_coupled_cluster/linear_response/naive.py), built from device panels.

Reference structure (naive.py:26-304): a Python double loop over operator pairs, every matrix element assembled
from ``propagate_state`` / ``expectation_value`` calls.  Here:

* the six panels  G|0>, G^d|0>, H G|0>, H G^d|0>, G H|0>, G^d H|0>  are built once (2 N_G gather launches and 2 N_G
  sigma builds instead of O(N_G^2) of each) and stay in HBM;
* the G-G blocks of A, B and Sigma are sums of Gram matrices of those panels (fp64 GEMMs over the determinant index);
* the q-G blocks keep the reference's symbolic route (operator products folded onto the active space on the host --
  an orbital rotation leaves the reference sector, so it cannot be a panel row), with the terms that factor through the
  reference sector taken from panels as well;
* the q-q blocks are the RDM contractions of density_matrix.py.
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
        pn.orbital_blocks(self)                                         # guard + q-q blocks (naive.py:42-54, 93-123)
        # ---- panels (device resident) ----
        psi = pn.state_on_device(wf.ci_coeffs, ci_info)
        H0 = pn.apply(self.H_0i_0a, psi, ci_info)                       # H|0>
        G_dag = [op.dagger for op in self.G_ops]
        Gk = pn.panel_from_operators(self.G_ops, psi, ci_info)         # G_I|0>
        Gdk = pn.panel_from_operators(G_dag, psi, ci_info)             # G_I^d|0>
        self._Gk, self._Gdk, self._psi = Gk, Gdk, psi
        # <0|[H, G]|0> and <0|[G^d, H]|0> must vanish at a converged wave function (naive.py:56-92)
        if nG != 0:
            hg = (Gk @ H0).cpu().numpy()
            hgd = (Gdk @ H0).cpu().numpy()
            pn.check_active_gradient(np.concatenate([hg - hgd, hg - hgd]))
        # ---- q-G blocks (naive.py:124-194) ----
        if nq != 0 and nG != 0:
            H1 = SectorSplit(self.H_1i_1a, nI, nA)   # only the strings of H that can survive the fold are multiplied
            Hq = pn.panel_from_operators([H1.times(q) for q in self.q_ops], psi, ci_info)               # H q_J|0>
            qdH = pn.panel_from_operators([H1.rtimes(q.dagger) for q in self.q_ops], psi, ci_info)      # q_J^d H|0>
            # terms that are overlaps of reference-sector vectors: one GEMM each
            A_Gq = pn.gram(Gk, Hq) - 0.5 * pn.gram(Gdk, qdH)           # <0|G^d H q|0> - 1/2 <0|H q G^d|0>
            B_Gq = pn.gram(Gdk, Hq) - 0.5 * pn.gram(Gk, qdH)           # <0|q^d H G^d|0> - 1/2 <0|G^d q^d H|0>
            # terms with the rotation next to |0>: the rotated state is outside the reference sector, so the
            # product is folded symbolically (as the reference does) and its expectation value taken on the device
            tmp = torch.empty_like(psi)
            A3 = np.zeros((nG, nq))
            B3 = np.zeros((nG, nq))
            for j, qJ in enumerate(self.q_ops):
                qJd = qJ.dagger
                for i, GId in enumerate(G_dag):
                    pn.apply_into(H1.times(GId * qJ), psi, tmp, ci_info)      # <0|H G^d q|0>
                    A3[i, j] = float(torch.dot(psi, tmp))
                    pn.apply_into(H1.rtimes(qJd * GId), psi, tmp, ci_info)    # <0|q^d G^d H|0>
                    B3[i, j] = float(torch.dot(psi, tmp))
            A_Gq = A_Gq.cpu().numpy() - 0.5 * A3
            B_Gq = B_Gq.cpu().numpy() - 0.5 * B3
            self.A[nq:, :nq] = A_Gq
            self.A[:nq, nq:] = A_Gq.T
            self.B[nq:, :nq] = B_Gq
            self.B[:nq, nq:] = B_Gq.T
        # ---- G-G blocks (naive.py:195-304) ----
        if nG != 0:
            HG = pn.panel_from_rows(self.H_0i_0a, Gk, ci_info)          # H G_J|0>
            HGd = pn.panel_from_rows(self.H_0i_0a, Gdk, ci_info)        # H G_J^d|0>
            GH = pn.panel_from_operators(self.G_ops, H0, ci_info)       # G_J H|0>
            GdH = pn.panel_from_operators(G_dag, H0, ci_info)           # G_J^d H|0>
            g = pn.gram
            A_GG = g(Gk, HG) + g(Gdk, HGd) - 0.5 * (g(Gk, GH) + g(Gdk, GdH) + g(GdH, Gdk) + g(GH, Gk))
            B_GG = g(Gk, HGd) - g(Gk, GdH) - g(Gdk, GH) + g(Gdk, HG)
            S_GG = g(Gk, Gk) - g(Gdk, Gdk)
            self.A[nq:, nq:] = pn.mirror_lower(A_GG)
            self.B[nq:, nq:] = pn.mirror_lower(B_GG)
            self.Sigma[nq:, nq:] = pn.mirror_lower(S_GG)

    def get_transition_dipole(self) -> np.ndarray:
        """<0|[mu, O_n]|0> for every excited state (naive.py:306-429); the active part for all states at once:
        the transfer states are two GEMMs of the response amplitudes with the G panels."""
        wf = self.wf
        ci_info = wf.ci_info
        nI, nA = wf.num_inactive_orbs, wf.num_active_orbs
        number_excitations = len(self.excitation_energies)
        dev = self._psi.device
        Z = torch.from_numpy(np.ascontiguousarray(self.Z_G_normed.T)).to(dev)   # [n_exc, N_G]
        Y = torch.from_numpy(np.ascontiguousarray(self.Y_G_normed.T)).to(dev)
        transfer = Z @ self._Gdk + Y @ self._Gk                                # (sum_i Z_i G_i^d + Y_i G_i)|0>
        transfer_d = Z @ self._Gk + Y @ self._Gdk                              # its adjoint on |0>
        dipole_integrals = wf.int_gen.electric_dipole
        transition_dipoles = np.zeros((number_excitations, 3))
        for axis in range(3):
            mu = one_electron_integral_transform(wf.c_mo, dipole_integrals[axis])
            mu_op = one_elec_op_0i_0a(mu, nI, nA)
            mu_ket = pn.apply(mu_op, self._psi, ci_info)
            mud_ket = pn.apply(mu_op.dagger, self._psi, ci_info)
            active = (transfer @ mud_ket - transfer_d @ mu_ket).cpu().numpy()  # <0|mu T|0> - <0|T mu|0>
            for state_number in range(number_excitations):
                q_part = pn.orbital_property_part(self, mu, state_number, number_excitations)
                transition_dipoles[state_number, axis] = q_part + active[state_number]
        return transition_dipoles
