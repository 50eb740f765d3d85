"""Shared part of the linear-response drivers: operator pools, the generalised eigenproblem, norms and
oscillator strengths.  Same attribute and method names as the reference's
slowquant/unitary_coupled_cluster/linear_response/lr_baseclass.py (cited per method); the matrices are filled by
the subclasses from device panels (see _panels.py).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from slowquant_b200.fermionic_operator import FermionicOperator
from slowquant_b200.operators import G3, G4, G5, G6, G1_sa, G2_sa, hamiltonian_0i_0a, hamiltonian_1i_1a
from slowquant_b200.ucc_wavefunction import WaveFunctionUCC
from slowquant_b200.ups_wavefunction import WaveFunctionUPS
from slowquant_b200.util import iterate_t1_sa, iterate_t2_sa, iterate_t3, iterate_t4, iterate_t5, iterate_t6


class LinearResponseBaseClass:
    def __init__(self, wave_function: WaveFunctionUCC | WaveFunctionUPS, excitations: str) -> None:
        """Operator pools G (active space) and q (orbital rotations), empty A/B/Sigma/Delta, the two
        Hamiltonians (lr_baseclass.py:33-113)."""
        self.wf = wave_function
        if isinstance(self.wf, WaveFunctionUCC):
            self.index_info = (self.wf.ci_info, self.wf.thetas, self.wf.ucc_layout)
        elif isinstance(self.wf, WaveFunctionUPS):
            self.index_info = (self.wf.ci_info, self.wf.thetas, self.wf.ups_layout)
        else:
            raise ValueError(f"Got incompatible wave function type, {type(self.wf)}")
        self.G_ops: list[FermionicOperator] = []
        self.q_ops: list[FermionicOperator] = []
        excitations = excitations.lower()
        occ, unocc = self.wf.active_occ_idx, self.wf.active_unocc_idx
        occ_s, unocc_s = self.wf.active_occ_spin_idx, self.wf.active_unocc_spin_idx
        if "s" in excitations:
            for a, i, _ in iterate_t1_sa(occ, unocc):
                self.G_ops.append(G1_sa(i, a))
        if "d" in excitations:
            for a, i, b, j, _, op_type in iterate_t2_sa(occ, unocc):
                self.G_ops.append(G2_sa(i, j, a, b, op_type))
        if "t" in excitations:
            for a, i, b, j, c, k in iterate_t3(occ_s, unocc_s):
                self.G_ops.append(G3(i, j, k, a, b, c))
        if "q" in excitations:
            for a, i, b, j, c, k, d, l in iterate_t4(occ_s, unocc_s):
                self.G_ops.append(G4(i, j, k, l, a, b, c, d))
        if "5" in excitations:
            for a, i, b, j, c, k, d, l, e, m in iterate_t5(occ_s, unocc_s):
                self.G_ops.append(G5(i, j, k, l, m, a, b, c, d, e))
        if "6" in excitations:
            for a, i, b, j, c, k, d, l, e, m, f, n in iterate_t6(occ_s, unocc_s):
                self.G_ops.append(G6(i, j, k, l, m, n, a, b, c, d, e, f))
        for p, q in self.wf.kappa_no_activeactive_idx:
            self.q_ops.append(G1_sa(int(p), int(q)))
        num_parameters = len(self.G_ops) + len(self.q_ops)
        self.A = np.zeros((num_parameters, num_parameters))
        self.B = np.zeros((num_parameters, num_parameters))
        self.Sigma = np.zeros((num_parameters, num_parameters))
        self.Delta = np.zeros((num_parameters, num_parameters))
        nI, nA, nV = self.wf.num_inactive_orbs, self.wf.num_active_orbs, self.wf.num_virtual_orbs
        self.H_1i_1a = hamiltonian_1i_1a(self.wf.h_mo, self.wf.g_mo, nI, nA, nV)
        self.H_0i_0a = hamiltonian_0i_0a(self.wf.h_mo, self.wf.g_mo, nI, nA)

    def calc_excitation_energies(self) -> None:
        """Solve E2 X = w S X (lr_baseclass.py:115-166): positive half of the spectrum, response vectors split into
        Z_q, Z_G, Y_q, Y_G and normalised with the metric."""
        size = len(self.A)
        E2 = np.block([[self.A, self.B], [self.B, self.A]])
        hess_eigval = np.real(np.linalg.eigvals(E2))  # E2 is symmetric; rounding can leave 1e-17 imaginary parts
        print(f"Smallest Hessian eigenvalue: {np.min(hess_eigval)}")
        if np.abs(np.min(hess_eigval)) < 10**-8:
            print("WARNING: Small eigenvalue in Hessian")
        elif np.min(hess_eigval) < 0:
            raise ValueError("Negative eigenvalue in Hessian.")
        S = np.block([[self.Sigma, self.Delta], [-self.Delta, -self.Sigma]])
        print(f"Smallest diagonal element in the metric: {np.min(np.abs(np.diagonal(self.Sigma)))}")
        self.hessian = E2
        self.metric = S
        eigval, eigvec = scipy.linalg.eig(self.hessian, self.metric)
        sorting = np.argsort(eigval)
        self.excitation_energies = np.real(eigval[sorting][size:])
        self.response_vectors = np.real(eigvec[:, sorting][:, size:])
        self.normed_response_vectors = np.zeros_like(self.response_vectors)
        self.num_q = len(self.q_ops)
        self.num_G = size - self.num_q
        nq, nG = self.num_q, self.num_G
        self.Z_q = self.response_vectors[:nq, :]
        self.Z_G = self.response_vectors[nq : nq + nG, :]
        self.Y_q = self.response_vectors[nq + nG : 2 * nq + nG]
        self.Y_G = self.response_vectors[2 * nq + nG :]
        self.Z_q_normed = np.zeros_like(self.Z_q)
        self.Z_G_normed = np.zeros_like(self.Z_G)
        self.Y_q_normed = np.zeros_like(self.Y_q)
        self.Y_G_normed = np.zeros_like(self.Y_G)
        norms = self.get_excited_state_norm()
        for state_number, norm in enumerate(norms):
            if norm < 10**-10:
                print(f"WARNING: State number {state_number} could not be normalized. Norm of {norm}.")
                continue
            scale = (1 / norm) ** 0.5
            self.Z_q_normed[:, state_number] = self.Z_q[:, state_number] * scale
            self.Z_G_normed[:, state_number] = self.Z_G[:, state_number] * scale
            self.Y_q_normed[:, state_number] = self.Y_q[:, state_number] * scale
            self.Y_G_normed[:, state_number] = self.Y_G[:, state_number] * scale
            self.normed_response_vectors[:, state_number] = self.response_vectors[:, state_number] * scale

    def get_excited_state_norm(self) -> np.ndarray:
        """Z^T S Z - Y^T S Y with the q-q and G-G diagonal blocks of the metric (lr_baseclass.py:168-188)."""
        nq, nG = self.num_q, self.num_G
        S_qq = self.metric[:nq, :nq]
        S_GG = self.metric[nq : nq + nG, nq : nq + nG]
        norms = np.einsum("is,ij,js->s", self.Z_q, S_qq, self.Z_q) - np.einsum("is,ij,js->s", self.Y_q, S_qq, self.Y_q)
        norms += np.einsum("is,ij,js->s", self.Z_G, S_GG, self.Z_G) - np.einsum("is,ij,js->s", self.Y_G, S_GG, self.Y_G)
        return norms

    def get_transition_dipole(self) -> np.ndarray:
        raise NotImplementedError

    def get_oscillator_strength(self) -> np.ndarray:
        r""":math:`f_n = \tfrac23 e_n |\langle 0|\hat\mu|n\rangle|^2` (lr_baseclass.py:198-220)."""
        transition_dipoles = self.get_transition_dipole()
        osc_strs = 2 / 3 * self.excitation_energies[: len(transition_dipoles)] * np.sum(np.asarray(transition_dipoles) ** 2, axis=1)
        self.oscillator_strengths = osc_strs
        return osc_strs

    def get_formatted_oscillator_strength(self) -> str:
        """Table of excitation energies and oscillator strengths (lr_baseclass.py:222-240)."""
        if not hasattr(self, "oscillator_strengths"):
            raise ValueError("Oscillator strengths have not been calculated. Run get_oscillator_strength() first.")
        output = "Excitation # | Excitation energy [Hartree] | Excitation energy [eV] | Oscillator strengths\n"
        for i, (exc_energy, osc_strength) in enumerate(zip(self.excitation_energies, self.oscillator_strengths)):
            exc_str = f"{exc_energy:2.6f}"
            exc_str_ev = f"{exc_energy * 27.2114079527:3.6f}"
            osc_str = f"{osc_strength:1.6f}"
            output += f"{str(i + 1).center(12)} | {exc_str.center(27)} | {exc_str_ev.center(22)} | {osc_str.center(20)}\n"
        return output
