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
        occ, unocc = self.wf.active_occ_idx, self.wf.active_unocc_idx
        occ_s, unocc_s = self.wf.active_occ_spin_idx, self.wf.active_unocc_spin_idx
        # excitation pools by the letters of `excitations`; the iterators yield (a, i, b, j, ...) = particle, hole, ...
        # and the operator factories take all holes first, then all particles (lr_baseclass.py:61-91)
        spin_orbital_pools = (("t", iterate_t3, G3), ("q", iterate_t4, G4), ("5", iterate_t5, G5), ("6", iterate_t6, G6))
        excitations = excitations.lower()
        self.G_ops: list[FermionicOperator] = []
        if "s" in excitations:
            self.G_ops += [G1_sa(i, a) for a, i, _ in iterate_t1_sa(occ, unocc)]
        if "d" in excitations:
            self.G_ops += [G2_sa(i, j, a, b, case) for a, i, b, j, _, case in iterate_t2_sa(occ, unocc)]
        for letter, iterator, factory in spin_orbital_pools:
            if letter in excitations:
                for idx in iterator(occ_s, unocc_s):
                    self.G_ops.append(factory(*idx[1::2], *idx[0::2]))
        self.q_ops: list[FermionicOperator] = [G1_sa(int(p), int(q)) for p, q in self.wf.kappa_no_activeactive_idx]
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
        self.num_q = nq = len(self.q_ops)
        self.num_G = nG = size - nq
        # rows of a response vector: [Z_q, Z_G, Y_q, Y_G]
        bounds = np.cumsum([0, nq, nG, nq, nG])
        self.Z_q, self.Z_G, self.Y_q, self.Y_G = (self.response_vectors[bounds[k] : bounds[k + 1]] for k in range(4))
        norms = self.get_excited_state_norm()
        scale = np.zeros_like(norms)
        for state_number, norm in enumerate(norms):
            if norm < 10**-10:    # such a state keeps zero normalised vectors (lr_baseclass.py:150-153)
                print(f"WARNING: State number {state_number} could not be normalized. Norm of {norm}.")
            else:
                scale[state_number] = (1 / norm) ** 0.5
        self.normed_response_vectors = self.response_vectors * scale[None, :]
        self.Z_q_normed, self.Z_G_normed = self.Z_q * scale[None, :], self.Z_G * scale[None, :]
        self.Y_q_normed, self.Y_G_normed = self.Y_q * scale[None, :], self.Y_G * scale[None, :]

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
