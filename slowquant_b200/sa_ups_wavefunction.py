"""State-averaged unitary product state wave function on the B200 engine.

Same constructor, properties and results as the reference's
slowquant/unitary_coupled_cluster/sa_ups_wavefunction.py (``WaveFunctionSAUPS``).  The states live on the device as
one ``[n_states, N_det]`` fp64 matrix; every state shares the ansatz unitary, so

* the state-averaged energy is n_states sigma builds and dots on resident vectors (sa_ups_wavefunction.py:496-515),
* the state-averaged 1-/2-RDMs are the mean of per-state ``sq_rdm12`` passes (the reference: ~n^4/4
  ``expectation_value_SA`` calls, :398-479),
* the theta gradient is the mean of per-state fused reverse sweeps (``sq_ups_grad_sweep``; the reference loops over
  operators with three `_SA` kernel calls each, :958-997),
* the subspace Hamiltonian and transition-property matrices are Gram matrices C (O C)^T (:741-769, :797-833).
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import Any

import numpy as np
import scipy.linalg
import torch

from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.density_matrix import get_orbital_gradient
from slowquant_b200.integral_manager import one_electron_integral_transform
from slowquant_b200.operators import hamiltonian_0i_0a, one_elec_op_0i_0a
from slowquant_b200.ups_wavefunction import WaveFunctionUPS, symmetrize_rdm2_like_reference

_SA_ANSATZE = ("tups", "qnp", "fucc", "ksafupccgsd", "safuccsd", "ksasdsfupccgsd")


class WaveFunctionSAUPS(WaveFunctionUPS):
    def __init__(
        self,
        cas: Sequence[int],
        mo_coeffs: np.ndarray,
        integral_generator,
        states: tuple[list[list[float]], list[list[str]]],
        ansatz: str,
        ansatz_options: dict[str, Any] | None = None,
        include_active_kappa: bool = False,
        device: int | None = None,
    ) -> None:
        """Arguments of sa_ups_wavefunction.py:40-50 (+ optional CUDA ``device``).  ``states`` = (weights, determinants):
        every state of the average is a combination of determinants written as occupation strings in the interleaved
        alpha0 beta0 alpha1 beta1 ... order (:221-236)."""
        options = dict(ansatz_options or {})
        a_low = ansatz.lower()
        if a_low not in _SA_ANSATZE:
            raise ValueError(f"Got unknown ansatz, {ansatz}")
        if a_low in ("tups", "qnp") and options.get("do_pp"):
            raise ValueError("perfect pairing is not supported for Ansatz in SA UPS wave functions.")
        if a_low == "safuccsd":
            options["SAS"] = True
            options["SAD"] = True
        # orbital spaces, kappa bookkeeping, CI space and ansatz layout are those of the single-state class
        super().__init__(cas, mo_coeffs, integral_generator, ansatz, options, include_active_kappa, device)
        self.kappa_idx_dagger = self.kappa_idx[:, ::-1].copy()
        self.num_states = len(states[0])
        self.csf_coeffs = np.zeros((self.num_states, self.num_det))
        for i, (coeffs, on_vecs) in enumerate(zip(states[0], states[1])):
            if len(coeffs) != len(on_vecs):
                raise ValueError(
                    f"Mismatch in number of coefficients, {len(coeffs)}, and number of determinants, {len(on_vecs)}. For {coeffs} and {on_vecs}"
                )
            for coeff, on_vec in zip(coeffs, on_vecs):
                if len(on_vec) != self.num_active_spin_orbs:
                    raise ValueError(
                        f"Length of determinant, {len(on_vec)}, does not match number of active spin orbitals, {self.num_active_spin_orbs}. For determinant, {on_vec}"
                    )
                self.csf_coeffs[i, self.ci_info.det2idx[int(on_vec, 2)]] = coeff
        overlap = self.csf_coeffs @ self.csf_coeffs.T
        for i in range(self.num_states):
            for j in range(self.num_states):
                if i == j:
                    if abs(1 - overlap[i, i]) > 10**-10:
                        raise ValueError(f"state {i} is not normalized got overlap of {overlap[i, i]}")
                elif abs(overlap[i, j]) > 10**-10:
                    raise ValueError(f"state {i} and {j} are not orthogonal got overlap of {overlap[i, j]}")
        dev = torch.device("cuda", self.ci_info.device)
        self._csf_dev = torch.from_numpy(self.csf_coeffs).to(dev)
        self._ci_dev = self._csf_dev.clone()
        self._ci_host = None
        self._sa_energy: float | None = None
        self._state_energies: np.ndarray | None = None
        self._state_ci_coeffs: np.ndarray | None = None

    # ---- parameters -----------------------------------------------------------------------------
    @property
    def kappa(self) -> list[float]:
        return self._kappa.copy()

    @kappa.setter
    def kappa(self, k: list[float]) -> None:
        """sa_ups_wavefunction.py:296-311."""
        self._h_mo = None
        self._g_mo = None
        self._sa_energy = None
        self._state_energies = None
        self._kappa = list(k)
        self._c_mo = self.c_mo
        self._kappa_old = self.kappa
        self._state_ci_coeffs = None

    @property
    def thetas(self) -> list[float]:
        return self._thetas.copy()

    @thetas.setter
    def thetas(self, theta_vals: list[float]) -> None:
        """Set ansatz parameters; all states are rebuilt on the device (sa_ups_wavefunction.py:338-353)."""
        if len(theta_vals) != len(self._thetas):
            raise ValueError(f"Expected {len(self._thetas)} theta1 values got {len(theta_vals)}")
        self._rdm1 = self._rdm2 = None
        self._sa_energy = None
        self._state_energies = None
        self._state_ci_coeffs = None
        self._thetas = [float(x) for x in theta_vals]
        self._ci_dev = osa.construct_ups_state_SA(self._csf_dev, self.ci_info, self._thetas, self.ups_layout)
        self._ci_host = None

    @property
    def ci_coeffs(self) -> np.ndarray:
        """[n_states, N_det] (sa_ups_wavefunction.py:313-327)."""
        if self._ci_host is None:
            self._ci_host = self._ci_dev.cpu().numpy()
        return self._ci_host

    # ---- densities and energies -------------------------------------------------------------------
    def _build_rdms(self, want_rdm2: bool) -> None:
        n = self.num_active_orbs
        d1 = np.zeros((n, n))
        d2 = np.zeros((n, n, n, n)) if want_rdm2 else None
        for s in range(self.num_states):
            a1, a2 = osa.reduced_density_matrices(self._ci_dev[s], self._ci_dev[s], self.ci_info, want_rdm2=want_rdm2)
            d1 += a1 / self.num_states
            if want_rdm2:
                d2 += a2 / self.num_states
        low = np.tril(d1)
        self._rdm1 = low + low.T - np.diag(np.diag(d1))
        if want_rdm2:
            self._rdm2 = symmetrize_rdm2_like_reference(d2)

    def _hamiltonian(self):
        return hamiltonian_0i_0a(self.h_mo, self.g_mo, self.num_inactive_orbs, self.num_active_orbs)

    def _state_energy_list(self) -> list[float]:
        H = self._hamiltonian()
        return [osa.expectation_value(self._ci_dev[s], [H], self._ci_dev[s], self.ci_info) for s in range(self.num_states)]

    @property
    def sa_energy(self) -> float:
        """State-averaged electronic energy (sa_ups_wavefunction.py:495-515)."""
        if self._sa_energy is None:
            self._sa_energy = float(sum(self._state_energy_list()) / self.num_states)
        return self._sa_energy

    @property
    def energy_elec(self) -> float:
        return self.sa_energy

    def _operator_matrix(self, op) -> np.ndarray:
        """M[i, j] = <ci_i| op |ci_j> as one Gram matrix of the resident states with the op-applied states."""
        applied = torch.stack([osa.propagate_state([op], self._ci_dev[s], self.ci_info) for s in range(self.num_states)])
        return (self._ci_dev @ applied.T).cpu().numpy()

    def _do_state_ci(self) -> None:
        """Subspace diagonalisation (sa_ups_wavefunction.py:741-769); j <= i evaluated and mirrored."""
        M = self._operator_matrix(self._hamiltonian())
        low = np.tril(M)
        state_H = low + np.tril(M, -1).T
        eigval, eigvec = scipy.linalg.eig(state_H)
        sorting = np.argsort(eigval)
        self._state_energies = np.real(eigval[sorting])
        self._state_ci_coeffs = np.real(eigvec[:, sorting])

    @property
    def energy_states(self) -> np.ndarray:
        if self._state_energies is None:
            self._do_state_ci()
        return self._state_energies

    @property
    def excitation_energies(self) -> np.ndarray:
        e = self.energy_states
        return np.asarray(e[1:] - e[0])

    def get_transition_property(self, ao_integral: np.ndarray) -> np.ndarray:
        r""":math:`t_n = \langle 0|\hat O|n\rangle` between the diagonalised states (sa_ups_wavefunction.py:797-833)."""
        if self._state_ci_coeffs is None:
            self._do_state_ci()
        mo_integral = one_electron_integral_transform(self.c_mo, ao_integral)
        op = one_elec_op_0i_0a(mo_integral, self.num_inactive_orbs, self.num_active_orbs)
        state_op = self._operator_matrix(op)
        V = self._state_ci_coeffs
        return np.array([V[:, i + 1] @ state_op @ V[:, 0] for i in range(self.num_states - 1)])

    def get_oscillator_strenghts(self) -> np.ndarray:
        """2/3 e_n |<0|mu|n>|^2 (sa_ups_wavefunction.py:835-853; the reference's spelling of the method name)."""
        dip = self.int_gen.electric_dipole
        t = np.array([self.get_transition_property(dip[k]) for k in range(3)])
        return 2 / 3 * self.excitation_energies * np.sum(t**2, axis=0)

    # ---- optimisation callables (sa_ups_wavefunction.py:855-1082) -----------------------------------
    def _calc_energy_optimization(self, parameters, theta_optimization: bool, kappa_optimization: bool, return_all_states: bool = False):
        if np.max(np.abs(np.array(self._old_opt_parameters) - np.array(parameters))) < 10**-14:
            return self._E_opt_old
        num_kappa = 0
        if kappa_optimization:
            num_kappa = len(self.kappa_idx)
            self.kappa = list(parameters[:num_kappa])
        if theta_optimization:
            self.thetas = list(parameters[num_kappa:])
        energies = np.array(self._state_energy_list())
        if return_all_states:
            self._E_opt_old = np.copy(energies)
            self._old_opt_parameters = np.copy(parameters)
            return energies
        E = float(np.sum(energies) / self.num_states)
        self._E_opt_old = E
        self._old_opt_parameters = np.copy(parameters)
        self.num_energy_evals += self.num_states
        return E

    def _calc_gradient_optimization(self, parameters, theta_optimization: bool, kappa_optimization: bool) -> np.ndarray:
        gradient = np.zeros(len(parameters))
        num_kappa = 0
        if kappa_optimization:
            num_kappa = len(self.kappa_idx)
            self.kappa = list(parameters[:num_kappa])
        if theta_optimization:
            self.thetas = list(parameters[num_kappa:])
        if kappa_optimization:
            gradient[:num_kappa] = get_orbital_gradient(
                self.h_mo, self.g_mo, self.kappa_idx, self.num_inactive_orbs, self.num_active_orbs, self.rdm1, self.rdm2
            )
        if theta_optimization:
            H = self._hamiltonian()
            for s in range(self.num_states):
                # backwards through the circuit from (H|psi_s>, |psi_s>): the numbers of the reference's forward loop without its
                # adjoint pass (operator_state_algebra.ups_gradient_sweep_backward)
                bra = osa.propagate_state([H], self._ci_dev[s], self.ci_info)
                g, _, _ = osa.ups_gradient_sweep_backward(bra, self._ci_dev[s], self.ci_info, self._thetas, self.ups_layout)
                gradient[num_kappa:] += g / self.num_states
            self.num_energy_evals += 2 * int(np.sum(list(self.ups_layout.grad_param_R.values()))) * self.num_states
        return gradient

    def _calc_energy_rotosolve_optimization(self, parameters: list[float], theta_diffs: list[float], theta_idx: int) -> np.ndarray:
        """Summed state energies at every shifted theta[theta_idx] (sa_ups_wavefunction.py:1005-1082: the reference
        returns the SUM over states, not the mean)."""
        th = np.asarray(parameters, dtype=np.float64).copy()
        n = len(th)
        H = self._hamiltonian()
        energies = np.zeros(len(theta_diffs))
        for s in range(self.num_states):
            prefix = self._csf_dev[s].clone()
            if theta_idx > 0:
                osa._ups_apply_inplace(prefix, self.ci_info, th, self.ups_layout, 0, theta_idx, False)
            kets = prefix.unsqueeze(0).repeat(len(theta_diffs), 1).contiguous()
            for j, shift in enumerate(theta_diffs):
                th_s = th.copy()
                th_s[theta_idx] = shift
                osa._ups_apply_inplace(kets[j], self.ci_info, th_s, self.ups_layout, theta_idx, theta_idx + 1, False)
            if theta_idx + 1 < n:   # common tail of all shifted states: one batched launch sequence
                osa._ups_apply_batch_inplace(kets, self.ci_info, th, self.ups_layout, theta_idx + 1, n, False)
            for j in range(len(theta_diffs)):
                energies[j] += osa._dot(osa.propagate_state([H], kets[j], self.ci_info), kets[j], self.ci_info)
        self.num_energy_evals += self.num_states
        return energies

    def run_wf_optimization_1step(self, optimizer_name: str, orbital_optimization: bool = False, tol: float = 1e-10, maxiter: int = 1000) -> None:
        """sa_ups_wavefunction.py:640-739."""
        if optimizer_name.lower() == "rotosolve" and orbital_optimization and len(self.kappa) != 0:
            raise ValueError("Cannot use RotoSolve together with orbital optimization in the one-step solver.")
        theta_opt = len(self.thetas) > 0 or not orbital_optimization
        parameters = (self.kappa if orbital_optimization else []) + (self.thetas if theta_opt else [])
        optimizer = self._optimizer(optimizer_name, theta_opt, orbital_optimization, tol, maxiter)
        self._old_opt_parameters = np.zeros(len(parameters)) + 10**20
        self._E_opt_old = 0.0
        res = optimizer.minimize(parameters, extra_options=self._rotosolve_options(optimizer_name))
        if orbital_optimization:
            self.thetas = res.x[len(self.kappa) :].tolist()
            for i in range(len(self._kappa)):
                self._kappa[i] = 0.0
                self._kappa_old[i] = 0.0
        else:
            self.thetas = res.x.tolist()
        self._do_state_ci()
        self._sa_energy = res.fun

    def _finish_optimization(self, energy: float) -> None:
        """After the shared two-step driver (ups_wavefunction.run_wf_optimization_2step): subspace diagonalisation, then
        the state-averaged energy of the optimiser (sa_ups_wavefunction.py:636-638)."""
        self._do_state_ci()
        self._sa_energy = energy
