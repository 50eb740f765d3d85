"""Non-factorised unitary coupled cluster wave function on the B200 engine.

Surface of the reference's ``WaveFunctionUCC`` (slowquant/unitary_coupled_cluster/ucc_wavefunction.py:35-1098)
for the state-vector path: constructor ``(cas, mo_coeffs, integral_generator, excitations,
include_active_kappa)``, ``thetas`` setter, lazily built ``ci_coeffs`` = exp(T - T^dagger)|CSF>, RDMs, energy,
orbital gradient and the forward finite-difference theta gradient of ucc_wavefunction.py:1062-1097.  The
exponential is applied matrix free (``ucc_state.py``) instead of through a dense N_det x N_det matrix.
"""
from __future__ import annotations

from collections.abc import Sequence

import numpy as np

from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.density_matrix import get_orbital_gradient
from slowquant_b200.operators import hamiltonian_0i_0a
from slowquant_b200.ucc_state import expm_multiply_operator, get_ucc_T
from slowquant_b200.ups_wavefunction import WaveFunctionUPS
from slowquant_b200.util import UccStructure


class WaveFunctionUCC(WaveFunctionUPS):
    def __init__(
        self,
        cas: Sequence[int],
        mo_coeffs: np.ndarray,
        integral_generator,
        excitations: str,
        include_active_kappa: bool = False,
        device: int | None = None,
    ) -> None:
        # orbital spaces, kappa bookkeeping, CI space and reference determinant are shared with the UPS class
        super().__init__(cas, mo_coeffs, integral_generator, "fucc", {"n_layers": 1, "S": True}, include_active_kappa, device)
        self._excitations = excitations
        exc = excitations.lower()
        self.ucc_layout = UccStructure()
        if "s" in exc:
            self.ucc_layout.add_sa_singles(self.active_occ_idx_shifted, self.active_unocc_idx_shifted)
        if "d" in exc:
            self.ucc_layout.add_sa_doubles(self.active_occ_idx_shifted, self.active_unocc_idx_shifted)
        if "t" in exc:
            self.ucc_layout.add_triples(self.active_occ_spin_idx_shifted, self.active_unocc_spin_idx_shifted)
        if "q" in exc:
            self.ucc_layout.add_quadruples(self.active_occ_spin_idx_shifted, self.active_unocc_spin_idx_shifted)
        if "5" in exc:
            self.ucc_layout.add_quintuples(self.active_occ_spin_idx_shifted, self.active_unocc_spin_idx_shifted)
        if "6" in exc:
            self.ucc_layout.add_sextuples(self.active_occ_spin_idx_shifted, self.active_unocc_spin_idx_shifted)
        self.ups_layout = None
        self._thetas = np.zeros(self.ucc_layout.n_params).tolist()
        self._old_opt_parameters = np.zeros(len(self._thetas) + len(self._kappa)) + 10**20

    @property
    def thetas(self) -> list[float]:
        return self._thetas.copy()

    @thetas.setter
    def thetas(self, theta: list[float]) -> None:
        if len(theta) != len(self._thetas):
            raise ValueError(f"Expected {len(self._thetas)} theta1 values got {len(theta)}")
        self._rdm1 = self._rdm2 = None
        self._rdm3 = self._rdm4 = None
        self._energy_elec = None
        self._thetas = [float(x) for x in theta]
        self._ci_dev = osa.construct_ucc_state(self._csf_dev, self.ci_info, self._thetas, self.ucc_layout)
        self._ci_host = None

    def _calc_gradient_optimization(self, parameters, theta_optimization: bool, kappa_optimization: bool) -> np.ndarray:
        """Orbital gradient from the RDMs; theta gradient by forward differences (ucc_wavefunction.py:1062-1097)."""
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
            H = hamiltonian_0i_0a(self.h_mo, self.g_mo, self.num_inactive_orbs, self.num_active_orbs)
            eps = np.finfo(np.float64).eps ** (1 / 2)
            Hket = osa.propagate_state([H], self._ci_dev, self.ci_info)
            E = osa._dot(self._ci_dev, Hket, self.ci_info)
            T0 = get_ucc_T(self._thetas, self.ucc_layout)
            theta_params = np.zeros(len(self._thetas))
            for i in range(len(theta_params)):
                step_size = eps  # theta_params[i] == 0: sign +1, max(1, 0) = 1
                theta_params[i] += step_size
                # exp(Tmat + Tmat_plus)|CSF>: the reference sums the two matrices (ucc_wavefunction.py:1093-1094)
                bra = expm_multiply_operator(T0 + get_ucc_T(theta_params, self.ucc_layout), self._csf_dev, self.ci_info)
                E_plus = osa._dot(bra, Hket, self.ci_info)
                theta_params[i] -= step_size
                gradient[i + num_kappa] = 2 * (E_plus - E) / step_size
        return gradient
