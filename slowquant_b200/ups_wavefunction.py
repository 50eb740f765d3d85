"""Unitary product state wave function on the B200 engine.

Keeps the constructor and property surface of the reference's ``WaveFunctionUPS``
(slowquant/unitary_coupled_cluster/ups_wavefunction.py:38-1194): ``thetas`` / ``kappa`` setters that
invalidate caches, ``ci_coeffs``, ``c_mo``, ``h_mo``, ``g_mo``, ``rdm1``, ``rdm2``, ``energy_elec``,
``_calc_energy_optimization`` / ``_calc_gradient_optimization`` and ``run_wf_optimization_1step``.  The CI
vector lives on the device; numpy views are made on demand.
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import Any

import numpy as np
import scipy.linalg
import scipy.optimize
import torch

from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.ci_spaces import get_indexing
from slowquant_b200.density_matrix import get_electronic_energy, get_orbital_gradient
from slowquant_b200.integral_manager import (
    IntegralManager,
    one_electron_integral_transform,
    two_electron_integral_transform,
)
from slowquant_b200.operators import hamiltonian_0i_0a
from slowquant_b200.util import UpsStructure


def symmetrize_rdm2_like_reference(full: np.ndarray) -> np.ndarray:
    """Keep the unique quadruples the reference evaluates and copy each to its 4 symmetric positions
    (ups_wavefunction.py:446-475), so symmetric partners are bit-identical as they are there."""
    n = full.shape[0]
    out = np.zeros_like(full)
    for p in range(n):
        for q in range(p + 1):
            for r in range(p + 1):
                if p == q:
                    s_lim = r + 1
                elif p == r:
                    s_lim = q + 1
                elif q < r:
                    s_lim = p
                else:
                    s_lim = p + 1
                for s in range(s_lim):
                    val = full[p, q, r, s]
                    out[p, q, r, s] = val
                    out[r, s, p, q] = val
                    out[q, p, s, r] = val
                    out[s, r, q, p] = val
    return out


class WaveFunctionUPS:
    def __init__(
        self,
        cas: Sequence[int],
        mo_coeffs: np.ndarray,
        integral_generator,
        ansatz: str,
        ansatz_options: dict[str, Any] | None = None,
        include_active_kappa: bool = False,
        device: int | None = None,
    ) -> None:
        """Same arguments as ups_wavefunction.py:39-47 (+ optional CUDA ``device``)."""
        if ansatz_options is None:
            ansatz_options = {}
        if len(cas) != 2:
            raise ValueError(f"cas must have two elements, got {len(cas)} elements.")
        self.int_gen = IntegralManager(integral_generator)
        self.num_orbs = len(self.int_gen.h_ao)
        self.num_spin_orbs = 2 * self.num_orbs
        self.ansatz_options = dict(ansatz_options)
        self.num_energy_evals = 0
        self._include_active_kappa = include_active_kappa
        self._rdm1 = self._rdm2 = None
        self._h_mo = self._g_mo = None
        self._energy_elec: float | None = None
        # ---- orbital spaces (ups_wavefunction.py:95-173) ----
        num_elec = self.int_gen.num_elec
        n_act_elec, n_act_orbs = int(cas[0]), int(cas[1])
        active_space = list(range(num_elec - n_act_elec, num_elec))
        active_space += list(range(num_elec, num_elec + 2 * n_act_orbs - len(active_space)))
        active_set = set(active_space)
        self.inactive_spin_idx = [i for i in range(num_elec) if i not in active_set]
        self.active_occ_spin_idx = [i for i in range(num_elec) if i in active_set]
        self.active_unocc_spin_idx = [i for i in range(num_elec, self.num_spin_orbs) if i in active_set]
        self.virtual_spin_idx = [i for i in range(num_elec, self.num_spin_orbs) if i not in active_set]
        self.active_spin_idx = self.active_occ_spin_idx + self.active_unocc_spin_idx
        self.num_active_elec = len(self.active_occ_spin_idx)
        self.num_active_spin_orbs = len(self.active_spin_idx)
        self.num_inactive_spin_orbs = len(self.inactive_spin_idx)
        self.num_virtual_spin_orbs = len(self.virtual_spin_idx)
        if self.num_active_elec % 2 != 0:
            raise ValueError("Number of active electrons has to be even")
        self.num_active_elec_alpha = self.num_active_elec // 2
        self.num_active_elec_beta = self.num_active_elec // 2
        self.num_inactive_orbs = self.num_inactive_spin_orbs // 2
        self.num_active_orbs = self.num_active_spin_orbs // 2
        self.num_virtual_orbs = self.num_virtual_spin_orbs // 2

        def spatial(spin_list):
            seen: list[int] = []
            for idx in spin_list:
                if idx // 2 not in seen:
                    seen.append(idx // 2)
            return seen

        self.inactive_idx = spatial(self.inactive_spin_idx)
        self.active_idx = spatial(self.active_spin_idx)
        self.virtual_idx = spatial(self.virtual_spin_idx)
        self.active_occ_idx = spatial(self.active_occ_spin_idx)
        self.active_unocc_idx = spatial(self.active_unocc_spin_idx)
        s_shift = min(self.active_spin_idx) if self.active_spin_idx else 0
        self.active_spin_idx_shifted = [i - s_shift for i in self.active_spin_idx]
        self.active_occ_spin_idx_shifted = [i - s_shift for i in self.active_occ_spin_idx]
        self.active_unocc_spin_idx_shifted = [i - s_shift for i in self.active_unocc_spin_idx]
        o_shift = min(self.active_idx) if self.active_idx else 0
        self.active_idx_shifted = [i - o_shift for i in self.active_idx]
        self.active_occ_idx_shifted = [i - o_shift for i in self.active_occ_idx]
        self.active_unocc_idx_shifted = [i - o_shift for i in self.active_unocc_idx]
        # ---- orbital-rotation parameters (ups_wavefunction.py:175-213) ----
        self._kappa: list[float] = []
        self._kappa_old: list[float] = []
        kappa_idx, no_aa, no_aa_dagger, redundant, hf_like = [], [], [], [], []
        inact, act, virt = set(self.inactive_idx), set(self.active_idx), set(self.virtual_idx)
        for p in range(self.num_orbs):
            for q in range(p + 1, self.num_orbs):
                if (p in inact and q in inact) or (p in virt and q in virt):
                    redundant.append((p, q))
                    continue
                if not include_active_kappa and p in act and q in act:
                    redundant.append((p, q))
                    continue
                if not (p in act and q in act):
                    no_aa.append((p, q))
                    no_aa_dagger.append((q, p))
                self._kappa.append(0.0)
                self._kappa_old.append(0.0)
                kappa_idx.append((p, q))
        occ, unocc = set(self.active_occ_idx), set(self.active_unocc_idx)
        for p in range(self.num_orbs):
            for q in range(p + 1, self.num_orbs):
                if (p in inact and q in virt) or (p in inact and q in unocc) or (p in occ and q in virt):
                    hf_like.append((p, q))
        self.kappa_idx = np.array(kappa_idx, dtype=int).reshape(-1, 2)
        self.kappa_no_activeactive_idx = np.array(no_aa, dtype=int)
        self.kappa_no_activeactive_idx_dagger = np.array(no_aa_dagger, dtype=int)
        self.kappa_redundant_idx = np.array(redundant, dtype=int)
        self.kappa_hf_like_idx = np.array(hf_like, dtype=int)
        # ---- CI space and reference determinant (ups_wavefunction.py:207-261) ----
        self.ci_info = get_indexing(
            self.num_inactive_orbs,
            self.num_active_orbs,
            self.num_virtual_orbs,
            self.num_active_elec_alpha,
            self.num_active_elec_beta,
            device=device,
        )
        self.num_det = self.ci_info.num_det
        self.csf_coeffs = np.zeros(self.num_det)
        hf_det = "1" * self.num_active_elec + "0" * (self.num_active_spin_orbs - self.num_active_elec)
        self._pp = False
        a_low = ansatz.lower()
        if a_low == "tups" and self.ansatz_options.get("do_pp", False):
            pp_det = self._perfect_pairing_det()
            hole = [i for i, (h, p) in enumerate(zip(hf_det, pp_det)) if h == "1" and p == "0"]
            part = [i for i, (h, p) in enumerate(zip(hf_det, pp_det)) if h == "0" and p == "1"]
            hole_spatial = sorted(set(i // 2 + self.num_inactive_orbs for i in hole))
            part_spatial = sorted(set(i // 2 + self.num_inactive_orbs for i in part))
            pp_mo = np.array(mo_coeffs, copy=True)
            pp_mo[:, hole_spatial + part_spatial] = pp_mo[:, part_spatial + hole_spatial]
            self._c_mo = pp_mo
            self.csf_coeffs[self.ci_info.det2idx[int(pp_det, 2)]] = 1
            self._pp = True
        else:
            self.csf_coeffs[self.ci_info.det2idx[int(hf_det, 2)]] = 1
            self._c_mo = mo_coeffs
        # ---- ansatz layout (ups_wavefunction.py:262-313) ----
        self.ups_layout = UpsStructure()
        fucc_args = (
            self.active_occ_idx_shifted,
            self.active_unocc_idx_shifted,
            self.active_occ_spin_idx_shifted,
            self.active_unocc_spin_idx_shifted,
            self.num_active_orbs,
            self.ansatz_options,
        )
        if a_low in ("tups", "qnp"):
            self.ansatz_options["do_tups" if a_low == "tups" else "do_qnp"] = True
            self.ups_layout.create_tiled(self.num_active_orbs, self.ansatz_options)
        elif a_low in ("fucc", "fuccsd", "ksafupccgsd", "fuccpd", "safuccsd"):
            if a_low == "fuccsd":
                self.ansatz_options["S"] = True
                self.ansatz_options["D"] = True
            elif a_low == "ksafupccgsd":
                self.ansatz_options["SAGS"] = True
                self.ansatz_options["GpD"] = True
            self.ansatz_options.setdefault("n_layers", 1)
            self.ups_layout.create_fUCC(*fucc_args)
        elif a_low in ("sdsfuccsd", "ksasdsfupccgsd"):
            self.ansatz_options["D" if a_low == "sdsfuccsd" else "GpD"] = True
            self.ansatz_options.setdefault("n_layers", 1)
            self.ups_layout.create_SDSfUCC(*fucc_args)
        else:
            raise ValueError(f"Got unknown ansatz, {ansatz}")
        self._thetas = np.zeros(self.ups_layout.n_params).tolist()
        # a single-determinant reference: the `thetas` setter runs the head of the circuit in the orbital window it can reach
        # (operator_state_algebra.construct_ups_state_from_determinant); `light_cone = False` takes the plain route
        nz = np.flatnonzero(self.csf_coeffs)
        self._ref_det: int | None = int(nz[0]) if nz.size == 1 and self.csf_coeffs[nz[0]] == 1.0 else None
        self.light_cone = True
        self._sigma_dev: torch.Tensor | None = None   # H|psi> of the current state and integrals (shared by energy and gradient)
        self._ci_from_thetas = True                   # the device state is U(thetas)|reference>
        dev = torch.device("cuda", self.ci_info.device)
        self._csf_dev = torch.from_numpy(self.csf_coeffs).to(dev)
        self._ci_dev = self._csf_dev.clone()
        self._ci_host: np.ndarray | None = None
        self._old_opt_parameters = np.zeros(len(self._thetas) + len(self._kappa)) + 10**20
        self._E_opt_old = 0.0

    def _perfect_pairing_det(self) -> str:
        """Perfect-pairing reference determinant string (ups_wavefunction.py:224-247)."""
        pp_det = ""
        spin_orb = 0
        elec_count = self.num_active_elec
        while spin_orb < self.num_active_spin_orbs:
            left = self.num_active_spin_orbs - spin_orb
            if elec_count >= 2 and left >= 4 and elec_count <= left - 2:
                pp_det += "1100"
                elec_count -= 2
                spin_orb += 4
            elif elec_count == 0:
                pp_det += "0"
                spin_orb += 1
            else:
                pp_det += "1"
                spin_orb += 1
                elec_count -= 1
        print("perfect-pairing determinant found as:", pp_det)
        if len(pp_det) != self.num_active_spin_orbs or pp_det.count("1") != self.num_active_elec:
            raise ValueError("Perfect pairing determinant violates orbital or electron numbers")
        return pp_det

    # ---- parameters -------------------------------------------------------------------------
    @property
    def kappa(self) -> list[float]:
        return self._kappa.copy()

    @kappa.setter
    def kappa(self, k: list[float]) -> None:
        """Set orbital-rotation parameters and move the expansion point (ups_wavefunction.py:320-334)."""
        self._h_mo = None
        self._g_mo = None
        self._energy_elec = None
        self._sigma_dev = None
        self._kappa = list(k)
        self._c_mo = self.c_mo
        self._kappa_old = self.kappa

    @property
    def thetas(self) -> list[float]:
        return self._thetas.copy()

    @thetas.setter
    def thetas(self, theta_vals: list[float]) -> None:
        """Set ansatz parameters and rebuild the state on the device (ups_wavefunction.py:345-365)."""
        if len(theta_vals) != len(self._thetas):
            raise ValueError(f"Expected {len(self._thetas)} theta1 values got {len(theta_vals)}")
        self._rdm1 = self._rdm2 = None
        self._rdm3 = self._rdm4 = None
        self._energy_elec = None
        self._sigma_dev = None
        self._ci_from_thetas = True
        self._thetas = [float(x) for x in theta_vals]
        if self._ref_det is not None and self.light_cone:
            self._ci_dev = osa.construct_ups_state_from_determinant(self._ref_det, self.ci_info, self._thetas, self.ups_layout)
        else:
            self._ci_dev = osa.construct_ups_state(self._csf_dev, self.ci_info, self._thetas, self.ups_layout)
        self._ci_host = None

    @property
    def ci_coeffs(self) -> np.ndarray:
        if self._ci_host is None:
            self._ci_host = self._ci_dev.cpu().numpy()
        return self._ci_host

    @ci_coeffs.setter
    def ci_coeffs(self, value) -> None:
        dev = torch.device("cuda", self.ci_info.device)
        self._ci_dev = torch.as_tensor(np.asarray(value, dtype=np.float64)).to(dev).clone()
        self._ci_host = None
        self._rdm1 = self._rdm2 = None
        self._rdm3 = self._rdm4 = None
        self._energy_elec = None
        self._sigma_dev = None
        self._ci_from_thetas = False

    @property
    def ci_coeffs_device(self) -> torch.Tensor:
        return self._ci_dev

    # ---- integrals ----------------------------------------------------------------------------
    @property
    def c_mo(self) -> np.ndarray:
        """MO coefficients rotated by exp(-kappa) relative to the expansion point (ups_wavefunction.py:367-385)."""
        kappa_mat = np.zeros_like(self._c_mo)
        if len(self._kappa) != 0:
            diff = np.array(self._kappa) - np.array(self._kappa_old)
            if np.max(np.abs(diff)) > 0.0:
                for d, (p, q) in zip(diff, self.kappa_idx):
                    kappa_mat[p, q] = d
                    kappa_mat[q, p] = -d
        return np.matmul(self._c_mo, scipy.linalg.expm(-kappa_mat))

    @property
    def h_mo(self) -> np.ndarray:
        if self._h_mo is None:
            self._h_mo = one_electron_integral_transform(self.c_mo, self.int_gen.h_ao)
        return self._h_mo

    @property
    def g_mo(self) -> np.ndarray:
        if self._g_mo is None:
            self._g_mo = two_electron_integral_transform(self.c_mo, self.int_gen.electron_electron_repulsion)
        return self._g_mo

    # ---- densities and energy -------------------------------------------------------------------
    def _build_rdms(self, want_rdm2: bool) -> None:
        d1, d2 = osa.reduced_density_matrices(self._ci_dev, self._ci_dev, self.ci_info, want_rdm2=want_rdm2)
        # the reference evaluates q <= p and mirrors (ups_wavefunction.py:416-429)
        low = np.tril(d1)
        self._rdm1 = low + low.T - np.diag(np.diag(d1))
        if want_rdm2:
            self._rdm2 = symmetrize_rdm2_like_reference(d2)

    @property
    def rdm1(self) -> np.ndarray:
        if self._rdm1 is None:
            self._build_rdms(False)
        return self._rdm1

    @property
    def rdm2(self) -> np.ndarray:
        if self._rdm2 is None:
            self._build_rdms(True)
        return self._rdm2

    def _build_higher_rdms(self, want_rdm4: bool) -> None:
        d3, d4 = osa.higher_reduced_density_matrices(self._ci_dev, self.ci_info, self.rdm1, self.rdm2, want_rdm4=want_rdm4)
        self._rdm3 = d3
        if want_rdm4:
            self._rdm4 = d4

    @property
    def rdm3(self) -> np.ndarray:
        """Three-electron RDM in the active space (ups_wavefunction.py:478-545)."""
        if getattr(self, "_rdm3", None) is None:
            self._build_higher_rdms(False)
        return self._rdm3

    @property
    def rdm4(self) -> np.ndarray:
        """Four-electron RDM in the active space (ups_wavefunction.py:547-754)."""
        if getattr(self, "_rdm4", None) is None:
            self._build_higher_rdms(True)
        return self._rdm4

    @property
    def energy_elec(self) -> float:
        """<Psi|H|Psi> through the sigma kernel (ups_wavefunction.py:770-784)."""
        if self._energy_elec is None:
            self._energy_elec = osa._dot(self._ci_dev, self._sigma(), self.ci_info)   # <psi|H|psi> with the kept H|psi>
        return self._energy_elec

    def _sigma(self) -> torch.Tensor:
        """H|psi> of the current state with the current integrals, kept until the state or the integrals change."""
        if self._sigma_dev is None:
            H = hamiltonian_0i_0a(self.h_mo, self.g_mo, self.num_inactive_orbs, self.num_active_orbs)
            self._sigma_dev = osa.propagate_state([H], self._ci_dev, self.ci_info)
        return self._sigma_dev

    def _set_thetas_if_changed(self, theta_vals) -> None:
        """The optimiser asks for the energy and then for the gradient at the same parameters: one state construction, not two."""
        new = [float(x) for x in theta_vals]
        if new != self._thetas or not self._ci_from_thetas:
            self.thetas = new

    # ---- optimisation callables (ups_wavefunction.py:1019-1142) -----------------------------------
    def _calc_energy_optimization(self, parameters, theta_optimization: bool, kappa_optimization: bool) -> float:
        if np.max(np.abs(np.array(self._old_opt_parameters) - np.array(parameters))) < 10**-14:
            return self._E_opt_old
        num_kappa = 0
        if kappa_optimization:
            num_kappa = len(self.kappa_idx)
            self.kappa = list(parameters[:num_kappa])
        if theta_optimization:
            self._set_thetas_if_changed(parameters[num_kappa:])
        if kappa_optimization:
            E = get_electronic_energy(
                self.h_mo, self.g_mo, self.num_inactive_orbs, self.num_active_orbs, self.rdm1, self.rdm2
            )
        else:
            E = osa._dot(self._ci_dev, self._sigma(), self.ci_info)
        self._E_opt_old = E
        self._old_opt_parameters = np.copy(parameters)
        self.num_energy_evals += 1
        return E

    def _calc_gradient_optimization(self, parameters, theta_optimization: bool, kappa_optimization: bool) -> np.ndarray:
        gradient = np.zeros(len(parameters))
        num_kappa = 0
        if kappa_optimization:
            num_kappa = len(self.kappa_idx)
            self.kappa = list(parameters[:num_kappa])
        if theta_optimization:
            self._set_thetas_if_changed(parameters[num_kappa:])
        if kappa_optimization:
            gradient[:num_kappa] = get_orbital_gradient(
                self.h_mo, self.g_mo, self.kappa_idx, self.num_inactive_orbs, self.num_active_orbs, self.rdm1, self.rdm2
            )
        if theta_optimization:
            # the loop of ups_wavefunction.py:1114-1138 run BACKWARDS through the circuit from (H|psi>, |psi>): 2 <bra|T_k|ket>, then
            # both vectors <- U_k^dagger (T_k commutes with its own rotation: the same numbers, without the adjoint pass U^dagger H|psi>
            # and without rebuilding the state from the reference); H|psi> is shared with the energy evaluation at the same parameters
            g, _, _ = osa.ups_gradient_sweep_backward(self._sigma(), self._ci_dev, self.ci_info, self._thetas, self.ups_layout)
            gradient[num_kappa:] += g
            self.num_energy_evals += 2 * int(np.sum(list(self.ups_layout.grad_param_R.values())))
        return gradient

    def run_wf_optimization_1step(
        self, optimizer_name: str, orbital_optimization: bool = False, tol: float = 1e-10, maxiter: int = 1000
    ) -> None:
        """One-step optimisation of thetas (and kappa) with a SciPy optimiser (ups_wavefunction.py:916-1017).

        The optimiser loop is host control flow; every energy / gradient evaluation runs on the device.
        """
        if optimizer_name.lower() in ("rotosolve", "cobyla", "cobyqa"):
            self._run_optimizer_by_name(optimizer_name, orbital_optimization, tol, maxiter)
            return
        method = {"bfgs": "BFGS", "l-bfgs-b": "L-BFGS-B", "slsqp": "SLSQP"}.get(optimizer_name.lower())
        if method is None:
            raise ValueError(f"Unknown optimizer: {optimizer_name}")
        x0 = (self.kappa if orbital_optimization else []) + self.thetas
        self._old_opt_parameters = np.zeros(len(x0)) + 10**20
        res = scipy.optimize.minimize(
            lambda x: self._calc_energy_optimization(list(x), True, orbital_optimization),
            np.array(x0, dtype=float),
            jac=lambda x: self._calc_gradient_optimization(list(x), True, orbital_optimization),
            method=method,
            tol=tol,
            options={"maxiter": maxiter},
        )
        nk = len(self.kappa_idx) if orbital_optimization else 0
        if orbital_optimization:
            self.kappa = list(res.x[:nk])
        self.thetas = list(res.x[nk:])
        self._energy_elec = None

    # ---- pieces shared by the one- and two-step drivers (and by the state-averaged subclass) ----
    def _optimizer(self, name: str, theta: bool, kappa: bool, tol: float, maxiter: int, silent: bool = False):
        from functools import partial

        from slowquant_b200.optimizers import Optimizers

        return Optimizers(
            partial(self._calc_energy_optimization, theta_optimization=theta, kappa_optimization=kappa),
            name,
            grad=partial(self._calc_gradient_optimization, theta_optimization=theta, kappa_optimization=kappa),
            maxiter=maxiter,
            tol=tol,
            is_silent=silent,
            energy_eval_callback=lambda: self.num_energy_evals,
        )

    def _rotosolve_options(self, name: str):
        if name.lower() != "rotosolve":
            return None
        if self.ups_layout is None:
            raise ValueError("RotoSolve needs a product ansatz (it is not defined for the non-factorised UCC)")
        return {
            "R": self.ups_layout.grad_param_R,
            "param_names": self.ups_layout.param_names,
            "f_rotosolve_optimized": self._calc_energy_rotosolve_optimization,
        }

    def _finish_optimization(self, energy: float) -> None:
        self._energy_elec = energy

    def run_wf_optimization_2step(
        self, optimizer_name: str, orbital_optimization: bool = False, tol: float = 1e-10, maxiter: int = 1000,
        is_silent_subiterations: bool = False,
    ) -> None:
        """Alternate ansatz (theta) and orbital (kappa, L-BFGS-B) optimisations until the energy stops changing
        (ups_wavefunction.py:799-918; sa_ups_wavefunction.py:517-638 for the state-averaged class)."""
        import time

        print("### Parameters information:")
        if orbital_optimization:
            print(f"### Number kappa: {len(self.kappa)}")
        print(f"### Number theta: {len(self._thetas)}")
        print("Full optimization")
        print("Iteration # | Iteration time [s] | Electronic energy [Hartree] | Energy measurement #")
        e_old = 1e12
        res = None
        for full_iter in range(int(maxiter)):
            full_start = time.time()
            optimizer = self._optimizer(optimizer_name, True, False, tol, maxiter, is_silent_subiterations)
            self._old_opt_parameters = np.zeros(len(self._thetas)) + 10**20
            self._E_opt_old = 0.0
            res = optimizer.minimize(self.thetas, extra_options=self._rotosolve_options(optimizer_name))
            self.thetas = res.x.tolist()
            if not (orbital_optimization and len(self.kappa) != 0):
                # without orbital parameters the ansatz optimisation already is the answer
                if orbital_optimization:
                    print("WARNING: No orbital optimization performed, because there is no non-redundant orbital parameters.")
                break
            optimizer = self._optimizer("l-bfgs-b", False, True, tol, maxiter, is_silent_subiterations)
            self._old_opt_parameters = np.zeros(len(self.kappa_idx)) + 10**20
            self._E_opt_old = 0.0
            res = optimizer.minimize([0.0] * len(self.kappa_idx))
            for i in range(len(self._kappa)):   # the expansion point has moved with every evaluation (kappa setter)
                self._kappa[i] = 0.0
                self._kappa_old[i] = 0.0
            e_new = res.fun
            print(f"{str(full_iter + 1).center(11)} | {f'{time.time() - full_start:7.2f}'.center(18)} | {f'{e_new:3.12f}'.center(27)} | {str(self.num_energy_evals).center(11)}")
            if abs(e_new - e_old) < tol:
                break
            e_old = e_new
        self._finish_optimization(res.fun)

    def check_orthonormality(self, overlap_integral: np.ndarray) -> None:
        """Print max|C^T S C - 1| (ups_wavefunction.py:756-768)."""
        S_ortho = one_electron_integral_transform(self.c_mo, overlap_integral)
        print("Max ortho-normal diff:", np.max(np.abs(S_ortho - np.identity(len(S_ortho)))))

    def _get_hamiltonian(self, qiskit_form: bool = False):
        """Energy Hamiltonian folded onto the active space as explicit strings (ups_wavefunction.py:786-797)."""
        if qiskit_form:
            raise NotImplementedError("the Qiskit form belongs to the Qiskit interface, which is outside this engine's scope")
        H = hamiltonian_0i_0a(self.h_mo, self.g_mo, self.num_inactive_orbs, self.num_active_orbs)
        return H.get_folded_operator(self.num_inactive_orbs, self.num_active_orbs, self.num_virtual_orbs)

    def _run_optimizer_by_name(self, optimizer_name: str, orbital_optimization: bool, tol: float, maxiter: int) -> None:
        """RotoSolve / gradient-free SciPy methods through the Optimizers front end (ups_wavefunction.py:934-1017)."""
        from functools import partial

        from slowquant_b200.optimizers import Optimizers

        if optimizer_name.lower() == "rotosolve" and orbital_optimization and len(self.kappa) != 0:
            raise ValueError("Cannot use RotoSolve together with orbital optimization in the one-step solver.")
        theta_opt = len(self.thetas) > 0 or not orbital_optimization
        energy = partial(self._calc_energy_optimization, theta_optimization=theta_opt, kappa_optimization=orbital_optimization)
        gradient = partial(self._calc_gradient_optimization, theta_optimization=theta_opt, kappa_optimization=orbital_optimization)
        parameters = (self.kappa if orbital_optimization else []) + (self.thetas if theta_opt else [])
        optimizer = Optimizers(
            energy, optimizer_name, grad=gradient, maxiter=maxiter, tol=tol, energy_eval_callback=lambda: self.num_energy_evals
        )
        self._old_opt_parameters = np.zeros(len(parameters)) + 10**20
        self._E_opt_old = 0.0
        extra = None
        if optimizer_name.lower() == "rotosolve":
            extra = {
                "R": self.ups_layout.grad_param_R,
                "param_names": self.ups_layout.param_names,
                "f_rotosolve_optimized": self._calc_energy_rotosolve_optimization,
            }
        res = optimizer.minimize(parameters, extra_options=extra)
        nk = len(self.kappa_idx) if orbital_optimization else 0
        if orbital_optimization:
            self.kappa = list(res.x[:nk])
        self.thetas = list(res.x[nk:])
        self._energy_elec = None

    def _calc_energy_rotosolve_optimization(self, parameters: list[float], theta_diffs: list[float], theta_idx: int) -> list[float]:
        """Energies at all shifted values of theta[theta_idx] (ups_wavefunction.py:1144-1194), device resident.

        The reference propagates operator by operator through its `_SA` kernels; here the prefix
        U_{idx-1}..U_0|CSF> is built once, every shift is one single-unitary launch, and the remaining operator range
        runs once for the whole batch of shifted states (sq_ups_apply_batch), followed by a sigma build and a dot per state.
        """
        th = np.asarray(parameters, dtype=np.float64).copy()
        n = len(th)
        prefix = self._csf_dev.clone()
        if theta_idx > 0:
            osa._ups_apply_inplace(prefix, self.ci_info, th, self.ups_layout, 0, theta_idx, False)
        H = hamiltonian_0i_0a(self.h_mo, self.g_mo, self.num_inactive_orbs, self.num_active_orbs)
        # the shifted operator is applied per shift; the remaining operators are the same for every shifted state, so the whole
        # [n_shifts, N_det] batch goes through ONE launch sequence (the reference batches them through its _SA kernels, :1183-1187)
        kets = prefix.unsqueeze(0).repeat(len(theta_diffs), 1).contiguous()
        for j, shift in enumerate(theta_diffs):
            th_s = th.copy()
            th_s[theta_idx] = shift
            osa._ups_apply_inplace(kets[j], self.ci_info, th_s, self.ups_layout, theta_idx, theta_idx + 1, False)
        if theta_idx + 1 < n:
            osa._ups_apply_batch_inplace(kets, self.ci_info, th, self.ups_layout, theta_idx + 1, n, False)
        energies = []
        for j in range(len(theta_diffs)):
            Hket = osa.propagate_state([H], kets[j], self.ci_info)
            energies.append(osa._dot(Hket, kets[j], self.ci_info))
        self.num_energy_evals += len(energies)
        return energies
