"""Self-consistent linear response (same class name and results as the reference's
slowquant/unitary_coupled_cluster/linear_response/selfconsistent.py), built from panels in the extended CI space.

The orbital rotations act on the correlated state before the adjoint ansatz is applied (U^d q|0>), which leaves the CAS:
the vectors live in ``get_indexing_extended(..., order=1)`` (CAS + singles into inactive / virtual orbitals), where
q|0> is an ordinary vector.  With the panels

    GC[I] = G_I|CSF>,  X[J] = U^d H U G_J|CSF>,  GUdH[I] = G_I U^d H|0>,  GdUdH[I] = G_I^d U^d H|0>,
    V[J] = U^d q_J|0>,  UdHq[J] = U^d H q_J|0>,  UdqdH[J] = U^d q_J^d H|0>,
    GZ[I] = G_I U^d H^d|0>,  GdZ[I] = G_I^d U^d H|0>          (H = hamiltonian_1i_1a for the last two rows)

the blocks of selfconsistent.py:140-282 are Gram matrices:

    A_GG = GC X^T - 1/2 (GC GUdH^T + GUdH GC^T),  B_GG = -GC GdUdH^T,  Sigma_GG = 1,
    A_Gq = GC UdHq^T - 1/2 GZ V^T,                B_Gq = -1/2 (GC UdqdH^T + GdZ V^T).

Operators that leave the extended space are applied with the reference's ``do_unsafe=True`` semantics (the part outside
the space is dropped after every operator).
"""
from __future__ import annotations

import numpy as np
import torch

from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.ci_spaces import get_indexing_extended
from slowquant_b200.integral_manager import one_electron_integral_transform
from slowquant_b200.linear_response import _panels as pn
from slowquant_b200.linear_response.lr_baseclass import LinearResponseBaseClass
from slowquant_b200.operators import one_elec_op_0i_0a
from slowquant_b200.ucc_wavefunction import WaveFunctionUCC
from slowquant_b200.ups_wavefunction import WaveFunctionUPS


class LinearResponse(LinearResponseBaseClass):
    def __init__(self, wave_function: WaveFunctionUCC | WaveFunctionUPS, excitations: str) -> None:
        super().__init__(wave_function, excitations)
        wf = self.wf
        ci_ext = get_indexing_extended(
            wf.num_inactive_orbs, wf.num_active_orbs, wf.num_virtual_orbs, wf.num_active_elec_alpha, wf.num_active_elec_beta, 1,
            device=wf.ci_info.device,
        )
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
        nq, nG = len(self.q_ops), len(self.G_ops)
        print("Gs", nG)
        print("qs", nq)
        pn.orbital_blocks(self)

        def panel(ops_lists, src, unsafe=False):
            rows = [osa.propagate_state(ops, src, *ext, do_unsafe=unsafe) for ops in ops_lists]
            return torch.stack(rows) if rows else torch.zeros((0, ci_ext.num_det), dtype=torch.float64, device=dev)

        UdH0 = osa.propagate_state(["Ud", self.H_0i_0a], psi, *ext)                # U^d H|0>
        GC = panel([[G] for G in self.G_ops], csf)                                 # G_I|CSF>
        self._GC = GC
        if nG != 0:
            gh = (GC @ UdH0).cpu().numpy()
            pn.check_active_gradient(np.concatenate([-gh, gh]))
        if nq != 0 and nG != 0:
            H1 = self.H_1i_1a
            UdHq = panel([["Ud", H1, q] for q in self.q_ops], psi, unsafe=True)
            UdqdH = panel([["Ud", q.dagger, H1] for q in self.q_ops], psi, unsafe=True)
            V = panel([["Ud", q] for q in self.q_ops], psi, unsafe=True)           # U^d q_J|0>
            zA = osa.propagate_state(["Ud", H1.dagger], psi, *ext, do_unsafe=True)  # <0|H U ... = <U^d H^d 0| ...
            zB = osa.propagate_state(["Ud", H1], psi, *ext, do_unsafe=True)
            GZ = panel([[G] for G in self.G_ops], zA)
            GdZ = panel([[G.dagger] for G in self.G_ops], zB)
            A_Gq = (pn.gram(GC, UdHq) - 0.5 * pn.gram(GZ, V)).cpu().numpy()
            B_Gq = (-0.5 * (pn.gram(GC, UdqdH) + pn.gram(GdZ, V))).cpu().numpy()
            self.A[nq:, :nq], self.A[:nq, nq:] = A_Gq, A_Gq.T
            self.B[nq:, :nq], self.B[:nq, nq:] = B_Gq, B_Gq.T
        if nG != 0:
            X = panel([["Ud", self.H_0i_0a, "U", G] for G in self.G_ops], csf)
            GUdH = panel([[G] for G in self.G_ops], UdH0)
            GdUdH = panel([[G.dagger] for G in self.G_ops], UdH0)
            A_GG = pn.gram(GC, X) - 0.5 * (pn.gram(GC, GUdH) + pn.gram(GUdH, GC))
            B_GG = -pn.gram(GC, GdUdH)
            self.A[nq:, nq:] = pn.mirror_lower(A_GG)
            self.B[nq:, nq:] = pn.mirror_lower(B_GG)
            self.Sigma[nq:, nq:] = np.eye(nG)

    def get_transition_dipole(self) -> np.ndarray:
        """selfconsistent.py:284-399: -Z_i <0|mu U G_i|CSF> + Y_i <CSF|G_i^d U^d mu|0> + orbital part."""
        wf = self.wf
        ext = self.index_info_extended
        nI, nA = wf.num_inactive_orbs, wf.num_active_orbs
        number_excitations = len(self.excitation_energies)
        dev = self._psi.device
        Z = torch.from_numpy(np.ascontiguousarray(self.Z_G_normed.T)).to(dev)
        Y = torch.from_numpy(np.ascontiguousarray(self.Y_G_normed.T)).to(dev)
        dipole_integrals = wf.int_gen.electric_dipole
        transition_dipoles = np.zeros((number_excitations, 3))
        for axis in range(3):
            mu = one_electron_integral_transform(wf.c_mo, dipole_integrals[axis])
            mu_op = one_elec_op_0i_0a(mu, nI, nA)
            Ud_mud = osa.propagate_state(["Ud", mu_op.dagger], self._psi, *ext)
            Ud_mu = osa.propagate_state(["Ud", mu_op], self._psi, *ext)
            active = (-Z @ (self._GC @ Ud_mud) + Y @ (self._GC @ Ud_mu)).cpu().numpy()
            for s in range(number_excitations):
                transition_dipoles[s, axis] = pn.orbital_property_part(self, mu, s, number_excitations) + active[s]
        return transition_dipoles
