"""Golden vectors for config 2 of BASELINE.json: N2 / cc-pVDZ, oo-tUPS CAS(10,10), 63 504 determinants.

Runs the REFERENCE (imported from /root/reference) at fixed (theta, c_mo = RHF orbitals): state, energy by
the string path and by the RDM path, rdm1, rdm2, orbital gradient (236 kappa) and theta gradient.  Only the
integral blocks the path touches are stored (indices below nI+nA on three of the four g axes).

    python tests/golden/make_golden_n2.py        # build container only; ~2 minutes
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up the pyscf stub and the reference import path)

import slowquant.SlowQuant as sq  # noqa: E402
from slowquant.unitary_coupled_cluster.density_matrix import get_electronic_energy, get_orbital_gradient  # noqa: E402
from slowquant.unitary_coupled_cluster.ups_wavefunction import WaveFunctionUPS  # noqa: E402


def main():
    SQobj = sq.SlowQuant()
    SQobj.set_molecule("""N 0.0 0.0 0.0; N 0.0 0.0 1.1;""", distance_unit="angstrom")
    SQobj.set_basis_set("cc-pVDZ")
    SQobj.init_hartree_fock()
    SQobj.hartree_fock.run_restricted_hartree_fock()
    WF = WaveFunctionUPS((10, 10), SQobj.hartree_fock.mo_coeff, SQobj, "tUPS", {"n_layers": 2}, include_active_kappa=False)
    rng = np.random.default_rng(2024)
    thetas = rng.uniform(-0.3, 0.3, WF.ups_layout.n_params).tolist()
    WF.thetas = thetas
    nI, nA = WF.num_inactive_orbs, WF.num_active_orbs
    M = nI + nA
    h, g = np.array(WF.h_mo), np.array(WF.g_mo)
    out = {
        "thetas": np.array(thetas),
        "dims": np.array([nI, nA, WF.num_virtual_orbs, WF.num_active_elec_alpha, WF.num_active_elec_beta]),
        "kappa_idx": np.array(WF.kappa_idx, dtype=np.int64),
        "h_mo": h,
        # blocks of g the energy / orbital-gradient formulas read: g[n, p, q, r] and g[p, n, q, r], p, q, r < M
        "g_npqr": np.ascontiguousarray(g[:, :M, :M, :M]),
        "g_pnqr": np.ascontiguousarray(g[:M, :, :M, :M]),
        "energy_strings": np.array(WF.energy_elec),
        "rdm1": np.array(WF.rdm1),
        "rdm2": np.array(WF.rdm2),
    }
    out["energy_rdm"] = np.array(get_electronic_energy(h, g, nI, nA, WF.rdm1, WF.rdm2))
    out["orbital_gradient"] = get_orbital_gradient(h, g, WF.kappa_idx, nI, nA, WF.rdm1, WF.rdm2)
    WF._old_opt_parameters = np.zeros(len(thetas)) + 10**20
    out["theta_gradient"] = WF._calc_gradient_optimization(thetas, True, False)
    ci = np.array(WF.ci_coeffs)
    out["ci_nonzero_idx"] = np.nonzero(ci)[0].astype(np.int64)
    out["ci_nonzero_val"] = ci[out["ci_nonzero_idx"]]
    np.savez_compressed(os.path.join(HERE, "golden_n2.npz"), **out)
    print("E(strings) =", float(out["energy_strings"]), " E(rdm) =", float(out["energy_rdm"]),
          " n_kappa =", len(WF.kappa_idx), " nnz(ci) =", len(out["ci_nonzero_idx"]))


if __name__ == "__main__":
    main()
