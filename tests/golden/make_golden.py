"""Generate the golden vectors in tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference; numba, scipy):

    python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference; its unconditional ``import pyscf`` (used only
for isinstance checks, integral_manager.py:37-120) is satisfied by a 3-file stub written to a temp dir.
Outputs (committed): tests/golden/*.npz + tests/golden/*.json.  Nothing here is used at run time by the
product; the GPU box never sees /root/reference.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    stub = tempfile.mkdtemp(prefix="pyscf_stub_")
    os.makedirs(os.path.join(stub, "pyscf", "gto"))
    with open(os.path.join(stub, "pyscf", "__init__.py"), "w") as f:
        f.write("from . import gto\n")
    with open(os.path.join(stub, "pyscf", "gto", "__init__.py"), "w") as f:
        f.write("from . import mole\n")
    with open(os.path.join(stub, "pyscf", "gto", "mole.py"), "w") as f:
        f.write("class Mole:\n    pass\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, "/root/reference")


_import_reference()

import slowquant.SlowQuant as sq  # noqa: E402
from slowquant.unitary_coupled_cluster import operators as rops  # noqa: E402
from slowquant.unitary_coupled_cluster.ci_spaces import get_indexing  # noqa: E402
from slowquant.unitary_coupled_cluster.density_matrix import (  # noqa: E402
    get_electronic_energy,
    get_orbital_gradient,
)
from slowquant.unitary_coupled_cluster.fermionic_operator import FermionicOperator  # noqa: E402
from slowquant.unitary_coupled_cluster.operator_state_algebra import (  # noqa: E402
    construct_ups_state,
    expectation_value,
    get_grad_action,
    propagate_state,
    propagate_unitary,
)
from slowquant.unitary_coupled_cluster.ups_wavefunction import WaveFunctionUPS  # noqa: E402
from slowquant.unitary_coupled_cluster.util import UpsStructure  # noqa: E402


def op_to_json(op: FermionicOperator) -> dict:
    return {
        "labels": [[[int(i), bool(d)] for i, d in label] for label in op.operators.keys()],
        "coeffs": [float(v) for v in op.operators.values()],
    }


def sym_integrals(n: int, seed: int):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    return h, g


def golden_indexing(out: dict) -> None:
    for n, na, nb in [(2, 1, 1), (3, 1, 2), (4, 2, 2), (5, 2, 3), (6, 3, 3), (6, 4, 1), (7, 3, 4)]:
        ci = get_indexing(0, n, 0, na, nb)
        out[f"idx2det_{n}_{na}_{nb}"] = np.asarray(ci.idx2det, dtype=np.int64)


def golden_operators() -> dict:
    ops = {}
    ops["Epq_2_0"] = op_to_json(rops.Epq(2, 0))
    ops["Epq_1_1"] = op_to_json(rops.Epq(1, 1))
    ops["Epq_3_1*Epq_1_2"] = op_to_json(rops.Epq(3, 1) * rops.Epq(1, 2))
    ops["Epq_0_1*Epq_1_0"] = op_to_json(rops.Epq(0, 1) * rops.Epq(1, 0))
    ops["epqrs_0_1_1_2"] = op_to_json(rops.epqrs(0, 1, 1, 2))
    ops["epqrs_2_2_2_2"] = op_to_json(rops.epqrs(2, 2, 2, 2))
    ops["G1_1_4_AH"] = op_to_json(rops.G1(1, 4, True))
    ops["G2_0_3_4_7_AH"] = op_to_json(rops.G2(0, 3, 4, 7, True))
    ops["G2_0_1_6_7_AH"] = op_to_json(rops.G2(0, 1, 6, 7, True))
    ops["G2_2_3_4_5_H"] = op_to_json(rops.G2(2, 3, 4, 5, False))
    ops["G3_0_1_2_5_6_7_AH"] = op_to_json(rops.G3(0, 1, 2, 5, 6, 7, True))
    ops["G4_0_1_2_3_4_5_6_7_AH"] = op_to_json(rops.G4(0, 1, 2, 3, 4, 5, 6, 7, True))
    ops["G1_sa_0_2_AH"] = op_to_json(rops.G1_sa(0, 2, True))
    ops["G2_sa_0_0_2_2_c1_AH"] = op_to_json(rops.G2_sa(0, 0, 2, 2, 1, True))
    ops["G2_sa_0_0_2_3_c2_AH"] = op_to_json(rops.G2_sa(0, 0, 2, 3, 2, True))
    ops["G2_sa_0_1_2_2_c3_AH"] = op_to_json(rops.G2_sa(0, 1, 2, 2, 3, True))
    ops["G2_sa_0_1_2_3_c4_AH"] = op_to_json(rops.G2_sa(0, 1, 2, 3, 4, True))
    ops["G2_sa_0_1_2_3_c5_AH"] = op_to_json(rops.G2_sa(0, 1, 2, 3, 5, True))
    ops["commutator_E01_E12"] = op_to_json(rops.commutator(rops.Epq(0, 1), rops.Epq(1, 2)))
    # folded Hamiltonian, nI=1, nA=3, nV=1
    h, g = sym_integrals(5, 7)
    H = rops.hamiltonian_0i_0a(h, g, 1, 3)
    ops["H0i0a_seed7_1_3_unfolded_count"] = {"n": len(H.operators)}
    ops["H0i0a_seed7_1_3_folded"] = op_to_json(H.get_folded_operator(1, 3, 1))
    return ops


def golden_propagate(out: dict, meta: dict) -> None:
    rng = np.random.default_rng(11)
    # (a) unfolded operators on spaces with unequal spin counts
    cases = [
        ("p0", (0, 4, 0, 2, 2), rops.Epq(3, 1) * rops.Epq(1, 2) + 0.3 * rops.Epq(2, 2)),
        ("p1", (0, 5, 0, 2, 3), rops.G2(0, 3, 4, 7, True) + 0.5 * rops.G1(1, 5, True)),
        ("p2", (0, 5, 0, 3, 2), rops.G2_sa(0, 1, 2, 4, 5, True)),
        ("p3", (0, 4, 0, 2, 2), rops.Epq(0, 1) * rops.Epq(2, 3) * rops.Epq(1, 0)),
    ]
    for name, dims, op in cases:
        ci = get_indexing(*dims)
        state = rng.normal(size=len(ci.idx2det))
        res = propagate_state([op], state, ci, do_folding=False)
        out[f"{name}_state"] = state
        out[f"{name}_result"] = res
        meta[name] = {"dims": list(dims), "op": op_to_json(op)}
    # (b) folded Hamiltonian expectation value, nI=1, nA=3, nV=1, (2,1) active electrons
    h, g = sym_integrals(5, 7)
    ci = get_indexing(1, 3, 1, 2, 1)
    state = rng.normal(size=len(ci.idx2det))
    state /= np.linalg.norm(state)
    H = rops.hamiltonian_0i_0a(h, g, 1, 3)
    out["fold_h"] = h
    out["fold_g"] = g
    out["fold_state"] = state
    out["fold_Hstate"] = propagate_state([H], state, ci)
    out["fold_energy"] = np.array(expectation_value(state, [H], state, ci))
    # (c) a two-operator list (right-to-left order) with folding
    op_a = rops.Epq(1, 2) + 0.25 * rops.Epq(0, 0)
    op_b = rops.Epq(3, 1) * rops.Epq(1, 3)
    out["fold2_result"] = propagate_state([op_a, op_b], state, ci)
    meta["fold2"] = {"op_a": op_to_json(op_a), "op_b": op_to_json(op_b)}


def _h2o():
    SQobj = sq.SlowQuant()
    SQobj.set_molecule(
        """O   0.0  0.0           0.1035174918;
        H   0.0  0.7955612117 -0.4640237459;
        H   0.0 -0.7955612117 -0.4640237459;""",
        distance_unit="angstrom",
    )
    SQobj.set_basis_set("sto-3g")
    SQobj.init_hartree_fock()
    SQobj.hartree_fock.run_restricted_hartree_fock()
    return SQobj


def golden_wavefunctions(out: dict, meta: dict) -> None:
    SQobj = _h2o()
    c_mo = SQobj.hartree_fock.mo_coeff
    specs = [
        # name, cas, ansatz, options, include_active_kappa, theta scale, what to record
        ("tups44", (4, 4), "tUPS", {"n_layers": 2}, True, np.pi, {"rdm", "grad", "steps"}),
        ("qnp44", (4, 4), "QNP", {"n_layers": 2, "skip_last_singles": True}, False, np.pi, {"steps"}),
        ("fuccsd44", (4, 4), "fUCCSD", {}, False, 1.0, {"rdm", "grad", "steps"}),
        ("sa44", (4, 4), "fUCC", {"n_layers": 1, "SAS": True, "SAD": True}, False, 0.7, {"grad", "steps"}),
        ("tq44", (4, 4), "fUCC", {"n_layers": 1, "T": True, "Q": True}, False, 0.9, {"steps"}),
        ("gsd44", (4, 4), "fUCC", {"n_layers": 1, "GS": True, "GD": True}, False, 0.5, set()),
        ("ksa44", (4, 4), "kSAfUpCCGSD", {"n_layers": 2}, False, 0.8, {"grad"}),
        ("sds44", (4, 4), "SDSfUCCSD", {}, False, 0.6, set()),
        ("sad65", (6, 6), "fUCC", {"n_layers": 1, "SAD": True, "pD": True}, False, 0.4, set()),
    ]
    for name, cas, ansatz, options, iak, scale, record in specs:
        WF = WaveFunctionUPS(cas, c_mo, SQobj, ansatz, ansatz_options=dict(options), include_active_kappa=iak)
        rng = np.random.default_rng(abs(hash(name)) % (2**31) if False else sum(map(ord, name)))
        thetas = (scale * rng.uniform(-1, 1, WF.ups_layout.n_params)).tolist()
        if len(thetas) > 3:
            thetas[2] = 0.0  # exercise the |theta| < 1e-28 skip
        WF.thetas = thetas
        layout = WF.ups_layout
        meta[name] = {
            "cas": list(cas),
            "ansatz": ansatz,
            "options": {k: v for k, v in WF.ansatz_options.items()},
            "include_active_kappa": iak,
            "num_inactive_orbs": WF.num_inactive_orbs,
            "num_active_orbs": WF.num_active_orbs,
            "num_virtual_orbs": WF.num_virtual_orbs,
            "n_alpha": WF.num_active_elec_alpha,
            "n_beta": WF.num_active_elec_beta,
            "types": list(layout.excitation_operator_type),
            "indices": [[int(x) for x in t] for t in layout.excitation_indices],
            "active_occ_idx_shifted": [int(x) for x in WF.active_occ_idx_shifted],
            "active_unocc_idx_shifted": [int(x) for x in WF.active_unocc_idx_shifted],
            "active_occ_spin_idx_shifted": [int(x) for x in WF.active_occ_spin_idx_shifted],
            "active_unocc_spin_idx_shifted": [int(x) for x in WF.active_unocc_spin_idx_shifted],
            "grad_param_R": dict(layout.grad_param_R),
        }
        out[f"{name}_thetas"] = np.array(thetas)
        out[f"{name}_csf"] = np.array(WF.csf_coeffs)
        out[f"{name}_ci"] = np.array(WF.ci_coeffs)
        out[f"{name}_ci_dagger"] = construct_ups_state(
            np.array(WF.ci_coeffs), WF.ci_info, thetas, layout, dagger=True
        )
        out[f"{name}_energy"] = np.array(WF.energy_elec)
        if name == "tups44":
            out["h2o_h_mo"] = np.array(WF.h_mo)
            out["h2o_g_mo"] = np.array(WF.g_mo)
            out["h2o_c_mo"] = np.array(c_mo)
            out["tups44_kappa_idx"] = np.array(WF.kappa_idx, dtype=np.int64)
        if "rdm" in record:
            out[f"{name}_rdm1"] = np.array(WF.rdm1)
            out[f"{name}_rdm2"] = np.array(WF.rdm2)
            out[f"{name}_energy_rdm"] = np.array(
                get_electronic_energy(WF.h_mo, WF.g_mo, WF.num_inactive_orbs, WF.num_active_orbs, WF.rdm1, WF.rdm2)
            )
        if "grad" in record:
            WF._old_opt_parameters = np.zeros(len(thetas) + len(WF.kappa_idx)) + 10**20
            if iak:
                params = WF.kappa + thetas
                out[f"{name}_gradient"] = WF._calc_gradient_optimization(params, True, True)
            else:
                out[f"{name}_gradient"] = WF._calc_gradient_optimization(thetas, True, False)
        if name == "tups44":
            out["tups44_orbital_gradient"] = get_orbital_gradient(
                WF.h_mo, WF.g_mo, WF.kappa_idx, WF.num_inactive_orbs, WF.num_active_orbs, WF.rdm1, WF.rdm2
            )
        if "steps" in record:
            rng2 = np.random.default_rng(5)
            probe = rng2.normal(size=WF.num_det)
            out[f"{name}_probe"] = probe
            picks = sorted(set([0, 1, len(thetas) // 2, len(thetas) - 1]))
            meta[name]["picks"] = picks
            for k in picks:
                out[f"{name}_unitary_{k}"] = propagate_unitary(probe, k, WF.ci_info, thetas, layout)
                out[f"{name}_gradaction_{k}"] = get_grad_action(probe, k, WF.ci_info, layout)
        print(name, "P =", len(thetas), "E =", WF.energy_elec, flush=True)


def golden_synthetic_ups(out: dict, meta: dict) -> None:
    """tUPS on spaces the wave-function class cannot make (unequal spins, more orbitals)."""
    for name, n, na, nb, L in [("syn_tups_5_23", 5, 2, 3, 2), ("syn_tups_6_33", 6, 3, 3, 3), ("syn_qnp_6_24", 6, 2, 4, 2)]:
        ci = get_indexing(0, n, 0, na, nb)
        lay = UpsStructure()
        opts = {"n_layers": L, "do_qnp": True} if "qnp" in name else {"n_layers": L, "do_tups": True}
        lay.create_tiled(n, opts)
        rng = np.random.default_rng(n * 100 + na * 10 + nb)
        thetas = rng.uniform(-np.pi, np.pi, lay.n_params).tolist()
        state = rng.normal(size=len(ci.idx2det))
        state /= np.linalg.norm(state)
        out[f"{name}_thetas"] = np.array(thetas)
        out[f"{name}_state"] = state
        out[f"{name}_result"] = construct_ups_state(state, ci, thetas, lay)
        out[f"{name}_result_dagger"] = construct_ups_state(state, ci, thetas, lay, dagger=True)
        meta[name] = {
            "n": n, "na": na, "nb": nb, "options": opts,
            "types": list(lay.excitation_operator_type),
            "indices": [[int(x) for x in t] for t in lay.excitation_indices],
        }
    # non-adjacent sa_single / pair doubles / generic doubles on (6; 3,3)
    ci = get_indexing(0, 6, 0, 3, 3)
    lay = UpsStructure()
    lay.create_fUCC([0, 1, 2], [3, 4, 5], [0, 1, 2, 3, 4, 5], [6, 7, 8, 9, 10, 11], 6,
                    {"n_layers": 1, "SAGS": True, "GpD": True, "S": True})
    rng = np.random.default_rng(77)
    thetas = rng.uniform(-1.0, 1.0, lay.n_params).tolist()
    state = rng.normal(size=len(ci.idx2det))
    state /= np.linalg.norm(state)
    out["syn_gsd_6_33_thetas"] = np.array(thetas)
    out["syn_gsd_6_33_state"] = state
    out["syn_gsd_6_33_result"] = construct_ups_state(state, ci, thetas, lay)
    meta["syn_gsd_6_33"] = {
        "n": 6, "na": 3, "nb": 3,
        "types": list(lay.excitation_operator_type),
        "indices": [[int(x) for x in t] for t in lay.excitation_indices],
    }


    # quintuple / sextuple / quadruple generators on (6; 3,3)
    lay = UpsStructure()
    lay.create_fUCC([0, 1, 2], [3, 4, 5], [0, 1, 2, 3, 4, 5], [6, 7, 8, 9, 10, 11], 6,
                    {"n_layers": 1, "Q": True, "5": True, "6": True})
    rng = np.random.default_rng(78)
    thetas = rng.uniform(-1.0, 1.0, lay.n_params).tolist()
    out["syn_q56_6_33_thetas"] = np.array(thetas)
    out["syn_q56_6_33_state"] = state
    out["syn_q56_6_33_result"] = construct_ups_state(state, ci, thetas, lay)
    meta["syn_q56_6_33"] = {
        "n": 6, "na": 3, "nb": 3,
        "types": list(lay.excitation_operator_type),
        "indices": [[int(x) for x in t] for t in lay.excitation_indices],
    }
    print("syn_q56 P =", lay.n_params, flush=True)


def golden_ucc(out: dict, meta: dict) -> None:
    """Non-factorised UCC state: expm_multiply of the dense T matrix (operator_state_algebra.py:870-896)."""
    from slowquant.unitary_coupled_cluster.operator_state_algebra import construct_ucc_state
    from slowquant.unitary_coupled_cluster.util import UccStructure

    ci = get_indexing(0, 4, 0, 2, 2)
    st = UccStructure()
    st.add_sa_singles([0, 1], [2, 3])
    st.add_sa_doubles([0, 1], [2, 3])
    st.add_triples([0, 1, 2, 3], [4, 5, 6, 7])
    st.add_quadruples([0, 1, 2, 3], [4, 5, 6, 7])
    rng = np.random.default_rng(4242)
    thetas = rng.uniform(-0.8, 0.8, st.n_params).tolist()
    hf = np.zeros(len(ci.idx2det))
    hf[0] = 1.0
    out["ucc44_thetas"] = np.array(thetas)
    out["ucc44_result"] = construct_ucc_state(hf, ci, thetas, st)
    out["ucc44_result_dagger"] = construct_ucc_state(out["ucc44_result"], ci, thetas, st, dagger=True)
    meta["ucc44"] = {
        "types": list(st.excitation_operator_type),
        "indices": [[int(x) for x in t] for t in st.excitation_indices],
    }
    print("ucc44 P =", st.n_params, flush=True)
    # wave-function object: H2O/STO-3G UCCSD(4,4) at fixed thetas (energy, RDMs, finite-difference theta gradient)
    from slowquant.unitary_coupled_cluster.ucc_wavefunction import WaveFunctionUCC

    SQobj = _h2o()
    WF = WaveFunctionUCC((4, 4), SQobj.hartree_fock.mo_coeff, SQobj, "SD")
    th = np.random.default_rng(99).uniform(-0.4, 0.4, len(WF.thetas)).tolist()
    WF.thetas = th
    out["uccwf_thetas"] = np.array(th)
    out["uccwf_ci"] = np.array(WF.ci_coeffs)
    out["uccwf_energy"] = np.array(WF.energy_elec)
    out["uccwf_rdm1"] = np.array(WF.rdm1)
    out["uccwf_rdm2"] = np.array(WF.rdm2)
    WF._old_opt_parameters = np.zeros(len(th)) + 10**20
    out["uccwf_gradient"] = WF._calc_gradient_optimization(th, True, False)
    meta["uccwf"] = {
        "types": list(WF.ucc_layout.excitation_operator_type),
        "indices": [[int(x) for x in t] for t in WF.ucc_layout.excitation_indices],
    }
    print("uccwf P =", len(th), "E =", WF.energy_elec, flush=True)


def main() -> None:
    arrays: dict = {}
    meta: dict = {}
    golden_indexing(arrays)
    ops = golden_operators()
    golden_propagate(arrays, meta)
    golden_synthetic_ups(arrays, meta)
    golden_wavefunctions(arrays, meta)
    golden_ucc(arrays, meta)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump({"meta": meta, "operators": ops}, f, indent=0)
    print("wrote", len(arrays), "arrays")


if __name__ == "__main__":
    main()
