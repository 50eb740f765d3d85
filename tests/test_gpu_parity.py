"""Parity of the CUDA path (through the Python shim -> C ABI -> sm_100a kernels) with the reference.

Three layers of evidence:
  1. golden vectors produced by running the reference itself (tests/golden/),
  2. the pinned CPU oracle (oracle/) on seeded inputs at sizes it finishes in seconds,
  3. size-independent properties at larger CAS (unitarity, adjoint round trip, antisymmetry of T).
Tolerances: fp64 amplitudes 1e-12 (north_star asks 1e-10 on energies / RDM elements).
"""
import numpy as np
import pytest
import torch

from conftest import json_to_opdict
from oracle import sq_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def sq():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import slowquant_b200.ci_spaces as ci
    import slowquant_b200.operator_state_algebra as osa
    from slowquant_b200 import _lib
    from slowquant_b200.fermionic_operator import FermionicOperator
    from slowquant_b200.util import UpsStructure

    class NS:
        pass

    ns = NS()
    ns.ci, ns.osa, ns.lib, ns.FermionicOperator, ns.UpsStructure = ci, osa, _lib, FermionicOperator, UpsStructure
    return ns


def _layout(sq, types, indices):
    lay = sq.UpsStructure()
    lay.excitation_operator_type = list(types)
    lay.excitation_indices = [tuple(t) for t in indices]
    lay.n_params = len(types)
    return lay


def test_native_library_is_loaded(sq):
    lib = sq.lib.load()
    before = lib.sq_launch_count()
    info = sq.ci.get_indexing(0, 4, 0, 2, 2)
    lay = sq.UpsStructure()
    lay.create_tiled(4, {"n_layers": 1, "do_tups": True})
    st = np.zeros(info.num_det)
    st[0] = 1.0
    out = sq.osa.construct_ups_state(st, info, [0.3] * lay.n_params, lay)
    assert abs(np.linalg.norm(out) - 1.0) < 1e-14
    assert lib.sq_launch_count() > before, "no CUDA kernel was launched"


def test_idx2det_view(sq, golden):
    arrays, _, _ = golden
    for key in arrays.files:
        if key.startswith("idx2det_"):
            n, na, nb = (int(x) for x in key.split("_")[1:])
            info = sq.ci.get_indexing(0, n, 0, na, nb)
            assert np.array_equal(info.idx2det, arrays[key])
            assert info.det2idx[int(arrays[key][-1])] == len(arrays[key]) - 1
            if 0 < na + nb < 2 * n:
                assert 0 not in info.det2idx
                with pytest.raises(KeyError):
                    info.det2idx[0]


SYN = ["syn_tups_5_23", "syn_tups_6_33", "syn_qnp_6_24", "syn_gsd_6_33", "syn_q56_6_33"]


@pytest.mark.parametrize("name", SYN)
def test_synthetic_ups_states(sq, golden, name):
    arrays, meta, _ = golden
    m = meta[name]
    info = sq.ci.get_indexing(0, m["n"], 0, m["na"], m["nb"])
    lay = _layout(sq, m["types"], m["indices"])
    th = arrays[f"{name}_thetas"].tolist()
    st = arrays[f"{name}_state"]
    keep = st.copy()
    res = sq.osa.construct_ups_state(st, info, th, lay)
    assert np.array_equal(st, keep), "input state was modified"
    assert np.max(np.abs(res - arrays[f"{name}_result"])) < TOL
    if f"{name}_result_dagger" in arrays.files:
        res = sq.osa.construct_ups_state(st, info, th, lay, dagger=True)
        assert np.max(np.abs(res - arrays[f"{name}_result_dagger"])) < TOL


WF = ["tups44", "qnp44", "fuccsd44", "sa44", "tq44", "gsd44", "ksa44", "sds44", "sad65"]


@pytest.mark.parametrize("name", WF)
def test_wavefunction_states(sq, golden, name):
    arrays, meta, _ = golden
    m = meta[name]
    info = sq.ci.get_indexing(m["num_inactive_orbs"], m["num_active_orbs"], m["num_virtual_orbs"], m["n_alpha"], m["n_beta"])
    lay = _layout(sq, m["types"], m["indices"])
    th = arrays[f"{name}_thetas"].tolist()
    ci = sq.osa.construct_ups_state(arrays[f"{name}_csf"], info, th, lay)
    assert np.max(np.abs(ci - arrays[f"{name}_ci"])) < TOL
    back = sq.osa.construct_ups_state(arrays[f"{name}_ci"], info, th, lay, dagger=True)
    assert np.max(np.abs(back - arrays[f"{name}_ci_dagger"])) < TOL
    # "U" / "Ud" entries of propagate_state (osa.py:525-552)
    via_u = sq.osa.propagate_state(["U"], arrays[f"{name}_csf"], info, th, lay)
    assert np.max(np.abs(via_u - arrays[f"{name}_ci"])) < TOL
    via_ud = sq.osa.propagate_state(["Ud"], arrays[f"{name}_ci"], info, th, lay)
    assert np.max(np.abs(via_ud - arrays[f"{name}_ci_dagger"])) < TOL
    for k in m.get("picks", []):
        probe = arrays[f"{name}_probe"]
        u = sq.osa.propagate_unitary(probe, k, info, th, lay)
        assert np.max(np.abs(u - arrays[f"{name}_unitary_{k}"])) < TOL, (name, k)
        ga = sq.osa.get_grad_action(probe, k, info, lay)
        assert np.max(np.abs(ga - arrays[f"{name}_gradaction_{k}"])) < TOL, (name, k)


def test_propagate_state_generic_operators(sq, golden):
    arrays, meta, _ = golden
    for name in ("p0", "p1", "p2", "p3"):
        info = sq.ci.get_indexing(*meta[name]["dims"])
        op = sq.FermionicOperator(json_to_opdict(meta[name]["op"]))
        res = sq.osa.propagate_state([op], arrays[f"{name}_state"], info, do_folding=False)
        assert np.max(np.abs(res - arrays[f"{name}_result"])) < TOL, name
    from slowquant_b200.operators import hamiltonian_0i_0a

    info = sq.ci.get_indexing(1, 3, 1, 2, 1)
    st = arrays["fold_state"]
    H = hamiltonian_0i_0a(arrays["fold_h"], arrays["fold_g"], 1, 3)
    assert np.max(np.abs(sq.osa.propagate_state([H], st, info) - arrays["fold_Hstate"])) < TOL
    assert abs(sq.osa.expectation_value(st, [H], st, info) - float(arrays["fold_energy"])) < TOL
    # same through explicit strings (generic gather kernel instead of the sigma kernel)
    Hs = sq.FermionicOperator(dict(H.operators))
    assert np.max(np.abs(sq.osa.propagate_state([Hs], st, info) - arrays["fold_Hstate"])) < TOL
    a = sq.FermionicOperator(json_to_opdict(meta["fold2"]["op_a"]))
    b = sq.FermionicOperator(json_to_opdict(meta["fold2"]["op_b"]))
    assert np.max(np.abs(sq.osa.propagate_state([a, b], st, info) - arrays["fold2_result"])) < TOL


def test_error_behaviour(sq):
    info = sq.ci.get_indexing(0, 3, 0, 1, 1)
    st = np.arange(info.num_det, dtype=float)
    out = sq.osa.propagate_state([], st, info)
    assert np.array_equal(out, st) and out is not st
    spin_flip = sq.FermionicOperator({((1, True), (0, False)): 1.0})
    with pytest.raises(KeyError):
        sq.osa.propagate_state([spin_flip], st, info, do_folding=False)
    assert np.all(sq.osa.propagate_state([spin_flip], st, info, do_folding=False, do_unsafe=True) == 0.0)
    with pytest.raises(ValueError):
        sq.osa.propagate_state(["X"], st, info)
    lay = sq.UpsStructure()
    lay.create_tiled(3, {"n_layers": 1, "do_tups": True})
    with pytest.raises(ValueError):
        sq.osa.propagate_state(["U"], st, info, None, lay)
    with pytest.raises(TypeError):
        sq.osa.propagate_state(["U"], st, info, [0.1], object())
    bad = _layout(sq, ["nonsense"], [(0, 1)])
    with pytest.raises(ValueError):
        sq.osa.construct_ups_state(st, info, [0.1], bad)
    with pytest.raises(ValueError):
        sq.osa.construct_ups_state(st[:-1], info, [0.1] * lay.n_params, lay)
    val = sq.osa.expectation_value(st, [], st, info)
    assert isinstance(val, float) and abs(val - float(st @ st)) < 1e-12


def test_gradient_sweep_matches_reference(sq, golden):
    arrays, meta, _ = golden
    from slowquant_b200.operators import hamiltonian_0i_0a

    for name in ("tups44", "fuccsd44", "sa44", "ksa44"):
        m = meta[name]
        nI, nA, nV = m["num_inactive_orbs"], m["num_active_orbs"], m["num_virtual_orbs"]
        info = sq.ci.get_indexing(nI, nA, nV, m["n_alpha"], m["n_beta"])
        lay = _layout(sq, m["types"], m["indices"])
        th = arrays[f"{name}_thetas"].tolist()
        H = hamiltonian_0i_0a(arrays["h2o_h_mo"], arrays["h2o_g_mo"], nI, nA)
        ci = arrays[f"{name}_ci"]
        e = sq.osa.expectation_value(ci, [H], ci, info)
        assert abs(e - float(arrays[f"{name}_energy"])) < 1e-10
        bra = sq.osa.propagate_state([H], ci, info)
        bra = sq.osa.construct_ups_state(bra, info, th, lay, dagger=True)
        g, bra_end, ket_end = sq.osa.ups_gradient_sweep(bra, arrays[f"{name}_csf"], info, th, lay)
        ref = arrays[f"{name}_gradient"]
        assert np.max(np.abs(g - ref[len(ref) - len(th):])) < 1e-10, name
        assert np.max(np.abs(ket_end - ci)) < TOL


def _seeded_case(n, na, nb, L, seed, qnp=False):
    types, idx = orc.tiled_layout(n, L, do_qnp=qnp)
    rng = np.random.default_rng(seed)
    th = rng.uniform(-np.pi, np.pi, len(types))
    return types, idx, th, rng


@pytest.mark.parametrize("n,na,nb,L", [(7, 3, 4, 2), (8, 4, 4, 3), (9, 5, 3, 2), (10, 5, 5, 1)])
def test_tups_against_oracle(sq, n, na, nb, L):
    types, idx, th, rng = _seeded_case(n, na, nb, L, 100 + n)
    sp = orc.get_indexing(0, n, 0, na, nb)
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    lay = _layout(sq, types, idx)
    st = rng.normal(size=sp.num_det)
    st /= np.linalg.norm(st)
    ref = orc.construct_ups_state(st, sp, th, types, idx, threaded=True)
    res = sq.osa.construct_ups_state(st, info, th.tolist(), lay)
    assert np.max(np.abs(res - ref)) < TOL
    # HF start (sparse light cone) and device-resident tensors
    hf = np.zeros(sp.num_det)
    hf[0] = 1.0
    ref = orc.construct_ups_state(hf, sp, th, types, idx)
    t = torch.from_numpy(hf).cuda()
    res_t = sq.osa.construct_ups_state(t, info, th.tolist(), lay)
    assert isinstance(res_t, torch.Tensor) and res_t.is_cuda
    assert torch.equal(t.cpu(), torch.from_numpy(hf)), "device input was modified"
    assert np.max(np.abs(res_t.cpu().numpy() - ref)) < TOL


def test_generic_generators_against_oracle(sq):
    n, na, nb = 7, 3, 3
    sp = orc.get_indexing(0, n, 0, na, nb)
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    rng = np.random.default_rng(17)
    types = ["single", "single", "double", "double", "double", "triple", "sa_single", "double", "sa_double_4", "sa_double_5"]
    idx = [(0, 8), (3, 13), (0, 1, 8, 13), (2, 4, 6, 12), (1, 5, 9, 11), (0, 1, 2, 8, 11, 12), (1, 5), (2, 3, 10, 11),
           (0, 1, 4, 6), (1, 2, 3, 5)]
    th = rng.uniform(-1.5, 1.5, len(types))
    st = rng.normal(size=sp.num_det)
    st /= np.linalg.norm(st)
    lay = _layout(sq, types, idx)
    ref = orc.construct_ups_state(st, sp, th, types, idx)
    res = sq.osa.construct_ups_state(st, info, th.tolist(), lay)
    assert np.max(np.abs(res - ref)) < 1e-11
    for k in range(len(types)):
        ga_ref = orc.get_grad_action(st, k, sp, types, idx)
        ga = sq.osa.get_grad_action(st, k, info, lay)
        assert np.max(np.abs(ga - ga_ref)) < TOL, (k, types[k])
        u_ref = orc.propagate_unitary(st, k, sp, th, types, idx)
        u = sq.osa.propagate_unitary(st, k, info, th.tolist(), lay)
        assert np.max(np.abs(u - u_ref)) < 1e-11, (k, types[k])


@pytest.mark.parametrize("n,ne", [(12, 6), (14, 7)])
def test_size_independent_properties(sq, n, ne):
    """Unitarity, adjoint round trip and <x|T|x> = 0 at sizes the oracle does not reach in seconds."""
    info = sq.ci.get_indexing(0, n, 0, ne, ne)
    lay = sq.UpsStructure()
    lay.create_tiled(n, {"n_layers": 2, "do_tups": True})
    rng = np.random.default_rng(n)
    th = rng.uniform(-np.pi, np.pi, lay.n_params).tolist()
    dev = torch.device("cuda", info.device)
    x = torch.randn(info.num_det, dtype=torch.float64, device=dev)
    x /= torch.linalg.norm(x)
    y = sq.osa.construct_ups_state(x, info, th, lay)
    assert abs(float(torch.linalg.norm(y)) - 1.0) < 1e-12
    back = sq.osa.construct_ups_state(y, info, th, lay, dagger=True)
    assert float(torch.max(torch.abs(back - x))) < 1e-12
    # overlap is preserved: <Ux|Uz> = <x|z>
    z = torch.randn(info.num_det, dtype=torch.float64, device=dev)
    uz = sq.osa.construct_ups_state(z, info, th, lay)
    assert abs(float(torch.dot(y, uz)) - float(torch.dot(x, z))) < 1e-10
    # generators are antisymmetric
    for k in (0, 1, lay.n_params - 1):
        tx = sq.osa.get_grad_action(x, k, info, lay)
        assert abs(float(torch.dot(x, tx))) < 1e-12
    # one unitary at a time equals the whole product
    w = x
    for k in range(6):
        w = sq.osa.propagate_unitary(w, k, info, th, lay)
    part = _layout(sq, lay.excitation_operator_type[:6], lay.excitation_indices[:6])
    w2 = sq.osa.construct_ups_state(x, info, th[:6], part)
    assert float(torch.max(torch.abs(w - w2))) < 1e-13


def test_rdms_against_reference(sq, golden):
    arrays, meta, _ = golden
    from slowquant_b200.density_matrix import get_electronic_energy, get_orbital_gradient
    from slowquant_b200.ups_wavefunction import symmetrize_rdm2_like_reference

    for name in ("tups44", "fuccsd44"):
        m = meta[name]
        nI, nA, nV = m["num_inactive_orbs"], m["num_active_orbs"], m["num_virtual_orbs"]
        info = sq.ci.get_indexing(nI, nA, nV, m["n_alpha"], m["n_beta"])
        ci = arrays[f"{name}_ci"]
        d1, d2 = sq.osa.reduced_density_matrices(ci, ci, info)
        assert np.max(np.abs(d1 - arrays[f"{name}_rdm1"])) < 1e-12
        assert np.max(np.abs(d2 - arrays[f"{name}_rdm2"])) < 1e-12
        assert np.max(np.abs(symmetrize_rdm2_like_reference(d2) - arrays[f"{name}_rdm2"])) < 1e-12
        assert abs(np.trace(d1) - (m["n_alpha"] + m["n_beta"])) < 1e-12
        e = get_electronic_energy(arrays["h2o_h_mo"], arrays["h2o_g_mo"], nI, nA, d1, d2)
        assert abs(e - float(arrays[f"{name}_energy_rdm"])) < 1e-10
        assert abs(e - float(arrays[f"{name}_energy"])) < 1e-10
    og = get_orbital_gradient(
        arrays["h2o_h_mo"], arrays["h2o_g_mo"], arrays["tups44_kappa_idx"], 3, 4, arrays["tups44_rdm1"], arrays["tups44_rdm2"]
    )
    assert np.max(np.abs(og - arrays["tups44_orbital_gradient"])) < 1e-11


def test_transition_rdm_against_oracle(sq):
    n, na, nb = 5, 2, 3
    sp = orc.get_indexing(0, n, 0, na, nb)
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    rng = np.random.default_rng(23)
    bra = rng.normal(size=sp.num_det)
    ket = rng.normal(size=sp.num_det)
    d1, d2 = sq.osa.reduced_density_matrices(bra, ket, info)
    for p, q in [(0, 0), (1, 3), (4, 2), (2, 4)]:
        ref = orc.expectation_value(bra, [orc.Epq(p, q)], ket, sp)
        assert abs(d1[p, q] - ref) < 1e-12
    for p, q, r, s in [(0, 1, 1, 0), (3, 1, 2, 4), (4, 4, 2, 2), (1, 2, 2, 3), (0, 3, 4, 1)]:
        ref = orc.expectation_value(bra, [orc.op_mul(orc.Epq(p, q), orc.Epq(r, s))], ket, sp)
        if q == r:
            ref -= orc.expectation_value(bra, [orc.Epq(p, s)], ket, sp)
        assert abs(d2[p, q, r, s] - ref) < 1e-12, (p, q, r, s)


@pytest.mark.parametrize("n,na,nb", [(6, 3, 3), (7, 4, 2), (8, 4, 4), (13, 2, 1)])
def test_sigma_against_oracle(sq, n, na, nb):
    """sigma and <H> against the oracle's string-by-string restatement (osa.py:596-628).  n = 7: no permutational symmetry;
    n = 13 (unsymmetric, 169 generator rows): more rows than one CTA of the DMMA kernel holds (row tiles along gridDim.y), an odd
    row count (padded leading dimension); its RDMs use a 2 x 2-tile Gram matrix (upper triangle + mirror)."""
    from slowquant_b200.operators import hamiltonian_0i_0a

    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    if n in (7, 13):  # no permutational symmetry at all: the kernel must not assume any
        h = rng.normal(size=(n, n))
        g = 0.1 * rng.normal(size=(n, n, n, n))
    sp = orc.get_indexing(0, n, 0, na, nb)
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    st = rng.normal(size=sp.num_det)
    st /= np.linalg.norm(st)
    H = hamiltonian_0i_0a(h, g, 0, n)
    sig = sq.osa.propagate_state([H], st, info)
    ref = orc.propagate_state([orc.hamiltonian_0i_0a(h, g, 0, n)], st, sp, threaded=True)
    assert np.max(np.abs(sig - ref)) < 1e-11
    e = sq.osa.expectation_value(st, [H], st, info)
    assert abs(e - float(st @ ref)) < 1e-11
    if n != 7:
        if n == 13:   # symmetric integrals again (91 symmetrised rows): <H> through sigma against <H> through the RDMs
            h, g = A + A.T, B + B.transpose(1, 0, 2, 3)
            g = g + g.transpose(0, 1, 3, 2)
            g = g + g.transpose(2, 3, 0, 1)
            e = sq.osa.expectation_value(st, [hamiltonian_0i_0a(h, g, 0, n)], st, info)
        d1, d2 = sq.osa.reduced_density_matrices(st, st, info)
        e_rdm = float(np.sum(h * d1) + 0.5 * np.sum(g * d2))
        assert abs(e - e_rdm) < 1e-10
        other = rng.normal(size=sp.num_det)
        t1, t2 = sq.osa.reduced_density_matrices(other, st, info)      # transition RDM: all Gram tiles, no mirror
        assert abs(np.trace(t1) - (na + nb) * float(other @ st)) < 1e-10
        assert abs(float(np.einsum("ppqq->", t2)) - (na + nb) * (na + nb - 1) * float(other @ st)) < 1e-9


def test_wavefunction_object(sq, golden):
    """WaveFunctionUPS surface on the H2O/STO-3G CAS(4,4) integrals exported from the reference run."""
    arrays, meta, _ = golden
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.ups_wavefunction import WaveFunctionUPS

    ints = ArrayIntegrals(arrays["h2o_h_mo"], arrays["h2o_g_mo"], num_elec=10)
    eye = np.eye(arrays["h2o_h_mo"].shape[0])
    WF = WaveFunctionUPS((4, 4), eye, ints, "tUPS", {"n_layers": 2}, include_active_kappa=True)
    m = meta["tups44"]
    assert WF.ups_layout.excitation_operator_type == m["types"]
    assert np.array_equal(WF.kappa_idx, arrays["tups44_kappa_idx"])
    assert np.array_equal(WF.csf_coeffs, arrays["tups44_csf"])
    th = arrays["tups44_thetas"].tolist()
    WF.thetas = th
    assert WF.thetas == th
    assert np.max(np.abs(WF.ci_coeffs - arrays["tups44_ci"])) < TOL
    assert abs(WF.energy_elec - float(arrays["tups44_energy"])) < 1e-10
    assert np.max(np.abs(WF.rdm1 - arrays["tups44_rdm1"])) < 1e-12
    assert np.max(np.abs(WF.rdm2 - arrays["tups44_rdm2"])) < 1e-12
    params = WF.kappa + th
    grad = WF._calc_gradient_optimization(params, True, True)
    assert np.max(np.abs(grad - arrays["tups44_gradient"])) < 1e-10
    e_opt = WF._calc_energy_optimization(params, True, True)
    assert abs(e_opt - float(arrays["tups44_energy_rdm"])) < 1e-10
    with pytest.raises(ValueError):
        WF.thetas = th[:-1]
    with pytest.raises(ValueError):
        WaveFunctionUPS((4, 4), eye, ints, "nonsense")
    # gradient is consistent with a central finite difference of the energy
    k = 4
    step = 1e-5
    tp, tm = list(th), list(th)
    tp[k] += step
    tm[k] -= step
    WF.thetas = tp
    ep = WF.energy_elec
    WF.thetas = tm
    em = WF.energy_elec
    assert abs((ep - em) / (2 * step) - grad[len(WF.kappa) + k]) < 1e-7


def test_wavefunction_optimisation_reaches_reference_energy(sq, golden):
    """Reference known answer (tests/test_unitary_product_state.py:362-394): pp-tUPS(4,4), 1 layer, on
    H2O/STO-3G converges to -83.96387402720552 Eh (tolerance 1e-6 there)."""
    arrays, _, _ = golden
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.ups_wavefunction import WaveFunctionUPS

    ints = ArrayIntegrals(arrays["h2o_h_mo"], arrays["h2o_g_mo"], num_elec=10)
    eye = np.eye(arrays["h2o_h_mo"].shape[0])
    WF = WaveFunctionUPS((4, 4), eye, ints, "tUPS", {"n_layers": 1, "do_pp": True})
    WF.run_wf_optimization_1step("bfgs", orbital_optimization=False)
    assert abs(WF.energy_elec + 83.96387402720552) < 1e-6


def test_ucc_state_against_reference(sq, golden):
    arrays, meta, _ = golden
    from slowquant_b200.util import UccStructure

    info = sq.ci.get_indexing(0, 4, 0, 2, 2)
    st = UccStructure()
    st.add_sa_singles([0, 1], [2, 3])
    st.add_sa_doubles([0, 1], [2, 3])
    st.add_triples([0, 1, 2, 3], [4, 5, 6, 7])
    st.add_quadruples([0, 1, 2, 3], [4, 5, 6, 7])
    m = meta["ucc44"]
    assert st.excitation_operator_type == m["types"]
    assert [list(t) for t in st.excitation_indices] == m["indices"]
    th = arrays["ucc44_thetas"].tolist()
    hf = np.zeros(info.num_det)
    hf[0] = 1.0
    res = sq.osa.construct_ucc_state(hf, info, th, st)
    assert np.max(np.abs(res - arrays["ucc44_result"])) < 1e-12
    back = sq.osa.construct_ucc_state(arrays["ucc44_result"], info, th, st, dagger=True)
    assert np.max(np.abs(back - arrays["ucc44_result_dagger"])) < 1e-12
    via_u = sq.osa.propagate_state(["U"], hf, info, th, st)
    assert np.max(np.abs(via_u - arrays["ucc44_result"])) < 1e-12


def test_build_operator_matrix(sq):
    from slowquant_b200.operators import G2_sa

    info = sq.ci.get_indexing(0, 4, 0, 2, 2)
    sp = orc.get_indexing(0, 4, 0, 2, 2)
    T = G2_sa(0, 1, 2, 3, 4, True)
    mat = sq.osa.build_operator_matrix(T, info)
    assert np.max(np.abs(mat + mat.T)) < 1e-15
    v = np.random.default_rng(2).normal(size=sp.num_det)
    ref = orc.propagate_state([orc.G2_sa(0, 1, 2, 3, 4)], v, sp, do_folding=False)
    assert np.max(np.abs(mat @ v - ref)) < 1e-13


def test_config2_n2_ccpvdz_cas1010(sq):
    """BASELINE.json config 2: N2 / cc-pVDZ oo-tUPS CAS(10,10), 63 504 determinants, real integrals exported
    from a run of the reference (tests/golden/make_golden_n2.py): state, energy (string and RDM paths),
    1-/2-RDM, orbital gradient (236 kappa), theta gradient."""
    import os

    from conftest import ROOT
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.ups_wavefunction import WaveFunctionUPS

    d = np.load(os.path.join(ROOT, "tests", "golden", "golden_n2.npz"))
    nI, nA, nV, na, nb = (int(x) for x in d["dims"])
    N, M = nI + nA + nV, nI + nA
    g = np.zeros((N, N, N, N))
    g[:, :M, :M, :M] = d["g_npqr"]
    g[:M, :, :M, :M] = d["g_pnqr"]
    WF = WaveFunctionUPS((10, 10), np.eye(N), ArrayIntegrals(d["h_mo"], g, num_elec=14), "tUPS", {"n_layers": 2})
    assert (WF.num_inactive_orbs, WF.num_active_orbs, WF.num_virtual_orbs) == (nI, nA, nV)
    assert WF.num_det == 63504
    assert np.array_equal(WF.kappa_idx, d["kappa_idx"])
    th = d["thetas"].tolist()
    WF.thetas = th
    ci = WF.ci_coeffs
    ref_ci = np.zeros(WF.num_det)
    ref_ci[d["ci_nonzero_idx"]] = d["ci_nonzero_val"]
    assert np.max(np.abs(ci - ref_ci)) < 1e-13
    assert abs(WF.energy_elec - float(d["energy_strings"])) < 1e-10
    assert np.max(np.abs(WF.rdm1 - d["rdm1"])) < 1e-11
    assert np.max(np.abs(WF.rdm2 - d["rdm2"])) < 1e-11
    from slowquant_b200.density_matrix import get_electronic_energy, get_orbital_gradient

    e_rdm = get_electronic_energy(WF.h_mo, WF.g_mo, nI, nA, WF.rdm1, WF.rdm2)
    assert abs(e_rdm - float(d["energy_rdm"])) < 1e-10
    og = get_orbital_gradient(WF.h_mo, WF.g_mo, WF.kappa_idx, nI, nA, WF.rdm1, WF.rdm2)
    assert np.max(np.abs(og - d["orbital_gradient"])) < 1e-10
    tg = WF._calc_gradient_optimization(th, True, False)
    assert np.max(np.abs(tg - d["theta_gradient"])) < 1e-10


def test_config3_fuccsd_against_oracle_and_invariants(sq):
    """BASELINE.json config 3 shape (fUCCSD: all singles then all doubles through the generic Givens kernel):
    a 60-operator slice against the oracle at CAS(10,10), and invariants of the full 3381-operator ansatz at
    CAS(14,14) (11.8M determinants): unit norm, adjoint round trip, E_sigma == E_RDM, Tr rdm1 = N_e."""
    from slowquant_b200.operators import hamiltonian_0i_0a

    # (a) oracle slice
    n, ne = 10, 5
    lay = sq.UpsStructure()
    occ_s, unocc_s = list(range(2 * ne)), list(range(2 * ne, 2 * n))
    lay.create_fUCC(list(range(ne)), list(range(ne, n)), occ_s, unocc_s, n, {"n_layers": 1, "S": True, "D": True})
    pick = list(range(0, 30)) + list(range(lay.n_params - 30, lay.n_params))
    types = [lay.excitation_operator_type[k] for k in pick]
    idx = [lay.excitation_indices[k] for k in pick]
    rng = np.random.default_rng(31)
    th = rng.uniform(-1.0, 1.0, len(pick))
    sp = orc.get_indexing(0, n, 0, ne, ne)
    info = sq.ci.get_indexing(0, n, 0, ne, ne)
    st = rng.normal(size=sp.num_det)
    st /= np.linalg.norm(st)
    ref = orc.construct_ups_state(st, sp, th, types, idx, threaded=True)
    res = sq.osa.construct_ups_state(st, info, th.tolist(), _layout(sq, types, idx))
    assert np.max(np.abs(res - ref)) < 1e-12
    # (b) full ansatz at CAS(14,14)
    n, ne = 14, 7
    lay = sq.UpsStructure()
    lay.create_fUCC(list(range(ne)), list(range(ne, n)), list(range(2 * ne)), list(range(2 * ne, 2 * n)), n,
                    {"n_layers": 1, "S": True, "D": True})
    assert lay.n_params == 3381 and lay.excitation_operator_type.count("single") == 98
    info = sq.ci.get_indexing(0, n, 0, ne, ne)
    th = (0.05 * np.random.default_rng(14).uniform(-1, 1, lay.n_params)).tolist()
    dev = torch.device("cuda", info.device)
    hf = torch.zeros(info.num_det, dtype=torch.float64, device=dev)
    hf[0] = 1.0
    psi = sq.osa.construct_ups_state(hf, info, th, lay)
    assert abs(float(torch.linalg.norm(psi)) - 1.0) < 1e-12
    back = sq.osa.construct_ups_state(psi, info, th, lay, dagger=True)
    assert float(torch.max(torch.abs(back - hf))) < 1e-12
    rng = np.random.default_rng(2024)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    e = sq.osa.expectation_value(psi, [hamiltonian_0i_0a(h, g, 0, n)], psi, info)
    d1, d2 = sq.osa.reduced_density_matrices(psi, psi, info)
    assert abs(np.trace(d1) - 2 * ne) < 1e-11
    assert abs(e - float(np.sum(h * d1) + 0.5 * np.sum(g * d2))) < 1e-10
    assert np.max(np.abs(d2 - d2.transpose(2, 3, 0, 1))) < 1e-12


def test_ucc_wavefunction_object(sq, golden):
    """WaveFunctionUCC surface (ucc_wavefunction.py): UCCSD(4,4) on the H2O integrals at fixed thetas."""
    arrays, meta, _ = golden
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.ucc_wavefunction import WaveFunctionUCC

    ints = ArrayIntegrals(arrays["h2o_h_mo"], arrays["h2o_g_mo"], num_elec=10)
    WF = WaveFunctionUCC((4, 4), np.eye(arrays["h2o_h_mo"].shape[0]), ints, "SD")
    assert WF.ucc_layout.excitation_operator_type == meta["uccwf"]["types"]
    assert [list(t) for t in WF.ucc_layout.excitation_indices] == meta["uccwf"]["indices"]
    th = arrays["uccwf_thetas"].tolist()
    WF.thetas = th
    assert np.max(np.abs(WF.ci_coeffs - arrays["uccwf_ci"])) < 1e-12
    assert abs(WF.energy_elec - float(arrays["uccwf_energy"])) < 1e-10
    assert np.max(np.abs(WF.rdm1 - arrays["uccwf_rdm1"])) < 1e-11
    assert np.max(np.abs(WF.rdm2 - arrays["uccwf_rdm2"])) < 1e-11
    # forward finite differences with step sqrt(eps) ~ 1.5e-8 on E ~ -84 Eh: both the reference's numbers and
    # ours are quantised in units of 2 ulp(E)/step ~ 1.9e-6, so agreement is only meaningful to a few units
    grad = WF._calc_gradient_optimization(th, True, False)
    assert np.max(np.abs(grad - arrays["uccwf_gradient"])) < 3e-5


WIN_VARIANTS = ["1", "6:5:4,113,5,16,3", "6:5:4,72,5,16,3", "6:4:0,60,2,16,2", "4:0:0,40,1,5,2", "5:0:0,200,4,16,1", "6:0:0,100,0,7,1",
                "5:3:0,100,3,16,2", "8:7:3,220,0,9,1"]


@pytest.mark.parametrize("n,na,nb,L,qnp", [(8, 4, 4, 3, False), (9, 4, 5, 2, True), (12, 6, 6, 3, False), (13, 6, 5, 2, False)])
def test_window_sweeps_equal_single_brick_launches(sq, n, na, nb, L, qnp):
    """Every window configuration of the planner (alpha/beta windows, run and block modes, brick caps) gives
    the state of the one-brick-per-launch path; at small sizes that path is also checked against the oracle."""
    import ctypes as C

    lib = sq.lib.load()
    types, idx, th, rng = _seeded_case(n, na, nb, L, 500 + n, qnp=qnp)
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    lay = _layout(sq, types, idx)
    handle = sq.osa.compile_layout(info, lay)
    dev = torch.device("cuda", info.device)
    x = torch.randn(info.num_det, dtype=torch.float64, device=dev)
    x /= torch.linalg.norm(x)
    try:
        sq.lib.check(lib.sq_set_option(b"win", b"0"))
        ref = sq.osa.construct_ups_state(x, info, th.tolist(), lay)
        ref_d = sq.osa.construct_ups_state(x, info, th.tolist(), lay, dagger=True)
        if n <= 9:
            sp = orc.get_indexing(0, n, 0, na, nb)
            o = orc.construct_ups_state(x.cpu().numpy(), sp, th, types, idx, threaded=True)
            assert np.max(np.abs(ref.cpu().numpy() - o)) < TOL
        used = 0
        # both window kernels: win_kernel (default) and win3_kernel (register blocks over orbital triples, merged tiles)
        for cfg, w3 in [(c, w) for c in WIN_VARIANTS for w in (b"1", b"0")]:
            sq.lib.check(lib.sq_set_option(b"win3", w3))
            sq.lib.check(lib.sq_set_option(b"win", cfg.encode()))
            stats = (C.c_int64 * 6)()
            sq.lib.check(lib.sq_layout_plan_stats(handle, 0, len(types), stats))
            used += int(stats[1])
            res = sq.osa.construct_ups_state(x, info, th.tolist(), lay)
            assert float(torch.max(torch.abs(res - ref))) < 1e-13, (cfg, w3)
            res_d = sq.osa.construct_ups_state(x, info, th.tolist(), lay, dagger=True)
            assert float(torch.max(torch.abs(res_d - ref_d))) < 1e-13, (cfg, w3)
            # a sub-range of the circuit (propagate_unitary-style first/last)
            k0, k1 = 2, len(types) - 1
            part = _layout(sq, types[k0:k1], idx[k0:k1])
            sq.lib.check(lib.sq_set_option(b"win", b"0"))
            ref_p = sq.osa.construct_ups_state(x, info, th[k0:k1].tolist(), part)
            sq.lib.check(lib.sq_set_option(b"win", cfg.encode()))
            res_p = sq.osa.construct_ups_state(x, info, th[k0:k1].tolist(), part)
            assert float(torch.max(torch.abs(res_p - ref_p))) < 1e-13, cfg
        assert used > 0, "no window sweep was planned"
    finally:
        sq.lib.check(lib.sq_set_option(b"win", b"1"))
        sq.lib.check(lib.sq_set_option(b"win3", b"0"))


@pytest.mark.parametrize("n,na,nb,L,qnp", [(8, 4, 4, 2, False), (9, 5, 4, 2, True), (12, 6, 6, 2, False)])
def test_window_gradient_sweep(sq, n, na, nb, L, qnp):
    """theta-gradient sweep through the window kernel (bricks differentiated inside the sweeps, commuting bricks
    reordered) against the one-brick-per-launch path; at small sizes that path is checked against the oracle's
    literal loop of ups_wavefunction.py:1114-1138.  Includes a zero theta (gradient without rotation)."""
    lib = sq.lib.load()
    types, idx, th, rng = _seeded_case(n, na, nb, L, 900 + n, qnp=qnp)
    th[3] = 0.0
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    lay = _layout(sq, types, idx)
    dev = torch.device("cuda", info.device)
    bra = torch.randn(info.num_det, dtype=torch.float64, device=dev)
    ket = torch.randn(info.num_det, dtype=torch.float64, device=dev)
    bra /= torch.linalg.norm(bra)
    ket /= torch.linalg.norm(ket)
    try:
        sq.lib.check(lib.sq_set_option(b"wingrad", b"0"))
        g0, b0, k0 = sq.osa.ups_gradient_sweep(bra, ket, info, th.tolist(), lay)
        if n <= 9:
            sp = orc.get_indexing(0, n, 0, na, nb)
            b_np, k_np = bra.cpu().numpy().copy(), ket.cpu().numpy().copy()
            gr = np.zeros(len(types))
            for k in range(len(types)):   # the literal loop of ups_wavefunction.py:1114-1138
                gr[k] = 2.0 * float(b_np @ orc.get_grad_action(k_np, k, sp, types, idx))
                b_np = orc.propagate_unitary(b_np, k, sp, th, types, idx)
                k_np = orc.propagate_unitary(k_np, k, sp, th, types, idx)
            assert np.max(np.abs(g0 - gr)) < 1e-11
            assert np.max(np.abs(k0.cpu().numpy() - k_np)) < TOL
        # g0 came from the default plan (quad_grad_kernel: two commuting bricks per launch); one brick per launch must agree
        sq.lib.check(lib.sq_set_option(b"quadgrad", b"0"))
        launches_before = lib.sq_launch_count()
        gs, bs, ks = sq.osa.ups_gradient_sweep(bra, ket, info, th.tolist(), lay)
        launches_single = lib.sq_launch_count() - launches_before
        sq.lib.check(lib.sq_set_option(b"quadgrad", b"1"))
        launches_before = lib.sq_launch_count()
        sq.osa.ups_gradient_sweep(bra, ket, info, th.tolist(), lay)
        assert lib.sq_launch_count() - launches_before < launches_single, "no quad gradient launch was used"
        assert np.max(np.abs(gs - g0)) < 1e-12
        assert float(torch.max(torch.abs(bs - b0))) < 1e-13 and float(torch.max(torch.abs(ks - k0))) < 1e-13
        sq.lib.check(lib.sq_set_option(b"wingrad", b"1"))
        # the gradient sweep plans with its own window configuration (two vectors per batch): the default, wide windows (one
        # CTA per SM), capped brick counts, narrow windows
        for cfg in ("5:4:0,40,3,16,2", "1", "5:4:0,72,2,16,2", "6:0:0,100,0,4,1", "4:3:0,40,1,16,1"):
            sq.lib.check(lib.sq_set_option(b"wingrad_win", cfg.encode()))
            g1, b1, k1 = sq.osa.ups_gradient_sweep(bra, ket, info, th.tolist(), lay)
            assert np.max(np.abs(g1 - g0)) < 1e-12, cfg
            assert float(torch.max(torch.abs(b1 - b0))) < 1e-13 and float(torch.max(torch.abs(k1 - k0))) < 1e-13, cfg
    finally:
        sq.lib.check(lib.sq_set_option(b"wingrad_win", b"5:4:0,40,3,16,2"))
        sq.lib.check(lib.sq_set_option(b"wingrad", b"0"))
        sq.lib.check(lib.sq_set_option(b"quadgrad", b"1"))


def test_state_averaged_twins(sq):
    """`_SA` batch functions (osa.py:633-781, 827-867, 1415-1864, 2312-2976) on [n_states, N_det] batches: host batches
    take the stream-pipelined path, device batches ONE launch sequence with the state index as a batch dimension of the window
    and gauge sweeps (sq_ups_apply_batch); both against the oracle state by state."""
    n, na, nb = 8, 4, 4
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    sp = orc.get_indexing(0, n, 0, na, nb)
    lay = sq.UpsStructure()
    lay.create_tiled(n, {"n_layers": 2, "do_tups": True})
    rng = np.random.default_rng(99)
    th = rng.uniform(-np.pi, np.pi, lay.n_params)
    S = 5  # more states than pipeline buffers
    states = rng.normal(size=(S, info.num_det))
    states /= np.linalg.norm(states, axis=1)[:, None]
    keep = states.copy()
    ref = np.array([orc.construct_ups_state(s, sp, th, lay.excitation_operator_type, lay.excitation_indices) for s in states])
    out = sq.osa.construct_ups_state_SA(states, info, th.tolist(), lay)
    assert out.shape == ref.shape and np.max(np.abs(out - ref)) < TOL
    assert np.array_equal(states, keep), "inputs must not be modified"
    dev_in = torch.from_numpy(states).cuda()
    lib = sq.lib.load()
    before = lib.sq_launch_count()
    out_dev = sq.osa.construct_ups_state_SA(dev_in, info, th.tolist(), lay)
    batched_launches = lib.sq_launch_count() - before
    assert isinstance(out_dev, torch.Tensor) and np.max(np.abs(out_dev.cpu().numpy() - ref)) < TOL
    assert torch.equal(dev_in.cpu(), torch.from_numpy(keep)), "device inputs must not be modified"
    before = lib.sq_launch_count()
    single = sq.osa.construct_ups_state(dev_in[2], info, th.tolist(), lay)
    single_launches = lib.sq_launch_count() - before
    assert torch.equal(single, out_dev[2]), "a batched state must be bit-identical to the single-state call"
    assert batched_launches == single_launches, "the batch must ride in the launches of one state (batch dimension inside the kernels)"
    u_dev = sq.osa.propagate_unitary_SA(dev_in, 4, info, th.tolist(), lay)
    assert u_dev.is_cuda and u_dev.shape == dev_in.shape
    back = sq.osa.construct_ups_state_SA(out, info, th.tolist(), lay, dagger=True)
    assert np.max(np.abs(back - states)) < TOL
    one = sq.osa.construct_ups_state_SA(states[:1], info, th.tolist(), lay)
    assert one.shape == (1, info.num_det) and np.max(np.abs(one[0] - ref[0])) < TOL
    k = 4
    ref_u = np.array([orc.propagate_unitary(s, k, sp, th, lay.excitation_operator_type, lay.excitation_indices) for s in states])
    assert np.max(np.abs(sq.osa.propagate_unitary_SA(states, k, info, th.tolist(), lay) - ref_u)) < TOL
    assert np.max(np.abs(u_dev.cpu().numpy() - ref_u)) < TOL
    ref_g = np.array([orc.get_grad_action(s, k, sp, lay.excitation_operator_type, lay.excitation_indices) for s in states])
    assert np.max(np.abs(sq.osa.get_grad_action_SA(states, k, info, lay) - ref_g)) < TOL
    from slowquant_b200 import operators as mops

    op = mops.Epq(1, 2) * mops.Epq(5, 3) + 0.3 * mops.Epq(0, 6)
    ref_p = np.array([orc.propagate_state([dict(op.operators)], s, sp) for s in states])
    got_p = sq.osa.propagate_state_SA([op], states, info)
    assert np.max(np.abs(got_p - ref_p)) < TOL
    ev = sq.osa.expectation_value_SA(states, [op], states, info)
    assert abs(ev - float(np.mean([s @ r for s, r in zip(states, ref_p)]))) < 1e-12


def test_sigma_and_rdm_kernel_variants_agree(sq):
    """The sigma / RDM path has run-time variants (sq_set_option): row-per-CTA or determinant-per-thread panel kernels,
    E table in shared or constant memory or per-string partner tables (etab = tab), panels pipelined over internal streams or
    one at a time.  With small panels (so
    that several panels, a partial last panel and rows split between panels all occur) every variant must give the oracle's
    sigma vector and identical RDMs."""
    from slowquant_b200.operators import hamiltonian_0i_0a

    lib = sq.lib.load()
    n, na, nb = 8, 4, 3
    rng = np.random.default_rng(7)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    g_unsym = g + 0.05 * rng.normal(size=(n, n, n, n))      # takes the general n^2 path
    sp = orc.get_indexing(0, n, 0, na, nb)
    state = rng.normal(size=sp.num_det)
    state /= np.linalg.norm(state)
    refs = {}
    for name, gg in (("sym", g), ("unsym", g_unsym)):
        refs[name] = orc.propagate_state([orc.hamiltonian_0i_0a(h, gg, 0, n)], state, sp)
    results = []
    other = rng.normal(size=sp.num_det)
    try:
        for rows, etab, pipe, fused in ((b"1", b"smem", b"1", b"0"), (b"1", b"smem", b"0", b"0"), (b"0", b"smem", b"1", b"0"),
                                        (b"0", b"const", b"0", b"0"), (b"0", b"smem", b"1", b"1"), (b"0", b"tab", b"1", b"0"),
                                        (b"0", b"tab", b"0", b"0")):
            lib.sq_set_option(b"panel", b"768")            # 3920 determinants -> 6 panels, the last one partial
            # 2-RDM of one vector: the plain n^2 x n^2 Gram matrix in the first two rounds, the two symmetric S / A Gram matrices
            # (the default) in the others -- all must agree to 1e-12
            lib.sq_set_option(b"rdm_sym", b"0" if rows == b"1" else b"1")
            lib.sq_set_option(b"sigma_fused", fused)       # "1": the fused gather -> DMMA -> scatter kernel (off by default: slower)
            lib.sq_set_option(b"rows", rows)
            lib.sq_set_option(b"etab", etab)
            lib.sq_set_option(b"pipeline", pipe)
            info = sq.ci.get_indexing(0, n, 0, na, nb)      # a fresh space picks up the panel width
            for name, gg in (("sym", g), ("unsym", g_unsym)):
                out = sq.osa.propagate_state([hamiltonian_0i_0a(h, gg, 0, n)], state, info)
                assert np.max(np.abs(out - refs[name])) < 1e-11, (rows, etab, pipe, fused, name)
            d1, d2 = sq.osa.reduced_density_matrices(state, state, info)
            t1, t2 = sq.osa.reduced_density_matrices(other, state, info)
            results.append((d1, d2, t1, t2))
            assert abs(np.trace(d1) - (na + nb)) < 1e-12
    finally:
        for name, val in ((b"panel", b"0"), (b"rows", b"0"), (b"etab", b"smem"), (b"pipeline", b"1"), (b"sigma_fused", b"0"), (b"rdm_sym", b"1")):
            lib.sq_set_option(name, val)
    for other in results[1:]:
        for x, y in zip(results[0], other):
            assert np.max(np.abs(x - y)) < 1e-12


def test_rdm3_rdm4_against_reference(sq, golden):
    """rdm3 / rdm4 (ups_wavefunction.py:478-754) from the E E|psi> panel against the reference's loops: CAS(4,4) rdm3,
    CAS(4,3) rdm3 + rdm4, and CAS(2,3) where both must vanish (every contraction term has to cancel)."""
    import os

    from conftest import ROOT
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.ups_wavefunction import WaveFunctionUPS

    arrays, _, _ = golden
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_rdm34.npz"))
    ints = ArrayIntegrals(arrays["h2o_h_mo"], arrays["h2o_g_mo"], num_elec=10)
    eye = np.eye(arrays["h2o_h_mo"].shape[0])
    WF = WaveFunctionUPS((4, 4), eye, ints, "tUPS", {"n_layers": 2}, include_active_kappa=True)
    WF.thetas = arrays["tups44_thetas"].tolist()
    assert np.max(np.abs(WF.rdm3 - g["cas44_rdm3"])) < 1e-10
    # partial trace: sum_t Gamma3[pq rs tt] = (N - 2) Gamma2[pq rs]
    assert np.max(np.abs(np.einsum("pqrstt->pqrs", WF.rdm3) - 2 * WF.rdm2)) < 1e-10
    WF3 = WaveFunctionUPS((4, 3), eye, ints, "tUPS", {"n_layers": 2}, include_active_kappa=True)
    WF3.thetas = g["cas43_thetas"].tolist()
    assert np.max(np.abs(WF3.rdm3 - g["cas43_rdm3"])) < 1e-10
    assert np.max(np.abs(WF3.rdm4 - g["cas43_rdm4"])) < 1e-10
    assert np.max(np.abs(np.einsum("pqrstumm->pqrstu", WF3.rdm4) - 1 * WF3.rdm3)) < 1e-10
    WF2 = WaveFunctionUPS((2, 3), eye, ints, "tUPS", {"n_layers": 2}, include_active_kappa=True)
    WF2.thetas = g["cas23_thetas"].tolist()
    assert np.max(np.abs(WF2.ci_coeffs - g["cas23_ci"])) < TOL
    assert np.max(np.abs(WF2.rdm3)) < 1e-12 and np.max(np.abs(WF2.rdm4)) < 1e-12
    assert np.max(np.abs(WF2.rdm3 - g["cas23_rdm3"])) < 1e-12 and np.max(np.abs(WF2.rdm4 - g["cas23_rdm4"])) < 1e-12
    th = WF3.thetas
    WF3.thetas = [t + 0.1 for t in th]      # the setter drops the cached higher RDMs
    assert np.max(np.abs(WF3.rdm3 - g["cas43_rdm3"])) > 1e-4


def test_fused_energy_gradient_call(sq, golden):
    """sq_ups_energy_grad (one C-ABI call for _calc_energy_optimization + _calc_gradient_optimization,
    ups_wavefunction.py:1019-1142) against the reference's energy and analytic theta gradient on H2O tUPS(4,4) and
    fUCCSD(4,4) (sa-free generic excitations), and against the separately composed calls."""
    arrays, meta, _ = golden
    from slowquant_b200.integral_manager import ArrayIntegrals
    from slowquant_b200.operators import hamiltonian_0i_0a
    from slowquant_b200.ups_wavefunction import WaveFunctionUPS

    ints = ArrayIntegrals(arrays["h2o_h_mo"], arrays["h2o_g_mo"], num_elec=10)
    eye = np.eye(arrays["h2o_h_mo"].shape[0])
    for name, ansatz, options in (("tups44", "tUPS", {"n_layers": 2}), ("fuccsd44", "fUCCSD", {})):
        WF = WaveFunctionUPS((4, 4), eye, ints, ansatz, dict(options), include_active_kappa=(name == "tups44"))
        th = arrays[f"{name}_thetas"].tolist()
        H = hamiltonian_0i_0a(WF.h_mo, WF.g_mo, WF.num_inactive_orbs, WF.num_active_orbs)
        E, g = sq.osa.ups_energy_and_gradient(WF.csf_coeffs, WF.ci_info, th, WF.ups_layout, H)
        assert abs(E - float(arrays[f"{name}_energy"])) < 1e-10
        ref_grad = WF._calc_gradient_optimization(th, True, False)
        assert np.max(np.abs(g - ref_grad)) < 1e-11
        ref = arrays[f"{name}_gradient"]              # reference: [kappa..., theta...] when active kappa is included
        assert np.max(np.abs(g - ref[len(ref) - len(g) :])) < 1e-10
        E2, g2 = sq.osa.ups_energy_and_gradient(torch.from_numpy(WF.csf_coeffs).cuda(), WF.ci_info, th, WF.ups_layout, H, want_gradient=False)
        assert g2 is None and abs(E2 - E) < 1e-13


def test_backwards_gradient_sweep_on_a_mixed_circuit(sq):
    """sq_ups_energy_grad runs the gradient loop BACKWARDS from (H|psi>, |psi>) (no adjoint pass).  On a circuit that mixes every
    operator family -- tUPS bricks (quad and single gradient launches), generic singles / doubles / a triple, the five sa_double
    cases (polynomial route) -- with a zero angle in the middle, energy and gradient must equal the step-by-step route
    (state, sigma, adjoint, forward sweep of ups_wavefunction.py:1114-1138) and the oracle's literal gradient loop."""
    from slowquant_b200.operators import hamiltonian_0i_0a

    n, na, nb = 7, 3, 3
    sp = orc.get_indexing(0, n, 0, na, nb)
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    rng = np.random.default_rng(314)
    types, idx = orc.tiled_layout(n, 1)
    types = list(types) + ["single", "double", "sa_double_1", "sa_double_2", "sa_single", "sa_double_3", "triple", "sa_double_4",
                           "sa_double_5", "double"]
    # sa_double (i, j, a, b): case 1 i = j and a = b, case 2 i = j, case 3 a = b, cases 4 / 5 all distinct (util.py:26-42)
    idx = list(idx) + [(3, 13), (0, 1, 8, 13), (0, 0, 4, 4), (1, 1, 3, 5), (1, 5), (0, 2, 5, 5), (0, 1, 2, 8, 11, 12), (0, 1, 4, 6),
                       (1, 2, 3, 5), (2, 4, 6, 12)]
    more_t, more_i = orc.tiled_layout(n, 1)
    types += list(more_t)
    idx += list(more_i)
    th = rng.uniform(-1.2, 1.2, len(types))
    th[len(types) // 2] = 0.0
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    lay = _layout(sq, types, idx)
    H = hamiltonian_0i_0a(h, g, 0, n)
    ref_state = np.zeros(sp.num_det)
    ref_state[0] = 1.0
    E, grad = sq.osa.ups_energy_and_gradient(ref_state, info, th.tolist(), lay, H)
    # step by step through the public functions (forward sweep)
    psi = sq.osa.construct_ups_state(ref_state, info, th.tolist(), lay)
    hpsi = sq.osa.propagate_state([H], psi, info)
    assert abs(E - float(psi @ hpsi)) < 1e-12
    bra = sq.osa.construct_ups_state(hpsi, info, th.tolist(), lay, dagger=True)
    g_fwd, _, ket_end = sq.osa.ups_gradient_sweep(bra, ref_state, info, th.tolist(), lay)
    assert np.max(np.abs(ket_end - psi)) < 1e-12
    assert np.max(np.abs(grad - g_fwd)) < 1e-11
    # the oracle's literal loop
    g_orc = orc.theta_gradient(ref_state, th, types, idx, h, g, sp)
    assert np.max(np.abs(grad - g_orc)) < 1e-10
    assert abs(E - orc.energy_elec(orc.construct_ups_state(ref_state, sp, th, types, idx), h, g, sp)) < 1e-10


@pytest.mark.parametrize("n,ne,L,panel", [(8, 4, 2, 512), (6, 3, 2, 512), (7, 3, 1, 512), (8, 4, 2, 2048), (8, 4, 2, 1536), (7, 3, 2, 1024),
                                          (10, 5, 2, 0), (9, 4, 2, 3072)])
def test_sigma_of_spin_flip_symmetric_vectors(sq, n, ne, L, panel):
    """sq_sigma builds H|psi> of a spin-flip symmetric vector (c[B,A] = lambda phi(A,B) c[A,B]: every tUPS state on a closed-shell
    reference; lambda = +1 for an even, -1 for an odd number of electron pairs) from the determinants above the diagonal only.
    The half build must equal the full build (switch off) and the oracle's sigma; a vector without the symmetry -- a random one, or
    a symmetric one disturbed at 1e-8 -- must take the full build (same launches as with the switch off, plus the check).
    Panels of fewer than 1024 determinants take the determinant-per-thread kernels of the half build, wider ones the blocked
    kernels (32 x 32 blocks of determinants; 1536 and 3072: panels whose width is no multiple of a block, 0: the default width)."""
    from slowquant_b200.operators import hamiltonian_0i_0a

    lib = sq.lib.load()
    rng = np.random.default_rng(40 + n)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    sp = orc.get_indexing(0, n, 0, ne, ne)
    types, idx, th, _ = _seeded_case(n, ne, ne, L, 4000 + n)
    hf = np.zeros(sp.num_det)
    hf[0] = 1.0
    psi = orc.construct_ups_state(hf, sp, th, types, idx)
    H_orc = orc.hamiltonian_0i_0a(h, g, 0, n)
    try:
        lib.sq_set_option(b"panel", str(panel).encode())          # several panels, the last one partial, in both builds
        info = sq.ci.get_indexing(0, n, 0, ne, ne)
        H = hamiltonian_0i_0a(h, g, 0, n)
        for label, vec, symmetric in (("tUPS state", psi, True), ("random", rng.normal(size=sp.num_det), False),
                                      ("disturbed", psi + 1e-8 * rng.normal(size=sp.num_det), False)):
            ref = orc.propagate_state([H_orc], vec, sp)
            lib.sq_set_option(b"sigma_spinsym", b"0")
            l0 = lib.sq_launch_count()
            full = sq.osa.propagate_state([H], vec, info)
            launches_full = lib.sq_launch_count() - l0
            lib.sq_set_option(b"sigma_spinsym", b"1")
            l0 = lib.sq_launch_count()
            half = sq.osa.propagate_state([H], vec, info)
            launches_half = lib.sq_launch_count() - l0
            assert np.max(np.abs(full - ref)) < 1e-11, label
            assert np.max(np.abs(half - ref)) < 1e-11, label
            assert np.max(np.abs(half - full)) < 1e-12, label
            # fall-back = the full build plus the symmetry check; the half build adds the mirror pass and has about half the panels
            if symmetric:
                assert launches_half != launches_full + 1, (label, launches_half, launches_full)
                if panel and sp.num_det > 8 * panel:
                    assert launches_half < launches_full, (label, launches_half, launches_full)
            else:
                assert launches_half == launches_full + 1, (label, launches_half, launches_full)
        # the 1-/2-RDM of the symmetric state: Gram matrices over the determinants above the diagonal (weight 2) against all of them
        lib.sq_set_option(b"sigma_spinsym", b"0")
        d1_full, d2_full = sq.osa.reduced_density_matrices(psi, psi, info)
        lib.sq_set_option(b"sigma_spinsym", b"1")
        d1_half, d2_half = sq.osa.reduced_density_matrices(psi, psi, info)
        assert np.max(np.abs(d1_half - d1_full)) < 1e-12 and np.max(np.abs(d2_half - d2_full)) < 1e-12
        assert abs(np.trace(d1_half) - 2 * ne) < 1e-12
        e_rdm = float(np.sum(h * d1_half) + 0.5 * np.sum(g * d2_half))
        assert abs(e_rdm - float(psi @ orc.propagate_state([H_orc], psi, sp))) < 1e-10
        if n <= 6:
            r1 = orc.rdm1(psi, sp)
            assert np.max(np.abs(d1_half - r1)) < 1e-11
            assert np.max(np.abs(d2_half - orc.rdm2(psi, sp, r1))) < 1e-11
    finally:
        lib.sq_set_option(b"panel", b"0")
        lib.sq_set_option(b"sigma_spinsym", b"1")


def test_sigma_half_builds_at_cas14(sq):
    """At a size with many panels of whole CTA waves (CAS(14,14): 11 778 624 determinants, 5 886 blocks of 32 x 32 determinants
    above the diagonal, default panel width): the blocked half build, the determinant-per-thread half build and the full build
    of H|psi> for a tUPS state agree to 1e-13, and so do the energies <psi|H|psi> from sigma and from the 1-/2-RDMs of the
    blocked route (trace of the 1-RDM = number of electrons)."""
    from slowquant_b200.operators import hamiltonian_0i_0a

    lib = sq.lib.load()
    n, ne = 14, 7
    rng = np.random.default_rng(1414)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    info = sq.ci.get_indexing(0, n, 0, ne, ne)
    dev = torch.device("cuda", info.device)
    types, idx, th, _ = _seeded_case(n, ne, ne, 3, 1407)
    lay = _layout(sq, types, idx)
    hf = torch.zeros(info.num_det, dtype=torch.float64, device=dev)
    hf[0] = 1.0
    psi = sq.osa.construct_ups_state(hf, info, th.tolist(), lay)
    H = hamiltonian_0i_0a(h, g, 0, n)
    out = {}
    try:
        for mode in (b"0", b"tri", b"1"):
            lib.sq_set_option(b"sigma_spinsym", mode)
            l0 = lib.sq_launch_count()
            out[mode] = sq.osa.propagate_state([H], psi, info)
            out[mode, "launches"] = lib.sq_launch_count() - l0
        scale = float(torch.max(torch.abs(out[b"0"])))
        assert float(torch.max(torch.abs(out[b"1"] - out[b"0"]))) < 1e-13 * max(1.0, scale)
        assert float(torch.max(torch.abs(out[b"tri"] - out[b"0"]))) < 1e-13 * max(1.0, scale)
        assert out[b"1", "launches"] < out[b"0", "launches"] and out[b"tri", "launches"] < out[b"0", "launches"]
        e_sigma = float(torch.dot(psi, out[b"1"]))
        d1, d2 = sq.osa.reduced_density_matrices(psi, psi, info)
        assert abs(np.trace(d1) - 2 * ne) < 1e-11
        assert abs(float(np.sum(h * d1) + 0.5 * np.sum(g * d2)) - e_sigma) < 1e-10
    finally:
        lib.sq_set_option(b"sigma_spinsym", b"1")


@pytest.mark.parametrize("n,ne,L,qnp", [(8, 4, 3, False), (9, 4, 3, True), (10, 5, 4, False), (12, 6, 5, False), (16, 8, 16, False)])
def test_light_cone_state_construction(sq, n, ne, L, qnp):
    """construct_ups_state_from_determinant: the head of the circuit in the orbital window the reference determinant's light cone
    reaches (a smaller CAS space, same kernels), embedded, then the tail in the full space -- against the plain route (all
    operators on the full vector) to 1e-13, against the oracle at the small sizes; (16, 8, 16) is the circuit of the bench."""
    types, idx, th, _ = _seeded_case(n, ne, ne, L, 7700 + n, qnp=qnp)
    lay = _layout(sq, types, idx)
    info = sq.ci.get_indexing(0, n, 0, ne, ne)
    a = int(info.strings(0)[0])
    plan = sq.osa._light_cone_plan(info, lay, a, a)
    assert plan is not None and plan["window"][1] - plan["window"][0] < n and plan["k0"] > len(plan["head"])
    plain = sq.osa.construct_ups_state_from_determinant(0, info, th.tolist(), lay, light_cone=False)
    l0 = sq.lib.load().sq_launch_count()
    cone = sq.osa.construct_ups_state_from_determinant(0, info, th.tolist(), lay, light_cone=True)
    assert sq.lib.load().sq_launch_count() > l0
    assert float(torch.max(torch.abs(cone - plain))) < 1e-13
    assert abs(float(torch.linalg.norm(cone)) - 1.0) < 1e-12
    if n <= 10:
        sp = orc.get_indexing(0, n, 0, ne, ne)
        hf = np.zeros(sp.num_det)
        hf[0] = 1.0
        assert np.max(np.abs(cone.cpu().numpy() - orc.construct_ups_state(hf, sp, th, types, idx))) < 1e-12


@pytest.mark.parametrize("n,ne,L,qnp", [(6, 3, 2, False), (8, 4, 2, True), (9, 4, 2, False)])
def test_backward_gradient_sweep_entry(sq, n, ne, L, qnp):
    """operator_state_algebra.ups_gradient_sweep_backward (what WaveFunctionUPS._calc_gradient_optimization calls): from
    (H|psi>, |psi>) backwards through the circuit -- against the reference's forward loop from (U^dagger H|psi>, |ref>)
    (ups_wavefunction.py:1114-1138) and the oracle's gradient; the vectors come back as (U^dagger H|psi>, |ref>), the inputs are
    left alone."""
    from slowquant_b200.operators import hamiltonian_0i_0a

    rng = np.random.default_rng(900 + n)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    types, idx, th, _ = _seeded_case(n, ne, ne, L, 9100 + n, qnp=qnp)
    lay = _layout(sq, types, idx)
    info = sq.ci.get_indexing(0, n, 0, ne, ne)
    dev = torch.device("cuda", info.device)
    ref = torch.zeros(info.num_det, dtype=torch.float64, device=dev)
    ref[0] = 1.0
    psi = sq.osa.construct_ups_state(ref, info, th.tolist(), lay)
    sigma = sq.osa.propagate_state([hamiltonian_0i_0a(h, g, 0, n)], psi, info)
    psi0, sigma0 = psi.clone(), sigma.clone()
    g_b, bra_b, ket_b = sq.osa.ups_gradient_sweep_backward(sigma, psi, info, th.tolist(), lay)
    assert torch.equal(psi, psi0) and torch.equal(sigma, sigma0)
    bra_f = sq.osa.construct_ups_state(sigma, info, th.tolist(), lay, dagger=True)
    g_f, _, _ = sq.osa.ups_gradient_sweep(bra_f, ref, info, th.tolist(), lay)
    assert np.max(np.abs(g_b - g_f)) < 1e-12
    assert float(torch.max(torch.abs(ket_b - ref))) < 1e-12 and float(torch.max(torch.abs(bra_b - bra_f))) < 1e-12
    sp = orc.get_indexing(0, n, 0, ne, ne)
    assert np.max(np.abs(g_b - orc.theta_gradient(ref.cpu().numpy(), th, types, idx, h, g, sp))) < 1e-10


def test_per_string_kernels_against_reference_outputs(sq):
    """The reference's per-string entry points (osa.py:33-410: apply_operator_serial / _threaded, their _SA twins,
    add_operator_matrix) through the gather kernel, on CAS(4,5) with 3 alpha / 1 beta electrons, against outputs of the
    reference's numba kernels (tests/golden/make_golden_strings.py): host arrays accumulate in place like the reference,
    device tensors stay resident."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_strings.npz"))
    nI, nA, nV, na, nb = (int(x) for x in g["space"])
    info = sq.ci.get_indexing(nI, nA, nV, na, nb)
    assert np.array_equal(info.idx2det, g["idx2det"])
    osa, pc, st, sts = sq.osa, g["parity_check"], g["state"], g["states"]
    dev_state = torch.from_numpy(st).cuda()
    for c in range(int(g["n_cases"])):
        pre = f"c{c}_"
        f = float(g[pre + "factor"])
        ser = (g[pre + "a_serial"], g[pre + "create_screen"], g[pre + "anni_idx"], nA, pc, info.idx2det, info.det2idx, False)
        thr = (g[pre + "a_threaded"], g[pre + "create_idx"], g[pre + "anni_screen"], nA, pc, info.idx2det, info.det2idx, False)
        tmp = g["tmp0"].copy()
        out = osa.apply_operator_serial(st, *ser, tmp, f)
        assert out is tmp and np.max(np.abs(out - g[pre + "serial"])) < 1e-14
        assert np.max(np.abs(osa.apply_operator_threaded(st, *thr, g["tmp0"].copy(), f) - g[pre + "threaded"])) < 1e-14
        assert np.max(np.abs(osa.apply_operator_SA_serial(sts, *ser, g["tmps0"].copy(), f) - g[pre + "sa_serial"])) < 1e-14
        assert np.max(np.abs(osa.apply_operator_SA_threaded(sts, *thr, g["tmps0"].copy(), f) - g[pre + "sa_threaded"])) < 1e-14
        n = len(st)
        assert np.max(np.abs(osa.add_operator_matrix(np.zeros((n, n)), *ser, f) - g[pre + "matrix"])) < 1e-14
        dev_tmp = torch.from_numpy(g["tmp0"]).cuda()
        dev_out = osa.apply_operator_serial(dev_state, *ser, dev_tmp, f)
        assert dev_out is dev_tmp and dev_out.is_cuda
        assert float(torch.max(torch.abs(dev_out.cpu() - torch.from_numpy(g[pre + "serial"])))) < 1e-14
    assert [osa.bitcount(int(x)) for x in g["bitcount_in"]] == [int(x) for x in g["bitcount_out"]]


def test_rdm_triangle_variant_agrees(sq):
    """sq_set_option("rdm_tri", "1"): the symmetric Gram matrix of sq_rdm12 (bra == ket) from three half-size DGEMMs + a host mirror
    must give the same RDMs as the single DGEMM; transition RDMs (bra != ket) are not affected by the switch."""
    lib = sq.lib.load()
    n, na, nb = 7, 4, 3                                         # n^2 = 49: odd, the two half blocks differ in size
    rng = np.random.default_rng(17)
    info = sq.ci.get_indexing(0, n, 0, na, nb)
    state = rng.normal(size=info.num_det)
    state /= np.linalg.norm(state)
    other = rng.normal(size=info.num_det)
    try:
        lib.sq_set_option(b"panel", b"512")
        info = sq.ci.get_indexing(0, n, 0, na, nb)              # a fresh space picks up the panel width (several panels)
        lib.sq_set_option(b"rdm_tri", b"0")
        d1, d2 = sq.osa.reduced_density_matrices(state, state, info)
        t1, t2 = sq.osa.reduced_density_matrices(other, state, info)
        lib.sq_set_option(b"rdm_tri", b"1")
        e1, e2 = sq.osa.reduced_density_matrices(state, state, info)
        u1, u2 = sq.osa.reduced_density_matrices(other, state, info)
    finally:
        lib.sq_set_option(b"rdm_tri", b"0")
        lib.sq_set_option(b"panel", b"0")
    assert np.max(np.abs(d1 - e1)) < 1e-12 and np.max(np.abs(d2 - e2)) < 1e-12
    assert np.max(np.abs(t1 - u1)) < 1e-12 and np.max(np.abs(t2 - u2)) < 1e-12
    assert abs(np.trace(e1) - (na + nb)) < 1e-12


def test_table_free_panel_kernels_agree(sq):
    """sq_set_option("etab", "alu"): sigma (symmetric and unsymmetric integrals) against the oracle and RDMs / transition RDMs
    against the table kernels, with small panels (several panels, a partial last one)."""
    from slowquant_b200.operators import hamiltonian_0i_0a

    lib = sq.lib.load()
    n, na, nb = 8, 4, 3
    rng = np.random.default_rng(7)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    g_unsym = g + 0.05 * rng.normal(size=(n, n, n, n))
    sp = orc.get_indexing(0, n, 0, na, nb)
    state = rng.normal(size=sp.num_det)
    state /= np.linalg.norm(state)
    other = rng.normal(size=sp.num_det)
    results = []
    try:
        for etab in (b"smem", b"alu"):
            lib.sq_set_option(b"panel", b"768")
            lib.sq_set_option(b"etab", etab)
            info = sq.ci.get_indexing(0, n, 0, na, nb)
            for gg in (g, g_unsym):
                ref = orc.propagate_state([orc.hamiltonian_0i_0a(h, gg, 0, n)], state, sp)
                out = sq.osa.propagate_state([hamiltonian_0i_0a(h, gg, 0, n)], state, info)
                assert np.max(np.abs(out - ref)) < 1e-11, etab
            results.append(sq.osa.reduced_density_matrices(state, state, info) + sq.osa.reduced_density_matrices(other, state, info))
    finally:
        lib.sq_set_option(b"panel", b"0")
        lib.sq_set_option(b"etab", b"smem")
    for x, y in zip(*results):
        assert np.max(np.abs(x - y)) < 1e-12


def test_cas16_bench_config(sq):
    """The benchmarked configuration itself (BASELINE.json config 4: CAS(16,16), 165 636 900 determinants, dense random vector,
    tUPS): (i) the default window plan (H = 6 windows, time-skewed schedule, 20 x 20 tiles, top-window path) against the
    one-brick-per-launch path, (ii) the adjoint round trip, (iii) one brick (3 operators, 5 rotations) against the oracle's
    restatement of the reference loop (osa.py:1002-1085) -- the same call bench.py times as the CPU baseline, (iv) the L = 16
    circuit of the bench from the HF determinant, window plan against bricks."""
    lib = sq.lib.load()
    n, ne = 16, 8
    info = sq.ci.get_indexing(0, n, 0, ne, ne)
    dev = torch.device("cuda", info.device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1604)
    x = torch.randn(info.num_det, dtype=torch.float64, device=dev, generator=gen)
    x /= torch.linalg.norm(x)
    types, idx, th, rng = _seeded_case(n, ne, ne, 2, 1616)
    lay = _layout(sq, types, idx)
    try:
        res = sq.osa.construct_ups_state(x, info, th.tolist(), lay)
        back = sq.osa.construct_ups_state(res, info, th.tolist(), lay, dagger=True)
        assert float(torch.max(torch.abs(back - x))) < 1e-12                     # (ii)
        del back
        sq.lib.check(lib.sq_set_option(b"win", b"0"))
        ref = sq.osa.construct_ups_state(x, info, th.tolist(), lay)
        assert float(torch.max(torch.abs(res - ref))) < 1e-13                    # (i)
        del res, ref
        # (iv) the bench circuit: L = 16 from the HF determinant (sparse light cone at first, dense at the end)
        types16, idx16, th16, _ = _seeded_case(n, ne, ne, 16, 1234)
        lay16 = _layout(sq, types16, idx16)
        hf = torch.zeros(info.num_det, dtype=torch.float64, device=dev)
        hf[0] = 1.0
        ref16 = sq.osa.construct_ups_state(hf, info, th16.tolist(), lay16)
        sq.lib.check(lib.sq_set_option(b"win", b"1"))
        res16 = sq.osa.construct_ups_state(hf, info, th16.tolist(), lay16)
        assert float(torch.max(torch.abs(res16 - ref16))) < 1e-13
        assert abs(float(torch.linalg.norm(res16)) - 1.0) < 1e-12
        del res16, ref16, hf
        # (iii) one brick against the oracle (16 host threads: about 15 s)
        one = _layout(sq, types[0:3], idx[0:3])
        got = sq.osa.construct_ups_state(x, info, [0.7, -0.4, 0.3], one).cpu().numpy()
        sp = orc.get_indexing(0, n, 0, ne, ne)
        want = orc.construct_ups_state(x.cpu().numpy(), sp, [0.7, -0.4, 0.3], types[0:3], idx[0:3], threaded=True)
        assert np.max(np.abs(got - want)) < 1e-12
    finally:
        sq.lib.check(lib.sq_set_option(b"win", b"1"))


@pytest.mark.parametrize("nb_cols,n_rows", [(12870, 37), (924, 200), (35, 50), (40000, 5)])
def test_reshard_rows_kernel_on_one_device(sq, nb_cols, n_rows):
    """sq_reshard_rows (the all-to-all step of the sharded engine) as a row permutation on ONE device: bulk-copy engine
    (cp.async.bulk through shared memory, rows longer and shorter than a chunk, partial last stage) and the vector load/store
    kernel (odd row length: no 16-byte alignment), bit-exact against torch indexing."""
    import ctypes as C

    lib = sq.lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device=dev)
    gen.manual_seed(nb_cols)
    src = torch.randn(n_rows, nb_cols, dtype=torch.float64, device=dev, generator=gen)
    perm = torch.randperm(n_rows, device=dev, generator=gen).to(torch.int32)
    rank = torch.zeros(n_rows, dtype=torch.int32, device=dev)
    try:
        for mode in (b"tma", b"lsu"):
            sq.lib.check(lib.sq_set_option(b"reshard", mode))
            dst = torch.full((n_rows, nb_cols), float("nan"), dtype=torch.float64, device=dev)
            ptrs = (C.c_void_p * 1)(dst.data_ptr())
            sq.lib.check(lib.sq_reshard_rows(dev.index, n_rows, nb_cols, C.c_void_p(src.data_ptr()), C.c_void_p(rank.data_ptr()),
                                             C.c_void_p(perm.data_ptr()), ptrs, 1, None))
            torch.cuda.synchronize()
            want = torch.empty_like(src)
            want[perm.long()] = src
            assert torch.equal(dst, want), mode
    finally:
        sq.lib.check(lib.sq_set_option(b"reshard", b"tma"))
