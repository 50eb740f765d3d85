"""Subprocess body of tests/test_host_callers.py: host logic of the engine's callers on the oracle-backed stand-in
(tests/host_standin.py), against the reference goldens.  Prints one line per check and HOST_CALLERS_OK at the end."""
import contextlib
import importlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import host_standin  # noqa: E402  (patches slowquant_b200 in this process)

import slowquant_b200.linear_response._panels as pn  # noqa: E402
import slowquant_b200.operator_state_algebra as osa  # noqa: E402
import slowquant_b200.sa_ups_wavefunction as sam  # noqa: E402
from slowquant_b200 import operators as mops  # noqa: E402
from slowquant_b200.ci_spaces import get_indexing_extended  # noqa: E402
from slowquant_b200.integral_manager import ArrayIntegrals  # noqa: E402
from slowquant_b200.ups_wavefunction import WaveFunctionUPS  # noqa: E402
from slowquant_b200.util import UpsStructure  # noqa: E402

pn.torch = sam.torch = host_standin.TorchProxy()
G = os.path.join(HERE, "golden")
worst = {}


def record(name, value, tol):
    value = float(value)
    print(f"{name:48s} {value:9.2e}  (tol {tol:.0e})", flush=True)
    worst[name] = (value, tol)


def d(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


which = sys.argv[1] if len(sys.argv) > 1 else "all"

# ---- linear response: all eight parametrisations on LiH tUPS(2,2) (and naive on H2O tUPS(4,4) with "h2o") ----
if which in ("all", "lr", "h2o"):
    g1 = np.load(os.path.join(G, "golden_config1.npz"))
    gv = np.load(os.path.join(G, "golden_lr_variants.npz"))
    systems = [("lih", {"n_layers": 1, "skip_last_singles": True})] + ([("h2o", {"n_layers": 3})] if which == "h2o" else [])
    variants = [("naive", "naive"), ("proj", "projected"), ("st", "statetransfer"), ("sc", "selfconsistent"), ("allst", "allstatetransfer"),
                ("allsc", "allselfconsistent"), ("allproj", "allprojected"), ("projst", "projected_statetransfer")]
    for name, options in systems:
        pre = name + "_"
        ints = ArrayIntegrals(g1[pre + "h_ao"], g1[pre + "eri_ao"], int(g1[pre + "num_elec"]), dipole=tuple(g1[pre + "dipole_ao"]))
        WF = WaveFunctionUPS(tuple(int(x) for x in g1[pre + "cas"]), g1[pre + "c_mo"], ints, "tUPS", dict(options), include_active_kappa=True)
        WF.thetas = g1[pre + "thetas"].tolist()
        record(f"{name} energy", abs(WF.energy_elec - float(g1[pre + "energy"])), 1e-10)
        for tag, modname in variants:
            keys = ("A", "B", "Sigma", "Delta", "excitation_energies", "norms", "oscillator_strengths")
            gold = {k: (g1[pre + k] if tag == "naive" else gv[f"{name}_{tag}_{k}"]) for k in keys}
            mod = importlib.import_module("slowquant_b200.linear_response." + modname)
            mod.torch = host_standin.TorchProxy()
            with contextlib.redirect_stdout(io.StringIO()):
                LR = mod.LinearResponse(WF, "SD")
                LR.calc_excitation_energies()
                osc = LR.get_oscillator_strength()
            record(f"{name} {modname} A/B/Sigma/Delta", max(d(getattr(LR, k), gold[k]) for k in ("A", "B", "Sigma", "Delta")), 1e-10)
            record(f"{name} {modname} excitation energies", d(LR.excitation_energies, gold["excitation_energies"]), 1e-8)
            record(f"{name} {modname} norms / osc. strengths", max(d(LR.get_excited_state_norm(), gold["norms"]), d(osc, gold["oscillator_strengths"])), 1e-8)

# ---- state-averaged UPS at the reference's parameters ----
if which in ("all", "saups"):
    g = np.load(os.path.join(G, "golden_saups.npz"))
    s2 = 2 ** (-1 / 2)
    cases = {
        "h2": (([[1], [s2, -s2], [1]], [["1100"], ["1001", "0110"], ["0011"]]), {"n_layers": 1, "skip_last_singles": True}),
        "h3": (([[1], [s2, -s2], [s2, -s2]], [["110000"], ["100100", "011000"], ["100001", "010010"]]), {"n_layers": 2, "skip_last_singles": True}),
    }
    for name, (states, options) in cases.items():
        pre = name + "_"
        ints = ArrayIntegrals(g[pre + "h_ao"], g[pre + "eri_ao"], int(g[pre + "num_elec"]), dipole=tuple(g[pre + "dipole_ao"]))
        WF = sam.WaveFunctionSAUPS(tuple(int(x) for x in g[pre + "cas"]), g[pre + "c_mo"], ints, states, "tUPS", dict(options), include_active_kappa=True)
        WF.thetas = g[pre + "thetas"].tolist()
        record(f"saups {name} ci / rdm1 / rdm2", max(d(WF.ci_coeffs, g[pre + "ci"]), d(WF.rdm1, g[pre + "rdm1"]), d(WF.rdm2, g[pre + "rdm2"])), 1e-10)
        record(f"saups {name} energies", max(abs(WF.sa_energy - float(g[pre + "sa_energy"])), d(WF.energy_states, g[pre + "energy_states"])), 1e-10)
        record(f"saups {name} oscillator strengths", d(WF.get_oscillator_strenghts(), g[pre + "oscillator_strengths"]), 1e-9)
        th = g[pre + "pert_thetas"].tolist()
        params = [0.0] * len(WF.kappa_idx) + th
        WF._old_opt_parameters = np.zeros(len(params)) + 10**20
        e = WF._calc_energy_optimization(params, True, True)
        grad = WF._calc_gradient_optimization(params, True, True)
        record(f"saups {name} energy / gradient at perturbed theta", max(abs(e - float(g[pre + "pert_energy"])), d(grad, g[pre + "pert_gradient"])), 1e-10)
        roto = WF._calc_energy_rotosolve_optimization(th, g[pre + "rs_shifts"].tolist(), int(g[pre + "rs_idx"]))
        record(f"saups {name} rotosolve energies", d(roto, g[pre + "rs_energies"]), 1e-10)

# ---- extended spaces: embedding, projection (do_unsafe), operator lists ----
if which in ("all", "extended"):
    g = np.load(os.path.join(G, "golden_extended.npz"))
    for tag in ("a", "b"):
        pre = tag + "_"
        sp = tuple(int(x) for x in g[pre + "space"])
        nI, nA, nV = sp[:3]
        ci = get_indexing_extended(*sp)
        lay = UpsStructure()
        lay.create_tiled(nA, {"n_layers": int(g[pre + "n_layers"]), "do_tups": True})
        th, state = g[pre + "thetas"].tolist(), g[pre + "state"]
        q = mops.G1_sa(0, nI + nA + nV - 1)
        E = mops.Epq(nI, nI + 1) * mops.Epq(nI + 1, nI) + 0.5 * mops.Epq(nI, nI)
        errs = [
            d(osa.construct_ups_state(state, ci, th, lay), g[pre + "U_state"]),
            d(osa.construct_ups_state(state, ci, th, lay, dagger=True), g[pre + "Ud_state"]),
            d(osa.propagate_unitary(state, 3, ci, th, lay), g[pre + "unitary3"]),
            d(osa.get_grad_action(state, 1, ci, lay), g[pre + "grad1"]),
            d(osa.propagate_state([q.dagger, q], state, ci, do_unsafe=True), g[pre + "qd_q_state"]),
            d(osa.propagate_state(["U", q], state, ci, th, lay, do_unsafe=True), g[pre + "U_q_state"]),
            abs(osa.expectation_value(state, ["Ud", E, "U"], state, ci, th, lay) - float(g[pre + "expval"])),
        ]
        record(f"extended space {tag}", max(errs), 1e-12)
        raised = False
        try:
            osa.propagate_state([q], state, ci)
        except KeyError:
            raised = True
        record(f"extended space {tag} KeyError as the reference", float(raised != bool(int(g[pre + "q_raises"]))), 0.5)

# ---- the reference's per-string kernels (osa.py:33-410) and generate_spin_strings, against reference outputs ----
if which in ("all", "strings"):
    from slowquant_b200.ci_spaces import generate_spin_strings, get_indexing

    g = np.load(os.path.join(G, "golden_strings.npz"))
    nI, nA, nV, na, nb = (int(x) for x in g["space"])
    ci = get_indexing(nI, nA, nV, na, nb)
    record("strings: idx2det bit-exact", float(not np.array_equal(ci.idx2det, g["idx2det"])), 0.5)
    record("strings: bitcount", float(any(osa.bitcount(int(x)) != int(y) for x, y in zip(g["bitcount_in"], g["bitcount_out"]))), 0.5)
    bad_strings = 0
    for n, k in ((5, 3), (5, 1), (4, 0), (3, 3)):
        bad_strings += not np.array_equal(np.array(list(generate_spin_strings(n, k)), dtype=np.int64).reshape(-1, n), g[f"strings_{n}_{k}"])
    bad_strings += list(generate_spin_strings(3, -1)) != []
    masks = [sum(b << o for o, b in enumerate(s_)) for s_ in generate_spin_strings(nA, na)]
    bad_strings += not np.array_equal(np.array(masks, dtype=np.uint32), ci.strings(0))
    record("strings: generate_spin_strings == reference == engine tables", float(bad_strings), 0.5)
    pc, st, sts = g["parity_check"], g["state"], g["states"]
    errs = {"serial": 0.0, "threaded": 0.0, "sa_serial": 0.0, "sa_threaded": 0.0, "matrix": 0.0}
    for c in range(int(g["n_cases"])):
        pre = f"c{c}_"
        f = float(g[pre + "factor"])
        t = g["tmp0"].copy()
        r = osa.apply_operator_serial(st, g[pre + "a_serial"], g[pre + "create_screen"], g[pre + "anni_idx"], nA, pc, ci.idx2det, ci.det2idx, False, t, f)
        assert r is t  # accumulates in place and returns tmp_state, as the reference
        errs["serial"] = max(errs["serial"], d(r, g[pre + "serial"]))
        r = osa.apply_operator_threaded(st, g[pre + "a_threaded"], g[pre + "create_idx"], g[pre + "anni_screen"], nA, pc, ci.idx2det, ci.det2idx, False, g["tmp0"].copy(), f)
        errs["threaded"] = max(errs["threaded"], d(r, g[pre + "threaded"]))
        r = osa.apply_operator_SA_serial(sts, g[pre + "a_serial"], g[pre + "create_screen"], g[pre + "anni_idx"], nA, pc, ci.idx2det, ci.det2idx, False, g["tmps0"].copy(), f)
        errs["sa_serial"] = max(errs["sa_serial"], d(r, g[pre + "sa_serial"]))
        r = osa.apply_operator_SA_threaded(sts, g[pre + "a_threaded"], g[pre + "create_idx"], g[pre + "anni_screen"], nA, pc, ci.idx2det, ci.det2idx, False, g["tmps0"].copy(), f)
        errs["sa_threaded"] = max(errs["sa_threaded"], d(r, g[pre + "sa_threaded"]))
        r = osa.add_operator_matrix(np.zeros((len(st), len(st))), g[pre + "a_serial"], g[pre + "create_screen"], g[pre + "anni_idx"], nA, pc, ci.idx2det, ci.det2idx, False, f)
        errs["matrix"] = max(errs["matrix"], d(r, g[pre + "matrix"]))
    for k, v in errs.items():
        record(f"strings: apply_operator / add_operator_matrix [{k}]", v, 1e-14)
    raised = 0
    try:
        osa.apply_operator_serial(st, g["c0_a_serial"], g["c0_create_screen"], g["c0_anni_idx"], nA, pc, ci.idx2det, {}, False, g["tmp0"].copy(), 1.0)
    except TypeError:
        raised += 1
    try:
        osa.apply_operator_serial(st, g["c0_a_serial"], g["c0_anni_idx"], g["c0_anni_idx"], nA, pc, ci.idx2det, ci.det2idx, False, g["tmp0"].copy(), 1.0)
    except ValueError:
        raised += 1
    record("strings: foreign det2idx -> TypeError, inconsistent screen -> ValueError", float(raised != 2), 0.5)

# ---- constructor bookkeeping of the wave-function classes (orbital partitions, index lists, kappa lists) ----
if which in ("all", "attributes"):
    import json

    from slowquant_b200.ucc_wavefunction import WaveFunctionUCC

    gj = json.load(open(os.path.join(G, "golden_wf_attributes.json")))
    ints = ArrayIntegrals(np.array(gj["h_ao"]), np.array(gj["eri_ao"]), int(gj["num_elec"]))
    c_mo = np.array(gj["c_mo"])

    def norm(v):
        if isinstance(v, np.ndarray):
            return v.tolist()
        if isinstance(v, (list, tuple)):
            return [norm(x) for x in v]
        if isinstance(v, (bool, np.bool_)):
            return bool(v)
        if isinstance(v, (int, np.integer)):
            return int(v)
        return v

    for case in gj["cases"]:
        with contextlib.redirect_stdout(io.StringIO()):
            if case["kind"] == "ups":
                WF = WaveFunctionUPS(tuple(case["cas"]), c_mo, ints, case["ansatz"], dict(case["options"]), include_active_kappa=case["include_active_kappa"])
            elif case["kind"] == "saups":
                WF = sam.WaveFunctionSAUPS(tuple(case["cas"]), c_mo, ints, (case["states"][0], case["states"][1]), case["ansatz"], dict(case["options"]),
                                           include_active_kappa=case["include_active_kappa"])
            else:
                WF = WaveFunctionUCC(tuple(case["cas"]), c_mo, ints, case["ansatz"], include_active_kappa=case["include_active_kappa"])
        missing = [k for k in case["attributes"] if not hasattr(WF, k)]
        wrong = [k for k, v in case["attributes"].items() if hasattr(WF, k) and norm(getattr(WF, k)) != v]
        tag = f"{case['kind']} {tuple(case['cas'])} {case['ansatz']} kappa_aa={case['include_active_kappa']}"
        if missing or wrong:
            print("   missing:", missing, " wrong:", wrong, flush=True)
        record(f"attributes {tag}: {len(case['attributes'])} reference attributes", float(len(missing) + len(wrong)), 0.5)

# ---- rdm3 / rdm4 contractions ----
if which in ("all", "rdm34"):
    g0 = np.load(os.path.join(G, "golden.npz"))
    g = np.load(os.path.join(G, "golden_rdm34.npz"))
    ints = ArrayIntegrals(g0["h2o_h_mo"], g0["h2o_g_mo"], num_elec=10)
    eye = np.eye(g0["h2o_h_mo"].shape[0])
    WF3 = WaveFunctionUPS((4, 3), eye, ints, "tUPS", {"n_layers": 2}, include_active_kappa=True)
    WF3.thetas = g["cas43_thetas"].tolist()
    record("rdm3 / rdm4 CAS(4,3)", max(d(WF3.rdm3, g["cas43_rdm3"]), d(WF3.rdm4, g["cas43_rdm4"])), 1e-10)

# ---- UCC wave function + linear response (UccStructure path of the "U" / "Ud" panels) ----
if which in ("all", "ucc"):
    from slowquant_b200.ucc_wavefunction import WaveFunctionUCC

    g = np.load(os.path.join(G, "golden_h4_ucc.npz"))
    ints = ArrayIntegrals(g["h_ao"], g["eri_ao"], 4, dipole=tuple(g["dipole_ao"]))
    WF = WaveFunctionUCC((4, 4), g["c_mo_rhf"], ints, "SD")
    WF.thetas = g["thetas"].tolist()
    record("ucc H4 ci / energy", max(d(WF.ci_coeffs, g["ci"]), abs(WF.energy_elec - float(g["energy"]))), 1e-10)
    for tag, modname in (("naive", "naive"), ("sc", "selfconsistent")):
        mod = importlib.import_module("slowquant_b200.linear_response." + modname)
        mod.torch = host_standin.TorchProxy()
        with contextlib.redirect_stdout(io.StringIO()):
            LR = mod.LinearResponse(WF, "SD")
            LR.calc_excitation_energies()
            osc = LR.get_oscillator_strength()
        record(f"ucc H4 {modname} A/B/Sigma/Delta", max(d(getattr(LR, k), g[f"{tag}_{k}"]) for k in ("A", "B", "Sigma", "Delta")), 1e-9)
        record(f"ucc H4 {modname} excitation energies", d(LR.excitation_energies, g[f"{tag}_excitation_energies"]), 1e-8)
        record(f"ucc H4 {modname} oscillator strengths", d(osc, g[f"{tag}_oscillator_strengths"]), 1e-7)

# ---- two-step optimisation drivers (host control flow around the engine) ----
if which in ("all", "opt2"):
    g1 = np.load(os.path.join(G, "golden_config1.npz"))
    ints = ArrayIntegrals(g1["lih_h_ao"], g1["lih_eri_ao"], int(g1["lih_num_elec"]), dipole=tuple(g1["lih_dipole_ao"]))
    WF = WaveFunctionUPS((2, 2), g1["lih_c_mo_rhf"], ints, "tUPS", {"n_layers": 1, "skip_last_singles": True}, include_active_kappa=True)
    with contextlib.redirect_stdout(io.StringIO()):
        WF.run_wf_optimization_2step("BFGS", True, is_silent_subiterations=True)
        WF.check_orthonormality(np.linalg.inv(g1["lih_c_mo_rhf"] @ g1["lih_c_mo_rhf"].T))
    # the reference's ONE-step optimisation of the same wave function (golden_config1) reaches the same minimum
    record("UPS two-step energy vs reference one-step minimum", abs(WF.energy_elec - float(g1["lih_energy"])), 1e-7)
    H = WF._get_hamiltonian()
    record("_get_hamiltonian expectation value", abs(osa.expectation_value(WF.ci_coeffs, [H], WF.ci_coeffs, WF.ci_info, do_folding=False) - WF.energy_elec), 1e-10)
    g = np.load(os.path.join(G, "golden_saups.npz"))
    s2 = 2 ** (-1 / 2)
    states = ([[1], [s2, -s2], [s2, -s2]], [["110000"], ["100100", "011000"], ["100001", "010010"]])
    ints = ArrayIntegrals(g["h3_h_ao"], g["h3_eri_ao"], int(g["h3_num_elec"]), dipole=tuple(g["h3_dipole_ao"]))
    WS = sam.WaveFunctionSAUPS((2, 3), g["h3_c_mo"], ints, states, "tUPS", {"n_layers": 2, "skip_last_singles": True}, include_active_kappa=True)
    with contextlib.redirect_stdout(io.StringIO()):
        WS.run_wf_optimization_2step("BFGS", True, is_silent_subiterations=True)
    # tests/test_unitary_product_state.py:238-242
    record("SA-UPS two-step excitation energies (reference literals)", d(WS.excitation_energies, [0.838466, 0.838466]), 1e-6)
    record("SA-UPS two-step oscillator strengths (reference literals)", d(WS.get_oscillator_strenghts(), [0.7569, 0.7569]), 1e-3)

# ---- one optimiser iteration of WaveFunctionUPS: what the class asks of the engine per parameter set ----
if which in ("all", "iteration"):
    n = 6
    rng = np.random.default_rng(0)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g4 = B + B.transpose(1, 0, 2, 3)
    g4 = g4 + g4.transpose(0, 1, 3, 2)
    g4 = g4 + g4.transpose(2, 3, 0, 1)
    WF = WaveFunctionUPS((n, n), np.eye(n), ArrayIntegrals(h, g4, num_elec=n), "tUPS", {"n_layers": 2})
    calls = {"sigma": 0, "cone": 0, "plain": 0, "backward": 0}

    def count(name, key):
        fn = getattr(osa, name)

        def wrapped(*a, **k):
            calls[key] += 1
            return fn(*a, **k)

        setattr(osa, name, wrapped)

    for name, key in (("propagate_state", "sigma"), ("construct_ups_state_from_determinant", "cone"), ("construct_ups_state", "plain"),
                      ("ups_gradient_sweep_backward", "backward")):
        count(name, key)
    th = list(np.random.default_rng(1).uniform(-1, 1, len(WF.thetas)))
    E = WF._calc_energy_optimization(th, True, False)
    grad = WF._calc_gradient_optimization(th, True, False)
    # fun(x) + jac(x) at the same x: ONE state construction (light cone of the HF determinant), ONE H|psi>, ONE backwards sweep
    record("iteration: engine calls per parameter set", abs(calls["sigma"] - 1) + abs(calls["cone"] - 1) + calls["plain"] + abs(calls["backward"] - 1), 0)
    record("iteration: energy_elec reuses H|psi>", abs(WF.energy_elec - E) + abs(calls["sigma"] - 1), 1e-12)
    fd = []
    for k in (1, 4, 9, len(th) - 1):
        tp, tm = list(th), list(th)
        tp[k] += 1e-5
        tm[k] -= 1e-5
        fd.append((WF._calc_energy_optimization(tp, True, False) - WF._calc_energy_optimization(tm, True, False)) / 2e-5 - grad[k])
    record("iteration: backwards-sweep gradient vs finite differences", max(abs(x) for x in fd), 1e-7)
    # the kept H|psi> follows the state: new parameters and a directly set state both invalidate it
    before = calls["sigma"]
    E2 = WF._calc_energy_optimization(th, True, False)
    record("iteration: new parameters rebuild state and H|psi>", abs(E2 - E) + abs(calls["sigma"] - before - 1), 1e-12)
    WF.light_cone = False
    WF.thetas = th
    plain = np.array(WF.ci_coeffs, copy=True)
    WF.light_cone = True
    WF.thetas = th
    record("iteration: light-cone state vs every operator on the full vector", d(WF.ci_coeffs, plain), 1e-13)
    plans = [v for v in WF.ci_info.__dict__.get("_light_cone", {}).values()]
    record("iteration: the light-cone plan is a proper window", 0.0 if plans and plans[0] is not None and plans[0]["window"][1] - plans[0]["window"][0] < n else 1.0, 0)
    WF.ci_coeffs = plain[::-1].copy()
    e_rev = WF.energy_elec
    WF.ci_coeffs = plain
    record("iteration: ci_coeffs setter invalidates H|psi>", abs(WF.energy_elec - E) + (0.0 if abs(e_rev - E) > 1e-6 else 1.0), 1e-10)

bad = {k: v for k, v in worst.items() if not v[0] <= v[1]}
if bad:
    print("HOST_CALLERS_FAILED", bad, flush=True)
    sys.exit(1)
print("HOST_CALLERS_OK", len(worst), "checks", flush=True)
