#!/usr/bin/env python
"""Benchmark of the hot path: tUPS layer applications per second at CAS(16,16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cas N_ORB] [--layers L]

A *step* applies one L-layer tUPS circuit (L = 16 -> 720 ansatz operators, reference util.py:694-745) to
the resident 165 636 900-amplitude fp64 CI vector.  `value` = layers/s with the vector resident in HBM;
`e2e` = the same through the public `construct_ups_state(numpy_state, ...)` call with pinned HOST buffers
(H2D of the state + D2H of the result inside the timed region).  One JSON line is printed by rank 0.

N > 1 (torchrun): ONE vector sharded by alpha string over the N GPUs (strong scaling, the multi-GPU path BASELINE.json
names): the circuit runs as a few local phases of window sweeps in two row layouts with an all-to-all re-shard over NVLink
peer memory between them; the line carries the re-shard's NVLink bandwidth and a sharded-vs-single-GPU parity check.
--mode replicas runs independent vectors per GPU instead (RotoSolve shifts / finite-difference columns; weak scaling,
no data-path collective).

--impl reference times the CPU restatement of the reference algorithm (oracle/, OpenMP over all host
cores) on a bounded sample of the same workload; /root/reference itself is pure Python + numba and does
not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "tUPS layer applications/s at CAS(16,16)"
UNIT = "layers/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cas", type=int, default=16, help="active orbitals n; CAS(n,n)")
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--mode", default="auto", choices=["auto", "sharded", "replicas"],
                    help="N > 1: independent replicas (weak scaling; auto while the vector fits one GPU) or one "
                    "alpha-sharded vector (strong scaling; auto when 8*N_det > 64 GB)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the energy / RDM / theta-gradient timings")
    ap.add_argument("--profile", action="store_true",
                    help="for nsys / ncu runs: cudaProfilerStart/Stop and an NVTX range around the timed region (the library's own NVTX "
                    "ranges -- sq_ups_apply, sq_sigma, sq_rdm12, sq_ups_grad_sweep, sq_reshard_rows -- nest inside it), no e2e / extras / CPU baseline")
    return ap.parse_args()


def vector_bytes(n: int) -> float:
    from math import comb

    return 8.0 * comb(n, n // 2) ** 2


def workload_name(n: int, L: int) -> str:
    from math import comb

    nd = comb(n, n // 2) ** 2
    return f"synthetic random-parameter tUPS CAS({n},{n}) state construction, {nd} determinants, L={L} layers ({3 * (n - 1) * L} operators) per step"


def measured_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown",
            nv.nvmlClocksThrottleReasonApplicationsClocksSetting: "applications_clocks_setting",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self) -> dict:
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference algorithm on a bounded sample
# ---------------------------------------------------------------------------------------------
_CPU_CACHE: dict = {}


def cpu_sample(n: int) -> dict:
    """Time ONE tUPS brick -- [sa_single, double, sa_single] on orbitals (0,1): 5 exp(theta T) rotations (an sa_single
    is an alpha and a beta rotation in sequence, osa.py:1009-1042), 6 string passes + the numpy closed form each
    (reference osa.py:1002-1085) -- on the dense CAS(n,n) vector with the oracle's OpenMP restatement of
    apply_operator_threaded (osa.py:139-219), and scale to one layer = (n-1) bricks (util.py:694-745)."""
    from oracle import sq_oracle as orc

    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    ne = n // 2
    if n not in _CPU_CACHE:
        sp = orc.get_indexing(0, n, 0, ne, ne)
        rng = np.random.default_rng(1234)
        state = rng.normal(size=sp.num_det)
        state /= np.linalg.norm(state)
        _CPU_CACHE[n] = (sp, state)
    sp, state = _CPU_CACHE[n]
    types, idx = orc.tiled_layout(n, 1)
    assert list(types[:3]) == ["sa_single", "double", "sa_single"]
    t0 = time.perf_counter()
    orc.construct_ups_state(state, sp, [0.7, -0.4, 0.3], types[0:3], idx[0:3], threaded=True)
    dt = time.perf_counter() - t0
    layer_s = dt * BRICKS_PER_LAYER(n)
    return {
        "value": 1.0 / layer_s,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"1 of the {BRICKS_PER_LAYER(n)} bricks of one tUPS layer ([sa_single, double, sa_single] on orbitals 0,1 = 3 "
        f"ansatz operators = 5 rotations, 30 string passes) on the dense CAS({n},{n}) vector ({sp.num_det} determinants): "
        f"{dt:.2f} s, scaled x{BRICKS_PER_LAYER(n)} to one layer",
        "seconds_per_sample": dt,
    }


def BRICKS_PER_LAYER(n: int) -> int:
    return n - 1


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, L = args.cas, args.layers
    for _ in range(min(args.warmup, 1)):  # one untimed pass (page faults, OpenMP pool); more would only burn host minutes
        cpu_sample(n)
    t_total = 0.0
    res = None
    timed = 0
    budget_s = 150.0   # host time for the timed samples: the whole arm has to end within a few minutes whatever K is
    for _ in range(max(args.steps, 1)):
        res = cpu_sample(n)
        t_total += res["seconds_per_sample"]
        timed += 1
        if t_total + res["seconds_per_sample"] > budget_s:
            break
    mean_sample = t_total / timed
    layer_s = mean_sample * BRICKS_PER_LAYER(n)
    value = 1.0 / layer_s
    cpu = dict(res)
    cpu["value"] = value
    cpu.pop("seconds_per_sample", None)
    line = {
        "impl": "reference",
        "metric": METRIC if n == 16 else f"tUPS layer applications/s at CAS({n},{n})",
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        # what was actually timed per step: ONE brick sample (the whole L-layer step is that sample x bricks-per-layer x L, see
        # `extrapolated_ms_per_full_step`; the driver's clock around this run can only be compared with ms_per_step x steps)
        "ms_per_step": 1e3 * mean_sample,
        "step_scale_to_full": BRICKS_PER_LAYER(n) * L,
        "extrapolated_ms_per_full_step": 1e3 * layer_s * L,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(n, L), "step_is": "bounded sample: one tUPS brick (3 operators, 5 rotations), scaled to L layers; at most one untimed warm-up sample"},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "samples_timed": timed,   # == steps unless the 150 s host-time budget ended the loop earlier (the samples are identical)
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def energy_gradient_extras(info, lay, thetas, state, dev) -> dict:
    """The rest of north_star's "energy + gradient" at the same CAS, timed with CUDA events on the resident state:
    sigma build H|psi> (ups_wavefunction.py:770-784), 1-/2-RDM (ups_wavefunction.py:409-476) and the reverse
    theta-gradient sweep (ups_wavefunction.py:1114-1138: per operator g_k = 2<bra|T_k|ket>, then both vectors <- U_k).
    Algorithmic bytes of the sweep: 32 B x touched amplitudes per LAUNCH (read + write of bra and ket; one or two bricks)."""
    import torch

    from slowquant_b200 import _lib
    from slowquant_b200 import operator_state_algebra as osa
    from slowquant_b200.operators import hamiltonian_0i_0a

    lib = _lib.load()
    n = info.num_active_orbs
    P = lay.n_params
    rng = np.random.default_rng(2024)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    H = hamiltonian_0i_0a(h, g, 0, n)
    peak, _ = measured_peak()

    def timed(fn, reps=2):
        best, res = None, None
        for _ in range(reps + 1):  # first pass is the warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            res = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best, res

    # the bench state is a tUPS state on the closed-shell reference: spin-flip symmetric, so sq_sigma / sq_rdm12 (which MEASURE the
    # symmetry of their input on every call) work on the determinants above the diagonal only; the full builds are timed beside them
    _lib.check(lib.sq_set_option(b"sigma_spinsym", b"0"))
    ms_sigma_full, sig_full = timed(lambda: osa.propagate_state([H], state, info), reps=1)
    ms_rdm_full, _ = timed(lambda: osa.reduced_density_matrices(state, state, info), reps=1)
    _lib.check(lib.sq_set_option(b"sigma_spinsym", b"1"))
    ms_sigma, sig = timed(lambda: osa.propagate_state([H], state, info))
    sigma_half_vs_full = float(torch.max(torch.abs(sig - sig_full)))
    del sig_full
    energy = float(torch.dot(state, sig))
    ms_rdm, (d1, d2) = timed(lambda: osa.reduced_density_matrices(state, state, info))
    e_rdm = float(np.sum(h * d1) + 0.5 * np.sum(g * d2))
    handle = osa.compile_layout(info, lay)
    # the sweep runs one launch per brick [sa_single, double, sa_single] on an orbital pair (p, p+1); a brick touches the
    # amplitudes whose alpha or beta string has exactly one of the two orbitals occupied (78.2 % at n = 16)
    from math import comb

    na, nb = info.num_active_elec_alpha, info.num_active_elec_beta
    inert_a = 1.0 - 2.0 * comb(n - 2, na - 1) / comb(n, na)
    inert_b = 1.0 - 2.0 * comb(n - 2, nb - 1) / comb(n, nb)
    touched = (P // 3) * (1.0 - inert_a * inert_b) * info.num_det
    bra = osa.construct_ups_state(sig, info, thetas, lay, dagger=True)   # U^d H|psi>
    ket = torch.zeros_like(state)
    ket[0] = 1.0
    g_out = np.zeros(P)
    import ctypes as C

    PD = C.POINTER(C.c_double)
    th = np.ascontiguousarray(thetas, dtype=np.float64)

    sweep_launches = [0]

    def sweep():
        b, k = bra.clone(), ket.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        l0 = lib.sq_launch_count()
        e0.record()
        _lib.check(lib.sq_ups_grad_sweep(info._handle, handle, th.ctypes.data_as(PD), 0, P, C.c_void_p(b.data_ptr()), C.c_void_p(k.data_ptr()),
                                         g_out.ctypes.data_as(PD), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        e1.record()
        torch.cuda.synchronize()
        sweep_launches[0] = int(lib.sq_launch_count() - l0)
        return e0.elapsed_time(e1)

    sweep()
    ms_sweep = min(sweep() for _ in range(2))
    # Algorithmic bytes of the sweep as it is launched: quad_grad_kernel takes TWO commuting bricks per read + write of bra and
    # ket and touches the amplitudes whose alpha or beta string is active in either pair; a single-brick launch touches those of
    # one pair.  `one_brick_equivalent` keeps round 1's accounting (32 B x touched per brick) for comparison.
    n_bricks = P // 3

    def inert2(ne):   # fraction of strings with 0 or 2 electrons on each of two disjoint orbital pairs
        return sum(comb(n - 4, ne - o1 - o2) for o1 in (0, 2) for o2 in (0, 2) if 0 <= ne - o1 - o2 <= n - 4) / comb(n, ne)

    # every gradient launch is a kernel + its reduction (2 counted launches): L = 2 (quads + singles), bricks = 2 quads + singles
    n_quads = max(0, min(n_bricks // 2, n_bricks - sweep_launches[0] // 2)) if sweep_launches[0] else 0
    n_single = n_bricks - 2 * n_quads
    touched_plan = (n_quads * (1.0 - inert2(na) * inert2(nb)) + n_single * (1.0 - inert_a * inert_b)) * info.num_det
    sweep_gbs = 32.0 * touched_plan / (ms_sweep * 1e-3) / 1e9
    sweep_gbs_one_brick = 32.0 * touched / (ms_sweep * 1e-3) / 1e9
    del bra
    # one whole evaluation through the call the wave-function object makes (sq_ups_energy_grad: state, sigma, dot, and the
    # gradient sweep run BACKWARDS from (H|psi>, |psi>) -- no adjoint pass), checked at size against the step-by-step route
    # psi = U|HF>, E = <psi|H|psi>, forward sweep from (U^d H psi, |HF>)
    ms_call, (e_call, g_call) = timed(lambda: osa.ups_energy_and_gradient(ket, info, thetas, lay, H), reps=1)
    psi = osa.construct_ups_state(ket, info, thetas, lay)
    hpsi = osa.propagate_state([H], psi, info)
    e_steps = float(torch.dot(psi, hpsi))
    del psi
    bra = osa.construct_ups_state(hpsi, info, thetas, lay, dagger=True)
    del hpsi
    sweep()
    g_steps = g_out.copy()
    del bra
    return {
        "energy_and_gradient_call_ms": ms_call,
        "energy_call_minus_stepwise": e_call - e_steps,
        "gradient_call_vs_forward_sweep_maxdiff": float(np.max(np.abs(g_call - g_steps))),
        "gradient_call_norm": float(np.linalg.norm(g_call)),
        "sigma_ms": ms_sigma,
        "rdm12_ms": ms_rdm,
        "sigma_full_build_ms": ms_sigma_full,
        "rdm12_full_build_ms": ms_rdm_full,
        "sigma_half_vs_full_build_maxdiff": sigma_half_vs_full,
        "energy_sigma": energy,
        "energy_rdm": e_rdm,
        "energy_diff": energy - e_rdm,
        "trace_rdm1": float(np.trace(d1)),
        "gradient_sweep_ms": ms_sweep,
        "gradient_sweep_ms_per_operator": ms_sweep / P,
        "gradient_sweep_ms_per_brick": ms_sweep / (P // 3),
        "gradient_sweep_launches": {"two_bricks": n_quads, "one_brick": n_single},
        "gradient_sweep_algorithmic_GBps": sweep_gbs,
        "gradient_sweep_frac_of_measured_hbm_peak": sweep_gbs / peak,
        "gradient_sweep_one_brick_equivalent_GBps": sweep_gbs_one_brick,
        "gradient_norm": float(np.linalg.norm(g_out)),
        "note": "synthetic symmetric integrals (default_rng(2024)); sigma / RDM: D-panel gathers + hand-written DMMA kernels (fp64 tensor pipe) over the determinants above the diagonal of the spin-flip symmetric state (full builds timed beside them); the sweep is HBM-bound",
    }


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the engine")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from slowquant_b200 import _lib
    from slowquant_b200.ci_spaces import get_indexing
    from slowquant_b200.operator_state_algebra import _ups_apply_inplace, compile_layout, construct_ups_state
    from slowquant_b200.util import UpsStructure

    lib = _lib.load()
    n, L = args.cas, args.layers
    ne = n // 2
    info = get_indexing(0, n, 0, ne, ne, device=local_rank)
    lay = UpsStructure()
    lay.create_tiled(n, {"n_layers": L, "do_tups": True})
    P = lay.n_params
    rng = np.random.default_rng(1234 + rank)
    thetas = rng.uniform(-np.pi, np.pi, P)
    handle = compile_layout(info, lay)
    launches_per_step = int(lib.sq_layout_num_launches(handle, 0, P))
    touched_per_step = int(lib.sq_layout_touched_amplitudes(handle, 0, P))
    import ctypes as C

    plan = (C.c_int64 * 6)()
    _lib.check(lib.sq_layout_plan_stats(handle, 0, P, plan))
    plan = [int(x) for x in plan]  # launches, window sweeps, bricks in window sweeps, quad, single-brick, other

    state = torch.zeros(info.num_det, dtype=torch.float64, device=dev)
    state[0] = 1.0  # HF determinant ("1"*ne*2 + "0"*...), index 0

    def step():
        _ups_apply_inplace(state, info, thetas, lay, 0, P, False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = int(lib.sq_launch_count())
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = int(lib.sq_launch_count()) - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    norm = float(torch.linalg.norm(state))

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * L * args.steps / (ms_max * 1e-3)

    # ---- end to end through the public API with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        host_in = torch.zeros(info.num_det, dtype=torch.float64).pin_memory()
        host_in[0] = 1.0
        np_in = host_in.numpy()
        out = construct_ups_state(np_in, info, thetas, lay)  # warm-up (pinned pool, layout cache)
        del out
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            out = None  # release the previous result so its page-locked block is reused, as a caller loop would
            out = construct_ups_state(np_in, info, thetas, lay)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {
            "value": world * L * n_e2e / float(t2.item()),
            "unit": UNIT,
            "h2d_bytes_per_step": int(8 * info.num_det + 8 * P),
            "d2h_bytes_per_step": int(8 * info.num_det),
            "steps": n_e2e,
            "norm_check": float(np.linalg.norm(out)),
        }
        # the batch call of the same API (construct_ups_state_SA on [S, N_det] host states: RotoSolve shifts, state-averaged
        # ensembles): copies of neighbouring states overlap the kernels on three streams
        try:
            if world > 1:
                raise RuntimeError("skipped for N > 1 (one pinned batch per rank)")
            from slowquant_b200.operator_state_algebra import construct_ups_state_SA

            S = 6
            batch_in = torch.zeros((S, info.num_det), dtype=torch.float64).pin_memory()
            batch_in[:, 0] = 1.0
            outb = construct_ups_state_SA(batch_in.numpy(), info, thetas, lay)  # warm-up: streams, and the page-locked result block the timed call reuses
            del outb
            barrier()
            t0 = time.perf_counter()
            outb = construct_ups_state_SA(batch_in.numpy(), info, thetas, lay)
            torch.cuda.synchronize()
            dtb = time.perf_counter() - t0
            t3 = torch.tensor([dtb], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            e2e["batched"] = {
                "value": world * L * S / float(t3.item()),
                "unit": UNIT,
                "states_per_call": S,
                "api": "construct_ups_state_SA(host_states[S, N_det]) -- H2D of state k+1 and D2H of state k-1 overlap the kernels of state k",
                "norm_check": float(np.linalg.norm(outb[S - 1])),
                "max_diff_vs_single_call": float(np.max(np.abs(outb[S - 1] - out))),
            }
            del outb, batch_in
        except Exception as exc:  # an extra, never a gate
            e2e["batched"] = {"error": repr(exc)}
        # the call an optimiser makes: `WF.thetas = x` on a WaveFunctionUPS whose reference state is resident (H2D: the
        # parameters; D2H: one amplitude as the completion fence) -- what a user pays per state construction when the
        # vector never has to leave the device
        try:
            if world > 1:
                raise RuntimeError("skipped for N > 1")
            from slowquant_b200.integral_manager import ArrayIntegrals
            from slowquant_b200.ups_wavefunction import WaveFunctionUPS

            rng_i = np.random.default_rng(2024)
            h_syn = rng_i.normal(size=(n, n))
            g_syn = np.zeros((n, n, n, n))
            WF = WaveFunctionUPS((n, n), np.eye(n), ArrayIntegrals(h_syn + h_syn.T, g_syn, num_elec=n), "tUPS", {"n_layers": L}, device=local_rank)
            th_list = thetas.tolist()
            n_set = max(n_e2e, 8)     # 42 ms per call: enough calls that one allocator hiccup does not decide the number

            def time_setter(light_cone):
                WF.light_cone = light_cone
                for _ in range(3):   # warm-up calls: the setter keeps the old state alive until the new one exists, so the caching
                    WF.thetas = th_list   # allocator needs two 1.3 GB blocks (carved out of what the earlier stages left) before it stops calling cudaMalloc
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(n_set):
                    WF.thetas = th_list
                    pr = float(WF.ci_coeffs_device[0].item())
                return time.perf_counter() - t0, pr, WF.ci_coeffs_device.clone()

            dtw, probe, ref_state = time_setter(False)     # every operator on the full vector: the route of `value`
            dtl, probe_l, cone_state = time_setter(True)   # the default of the class: head of the circuit in its light-cone window
            cone_diff = float(torch.max(torch.abs(cone_state - ref_state)))
            del ref_state, cone_state
            e2e["wavefunction_setter"] = {
                "value": L * n_set / dtw,
                "calls": n_set,
                "light_cone": {"value": L * n_set / dtl, "unit": UNIT, "max_diff_vs_full_route": cone_diff,
                               "note": "class default: the operators that are identities on the HF determinant are dropped and the first 4 layers run in the CAS(14,14) window their light cone reaches (11.8 M determinants), embedded, then 12 layers on the full vector; same state"},
                "unit": UNIT,
                "api": "WaveFunctionUPS.thetas = x (reference state resident on the device), one amplitude read back",
                "h2d_bytes_per_step": int(8 * P),
                "d2h_bytes_per_step": 8,
                "probe": probe,
            }
            # one optimiser iteration through the class: fun(x) and jac(x) at the same (new) parameters -- one state construction,
            # one H|psi>, the backwards gradient sweep
            try:
                th_it = [x + 1e-3 for x in th_list]
                for rep in range(2):   # first pass: warm-up (sigma work buffers, allocator)
                    th_it = [x + 1e-3 for x in th_it]
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    e_it = WF._calc_energy_optimization(th_it, True, False)
                    g_it = WF._calc_gradient_optimization(th_it, True, False)
                    torch.cuda.synchronize()
                    dti = time.perf_counter() - t0
                e2e["wavefunction_setter"]["optimizer_iteration"] = {
                    "ms": 1e3 * dti, "energy": float(e_it), "gradient_norm": float(np.linalg.norm(g_it)),
                    "api": "WaveFunctionUPS._calc_energy_optimization(x) + _calc_gradient_optimization(x), what scipy.optimize.minimize calls per iteration",
                }
            except Exception as exc:  # an extra of an extra
                e2e["wavefunction_setter"]["optimizer_iteration"] = {"error": repr(exc)}
            del WF
        except Exception as exc:  # an extra, never a gate
            e2e["wavefunction_setter"] = {"error": repr(exc)}
        del out, host_in, np_in

    # ---- the same circuit on a resident batch of S vectors in ONE launch sequence (state-averaged wave functions, the common
    # tail of RotoSolve's shifted states): the window / gauge sweeps carry the state index as a batch dimension ----
    batched = None
    if not args.no_extras and world == 1:
        try:
            from slowquant_b200.operator_state_algebra import _ups_apply_batch_inplace

            S = 4
            batch = torch.zeros((S, info.num_det), dtype=torch.float64, device=dev)
            batch[:, 0] = 1.0
            _ups_apply_batch_inplace(batch, info, thetas, lay, 0, P, False)   # warm-up
            torch.cuda.synchronize()
            l0 = int(lib.sq_launch_count())
            ev0.record()
            reps = max(1, min(args.steps, 3))
            for _ in range(reps):
                _ups_apply_batch_inplace(batch, info, thetas, lay, 0, P, False)
            ev1.record()
            torch.cuda.synchronize()
            msb = ev0.elapsed_time(ev1)
            batched = {
                "value": S * L * reps / (msb * 1e-3),
                "unit": UNIT,
                "states_per_call": S,
                "launches_per_call": (int(lib.sq_launch_count()) - l0) // reps,
                "api": "sq_ups_apply_batch on a resident [S, N_det] batch (construct_ups_state_SA with a device tensor)",
                "max_diff_vs_single_state": float(torch.max(torch.abs(batch[S - 1] - batch[0]))),
            }
            del batch
        except Exception as exc:  # an extra, never a gate
            batched = {"error": repr(exc)}

    extras = None
    if not args.no_extras and world == 1:
        try:
            extras = energy_gradient_extras(info, lay, thetas, state, dev)
        except Exception as exc:  # extras never gate the headline line
            extras = {"error": repr(exc)}

    if rank == 0:
        peak, peak_src = measured_peak()
        launch_ms = ms_max / max(launches, 1)
        bytes_per_launch = 16.0 * touched_per_step / max(launches_per_step, 1)
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
            if n == 16 and plan[1]:
                with open(os.path.join(ROOT, "profiles", "r2_win_kernel_traffic.json")) as f:
                    traffic = float(json.load(f)["dram_bytes_per_launch"])
        except Exception:
            traffic = None
        roofline = {
            "kernel": "win_kernel (one read + one write of the vector per sweep; %.1f bricks = %.1f ansatz operators per sweep)"
            % (plan[2] / max(plan[1], 1), 3.0 * plan[2] / max(plan[1], 1)) if plan[1] else "tile/quad kernels (one or two bricks per launch)",
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "peak_source": peak_src,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": traffic,
            "traffic_source": "ncu --set full capture of win_kernel, profiles/r2_win_kernel_ncu_full_summary.csv" if traffic else None,
            "algorithmic_bytes_per_launch": bytes_per_launch,
            "avg_launch_ms": launch_ms,
            "full_sweep_equiv_GBps": 16.0 * info.num_det / (launch_ms * 1e-3) / 1e9,
            # the traffic the reference's algorithm needs for the same work (one read + one write of the vector per
            # ansatz operator) divided by our time: how far operator fusion lifts the path above the
            # per-operator HBM roofline
            "per_operator_algorithm_GBps": 16.0 * info.num_det * P * args.steps / (ms_max * 1e-3) / 1e9,
        }
        # the roofline that actually binds win_kernel: every brick moves its touched amplitudes through shared memory
        # once (LDS + STS, 16 B each); peak = 128 B per clock and SM (B300_MICROARCH.md) at the sampled SM clock
        if plan[1]:
            from math import comb

            inert = 1.0 - 2.0 * comb(n - 2, ne - 1) / comb(n, ne)
            smem_bytes = 16.0 * plan[2] * (1.0 - inert * inert) * info.num_det * args.steps
            smem_peak = 148 * 128 * (clocks.get("sm_mhz") or 1965) * 1e6 / 1e9
            roofline["shared_memory"] = {
                "bound": "shared memory bandwidth (brick phases of win_kernel)",
                "achieved": smem_bytes / (ms_max * 1e-3) / 1e9,
                "peak": smem_peak,
                "unit": "GB/s",
                "frac": smem_bytes / (ms_max * 1e-3) / 1e9 / smem_peak,
                "note": "whole-step average including the HBM phases of every sweep; ncu: LSU data pipe 70-75 % over the kernel",
            }
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cpu = cpu_sample(n)
                cpu.pop("seconds_per_sample", None)
            except Exception as exc:  # the baseline is a reported number, never a gate
                cpu = {"error": repr(exc)}
        line = {
            "metric": METRIC if n == 16 else f"tUPS layer applications/s at CAS({n},{n})",
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": workload_name(n, L),
                "l2_policy": "inputs larger than L2 (1.325 GB vector vs 126 MB L2)",
                "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (one CAS vector per GPU), no collective",
                "fusion": "%d launches per step: %d window sweeps holding %d bricks (3 operators each), %d quad, %d single-brick, "
                "2 gauge sweeps" % (launches_per_step, plan[1], plan[2], plan[3], plan[4]),
            },
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "state_norm": norm,
            "batched_resident": batched,
            "energy_gradient": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sharded(args) -> None:
    """N > 1: ONE CAS(n,n) vector sharded by alpha string over the N GPUs (strong scaling).  The circuit runs as a few local
    phases (window sweeps on the shard) in two row layouts, with an all-to-all re-shard over NVLink peer memory between them
    (slowquant_b200/distributed.py)."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the engine")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)

    from slowquant_b200 import _lib
    from slowquant_b200 import distributed as sqd
    from slowquant_b200.distributed import ShardedSpace, construct_ups_state_sharded, dot_sharded, reshard_schedule
    from slowquant_b200.operator_state_algebra import compile_layout
    from slowquant_b200.util import UpsStructure

    lib = _lib.load()
    n, L = args.cas, args.layers
    ne = n // 2
    sp = ShardedSpace(0, n, 0, ne, ne, device=local_rank)
    lay = UpsStructure()
    lay.create_tiled(n, {"n_layers": L, "do_tups": True})
    P = lay.n_params
    thetas = np.random.default_rng(1234).uniform(-np.pi, np.pi, P)   # same parameters on every rank
    handle = compile_layout(sp.ci_info, lay)
    use_reshard = sp.reshard_ok and sqd._RESHARD_DEFAULT
    st = sp.alloc_state()
    st.set_determinant(0)
    nb = sp.ci_info.num_beta_strings

    # ---- plan accounting (per rank): phases, sweeps, amplitudes the kernels read and write, rows that cross NVLink ----
    pstats = [0] * 8
    n_reshards = 0
    nvlink_bytes = 0.0   # bytes this rank writes into OTHER ranks' buffers per step
    gauge_amps = 0
    if use_reshard:
        phases = reshard_schedule(lay.excitation_operator_type, lay.excitation_indices, n, world)
        handle_B = compile_layout(sp.ci_info_B, lay) if any(nm == "B" for nm, _ in phases) else None
        a2b, b2a = sp.reshard_tables_host
        where = "A"
        for nm, ops in phases:
            arr = np.asarray(ops, dtype=np.int32)
            one = (C.c_int64 * 8)()
            tgt = "A" if nm == "X" else nm
            if tgt != where:
                tab = a2b if tgt == "B" else b2a
                nvlink_bytes += 8.0 * nb * float(np.count_nonzero(tab[0] != rank))
                n_reshards += 1
                where = tgt
            if nm == "X":
                continue
            _lib.check(lib.sq_layout_plan_stats_list(handle if nm == "A" else handle_B, len(arr), arr.ctypes.data_as(C.POINTER(C.c_int32)), one))
            pstats = [a + int(b) for a, b in zip(pstats, one)]
        if where == "B":
            nvlink_bytes += 8.0 * nb * float(np.count_nonzero(b2a[0] != rank))
            n_reshards += 1
        gauge_amps = 2 * max(sp.local_len, sp.local_len_B)   # into the sign-free gauge once, out of it once
        n_phase = len(phases)
        plan_desc = (f"{n_phase} local phases per step in two row layouts (rows grouped by the first / the last {world.bit_length() - 1} "
                     f"orbitals), {n_reshards} all-to-all re-shards over NVLink peer memory (bulk-copy engine), one device-wide barrier each")
    else:
        plan = sp.exchange_plan(lay, 0, P, False)
        for first, last, _ in plan:
            one = (C.c_int64 * 6)()
            _lib.check(lib.sq_layout_plan_stats(handle, first, last, one))
            pstats = [a + int(b) for a, b in zip(pstats[:6], one)] + [0, 0]
        pstats[7] = int(lib.sq_layout_touched_amplitudes(handle, 0, P))
        n_exchange = sum(1 for _, _, x in plan if x)
        plan_desc = f"{n_exchange} of {len(plan)} operator ranges per step exchange tiles over NVLink peer memory, the rest are local"
    touched_per_step = pstats[7] + gauge_amps

    def step():
        construct_ups_state_sharded(st, thetas, lay)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = int(lib.sq_launch_count())
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = int(lib.sq_launch_count()) - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    norm2 = dot_sharded(st, st)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = L * args.steps / (ms_max * 1e-3)

    # ---- the re-shard alone: NVLink bytes per second of one A -> B -> A round trip (timed on the device, max over ranks) ----
    nvlink = None
    if use_reshard and st._ptr_B is not None:
        reps = 5
        barrier()
        ev0.record()
        for _ in range(reps):
            sqd._reshard(st, to_B=True)
            sqd._reshard(st, to_B=False)
        ev1.record()
        barrier()
        tr = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        out_b = 8.0 * nb * float(np.count_nonzero(sp.reshard_tables_host[0][0] != rank))
        ob = torch.tensor([out_b], dtype=torch.float64, device=dev)
        dist.all_reduce(ob, op=dist.ReduceOp.MAX)
        ms_one = float(tr.item()) / (2 * reps)
        nvlink = {
            "kernel": "reshard_bulk_kernel (cp.async.bulk global -> shared -> peer global), one launch + one barrier per re-shard",
            "bound": "nvlink",
            "ms_per_reshard": ms_one,
            "bytes_out_per_gpu": float(ob.item()),
            "achieved": float(ob.item()) / (ms_one * 1e-3) / 1e9,
            "peak": 770.0,
            "peak_source": "measured peer copy per direction (B200_PROFILING.md); nominal 900",
            "unit": "GB/s",
            "frac": float(ob.item()) / (ms_one * 1e-3) / 1e9 / 770.0,
            "reshards_per_step": n_reshards,
            "share_of_step": n_reshards * ms_one / (ms_max / args.steps),
        }

    # ---- multi-GPU parity inside the driver's run: the sharded engine against the single-GPU engine on this rank's own GPU,
    # CAS(12,12), dense random vector, L = 4 (the single-GPU engine is pinned to the oracle / reference goldens in tests/) ----
    parity = None
    try:
        from slowquant_b200.ci_spaces import get_indexing
        from slowquant_b200.operator_state_algebra import construct_ups_state

        pn, pe, pL = 12, 6, 4
        psp = ShardedSpace(0, pn, 0, pe, pe, device=local_rank)
        pinfo = get_indexing(0, pn, 0, pe, pe, device=local_rank)
        play = UpsStructure()
        play.create_tiled(pn, {"n_layers": pL, "do_tups": True})
        prng = np.random.default_rng(4321)
        pth = prng.uniform(-np.pi, np.pi, play.n_params)
        full = torch.from_numpy(prng.normal(size=pinfo.num_det)).to(dev)
        full /= torch.linalg.norm(full)
        ref = construct_ups_state(full, pinfo, pth.tolist(), play)
        pst = psp.alloc_state()
        lo, hi = psp.row_begin * pinfo.num_beta_strings, psp.row_end * pinfo.num_beta_strings
        pst.local.copy_(full[lo:hi])
        construct_ups_state_sharded(pst, pth, play)
        torch.cuda.synchronize()
        perr = torch.tensor([float(torch.max(torch.abs(pst.local - ref[lo:hi]))) if hi > lo else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(perr, op=dist.ReduceOp.MAX)
        pst.close()
        parity = {"check": f"sharded vs single-GPU engine, CAS({pn},{pn}) L={pL}, dense random vector, max over ranks of max|diff| on the rank's rows",
                  "max_abs_diff": float(perr.item()), "tolerance": 1e-12, "ok": bool(float(perr.item()) < 1e-12)}
    except Exception as exc:  # an extra, never a gate
        parity = {"error": repr(exc)}

    e2e = None
    if not args.no_e2e:
        host_in = torch.zeros(sp.local_len, dtype=torch.float64).pin_memory()
        host_out = torch.empty(sp.local_len, dtype=torch.float64).pin_memory()
        if rank == 0:
            host_in[0] = 1.0
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            st.local.copy_(host_in, non_blocking=True)
            construct_ups_state_sharded(st, thetas, lay)
            host_out.copy_(st.local)
        torch.cuda.synchronize()
        t2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        sq = torch.tensor([float(torch.dot(host_out, host_out))], dtype=torch.float64, device=dev)
        dist.all_reduce(sq)
        e2e = {
            "value": L * n_e2e / float(t2.item()),
            "unit": UNIT,
            "h2d_bytes_per_step": int(8 * sp.ci_info.num_det + 8 * P * world),
            "d2h_bytes_per_step": int(8 * sp.ci_info.num_det),
            "steps": n_e2e,
            "norm_check": float(sq.item()) ** 0.5,
            "api": "every rank copies its rows from pinned host memory into its shard, construct_ups_state_sharded, rows back to the host",
        }
    touched_all = torch.tensor([float(touched_per_step)], dtype=torch.float64, device=dev)
    dist.all_reduce(touched_all)
    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per_step_all = 16.0 * float(touched_all.item())
        achieved = bytes_per_step_all * args.steps / (ms_max * 1e-3) / 1e9 / world
        line = {
            "metric": METRIC if n == 16 else f"tUPS layer applications/s at CAS({n},{n})",
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": workload_name(n, L),
                "l2_policy": "inputs larger than L2" if 8 * sp.local_len > 126e6 else "per-GPU shard may fit L2 (strong scaling of a fixed vector)",
                "parallelism": f"one vector sharded by alpha string over {world} GPUs; " + plan_desc,
                "fusion": "per rank and step: %d window sweeps holding %d bricks (3 operators each), %d quad, %d single-brick launches"
                % (pstats[1], pstats[2], pstats[3], pstats[4]),
            },
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "win_kernel on the rank's shard (HBM bytes of all sweeps of a step over the WHOLE step time, re-shards and barriers included; per-GPU average)",
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "peak_source": peak_src,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": None,
                "nvlink": nvlink,
            },
            "cpu_baseline": None,
            "state_norm": norm2 ** 0.5,
            "multi_gpu_parity": parity,
        }
        print(json.dumps(line), flush=True)
    st.close()
    dist.destroy_process_group()


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args)
    elif world > 1 and args.mode != "replicas":
        # the multi-GPU path north_star names: ONE vector sharded by alpha string (strong scaling).  --mode replicas keeps the
        # independent-vectors-per-GPU run (RotoSolve shifts, finite-difference columns) for comparison.
        run_sharded(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
