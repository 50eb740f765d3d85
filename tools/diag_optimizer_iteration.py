"""Where do the 886 ms of one optimiser iteration through WaveFunctionUPS go (open item)?

    python tools/diag_optimizer_iteration.py [n] [L]

Times, with a device synchronisation around every part, the pieces of fun(x) + jac(x) at CAS(n,n): the `thetas` setter (light cone
on / off), H|psi> through `propagate_state`, the dot product, the two clones and the backwards sweep of
`ups_gradient_sweep_backward`, the fused library call on the same objects, and the two optimisation callables themselves; every
part three times (the first pass warms tables and the allocator).  Also prints the kernel-launch counts of the parts and the
allocator statistics (cudaMalloc calls of the caching allocator) per pass."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from slowquant_b200 import _lib  # noqa: E402
from slowquant_b200 import operator_state_algebra as osa  # noqa: E402
from slowquant_b200.integral_manager import ArrayIntegrals  # noqa: E402
from slowquant_b200.operators import hamiltonian_0i_0a  # noqa: E402
from slowquant_b200.ups_wavefunction import WaveFunctionUPS  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rng = np.random.default_rng(2024)
A = rng.normal(size=(n, n))
h = A + A.T
B = 0.1 * rng.normal(size=(n, n, n, n))
g = B + B.transpose(1, 0, 2, 3)
g = g + g.transpose(0, 1, 3, 2)
g = g + g.transpose(2, 3, 0, 1)
WF = WaveFunctionUPS((n, n), np.eye(n), ArrayIntegrals(h, g, num_elec=n), "tUPS", {"n_layers": L})
lib = _lib.load()
th = np.random.default_rng(1234).uniform(-np.pi, np.pi, len(WF.thetas))


def timed(label, fn):
    torch.cuda.synchronize()
    l0, m0 = lib.sq_launch_count(), torch.cuda.memory_stats().get("num_device_alloc", 0)
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"  {label:58s} {1e3 * dt:9.1f} ms  {lib.sq_launch_count() - l0:5d} launches  "
          f"{torch.cuda.memory_stats().get('num_device_alloc', 0) - m0:3d} cudaMalloc", flush=True)
    return out


for rep in range(3):
    print(f"pass {rep}", flush=True)
    x = (th + 1e-3 * rep).tolist()
    WF.light_cone = False
    timed("thetas setter, every operator on the full vector", lambda: setattr(WF, "thetas", x))
    WF.light_cone = True
    timed("thetas setter, light cone", lambda: setattr(WF, "thetas", x))
    H = hamiltonian_0i_0a(WF.h_mo, WF.g_mo, WF.num_inactive_orbs, WF.num_active_orbs)
    sigma = timed("H|psi> (propagate_state)", lambda: osa.propagate_state([H], WF.ci_coeffs_device, WF.ci_info))
    timed("<psi|H|psi> (dot)", lambda: osa._dot(WF.ci_coeffs_device, sigma, WF.ci_info))
    timed("two clones (bra, ket)", lambda: (sigma.clone(), WF.ci_coeffs_device.clone()))
    timed("ups_gradient_sweep_backward", lambda: osa.ups_gradient_sweep_backward(sigma, WF.ci_coeffs_device, WF.ci_info, x, WF.ups_layout))
    timed("fused call ups_energy_and_gradient (from the reference)", lambda: osa.ups_energy_and_gradient(WF._csf_dev, WF.ci_info, x, WF.ups_layout, H))
    y = (th + 1e-3 * rep + 5e-4).tolist()
    timed("WF._calc_energy_optimization(y)", lambda: WF._calc_energy_optimization(y, True, False))
    timed("WF._calc_gradient_optimization(y)", lambda: WF._calc_gradient_optimization(y, True, False))
    del sigma
