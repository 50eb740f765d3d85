"""Times energy (sigma), 1/2-RDM and the fused gradient sweep at a given CAS (diagnostic)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from slowquant_b200 import operator_state_algebra as osa  # noqa: E402
from slowquant_b200.ci_spaces import get_indexing  # noqa: E402
from slowquant_b200.operators import hamiltonian_0i_0a  # noqa: E402
from slowquant_b200.util import UpsStructure  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ne = n // 2
info = get_indexing(0, n, 0, ne, ne)
rng = np.random.default_rng(2024)
A = rng.normal(size=(n, n))
h = A + A.T
B = 0.1 * rng.normal(size=(n, n, n, n))
g = B + B.transpose(1, 0, 2, 3)
g = g + g.transpose(0, 1, 3, 2)
g = g + g.transpose(2, 3, 0, 1)
lay = UpsStructure()
lay.create_tiled(n, {"n_layers": L, "do_tups": True})
th = rng.uniform(-np.pi, np.pi, lay.n_params)
dev = torch.device("cuda", info.device)
csf = torch.zeros(info.num_det, dtype=torch.float64, device=dev)
csf[0] = 1.0


def timed(label, fn, reps=2):
    best = 1e9
    r = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"CAS({n},{n}) {label:28s} {best*1e3:10.1f} ms", flush=True)
    return r


ci = timed(f"state L={L} ({lay.n_params} ops)", lambda: osa.construct_ups_state(csf, info, th, lay))
H = hamiltonian_0i_0a(h, g, 0, n)
sig = timed("sigma H|psi>", lambda: osa.propagate_state([H], ci, info))
e = float(torch.dot(ci, sig))
d1, d2 = timed("rdm1+rdm2", lambda: osa.reduced_density_matrices(ci, ci, info))
e_rdm = float(np.sum(h * d1) + 0.5 * np.sum(g * d2))
print("E(sigma) =", e, " E(rdm) =", e_rdm, " diff =", e - e_rdm, " tr rdm1 =", np.trace(d1))


def grad():
    bra = osa.construct_ups_state(sig, info, th, lay, dagger=True)
    return osa.ups_gradient_sweep(bra, csf, info, th, lay)[0]


gth = timed("theta gradient sweep", grad)
k = lay.n_params - 2
step = 1e-5
tp, tm = th.copy(), th.copy()
tp[k] += step
tm[k] -= step
cp = osa.construct_ups_state(csf, info, tp, lay)
cm = osa.construct_ups_state(csf, info, tm, lay)
ep = osa.expectation_value(cp, [H], cp, info)
em = osa.expectation_value(cm, [H], cm, info)
print(f"grad[{k}] =", gth[k], " finite diff =", (ep - em) / (2 * step))

from slowquant_b200 import _lib  # noqa: E402

lib = _lib.load()
h_lay = osa.compile_layout(info, lay)
touched = lib.sq_layout_touched_amplitudes(h_lay, 0, lay.n_params)
bra = osa.construct_ups_state(sig, info, th, lay, dagger=True)


def sweep_only():
    b = bra.clone()
    k_ = csf.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    osa.ups_gradient_sweep(b, k_, info, th, lay)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


best = min(sweep_only() for _ in range(3))
# ups_gradient_sweep copies its inputs (2 x 16 B/amplitude extra); subtract nothing, report as is
print(f"gradient sweep (incl. input copies): {best*1e3:.1f} ms; algorithmic 32 B x touched = {32*touched/1e9:.2f} GB "
      f"-> {32*touched/best/1e9:.0f} GB/s")
