"""A/B of a run-time switch of the sigma / RDM path at a given CAS:  python tools/ab_option.py [n] [option] [value_a] [value_b]
(defaults: 16 pipeline 0 1; e.g. `16 etab smem const`; a fifth argument `tups` takes a spin-flip symmetric tUPS state instead of a
random vector).  Prints timings of both settings (twice) and the result differences."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from slowquant_b200 import _lib  # noqa: E402
from slowquant_b200 import operator_state_algebra as osa  # noqa: E402
from slowquant_b200.ci_spaces import get_indexing  # noqa: E402
from slowquant_b200.operators import hamiltonian_0i_0a  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
opt = (sys.argv[2] if len(sys.argv) > 2 else "pipeline").encode()
va = (sys.argv[3] if len(sys.argv) > 3 else "0").encode()
vb = (sys.argv[4] if len(sys.argv) > 4 else "1").encode()
ne = n // 2
info = get_indexing(0, n, 0, ne, ne)
rng = np.random.default_rng(2024)
A = rng.normal(size=(n, n))
h = A + A.T
B = 0.1 * rng.normal(size=(n, n, n, n))
g = B + B.transpose(1, 0, 2, 3)
g = g + g.transpose(0, 1, 3, 2)
g = g + g.transpose(2, 3, 0, 1)
H = hamiltonian_0i_0a(h, g, 0, n)
dev = torch.device("cuda", info.device)
if len(sys.argv) > 5 and sys.argv[5] == "tups":
    from slowquant_b200.util import UpsStructure  # noqa: E402

    lay = UpsStructure()
    lay.create_tiled(n, {"n_layers": 4, "do_tups": True})
    hf = torch.zeros(info.num_det, dtype=torch.float64, device=dev)
    hf[0] = 1.0
    ci = osa.construct_ups_state(hf, info, np.random.default_rng(3).uniform(-np.pi, np.pi, lay.n_params).tolist(), lay)
else:
    ci = torch.randn(info.num_det, dtype=torch.float64, device=dev)
    ci /= torch.linalg.norm(ci)
lib = _lib.load()
res = {}
for mode in (va, vb, va, vb):
    lib.sq_set_option(opt, mode)
    for label, fn in (("sigma", lambda: osa.propagate_state([H], ci, info)), ("rdm12", lambda: osa.reduced_density_matrices(ci, ci, info))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"CAS({n},{n}) {opt.decode()}={mode.decode():5s} {label:6s} {dt*1e3:9.1f} ms", flush=True)
        res[(mode, label)] = out
s0, s1 = res[(va, "sigma")], res[(vb, "sigma")]
print("sigma max|b - a| =", float(torch.max(torch.abs(s0 - s1))), " |sigma| =", float(torch.linalg.norm(s0)))
(d1a, d2a), (d1b, d2b) = res[(va, "rdm12")], res[(vb, "rdm12")]
print("rdm1 diff", float(np.max(np.abs(d1a - d1b))), "rdm2 diff", float(np.max(np.abs(d2a - d2b))), "tr", float(np.trace(d1b)))
