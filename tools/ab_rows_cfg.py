"""Scan of the row-kernel geometry (threads per CTA, column chunks per row) for the sigma build at a given CAS."""
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
ci = torch.randn(info.num_det, dtype=torch.float64, device=torch.device("cuda", info.device))
ci /= torch.linalg.norm(ci)
lib = _lib.load()


def run(label):
    for what, fn in (("sigma", lambda: osa.propagate_state([H], ci, info)), ("rdm12", lambda: osa.reduced_density_matrices(ci, ci, info))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        print(f"CAS({n},{n}) {label:28s} {what:6s} {(time.perf_counter() - t0) * 1e3:9.1f} ms", flush=True)


lib.sq_set_option(b"rows", b"0")
run("rows=0")
lib.sq_set_option(b"rows", b"1")
for cfg in sys.argv[2:] or ["1024,0", "1024,13", "512,0", "512,14", "256,0", "1024,4"]:
    lib.sq_set_option(b"rows_cfg", cfg.encode())
    run("rows=1 cfg=" + cfg)
for pipe in (b"0",):
    lib.sq_set_option(b"pipeline", pipe)
    lib.sq_set_option(b"rows_cfg", b"1024,0")
    run("rows=1 cfg=1024,0 pipeline=0")
