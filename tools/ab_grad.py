"""A/B of the theta-gradient sweep at CAS(n,n), L layers: one fused brick per launch (wingrad=0) against the window gradient
kernel with several window configurations.   python tools/ab_grad.py [n] [L] [cfg ...]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from slowquant_b200 import _lib  # noqa: E402
from slowquant_b200 import operator_state_algebra as osa  # noqa: E402
from slowquant_b200.ci_spaces import get_indexing  # noqa: E402
from slowquant_b200.util import UpsStructure  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfgs = sys.argv[3:] or ["5:4:0,40,3,16,2", "5:0:0,40,3,16,2", "6:5:4,72,3,16,3", "4:0:0,40,3,16,2"]
ne = n // 2
info = get_indexing(0, n, 0, ne, ne)
lay = UpsStructure()
lay.create_tiled(n, {"n_layers": L, "do_tups": True})
th = np.random.default_rng(1).uniform(-np.pi, np.pi, lay.n_params)
dev = torch.device("cuda", info.device)
gen = torch.Generator(device=dev)
gen.manual_seed(5)
bra = torch.randn(info.num_det, dtype=torch.float64, device=dev, generator=gen)
ket = torch.randn(info.num_det, dtype=torch.float64, device=dev, generator=gen)
bra /= torch.linalg.norm(bra)
ket /= torch.linalg.norm(ket)
lib = _lib.load()
lay_h = osa.compile_layout(info, lay)
PD = __import__("ctypes").POINTER(__import__("ctypes").c_double)


def sweep(label):
    best, g = 1e9, None
    for _ in range(3):
        b, k = bra.clone(), ket.clone()
        out = np.zeros(lay.n_params)
        torch.cuda.synchronize()
        l0 = lib.sq_launch_count()
        t0 = time.perf_counter()
        _lib.check(lib.sq_ups_grad_sweep(info._handle, lay_h, th.ctypes.data_as(PD), 0, lay.n_params, osa._ptr(b), osa._ptr(k),
                                         out.ctypes.data_as(PD), osa._stream()))
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
        g, launches = out, lib.sq_launch_count() - l0
    print(f"CAS({n},{n}) L={L} {label:34s} {best*1e3:9.1f} ms  {launches:4d} launches", flush=True)
    return g, b, k


lib.sq_set_option(b"wingrad", b"0")
lib.sq_set_option(b"quadgrad", b"0")
g0, b0, k0 = sweep("one brick per launch")
lib.sq_set_option(b"quadgrad", b"1")
gq, bq, kq = sweep("two bricks per launch (quad)")
print("    max|grad diff| %.2e  max|bra diff| %.2e  max|ket diff| %.2e" % (
    float(np.max(np.abs(gq - g0))), float(torch.max(torch.abs(bq - b0))), float(torch.max(torch.abs(kq - k0)))), flush=True)
lib.sq_set_option(b"wingrad", b"1")
for cfg in ([] if cfgs == ["none"] else cfgs):
    lib.sq_set_option(b"wingrad_win", cfg.encode())
    g1, b1, k1 = sweep("window kernel " + cfg)
    print("    max|grad diff| %.2e  max|bra diff| %.2e  max|ket diff| %.2e" % (
        float(np.max(np.abs(g1 - g0))), float(torch.max(torch.abs(b1 - b0))), float(torch.max(torch.abs(k1 - k0)))), flush=True)
lib.sq_set_option(b"wingrad", b"0")
