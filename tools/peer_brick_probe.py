"""Single-process, two-GPU probe of the cross-device brick kernel (diagnostic).

Builds the rank-0 and rank-1 spaces of a 2-way alpha partition in ONE process (direct peer access, no
IPC), and times the exchange brick (0,1) on device 0 alone and on both devices concurrently.
Single process => can run under ncu.
"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from slowquant_b200 import _lib  # noqa: E402
from slowquant_b200 import operator_state_algebra as osa  # noqa: E402
from slowquant_b200.ci_spaces import CI_Info  # noqa: E402
from slowquant_b200.distributed import partition_prefix  # noqa: E402
from slowquant_b200.util import UpsStructure  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ne = n // 2
lib = _lib.load()
cudart = C.CDLL("libcudart.so.12")
for a, b in ((0, 1), (1, 0)):
    cudart.cudaSetDevice(a)
    r = cudart.cudaDeviceEnablePeerAccess(b, 0)
    print("enable peer", a, "->", b, "rc", r)
starts = partition_prefix(n, ne, 2)
spaces, shards, lays = [], [], []
lay = UpsStructure()
lay.create_tiled(n, {"n_layers": 1, "do_tups": True})
th = np.random.default_rng(0).uniform(-1, 1, lay.n_params)
for r in range(2):
    torch.cuda.set_device(r)
    info = CI_Info(0, n, 0, ne, ne, device=r, row_range=(int(starts[r]), int(starts[r + 1])))
    _lib.check(lib.sq_space_set_partition(info._handle, 2, r, starts.ctypes.data_as(C.POINTER(C.c_int64))))
    spaces.append(info)
    lays.append(osa.compile_layout(info, lay))
    t = torch.randn(info.local_len, dtype=torch.float64, device=f"cuda:{r}")
    shards.append(t)
ptrs = (C.c_void_p * 2)(shards[0].data_ptr(), shards[1].data_ptr())
thp = th.ctypes.data_as(C.POINTER(C.c_double))
streams = [torch.cuda.Stream(device=r) for r in range(2)]


def run(r, f, l):
    torch.cuda.set_device(r)
    _lib.check(lib.sq_ups_apply_dist(spaces[r]._handle, lays[r], thp, f, l, 0, ptrs, C.c_void_p(streams[r].cuda_stream)))


def sync():
    for r in range(2):
        torch.cuda.synchronize(r)


nb = spaces[0].num_beta_strings
for label, (f, l) in {"exchange brick (0,1)": (0, 3), "local brick (2,3)": (3, 6), "local brick (14,15)": (21, 24)}.items():
    for who in ([0], [0, 1]):
        run(0, f, l); run(1, f, l); sync()
        best = 1e9
        for _ in range(5):
            sync()
            t0 = time.perf_counter()
            for r in who:
                run(r, f, l)
            sync()
            best = min(best, time.perf_counter() - t0)
        stats = np.zeros(6, dtype=np.int64)
        lib.sq_layout_op_stats(lays[0], f, stats.ctypes.data_as(C.POINTER(C.c_int64)))
        remote_mb = stats[3] * nb / 2 * 8 / 1e6
        print(f"CAS({n},{n}) {label:22s} devices {who}: {best*1e3:7.3f} ms   cross items {stats[3]}  remote read {remote_mb:.0f} MB"
              + (f" -> {remote_mb/best/1e3:.0f} GB/s each way per GPU" if stats[3] else ""), flush=True)
