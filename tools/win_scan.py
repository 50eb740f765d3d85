#!/usr/bin/env python
"""Time sq_ups_apply for several launch-planner configurations in one process (B200).

    python tools/win_scan.py [--cas 16] [--layers 16] [--reps 3] cfg1 cfg2 ...

cfg uses the SQ_WIN syntax ("0" = no window sweeps, "1" = defaults,
"w1:w2:w3,smem_kb,min_suffix,max_bricks,min_bricks").  Prints ms per step, launches, window
sweeps, bricks inside windows, ms per launch and layers/s.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slowquant_b200 import _lib  # noqa: E402
from slowquant_b200.ci_spaces import get_indexing  # noqa: E402
from slowquant_b200.operator_state_algebra import _ups_apply_inplace, compile_layout  # noqa: E402
from slowquant_b200.util import UpsStructure  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cas", type=int, default=16)
ap.add_argument("--layers", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("cfgs", nargs="*", default=["1"])
args = ap.parse_args()

lib = _lib.load()
n, L = args.cas, args.layers
info = get_indexing(0, n, 0, n // 2, n // 2, device=0)
lay = UpsStructure()
lay.create_tiled(n, {"n_layers": L, "do_tups": True})
P = lay.n_params
thetas = np.random.default_rng(1234).uniform(-np.pi, np.pi, P)
handle = compile_layout(info, lay)
state = torch.zeros(info.num_det, dtype=torch.float64, device="cuda:0")
state[0] = 1.0
ref_norm = None
for cfg in args.cfgs:
    _lib.check(lib.sq_set_option(b"win", cfg.encode()))
    stats = (C.c_int64 * 6)()
    _lib.check(lib.sq_layout_plan_stats(handle, 0, P, stats))
    for _ in range(2):
        _ups_apply_inplace(state, info, thetas, lay, 0, P, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        _ups_apply_inplace(state, info, thetas, lay, 0, P, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    s = list(stats)
    print(f"{cfg:40s} ms/step {ms:8.2f}  launches {s[0]:4d} win {s[1]:4d} bricks_in_win {s[2]:4d} quad {s[3]:3d} single {s[4]:3d}"
          f"  ms/launch {ms / max(s[0], 1):6.3f}  layers/s {L / ms * 1e3:8.1f}  norm {float(torch.linalg.norm(state)):.12f}", flush=True)
