#!/usr/bin/env python
"""Print the launch plan statistics of a tUPS circuit (host-only: no GPU needed).

    python tools/plan_stats.py [n_orb] [layers]       (SQ_WIN=... selects the window configuration,
                                                       SQ_PLAN_DEBUG=1 lists every launch)
"""
import ctypes as C
import os
import sys
import time


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slowquant_b200 import _lib  # noqa: E402
from slowquant_b200.ci_spaces import get_indexing  # noqa: E402
from slowquant_b200.operator_state_algebra import compile_layout  # noqa: E402
from slowquant_b200.util import UpsStructure  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16
info = get_indexing(0, n, 0, n // 2, n // 2, device=-1)
lay = UpsStructure()
lay.create_tiled(n, {"n_layers": L, "do_tups": True})
h = compile_layout(info, lay)
lib = _lib.load()
out = (C.c_int64 * 6)()
t0 = time.time()
_lib.check(lib.sq_layout_plan_stats(h, 0, lay.n_params, out))
t1 = time.time()
o = list(out)
print(f"CAS({n},{n}) L={L}: launches={o[0]} window sweeps={o[1]} bricks in windows={o[2]} quads={o[3]} singles={o[4]} other={o[5]}")
print(f"launches per layer = {o[0] / L:.2f}; planning {1e3 * (t1 - t0):.1f} ms")
t0 = time.time()
_lib.check(lib.sq_layout_plan_stats(h, 0, lay.n_params, out))
print(f"second planning {1e3 * (time.time() - t0):.1f} ms")
