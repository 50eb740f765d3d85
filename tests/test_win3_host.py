"""CPU checks of the host logic behind win3_kernel (slowquant_b200/csrc/sqsv_win3.cu): the orbital-triple item lists, the step
grouping, the merged tiles and the 3 x 3 block algebra, through the library's host emulation of the kernel
(sq_debug_win3_emulate, test infrastructure) on host-only spaces, against the oracle's restatement of the reference loop
(operator_state_algebra.py:963-1085)."""
import ctypes as C

import numpy as np
import pytest

from oracle import sq_oracle as orc
from slowquant_b200 import _lib


def _space_and_layout(n, na, nb, types, idx):
    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.sq_space_create(n, na, nb, -1, 0, -1, C.byref(h)))
    codes = np.array([_lib.EXC_CODES[t] for t in types], dtype=np.int32)
    offs = np.zeros(len(codes) + 1, dtype=np.int32)
    flat = []
    for k, ix in enumerate(idx):
        flat.extend(int(x) for x in ix)
        offs[k + 1] = len(flat)
    flat = np.asarray(flat, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    lay = C.c_void_p()
    _lib.check(lib.sq_layout_create(h, len(codes), p(codes), p(offs), p(flat), C.byref(lay)))
    return lib, h, lay


def _emulate(lib, h, lay, th, state, dagger=0):
    out = np.ascontiguousarray(state, dtype=np.float64).copy()
    pd = C.POINTER(C.c_double)
    th = np.ascontiguousarray(th, dtype=np.float64)
    _lib.check(lib.sq_debug_win3_emulate(h, lay, th.ctypes.data_as(pd), 0, len(th), dagger, out.ctypes.data_as(pd)))
    return out


# windows "w1:w2:w3,smem_kb,min_suffix,max_bricks,min_bricks": min_bricks = 1 sends every brick through a window sweep
@pytest.mark.parametrize(
    "n,na,nb,L,qnp,win",
    [
        (6, 3, 3, 2, False, "6:5:4,72,0,16,1"),     # one window = the whole space (top window, column-wise copies)
        (8, 4, 4, 3, False, "6:5:4,72,1,16,1"),     # windows with suffix runs and the top window
        (8, 3, 5, 2, False, "5:4:3,72,1,16,1"),     # unequal spin counts, narrower windows
        (9, 5, 3, 2, True, "6:5:4,72,2,16,1"),      # QNP bricks (pair double + one sa_single)
        (10, 5, 5, 2, False, "6:5:4,72,3,16,1"),    # the default configuration of the planner
        (7, 1, 6, 2, False, "4:3:0,72,1,16,1"),     # nearly empty / nearly full spins: inert classes dominate
    ],
)
def test_win3_emulation_matches_oracle(n, na, nb, L, qnp, win):
    types, idx = orc.tiled_layout(n, L, do_qnp=qnp)
    rng = np.random.default_rng(1000 + 17 * n + na)
    th = rng.uniform(-np.pi, np.pi, len(types))
    sp = orc.get_indexing(0, n, 0, na, nb)
    st = rng.normal(size=sp.num_det)
    st /= np.linalg.norm(st)
    lib, h, lay = _space_and_layout(n, na, nb, types, idx)
    try:
        _lib.check(lib.sq_set_option(b"win", win.encode()))
        stats = (C.c_int64 * 6)()
        _lib.check(lib.sq_layout_plan_stats(lay, 0, len(types), stats))
        assert int(stats[1]) > 0 and int(stats[3]) == 0 and int(stats[4]) == 0, list(stats)
        ref = orc.construct_ups_state(st, sp, th, types, idx, threaded=True)
        res = _emulate(lib, h, lay, th, st)
        assert np.max(np.abs(res - ref)) < 1e-12
        ref_d = orc.construct_ups_state(st, sp, th, types, idx, dagger=True, threaded=True)
        res_d = _emulate(lib, h, lay, th, st, dagger=1)
        assert np.max(np.abs(res_d - ref_d)) < 1e-12
    finally:
        _lib.check(lib.sq_set_option(b"win", b"1"))
        lib.sq_layout_destroy(lay)
        lib.sq_space_destroy(h)
