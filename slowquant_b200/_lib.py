"""ctypes binding of libsqsv.so (C ABI declared in include/sqsv.h).

There is no CPU fallback: if the CUDA library cannot be loaded this module raises, and every compute
entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsqsv.so")

SQ_OK, SQ_ERR_INVALID, SQ_ERR_CUDA, SQ_ERR_OUTSIDE, SQ_ERR_UNSUPPORTED, SQ_ERR_NOMEM = range(6)

EXC_CODES = {
    "sa_single": 0,
    "single": 1,
    "double": 2,
    "triple": 3,
    "quadruple": 4,
    "quintuple": 5,
    "sextuple": 6,
    "sa_double_1": 7,
    "sa_double_2": 8,
    "sa_double_3": 9,
    "sa_double_4": 10,
    "sa_double_5": 11,
}

_lib = None


def _declare(lib: C.CDLL) -> None:
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    pi32, pi64, pdbl, pu32 = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    sig = {
        "sq_last_error": (C.c_char_p, []),
        "sq_version": (i32, []),
        "sq_space_create": (i32, [i32, i32, i32, i32, i64, i64, C.POINTER(vp)]),
        "sq_space_create_constrained": (i32, [i32, i32, i32, i32, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
        "sq_space_destroy": (i32, [vp]),
        "sq_space_num_det": (i64, [vp]),
        "sq_space_num_strings": (i64, [vp, i32]),
        "sq_space_local_rows": (i64, [vp]),
        "sq_space_export_strings": (i32, [vp, i32, pu32]),
        "sq_space_export_idx2det": (i32, [vp, i64, i64, pi64]),
        "sq_space_det2idx": (i32, [vp, i64, pi64, pi64]),
        "sq_layout_create": (i32, [vp, i32, pi32, pi32, pi32, C.POINTER(vp)]),
        "sq_layout_attach_generator": (i32, [vp, i32, i32, pi32, pi32, pdbl]),
        "sq_layout_destroy": (i32, [vp]),
        "sq_layout_num_ops": (i32, [vp]),
        "sq_layout_num_launches": (i32, [vp, i32, i32]),
        "sq_layout_touched_amplitudes": (i64, [vp, i32, i32]),
        "sq_layout_plan_stats": (i32, [vp, i32, i32, pi64]),
        "sq_set_option": (i32, [C.c_char_p, C.c_char_p]),
        "sq_ups_apply": (i32, [vp, vp, pdbl, i32, i32, i32, vp, vp]),
        "sq_ups_apply_batch": (i32, [vp, vp, pdbl, i32, i32, i32, vp, i32, i64, vp]),
        "sq_ups_apply_list": (i32, [vp, vp, pdbl, i32, pi32, i32, i32, vp, vp]),
        "sq_reshard_rows": (i32, [i32, i64, i64, vp, vp, vp, C.POINTER(vp), i32, vp]),
        "sq_layout_op_blocked": (i32, [vp, i32]),
        "sq_layout_plan_stats_list": (i32, [vp, i32, pi32, pi64]),
        "sq_grad_action": (i32, [vp, vp, i32, vp, vp, vp]),
        "sq_ups_grad_sweep": (i32, [vp, vp, pdbl, i32, i32, vp, vp, pdbl, vp]),
        "sq_ups_grad_sweep_list": (i32, [vp, vp, pdbl, i32, pi32, vp, vp, pdbl, vp]),
        "sq_ups_grad_sweep_list_rev": (i32, [vp, vp, pdbl, i32, pi32, vp, vp, pdbl, vp]),
        "sq_ups_energy_grad": (i32, [vp, vp, pdbl, dbl, pdbl, pdbl, vp, vp, vp, pdbl, pdbl, vp]),
        "sq_apply_strings": (i32, [vp, i32, pi32, pi32, pdbl, vp, vp, i32, i32, vp]),
        "sq_dot": (i32, [vp, vp, vp, pdbl, vp]),
        "sq_axpy": (i32, [vp, dbl, vp, vp, vp]),
        "sq_scale_copy": (i32, [vp, dbl, vp, vp, vp]),
        "sq_sigma": (i32, [vp, dbl, pdbl, pdbl, vp, vp, vp]),
        "sq_rdm12": (i32, [vp, vp, vp, pdbl, pdbl, vp]),
        "sq_debug_string_action": (
            i32,
            [vp, pi32, i32, C.c_uint32, C.c_uint32, C.POINTER(i32), pu32, pu32, C.POINTER(i32)],
        ),
        "sq_debug_etab_closed_form": (i32, [vp, C.POINTER(C.c_int)]),
        "sq_debug_win3_emulate": (i32, [vp, vp, pdbl, i32, i32, i32, pdbl]),
        "sq_launch_count": (i64, []),
        "sq_partition_prefix": (i32, [i32, i32, i32, pi64]),
        "sq_space_set_partition": (i32, [vp, i32, i32, pi64]),
        "sq_dist_alloc": (i32, [i32, i64, C.POINTER(vp)]),
        "sq_dist_free": (i32, [vp]),
        "sq_ipc_export": (i32, [vp, C.c_char_p]),
        "sq_ipc_import": (i32, [i32, C.c_char_p, C.POINTER(vp)]),
        "sq_ipc_close": (i32, [vp]),
        "sq_layout_needs_exchange": (i32, [vp, i32, i32]),
        "sq_layout_op_stats": (i32, [vp, i32, pi64]),
        "sq_ups_apply_dist": (i32, [vp, vp, pdbl, i32, i32, i32, C.POINTER(vp), vp]),
        "sq_rdm12_dist": (i32, [vp, C.POINTER(vp), C.POINTER(vp), pdbl, pdbl, vp]),
        "sq_rdm12_dist_sym": (i32, [vp, C.POINTER(vp), C.POINTER(vp), dbl, pdbl, pdbl, vp]),
        "sq_sigma_dist": (i32, [vp, pdbl, pdbl, C.POINTER(vp), C.POINTER(vp), vp]),
        "sq_spinsym_measure_dist": (i32, [vp, C.POINTER(vp), pdbl, vp]),
        "sq_sigma_dist_sym": (i32, [vp, pdbl, pdbl, C.POINTER(vp), C.POINTER(vp), dbl, C.POINTER(C.c_int), vp]),
        "sq_spinsym_mirror_dist": (i32, [vp, C.POINTER(vp), dbl, vp]),
        "sq_ups_grad_sweep_dist": (i32, [vp, vp, pdbl, i32, i32, C.POINTER(vp), C.POINTER(vp), pdbl, vp]),
        "sq_layout_plan_export": (i32, [vp, pdbl, i32, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i32, C.POINTER(C.c_int)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args


EXPORTED_SYMBOLS = (
    "sq_last_error sq_version sq_space_create sq_space_destroy sq_space_num_det sq_space_num_strings "
    "sq_space_local_rows sq_space_export_strings sq_space_export_idx2det sq_space_det2idx sq_layout_create "
    "sq_layout_attach_generator sq_layout_destroy sq_layout_num_ops sq_layout_num_launches "
    "sq_layout_touched_amplitudes sq_layout_plan_stats sq_set_option sq_ups_apply "
    "sq_grad_action sq_ups_grad_sweep sq_ups_energy_grad sq_apply_strings sq_dot sq_axpy sq_scale_copy sq_sigma sq_rdm12 "
    "sq_debug_string_action sq_launch_count sq_partition_prefix sq_space_set_partition sq_dist_alloc sq_dist_free "
    "sq_ipc_export sq_ipc_import sq_ipc_close sq_layout_needs_exchange sq_ups_apply_dist sq_rdm12_dist sq_sigma_dist sq_layout_op_stats sq_layout_plan_export sq_debug_etab_closed_form sq_debug_win3_emulate sq_ups_grad_sweep_dist "
    "sq_space_create_constrained sq_ups_apply_list sq_reshard_rows sq_layout_op_blocked sq_layout_plan_stats_list sq_ups_apply_batch sq_ups_grad_sweep_list sq_ups_grad_sweep_list_rev sq_spinsym_measure_dist sq_sigma_dist_sym sq_spinsym_mirror_dist sq_rdm12_dist_sym"
).split()


def load() -> C.CDLL:
    """Load libsqsv.so, building it with nvcc if it is not there yet.  Raises if that fails."""
    global _lib
    if _lib is not None:
        return _lib
    # rebuild when the library is missing OR older than its sources (digest of the .cu / .h files against the build
    # stamp): the .so is git-ignored and travels with repository snapshots, so a stale binary must never be picked up
    alt = os.environ.get("SQSV_LIB")   # developer switch: A/B of compile-time variants of the library (tools/ab_quad_rows.sh)
    if alt:
        lib = C.CDLL(alt, mode=C.RTLD_GLOBAL)
        _declare(lib)
        _lib = lib
        return lib
    if os.path.exists(os.path.join(_HERE, "csrc", "sqsv_api.cu")) or not os.path.exists(LIB_PATH):
        from slowquant_b200.build import build_library

        build_library()
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    _declare(lib)
    _lib = lib
    return lib


class SqsvError(RuntimeError):
    pass


def check(status: int) -> None:
    """Map a C status to the exception the reference would raise at that point."""
    if status == SQ_OK:
        return
    msg = load().sq_last_error().decode(errors="replace")
    if status == SQ_ERR_INVALID:
        raise ValueError(msg)
    if status == SQ_ERR_OUTSIDE:
        raise KeyError(msg)
    if status == SQ_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status == SQ_ERR_NOMEM:
        raise MemoryError(msg)
    raise SqsvError(msg)
