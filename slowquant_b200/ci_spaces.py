"""CI-space indexing on the device engine.

Mirror of slowquant/unitary_coupled_cluster/ci_spaces.py: ``get_indexing`` returns a ``CI_Info`` with the
same attributes (``idx2det``, ``det2idx``, ``num_*``, ``space_extension_offset``).  Instead of building a
Python list and a hash map determinant by determinant (ci_spaces.py:93-107) the space is the product of
the alpha and beta string lists held by libsqsv (``sq_space``); ``idx2det`` / ``det2idx`` are derived
views computed by combinatorial ranking, bit-identical to the reference's tables.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from slowquant_b200 import _lib


def _current_device() -> int:
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "slowquant_b200 needs a CUDA device (sm_100a); there is no CPU fallback for the state-vector engine"
        )
    return torch.cuda.current_device()


class _Det2Idx:
    """Mapping view determinant -> index (replaces the numba typed dict of ci_spaces.py:47-52)."""

    def __init__(self, info: "CI_Info") -> None:
        self._info = info

    def _lookup(self, det: int) -> int:
        lib = _lib.load()
        d = np.array([det], dtype=np.int64)
        out = np.empty(1, dtype=np.int64)
        _lib.check(
            lib.sq_space_det2idx(
                self._info._handle, 1, d.ctypes.data_as(C.POINTER(C.c_int64)), out.ctypes.data_as(C.POINTER(C.c_int64))
            )
        )
        return int(out[0])

    def __getitem__(self, det: int) -> int:
        idx = self._lookup(int(det))
        if idx < 0:
            raise KeyError(det)
        return idx

    def __contains__(self, det: int) -> bool:
        return self._lookup(int(det)) >= 0

    def __len__(self) -> int:
        return self._info.num_det

    def lookup_many(self, dets: np.ndarray) -> np.ndarray:
        lib = _lib.load()
        d = np.ascontiguousarray(dets, dtype=np.int64)
        out = np.empty(d.shape, dtype=np.int64)
        _lib.check(
            lib.sq_space_det2idx(
                self._info._handle, d.size, d.ctypes.data_as(C.POINTER(C.c_int64)), out.ctypes.data_as(C.POINTER(C.c_int64))
            )
        )
        return out


class CI_Info:
    """Configuration-space information; attribute names follow ci_spaces.py:9-53."""

    def __init__(
        self,
        num_inactive_orbs: int,
        num_active_orbs: int,
        num_virtual_orbs: int,
        num_active_elec_alpha: int,
        num_active_elec_beta: int,
        device: int | None = None,
        row_range: tuple[int, int] | None = None,
    ) -> None:
        self.num_inactive_orbs = num_inactive_orbs
        self.num_active_orbs = num_active_orbs
        self.num_virtual_orbs = num_virtual_orbs
        self.num_active_elec_alpha = num_active_elec_alpha
        self.num_active_elec_beta = num_active_elec_beta
        self.space_extension_offset = 0
        self.device = _current_device() if device is None else device
        lib = _lib.load()
        handle = C.c_void_p()
        rb, re = (0, -1) if row_range is None else row_range
        _lib.check(
            lib.sq_space_create(
                num_active_orbs, num_active_elec_alpha, num_active_elec_beta, self.device, rb, re, C.byref(handle)
            )
        )
        self._handle = handle
        self.num_det = int(lib.sq_space_num_det(handle))
        self.num_alpha_strings = int(lib.sq_space_num_strings(handle, 0))
        self.num_beta_strings = int(lib.sq_space_num_strings(handle, 1))
        self.local_rows = int(lib.sq_space_local_rows(handle))
        self.row_begin = rb
        self.local_len = self.local_rows * self.num_beta_strings
        self._idx2det: np.ndarray | None = None
        self.det2idx = _Det2Idx(self)
        self._layouts: dict = {}

    @property
    def idx2det(self) -> np.ndarray:
        """int64 determinant of every index (interleaved a0 b0 a1 b1 ..., ci_spaces.py:99-107); built lazily."""
        if self._idx2det is None:
            lib = _lib.load()
            out = np.empty(self.num_det, dtype=np.int64)
            _lib.check(lib.sq_space_export_idx2det(self._handle, 0, self.num_det, out.ctypes.data_as(C.POINTER(C.c_int64))))
            self._idx2det = out
        return self._idx2det

    def strings(self, spin: int) -> np.ndarray:
        """Occupation masks (bit o = orbital o) of the alpha (0) or beta (1) strings."""
        lib = _lib.load()
        n = self.num_beta_strings if spin else self.num_alpha_strings
        out = np.empty(n, dtype=np.uint32)
        _lib.check(lib.sq_space_export_strings(self._handle, spin, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def __del__(self) -> None:
        try:
            lib = _lib.load()
            for lay in getattr(self, "_layouts", {}).values():
                lib.sq_layout_destroy(lay)
            if getattr(self, "_handle", None):
                lib.sq_space_destroy(self._handle)
                self._handle = None
        except Exception:
            pass


def get_indexing(
    num_inactive_orbs: int,
    num_active_orbs: int,
    num_virtual_orbs: int,
    num_active_elec_alpha: int,
    num_active_elec_beta: int,
    device: int | None = None,
    row_range: tuple[int, int] | None = None,
) -> CI_Info:
    """Relation between index and determinant (same call as ci_spaces.py:76-116).

    ``device=-1`` gives a host-only space (integer tables only).  ``row_range`` selects the alpha rows
    resident on this device for sharded vectors.
    """
    return CI_Info(
        num_inactive_orbs,
        num_active_orbs,
        num_virtual_orbs,
        num_active_elec_alpha,
        num_active_elec_beta,
        device=device,
        row_range=row_range,
    )


def get_indexing_extended(*args, **kwargs):
    """Extended (CAS + singles/doubles) spaces of ci_spaces.py:119-259 are not a product of string lists;
    they are outside the round-1 scope."""
    raise NotImplementedError("extended CI spaces are not available in the B200 engine yet")
