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
        idx2det=None,
        det2idx=None,
        device: int | None = None,
        row_range: tuple[int, int] | None = None,
        alpha_constraint: tuple[int, int] | None = None,
    ) -> None:
        """``idx2det`` / ``det2idx`` occupy the positions they have in the reference's constructor (ci_spaces.py:12-21).  The engine
        derives both tables from the string lists, so a caller-supplied ``idx2det`` is only checked: it must be the determinant
        list of this product space in the reference's order (anything else -- e.g. an extended space -- needs
        ``get_indexing_extended``); ``det2idx`` is ignored.  ``alpha_constraint = (mask, pattern)`` keeps only the alpha strings
        with ``string & mask == pattern`` (second row layout of a re-sharded vector, ``slowquant_b200.distributed``)."""
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
        if alpha_constraint is not None:
            if row_range is not None:
                raise ValueError("alpha_constraint and row_range are mutually exclusive")
            _lib.check(
                lib.sq_space_create_constrained(
                    num_active_orbs, num_active_elec_alpha, num_active_elec_beta, self.device,
                    int(alpha_constraint[0]), int(alpha_constraint[1]), C.byref(handle),
                )
            )
        else:
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
        if idx2det is not None:
            given = np.asarray(idx2det, dtype=np.int64)
            if given.shape != (self.num_det,) or not np.array_equal(given, self.idx2det):
                raise ValueError(
                    "idx2det is not the determinant list of the (num_active_orbs, n_alpha, n_beta) product space in "
                    "get_indexing order; build extended spaces with get_indexing_extended"
                )

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


def generate_spin_strings(num_orbs: int, num_elec: int):
    """All occupation lists with `num_elec` ones in `num_orbs` places, in the order of the engine's string tables
    (= ``itertools.combinations`` order, ci_spaces.py:56-73).  Host helper; the tables themselves are built inside
    libsqsv (``sq_space_create``) and exported by ``CI_Info.strings``."""
    import itertools

    if num_elec < 0:
        return
    for occupied in itertools.combinations(range(num_orbs), num_elec):
        string = [0] * num_orbs
        for o in occupied:
            string[o] = 1
        yield string


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


# ---- extended spaces (ci_spaces.py:119-341) --------------------------------------------------------------------
def generate_singles(num_inactive_orbs: int, num_virtual_orbs: int):
    """(inactive occupation, virtual occupation) after at most one electron left the inactive orbitals and at most one
    entered the virtual orbitals, in the order of ci_spaces.py:262-294 (hole index outer, particle index inner, "no
    change" last in both loops)."""
    for i in range(num_inactive_orbs + 1):
        inactive = [1] * num_inactive_orbs
        if i != num_inactive_orbs:
            inactive[i] = 0
        for j in range(num_virtual_orbs + 1):
            virtual = [0] * num_virtual_orbs
            if j != num_virtual_orbs:
                virtual[j] = 1
            yield inactive, virtual


def generate_doubles(num_inactive_orbs: int, num_virtual_orbs: int):
    """Same with up to two holes and up to two particles (ci_spaces.py:297-341)."""
    nI, nV = num_inactive_orbs, num_virtual_orbs
    for i in range(nI + 1):
        for i2 in range(min(i + 1, nI), nI + 1):
            inactive = [1] * nI
            for h in (i, i2):
                if h != nI:
                    inactive[h] = 0
            for j in range(nV + 1):
                for j2 in range(min(j + 1, nV), nV + 1):
                    virtual = [0] * nV
                    for q in (j, j2):
                        if q != nV:
                            virtual[q] = 1
                    yield inactive, virtual


def _sector_strings(inactive: list[int], virtual: list[int], num_active_orbs: int, n_active_elec: int, shift: int) -> np.ndarray:
    """All strings (inactive pattern | any active string with n_active_elec electrons | virtual pattern) as int64 masks
    already spread to the interleaved determinant layout (orbital 0 = most significant bit pair; shift 1 = alpha, 0 = beta),
    active part in itertools.combinations order (ci_spaces.py:56-73)."""
    import itertools

    N = len(inactive) + num_active_orbs + len(virtual)
    fixed = 0
    for p, occ in enumerate(inactive):
        if occ:
            fixed |= 1 << (2 * (N - 1 - p) + shift)
    for v, occ in enumerate(virtual):
        if occ:
            fixed |= 1 << (2 * (N - 1 - (len(inactive) + num_active_orbs + v)) + shift)
    out = []
    if n_active_elec < 0:  # nothing to iterate (ci_spaces.py:66-68)
        return np.zeros(0, dtype=np.int64)
    for comb in itertools.combinations(range(num_active_orbs), n_active_elec):
        m = fixed
        for a in comb:
            m |= 1 << (2 * (N - 1 - (len(inactive) + a)) + shift)
        out.append(m)
    return np.asarray(out, dtype=np.int64)


def extended_idx2det(
    num_inactive_orbs: int, num_active_orbs: int, num_virtual_orbs: int, num_active_elec_alpha: int, num_active_elec_beta: int, order: int
) -> np.ndarray:
    """Determinant list of the extended space in the reference's order (ci_spaces.py:142-247): the CAS block, then alpha
    sectors x reference beta, reference alpha x beta sectors and (order 2) single alpha x single beta sectors; alpha outer,
    beta inner inside a block; determinants already present are skipped."""
    if order > 2:
        raise ValueError("Excitation order needs to be <= 2")
    nI, nA, nV = num_inactive_orbs, num_active_orbs, num_virtual_orbs
    singles = list(generate_singles(nI, nV))
    doubles = list(generate_doubles(nI, nV)) if order == 2 else []

    def strings(sector, n_elec, shift):
        inactive, virtual = sector
        return _sector_strings(inactive, virtual, nA, int(n_elec - sum(virtual) + nI - sum(inactive)), shift)

    reference = ([1] * nI, [0] * nV)
    ref_a = strings(reference, num_active_elec_alpha, 1)
    ref_b = strings(reference, num_active_elec_beta, 0)
    blocks = [(ref_a, ref_b)]
    for sector in singles + doubles:
        blocks.append((strings(sector, num_active_elec_alpha, 1), ref_b))
    for sector in singles + doubles:
        blocks.append((ref_a, strings(sector, num_active_elec_beta, 0)))
    if order == 2:
        for sa in singles:
            a_str = strings(sa, num_active_elec_alpha, 1)
            for sb in singles:
                blocks.append((a_str, strings(sb, num_active_elec_beta, 0)))
    dets = np.concatenate([(a[:, None] | b[None, :]).ravel() for a, b in blocks if a.size and b.size])
    _, first = np.unique(dets, return_index=True)
    return dets[np.sort(first)]


class ExtendedCI_Info:
    """CAS + singles(/doubles) into the inactive and virtual orbitals (ci_spaces.py:119-259), attribute for attribute.

    The determinant list is not a product of string lists, but it is a union of (alpha sector) x (beta sector) blocks of the
    product space of ALL orbitals with ``nI + n_alpha`` / ``nI + n_beta`` electrons.  The engine therefore keeps a private
    *parent* product space (every kernel works there unchanged) and this object holds the embedding: vectors of the
    extended space are scattered into parent vectors, operated on, and gathered back, which is exactly the reference's
    "skip what leaves the space" (``do_unsafe=True``) -- and a KeyError otherwise, when anything landed outside.
    Like the reference's version (it scans a Python list per determinant, ci_spaces.py:173), this is for small spaces.
    """

    is_extended = True

    def __init__(self, nI: int, nA: int, nV: int, n_alpha: int, n_beta: int, order: int, device: int | None = None) -> None:
        N = nI + nA + nV
        self.num_inactive_orbs = 0
        self.num_active_orbs = N
        self.num_virtual_orbs = 0
        self.num_active_elec_alpha = n_alpha + nI
        self.num_active_elec_beta = n_beta + nI
        self.space_extension_offset = nI
        self.idx2det = extended_idx2det(nI, nA, nV, n_alpha, n_beta, order)
        self.num_det = int(self.idx2det.size)
        self.local_len = self.num_det
        self.det2idx = {int(d): i for i, d in enumerate(self.idx2det)}
        self.parent = CI_Info(0, N, 0, n_alpha + nI, n_beta + nI, device=device)
        self.parent.space_extension_offset = nI          # ansatz indices are shifted in the parent space too
        self.device = self.parent.device
        self.embedding = self.parent.det2idx.lookup_many(self.idx2det)      # parent index of every determinant
        if np.any(self.embedding < 0):
            raise RuntimeError("extended determinant outside its parent product space")
        self._layouts = self.parent._layouts


def get_indexing_extended(
    num_inactive_orbs: int,
    num_active_orbs: int,
    num_virtual_orbs: int,
    num_active_elec_alpha: int,
    num_active_elec_beta: int,
    order: int,
    device: int | None = None,
) -> ExtendedCI_Info:
    """Same call as ci_spaces.py:119-126 (+ optional ``device``; -1 = host-only integer tables)."""
    return ExtendedCI_Info(
        num_inactive_orbs, num_active_orbs, num_virtual_orbs, num_active_elec_alpha, num_active_elec_beta, order, device
    )
