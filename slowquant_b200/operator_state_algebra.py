"""Operator-state algebra on the B200 engine.

Same function names, argument order and error behaviour as the reference's
slowquant/unitary_coupled_cluster/operator_state_algebra.py; the numerics run in libsqsv's CUDA kernels.

State vectors may be numpy arrays (host; copied to the device and back on every call, exactly the
reference's value semantics) or ``torch.float64`` CUDA tensors (device resident; results are new CUDA
tensors).  Inputs are never modified.
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import numpy as np
import torch

from slowquant_b200 import _lib
from slowquant_b200.ci_spaces import CI_Info, get_indexing
from slowquant_b200.fermionic_operator import FermionicOperator
from slowquant_b200.operators import ActiveSpaceHamiltonian, G2_sa
from slowquant_b200.util import UccStructure, UpsStructure

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)


# ---- plumbing -----------------------------------------------------------------------------------
def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device_of(ci_info: CI_Info) -> torch.device:
    return torch.device("cuda", ci_info.device)


def _to_device(state, ci_info: CI_Info, copy: bool = True) -> tuple[torch.Tensor, bool]:
    """Return (fp64 CUDA tensor holding a private copy of `state`, input_was_numpy)."""
    dev = _device_of(ci_info)
    if isinstance(state, torch.Tensor):
        if state.dtype != torch.float64:
            raise TypeError(f"state tensors must be float64, got {state.dtype}")
        t = state.to(dev)
        if copy and t.data_ptr() == state.data_ptr():
            t = t.clone()
        t = t.contiguous()
        was_numpy = False
    else:
        arr = np.ascontiguousarray(state, dtype=np.float64)
        host = torch.from_numpy(arr)
        # page-locked host arrays (e.g. results of earlier calls) take the DMA fast path
        t = host.to(dev, non_blocking=host.is_pinned())
        was_numpy = True
    if t.numel() != ci_info.local_len:
        raise ValueError(f"state has {t.numel()} elements, the CI space holds {ci_info.local_len}")
    return t, was_numpy


def _from_device(t: torch.Tensor, was_numpy: bool):
    if was_numpy:
        # fresh page-locked array from torch's caching host allocator: D2H at full PCIe rate
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t)
        return host.numpy()
    return t


def _ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


# ---- extended (non-product) spaces: embed into the parent product space, operate, project back ----------------
def _is_extended(ci_info) -> bool:
    return getattr(ci_info, "is_extended", False)


def _embedding(ci_info) -> tuple[torch.Tensor, torch.Tensor]:
    """(parent index of every determinant of the extended space, 0/1 mask over the parent vector), device resident."""
    cached = getattr(ci_info, "_embed_dev", None)
    if cached is None:
        dev = _device_of(ci_info)
        idx = torch.from_numpy(np.ascontiguousarray(ci_info.embedding, dtype=np.int64)).to(dev)
        mask = torch.zeros(ci_info.parent.local_len, dtype=torch.float64, device=dev)
        mask[idx] = 1.0
        cached = ci_info._embed_dev = (idx, mask)
    return cached


def _embed(state, ci_info) -> tuple[torch.Tensor, bool]:
    """Scatter a vector of the extended space into a zero-filled parent vector."""
    sub, was_numpy = _to_device(state, ci_info, copy=False)
    idx, _ = _embedding(ci_info)
    full = torch.zeros(ci_info.parent.local_len, dtype=torch.float64, device=sub.device)
    full[idx] = sub
    return full, was_numpy


def _restrict(full: torch.Tensor, ci_info) -> torch.Tensor:
    return full[_embedding(ci_info)[0]]


def _project(full: torch.Tensor, ci_info, do_unsafe: bool) -> torch.Tensor:
    """Zero whatever left the extended space: the reference's do_unsafe=True skips such determinants string by string
    (osa.py:131-135), which is this projection; without do_unsafe it raises KeyError."""
    _, mask = _embedding(ci_info)
    if not do_unsafe:
        leak = float(torch.max(torch.abs(full * (1.0 - mask))))
        if leak > 0.0:
            raise KeyError("operator maps a determinant outside the CI space (pass do_unsafe=True to skip such terms)")
    return full * mask


def encode_operator(op: FermionicOperator) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Flatten a FermionicOperator into (ops_flat, offsets, coeffs); entry = 2*spin_orbital + dagger."""
    labels = op.operators
    n = len(labels)
    offsets = np.zeros(n + 1, dtype=np.int32)
    flat: list[int] = []
    coeffs = np.empty(n, dtype=np.float64)
    for s, (label, fac) in enumerate(labels.items()):
        for idx, dag in label:
            flat.append(2 * int(idx) + (1 if dag else 0))
        offsets[s + 1] = len(flat)
        coeffs[s] = fac
    ops_flat = np.asarray(flat, dtype=np.int32) if flat else np.zeros(1, dtype=np.int32)
    return ops_flat, offsets, coeffs


def _apply_operator(op: FermionicOperator, src: torch.Tensor, dst: torch.Tensor, ci_info: CI_Info, do_unsafe: bool) -> None:
    """dst <- op|src> (strings applied by the gather kernel; replaces osa.py:596-628)."""
    lib = _lib.load()
    ops_flat, offsets, coeffs = encode_operator(op)
    _lib.check(
        lib.sq_apply_strings(
            ci_info._handle,
            len(coeffs),
            ops_flat.ctypes.data_as(_PI),
            offsets.ctypes.data_as(_PI),
            coeffs.ctypes.data_as(_PD),
            _ptr(src),
            _ptr(dst),
            0,
            1 if do_unsafe else 0,
            _stream(),
        )
    )


_SIGMA_AVAILABLE = True


def _apply_hamiltonian(op: ActiveSpaceHamiltonian, src: torch.Tensor, dst: torch.Tensor, ci_info: CI_Info) -> bool:
    """dst <- H|src> through the dedicated sigma kernel; False if it cannot be used for this operator."""
    global _SIGMA_AVAILABLE
    if not _SIGMA_AVAILABLE:
        return False
    if op._ops is not None:  # operator was modified / materialised: treat as generic strings
        return False
    if op.num_inactive_orbs != ci_info.num_inactive_orbs or op.num_active_orbs != ci_info.num_active_orbs:
        return False
    lib = _lib.load()
    e_core, h_eff, g_act = op.folded_integrals()
    status = lib.sq_sigma(
        ci_info._handle,
        e_core,
        h_eff.ctypes.data_as(_PD),
        g_act.ctypes.data_as(_PD),
        _ptr(src),
        _ptr(dst),
        _stream(),
    )
    if status == _lib.SQ_ERR_UNSUPPORTED:
        _SIGMA_AVAILABLE = False
        return False
    _lib.check(status)
    return True


_SA_CASE = {"sa_double_1": 1, "sa_double_2": 2, "sa_double_3": 3, "sa_double_4": 4, "sa_double_5": 5}


def compile_layout(ci_info: CI_Info, wf_struct: UpsStructure) -> C.c_void_p:
    """Compile (and cache on the CI_Info) the device layout for an ansatz structure."""
    types = wf_struct.excitation_operator_type
    indices = wf_struct.excitation_indices
    key = (id(wf_struct), len(types), hash(tuple(types)), hash(tuple(tuple(int(x) for x in t) for t in indices)))
    lay = ci_info._layouts.get(key)
    if lay is not None:
        return lay
    lib = _lib.load()
    n = len(types)
    codes = np.empty(max(n, 1), dtype=np.int32)
    offsets = np.zeros(n + 1, dtype=np.int32)
    flat: list[int] = []
    off = ci_info.space_extension_offset
    for k, (t, idx) in enumerate(zip(types, indices)):
        if t not in _lib.EXC_CODES:
            raise ValueError(f"Got unknown excitation type, {t}")
        codes[k] = _lib.EXC_CODES[t]
        shift = off if t.startswith("sa_") else 2 * off
        flat.extend(int(x) + shift for x in idx)
        offsets[k + 1] = len(flat)
    flat_arr = np.asarray(flat, dtype=np.int32) if flat else np.zeros(1, dtype=np.int32)
    handle = C.c_void_p()
    _lib.check(
        lib.sq_layout_create(
            ci_info._handle, n, codes.ctypes.data_as(_PI), offsets.ctypes.data_as(_PI), flat_arr.ctypes.data_as(_PI), C.byref(handle)
        )
    )
    # spin-adapted doubles: attach the normal-ordered strings of T = G - G^dagger (osa.py:1063-1065, 1087-1092, ...)
    gen_cache: dict[tuple, tuple] = {}
    for k, (t, idx) in enumerate(zip(types, indices)):
        if t in _SA_CASE:
            i, j, a, b = (int(x) + off for x in idx)
            ck = (i, j, a, b, t)
            if ck not in gen_cache:
                gen_cache[ck] = encode_operator(G2_sa(i, j, a, b, _SA_CASE[t], True))
            ops_flat, op_offsets, coeffs = gen_cache[ck]
            _lib.check(
                lib.sq_layout_attach_generator(
                    handle, k, len(coeffs), ops_flat.ctypes.data_as(_PI), op_offsets.ctypes.data_as(_PI), coeffs.ctypes.data_as(_PD)
                )
            )
    ci_info._layouts[key] = handle
    return handle


def _thetas_array(thetas: Sequence[float], n: int) -> np.ndarray:
    th = np.ascontiguousarray(np.asarray(thetas, dtype=np.float64))
    if th.size != n:
        raise ValueError(f"Expected {n} theta values got {th.size}")
    return th


def _ups_apply_inplace(
    t: torch.Tensor, ci_info: CI_Info, thetas: Sequence[float], ups_struct: UpsStructure, first: int, last: int, dagger: bool
) -> None:
    lib = _lib.load()
    lay = compile_layout(ci_info, ups_struct)
    th = _thetas_array(thetas, len(ups_struct.excitation_operator_type))
    _lib.check(
        lib.sq_ups_apply(ci_info._handle, lay, th.ctypes.data_as(_PD), first, last, 1 if dagger else 0, _ptr(t), _stream())
    )


def _ups_apply_batch_inplace(
    states: torch.Tensor, ci_info: CI_Info, thetas: Sequence[float], ups_struct: UpsStructure, first: int, last: int, dagger: bool
) -> None:
    """The same operators [first, last) on every row of a contiguous device batch ``[S, N_det]`` in ONE launch sequence
    (``sq_ups_apply_batch``: the window / gauge sweeps carry the state index as a batch dimension)."""
    if states.dim() != 2 or not states.is_contiguous() or states.shape[1] != ci_info.local_len:
        raise ValueError("batched application needs a contiguous [S, N_det] device tensor")
    lib = _lib.load()
    lay = compile_layout(ci_info, ups_struct)
    th = _thetas_array(thetas, len(ups_struct.excitation_operator_type))
    _lib.check(
        lib.sq_ups_apply_batch(ci_info._handle, lay, th.ctypes.data_as(_PD), first, last, 1 if dagger else 0, _ptr(states),
                               int(states.shape[0]), int(states.shape[1]), _stream())
    )


# ---- public surface -----------------------------------------------------------------------------
def construct_ups_state(state, ci_info: CI_Info, thetas: Sequence[float], ups_struct: UpsStructure, dagger: bool = False):
    r"""Apply the unitary product :math:`U_N \dots U_0` (or its adjoint) to `state` (osa.py:963-1412)."""
    if _is_extended(ci_info):  # the ansatz acts inside every inactive/virtual sector: the space is closed under U
        full, was_numpy = _embed(state, ci_info)
        _ups_apply_inplace(full, ci_info.parent, thetas, ups_struct, 0, len(ups_struct.excitation_operator_type), dagger)
        return _from_device(_restrict(full, ci_info), was_numpy)
    t, was_numpy = _to_device(state, ci_info)
    _ups_apply_inplace(t, ci_info, thetas, ups_struct, 0, len(ups_struct.excitation_operator_type), dagger)
    return _from_device(t, was_numpy)


# ---- light cone of a reference determinant -------------------------------------------------------
# An ansatz operator whose spatial orbitals are ALL doubly occupied, or ALL empty, in every determinant of the state is the
# identity on it (T|D> = 0: nothing to excite from / into; the reference's loop skips the zero amplitudes, osa.py:125).  Starting
# from ONE determinant the first operators of a brick-wall circuit therefore touch only a growing window [lo, hi) of orbitals
# around the Fermi level: the state lives in the CAS(hi - lo) space of that window (lower orbitals doubly occupied -- an even
# number of electrons in front of every active spin orbital, so no sign changes --, higher ones empty).  The head of the circuit
# runs in that small space with the same kernels, the result is embedded into the full vector, and the rest of the circuit runs
# as usual: tUPS on CAS(16,16) does its first 4 of 16 layers on 11.8 M instead of 165.6 M determinants.
def _spatial_orbitals(kind: str, idx) -> tuple[int, ...]:
    return tuple(int(x) for x in idx) if kind.startswith("sa_") else tuple(int(x) // 2 for x in idx)


def _light_cone_plan(ci_info: CI_Info, ups_struct: UpsStructure, mask_a: int, mask_b: int, max_fraction: float = 0.125):
    """Split the circuit for the reference determinant (alpha / beta occupation masks): returns None (no gain) or a dict with the
    window space, the operators of the head that act in it, the index of the first operator of the tail and the embedding."""
    import math

    n = ci_info.num_active_orbs
    na, nb = ci_info.num_active_elec_alpha, ci_info.num_active_elec_beta
    types, indices = ups_struct.excitation_operator_type, ups_struct.excitation_indices
    occ = {o for o in range(n) if (mask_a >> o) & 1 and (mask_b >> o) & 1}
    emp = {o for o in range(n) if not (mask_a >> o) & 1 and not (mask_b >> o) & 1}
    active = set(range(n)) - occ - emp

    def window(act):
        if not act:
            return None
        lo, hi = min(act), max(act) + 1
        if any(o not in occ for o in range(lo) if o not in act) or any(o not in emp for o in range(hi, n)):
            return None
        if any(o not in act and o not in occ and o not in emp for o in range(lo, hi)):
            return None
        return lo, hi

    def size(lo, hi):
        return math.comb(hi - lo, na - lo) * math.comb(hi - lo, nb - lo) if 0 <= na - lo <= hi - lo and 0 <= nb - lo <= hi - lo else 0

    head: list[int] = []
    k0, best = 0, None
    for k, (t, idx) in enumerate(zip(types, indices)):
        orbs = set(_spatial_orbitals(t, idx))
        if orbs <= occ or orbs <= emp:
            continue                       # identity on this state
        grown = active | orbs
        w = window(grown)
        if w is None or size(*w) == 0 or size(*w) > max_fraction * ci_info.num_det:
            k0 = k
            break
        active = grown
        occ -= orbs
        emp -= orbs
        head.append(k)
        best = w
    else:
        k0 = len(types)                    # the whole circuit stays inside the window
    if best is None or len(head) < 2:
        return None
    lo, hi = best
    sub = get_indexing(0, hi - lo, 0, na - lo, nb - lo, device=ci_info.device)
    sub_struct = UpsStructure()
    for k in head:
        t, idx = types[k], indices[k]
        shift = lo if t.startswith("sa_") else 2 * lo
        sub_struct._push(t, tuple(int(x) - shift for x in idx), None)
    # embedding: window string w -> full string (doubly occupied below lo) | w << lo, ranks from the full space's string tables
    low = (1 << lo) - 1

    def ranks(spin: int) -> np.ndarray:
        full = ci_info.strings(spin).astype(np.int64)
        order = np.argsort(full)
        want = (sub.strings(spin).astype(np.int64) << lo) | low
        pos = np.searchsorted(full[order], want)
        if np.any(pos >= full.size) or np.any(full[order][np.minimum(pos, full.size - 1)] != want):
            raise RuntimeError("light cone: a window string has no counterpart in the full space")
        return order[pos].astype(np.int64)

    ia, ib = ranks(0), ranks(1)
    embed = (ia.reshape(-1, 1) * ci_info.num_beta_strings + ib.reshape(1, -1)).reshape(-1)   # full index of every window determinant
    wmask = (1 << (hi - lo)) - 1
    sa, sb = (mask_a >> lo) & wmask, (mask_b >> lo) & wmask
    ref = int(np.nonzero(sub.strings(0) == sa)[0][0]) * sub.num_beta_strings + int(np.nonzero(sub.strings(1) == sb)[0][0])
    return {"sub": sub, "struct": sub_struct, "head": np.asarray(head, dtype=np.int64), "k0": k0, "embed": embed, "ref": ref,
            "window": (lo, hi)}


def construct_ups_state_from_determinant(det_index: int, ci_info: CI_Info, thetas: Sequence[float], ups_struct: UpsStructure,
                                         light_cone: bool = True) -> torch.Tensor:
    r"""`construct_ups_state` (osa.py:963-1412) for the reference determinant number `det_index`, returned as a device tensor.
    With ``light_cone`` the head of the circuit runs in the orbital window it can reach (see above); the result is the same
    vector (the dropped operators are identities on every intermediate state)."""
    if _is_extended(ci_info) or ci_info.local_len != ci_info.num_det:
        raise ValueError("construct_ups_state_from_determinant needs a full product space on one device")
    P = len(ups_struct.excitation_operator_type)
    th = _thetas_array(thetas, P)
    dev = _device_of(ci_info)
    plan = None
    if light_cone:
        ia, ib = divmod(int(det_index), ci_info.num_beta_strings)
        cache = ci_info.__dict__.setdefault("_light_cone", {})
        key = (id(ups_struct), P, hash(tuple(ups_struct.excitation_operator_type)),
               hash(tuple(tuple(int(x) for x in t) for t in ups_struct.excitation_indices)), int(det_index))
        if key not in cache:
            plan = _light_cone_plan(ci_info, ups_struct, int(ci_info.strings(0)[ia]), int(ci_info.strings(1)[ib]))
            if plan is not None:
                plan["embed"] = torch.from_numpy(plan["embed"]).to(dev)
            cache[key] = plan
        plan = cache[key]
    full = torch.zeros(ci_info.num_det, dtype=torch.float64, device=dev)
    if plan is None:
        full[int(det_index)] = 1.0
        _ups_apply_inplace(full, ci_info, th, ups_struct, 0, P, False)
        return full
    sub = torch.zeros(plan["sub"].num_det, dtype=torch.float64, device=dev)
    sub[plan["ref"]] = 1.0
    _ups_apply_inplace(sub, plan["sub"], th[plan["head"]], plan["struct"], 0, len(plan["head"]), False)
    full.index_copy_(0, plan["embed"], sub)
    if plan["k0"] < P:
        _ups_apply_inplace(full, ci_info, th, ups_struct, plan["k0"], P, False)
    return full


def propagate_unitary(state, idx: int, ci_info: CI_Info, thetas: Sequence[float], ups_struct: UpsStructure):
    """Apply the single unitary number `idx` of the layout (osa.py:1867-2309)."""
    n = len(ups_struct.excitation_operator_type)
    if not 0 <= idx < n:
        raise IndexError(f"unitary index {idx} out of range for {n} operators")
    if _is_extended(ci_info):
        full, was_numpy = _embed(state, ci_info)
        _ups_apply_inplace(full, ci_info.parent, thetas, ups_struct, idx, idx + 1, False)
        return _from_device(_restrict(full, ci_info), was_numpy)
    t, was_numpy = _to_device(state, ci_info)
    _ups_apply_inplace(t, ci_info, thetas, ups_struct, idx, idx + 1, False)
    return _from_device(t, was_numpy)


def get_grad_action(state, idx: int, ci_info: CI_Info, ups_struct: UpsStructure):
    r"""Apply the generator :math:`T_{idx}` to `state` (osa.py:2757-2865)."""
    n = len(ups_struct.excitation_operator_type)
    if not 0 <= idx < n:
        raise IndexError(f"operator index {idx} out of range for {n} operators")
    if _is_extended(ci_info):
        full, was_numpy = _embed(state, ci_info)
        return _from_device(_restrict(get_grad_action(full, idx, ci_info.parent, ups_struct), ci_info), was_numpy)
    lib = _lib.load()
    lay = compile_layout(ci_info, ups_struct)
    t, was_numpy = _to_device(state, ci_info, copy=False)
    out = torch.empty_like(t)
    _lib.check(lib.sq_grad_action(ci_info._handle, lay, idx, _ptr(t), _ptr(out), _stream()))
    return _from_device(out, was_numpy)


def construct_ucc_state(state, ci_info: CI_Info, thetas: Sequence[float], ucc_struct: UccStructure, dagger: bool = False):
    """exp(T - T^dagger)|state> for the non-factorised UCC (osa.py:870-896), matrix free."""
    from slowquant_b200.ucc_state import expm_multiply_operator, get_ucc_T

    if _is_extended(ci_info):
        full, was_numpy = _embed(state, ci_info)
        return _from_device(_restrict(construct_ucc_state(full, ci_info.parent, thetas, ucc_struct, dagger), ci_info), was_numpy)
    T = get_ucc_T(thetas, ucc_struct, ci_info.space_extension_offset)
    t, was_numpy = _to_device(state, ci_info, copy=False)
    out = expm_multiply_operator(T, t, ci_info, -1.0 if dagger else 1.0)
    return _from_device(out, was_numpy)


def propagate_state(
    operators: list[FermionicOperator | str],
    state,
    ci_info: CI_Info,
    thetas: Sequence[float] | None = None,
    wf_struct: UpsStructure | UccStructure | None = None,
    do_folding: bool = True,
    do_unsafe: bool = False,
):
    r"""Apply `operators` right to left to `state` (osa.py:472-630).

    Elements are FermionicOperators (folded to the active space unless ``do_folding=False``) or the strings
    ``"U"`` / ``"Ud"`` for the ansatz unitary and its adjoint.  Returns a new state.
    """
    if len(operators) == 0:
        return np.copy(state) if not isinstance(state, torch.Tensor) else state.clone()
    if _is_extended(ci_info):
        cur, was_numpy = _embed(state, ci_info)
        for op in operators[::-1]:
            cur = propagate_state([op], cur, ci_info.parent, thetas, wf_struct, do_folding=do_folding, do_unsafe=do_unsafe)
            if not isinstance(op, str):
                cur = _project(cur, ci_info, do_unsafe)
        return _from_device(_restrict(cur, ci_info), was_numpy)
    cur, was_numpy = _to_device(state, ci_info)
    tmp = None
    for op in operators[::-1]:
        if isinstance(op, str):
            if op not in ("U", "Ud"):
                raise ValueError(f"Unknown str operator, expected ('U', 'Ud') got {op}")
            dagger = op == "Ud"
            if isinstance(wf_struct, UpsStructure) or (
                not isinstance(wf_struct, UccStructure) and hasattr(wf_struct, "grad_param_R")
            ):
                if thetas is None:
                    raise ValueError("theta must be different from None")
                _ups_apply_inplace(cur, ci_info, thetas, wf_struct, 0, len(wf_struct.excitation_operator_type), dagger)
            elif isinstance(wf_struct, UccStructure):
                if thetas is None:
                    raise ValueError("theta must be different from None")
                cur = construct_ucc_state(cur, ci_info, thetas, wf_struct, dagger=dagger)
            else:
                raise TypeError(f"Got unknown wave function structure type, {type(wf_struct)}")
        else:
            if tmp is None:
                tmp = torch.empty_like(cur)
            done = False
            if do_folding and isinstance(op, ActiveSpaceHamiltonian):
                done = _apply_hamiltonian(op, cur, tmp, ci_info)
            if not done:
                if do_folding:
                    op_folded = op.get_folded_operator(
                        ci_info.num_inactive_orbs, ci_info.num_active_orbs, ci_info.num_virtual_orbs
                    )
                else:
                    op_folded = op
                _apply_operator(op_folded, cur, tmp, ci_info, do_unsafe)
            cur, tmp = tmp, cur
    return _from_device(cur, was_numpy)


def _dot(a: torch.Tensor, b: torch.Tensor, ci_info: CI_Info) -> float:
    if _is_extended(ci_info):
        return float(torch.dot(a, b))
    lib = _lib.load()
    out = C.c_double(0.0)
    _lib.check(lib.sq_dot(ci_info._handle, _ptr(a), _ptr(b), C.byref(out), _stream()))
    return float(out.value)


def expectation_value(
    bra,
    operators: list[FermionicOperator | str],
    ket,
    ci_info: CI_Info,
    thetas: Sequence[float] | None = None,
    wf_struct: UpsStructure | UccStructure | None = None,
    do_folding: bool = True,
    do_unsafe: bool = False,
) -> float:
    """<bra| operators |ket> as a Python float (osa.py:784-824)."""
    ket_t, _ = _to_device(ket, ci_info, copy=False)
    op_ket = propagate_state(operators, ket_t, ci_info, thetas, wf_struct, do_folding=do_folding, do_unsafe=do_unsafe)
    bra_t, _ = _to_device(bra, ci_info, copy=False)
    val = _dot(bra_t, op_ket, ci_info)
    if not isinstance(val, float):
        raise ValueError(f"Calculated expectation value is not a float, got type {type(val)}")
    return val


def ups_gradient_sweep(
    bra, ket, ci_info: CI_Info, thetas: Sequence[float], ups_struct: UpsStructure
) -> tuple[np.ndarray, object, object]:
    r"""Fused reverse sweep of ups_wavefunction.py:1114-1138.

    For every operator k (ascending): ``g[k] = 2 <bra|T_k|ket>``, then ``bra <- U_k bra``, ``ket <- U_k ket``.
    Returns (g, bra_final, ket_final).
    """
    lib = _lib.load()
    lay = compile_layout(ci_info, ups_struct)
    n = len(ups_struct.excitation_operator_type)
    th = _thetas_array(thetas, n)
    b, b_np = _to_device(bra, ci_info)
    k, k_np = _to_device(ket, ci_info)
    g = np.zeros(n, dtype=np.float64)
    _lib.check(
        lib.sq_ups_grad_sweep(ci_info._handle, lay, th.ctypes.data_as(_PD), 0, n, _ptr(b), _ptr(k), g.ctypes.data_as(_PD), _stream())
    )
    return g, _from_device(b, b_np), _from_device(k, k_np)


def ups_gradient_sweep_backward(bra, ket, ci_info: CI_Info, thetas: Sequence[float], ups_struct: UpsStructure):
    r"""The gradient loop of ups_wavefunction.py:1114-1138 run BACKWARDS through the circuit: started from
    (bra, ket) = (H|psi>, |psi>) with |psi> = U(theta)|ref>, for k = P-1 ... 0: g_k = 2 <bra|T_k|ket>, then both vectors
    <- U_k^dagger (``sq_ups_grad_sweep_list_rev``).  T_k commutes with its own rotation, so these are the numbers of the
    reference's forward loop from (U^dagger H|psi>, |ref>) without the adjoint pass.  Returns (gradient, bra, ket) with the two
    vectors rotated back to (U^dagger H|psi>, |ref>); the inputs are not modified."""
    lib = _lib.load()
    lay = compile_layout(ci_info, ups_struct)
    n = len(ups_struct.excitation_operator_type)
    th = _thetas_array(thetas, n)
    b, b_np = _to_device(bra, ci_info)
    k, k_np = _to_device(ket, ci_info)
    ops = np.arange(n - 1, -1, -1, dtype=np.int32)
    out = np.zeros(max(n, 1), dtype=np.float64)
    _lib.check(
        lib.sq_ups_grad_sweep_list_rev(ci_info._handle, lay, th.ctypes.data_as(_PD), n, ops.ctypes.data_as(_PI), _ptr(b), _ptr(k),
                                       out.ctypes.data_as(_PD), _stream())
    )
    g = np.zeros(n, dtype=np.float64)
    g[ops] = out[:n]
    return g, _from_device(b, b_np), _from_device(k, k_np)


def ups_energy_and_gradient(
    ref_state, ci_info: CI_Info, thetas: Sequence[float], ups_struct: UpsStructure, hamiltonian: ActiveSpaceHamiltonian,
    want_gradient: bool = True,
) -> tuple[float, np.ndarray | None]:
    r"""Energy and analytic theta gradient of :math:`U(\theta)|\text{ref}\rangle` in ONE library call
    (``sq_ups_energy_grad``: state construction, sigma build and the fused gradient sweep of ups_wavefunction.py:1019-1142,
    run backwards through the circuit from (H|psi>, |psi>) so that no adjoint pass is needed, stay on the device; only the scalar
    and the gradient come back)."""
    lib = _lib.load()
    lay = compile_layout(ci_info, ups_struct)
    n = len(ups_struct.excitation_operator_type)
    th = _thetas_array(thetas, n)
    ref, _ = _to_device(ref_state, ci_info, copy=False)
    e_core, h_eff, g_act = hamiltonian.folded_integrals()
    ket, bra = torch.empty_like(ref), torch.empty_like(ref)
    energy = C.c_double(0.0)
    grad = np.zeros(n, dtype=np.float64) if want_gradient else None
    _lib.check(
        lib.sq_ups_energy_grad(
            ci_info._handle, lay, th.ctypes.data_as(_PD), e_core, h_eff.ctypes.data_as(_PD), g_act.ctypes.data_as(_PD),
            _ptr(ref), _ptr(ket), _ptr(bra), C.byref(energy), grad.ctypes.data_as(_PD) if want_gradient else None, _stream(),
        )
    )
    return float(energy.value), grad


def reduced_density_matrices(bra, ket, ci_info: CI_Info, want_rdm2: bool = True):
    r"""Active-space (transition) RDMs in one pass over the vector (replaces the n^4/4 expectation_value
    calls of ups_wavefunction.py:409-476):

        rdm1[p,q] = <bra|E_pq|ket>,   rdm2[p,q,r,s] = <bra|E_pq E_rs|ket> - delta_qr rdm1[p,s].

    Returns (rdm1, rdm2 or None) as numpy arrays.
    """
    lib = _lib.load()
    n = ci_info.num_active_orbs
    k, _ = _to_device(ket, ci_info, copy=False)
    if bra is ket:
        b = k
    else:
        b, _ = _to_device(bra, ci_info, copy=False)
    d1 = np.zeros((n, n), dtype=np.float64)
    d2 = np.zeros((n, n, n, n), dtype=np.float64) if want_rdm2 else None
    _lib.check(
        lib.sq_rdm12(
            ci_info._handle, _ptr(b), _ptr(k), d1.ctypes.data_as(_PD), d2.ctypes.data_as(_PD) if want_rdm2 else None, _stream()
        )
    )
    return d1, d2


def higher_reduced_density_matrices(state, ci_info: CI_Info, rdm1: np.ndarray, rdm2: np.ndarray, want_rdm4: bool = True):
    r"""Active-space 3- and 4-RDMs of one state (replaces the n^6 / n^8 ``expectation_value`` loops of
    ups_wavefunction.py:478-754).

    One panel ``W[(a,b),(c,d)] = E_ab E_cd |psi>`` (n^2 + n^4 gather launches) carries both: with ``D[(a,b)] = E_ab|psi>``

        <E_pq E_rs E_tu>      = D[(q,p)] . W[(r,s),(t,u)]                 (one GEMM  n^2 x N_det x n^4)
        <E_pq E_rs E_tu E_mn> = W[(s,r),(q,p)] . W[(t,u),(m,n)]           (one Gram matrix  n^4 x N_det x n^4)

    followed by the contraction terms of the reference (:517-524, :575-606) as index-placement einsums with the
    (mirrored) rdm1 / rdm2.  Small active spaces only: the panel holds n^4 vectors.
    """
    from slowquant_b200.operators import Epq

    n = ci_info.num_active_orbs
    psi, _ = _to_device(state, ci_info, copy=False)
    ndet = psi.numel()
    if 8.0 * n**4 * ndet > 16e9:
        raise MemoryError(f"rdm3/rdm4 panel needs {8.0 * n**4 * ndet / 1e9:.1f} GB; implemented for small active spaces only")
    E = [[Epq(a, b) for b in range(n)] for a in range(n)]
    D = torch.empty((n * n, ndet), dtype=torch.float64, device=psi.device)
    for a in range(n):
        for b in range(n):
            _apply_operator(E[a][b], psi, D[a * n + b], ci_info, False)
    W = torch.empty((n * n, n * n, ndet), dtype=torch.float64, device=psi.device)
    for a in range(n):
        for b in range(n):
            for cd in range(n * n):
                _apply_operator(E[a][b], D[cd], W[a * n + b, cd], ci_info, False)
    Wf = W.reshape(n**4, ndet)
    # R3[(q,p),(r,s),(t,u)] -> [p,q,r,s,t,u]
    R3 = (D @ Wf.T).cpu().numpy().reshape(n, n, n, n, n, n).transpose(1, 0, 2, 3, 4, 5)
    I = np.eye(n)
    G1, G2 = np.asarray(rdm1), np.asarray(rdm2)
    rdm3 = (
        R3
        - np.einsum("ts,pqru->pqrstu", I, G2)
        - np.einsum("rq,pstu->pqrstu", I, G2)
        - np.einsum("tq,purs->pqrstu", I, G2)
        - np.einsum("ts,rq,pu->pqrstu", I, I, G1)
    )
    if not want_rdm4:
        return rdm3, None
    # R4[(s,r),(q,p),(t,u),(m,n)] -> [p,q,r,s,t,u,m,n]
    R4 = (Wf @ Wf.T).cpu().numpy().reshape(n, n, n, n, n, n, n, n).transpose(3, 2, 1, 0, 4, 5, 6, 7)
    rdm4 = (
        R4
        - np.einsum("rq,pstumn->pqrstumn", I, rdm3)
        - np.einsum("tq,pursmn->pqrstumn", I, rdm3)
        - np.einsum("mq,pnrstu->pqrstumn", I, rdm3)
        - np.einsum("mu,pqrstn->pqrstumn", I, rdm3)
        - np.einsum("ts,pqrumn->pqrstumn", I, rdm3)
        - np.einsum("ms,pqrntu->pqrstumn", I, rdm3)
        - np.einsum("mu,rq,pstn->pqrstumn", I, I, G2)
        - np.einsum("mu,tq,pnrs->pqrstumn", I, I, G2)
        - np.einsum("ts,mu,pqrn->pqrstumn", I, I, G2)
        - np.einsum("ts,rq,pumn->pqrstumn", I, I, G2)
        - np.einsum("ts,mq,pnru->pqrstumn", I, I, G2)
        - np.einsum("ms,rq,pntu->pqrstumn", I, I, G2)
        - np.einsum("ms,tq,purn->pqrstumn", I, I, G2)
        - np.einsum("mu,ts,rq,pn->pqrstumn", I, I, I, G1)
    )
    return rdm3, rdm4


# ---- state-averaged twins: batches [n_states, N_det] (osa.py:633-781, 827-867, 1415-1864, 2312-2976) ----
def _map_states(fn, states):
    if isinstance(states, torch.Tensor):
        return torch.stack([fn(s) for s in states])
    return np.array([fn(s) for s in states])


def propagate_state_SA(operators, state, ci_info, thetas=None, wf_struct=None, do_folding=True, do_unsafe=False):
    return _map_states(lambda s: propagate_state(operators, s, ci_info, thetas, wf_struct, do_folding, do_unsafe), state)


def expectation_value_SA(bra, operators, ket, ci_info, thetas=None, wf_struct=None, do_folding=True) -> float:
    val = 0.0
    for b, k in zip(bra, ket):
        val += expectation_value(b, operators, k, ci_info, thetas, wf_struct, do_folding=do_folding)
    return val / len(bra)


def _pipelined_host_batch(states, ci_info: CI_Info, run_inplace) -> np.ndarray:
    """``out[k] = run_inplace(states[k])`` for a batch of HOST vectors ``[S, N_det]``.

    Three device buffers rotate through three streams: while the kernels of state k run on the caller's stream, the
    H2D copy of state k+1 and the D2H copy of state k-1 are in flight (PCIe is full duplex), so a batch costs
    max(copy, compute) per state instead of their sum.  Copies overlap only for page-locked input; pageable input is
    staged by the driver and merely loses the overlap.  The result is a fresh page-locked array.
    """
    dev = _device_of(ci_info)
    src = torch.from_numpy(np.ascontiguousarray(states, dtype=np.float64))
    S, N = src.shape
    if N != ci_info.local_len:
        raise ValueError(f"states have {N} elements, the CI space holds {ci_info.local_len}")
    out = torch.empty((S, N), dtype=torch.float64, pin_memory=True)
    nb = min(S, 3)
    bufs = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(nb)]
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(nb)]
    ev_done = [torch.cuda.Event() for _ in range(nb)]
    ev_out = [torch.cuda.Event() for _ in range(nb)]
    s_in.wait_stream(main)
    for k in range(S):
        b = k % nb
        with torch.cuda.stream(s_in):
            if k >= nb:
                s_in.wait_event(ev_out[b])      # the buffer is free once its previous result has left
            bufs[b].copy_(src[k], non_blocking=True)
            ev_in[b].record(s_in)
        main.wait_event(ev_in[b])
        run_inplace(bufs[b])                    # launches on the caller's (current) stream
        ev_done[b].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[b])
            out[k].copy_(bufs[b], non_blocking=True)
            ev_out[b].record(s_out)
    s_out.synchronize()
    for buf in bufs:                            # the side streams used the buffers: tell the caching allocator
        buf.record_stream(s_in)
        buf.record_stream(s_out)
    return out.numpy()


def construct_ups_state_SA(state, ci_info, thetas, ups_struct, dagger=False):
    """Batch twin of construct_ups_state (osa.py:1415-1864); host batches are stream-pipelined."""
    n = len(ups_struct.excitation_operator_type)
    if not isinstance(state, torch.Tensor) and len(state) > 1:
        return _pipelined_host_batch(state, ci_info, lambda t: _ups_apply_inplace(t, ci_info, thetas, ups_struct, 0, n, dagger))
    if isinstance(state, torch.Tensor) and state.is_cuda and state.dim() == 2 and not _is_extended(ci_info):
        out = state.to(torch.float64).contiguous().clone()      # a fresh [S, N_det] batch; the input is not modified
        _ups_apply_batch_inplace(out, ci_info, thetas, ups_struct, 0, n, dagger)
        return out
    return _map_states(lambda s: construct_ups_state(s, ci_info, thetas, ups_struct, dagger), state)


def propagate_unitary_SA(state, idx, ci_info, thetas, ups_struct):
    """Batch twin of propagate_unitary (osa.py:2312-2754); host batches are stream-pipelined."""
    n = len(ups_struct.excitation_operator_type)
    if not 0 <= idx < n:
        raise IndexError(f"unitary index {idx} out of range for {n} operators")
    if not isinstance(state, torch.Tensor) and len(state) > 1:
        return _pipelined_host_batch(state, ci_info, lambda t: _ups_apply_inplace(t, ci_info, thetas, ups_struct, idx, idx + 1, False))
    if isinstance(state, torch.Tensor) and state.is_cuda and state.dim() == 2 and not _is_extended(ci_info):
        out = state.to(torch.float64).contiguous().clone()
        _ups_apply_batch_inplace(out, ci_info, thetas, ups_struct, idx, idx + 1, False)
        return out
    return _map_states(lambda s: propagate_unitary(s, idx, ci_info, thetas, ups_struct), state)


def get_grad_action_SA(state, idx, ci_info, ups_struct):
    return _map_states(lambda s: get_grad_action(s, idx, ci_info, ups_struct), state)


def build_operator_matrix(op: FermionicOperator, ci_info: CI_Info, do_unsafe: bool = False) -> np.ndarray:
    """Dense matrix of an (unfolded) operator, column j = op|j> (osa.py:413-469).  Small spaces only."""
    n = ci_info.num_det
    dev = _device_of(ci_info)
    mat = np.zeros((n, n), dtype=np.float64)
    if _is_extended(ci_info):
        for j in range(n):
            unit = torch.zeros(n, dtype=torch.float64, device=dev)
            unit[j] = 1.0
            mat[:, j] = propagate_state([op], unit, ci_info, do_folding=False, do_unsafe=do_unsafe).cpu().numpy()
        return mat
    unit = torch.zeros(n, dtype=torch.float64, device=dev)
    out = torch.empty_like(unit)
    for j in range(n):
        unit.zero_()
        unit[j] = 1.0
        _apply_operator(op, unit, out, ci_info, do_unsafe)
        mat[:, j] = out.cpu().numpy()
    return mat


# ---- the reference's per-string entry points (osa.py:33-410) ----------------------------------------
# The reference's propagate_state hands every ladder string to a numba kernel together with its idx2det / det2idx
# tables.  On the engine the per-string loop lives inside ``sq_apply_strings``; these wrappers keep the reference's
# names, argument order and in-place accumulation for code that calls the kernels directly.  The CI space is
# recovered from the ``det2idx`` view (``ci_info.det2idx``), the only table the engine needs.
def bitcount(x: int) -> int:
    """Number of set bits (osa.py:33-50; the kernels use ``__popc``)."""
    return int(x).bit_count() if x > 0 else 0


def get_ucc_T(thetas: Sequence[float], ucc_struct: UccStructure, offset: int = 0) -> FermionicOperator:
    """T = sum_k theta_k (G_k - G_k^dagger) (osa.py:899-960)."""
    from slowquant_b200.ucc_state import get_ucc_T as _get_ucc_T

    return _get_ucc_T(thetas, ucc_struct, offset)


def _space_of_tables(det2idx) -> CI_Info:
    info = getattr(det2idx, "_info", None)
    if info is None:
        raise TypeError(
            "det2idx must be the det2idx view of a slowquant_b200 CI_Info (the engine addresses determinants by "
            f"string rank, not through a hash map); got {type(det2idx)}"
        )
    return info


def _string_operator(kind: str, a_string, n_first: int, screen, factor: float) -> FermionicOperator:
    """One-string FermionicOperator from the index arrays the reference's kernels take.

    serial kernels (osa.py:610-612):   a_string = anni_idx + create_idx, screen = creators not also annihilated
    threaded kernels (osa.py:577-579): a_string = create_idx + anni_idx, screen = annihilators not also created
    Both split a label into its creators and annihilators in label order, so the label (creators, annihilators)
    reproduces exactly these arrays."""
    a_string = [int(x) for x in np.asarray(a_string).ravel()]
    if kind == "serial":
        anni, create = a_string[:n_first], a_string[n_first:]
        expect = [c for c in create if c not in anni]
    else:
        create, anni = a_string[:n_first], a_string[n_first:]
        expect = [a for a in anni if a not in create]
    if sorted(int(x) for x in np.asarray(screen).ravel()) != sorted(expect):
        raise ValueError("screening indices do not belong to a_string (expected the arrays propagate_state builds)")
    label = tuple((c, True) for c in create) + tuple((a, False) for a in anni)
    return FermionicOperator({label: float(factor)})


def _accumulate_string(op: FermionicOperator, state, det2idx, do_unsafe: bool, tmp_state, batched: bool):
    ci_info = _space_of_tables(det2idx)
    fn = propagate_state_SA if batched else propagate_state
    out = fn([op], state, ci_info, do_folding=False, do_unsafe=bool(do_unsafe))
    if isinstance(tmp_state, torch.Tensor):
        tmp_state += out if isinstance(out, torch.Tensor) else torch.from_numpy(np.asarray(out)).to(tmp_state.device)
    else:
        tmp_state += out.cpu().numpy() if isinstance(out, torch.Tensor) else out
    return tmp_state


def apply_operator_serial(
    state, a_string, create_screen, anni_idx, num_active_orbs, parity_check, idx2det, det2idx, do_unsafe, tmp_state, factor
):
    """tmp_state += factor * string|state> for one ladder string (osa.py:53-136).  ``parity_check`` / ``idx2det`` /
    ``num_active_orbs`` are accepted for signature compatibility; signs come from the closed form of DESIGN.md §2."""
    op = _string_operator("serial", a_string, len(np.asarray(anni_idx).ravel()), create_screen, factor)
    return _accumulate_string(op, state, det2idx, do_unsafe, tmp_state, batched=False)


def apply_operator_threaded(
    state, a_string, create_idx, anni_screen, num_active_orbs, parity_check, idx2det, det2idx, do_unsafe, tmp_state, factor
):
    """Gather-form twin of apply_operator_serial (osa.py:139-219); same result."""
    op = _string_operator("threaded", a_string, len(np.asarray(create_idx).ravel()), anni_screen, factor)
    return _accumulate_string(op, state, det2idx, do_unsafe, tmp_state, batched=False)


def apply_operator_SA_serial(
    state, a_string, create_screen, anni_idx, num_active_orbs, parity_check, idx2det, det2idx, do_unsafe, tmp_state, factor
):
    """Batch twin on ``state[n_states, N_det]`` (osa.py:282-348)."""
    op = _string_operator("serial", a_string, len(np.asarray(anni_idx).ravel()), create_screen, factor)
    return _accumulate_string(op, state, det2idx, do_unsafe, tmp_state, batched=True)


def apply_operator_SA_threaded(
    state, a_string, create_idx, anni_screen, num_active_orbs, parity_check, idx2det, det2idx, do_unsafe, tmp_state, factor
):
    """Batch twin of apply_operator_threaded (osa.py:351-410)."""
    op = _string_operator("threaded", a_string, len(np.asarray(create_idx).ravel()), anni_screen, factor)
    return _accumulate_string(op, state, det2idx, do_unsafe, tmp_state, batched=True)


def add_operator_matrix(
    op_mat, a_string, create_screen, anni_idx, num_active_orbs, parity_check, idx2det, det2idx, do_unsafe, factor
):
    """op_mat[tgt, src] += factor * sign for one ladder string (osa.py:222-279).  Small spaces only."""
    op = _string_operator("serial", a_string, len(np.asarray(anni_idx).ravel()), create_screen, factor)
    op_mat += build_operator_matrix(op, _space_of_tables(det2idx), do_unsafe=bool(do_unsafe))
    return op_mat


def get_determinant_expansion_from_operator_on_HF(
    operator: FermionicOperator, num_active_orbs: int, num_active_elec_alpha: int, num_active_elec_beta: int
) -> tuple[list[float], list[str]]:
    """Coefficients and determinants (bit strings a0 b0 a1 b1 ..., orbital 0 first) of ``operator|HF>``, one entry per ladder
    string that does not annihilate the Hartree-Fock determinant (osa.py:2979-3033).  Host only: every string goes through the
    closed-form string action of the engine (``sq_debug_string_action`` -- screens, flip masks and sign of DESIGN.md section 2)
    instead of the reference's operator-by-operator bit loop."""
    lib = _lib.load()
    n = int(num_active_orbs)
    handle = C.c_void_p()
    _lib.check(lib.sq_space_create(n, int(num_active_elec_alpha), int(num_active_elec_beta), -1, 0, -1, C.byref(handle)))
    occ_a, occ_b = (1 << int(num_active_elec_alpha)) - 1, (1 << int(num_active_elec_beta)) - 1
    coeffs: list[float] = []
    dets: list[str] = []
    valid, sign = C.c_int32(0), C.c_int32(0)
    tgt_a, tgt_b = C.c_uint32(0), C.c_uint32(0)
    try:
        for label, factor in operator.operators.items():
            ops = np.asarray([2 * int(i) + (1 if dag else 0) for i, dag in label] or [0], dtype=np.int32)
            _lib.check(
                lib.sq_debug_string_action(
                    handle, ops.ctypes.data_as(_PI), len(label), occ_a, occ_b, C.byref(valid), C.byref(tgt_a), C.byref(tgt_b), C.byref(sign)
                )
            )
            if not valid.value:
                continue
            coeffs.append(factor * sign.value)
            dets.append("".join(str((tgt_a.value >> o) & 1) + str((tgt_b.value >> o) & 1) for o in range(n)))
    finally:
        lib.sq_space_destroy(handle)
    return coeffs, dets
