"""Matrix-free exp(T)|state> for the non-factorised UCC wave function.

The reference builds the dense N_det x N_det matrix of T and calls scipy's ``expm_multiply``
(operator_state_algebra.py:870-896).  Here T is applied string-wise by the gather kernel and the
exponential is a scaled Taylor series (the algorithm family of Al-Mohy & Higham that scipy uses, with a
fixed conservative scaling): exp(T) v = (exp(T/s))^s v, each factor summed until the term is below
double-precision round-off.  T is anti-Hermitian, so the series is norm-stable.
"""
from __future__ import annotations

import math
from collections.abc import Sequence

import numpy as np
import torch

from slowquant_b200 import _lib
from slowquant_b200.fermionic_operator import FermionicOperator
from slowquant_b200.operators import G1, G2, G3, G4, G5, G6, G1_sa, G2_sa


def get_ucc_T(thetas: Sequence[float], ucc_struct, offset: int = 0) -> FermionicOperator:
    """T = sum_k theta_k (G_k - G_k^dagger) over the UCC layout (operator_state_algebra.py:899-960)."""
    T = FermionicOperator({})
    gens = {"single": G1, "double": G2, "triple": G3, "quadruple": G4, "quintuple": G5, "sextuple": G6}
    for exc_type, exc_indices, theta in zip(ucc_struct.excitation_operator_type, ucc_struct.excitation_indices, thetas):
        if abs(theta) < 10**-28:
            continue
        if exc_type == "sa_single":
            i, a = (int(x) + offset for x in exc_indices)
            T += float(theta) * G1_sa(i, a, True)
        elif exc_type.startswith("sa_double_"):
            i, j, a, b = (int(x) + offset for x in exc_indices)
            T += float(theta) * G2_sa(i, j, a, b, int(exc_type[-1]), True)
        elif exc_type in gens:
            idx = [int(x) + 2 * offset for x in exc_indices]
            T += float(theta) * gens[exc_type](*idx, True)
        else:
            raise ValueError(f"Got unknown excitation type, {exc_type}")
    return T


def expm_multiply_operator(T: FermionicOperator, state: torch.Tensor, ci_info, scale: float = 1.0) -> torch.Tensor:
    """exp(scale * T)|state> with T applied matrix-free on the device; returns a new tensor."""
    from slowquant_b200 import operator_state_algebra as osa

    lib = _lib.load()
    ops_flat, offsets, coeffs = osa.encode_operator(T)
    coeffs = coeffs * scale
    out = state.clone()
    if len(coeffs) == 0:
        return out
    # ||T|| <= sum |c_s| (every ladder string is a partial isometry); Taylor steps of norm <= 1
    bound = float(np.sum(np.abs(coeffs)))
    steps = max(1, int(math.ceil(bound)))
    term = torch.empty_like(out)
    nxt = torch.empty_like(out)
    PI, PD = osa._PI, osa._PD
    c = np.ascontiguousarray(coeffs / steps)
    for _ in range(steps):
        term.copy_(out)
        for k in range(1, 200):
            _lib.check(
                lib.sq_apply_strings(
                    ci_info._handle, len(c), ops_flat.ctypes.data_as(PI), offsets.ctypes.data_as(PI), c.ctypes.data_as(PD),
                    osa._ptr(term), osa._ptr(nxt), 0, 0, osa._stream(),
                )
            )
            nxt /= k
            out += nxt
            term, nxt = nxt, term
            if float(torch.linalg.norm(term)) <= 1e-17 * max(float(torch.linalg.norm(out)), 1e-300):
                break
    return out
