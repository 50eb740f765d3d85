"""Device panels for the linear-response matrix builds.

A *panel* is a row-major fp64 device matrix ``P[k, :] = O_k |psi>`` (one CI vector per operator).  The reference
builds every element of A, B and Sigma from ``propagate_state`` + ``expectation_value`` pairs inside a Python
double loop (linear_response/naive.py:232-304: 16 dots and 2 operator applications per (I, J)); here each
operator is applied once, the results stay in HBM, and the blocks are Gram matrices ``P Q^T`` -- fp64 GEMMs
(cuBLAS; the one GEMM-shaped step on this path).  Rows are filled in place by the same kernels
``propagate_state`` uses (``sq_apply_strings`` gather kernel, ``sq_sigma``).
"""
from __future__ import annotations

from collections.abc import Sequence

import torch

from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.ci_spaces import CI_Info
from slowquant_b200.fermionic_operator import FermionicOperator
from slowquant_b200.operators import ActiveSpaceHamiltonian


def state_on_device(state, ci_info: CI_Info) -> torch.Tensor:
    """fp64 device copy of a host/device CI vector (never aliases the caller's array)."""
    t, _ = osa._to_device(state, ci_info)
    return t


def fold(op: FermionicOperator, ci_info: CI_Info) -> FermionicOperator:
    return op.get_folded_operator(ci_info.num_inactive_orbs, ci_info.num_active_orbs, ci_info.num_virtual_orbs)


def apply_into(op: FermionicOperator, src: torch.Tensor, dst: torch.Tensor, ci_info: CI_Info) -> None:
    """dst <- fold(op)|src> (dst is a vector or a panel row; overwritten)."""
    if isinstance(op, ActiveSpaceHamiltonian) and osa._apply_hamiltonian(op, src, dst, ci_info):
        return
    osa._apply_operator(fold(op, ci_info), src, dst, ci_info, False)


def apply(op: FermionicOperator, src: torch.Tensor, ci_info: CI_Info) -> torch.Tensor:
    dst = torch.empty_like(src)
    apply_into(op, src, dst, ci_info)
    return dst


def panel_from_operators(ops: Sequence[FermionicOperator], src: torch.Tensor, ci_info: CI_Info) -> torch.Tensor:
    """P[k] = ops[k]|src>."""
    P = torch.empty((len(ops), src.numel()), dtype=torch.float64, device=src.device)
    for k, op in enumerate(ops):
        apply_into(op, src, P[k], ci_info)
    return P


def panel_from_rows(op: FermionicOperator, rows: torch.Tensor, ci_info: CI_Info) -> torch.Tensor:
    """P[k] = op|rows[k]> (one operator applied to every vector of a panel)."""
    P = torch.empty_like(rows)
    for k in range(rows.shape[0]):
        apply_into(op, rows[k], P[k], ci_info)
    return P


def gram(P: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
    """G[i, j] = <P_i|Q_j>  (one DGEMM over the determinant index)."""
    return P @ Q.T
