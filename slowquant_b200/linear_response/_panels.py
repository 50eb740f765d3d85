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


def apply_unitary_rows(rows: torch.Tensor, index_info, dagger: bool) -> None:
    """rows[k] <- U rows[k] (or U^dagger) in place for every vector of a panel: the "U" / "Ud" elements of the
    reference's operator lists (osa.py:525-552) with the ansatz of ``index_info = (ci_info, thetas, layout)``."""
    ci_info, thetas, layout = index_info
    from slowquant_b200.util import UccStructure

    if rows.dim() == 1:
        rows = rows.unsqueeze(0)
    if isinstance(layout, UccStructure):
        for k in range(rows.shape[0]):
            rows[k].copy_(osa.construct_ucc_state(rows[k], ci_info, thetas, layout, dagger=dagger))
        return
    n = len(layout.excitation_operator_type)
    for k in range(rows.shape[0]):
        osa._ups_apply_inplace(rows[k], ci_info, thetas, layout, 0, n, dagger)


def orbital_blocks(lr) -> None:
    """q-q blocks of A, B, Sigma from the RDMs and the orbital-gradient guard, shared by every LR parametrisation
    (naive.py:42-54, 93-123 and the identical code in projected.py / statetransfer.py)."""
    import numpy as np

    from slowquant_b200.density_matrix import (
        get_orbital_gradient_response,
        get_orbital_response_hessian_block,
        get_orbital_response_metric_sigma,
    )

    wf = lr.wf
    nq = len(lr.q_ops)
    if nq == 0:
        return
    nI, nA = wf.num_inactive_orbs, wf.num_active_orbs
    k, kd = wf.kappa_no_activeactive_idx, wf.kappa_no_activeactive_idx_dagger
    grad = get_orbital_gradient_response(wf.h_mo, wf.g_mo, k, nI, nA, wf.rdm1, wf.rdm2)
    print("idx, max(abs(grad orb)):", np.argmax(np.abs(grad)), np.max(np.abs(grad)))
    if np.max(np.abs(grad)) > 10**-3:
        raise ValueError("Large Gradient detected in q of ", np.max(np.abs(grad)))
    lr.A[:nq, :nq] = get_orbital_response_hessian_block(wf.h_mo, wf.g_mo, kd, k, nI, nA, wf.rdm1, wf.rdm2)
    lr.B[:nq, :nq] = get_orbital_response_hessian_block(wf.h_mo, wf.g_mo, kd, kd, nI, nA, wf.rdm1, wf.rdm2)
    lr.Sigma[:nq, :nq] = get_orbital_response_metric_sigma(k, nI, nA, wf.rdm1)


def check_active_gradient(grad) -> None:
    import numpy as np

    if len(grad) != 0:
        print("idx, max(abs(grad active)):", np.argmax(np.abs(grad)), np.max(np.abs(grad)))
        if np.max(np.abs(grad)) > 10**-3:
            raise ValueError("Large Gradient detected in G of ", np.max(np.abs(grad)))


def mirror_lower(M: torch.Tensor):
    """The reference evaluates val(i, j) for i >= j only and stores it in [i, j] and [j, i]."""
    return (torch.tril(M) + torch.tril(M, -1).T).cpu().numpy()


def orbital_property_part(lr, mu, state_number: int, number_excitations: int) -> float:
    from slowquant_b200.density_matrix import get_orbital_response_property_gradient

    if len(lr.q_ops) == 0:
        return 0.0
    wf = lr.wf
    return get_orbital_response_property_gradient(
        mu, wf.kappa_no_activeactive_idx, wf.num_inactive_orbs, wf.num_active_orbs, wf.rdm1, lr.normed_response_vectors,
        state_number, number_excitations,
    )
