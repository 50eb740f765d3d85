"""Energy and orbital gradient from the active-space 1-/2-RDMs.

Same functions and argument order as the reference's slowquant/unitary_coupled_cluster/density_matrix.py
(RDM1 :5-43, RDM2 :46-136, get_electronic_energy :139-178, get_orbital_gradient :181-230).  These are small
dense contractions over (inactive + active) orbitals -- O(K (nI+nA)^3) flops on arrays of a few hundred kB --
so they are evaluated as vectorised einsums on the host; the expensive inputs (rdm1, rdm2) come from the
CUDA kernels (sq_rdm12).
"""
from __future__ import annotations

import numpy as np


def RDM1(p: int, q: int, num_inactive_orbs: int, num_active_orbs: int, rdm1: np.ndarray) -> float:
    """Full-space 1-RDM element from the active block (density_matrix.py:5-43)."""
    virt_start = num_inactive_orbs + num_active_orbs
    if p >= virt_start or q >= virt_start:
        return 0
    if p >= num_inactive_orbs and q >= num_inactive_orbs:
        return rdm1[p - num_inactive_orbs, q - num_inactive_orbs]
    if p < num_inactive_orbs and q < num_inactive_orbs:
        return 2 if p == q else 0
    return 0


def full_rdm1(num_inactive_orbs: int, num_active_orbs: int, rdm1: np.ndarray) -> np.ndarray:
    """[M, M] array of RDM1(p, q) for p, q < M = nI + nA."""
    nI, nA = num_inactive_orbs, num_active_orbs
    M = nI + nA
    out = np.zeros((M, M))
    out[np.arange(nI), np.arange(nI)] = 2.0
    out[nI:, nI:] = rdm1
    return out


def full_rdm2(num_inactive_orbs: int, num_active_orbs: int, rdm1: np.ndarray, rdm2: np.ndarray) -> np.ndarray:
    """[M, M, M, M] array of RDM2(p, q, r, s) (density_matrix.py:46-136) for indices < M = nI + nA."""
    nI, nA = num_inactive_orbs, num_active_orbs
    M = nI + nA
    out = np.zeros((M, M, M, M))
    a = slice(nI, M)
    out[a, a, a, a] = rdm2
    eye = np.eye(nI)
    for i in range(nI):
        out[i, a, a, i] = -rdm1          # iuvj, p == s
        out[a, i, i, a] = -rdm1          # uijv, q == r
        out[a, a, i, i] = 2.0 * rdm1     # uvij, r == s
        out[i, i, a, a] = 2.0 * rdm1     # ijuv, p == q
    if nI:
        ii = slice(0, nI)
        out[ii, ii, ii, ii] = 4.0 * np.einsum("pq,rs->pqrs", eye, eye) - 2.0 * np.einsum("qr,ps->pqrs", eye, eye)
    return out


def RDM2(
    p: int, q: int, r: int, s: int, num_inactive_orbs: int, num_active_orbs: int, rdm1: np.ndarray, rdm2: np.ndarray
) -> float:
    """Full-space 2-RDM element (density_matrix.py:46-136)."""
    M = num_inactive_orbs + num_active_orbs
    if max(p, q, r, s) >= M:
        return 0
    nI = num_inactive_orbs
    ap, aq, ar, as_ = p >= nI, q >= nI, r >= nI, s >= nI
    if ap and aq and ar and as_:
        return rdm2[p - nI, q - nI, r - nI, s - nI]
    if (not ap) and aq and ar and (not as_):
        return -rdm1[q - nI, r - nI] if p == s else 0
    if ap and (not aq) and (not ar) and as_:
        return -rdm1[p - nI, s - nI] if q == r else 0
    if ap and aq and (not ar) and (not as_):
        return 2 * rdm1[p - nI, q - nI] if r == s else 0
    if (not ap) and (not aq) and ar and as_:
        return 2 * rdm1[r - nI, s - nI] if p == q else 0
    if not (ap or aq or ar or as_):
        val = 0
        if p == q and r == s:
            val += 4
        if q == r and p == s:
            val -= 2
        return val
    return 0


def get_electronic_energy(
    h_int: np.ndarray, g_int: np.ndarray, num_inactive_orbs: int, num_active_orbs: int, rdm1: np.ndarray, rdm2: np.ndarray
) -> float:
    r""":math:`E=\sum_{pq}h_{pq}\Gamma^{[1]}_{pq}+\tfrac12\sum_{pqrs}g_{pqrs}\Gamma^{[2]}_{pqrs}` (density_matrix.py:139-178)."""
    M = num_inactive_orbs + num_active_orbs
    d1 = full_rdm1(num_inactive_orbs, num_active_orbs, np.asarray(rdm1))
    d2 = full_rdm2(num_inactive_orbs, num_active_orbs, np.asarray(rdm1), np.asarray(rdm2))
    h = np.asarray(h_int)[:M, :M]
    g = np.asarray(g_int)[:M, :M, :M, :M]
    return float(np.sum(h * d1) + 0.5 * np.sum(g * d2))


def get_orbital_gradient(
    h_int: np.ndarray,
    g_int: np.ndarray,
    kappa_idx,
    num_inactive_orbs: int,
    num_active_orbs: int,
    rdm1: np.ndarray,
    rdm2: np.ndarray,
) -> np.ndarray:
    r"""Orbital gradient :math:`\langle 0|[\hat\kappa_{mn},\hat H]|0\rangle` for (m, n) in kappa_idx (density_matrix.py:181-230)."""
    h = np.asarray(h_int)
    g = np.asarray(g_int)
    N = h.shape[0]
    M = num_inactive_orbs + num_active_orbs
    d1 = np.zeros((N, M))
    d1[:M] = full_rdm1(num_inactive_orbs, num_active_orbs, np.asarray(rdm1))
    d2 = np.zeros((N, M, M, M))
    d2[:M] = full_rdm2(num_inactive_orbs, num_active_orbs, np.asarray(rdm1), np.asarray(rdm2))
    d1T = np.zeros((M, N))
    d1T[:, :M] = d1[:M]
    d2b = np.zeros((M, N, M, M))
    d2b[:, :M] = d2[:M]
    # one-electron: 2 sum_p h[n,p] G1[m,p] - 2 sum_p h[p,m] G1[p,n]
    one_a = h[:, :M] @ d1.T            # [n, m]
    one_b = h[:M, :].T @ d1T           # [m, n]
    # two-electron: A[n,m] = sum g[n,p,q,r] G2[m,p,q,r];  B[m,n] = sum g[p,m,q,r] G2[p,n,q,r]
    A = np.einsum("npqr,mpqr->nm", g[:, :M, :M, :M], d2, optimize=True)
    B = np.einsum("pmqr,pnqr->mn", g[:M, :, :M, :M], d2b, optimize=True)
    k = np.asarray(kappa_idx, dtype=np.int64).reshape(-1, 2)
    m, n = k[:, 0], k[:, 1]
    return 2.0 * one_a[n, m] - 2.0 * one_b[m, n] + A[n, m] - B[m, n] - A[m, n] + B[n, m]
