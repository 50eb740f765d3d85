"""Energy and orbital gradient from the active-space 1-/2-RDMs.

Same functions and argument order as the reference's slowquant/unitary_coupled_cluster/density_matrix.py
(RDM1 :5-43, RDM2 :46-136, get_electronic_energy :139-178, get_orbital_gradient :181-230, and the RDM-only
linear-response orbital blocks :233-563).  These are small
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


# ---------------------------------------------------------------------------------------------
# RDM-only orbital blocks of the linear-response equations (reference density_matrix.py:233-563;
# callers: linear_response/naive.py:43-120 and the other LR drivers).  Same signatures and return
# conventions as the reference; evaluated with dense full-space RDMs (zero on virtual indices) and
# einsum intermediates instead of the reference's per-element numba loops.
# ---------------------------------------------------------------------------------------------
def _padded_rdms(n_orb: int, num_inactive_orbs: int, num_active_orbs: int, rdm1, rdm2=None):
    """RDM1 / RDM2 on all n_orb orbitals (virtual rows and columns are zero, density_matrix.py:27-28, 86-87)."""
    M = num_inactive_orbs + num_active_orbs
    d1 = np.zeros((n_orb, n_orb))
    d1[:M, :M] = full_rdm1(num_inactive_orbs, num_active_orbs, np.asarray(rdm1))
    d2 = None
    if rdm2 is not None:
        d2 = np.zeros((n_orb, n_orb, n_orb, n_orb))
        d2[:M, :M, :M, :M] = full_rdm2(num_inactive_orbs, num_active_orbs, np.asarray(rdm1), np.asarray(rdm2))
    return d1, d2


def get_orbital_gradient_response(
    h_int: np.ndarray,
    g_int: np.ndarray,
    kappa_idx,
    num_inactive_orbs: int,
    num_active_orbs: int,
    rdm1: np.ndarray,
    rdm2: np.ndarray,
) -> np.ndarray:
    r""":math:`g^{\hat q}_{pq}=\langle0|[\hat q_{pq},\hat H]|0\rangle`, de-excitation part stacked behind the
    excitation part, times :math:`2^{-1/2}` (density_matrix.py:233-328)."""
    h = np.asarray(h_int)
    g = np.asarray(g_int)
    N = h.shape[0]
    M = num_inactive_orbs + num_active_orbs
    d1, d2 = _padded_rdms(N, num_inactive_orbs, num_active_orbs, rdm1, rdm2)
    one_a = h[:, :M] @ d1[:, :M].T                       # [n, m] = sum_p h[n,p] G1[m,p]
    one_b = h[:M, :].T @ d1[:M, :]                       # [m, n] = sum_p h[p,m] G1[p,n]
    A = np.einsum("npqr,mpqr->nm", g[:, :M, :M, :M], d2[:, :M, :M, :M], optimize=True)
    B = np.einsum("pmqr,pnqr->mn", g[:M, :, :M, :M], d2[:M, :, :M, :M], optimize=True)
    k = np.asarray(kappa_idx, dtype=np.int64).reshape(-1, 2)

    def block(m, n):
        return one_a[n, m] - one_b[m, n] + 0.5 * (A[n, m] - B[m, n] - A[m, n] + B[n, m])

    # second half: the loop variables are read as (n, m) (density_matrix.py:296)
    return 2 ** (-1 / 2) * np.concatenate([block(k[:, 0], k[:, 1]), block(k[:, 1], k[:, 0])])


def get_orbital_response_metric_sigma(kappa_idx, num_inactive_orbs: int, num_active_orbs: int, rdm1: np.ndarray) -> np.ndarray:
    r""":math:`\Sigma^{\hat q,\hat q}=\langle0|[\hat q^\dagger,\hat q]|0\rangle` (density_matrix.py:331-359)."""
    k = np.asarray(kappa_idx, dtype=np.int64).reshape(-1, 2)
    N = int(k.max()) + 1 if k.size else 0
    d1, _ = _padded_rdms(max(N, num_inactive_orbs + num_active_orbs), num_inactive_orbs, num_active_orbs, rdm1)
    n, m = k[:, 0][:, None], k[:, 1][:, None]
    p, q = k[:, 0][None, :], k[:, 1][None, :]
    sigma = (p == n) * d1[m, q] - (m == q) * d1[p, n]
    return -0.5 * sigma


def get_orbital_response_vector_norm(
    kappa_idx,
    num_inactive_orbs: int,
    num_active_orbs: int,
    rdm1: np.ndarray,
    response_vectors: np.ndarray,
    state_number: int,
    number_excitations: int,
) -> float:
    r"""Orbital part of the excited-state norm (density_matrix.py:362-416)."""
    k = np.asarray(kappa_idx, dtype=np.int64).reshape(-1, 2)
    K = len(k)
    N = int(k.max()) + 1 if k.size else 0
    d1, _ = _padded_rdms(max(N, num_inactive_orbs + num_active_orbs), num_inactive_orbs, num_active_orbs, rdm1)
    rv = np.asarray(response_vectors)
    z = rv[:K, state_number]
    y = rv[number_excitations : number_excitations + K, state_number]
    m, n = k[:, 0][:, None], k[:, 1][:, None]
    t, u = k[:, 0][None, :], k[:, 1][None, :]
    w = (n == u) * d1[m, t] - (m == t) * d1[n, u]
    return float(0.5 * (z @ w @ z - y @ w @ y))


def get_orbital_response_property_gradient(
    x_mo: np.ndarray,
    kappa_idx,
    num_inactive_orbs: int,
    num_active_orbs: int,
    rdm1: np.ndarray,
    response_vectors: np.ndarray,
    state_number: int,
    number_excitations: int,
) -> float:
    r"""Orbital part of the property gradient :math:`\sum_k\langle0|[\hat O_k,\hat X]|0\rangle` (density_matrix.py:419-461)."""
    x = np.asarray(x_mo)
    N = x.shape[0]
    M = num_inactive_orbs + num_active_orbs
    k = np.asarray(kappa_idx, dtype=np.int64).reshape(-1, 2)
    K = len(k)
    d1, _ = _padded_rdms(N, num_inactive_orbs, num_active_orbs, rdm1)
    xd = x[:, :M] @ d1[:, :M].T                          # [a, b] = sum_p x[a,p] G1[b,p]
    rv = np.asarray(response_vectors)
    z = rv[:K, state_number]
    y = rv[number_excitations : number_excitations + K, state_number]
    m, n = k[:, 0], k[:, 1]
    return float(2 ** (-1 / 2) * np.sum((y - z) * (xd[n, m] - xd[m, n])))


def get_orbital_response_hessian_block(
    h: np.ndarray,
    g: np.ndarray,
    kappa_idx1,
    kappa_idx2,
    num_inactive_orbs: int,
    num_active_orbs: int,
    rdm1: np.ndarray,
    rdm2: np.ndarray,
) -> np.ndarray:
    r""":math:`H^{\hat q,\hat q}_{tu,mn}=\langle0|[\hat q_{tu},[\hat H,\hat q_{mn}]]|0\rangle` (density_matrix.py:464-563).
    The twelve two-electron terms of the reference are six einsum intermediates used with two index orders each."""
    h = np.asarray(h)
    g = np.asarray(g)
    N = h.shape[0]
    M = num_inactive_orbs + num_active_orbs
    d1, d2 = _padded_rdms(N, num_inactive_orbs, num_active_orbs, rdm1, rdm2)
    S = slice(0, M)
    k1 = np.asarray(kappa_idx1, dtype=np.int64).reshape(-1, 2)
    k2 = np.asarray(kappa_idx2, dtype=np.int64).reshape(-1, 2)
    t, u = k1[:, 0][:, None], k1[:, 1][:, None]
    m, n = k2[:, 0][None, :], k2[:, 1][None, :]
    # one-electron part
    X1 = h[:, S] @ d1[:, S].T                            # [n, t] = sum_p h[n,p] G1[t,p]
    Y1 = h[S, :].T @ d1[S, :]                            # [m, u] = sum_p h[p,m] G1[p,u]
    A1 = h[n, t] * d1[m, u] + h[u, m] * d1[t, n] - (m == u) * X1[n, t] - (t == n) * Y1[m, u]
    # two-electron part: T_x[a,b,c,d] contract two summed indices p, q < M
    T1 = np.einsum("abpq,cdpq->abcd", g[:, :, S, S], d2[:, :, S, S], optimize=True)   # g[a,b,p,q] G2[c,d,p,q]
    T2 = np.einsum("apbq,cpdq->abcd", g[:, S, :, S], d2[:, S, :, S], optimize=True)   # g[a,p,b,q] G2[c,p,d,q]
    T3 = np.einsum("apqb,cpqd->abcd", g[:, S, S, :], d2[:, S, S, :], optimize=True)   # g[a,p,q,b] G2[c,p,q,d]
    T5 = np.einsum("pabq,pcdq->abcd", g[S, :, :, S], d2[S, :, :, S], optimize=True)   # g[p,a,b,q] G2[p,c,d,q]
    T6 = np.einsum("paqb,pcqd->abcd", g[S, :, S, :], d2[S, :, S, :], optimize=True)   # g[p,a,q,b] G2[p,c,q,d]
    T9 = np.einsum("pqab,pqcd->abcd", g[S, S, :, :], d2[S, S, :, :], optimize=True)   # g[p,q,a,b] G2[p,q,c,d]
    A2 = (
        T1[n, t, m, u] - T2[n, u, m, t] + T3[n, t, m, u] + T1[u, m, t, n] + T5[m, u, n, t] - T6[m, t, n, u]
        - T2[u, n, t, m] + T5[t, n, u, m] + T9[n, t, m, u] + T3[u, m, t, n] - T6[t, m, u, n] + T9[u, m, t, n]
    )
    Z1 = np.einsum("npqr,tpqr->nt", g[:, S, S, S], d2[:, S, S, S], optimize=True)
    Z2 = np.einsum("pmqr,puqr->mu", g[S, :, S, S], d2[S, :, S, S], optimize=True)
    Z3 = np.einsum("pqnr,pqtr->nt", g[S, S, :, S], d2[S, S, :, S], optimize=True)
    Z4 = np.einsum("pqrm,pqru->mu", g[S, S, S, :], d2[S, S, S, :], optimize=True)
    A2 = A2 - (m == u) * (Z1[n, t] + Z3[n, t]) - (t == n) * (Z2[m, u] + Z4[m, u])
    out = np.zeros((len(k1), len(k1)))                   # the reference allocates (len(idx1), len(idx1)) (:487-488)
    out[:, : len(k2)] = 0.5 * A1 + 0.25 * A2
    return out
