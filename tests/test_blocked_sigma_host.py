"""Host-side restatement of the blocked half sigma build (sqsv_hamiltonian.cu: build_blk_kernel / scatter_blk_kernel /
spinsym_symmetrize_kernel) in dense numpy algebra, with the oracle supplying the spin-resolved E_pq matrices: for a spin-flip
symmetric vector c (c[B,A] = lambda phi(A,B) c[A,B]),

    D_f = <J|E_f|c> with the beta partners read through their mirrors,      F = Gm D,      val = F + k c,
    Y   = sum_f  E^alpha_f (w val_f)  +  lambda U E^beta_f (w val_f)        (w = 1 above, 1/2 on, 0 below the diagonal),
    sigma = (1 + lambda U)(e_core / 2 * c + Y)

must equal H c.  This pins the algebra of the scheme (weights, mirrors, symmetrisation) on the CPU; the kernels' index
arithmetic is checked by the GPU tests."""
import numpy as np
import pytest

from oracle import sq_oracle as orc


def _dense(op, sp):
    m = np.zeros((sp.num_det, sp.num_det))
    for j in range(sp.num_det):
        e = np.zeros(sp.num_det)
        e[j] = 1.0
        m[:, j] = orc.propagate_state([op], e, sp, do_folding=False)
    return m


@pytest.mark.parametrize("n,ne,L", [(4, 2, 2), (5, 2, 2), (4, 1, 2), (5, 3, 2)])
def test_blocked_half_sigma_algebra(n, ne, L):
    from slowquant_b200.ci_spaces import get_indexing

    sp = orc.get_indexing(0, n, 0, ne, ne)
    strings = get_indexing(0, n, 0, ne, ne, device=-1).strings(0).astype(np.int64)
    N = len(strings)
    assert N * N == sp.num_det
    rng = np.random.default_rng(7 + n)
    A = rng.normal(size=(n, n))
    h = A + A.T
    B = 0.1 * rng.normal(size=(n, n, n, n))
    g = B + B.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    types, idx = orc.tiled_layout(n, L)
    th = rng.uniform(-np.pi, np.pi, len(types))
    hf = np.zeros(sp.num_det)
    hf[0] = 1.0
    c = orc.construct_ups_state(hf, sp, th, types, idx)
    want = orc.propagate_state([orc.hamiltonian_0i_0a(h, g, 0, n)], c, sp)
    # spin flip: (U x)[A,B] = phi(A,B) x[B,A], phi = (-1)^popcount(A & B)
    phi = np.array([[(-1.0) ** bin(int(a & b)).count("1") for b in strings] for a in strings])
    U = lambda x: (phi * x.reshape(N, N).T).reshape(-1)
    lam = 1.0 if np.max(np.abs(U(c) - c)) < 1e-12 else -1.0
    assert np.max(np.abs(U(c) - lam * c)) < 1e-12, "a tUPS state on a closed-shell reference is spin-flip symmetric"
    Ea = {(p, q): _dense({((2 * p, True), (2 * q, False)): 1.0}, sp) for p in range(n) for q in range(n)}
    Eb = {(p, q): _dense({((2 * p + 1, True), (2 * q + 1, False)): 1.0}, sp) for p in range(n) for q in range(n)}
    pairs = [(p, q) for p in range(n) for q in range(p + 1)]
    sym = lambda E, p, q: E[p, q] if p == q else E[p, q] + E[q, p]
    # gather: alpha partners directly, beta partners through the mirror c[A,B'] = lambda phi(A,B') c[B',A]
    mirror = lam * U(c)
    D = np.array([sym(Ea, p, q) @ c + sym(Eb, p, q) @ mirror for p, q in pairs])
    Gm = np.array([[0.5 * g[p, q, r, t] for r, t in pairs] for p, q in pairs])
    k = np.array([h[p, q] - 0.5 * sum(g[p, r, r, q] for r in range(n)) for p, q in pairs])
    val = Gm @ D + k[:, None] * c[None, :]
    ia, ib = np.divmod(np.arange(sp.num_det), N)
    w = np.where(ia < ib, 1.0, np.where(ia == ib, 0.5, 0.0))
    Y = np.zeros(sp.num_det)
    for f, (p, q) in enumerate(pairs):
        Y += sym(Ea, p, q) @ (w * val[f]) + lam * U(sym(Eb, p, q) @ (w * val[f]))
    got = Y + lam * U(Y)          # e_core = 0 for a space without inactive orbitals
    assert np.max(np.abs(got - want)) < 1e-11
    # the same with every source (no weights, no symmetrisation) is the textbook build
    full = sum((sym(Ea, p, q) + sym(Eb, p, q)) @ val[f] for f, (p, q) in enumerate(pairs))
    assert np.max(np.abs(full - want)) < 1e-11
