"""TEST INFRASTRUCTURE -- never imported by the product (slowquant_b200/), by bench.py or by the GPU tests.

Runs the Python HOST logic of the callers of the engine (linear-response matrix assembly from panels, state-averaged
bookkeeping, extended-space embedding, higher-RDM contractions, RotoSolve) on a machine WITHOUT a GPU: the handful of
functions of ``operator_state_algebra`` that call libsqsv are replaced, in THIS process only, by the oracle (the CPU
restatement of the reference) acting on CPU tensors.  What is checked this way is the host code around the kernels,
against the same reference goldens the GPU tests use; the kernels themselves are checked by ``-m gpu`` tests only.
Importing this module patches ``slowquant_b200`` globally, so it is only ever imported by ``tests/host_callers_check.py``,
which ``tests/test_host_callers.py`` runs in a subprocess.
"""
import sys
import numpy as np, torch
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sq_oracle as orc
import slowquant_b200.ci_spaces as cis
import slowquant_b200.operator_state_algebra as osa
import slowquant_b200.ups_wavefunction as upsm

class TorchProxy:
    def __getattr__(self, k): return getattr(torch, k)
    def device(self, *a, **k): return torch.device("cpu")
for m in (osa, upsm):
    m.torch = TorchProxy()
cis._current_device = lambda: -1
_spaces = {}
def space_of(ci):
    key = (ci.num_inactive_orbs, ci.num_active_orbs, ci.num_virtual_orbs, ci.num_active_elec_alpha, ci.num_active_elec_beta)
    if key not in _spaces: _spaces[key] = orc.get_indexing(*key)
    return _spaces[key]
def _to_device(state, ci, copy=True):
    if isinstance(state, torch.Tensor): return (state.clone() if copy else state), False
    return torch.from_numpy(np.array(state, dtype=np.float64)), True
def _from_device(t, was_numpy): return t.numpy().copy() if was_numpy else t
osa._to_device = _to_device; osa._from_device = _from_device
def _apply_operator(op, src, dst, ci, do_unsafe):
    out = orc.propagate_state([dict(op.operators)], src.numpy().copy(), space_of(ci), do_folding=False)
    dst.copy_(torch.from_numpy(np.asarray(out)))
osa._apply_operator = _apply_operator
def _apply_hamiltonian(op, src, dst, ci):
    H = orc.hamiltonian_0i_0a(op.h_mo, op.g_mo, op.num_inactive_orbs, op.num_active_orbs)
    out = orc.propagate_state([H], src.numpy().copy(), space_of(ci), do_folding=True)
    dst.copy_(torch.from_numpy(np.asarray(out))); return True
osa._apply_hamiltonian = _apply_hamiltonian
def _ups_apply_inplace(t, ci, thetas, lay, first, last, dagger):
    off = getattr(ci, "space_extension_offset", 0)
    types = lay.excitation_operator_type[first:last]
    idx = [tuple(int(x) + (off if t.startswith("sa_") else 2 * off) for x in ix) for t, ix in zip(types, lay.excitation_indices[first:last])]
    th = np.asarray(thetas, dtype=np.float64)[first:last]
    out = orc.construct_ups_state(t.numpy().copy(), space_of(ci), th, types, idx, dagger=dagger)
    t.copy_(torch.from_numpy(np.asarray(out)))
osa._ups_apply_inplace = _ups_apply_inplace
osa._dot = lambda a, b, ci: float(torch.dot(a, b))
def get_grad_action(state, idx, ci, lay):
    if osa._is_extended(ci):
        full, was = osa._embed(state, ci)
        return osa._from_device(osa._restrict(get_grad_action(full, idx, ci.parent, lay), ci), was)
    t, was = _to_device(state, ci, copy=False)
    off = getattr(ci, "space_extension_offset", 0)
    sh = [tuple(int(x) + (off if ty.startswith("sa_") else 2 * off) for x in ix) for ty, ix in zip(lay.excitation_operator_type, lay.excitation_indices)]
    out = orc.get_grad_action(t.numpy().copy(), idx, space_of(ci), lay.excitation_operator_type, sh)
    return _from_device(torch.from_numpy(np.asarray(out)), was)
osa.get_grad_action = get_grad_action
def ups_gradient_sweep(bra, ket, ci, thetas, lay):
    b, _ = _to_device(bra, ci); k, _ = _to_device(ket, ci)
    n = len(lay.excitation_operator_type); g = np.zeros(n)
    for i in range(n):
        g[i] = 2 * float(torch.dot(b, get_grad_action(k, i, ci, lay)))
        _ups_apply_inplace(b, ci, thetas, lay, i, i + 1, False); _ups_apply_inplace(k, ci, thetas, lay, i, i + 1, False)
    return g, b, k
osa.ups_gradient_sweep = ups_gradient_sweep
def ups_gradient_sweep_backward(bra, ket, ci, thetas, lay):
    b, _ = _to_device(bra, ci); k, _ = _to_device(ket, ci)
    n = len(lay.excitation_operator_type); g = np.zeros(n)
    for i in reversed(range(n)):
        g[i] = 2 * float(torch.dot(b, get_grad_action(k, i, ci, lay)))
        _ups_apply_inplace(b, ci, thetas, lay, i, i + 1, True); _ups_apply_inplace(k, ci, thetas, lay, i, i + 1, True)
    return g, b, k
osa.ups_gradient_sweep_backward = ups_gradient_sweep_backward
def reduced_density_matrices(bra, ket, ci, want_rdm2=True):
    from slowquant_b200 import operators as mops
    n = ci.num_active_orbs; sp = space_of(ci)
    b = bra.numpy() if isinstance(bra, torch.Tensor) else np.asarray(bra); k = ket.numpy() if isinstance(ket, torch.Tensor) else np.asarray(ket)
    E = [[np.asarray(orc.propagate_state([dict(mops.Epq(p, q).operators)], k.copy(), sp, do_folding=False)) for q in range(n)] for p in range(n)]
    Eb = [[np.asarray(orc.propagate_state([dict(mops.Epq(q, p).operators)], b.copy(), sp, do_folding=False)) for q in range(n)] for p in range(n)]
    d1 = np.array([[b @ E[p][q] for q in range(n)] for p in range(n)])
    d2 = None
    if want_rdm2:
        d2 = np.zeros((n,)*4)
        for p in range(n):
            for q in range(n):
                for r in range(n):
                    for s in range(n):
                        d2[p,q,r,s] = Eb[p][q] @ E[r][s] - (d1[p,s] if q == r else 0.0)
    return d1, d2
osa.reduced_density_matrices = reduced_density_matrices


# matrix-free exp(T)|state> of ucc_state.py -> dense matrix of T from the oracle + scipy (small spaces only)
import scipy.sparse.linalg  # noqa: E402

import slowquant_b200.ucc_state as _ucc_state  # noqa: E402
import slowquant_b200.ucc_wavefunction as _ucc_wf  # noqa: E402


def expm_multiply_operator(T, state, ci_info, scale=1.0):
    n = state.numel()
    sp = space_of(ci_info)
    M = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        M[:, j] = orc.propagate_state([dict(T.operators)], e, sp, do_folding=False)
    return torch.from_numpy(np.asarray(scipy.sparse.linalg.expm_multiply(scale * M, state.numpy())))


_ucc_state.expm_multiply_operator = expm_multiply_operator
_ucc_wf.expm_multiply_operator = expm_multiply_operator
for _m in (_ucc_state, _ucc_wf):
    if hasattr(_m, "torch"):
        _m.torch = TorchProxy()
