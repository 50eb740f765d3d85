"""A/B at CAS(16,16): sigma build of a tUPS state (spin-flip symmetric): full build ("0"), half build on 32 x 32 blocked panels ("1"),
half build with the determinant-per-thread kernels ("tri")."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from slowquant_b200 import _lib
from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.ci_spaces import get_indexing
from slowquant_b200.operators import hamiltonian_0i_0a
from slowquant_b200.util import UpsStructure

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L = 4
info = get_indexing(0, n, 0, n // 2, n // 2)
lay = UpsStructure(); lay.create_tiled(n, {"n_layers": L, "do_tups": True})
th = np.random.default_rng(3).uniform(-np.pi, np.pi, lay.n_params)
rng = np.random.default_rng(2024)
A = rng.normal(size=(n, n)); h = A + A.T
B = 0.1 * rng.normal(size=(n, n, n, n))
g = B + B.transpose(1, 0, 2, 3); g = g + g.transpose(0, 1, 3, 2); g = g + g.transpose(2, 3, 0, 1)
H = hamiltonian_0i_0a(h, g, 0, n)
hf = torch.zeros(info.num_det, dtype=torch.float64, device="cuda"); hf[0] = 1.0
psi = osa.construct_ups_state(hf, info, th.tolist(), lay)
lib = _lib.load()
res = {}
for mode in (b"0", b"tri", b"1", b"0", b"tri", b"1"):
    lib.sq_set_option(b"sigma_spinsym", mode)
    osa.propagate_state([H], psi, info); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = osa.propagate_state([H], psi, info); torch.cuda.synchronize()
    print(f"CAS({n},{n}) sigma_spinsym={mode.decode()} sigma {1e3*(time.perf_counter()-t0):8.1f} ms  E = {float(torch.dot(psi, out)):.12f}", flush=True)
    res[mode] = out
for mode in (b"0", b"tri", b"1", b"0", b"tri", b"1"):
    lib.sq_set_option(b"sigma_spinsym", mode)
    osa.reduced_density_matrices(psi, psi, info); torch.cuda.synchronize()
    t0 = time.perf_counter(); d1, d2 = osa.reduced_density_matrices(psi, psi, info); torch.cuda.synchronize()
    print(f"CAS({n},{n}) sigma_spinsym={mode.decode()} rdm12 {1e3*(time.perf_counter()-t0):8.1f} ms  E_rdm = {float(np.sum(h*d1)+0.5*np.sum(g*d2)):.12f}  tr = {np.trace(d1):.12f}", flush=True)
    res[(mode, 'rdm')] = (d1, d2)
print("rdm1 diff", float(np.max(np.abs(res[(b'1','rdm')][0]-res[(b'0','rdm')][0]))), "rdm2 diff", float(np.max(np.abs(res[(b'1','rdm')][1]-res[(b'0','rdm')][1]))))
print("max|half - full| =", float(torch.max(torch.abs(res[b'1'] - res[b'0']))), " max|tri - full| =", float(torch.max(torch.abs(res[b'tri'] - res[b'0']))), " |sigma| =", float(torch.linalg.norm(res[b'0'])))
lib.sq_set_option(b"sigma_spinsym", b"1")
