"""Where the time of `WF.thetas = x` goes at CAS(16,16), L=16: kernels vs clone vs allocation vs host work."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from slowquant_b200 import _lib
from slowquant_b200 import operator_state_algebra as osa
from slowquant_b200.ci_spaces import get_indexing
from slowquant_b200.util import UpsStructure

n, L = 16, 16
info = get_indexing(0, n, 0, 8, 8)
lay = UpsStructure(); lay.create_tiled(n, {"n_layers": L, "do_tups": True})
P = lay.n_params
th = np.random.default_rng(1234).uniform(-np.pi, np.pi, P)
thl = th.tolist()
hf = torch.zeros(info.num_det, dtype=torch.float64, device="cuda"); hf[0] = 1.0
dense = torch.randn(info.num_det, dtype=torch.float64, device="cuda")

def timeit(label, fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    print(f"{label:60s} {(time.perf_counter()-t0)/reps*1e3:8.2f} ms", flush=True)

work = dense.clone()
timeit("_ups_apply_inplace on a dense resident vector", lambda: osa._ups_apply_inplace(work, info, th, lay, 0, P, False))
w2 = hf.clone()
def hf_inplace():
    w2.copy_(hf); osa._ups_apply_inplace(w2, info, th, lay, 0, P, False)
timeit("copy_(HF) + _ups_apply_inplace", hf_inplace)
timeit("construct_ups_state(HF device tensor, ndarray thetas)", lambda: osa.construct_ups_state(hf, info, th, lay))
timeit("construct_ups_state(HF device tensor, list thetas)", lambda: osa.construct_ups_state(hf, info, thl, lay))
timeit("construct_ups_state(dense device tensor)", lambda: osa.construct_ups_state(dense, info, thl, lay))
timeit("clone only", lambda: dense.clone())
def with_item():
    r = osa.construct_ups_state(hf, info, thl, lay); return float(r[0].item())
timeit("construct_ups_state(HF) + .item()", with_item)
from slowquant_b200.integral_manager import ArrayIntegrals
from slowquant_b200.ups_wavefunction import WaveFunctionUPS
rng_i = np.random.default_rng(2024)
h_syn = rng_i.normal(size=(n, n))
WF = WaveFunctionUPS((n, n), np.eye(n), ArrayIntegrals(h_syn + h_syn.T, np.zeros((n, n, n, n)), num_elec=n), "tUPS", {"n_layers": L}, device=0)
def setter():
    WF.thetas = thl
timeit("WF.thetas = list", setter)
def setter_item():
    WF.thetas = thl
    return float(WF.ci_coeffs_device[0].item())
timeit("WF.thetas = list; ci_coeffs_device[0].item()", setter_item)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); setter(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
