"""Golden vectors for the RotoSolve path (reference optimizers.py:166-492, ups_wavefunction.py:1144-1194), produced by
RUNNING THE REFERENCE in the build container:

    python tests/golden/make_golden_rotosolve.py        ->  tests/golden/golden_rotosolve.npz

  * reconstructed_f / reconstructed_f_derivative on seeded inputs (single-state lists and state-averaged lists);
  * _calc_energy_rotosolve_optimization of H2O/STO-3G tUPS(4,4) L=2 at the `tups44` thetas of golden.npz for three
    parameter indices;
  * a 3-sweep RotoSolve optimisation of H2O/STO-3G tUPS(4,4) L=2 from seeded random thetas (final energy).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
stub = tempfile.mkdtemp(prefix="pyscf_stub_")
os.makedirs(os.path.join(stub, "pyscf", "gto"))
open(os.path.join(stub, "pyscf", "__init__.py"), "w").write("from . import gto\n")
open(os.path.join(stub, "pyscf", "gto", "__init__.py"), "w").write("from . import mole\n")
open(os.path.join(stub, "pyscf", "gto", "mole.py"), "w").write("class Mole:\n    pass\n")
sys.path.insert(0, stub)
sys.path.insert(0, "/root/reference")

import slowquant.SlowQuant as sq  # noqa: E402
from slowquant.unitary_coupled_cluster import optimizers as ropt  # noqa: E402
from slowquant.unitary_coupled_cluster.ups_wavefunction import WaveFunctionUPS  # noqa: E402

out = {}
rng = np.random.default_rng(31)
xs = rng.uniform(-np.pi, np.pi, 40)
xs[0] = 0.0  # a shift point itself (sinc limits)
out["x_vals"] = xs
for R in (1, 2, 4):
    e_single = [float(v) for v in rng.normal(size=2 * R + 1)]
    e_sa = [rng.normal(size=3) for _ in range(2 * R + 1)]
    out[f"R{R}_e_single"] = np.array(e_single)
    out[f"R{R}_e_sa"] = np.array(e_sa)
    out[f"R{R}_f_single"] = ropt.reconstructed_f(xs, e_single, R)
    out[f"R{R}_df_single"] = ropt.reconstructed_f_derivative(xs, e_single, R)
    out[f"R{R}_f_sa"] = ropt.reconstructed_f(xs, e_sa, R)
    out[f"R{R}_df_sa"] = ropt.reconstructed_f_derivative(xs, e_sa, R)

SQobj = sq.SlowQuant()
SQobj.set_molecule(
    """O   0.0  0.0           0.1035174918;
    H   0.0  0.7955612117 -0.4640237459;
    H   0.0 -0.7955612117 -0.4640237459;""",
    distance_unit="angstrom",
)
SQobj.set_basis_set("sto-3g")
SQobj.init_hartree_fock()
SQobj.hartree_fock.run_restricted_hartree_fock()
c_mo = SQobj.hartree_fock.mo_coeff
g0 = np.load(os.path.join(HERE, "golden.npz"))
assert np.max(np.abs(g0["h2o_c_mo"] - c_mo)) < 1e-10
WF = WaveFunctionUPS((4, 4), c_mo, SQobj, "tUPS", ansatz_options={"n_layers": 2}, include_active_kappa=True)
th = g0["tups44_thetas"].tolist()
for idx in (0, 4, len(th) - 1):
    name = WF.ups_layout.param_names[idx]
    R = WF.ups_layout.grad_param_R[name]
    shifts = [2 * mu / (2 * R + 1) * np.pi for mu in range(-R, R + 1)]
    out[f"rs_idx{idx}_shifts"] = np.array(shifts)
    out[f"rs_idx{idx}_energies"] = np.array(WF._calc_energy_rotosolve_optimization(th, shifts, idx))

WF1 = WaveFunctionUPS((4, 4), c_mo, SQobj, "tUPS", ansatz_options={"n_layers": 2})
th0 = np.random.default_rng(5).uniform(-0.4, 0.4, WF1.ups_layout.n_params)
WF1.thetas = th0.tolist()
out["opt_start_thetas"] = th0
out["opt_start_energy"] = np.array(WF1.energy_elec)
WF1.run_wf_optimization_1step("rotosolve", False, maxiter=3)
# parameters the energy does not depend on (e.g. the pair double on two occupied orbitals) are set from rounding noise:
# only the energy is a stable quantity (checked: start + 1e-13 changes such thetas by 0.1 and the energy by 3e-11)
out["opt_thetas"] = np.array(WF1.thetas)
out["opt_energy"] = np.array(WF1.energy_elec)
np.savez_compressed(os.path.join(HERE, "golden_rotosolve.npz"), **out)
print("opt energy", float(WF1.energy_elec), WF1.thetas)
