"""Golden matrices for the projected, state-transfer, self-consistent, all-state-transfer, all-self-consistent, all-projected and projected state-transfer linear-response parametrisations (reference
linear_response/projected.py, statetransfer.py, selfconsistent.py, allstatetransfer.py, allselfconsistent.py, allprojected.py, projected_statetransfer.py), produced by RUNNING THE REFERENCE in the build container at the FIXED
converged (theta, c_mo) of golden_config1.npz (no re-optimisation):

    python tests/golden/make_golden_lr_variants.py        ->  tests/golden/golden_lr_variants.npz
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
import slowquant.unitary_coupled_cluster.linear_response.allprojected as allprojlr  # noqa: E402
import slowquant.unitary_coupled_cluster.linear_response.allselfconsistent as allsclr  # noqa: E402
import slowquant.unitary_coupled_cluster.linear_response.allstatetransfer as allstlr  # noqa: E402
import slowquant.unitary_coupled_cluster.linear_response.projected as projlr  # noqa: E402
import slowquant.unitary_coupled_cluster.linear_response.projected_statetransfer as projstlr  # noqa: E402
import slowquant.unitary_coupled_cluster.linear_response.selfconsistent as sclr  # noqa: E402
import slowquant.unitary_coupled_cluster.linear_response.statetransfer as stlr  # noqa: E402
from slowquant.unitary_coupled_cluster.ups_wavefunction import WaveFunctionUPS  # noqa: E402

g1 = np.load(os.path.join(HERE, "golden_config1.npz"))
MOLECULES = {
    "h2o": (
        """O   0.0  0.0           0.1035174918;
        H   0.0  0.7955612117 -0.4640237459;
        H   0.0 -0.7955612117 -0.4640237459;""",
        (4, 4),
        {"n_layers": 3},
    ),
    "lih": (
        """Li 0.0           0.0  0.0;
           H  1.6717072740  0.0  0.0;""",
        (2, 2),
        {"n_layers": 1, "skip_last_singles": True},
    ),
}
out = {}
for name, (geom, cas, options) in MOLECULES.items():
    SQobj = sq.SlowQuant()
    SQobj.set_molecule(geom, distance_unit="angstrom")
    SQobj.set_basis_set("STO-3G")
    SQobj.init_hartree_fock()
    SQobj.hartree_fock.run_restricted_hartree_fock()
    WF = WaveFunctionUPS(cas, g1[name + "_c_mo"], SQobj, "tUPS", ansatz_options=dict(options), include_active_kappa=True)
    WF.thetas = g1[name + "_thetas"].tolist()
    assert abs(WF.energy_elec - float(g1[name + "_energy"])) < 1e-9
    for tag, mod in (("proj", projlr), ("st", stlr), ("sc", sclr), ("allst", allstlr), ("allsc", allsclr), ("allproj", allprojlr), ("projst", projstlr)):
        LR = mod.LinearResponse(WF, excitations="SD")
        LR.calc_excitation_energies()
        pre = f"{name}_{tag}_"
        for key in ("A", "B", "Sigma", "Delta"):
            out[pre + key] = np.array(getattr(LR, key))
        out[pre + "excitation_energies"] = np.array(LR.excitation_energies)
        out[pre + "norms"] = np.array(LR.get_excited_state_norm())
        out[pre + "oscillator_strengths"] = np.array(LR.get_oscillator_strength())
        print(name, tag, LR.excitation_energies[:4])
np.savez_compressed(os.path.join(HERE, "golden_lr_variants.npz"), **out)
