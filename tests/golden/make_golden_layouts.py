"""Golden ansatz layouts over MANY option combinations, produced by RUNNING THE REFERENCE's UpsStructure / UccStructure
(util.py:547-1073) in the build container:

    python tests/golden/make_golden_layouts.py        ->  tests/golden/golden_layouts.json.gz

create_fUCC with seeded random subsets of its 13 excitation flags (1-2 layers) on three occupied / unoccupied index sets (incl. an
open-shell-like spin-orbital set), create_SDSfUCC with every subset of {D, pD, GpD} (incl. the reference's pD entry that carries a
2-tuple, util.py:1042-1043), create_tiled with every combination of do_tups / do_qnp / skip_last_singles, and the UccStructure
builders.  Stored per case: options, types, indices, n_params, grad_param_R.
"""
from __future__ import annotations

import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from slowquant.unitary_coupled_cluster.util import UccStructure, UpsStructure  # noqa: E402

rng = np.random.default_rng(321)
SPACES = [
    # (num_active_orbs, occ, unocc, occ_spin, unocc_spin)
    (4, [0, 1], [2, 3], [0, 1, 2, 3], [4, 5, 6, 7]),
    (5, [0, 1, 2], [3, 4], [0, 1, 2, 3, 4, 5], [6, 7, 8, 9]),
    (4, [0, 1], [2, 3], [0, 1, 2], [3, 4, 5, 6, 7]),      # 2 alpha + 1 beta electrons: spin-orbital sets differ from 2 x spatial
]
FLAGS = ["S", "GS", "SAS", "SAGS", "D", "GD", "pD", "GpD", "T", "Q", "5", "6", "SAD"]


def dump(lay, extra):
    d = dict(extra)
    d["types"] = list(lay.excitation_operator_type)
    d["indices"] = [[int(x) for x in t] for t in lay.excitation_indices]
    d["n_params"] = int(lay.n_params)
    if hasattr(lay, "grad_param_R"):
        d["grad_param_R"] = {str(k): int(v) for k, v in dict(lay.grad_param_R).items()}
    return d


cases = []
for si, (n, occ, unocc, occs, unoccs) in enumerate(SPACES):
    combos = [[f] for f in FLAGS]
    for _ in range(14):
        k = int(rng.integers(2, 5))
        combos.append(sorted(rng.choice(FLAGS, size=k, replace=False).tolist(), key=FLAGS.index))
    for combo in combos:
        for L in (1, 2):
            if L == 2 and len(combo) == 1 and combo[0] in ("5", "6", "Q", "T"):
                continue
            opts = {"n_layers": L, **{f: True for f in combo}}
            lay = UpsStructure()
            try:
                lay.create_fUCC(occ, unocc, occs, unoccs, n, dict(opts))
            except Exception as e:  # noqa: BLE001
                cases.append({"kind": "fUCC", "space": si, "options": opts, "raises": type(e).__name__})
                continue
            cases.append(dump(lay, {"kind": "fUCC", "space": si, "options": opts}))
    for r in range(0, 4):
        for combo in itertools.combinations(["D", "pD", "GpD"], r):
            for L in (1, 2):
                opts = {"n_layers": L, **{f: True for f in combo}}
                lay = UpsStructure()
                try:
                    lay.create_SDSfUCC(occ, unocc, occs, unoccs, n, dict(opts))
                except Exception as e:  # noqa: BLE001
                    cases.append({"kind": "SDSfUCC", "space": si, "options": opts, "raises": type(e).__name__})
                    continue
                cases.append(dump(lay, {"kind": "SDSfUCC", "space": si, "options": opts}))
for n in (2, 3, 4, 5, 6, 7):
    for do_tups, do_qnp, skip in itertools.product((False, True), repeat=3):
        for L in (1, 3):
            opts = {"n_layers": L}
            if do_tups:
                opts["do_tups"] = True
            if do_qnp:
                opts["do_qnp"] = True
            if skip:
                opts["skip_last_singles"] = True
            lay = UpsStructure()
            try:
                lay.create_tiled(n, dict(opts))
            except Exception as e:  # noqa: BLE001
                cases.append({"kind": "tiled", "n": n, "options": opts, "raises": type(e).__name__})
                continue
            cases.append(dump(lay, {"kind": "tiled", "n": n, "options": opts}))
for si, (n, occ, unocc, occs, unoccs) in enumerate(SPACES):
    for builders in (["sa_singles"], ["sa_singles", "sa_doubles"], ["sa_doubles", "triples"], ["sa_singles", "sa_doubles", "triples", "quadruples"],
                     ["quintuples"], ["sextuples"]):
        lay = UccStructure()
        for b in builders:
            if b in ("sa_singles", "sa_doubles"):
                getattr(lay, "add_" + b)(occ, unocc)
            else:
                getattr(lay, "add_" + b)(occs, unoccs)
        cases.append(dump(lay, {"kind": "ucc", "space": si, "builders": builders}))

import gzip  # noqa: E402

with gzip.open(os.path.join(HERE, "golden_layouts.json.gz"), "wt", compresslevel=9) as f:
    json.dump({"spaces": SPACES, "cases": cases}, f)
print("wrote golden_layouts.json.gz:", len(cases), "cases,", sum("raises" in c for c in cases), "raising")
