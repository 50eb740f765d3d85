"""Adapter over integral providers (mirror of slowquant/unitary_coupled_cluster/integral_manager.py:9-125).

Accepts, by duck typing (no hard dependency on SlowQuant or PySCF):
  * a SlowQuant-like object (``.integral`` with ``kinetic_energy_matrix`` ..., ``.molecule``),
  * a PySCF-Mole-like object (``.intor``, ``.nelectron``),
  * :class:`ArrayIntegrals` for synthetic integrals given as arrays.
"""
from __future__ import annotations

import numpy as np


class ArrayIntegrals:
    """Integral provider from plain arrays (AO basis); used for synthetic benchmarks."""

    def __init__(self, h_ao: np.ndarray, eri_ao: np.ndarray, num_elec: int, nuclear_repulsion: float = 0.0, dipole=None):
        self.h_core = np.asarray(h_ao, dtype=np.float64)
        self.eri = np.asarray(eri_ao, dtype=np.float64)
        self.num_elec = int(num_elec)
        self.nuclear_repulsion = float(nuclear_repulsion)
        self.dipole = dipole


class IntegralManager:
    def __init__(self, integral_obj) -> None:
        self.int_obj = integral_obj
        self._cache: dict = {}

    def _kind(self) -> str:
        o = self.int_obj
        if isinstance(o, ArrayIntegrals):
            return "arrays"
        if hasattr(o, "integral") and hasattr(o, "molecule"):
            return "slowquant"
        if hasattr(o, "intor") and hasattr(o, "nelectron"):
            return "pyscf"
        raise ValueError(f"Got unknown integral object, {type(o)}")

    def _get(self, key: str, make):
        if key not in self._cache:
            self._cache[key] = make()
        return self._cache[key]

    @property
    def num_elec(self) -> int:
        k = self._kind()
        if k == "arrays":
            return self.int_obj.num_elec
        if k == "slowquant":
            return self.int_obj.molecule.number_electrons
        return self.int_obj.nelectron

    @property
    def kinetic_energy(self) -> np.ndarray:
        k = self._kind()
        if k == "arrays":
            return self._get("kin", lambda: np.zeros_like(self.int_obj.h_core))
        if k == "slowquant":
            return self._get("kin", lambda: self.int_obj.integral.kinetic_energy_matrix)
        return self._get("kin", lambda: self.int_obj.intor("int1e_kin"))

    @property
    def nuclear_electron_attraction(self) -> np.ndarray:
        k = self._kind()
        if k == "arrays":
            return self.int_obj.h_core
        if k == "slowquant":
            return self._get("nuc", lambda: self.int_obj.integral.nuclear_attraction_matrix)
        return self._get("nuc", lambda: self.int_obj.intor("int1e_nuc"))

    @property
    def electron_electron_repulsion(self) -> np.ndarray:
        k = self._kind()
        if k == "arrays":
            return self.int_obj.eri
        if k == "slowquant":
            return self._get("eri", lambda: self.int_obj.integral.electron_repulsion_tensor)
        return self._get("eri", lambda: self.int_obj.intor("int2e"))

    @property
    def nuclear_nuclear_repulsion(self) -> float:
        k = self._kind()
        if k == "arrays":
            return self.int_obj.nuclear_repulsion
        if k == "slowquant":
            return self.int_obj.molecule.nuclear_repulsion
        return self.int_obj.energy_nuc()

    @property
    def electric_dipole(self):
        k = self._kind()
        if k == "arrays":
            return self.int_obj.dipole
        if k == "slowquant":
            return self._get(
                "dip",
                lambda: tuple(self.int_obj.integral.get_multipole_matrix(np.array(v)) for v in ([1, 0, 0], [0, 1, 0], [0, 0, 1])),
            )
        return self._get("dip", lambda: tuple(self.int_obj.intor("int1e_r", comp=3)))

    @property
    def h_ao(self) -> np.ndarray:
        return self._get("h_ao", lambda: self.nuclear_electron_attraction + self.kinetic_energy)


def one_electron_integral_transform(C: np.ndarray, int1e: np.ndarray) -> np.ndarray:
    """AO -> MO for one-electron integrals (molecularintegrals/integralfunctions.py:186-196)."""
    return C.T @ int1e @ C


def two_electron_integral_transform(C: np.ndarray, int2e: np.ndarray) -> np.ndarray:
    """AO -> MO for two-electron integrals (molecularintegrals/integralfunctions.py:199-211)."""
    return np.einsum("ai,bj,ck,dl,abcd->ijkl", C, C, C, C, int2e, optimize=True)
