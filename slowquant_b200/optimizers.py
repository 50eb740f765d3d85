"""Optimizer front end with the reference's names (slowquant/unitary_coupled_cluster/optimizers.py): SciPy methods
by name and the RotoSolve sweep.  Host control flow only; every energy / gradient evaluation it asks for runs on the
device.  The trigonometric reconstruction is written as one [points x shifts] kernel matrix instead of the
reference's numba double loops.
"""
from __future__ import annotations

import time
from collections.abc import Callable, Sequence
from typing import Any

import numpy as np
import scipy.optimize


class Result:
    """x, fun, success (optimizers.py:11-19)."""

    x: np.ndarray
    fun: float
    success: bool


def _shift_points(R: int) -> np.ndarray:
    r""":math:`x_\mu = 2\mu\pi/(2R+1)`, :math:`\mu=-R..R`."""
    return 2 * np.arange(-R, R + 1) / (2 * R + 1) * np.pi


def _as_weights(energy_vals) -> tuple[np.ndarray, float]:
    """Per-shift weights of the reconstruction and its final divisor.  State-averaged input (one array per shift) is
    summed over states and divided by the number of SHIFTS, exactly as optimizers.py:331-343 does."""
    first = energy_vals[0]
    if isinstance(first, (float, np.floating)):
        return np.asarray(energy_vals, dtype=np.float64), 1.0
    return np.asarray([np.sum(e) for e in energy_vals], dtype=np.float64), float(len(energy_vals))


def reconstructed_f(x_vals: np.ndarray, energy_vals, R: int) -> np.ndarray:
    r""":math:`E(x)=\sum_\mu E(x_\mu)\,\mathrm{sinc}\big(\tfrac{2R+1}{2}(x-x_\mu)\big)/\mathrm{sinc}\big(\tfrac12(x-x_\mu)\big)`
    (optimizers.py:296-344; numpy's normalised sinc)."""
    w, div = _as_weights(energy_vals)
    delta = np.asarray(x_vals, dtype=np.float64)[:, None] - _shift_points(R)[None, :]
    kern = np.sinc((2 * R + 1) / 2 * delta / np.pi) / np.sinc(1 / 2 * delta / np.pi)
    return kern @ w / div


def _sinc_derivative(u: np.ndarray) -> np.ndarray:
    """d/du of numpy's sinc; 0 where |u| <= 1e-12 (optimizers.py:364-381)."""
    u = np.asarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    m = np.abs(u) > 1e-12
    um = u[m]
    out[m] = np.cos(np.pi * um) / um - np.sin(np.pi * um) / (np.pi * um**2)
    return out


def reconstructed_f_derivative(x_vals: np.ndarray, energy_vals, R: int) -> np.ndarray:
    """Derivative of reconstructed_f by the quotient rule (optimizers.py:384-492)."""
    w, div = _as_weights(energy_vals)
    A, B = (2 * R + 1) / 2.0, 0.5
    delta = np.asarray(x_vals, dtype=np.float64)[:, None] - _shift_points(R)[None, :]
    u, v = A * delta / np.pi, B * delta / np.pi
    s1, s2 = np.sinc(u), np.sinc(v)
    s1p, s2p = _sinc_derivative(u) * (A / np.pi), _sinc_derivative(v) * (B / np.pi)
    return ((s1p * s2 - s1 * s2p) / s2**2) @ w / div


def get_energy_evals(f: Callable[[list[float]], float | np.ndarray], x: list[float], idx: int, R: int) -> list:
    """f at the 2R+1 shifted values of parameter idx (optimizers.py:273-293)."""
    x = list(x)
    vals = []
    for x_mu in _shift_points(R):
        x[idx] = float(x_mu)
        vals.append(f(x))
    return vals


def get_energy_evals_optimized(f: Callable[[list[float], list[float], int], list[float]], x: list[float], idx: int, R: int) -> list[float]:
    """Same through a batched evaluator f(x, shifts, idx) (optimizers.py:347-361)."""
    return f(x, [float(t) for t in _shift_points(R)], idx)


class RotoSolve:
    """Coordinate-wise exact minimisation of the trigonometric energy curve (optimizers.py:166-270)."""

    def __init__(
        self,
        R: dict[str, int],
        param_names: Sequence[str],
        maxiter: int = 1000,
        tol: float = 1e-8,
        callback: Callable[[list[float]], None] | None = None,
    ) -> None:
        self._callback = callback
        self.max_iterations = maxiter
        self.threshold = tol
        self.max_fail = 6
        self._R = R
        self._param_names = param_names

    def minimize(self, f, x0: Sequence[float], f_rotosolve_optimized=None) -> Result:
        f_best = float(10**20)
        x = list(x0)
        x_best = list(x)
        fails = 0
        res = Result()
        success = False
        grid = np.linspace(-np.pi, np.pi, int(1e4))
        for _ in range(self.max_iterations):
            for i, par_name in enumerate(self._param_names):
                R = self._R[par_name]
                if f_rotosolve_optimized is not None:
                    e_vals = get_energy_evals_optimized(f_rotosolve_optimized, x, i, R)
                else:
                    e_vals = get_energy_evals(f, x, i, R)
                theta = grid[np.argmin(reconstructed_f(grid, e_vals, R))]
                fine = scipy.optimize.minimize(
                    lambda t: reconstructed_f(t, e_vals, R)[0],
                    x0=[theta],
                    jac=lambda t: reconstructed_f_derivative(t, e_vals, R),
                    method="BFGS",
                    tol=1e-12,
                )
                x[i] = float(fine.x[0])
                # wrap to (-pi, pi] with the reference's two loops (optimizers.py:244-247)
                while x[i] < np.pi:
                    x[i] += 2 * np.pi
                while x[i] > np.pi:
                    x[i] -= 2 * np.pi
            f_tmp = f(x)
            f_new = float(np.mean(f_tmp)) if isinstance(f_tmp, np.ndarray) else f_tmp
            if self._callback is not None:
                self._callback(x)
            if abs(f_best - f_new) < self.threshold:
                f_best = f_new
                x_best = list(x)
                success = True
                break
            if (f_new - f_best) > 0.0:
                fails += 1
            else:
                f_best = f_new
                x_best = list(x)
            if fails == self.max_fail:
                print("Three energy raises detected.")
                break
        res.x = np.array(x_best)
        res.fun = f_best
        res.success = success
        return res


class Optimizers:
    """Method dispatch by name (optimizers.py:22-163)."""

    def __init__(
        self,
        fun: Callable[[list[float]], float | np.ndarray],
        method: str,
        grad: Callable[[list[float]], np.ndarray] | None = None,
        maxiter: int = 1000,
        tol: float = 10e-8,
        is_silent: bool = False,
        energy_eval_callback: Callable[[], int] | None = None,
        std_callback: Callable[[], float] | None = None,
    ) -> None:
        self.fun = fun
        self.grad = grad
        self.method = method.lower()
        self.maxiter = maxiter
        self.tol = tol
        self.is_silent = is_silent
        self.energy_eval_callback = energy_eval_callback
        self.std_callback = std_callback
        self._start = 0.0
        self._iteration = 0

    def _print_progress(self, x: Sequence[float]) -> None:
        if self.is_silent:
            return
        e = self.fun(list(x))
        e_str = f"{np.mean(e):3.16f}" if isinstance(e, np.ndarray) else f"{e:3.16f}"
        time_str = f"{time.time() - self._start:7.2f}"
        evals_str = str(self.energy_eval_callback()) if self.energy_eval_callback else ""
        std_str = ""
        if self.std_callback is not None:
            var = self.std_callback()
            if var is not None:
                std_str = f" | {np.sqrt(var):.6e}"
        print(f"--------{str(self._iteration + 1).center(11)} | {time_str.center(18)} | {e_str.center(27)} | {evals_str.center(20)}{std_str}")
        self._iteration += 1
        self._start = time.time()

    def minimize(self, x0: Sequence[float], extra_options: dict[str, Any] | None = None) -> Result:
        self._start = time.time()
        self._iteration = 0
        if self.method in ("bfgs", "l-bfgs-b", "slsqp"):
            res = scipy.optimize.minimize(
                self.fun, x0, jac=self.grad, method=self.method, tol=self.tol, callback=self._print_progress,
                options={"maxiter": self.maxiter},
            )
        elif self.method in ("cobyla", "cobyqa"):
            res = scipy.optimize.minimize(
                self.fun, x0, method=self.method, tol=self.tol, callback=self._print_progress, options={"maxiter": self.maxiter}
            )
        elif self.method == "rotosolve":
            if not isinstance(extra_options, dict):
                raise TypeError("extra_options is not set, but is required for RotoSolve")
            if "R" not in extra_options:
                raise ValueError(f"Expected option 'R' in extra_options, got {extra_options.keys()}")
            if "param_names" not in extra_options:
                raise ValueError(f"Expected option 'param_names' in extra_options, got {extra_options.keys()}")
            optimizer = RotoSolve(
                extra_options["R"], extra_options["param_names"], maxiter=self.maxiter, tol=self.tol, callback=self._print_progress
            )
            res = optimizer.minimize(self.fun, x0, f_rotosolve_optimized=extra_options.get("f_rotosolve_optimized"))
        else:
            raise ValueError(f"Got an unkonwn optimizer {self.method}")
        result = Result()
        result.x = res.x
        result.fun = res.fun
        result.success = res.success
        if not result.success:
            print("Optimization failed.")
            if hasattr(res, "message"):
                print(res.message)
        return result
