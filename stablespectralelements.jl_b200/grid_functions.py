"""Initial-data functors (host side), vectorised over nodes.

Mirrors /root/reference/src/GridFunctions/GridFunctions.jl:21-201 and the Euler test states of
/root/reference/src/ConservationLaws/euler_navierstokes.jl:234-348.  ``evaluate(f, x, t)``
takes a tuple of coordinate arrays of identical shape S and returns an array (N_c,) + S.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Tuple

import numpy as np


@dataclass(frozen=True)
class InitialDataSine:
    A: float
    k: Tuple[float, ...]
    N_c: int = 1


@dataclass(frozen=True)
class InitialDataCosine:
    A: float
    k: Tuple[float, ...]
    N_c: int = 1


@dataclass(frozen=True)
class InitialDataGaussian:
    A: float
    sigma: float
    x0: Tuple[float, ...]
    N_c: int = 1


@dataclass(frozen=True)
class InitialDataGassner:
    k: float
    eps: float
    N_c: int = 1


@dataclass(frozen=True)
class ConstantFunction:
    c: float
    N_c: int = 1


@dataclass(frozen=True)
class IsentropicVortex:
    gamma: float = 1.4
    Ma: float = 0.4
    theta: float = math.pi / 4
    R: float = 1.0
    beta: float = 1.0
    sigma: float = 1.0
    x_0: Tuple[float, float] = (0.0, 0.0)
    N_c: int = 4


@dataclass(frozen=True)
class EulerPeriodicTest:
    d: int
    gamma: float = 1.4
    strength: float = 0.2
    L: float = 2.0

    @property
    def N_c(self):
        return self.d + 2


@dataclass(frozen=True)
class TaylorGreenVortex:
    gamma: float = 1.4
    Ma: float = 0.1
    N_c: int = 5


@dataclass(frozen=True)
class KelvinHelmholtzInstability:
    gamma: float = 1.4
    rho_0: float = 0.5
    N_c: int = 4


def evaluate(f, x, t: float = 0.0) -> np.ndarray:
    x = tuple(np.asarray(c, dtype=np.float64) for c in x)
    d = len(x)
    if isinstance(f, InitialDataSine):
        v = f.A * np.prod([np.sin(f.k[m] * x[m]) for m in range(d)], axis=0)
        return np.stack([v] * f.N_c)
    if isinstance(f, InitialDataCosine):
        v = f.A * np.prod([np.cos(f.k[m] * x[m]) for m in range(d)], axis=0)
        return np.stack([v] * f.N_c)
    if isinstance(f, InitialDataGaussian):
        r2 = sum((x[m] - f.x0[m]) ** 2 for m in range(d))
        return np.stack([f.A * np.exp(-r2 / (2.0 * f.sigma ** 2))] * f.N_c)
    if isinstance(f, InitialDataGassner):
        return np.stack([np.sin(f.k * x[0]) + f.eps])
    if isinstance(f, ConstantFunction):
        return np.stack([np.full_like(x[0], f.c)] * f.N_c)
    if isinstance(f, IsentropicVortex):
        g = f.gamma
        xr = ((x[0] - f.x_0[0]) / f.R, (x[1] - f.x_0[1]) / f.R)
        Om = f.beta * np.exp(-0.5 / f.sigma ** 2 * (xr[0] ** 2 + xr[1] ** 2))
        dv = (-xr[1] * Om, xr[0] * Om)
        dT = -0.5 * (g - 1) * Om ** 2
        rho = (1 + dT) ** (1 / (g - 1))
        v = (f.Ma * math.cos(f.theta) + dv[0], f.Ma * math.sin(f.theta) + dv[1])
        p = rho ** g / g
        E = p / (g - 1) + 0.5 * rho * (v[0] ** 2 + v[1] ** 2)
        return np.stack([rho, rho * v[0], rho * v[1], E])
    if isinstance(f, EulerPeriodicTest):
        rho = 1.0 + f.strength * np.sin(2 * math.pi * sum(x) / f.L)
        return np.stack([rho] + [rho] * d + [1.0 / (f.gamma - 1.0) + 0.5 * rho * d])
    if isinstance(f, TaylorGreenVortex):
        p = (1 / (f.Ma ** 2 * f.gamma)) + 0.0625 * (
            2 * np.cos(2 * x[0]) + 2 * np.cos(2 * x[1]) + np.cos(2 * x[0]) * np.cos(2 * x[2])
            + np.cos(2 * x[1]) * np.cos(2 * x[2]))
        u = np.sin(x[0]) * np.cos(x[1]) * np.cos(x[2])
        v = -np.cos(x[0]) * np.sin(x[1]) * np.cos(x[2])
        return np.stack([np.ones_like(u), u, v, np.zeros_like(u),
                         p / (f.gamma - 1) + 0.5 * (u ** 2 + v ** 2)])
    if isinstance(f, KelvinHelmholtzInstability):
        xr = (x[0] - 1, x[1] - 1)
        B = np.tanh(15 * xr[1] + 7.5) - np.tanh(15 * xr[1] - 7.5)
        rho = f.rho_0 + 0.75 * B
        u = 0.5 * (B - 1)
        v = 0.1 * np.sin(2 * math.pi * xr[0])
        return np.stack([rho, rho * u, rho * v, 1.0 / (f.gamma - 1) + 0.5 * rho * (u ** 2 + v ** 2)])
    if callable(f):
        out = f(*x, t)
        return np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), x[0].shape)
                         for c in out])
    raise TypeError(f"cannot evaluate {f!r}")
