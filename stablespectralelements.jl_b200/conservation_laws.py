"""Conservation-law, numerical-flux and two-point-flux *types* (host side).

Mirrors the type surface of /root/reference/src/ConservationLaws/ConservationLaws.jl:39-71,
linear_advection_diffusion.jl:1-52, burgers.jl:1-50 and euler_navierstokes.jl:26-37.  The
pointwise physics itself (fluxes, entropy maps, wave speeds) runs in CUDA (csrc/physics.cuh);
these objects only select it through the C-ABI config.  ``source_term`` exists on the
reference's laws but is never applied by any residual, so it is not carried here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


# ---- PDE types ------------------------------------------------------------------
class FirstOrder:
    pass


class SecondOrder:
    pass


@dataclass(frozen=True)
class LinearAdvectionEquation:
    a: Tuple[float, ...]
    pde_type = FirstOrder
    N_c = 1

    def __post_init__(self):
        if not isinstance(self.a, tuple):
            object.__setattr__(self, "a", (float(self.a),))

    @property
    def d(self):
        return len(self.a)


@dataclass(frozen=True)
class LinearAdvectionDiffusionEquation:
    a: Tuple[float, ...]
    b: float
    pde_type = SecondOrder
    N_c = 1

    def __post_init__(self):
        if not isinstance(self.a, tuple):
            object.__setattr__(self, "a", (float(self.a),))

    @property
    def d(self):
        return len(self.a)


@dataclass(frozen=True)
class InviscidBurgersEquation:
    a: Tuple[float, ...] = (1.0,)
    pde_type = FirstOrder
    N_c = 1

    @property
    def d(self):
        return len(self.a)


class ViscousBurgersEquation:
    """``ViscousBurgersEquation(a, b)`` / ``ViscousBurgersEquation(b)`` with a = (1.0,)
    (burgers.jl:20-48)."""
    pde_type = SecondOrder
    N_c = 1

    def __init__(self, a, b=None):
        if b is None:
            a, b = (1.0,), a
        self.a = tuple(float(x) for x in a)
        self.b = float(b)

    @property
    def d(self):
        return len(self.a)

    def __repr__(self):
        return f"ViscousBurgersEquation(a={self.a}, b={self.b})"


@dataclass(frozen=True)
class EulerEquations:
    d: int
    gamma: float = 1.4
    pde_type = FirstOrder

    @property
    def N_c(self):
        return self.d + 2


# ---- interface fluxes (ConservationLaws.jl:50-71) ---------------------------------
class NoInviscidFlux:
    pass


class LaxFriedrichsNumericalFlux:
    """Stores halfλ = λ/2 like the reference (ConservationLaws.jl:52-60)."""

    def __init__(self, lam: float = 1.0):
        self.half_lambda = 0.5 * lam


class EntropyConservativeNumericalFlux:
    pass


class CentralNumericalFlux:
    pass


class BR1:
    pass


class NoViscousFlux:
    pass


class ConservativeFlux:
    pass


class EntropyConservativeFlux:
    pass
