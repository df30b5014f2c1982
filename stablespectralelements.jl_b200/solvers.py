"""Solver front end: the reference's ``Solvers`` API surface over the B200 C-ABI.

Mirrors /root/reference/src/Solvers/Solvers.jl: residual forms (:82-115), strategies (:75-78),
mass-matrix solvers (mass_matrix.jl), ``Solver`` (:259-272, constructors :287-377),
``initialize``/``project_function`` (:389-428), ``semidiscretize`` (:430-454) and
``semi_discrete_residual!`` (:455-570).  The parallelism tag is always the device
(``B200``, alongside the reference's ``Serial``/``Threaded``): the residual is evaluated by
the CUDA kernels behind ``libsse_b200.so``; there is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import conservation_laws as cl
from . import grid_functions as gfn
from .mesh import run_chunks
from .geometric_factors import SpatialDiscretization, apply_reference_mapping
from .linear_maps import (IdentityMap, SelectionMap, WarpedTensorProductMap2D,
                          WarpedTensorProductMap3D)
from .reference_approximation import NoMapping, reference_derivative_operators


# ---- dispatch tags (Solvers.jl:67-80) ----------------------------------------------
class StandardMapping:
    pass


class SkewSymmetricMapping:
    pass


class PhysicalOperator:
    pass


class ReferenceOperator:
    pass


class B200:
    """Parallelism tag: evaluate on the device through the C-ABI (no Serial/Threaded here)."""


@dataclass
class StandardForm:
    """Solvers.jl:82-94."""
    mapping_form: object = field(default_factory=SkewSymmetricMapping)
    inviscid_numerical_flux: object = field(default_factory=cl.LaxFriedrichsNumericalFlux)
    viscous_numerical_flux: object = field(default_factory=cl.BR1)


@dataclass
class FluxDifferencingForm:
    """Solvers.jl:96-115."""
    mapping_form: object = field(default_factory=SkewSymmetricMapping)
    inviscid_numerical_flux: object = field(default_factory=cl.LaxFriedrichsNumericalFlux)
    viscous_numerical_flux: object = field(default_factory=cl.BR1)
    two_point_flux: object = field(default_factory=cl.EntropyConservativeFlux)


# ---- mass-matrix solvers (mass_matrix.jl:1-136) --------------------------------------
class CholeskySolver:
    kind = "cholesky"


class DiagonalSolver:
    kind = "diagonal"


class WeightAdjustedSolver:
    """mass_matrix.jl:59-115.  ``assume_orthonormal=True`` -> M⁻¹ = I."""
    kind = "weight_adjusted"

    def __init__(self, assume_orthonormal: bool = True, tol: float = 1.0e-13):
        self.assume_orthonormal = assume_orthonormal
        self.tol = tol


def default_mass_matrix_solver(sd: SpatialDiscretization):
    """mass_matrix.jl:19-24."""
    return WeightAdjustedSolver(True, 0.0)


def _resolve_mass_solver(ms, ra):
    """Nodal schemes (V = I) always collapse to the diagonal solver (mass_matrix.jl:26-28,
    41-57)."""
    if isinstance(ra.V, IdentityMap):
        return "diagonal", None
    if isinstance(ms, DiagonalSolver):
        raise ValueError("DiagonalSolver requires a nodal scheme (V = I)")
    if isinstance(ms, CholeskySolver):
        return "cholesky", None
    if ms.assume_orthonormal:
        return "weight_adjusted", None
    VDM = ra.V.to_dense()
    M = VDM.T @ (ra.W[:, None] * VDM)
    Md = np.diag(M)
    if np.max(np.abs(M - np.diag(Md))) < ms.tol:
        if np.max(np.abs(Md - 1.0)) < ms.tol:
            return "weight_adjusted", None
        return "weight_adjusted", np.diag(1.0 / Md)
    return "weight_adjusted", np.linalg.inv(M)


# ---- law / form descriptions shared by the C-ABI packer and the test bridge -----------
def describe_law(law) -> dict:
    if isinstance(law, cl.LinearAdvectionEquation):
        return dict(kind="advection", d=law.d, N_c=1, a=tuple(law.a))
    if isinstance(law, cl.LinearAdvectionDiffusionEquation):
        return dict(kind="advection_diffusion", d=law.d, N_c=1, a=tuple(law.a), b=law.b)
    if isinstance(law, cl.InviscidBurgersEquation):
        return dict(kind="burgers", d=law.d, N_c=1, a=tuple(law.a))
    if isinstance(law, cl.ViscousBurgersEquation):
        return dict(kind="viscous_burgers", d=law.d, N_c=1, a=tuple(law.a), b=law.b)
    if isinstance(law, cl.EulerEquations):
        return dict(kind="euler", d=law.d, N_c=law.d + 2, gamma=law.gamma)
    raise TypeError(law)


def describe_form(form, strategy, law) -> dict:
    inv = form.inviscid_numerical_flux
    if isinstance(inv, cl.LaxFriedrichsNumericalFlux):
        inviscid = ("lf", inv.half_lambda)
    elif isinstance(inv, cl.CentralNumericalFlux):
        inviscid = ("central",)
    elif isinstance(inv, cl.EntropyConservativeNumericalFlux):
        inviscid = ("ec",)
    else:
        raise TypeError(inv)
    second = law.pde_type is cl.SecondOrder
    out = dict(inviscid=inviscid,
               mapping_form="standard" if isinstance(form.mapping_form, StandardMapping)
               else "skew")
    if isinstance(form, FluxDifferencingForm):
        if second:
            raise ValueError("no flux-differencing form for second-order equations")
        out.update(kind="flux_differencing", strategy="reference",
                   two_point="ec" if isinstance(form.two_point_flux,
                                                cl.EntropyConservativeFlux) else "conservative")
    else:
        # second-order equations always use physical operators (Solvers.jl:357-377)
        phys = second or isinstance(strategy, PhysicalOperator)
        out.update(kind="standard", strategy="physical" if phys else "reference",
                   two_point="conservative")
    return out


def get_dof(sd: SpatialDiscretization, law) -> Tuple[int, int, int]:
    """Solvers.jl:379-387."""
    return sd.reference_approximation.N_p, law.N_c, sd.N_e


# ---- initial data (Solvers.jl:389-428) ------------------------------------------------
def project_function(initial_data, sd: SpatialDiscretization) -> np.ndarray:
    """Nodal: point values; modal: per-element L2 projection (VᵀWJV) \\ VᵀWJ u_q.
    Returns u0 as (N_e, N_c, N_p) -- memory-identical to Julia's (N_p, N_c, N_e).
    Elements are independent: chunks of them run on a thread pool (NumPy releases the GIL)."""
    ra = sd.reference_approximation
    nodal = isinstance(ra.V, IdentityMap)
    V = None if nodal else ra.V.to_dense()
    J_q = sd.geometric_factors.J_q
    xyzq = sd.mesh.xyzq                                       # d arrays (N_q, N_e)
    N_e = sd.N_e
    probe = gfn.evaluate(initial_data, tuple(x[:, :1].T for x in xyzq), 0.0)
    N_c = probe.shape[0]
    u0 = np.empty((N_e, N_c, ra.N_q if nodal else ra.N_p))

    def work(s, e):
        xq = tuple(x[:, s:e].T for x in xyzq)                 # (C, N_q)
        u_q = gfn.evaluate(initial_data, xq, 0.0)             # (N_c, C, N_q)
        if nodal:
            u0[s:e] = u_q.transpose(1, 0, 2)
            return
        VW = V.T[None, :, :] * (ra.W[None, :] * J_q[s:e])[:, None, :]   # (C, N_p, N_q)
        M = VW @ V
        rhs = VW @ u_q.transpose(1, 2, 0)                               # (C, N_p, N_c)
        u0[s:e] = np.linalg.solve(M, rhs).transpose(0, 2, 1)

    run_chunks(work, N_e, 8192)
    return u0


initialize = project_function


# ---- Solver -------------------------------------------------------------------------
class Solver:
    """Solvers.jl:259-377: bundles the conservation law, operators, mass solver,
    connectivity and form; here it owns the device handle created by ``sse_create``."""

    def __init__(self, conservation_law, spatial_discretization: SpatialDiscretization, form,
                 strategy=None, mass_solver=None, parallelism=None, device: int = 0,
                 lazy: bool = False):
        self.conservation_law = conservation_law
        self.spatial_discretization = spatial_discretization
        self.form = form
        self.strategy = strategy or ReferenceOperator()
        self.mass_solver = mass_solver or default_mass_matrix_solver(spatial_discretization)
        self.parallelism = parallelism or B200()
        self.connectivity = spatial_discretization.mesh.mapP
        self.law_desc = describe_law(conservation_law)
        self.form_desc = describe_form(form, self.strategy, conservation_law)
        ra = spatial_discretization.reference_approximation
        if ra.dim != self.law_desc["d"]:
            raise ValueError("dimension mismatch between conservation law and discretization")
        self.mass_kind, self.Minv = _resolve_mass_solver(self.mass_solver, ra)
        self.device = device
        self._handle = None
        if not lazy:
            self.handle  # create now

    # Base.size(solver) (Solvers.jl:274-285)
    def size(self):
        return get_dof(self.spatial_discretization, self.conservation_law)

    @property
    def handle(self):
        if self._handle is None:
            from .device import DeviceResidual
            self._handle = DeviceResidual(self, device=self.device)
        return self._handle

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None


@dataclass
class ODEProblem:
    """Stand-in for OrdinaryDiffEq.ODEProblem(f, u0, tspan, p) (Solvers.jl:453)."""
    f: object
    u0: np.ndarray
    tspan: Tuple[float, float]
    p: Solver


def semi_discrete_residual(dudt: np.ndarray, u: np.ndarray, solver: Solver, t: float = 0.0):
    """``semi_discrete_residual!(dudt, u, solver, t)`` (Solvers.jl:455-570) with host arrays:
    copies ``u`` to the device, evaluates the residual there and copies ``dudt`` back."""
    solver.handle.residual_host(u, dudt, t)
    return dudt


def semidiscretize(conservation_law, spatial_discretization, initial_data, form, tspan,
                   strategy=None, mass_matrix_solver=None, device: int = 0) -> ODEProblem:
    """Solvers.jl:430-454."""
    u0 = initialize(initial_data, spatial_discretization)
    solver = Solver(conservation_law, spatial_discretization, form, strategy,
                    mass_matrix_solver, device=device)
    return ODEProblem(semi_discrete_residual, u0, (float(tspan[0]), float(tspan[1])), solver)
