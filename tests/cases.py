"""Shared problem builders for the parity tests, smoke() and bench.py (BASELINE.json configs)."""
import math

import numpy as np

from sse_b200.conservation_laws import (BR1, CentralNumericalFlux,
                                        EntropyConservativeNumericalFlux, EulerEquations,
                                        InviscidBurgersEquation, LaxFriedrichsNumericalFlux,
                                        LinearAdvectionDiffusionEquation, LinearAdvectionEquation)
from sse_b200.geometric_factors import (ChanWilcoxMetrics, ExactMetrics,
                                        make_spatial_discretization)
from sse_b200.grid_functions import (EulerPeriodicTest, InitialDataSine, IsentropicVortex,
                                     TaylorGreenVortex)
from sse_b200.mesh import ChanWarping, DelReyWarping, uniform_periodic_mesh, warp_mesh
from sse_b200.reference_approximation import (Hex, Line, ModalMulti, ModalTensor, NodalTensor,
                                              Quad, Tet, Tri, make_reference_approximation)
from sse_b200.solvers import (FluxDifferencingForm, PhysicalOperator, ReferenceOperator, Solver,
                              StandardForm, StandardMapping, project_function)


def rough_state(solver, u0, seed=0, amp=0.05):
    """Deterministic rough perturbation of a smooth state (keeps Euler states admissible):
    exercises both logmean branches."""
    rng = np.random.default_rng(seed)
    u = u0 * (1.0 + amp * (rng.random(u0.shape) - 0.5))
    return np.ascontiguousarray(u)


def advection_tri_case(p=4, M=4, lazy=True, warp=0.2, lam=1.0):
    """BASELINE config 1: 2-D advection, curved triangles, StandardForm + ReferenceOperator."""
    law = LinearAdvectionEquation((1.0, 1.0))
    ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M)), ra, warp)
    sd = make_spatial_discretization(mesh, ra)
    solver = Solver(law, sd, StandardForm(inviscid_numerical_flux=LaxFriedrichsNumericalFlux(lam)),
                    ReferenceOperator(), lazy=lazy)
    return solver, project_function(InitialDataSine(1.0, (2 * math.pi,) * 2), sd)


def euler_tri_case(p=4, M=4, lazy=True, interface="lf", approx="modal"):
    """BASELINE config 2: 2-D Euler isentropic vortex, flux differencing (scaling_test_euler_2d)."""
    g = 1.4
    law = EulerEquations(2, g)
    ic = IsentropicVortex(gamma=g, Ma=0.4, theta=0.0, R=0.1,
                          beta=math.sqrt(2 / (g - 1) * (1 - 0.75 ** (g - 1))), x_0=(0.5, 0.5))
    at = ModalTensor(p) if approx == "modal" else NodalTensor(p)
    ra = make_reference_approximation(at, Tri(), mapping_degree=p)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M)), ra,
                     ChanWarping(1 / 16, (1.0, 1.0)))
    sd = make_spatial_discretization(mesh, ra)
    flux = LaxFriedrichsNumericalFlux() if interface == "lf" else EntropyConservativeNumericalFlux()
    solver = Solver(law, sd, FluxDifferencingForm(inviscid_numerical_flux=flux),
                    ReferenceOperator(), lazy=lazy)
    return solver, project_function(ic, sd)


def advection_tet_case(p=4, M=2, lazy=True, warp=0.1, mapping_degree=None):
    """BASELINE config 3: 3-D advection on curved tetrahedra, StandardForm + ReferenceOperator."""
    law = LinearAdvectionEquation((1.0, 1.0, 1.0))
    md = p if mapping_degree is None else mapping_degree
    ra = make_reference_approximation(ModalTensor(p), Tet(), mapping_degree=md)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 3, (M,) * 3)
    if warp:
        mesh = warp_mesh(mesh, ra, warp)
    sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics())
    solver = Solver(law, sd, StandardForm(), ReferenceOperator(), lazy=lazy)
    from sse_b200.grid_functions import InitialDataCosine
    return solver, project_function(InitialDataCosine(1.0, (2 * math.pi,) * 3), sd)


def euler_tet_case(p=4, M=2, lazy=True, warp=False, interface="lf", ic="tgv",
                   approx="modal", shard=None, device_geometry=None):
    """BASELINE config 4 (north star): 3-D Euler Taylor-Green vortex on tetrahedra, flux
    differencing, entropy-conservative two-point flux, LF or EC interface flux."""
    g = 1.4
    law = EulerEquations(3, g)
    L = 2 * math.pi
    at = ModalTensor(p) if approx == "modal" else NodalTensor(p)
    ra = make_reference_approximation(at, Tet(), mapping_degree=(min(p, 3) if warp else 1))
    mesh = uniform_periodic_mesh(ra, ((0.0, L),) * 3, (M,) * 3)
    if shard is not None:      # (rank, world): keep only this rank's elements from here on
        from sse_b200.distributed import element_ranges
        from sse_b200.mesh import mesh_subset
        mesh = mesh_subset(mesh, *element_ranges(mesh.N_e, shard[1])[shard[0]])
    if warp:
        mesh = warp_mesh(mesh, ra, ChanWarping(1 / 16, (L, L, L)))
        sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics(),
                                         device_geometry=device_geometry)
    else:
        sd = make_spatial_discretization(mesh, ra, device_geometry=device_geometry)
    flux = LaxFriedrichsNumericalFlux() if interface == "lf" else EntropyConservativeNumericalFlux()
    solver = Solver(law, sd, FluxDifferencingForm(inviscid_numerical_flux=flux),
                    ReferenceOperator(), lazy=lazy)
    data = TaylorGreenVortex(gamma=g, Ma=0.1) if ic == "tgv" else EulerPeriodicTest(3, g, 0.2, L)
    return solver, project_function(data, sd)


def advection_diffusion_case(d=1, p=4, M=4, lazy=True):
    """BASELINE config 5: advection-diffusion with BR1, PhysicalOperator."""
    if d == 1:
        law = LinearAdvectionDiffusionEquation((1.0,), 5.0e-2)
        ra = make_reference_approximation(ModalMulti(p), Line())
        sd = make_spatial_discretization(uniform_periodic_mesh(ra, (0.0, 1.0), M), ra)
    else:
        law = LinearAdvectionDiffusionEquation((1.0, 1.0), 5.0e-2)
        ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p)
        mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M)), ra, 0.1)
        sd = make_spatial_discretization(mesh, ra)
    form = StandardForm(StandardMapping() if d == 1 else None or StandardMapping(),
                        LaxFriedrichsNumericalFlux(), BR1())
    solver = Solver(law, sd, form, PhysicalOperator(), lazy=lazy)
    return solver, project_function(InitialDataSine(1.0, (2 * math.pi,) * d), sd)


def euler_hex_case(p=3, M=2, lazy=True, interface="ec"):
    """SURVEY §8(f) item 4 / runtests.jl:131-144: 3-D Euler on curved hexahedra, NodalTensor LGL
    collocation (diag-E: SelectionMap R, no facet correction), conservative-curl metrics."""
    g = 1.4
    law = EulerEquations(3, g)
    L = 2.0
    ra = make_reference_approximation(NodalTensor(p), Hex(), mapping_degree=p)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, L),) * 3, (M,) * 3), ra,
                     ChanWarping(1 / 16, (L, L, L)))
    sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics())
    flux = LaxFriedrichsNumericalFlux() if interface == "lf" else EntropyConservativeNumericalFlux()
    solver = Solver(law, sd, FluxDifferencingForm(inviscid_numerical_flux=flux),
                    ReferenceOperator(), lazy=lazy)
    return solver, project_function(EulerPeriodicTest(3, g, 0.2, L), sd)


def burgers_tri_case(p=3, M=3, lazy=True):
    """2-D inviscid Burgers, flux differencing with the EC flux on curved triangles."""
    law = InviscidBurgersEquation((1.0, 0.5))
    ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M)), ra, 0.1)
    sd = make_spatial_discretization(mesh, ra)
    form = FluxDifferencingForm(inviscid_numerical_flux=EntropyConservativeNumericalFlux())
    solver = Solver(law, sd, form, ReferenceOperator(), lazy=lazy)
    u0 = project_function(InitialDataSine(1.0, (2 * math.pi,) * 2), sd)
    return solver, u0 + 1.5
