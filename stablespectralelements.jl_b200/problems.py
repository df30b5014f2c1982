"""Problem builders for the BASELINE.json configurations and the dispatch branches of the
residual (used by bench.py, smoke() and, through tests/cases.py, the parity tests).  Each one
returns ``(solver, u0)`` the way ``semidiscretize`` of the reference's drivers would
(test/test_driver.jl, examples/*.ipynb)."""
import math

import numpy as np

from .conservation_laws import (BR1, CentralNumericalFlux,
                                        EntropyConservativeNumericalFlux, EulerEquations,
                                        InviscidBurgersEquation, LaxFriedrichsNumericalFlux,
                                        LinearAdvectionDiffusionEquation, LinearAdvectionEquation)
from .geometric_factors import (ChanWilcoxMetrics, ExactMetrics,
                                        make_spatial_discretization)
from .grid_functions import (EulerPeriodicTest, InitialDataSine, IsentropicVortex,
                                     TaylorGreenVortex)
from .mesh import ChanWarping, DelReyWarping, uniform_periodic_mesh, warp_mesh
from .reference_approximation import (Hex, Line, ModalMulti, ModalTensor, NodalTensor,
                                              Quad, Tet, Tri, make_reference_approximation)
from .solvers import (FluxDifferencingForm, PhysicalOperator, ReferenceOperator, Solver,
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


def advection_tet_case(p=4, M=2, lazy=True, warp=0.1, mapping_degree=None, shard=None):
    """BASELINE config 3: 3-D advection on curved tetrahedra, StandardForm + ReferenceOperator."""
    law = LinearAdvectionEquation((1.0, 1.0, 1.0))
    md = p if mapping_degree is None else mapping_degree
    ra = make_reference_approximation(ModalTensor(p), Tet(), mapping_degree=md)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 3, (M,) * 3)
    if shard is not None:      # (rank, world): keep only this rank's elements from here on
        from .distributed import element_ranges
        from .mesh import mesh_subset
        mesh = mesh_subset(mesh, *element_ranges(mesh.N_e, shard[1])[shard[0]])
    if warp:
        mesh = warp_mesh(mesh, ra, warp)
    sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics())
    solver = Solver(law, sd, StandardForm(), ReferenceOperator(), lazy=lazy)
    from .grid_functions import InitialDataCosine
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
        from .distributed import element_ranges
        from .mesh import mesh_subset
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


# ---- dispatch branches the round-1 suite did not reach (VERDICT rows a7, a12, a14, a15) ----
def advection_physical_case(d=2, p=3, M=3, lazy=True, mapping="skew", lam=1.0):
    """First-order law with PhysicalOperator (standard_form_first_order.jl:65-94): per-element
    dense VOL/FAC, skew-symmetric or standard mapping form (operators.jl:85-164)."""
    from .solvers import SkewSymmetricMapping
    mf = SkewSymmetricMapping() if mapping == "skew" else StandardMapping()
    if d == 1:
        law = LinearAdvectionEquation((1.0,))
        ra = make_reference_approximation(ModalMulti(p), Line())
        sd = make_spatial_discretization(uniform_periodic_mesh(ra, (0.0, 1.0), M), ra)
    else:
        law = LinearAdvectionEquation((1.0, 0.5))
        ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p)
        mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M)), ra, 0.1)
        sd = make_spatial_discretization(mesh, ra)
    form = StandardForm(mf, LaxFriedrichsNumericalFlux(lam))
    solver = Solver(law, sd, form, PhysicalOperator(), lazy=lazy)
    return solver, project_function(InitialDataSine(1.0, (2 * math.pi,) * d), sd)


def burgers_physical_case(p=3, M=3, lazy=True):
    """Inviscid Burgers, StandardForm + PhysicalOperator on curved triangles (nonlinear
    physical_flux! through VOL, burgers.jl:51-57)."""
    law = InviscidBurgersEquation((1.0, 0.5))
    ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M)), ra, 0.1)
    sd = make_spatial_discretization(mesh, ra)
    solver = Solver(law, sd, StandardForm(), PhysicalOperator(), lazy=lazy)
    return solver, project_function(InitialDataSine(1.0, (2 * math.pi,) * 2), sd) + 1.5


def euler_standard_case(d=2, p=3, M=3, lazy=True, interface="lf", strategy="reference",
                        approx="modal"):
    """Euler under StandardForm: physical_flux! (euler_navierstokes.jl:58-69) in the volume and
    the conservative two-point flux (:152-158) + Lax-Friedrichs at the interfaces."""
    g = 1.4
    law = EulerEquations(d, g)
    elem = Tri() if d == 2 else Tet()
    at = ModalTensor(p) if approx == "modal" else NodalTensor(p)
    L = 2.0
    ra = make_reference_approximation(at, elem, mapping_degree=min(p, 3))
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, L),) * d, (M,) * d), ra,
                     ChanWarping(1 / 16, (L,) * d))
    sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics() if d == 3 else ExactMetrics())
    flux = LaxFriedrichsNumericalFlux() if interface == "lf" else CentralNumericalFlux()
    strat = ReferenceOperator() if strategy == "reference" else PhysicalOperator()
    solver = Solver(law, sd, StandardForm(inviscid_numerical_flux=flux), strat, lazy=lazy)
    return solver, project_function(EulerPeriodicTest(d, g, 0.2, L), sd)


def euler_conservative_fluxdiff_case(p=3, M=3, lazy=True):
    """FluxDifferencingForm with the *conservative* two-point flux (Solvers.jl:96-115 with
    two_point_flux = ConservativeFlux; euler_navierstokes.jl:152-158 in the volume term)."""
    from .conservation_laws import ConservativeFlux
    g = 1.4
    law = EulerEquations(2, g)
    ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 2.0),) * 2, (M, M)), ra,
                     ChanWarping(1 / 16, (2.0, 2.0)))
    sd = make_spatial_discretization(mesh, ra)
    form = FluxDifferencingForm(inviscid_numerical_flux=LaxFriedrichsNumericalFlux(),
                                two_point_flux=ConservativeFlux())
    solver = Solver(law, sd, form, ReferenceOperator(), lazy=lazy)
    return solver, project_function(EulerPeriodicTest(2, g, 0.2, 2.0), sd)


def mass_solver_case(kind="cholesky", form="standard", p=3, M=3, lazy=True):
    """Mass-matrix solvers other than the default (mass_matrix.jl:41-115,169-196):
    kind = 'cholesky'  -> CholeskySolver (per-element M_k = V^T W J_k V factorised),
           'wa_full'   -> WeightAdjustedSolver(assume_orthonormal=False) with a dense M^-1,
           'wa_diag'   -> the same with an inexact (LGL) volume quadrature: diagonal M^-1 != I."""
    from .reference_approximation import LGLQuadrature, LGQuadrature
    from .solvers import CholeskySolver, WeightAdjustedSolver
    vrule = None
    if kind == "cholesky":
        ms = CholeskySolver()
    elif kind == "wa_full":
        ms = WeightAdjustedSolver(assume_orthonormal=False, tol=0.0)
    else:
        ms = WeightAdjustedSolver(assume_orthonormal=False, tol=1.0e-13)
        vrule = (LGLQuadrature(p), LGQuadrature(p))
    ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p,
                                      volume_quadrature_rule=vrule)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M)), ra, 0.15)
    sd = make_spatial_discretization(mesh, ra)
    if form == "standard":
        law = LinearAdvectionEquation((1.0, 1.0))
        solver = Solver(law, sd, StandardForm(), ReferenceOperator(), mass_solver=ms, lazy=lazy)
        return solver, project_function(InitialDataSine(1.0, (2 * math.pi,) * 2), sd)
    g = 1.4
    law = EulerEquations(2, g)
    solver = Solver(law, sd, FluxDifferencingForm(), ReferenceOperator(), mass_solver=ms,
                    lazy=lazy)
    return solver, project_function(EulerPeriodicTest(2, g, 0.2, 1.0), sd)


def viscous_burgers_case(d=1, p=4, M=4, lazy=True):
    """ViscousBurgersEquation with BR1 (burgers.jl:51-142): nonlinear inviscid flux + the
    auxiliary-gradient viscous flux, PhysicalOperator."""
    from .conservation_laws import ViscousBurgersEquation
    if d == 1:
        law = ViscousBurgersEquation((1.0,), 2.0e-2)
        ra = make_reference_approximation(ModalMulti(p), Line())
        sd = make_spatial_discretization(uniform_periodic_mesh(ra, (0.0, 1.0), M), ra)
    else:
        law = ViscousBurgersEquation((1.0, 0.5), 2.0e-2)
        ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p)
        mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M)), ra, 0.1)
        sd = make_spatial_discretization(mesh, ra)
    form = StandardForm(StandardMapping(), LaxFriedrichsNumericalFlux(), BR1())
    solver = Solver(law, sd, form, PhysicalOperator(), lazy=lazy)
    return solver, project_function(InitialDataSine(1.0, (2 * math.pi,) * d), sd) + 1.5
