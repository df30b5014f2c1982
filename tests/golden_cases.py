"""The reference test-suite cases (/root/reference/test/runtests.jl) rebuilt on the host package.

Each function returns (solver, u0, T, dt, exact(x..., t) -> tuple, golden_l2).
"""
import math

import numpy as np

from sse_b200.conservation_laws import (CentralNumericalFlux, EntropyConservativeNumericalFlux,
                                        EulerEquations, InviscidBurgersEquation,
                                        LaxFriedrichsNumericalFlux,
                                        LinearAdvectionDiffusionEquation,
                                        LinearAdvectionEquation, BR1)
from sse_b200.geometric_factors import ChanWilcoxMetrics, make_spatial_discretization
from sse_b200.grid_functions import (EulerPeriodicTest, InitialDataCosine, InitialDataGassner, InitialDataSine,
                                     IsentropicVortex, evaluate)
from sse_b200.mesh import ChanWarping, uniform_periodic_mesh, warp_mesh
from sse_b200.reference_approximation import (Hex, LGQuadrature, Line, ModalMulti, ModalTensor,
                                              NodalTensor, Quad, Tet, Tri,
                                              make_reference_approximation)
from sse_b200.solvers import (FluxDifferencingForm, PhysicalOperator, ReferenceOperator, Solver,
                              StandardForm, StandardMapping, SkewSymmetricMapping,
                              project_function)


def _test_discretization(L, M, ra, perturb):
    d = ra.dim
    if d == 1:
        return make_spatial_discretization(uniform_periodic_mesh(ra, (0.0, L), M), ra)
    mesh = uniform_periodic_mesh(ra, ((0.0, L),) * d, (M,) * d)
    return make_spatial_discretization(warp_mesh(mesh, ra, perturb), ra,
                                       project_jacobian_flag=True)


def advection_diffusion_1d(lazy=True):
    """runtests.jl:14-36."""
    law = LinearAdvectionDiffusionEquation((1.0,), 5.0e-2)
    ra = make_reference_approximation(ModalMulti(4), Line())
    sd = _test_discretization(1.0, 4, ra, 0.1)
    ic = InitialDataSine(1.0, (2 * math.pi,))
    form = StandardForm(StandardMapping(), LaxFriedrichsNumericalFlux(), BR1())
    solver = Solver(law, sd, form, PhysicalOperator(), lazy=lazy)
    k2 = (2 * math.pi) ** 2

    def exact(x, t):
        return (np.sin(2 * math.pi * (x - t)) * math.exp(-5.0e-2 * k2 * t),)
    return solver, project_function(ic, sd), 1.0, 1.0 / 100.0, exact, [6.988216111882884e-6]


def advection_2d_tri(lazy=True):
    """runtests.jl:38-60."""
    law = LinearAdvectionEquation((1.0, 1.0))
    ra = make_reference_approximation(ModalTensor(4), Tri(), mapping_degree=4)
    sd = _test_discretization(1.0, 2, ra, 0.1)
    ic = InitialDataSine(1.0, (2 * math.pi, 2 * math.pi))
    form = StandardForm(SkewSymmetricMapping(), LaxFriedrichsNumericalFlux(0.0))
    solver = Solver(law, sd, form, ReferenceOperator(), lazy=lazy)

    def exact(x, y, t):
        return (np.sin(2 * math.pi * (x - t)) * np.sin(2 * math.pi * (y - t)),)
    return solver, project_function(ic, sd), 1.0, 1.0 / 100.0, exact, [0.2660013939427627]


def advection_2d_quad_fluxdiff(lazy=True):
    """runtests.jl:62-80."""
    law = LinearAdvectionEquation((1.0, 1.0))
    ra = make_reference_approximation(NodalTensor(4), Quad(), mapping_degree=4)
    sd = _test_discretization(1.0, 2, ra, 0.1)
    ic = InitialDataSine(1.0, (2 * math.pi, 2 * math.pi))
    solver = Solver(law, sd, FluxDifferencingForm(), ReferenceOperator(), lazy=lazy)

    def exact(x, y, t):
        return (np.sin(2 * math.pi * (x - t)) * np.sin(2 * math.pi * (y - t)),)
    return solver, project_function(ic, sd), 1.0, 1.0 / 100.0, exact, [0.04790536605026519]


def burgers_fluxdiff_1d(lazy=True):
    """runtests.jl:82-87, test/burgers_fluxdiff_1d.jl (no L2 number: invariants only)."""
    law = InviscidBurgersEquation()
    ra = make_reference_approximation(NodalTensor(7), Line())
    sd = make_spatial_discretization(uniform_periodic_mesh(ra, (0.0, 2.0), 20), ra)
    ic = InitialDataGassner(math.pi, 0.01)
    form = FluxDifferencingForm(inviscid_numerical_flux=EntropyConservativeNumericalFlux())
    solver = Solver(law, sd, form, lazy=lazy)
    h = 2.0 / (ra.N_p * sd.N_e)
    return solver, project_function(ic, sd), 0.3, 0.1 * h, None, None


def euler_1d_gauss(lazy=True):
    """runtests.jl:89-96, test/euler_1d_gauss.jl."""
    law = EulerEquations(1, 1.4)
    ra = make_reference_approximation(NodalTensor(5), Line(),
                                      volume_quadrature_rule=LGQuadrature(5))
    sd = make_spatial_discretization(uniform_periodic_mesh(ra, (0.0, 2.0), 4), ra)
    form = FluxDifferencingForm(inviscid_numerical_flux=EntropyConservativeNumericalFlux())
    solver = Solver(law, sd, form, ReferenceOperator(), lazy=lazy)

    def exact(x, t):
        rho = 1.0 + 0.2 * np.sin(math.pi * x)
        return rho, rho * 1.0, 1.0 / 0.4 + 0.5 * rho
    T = 2.0
    return solver, project_function(exact, sd), T, T / 1000, exact, \
        [3.5808560177567635e-5, 5.2129828619609155e-5, 0.00012637647535378534]


def euler_vortex_2d_modal(lazy=True, M=4, p=3):
    """runtests.jl:111-121, test/euler_vortex_2d_modal.jl."""
    g = 1.4
    L = 1.0
    T = L / 0.4
    law = EulerEquations(2, g)
    strength = math.sqrt(2 / (g - 1) * (1 - 0.75 ** (g - 1)))
    ic = IsentropicVortex(gamma=g, Ma=0.4, theta=0.0, R=0.1, beta=strength, sigma=1.0,
                          x_0=(L / 2, L / 2))
    ra = make_reference_approximation(ModalTensor(p), Tri(), mapping_degree=p)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, L), (0.0, L)), (M, M)), ra,
                     ChanWarping(1.0 / 16.0, (L, L)))
    sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics())
    form = FluxDifferencingForm(inviscid_numerical_flux=LaxFriedrichsNumericalFlux())
    solver = Solver(law, sd, form, ReferenceOperator(), lazy=lazy)

    def exact(x, y, t):
        return tuple(evaluate(ic, (x, y), t))
    return solver, project_function(ic, sd), T, T / 1000, exact, \
        [0.015568197027072704, 0.040539693811761104, 0.04060141777050208, 0.043960971468832745]


def advection_3d_tet(lazy=True):
    """runtests.jl:123-129, test/advection_3d.jl."""
    law = LinearAdvectionEquation((1.0, 1.0, 1.0))
    ra = make_reference_approximation(ModalTensor(4), Tet(), mapping_degree=4,
                                      sum_factorize_vandermonde=False)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 3, (2, 2, 2)), ra, 0.1, 1.0)
    sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics())
    ic = InitialDataCosine(1.0, (2 * math.pi,) * 3)
    form = StandardForm(SkewSymmetricMapping(), CentralNumericalFlux())
    solver = Solver(law, sd, form, ReferenceOperator(), lazy=lazy)
    h = 1.0 / (ra.N_p * sd.N_e) ** (1 / 3)
    dt = 0.1 * h / math.sqrt(3.0)

    def exact(x, y, z, t):
        return (np.cos(2 * math.pi * x) * np.cos(2 * math.pi * y) * np.cos(2 * math.pi * z),)
    return solver, project_function(ic, sd), 1.0, dt, exact, [0.1876141674772107]


def euler_3d_hex(lazy=True):
    """runtests.jl:131-144, test/euler_3d.jl: 3-D Euler, NodalTensor(4) Hex, M = 2, ChanWarping,
    conservative-curl metrics, flux differencing with the EC interface flux; integrated with
    DP8, N_t = 25 M (p+1) fixed steps (the tuple's dt slot holds N_t)."""
    g, L, M, p = 1.4, 2.0, 2, 4
    law = EulerEquations(3, g)
    ra = make_reference_approximation(NodalTensor(p), Hex(), mapping_degree=p)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, L),) * 3, (M,) * 3), ra,
                     ChanWarping(1.0 / 16.0, (L, L, L)))
    sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics())
    form = FluxDifferencingForm(inviscid_numerical_flux=EntropyConservativeNumericalFlux())
    solver = Solver(law, sd, form, ReferenceOperator(), lazy=lazy)
    ic = EulerPeriodicTest(3, g, 0.2, L)

    def exact(x, y, z, t):
        return tuple(evaluate(ic, (x, y, z), t))
    return solver, project_function(ic, sd), L, 25 * M * (p + 1), exact, \
        [0.18342164491797003, 0.1834216449179776, 0.18342164491796725, 0.18342164491796784,
         0.2751324673769553]
