"""Pins the oracle (and the host-side setup) to the reference's own golden numbers:
the end-to-end L2 errors hard-coded in /root/reference/test/runtests.jl (atol 1e-10 there)."""
import numpy as np
import pytest

import golden_cases as gc
import sse_oracle as oc
from bridge import oracle_problem


def _run(case):
    solver, u0, T, dt, exact, gold = case
    prob = oracle_problem(solver)
    u = oc.ck54_integrate(lambda u, t: oc.semi_discrete_residual(prob, u, t), u0, (0.0, T), dt)
    xq = tuple(x.T for x in solver.spatial_discretization.mesh.xyzq)
    l2 = oc.l2_error(prob, u, np.stack(exact(*xq, T), axis=-1))
    du = oc.semi_discrete_residual(prob, u)
    return prob, u, du, l2, np.array(gold)


@pytest.mark.parametrize("name", ["advection_diffusion_1d", "euler_1d_gauss", "advection_2d_tri",
                                  "advection_2d_quad_fluxdiff"])
def test_reference_golden_l2(name):
    prob, u, du, l2, gold = _run(getattr(gc, name)())
    assert np.max(np.abs(l2 - gold)) < 1e-10, (l2, gold)
    assert np.max(np.abs(oc.conservation_residual(prob, du))) < 1e-10


@pytest.mark.slow
def test_reference_golden_l2_euler_vortex_modal_tri():
    """runtests.jl:111-121 -- the modal flux-differencing path with LF facets (north-star
    algorithm in 2-D): 5000 residual evaluations, ~15 s."""
    prob, u, du, l2, gold = _run(gc.euler_vortex_2d_modal())
    assert np.max(np.abs(l2 - gold)) < 1e-10, (l2, gold)
    assert np.max(np.abs(oc.conservation_residual(prob, du))) < 1e-10


def test_burgers_invariants():
    """runtests.jl:82-87: conservation and energy conservation with the EC interface flux."""
    solver, u0, T, dt, _, _ = gc.burgers_fluxdiff_1d()
    prob = oracle_problem(solver)
    u = oc.ck54_integrate(lambda u, t: oc.semi_discrete_residual(prob, u, t), u0, (0.0, 0.05), dt)
    du = oc.semi_discrete_residual(prob, u)
    assert abs(oc.conservation_residual(prob, du)[0]) < 1e-12
    assert abs(oc.energy_residual(prob, u, du)[0]) < 1e-12


def test_energy_conservation_central_flux_tri():
    """runtests.jl:59: λ = 0 Lax-Friedrichs == central flux conserves energy."""
    solver, u0, *_ = gc.advection_2d_tri()
    prob = oracle_problem(solver)
    du = oc.semi_discrete_residual(prob, u0)
    assert abs(oc.energy_residual(prob, u0, du)[0]) < 1e-12
    assert abs(oc.conservation_residual(prob, du)[0]) < 1e-12
