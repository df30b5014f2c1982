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


def test_reference_golden_l2_euler_3d_hex():
    """runtests.jl:131-144 -- 3-D Euler flux differencing (EC two-point and interface flux,
    conservative-curl metrics) on curved hexahedra, DP8 with 250 fixed steps (3000 residual
    evaluations, ~30 s): pins the 3-D Euler physics and 3-D metrics the north-star path shares."""
    solver, u0, T, n_steps, exact, gold = gc.euler_3d_hex()
    prob = oracle_problem(solver)
    u = oc.dp8_integrate(lambda u, t: oc.semi_discrete_residual(prob, u, t), u0, (0.0, T), n_steps)
    xq = tuple(x.T for x in solver.spatial_discretization.mesh.xyzq)
    l2 = oc.l2_error(prob, u, np.stack(exact(*xq, T), axis=-1))
    du = oc.semi_discrete_residual(prob, u)
    assert np.max(np.abs(l2 - np.array(gold))) < 1e-10, (l2, gold)
    assert np.max(np.abs(oc.conservation_residual(prob, du))) < 1e-10
    assert abs(oc.entropy_residual(prob, u, du)) < 1e-10


def _tet_error_quadrature(solver, p, n):
    """error.jl:21-40 with a collapsed Legendre-Gauss rule of n^3 nodes standing in for the
    un-vendored Jaskowiec-Sukumar table."""
    from sse_b200.reference_approximation import LGQuadrature, Tet, quadrature, vandermonde
    ra = solver.spatial_discretization.reference_approximation
    re = ra.reference_element
    r, s, t, w = (a.ravel() for a in quadrature(Tet(), LGQuadrature(n)))
    VDM = vandermonde(Tet(), p, *(np.asarray(a).ravel() for a in re.rstq))
    wq = np.asarray(re.wq).ravel()
    P = np.linalg.solve(VDM.T @ (wq[:, None] * VDM), VDM.T * wq[None, :])
    return vandermonde(Tet(), p, r, s, t) @ P, w


def test_reference_golden_l2_advection_3d_tet():
    """runtests.jl:123-129, test/advection_3d.jl -- the reference's only Tet test: ModalTensor(4)
    (dense V), curved mesh, conservative-curl metrics, skew-symmetric standard form, central
    flux.  Its L2 error is taken with JaskowiecSukumarQuadrature(11), a table that lives in
    un-vendored StartUpDG, so the comparison is limited by the error-quadrature's own accuracy:
    collapsed Gauss rules of the same degree class (6^3 nodes) and converged ones (8^3, 10^3)
    bracket the golden number to 4e-5, while every other choice of tetrahedral split / vertex
    numbering in the host mesh generator misses it by 6e-4 .. 3e-2 -- this is what identified
    StartUpDG's split (mesh.py).  Energy conservation and conservation hold to round-off."""
    solver, u0, T, dt, exact, gold = gc.advection_3d_tet()
    prob = oracle_problem(solver)
    u = oc.ck54_integrate(lambda u, t: oc.semi_discrete_residual(prob, u, t), u0, (0.0, T), dt)
    xyzq = tuple(x.T for x in solver.spatial_discretization.mesh.xyzq)
    l2 = {n: oc.l2_error_quadrature(prob, u, exact, xyzq, T, *_tet_error_quadrature(solver, 4, n))[0]
          for n in (6, 10)}
    assert abs(l2[6] - gold[0]) < 2e-5, l2        # observed -9.0e-6
    assert abs(l2[10] - gold[0]) < 5e-5, l2       # observed -3.6e-5 (the converged integral)
    du = oc.semi_discrete_residual(prob, u)
    assert abs(oc.conservation_residual(prob, du)[0]) < 1e-10
    assert abs(oc.energy_residual(prob, u, du)[0]) < 1e-10


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
