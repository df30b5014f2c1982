"""Round-off invariants of the north-star path (3-D Euler, tetrahedra, flux differencing) for
which the reference has no test at all: conservation, entropy conservation with the EC
interface flux (Analysis/conservation.jl:145-190), free-stream preservation, and agreement of
the C/OpenMP restatement with the NumPy oracle."""
import numpy as np
import pytest

import cases
import sse_oracle as oc
from bridge import oracle_problem


@pytest.fixture(scope="module")
def tet_ec():
    solver, u0 = cases.euler_tet_case(p=3, M=2, warp=True, interface="ec", ic="periodic")
    return oracle_problem(solver), cases.rough_state(solver, u0, seed=2)


def test_conservation_and_entropy_conservation_tet(tet_ec):
    prob, u = tet_ec
    du = oc.semi_discrete_residual(prob, u)
    scale = np.sum(np.abs(du)) * np.max(np.abs(u))
    assert np.max(np.abs(oc.conservation_residual(prob, du))) < 1e-13 * scale
    assert abs(oc.entropy_residual(prob, u, du)) < 1e-13 * scale


def test_lax_friedrichs_dissipates_entropy():
    solver, u0 = cases.euler_tet_case(p=3, M=2, warp=True, interface="lf", ic="periodic")
    prob = oracle_problem(solver)
    u = cases.rough_state(solver, u0, seed=2)
    assert oc.entropy_residual(prob, u, oc.semi_discrete_residual(prob, u)) < 0.0


def test_free_stream_preservation_tet():
    solver, u0 = cases.euler_tet_case(p=4, M=2, warp=True)
    prob = oracle_problem(solver)
    u = np.zeros_like(u0)
    const = np.array([1.2, 0.3, -0.2, 0.5, 3.0])
    V = prob["V"]
    # modal coefficients of a constant state: project the constant
    uq = np.broadcast_to(const, (u0.shape[0], V.shape[0], 5))
    u = oc.project_initial_data(prob, uq)
    du = oc.semi_discrete_residual(prob, u)
    assert np.max(np.abs(du)) < 1e-10


@pytest.mark.parametrize("builder", [
    lambda: cases.euler_tet_case(p=4, M=2, warp=True),
    lambda: cases.euler_tri_case(p=4, M=3),
    lambda: cases.euler_tri_case(p=3, M=3, approx="nodal", interface="ec"),
])
def test_c_oracle_matches_numpy_oracle(builder):
    c_oracle = pytest.importorskip("c_oracle")
    import os
    if not os.path.exists(c_oracle.LIB):
        pytest.skip("C oracle not built")
    solver, u0 = builder()
    prob = oracle_problem(solver)
    u = cases.rough_state(solver, u0, seed=5)
    ref = oc.semi_discrete_residual(prob, u)
    V = solver.spatial_discretization.reference_approximation.V
    warped = (V.A, V.B, getattr(V, "C", None), V.sigma_i) if hasattr(V, "sigma_i") else None
    for wp in (None, warped):
        r = c_oracle.make_residual(prob, wp)(u)
        assert np.max(np.abs(r - ref)) / np.max(np.abs(ref)) < 1e-12
