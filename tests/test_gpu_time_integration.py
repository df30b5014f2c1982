"""Device-resident general explicit Runge-Kutta (sse_erk_step): the reference's 3-D Euler test
(test/euler_3d.jl, runtests.jl:131-144 -- NodalTensor(4) Hex, flux differencing with the EC
interface flux, DP8 with 250 fixed steps) run entirely on the GPU reproduces its golden L2 errors."""
import numpy as np
import pytest

import golden_cases as gc
import sse_oracle as oc
from bridge import oracle_problem

pytestmark = pytest.mark.gpu


def test_dp8_reproduces_reference_golden_l2_euler_3d_hex():
    from sse_b200.solvers import ODEProblem, semi_discrete_residual as f
    from sse_b200.time_integration import DP8, solve
    solver, u0, T, n_steps, exact, gold = gc.euler_3d_hex(lazy=False)
    try:
        u = solve(ODEProblem(f, u0, (0.0, T), solver), DP8(), dt=T / n_steps)
        prob = oracle_problem(solver)
        xq = tuple(x.T for x in solver.spatial_discretization.mesh.xyzq)
        l2 = oc.l2_error(prob, u, np.stack(exact(*xq, T), axis=-1))
        assert np.max(np.abs(l2 - np.array(gold))) < 1e-10, (l2, gold)
        du = np.empty_like(u)
        f(du, u, solver, T)
        assert abs(oc.entropy_residual(prob, u, du)) < 1e-10
    finally:
        solver.close()
