"""GPU parity of the on-device geometric factors (sse_geometry_build) against the
host restatement of GeometricFactors (mesh.jl:213-509), which the reference's golden L2 errors pin
(tests/test_oracle_golden.py)."""
import math

import numpy as np
import pytest

import cases
from sse_b200.geometric_factors import (ChanWilcoxMetrics, ExactMetrics,
                                        make_spatial_discretization)
from sse_b200.mesh import ChanWarping, uniform_periodic_mesh, warp_mesh
from sse_b200.reference_approximation import (Hex, Line, ModalMulti, ModalTensor, NodalTensor,
                                              Tet, Tri, make_reference_approximation)

pytestmark = pytest.mark.gpu


def _meshes():
    L = 2 * math.pi
    ra = make_reference_approximation(ModalTensor(3), Tet(), mapping_degree=3)
    yield "tet_curl_warped", ra, warp_mesh(uniform_periodic_mesh(ra, ((0.0, L),) * 3, (3,) * 3),
                                           ra, ChanWarping(1 / 16, (L,) * 3)), ChanWilcoxMetrics()
    ra = make_reference_approximation(ModalTensor(4), Tet(), mapping_degree=1)
    yield "tet_exact_straight", ra, uniform_periodic_mesh(ra, ((0.0, L),) * 3, (2,) * 3), None
    ra = make_reference_approximation(ModalTensor(4), Tri(), mapping_degree=4)
    yield "tri_exact_projected", ra, warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2,
                                                                      (4, 4)), ra, 0.2), None
    ra = make_reference_approximation(NodalTensor(3), Hex(), mapping_degree=3)
    yield "hex_curl_warped", ra, warp_mesh(uniform_periodic_mesh(ra, ((0.0, 2.0),) * 3, (2,) * 3),
                                           ra, ChanWarping(1 / 16, (2.0,) * 3)), ChanWilcoxMetrics()
    ra = make_reference_approximation(ModalMulti(4), Line())
    yield "line_exact", ra, uniform_periodic_mesh(ra, (0.0, 1.0), 7), ExactMetrics()


@pytest.mark.parametrize("case", list(_meshes()), ids=lambda c: c[0])
def test_device_geometry_matches_host(case):
    name, ra, mesh, metric = case
    host = make_spatial_discretization(mesh, ra, metric).geometric_factors
    dev = make_spatial_discretization(mesh, ra, metric, device_geometry=0).geometric_factors
    try:
        for field in ("J_q", "Lambda_q", "J_f", "nJf"):
            a, b = getattr(dev, field), getattr(host, field)
            assert a.shape == b.shape, field
            assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), (name, field)
    finally:
        dev.free()


def test_residual_from_device_geometry_matches_host_geometry():
    """The north-star residual with device-built geometry (pointers handed to sse_create, nothing
    uploaded) equals the one with host-built geometry."""
    solver_h, u0 = cases.euler_tet_case(p=4, M=2, lazy=False, warp=True, ic="periodic")
    solver_d, _ = cases.euler_tet_case(p=4, M=2, lazy=False, warp=True, ic="periodic",
                                       device_geometry=0)
    try:
        from sse_b200.solvers import semi_discrete_residual
        u = cases.rough_state(solver_h, u0, seed=2)
        a, b = np.empty_like(u), np.empty_like(u)
        semi_discrete_residual(a, u, solver_h, 0.0)
        semi_discrete_residual(b, u, solver_d, 0.0)
        assert np.max(np.abs(a - b)) < 1e-12 * np.max(np.abs(a))
    finally:
        solver_h.close()
        solver_d.close()
