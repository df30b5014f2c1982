"""Mesh / metric self-checks mirroring check_normals, check_facet_nodes
(SpatialDiscretizations.jl:426-439) plus the discrete metric identities."""
import numpy as np
import pytest

from sse_b200.geometric_factors import (ChanWilcoxMetrics, ExactMetrics, check_facet_nodes,
                                        check_metric_identities, check_normals,
                                        make_spatial_discretization)
from sse_b200.mesh import ChanWarping, uniform_periodic_mesh, warp_mesh
from sse_b200.reference_approximation import (Hex, Line, ModalTensor, NodalTensor, Quad, Tet,
                                              Tri, make_reference_approximation)

CASES = [(Tri(), ModalTensor(3), (4, 3), 3), (Quad(), NodalTensor(4), (2, 3), 4),
         (Tet(), ModalTensor(4), (2, 2, 2), 4), (Tet(), ModalTensor(2), (2, 3, 2), 2),
         (Hex(), NodalTensor(2), (2, 2, 2), 2)]


@pytest.mark.parametrize("elem,approx,M,md", CASES, ids=lambda x: repr(x))
def test_watertight_and_conforming(elem, approx, M, md):
    d = elem.dim
    ra = make_reference_approximation(approx, elem, mapping_degree=md)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, 1.0),) * d, M), ra,
                     ChanWarping(1 / 16, (1.0,) * d))
    mp = mesh.mapP.ravel(order="F")
    assert np.array_equal(mp[mp], np.arange(mp.size))          # involution
    assert not np.any(mp == np.arange(mp.size))
    for mt in (ExactMetrics(), ChanWilcoxMetrics()):
        sd = make_spatial_discretization(mesh, ra, mt)
        assert check_normals(sd) < 1e-11
        assert check_facet_nodes(sd) < 1e-12
        assert sd.geometric_factors.J_q.min() > 0
        if d == 2 or isinstance(mt, ChanWilcoxMetrics):
            assert check_metric_identities(sd) < 1e-11      # free-stream preservation
    gf = make_spatial_discretization(mesh, ra, ExactMetrics(), False).geometric_factors
    assert abs((gf.J_q * ra.W).sum() - 1.0) < 1e-3             # domain volume


def test_line_mesh():
    ra = make_reference_approximation(NodalTensor(3), Line())
    mesh = uniform_periodic_mesh(ra, (0.0, 2.0), 5)
    sd = make_spatial_discretization(mesh, ra)
    assert np.allclose(sd.geometric_factors.J_q, 0.2)
    assert mesh.mapP[0, 0] == 1 + 2 * 4 and mesh.mapP[1, 4] == 0
