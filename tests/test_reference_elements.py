"""Known-answer checks of the reference-element construction."""
import numpy as np
import pytest

from sse_b200.reference_approximation import (Hex, LGQuadrature, Line, ModalMulti, ModalTensor,
                                              NodalTensor, Quad, Tet, Tri, check_sbp_property,
                                              make_reference_approximation, quadrature,
                                              reference_derivative_operators)

CASES = [(Tri(), ModalTensor(2)), (Tri(), ModalTensor(4)), (Tri(), NodalTensor(3)),
         (Tet(), ModalTensor(2)), (Tet(), ModalTensor(4)), (Tet(), NodalTensor(2)),
         (Quad(), NodalTensor(4)), (Hex(), NodalTensor(3)), (Line(), NodalTensor(5)),
         (Line(), ModalMulti(4))]


@pytest.mark.parametrize("elem,approx", CASES, ids=lambda x: repr(x))
def test_sbp_property(elem, approx):
    """W D_ξ + D_ξᵀ W = Rᵀ B diag(n̂) R (SpatialDiscretizations.jl:441-457)."""
    ra = make_reference_approximation(approx, elem, mapping_degree=2)
    assert max(check_sbp_property(ra)) < 5e-14


@pytest.mark.parametrize("elem,p", [(Tri(), 3), (Tri(), 4), (Tet(), 3), (Tet(), 4)])
def test_modal_basis_orthonormal_and_warped_equals_dense(elem, p):
    ra = make_reference_approximation(ModalTensor(p), elem)
    V = ra.V.to_dense()
    assert np.max(np.abs(V.T @ (ra.W[:, None] * V) - np.eye(ra.N_p))) < 1e-13
    rb = make_reference_approximation(ModalTensor(p), elem, sum_factorize_vandermonde=False)
    assert np.max(np.abs(V - rb.V.to_dense())) < 1e-12
    assert abs(ra.W.sum() - (2.0 if elem.dim == 2 else 4.0 / 3.0)) < 1e-13


def test_sizes_and_sparsity_tet_p4():
    """N_p, N_q, N_f = 35, 125, 100; nnz(S_m upper) = 250/500/750; nnz(C) = 1000."""
    ra = make_reference_approximation(ModalTensor(4), Tet())
    assert (ra.N_p, ra.N_q, ra.N_f) == (35, 125, 100)
    D_xi = reference_derivative_operators(ra.D, ra.reference_mapping)
    union = np.zeros((125, 125), dtype=bool)
    counts = []
    for m in range(3):
        S = 0.5 * (ra.W[:, None] * D_xi[m] - D_xi[m].T * ra.W[None, :])
        assert np.max(np.abs(S + S.T)) < 1e-14
        counts.append(int(np.count_nonzero(np.triu(S, 1))))
        union |= S != 0
    assert counts == [250, 500, 750]
    assert int(np.count_nonzero(np.triu(union, 1))) == 750
    R = ra.R.to_dense()
    assert int(np.count_nonzero(R)) == 1000
    per_face = [int(np.count_nonzero(R[25 * f:25 * (f + 1)])) for f in range(4)]
    assert per_face == [125, 125, 125, 625]


def test_sizes_and_sparsity_tri_p4():
    ra = make_reference_approximation(ModalTensor(4), Tri())
    assert (ra.N_p, ra.N_q, ra.N_f) == (15, 25, 15)
    D_xi = reference_derivative_operators(ra.D, ra.reference_mapping)
    counts = [int(np.count_nonzero(np.triu(0.5 * (ra.W[:, None] * D - D.T * ra.W[None, :]), 1)))
              for D in D_xi]
    assert counts == [50, 100]
    assert int(np.count_nonzero(ra.R.to_dense())) == 75


def test_quadrature_exactness():
    r, s, t, w = quadrature(Tet(), (LGQuadrature(3), LGQuadrature(3), LGQuadrature(3)))
    # ∫ r^2 s t over the reference tet, against the default (Jacobi) rule of higher degree
    ra = make_reference_approximation(ModalTensor(4), Tet())
    rq, sq, tq = ra.reference_element.rstq
    f = lambda a, b, c: a ** 2 * b * c
    assert abs(np.sum(w * f(r, s, t)) - np.sum(ra.W * f(rq, sq, tq))) < 1e-13
