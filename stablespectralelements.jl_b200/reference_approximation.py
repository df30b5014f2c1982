"""Reference-element discretisations (host-side setup; NumPy, runs once).

Mirrors the user-facing types of /root/reference/src/SpatialDiscretizations/:
element shapes, approximation types (SpatialDiscretizations.jl:69-142), quadrature rules
(quadrature_rules.jl), the collapsed-coordinate Tri/Tet ``RefElemData`` (ref_elem_data.jl),
and the ``ReferenceApproximation`` constructors for ``NodalTensor``/``ModalTensor`` on
Line/Quad/Hex (tensor_cartesian.jl) and Tri/Tet (tensor_simplex.jl:158-306), plus the 1-D
``ModalMulti``/``NodalMulti`` (multidimensional.jl:1-80).  Multidimensional (non-tensor)
Tri/Tet rules need quadrature tables from un-vendored packages and are not provided.

All indices are 0-based; node ordering follows the reference (first coordinate slowest,
quadrature_rules.jl:37-46).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import polynomials as poly
from . import nodes as nd
from .linear_maps import (BlockMap, DenseMap, IdentityMap, KroneckerMap, LinearMap,
                          SelectionMap, WarpedTensorProductMap2D, WarpedTensorProductMap3D)


# ------------------------------------------------------------------ element shapes
class AbstractElemShape:
    dim: int
    num_faces: int
    name: str

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return f"{type(self).__name__}()"


class Line(AbstractElemShape):
    dim, num_faces, name = 1, 2, "Line"


class Quad(AbstractElemShape):
    dim, num_faces, name = 2, 4, "Quad"


class Tri(AbstractElemShape):
    dim, num_faces, name = 2, 3, "Tri"


class Hex(AbstractElemShape):
    dim, num_faces, name = 3, 6, "Hex"


class Tet(AbstractElemShape):
    dim, num_faces, name = 3, 4, "Tet"


# --------------------------------------------------------------- approximation types
@dataclass(frozen=True)
class NodalTensor:
    p: int


@dataclass(frozen=True)
class ModalTensor:
    p: int


@dataclass(frozen=True)
class ModalMulti:
    p: int


@dataclass(frozen=True)
class NodalMulti:
    p: int


# ------------------------------------------------------------------ quadrature rules
@dataclass(frozen=True)
class GaussLobattoQuadrature:
    q: int
    a: int = 0
    b: int = 0


@dataclass(frozen=True)
class GaussQuadrature:
    q: int
    a: int = 0
    b: int = 0


@dataclass(frozen=True)
class GaussRadauQuadrature:
    q: int
    a: int = 0
    b: int = 0


@dataclass(frozen=True)
class DefaultQuadrature:
    degree: int


def LGLQuadrature(q: int):
    return GaussLobattoQuadrature(q, 0, 0)


def LGQuadrature(q: int):
    return GaussQuadrature(q, 0, 0)


def LGRQuadrature(q: int):
    return GaussRadauQuadrature(q, 0, 0)


def quadrature_line(rule):
    """quadrature_rules.jl:48-92."""
    if isinstance(rule, DefaultQuadrature):
        return quadrature_line(LGQuadrature(int(np.ceil((rule.degree - 1) / 2))))
    if isinstance(rule, GaussLobattoQuadrature):
        return poly.gauss_lobatto(rule.q + 1, rule.a, rule.b)
    if isinstance(rule, GaussQuadrature):
        return poly.gauss_jacobi(rule.q + 1, rule.a, rule.b)
    if isinstance(rule, GaussRadauQuadrature):
        return poly.gauss_radau(rule.q + 1, rule.a, rule.b)
    raise TypeError(f"unsupported 1-D quadrature rule {rule!r}")


def _grid2(x, y):
    """First coordinate slowest (quadrature_rules.jl:37-40 followed by ``[:]``)."""
    return np.repeat(x, len(y)), np.tile(y, len(x))


def _grid3(x, y, z):
    ny, nz = len(y), len(z)
    return (np.repeat(x, ny * nz), np.tile(np.repeat(y, nz), len(x)),
            np.tile(z, len(x) * ny))


def chi_tri(eta1, eta2):
    """Duffy map square -> triangle (tensor_simplex.jl:2-4)."""
    return 0.5 * (1.0 + eta1) * (1.0 - eta2) - 1.0, eta2


def chi_tet(eta1, eta2, eta3):
    """Duffy map cube -> tetrahedron (tensor_simplex.jl:7-11)."""
    xi_pri1 = 0.5 * (1.0 + eta1) * (1.0 - eta3) - 1.0
    xi_pyr2 = 0.5 * (1.0 + eta2) * (1.0 - eta3) - 1.0
    return 0.5 * (1.0 + xi_pri1) * (1.0 - eta2) - 1.0, xi_pyr2, eta3


def _rules(rule, d):
    return tuple(rule) if isinstance(rule, (tuple, list)) else tuple(rule for _ in range(d))


def quadrature(elem, rule):
    """``quadrature(elem, rule)`` -> (r[, s[, t]], w) (quadrature_rules.jl:94-166)."""
    if isinstance(elem, Line):
        return quadrature_line(rule)
    if isinstance(elem, Quad):
        r1, r2 = _rules(rule, 2)
        (x1, w1), (x2, w2) = quadrature_line(r1), quadrature_line(r2)
        g1, g2 = _grid2(x1, x2)
        wg1, wg2 = _grid2(w1, w2)
        return g1, g2, wg1 * wg2
    if isinstance(elem, Hex):
        r1, r2, r3 = _rules(rule, 3)
        (x1, w1), (x2, w2), (x3, w3) = (quadrature_line(r1), quadrature_line(r2),
                                        quadrature_line(r3))
        g = _grid3(x1, x2, x3)
        wg = _grid3(w1, w2, w3)
        return g[0], g[1], g[2], wg[0] * wg[1] * wg[2]
    if isinstance(elem, Tri):
        r1, r2 = _rules(rule, 2)
        (x1, w1), (x2, w2) = quadrature_line(r1), quadrature_line(r2)
        g1, g2 = _grid2(x1, x2)
        wg1, wg2 = _grid2(w1, w2)
        w2d = wg1 * wg2
        r, s = chi_tri(g1, g2)
        if (r1.a, r1.b, r2.a, r2.b) == (0, 0, 0, 0):
            return r, s, 0.5 * (1 - g2) * w2d
        if (r1.a, r1.b, r2.a, r2.b) == (0, 0, 1, 0):
            return r, s, 0.5 * w2d
        raise ValueError("Chosen Jacobi weight not supported")
    if isinstance(elem, Tet):
        r1, r2, r3 = _rules(rule, 3)
        (x1, w1), (x2, w2), (x3, w3) = (quadrature_line(r1), quadrature_line(r2),
                                        quadrature_line(r3))
        g = _grid3(x1, x2, x3)
        wg = _grid3(w1, w2, w3)
        w3d = wg[0] * wg[1] * wg[2]
        r, s, t = chi_tet(*g)
        ab = (r1.a, r1.b, r2.a, r2.b, r3.a, r3.b)
        if ab == (0, 0, 0, 0, 0, 0):
            return r, s, t, 0.125 * (1 - g[1]) * (1 - g[2]) ** 2 * w3d
        if ab == (0, 0, 0, 0, 1, 0):
            return r, s, t, 0.125 * (1 - g[1]) * (1 - g[2]) * w3d
        raise ValueError("Chosen Jacobi weight not supported")
    raise TypeError(f"unsupported element {elem!r}")


# ------------------------------------------------------------ mapping element data
def _basis(elem, N, *rst, grad=False):
    if isinstance(elem, Line):
        V = poly.vandermonde_1d(N, rst[0])
        return (V, poly.grad_vandermonde_1d(N, rst[0])) if grad else V
    if isinstance(elem, Tri):
        return poly.simplex_basis_2d(N, *rst, grad=grad)
    if isinstance(elem, Tet):
        return poly.simplex_basis_3d(N, *rst, grad=grad)
    if isinstance(elem, Quad):
        return poly.tensor_basis_2d(N, *rst, grad=grad)
    if isinstance(elem, Hex):
        return poly.tensor_basis_3d(N, *rst, grad=grad)
    raise TypeError(elem)


def vandermonde(elem, N, *rst):
    return _basis(elem, N, *rst)


def mapping_nodes(elem, N):
    if isinstance(elem, Line):
        return (nd.nodes_line(N),)
    if isinstance(elem, Tri):
        return nd.nodes_tri(N)
    if isinstance(elem, Tet):
        return nd.nodes_tet(N)
    if isinstance(elem, Quad):
        return nd.nodes_quad(N)
    if isinstance(elem, Hex):
        return nd.nodes_hex(N)
    raise TypeError(elem)


def reference_vertices(elem) -> np.ndarray:
    """Vertex coordinates of the reference element, (n_vertices, d)."""
    if isinstance(elem, Line):
        return np.array([[-1.0], [1.0]])
    if isinstance(elem, Tri):
        return np.array([[-1.0, -1.0], [1.0, -1.0], [-1.0, 1.0]])
    if isinstance(elem, Tet):
        return np.array([[-1.0, -1, -1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]])
    if isinstance(elem, Quad):
        return np.array([[-1.0, -1], [1, -1], [-1, 1], [1, 1]])
    if isinstance(elem, Hex):
        return np.array([[-1.0, -1, -1], [1, -1, -1], [-1, 1, -1], [1, 1, -1],
                         [-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]])
    raise TypeError(elem)


def face_vertices(elem):
    """Local vertex ids of each reference face, ordered like the facet quadrature blocks."""
    if isinstance(elem, Line):
        return [(0,), (1,)]
    if isinstance(elem, Tri):      # s=-1, hypotenuse, r=-1 (ref_elem_data.jl:24-27)
        return [(0, 1), (1, 2), (2, 0)]
    if isinstance(elem, Tet):      # s=-1, r+s+t=-1, r=-1, t=-1 (ref_elem_data.jl:88-100)
        return [(0, 1, 3), (1, 2, 3), (0, 2, 3), (0, 1, 2)]
    if isinstance(elem, Quad):     # r=-1, r=+1, s=-1, s=+1
        return [(0, 2), (1, 3), (0, 1), (2, 3)]
    if isinstance(elem, Hex):      # r=-1, r=+1, s=-1, s=+1, t=-1, t=+1
        return [(0, 2, 4, 6), (1, 3, 5, 7), (0, 1, 4, 5), (2, 3, 6, 7), (0, 1, 2, 3),
                (4, 5, 6, 7)]
    raise TypeError(elem)


def vertex_interpolation(elem, *rst) -> np.ndarray:
    """Linear (multilinear on Quad/Hex) vertex shape functions evaluated at ``rst``."""
    if isinstance(elem, Line):
        r, = rst
        return np.stack([0.5 * (1 - r), 0.5 * (1 + r)], axis=1)
    if isinstance(elem, Tri):
        r, s = rst
        return np.stack([-0.5 * (r + s), 0.5 * (1 + r), 0.5 * (1 + s)], axis=1)
    if isinstance(elem, Tet):
        r, s, t = rst
        return np.stack([-0.5 * (1 + r + s + t), 0.5 * (1 + r), 0.5 * (1 + s), 0.5 * (1 + t)],
                        axis=1)
    if isinstance(elem, Quad):
        r, s = rst
        return np.stack([0.25 * (1 - r) * (1 - s), 0.25 * (1 + r) * (1 - s),
                         0.25 * (1 - r) * (1 + s), 0.25 * (1 + r) * (1 + s)], axis=1)
    if isinstance(elem, Hex):
        r, s, t = rst
        cols = []
        for c in (-1, 1):
            for b in (-1, 1):
                for a in (-1, 1):
                    cols.append(0.125 * (1 + a * r) * (1 + b * s) * (1 + c * t))
        return np.stack(cols, axis=1)
    raise TypeError(elem)


@dataclass
class RefElemData:
    """The subset of StartUpDG's ``RefElemData`` the reference's hot-path setup reads."""
    element_type: AbstractElemShape
    N: int                      # mapping degree
    rst: Tuple[np.ndarray, ...]  # mapping nodes
    VDM: np.ndarray
    Drst: Tuple[np.ndarray, ...]
    V1: np.ndarray              # vertices -> mapping nodes
    rstq: Tuple[np.ndarray, ...]
    wq: np.ndarray
    Vq: np.ndarray              # mapping nodes -> volume quadrature nodes
    rstf: Tuple[np.ndarray, ...]
    wf: np.ndarray
    Vf: np.ndarray              # mapping nodes -> facet quadrature nodes
    nrstJ: Tuple[np.ndarray, ...]
    fv: list = field(default_factory=list)

    @property
    def dim(self):
        return self.element_type.dim


def _facet_nodes(elem, facet_rule):
    """Facet quadrature nodes / weights / scaled reference normals."""
    if isinstance(elem, Line):
        return (np.array([-1.0, 1.0]),), np.array([1.0, 1.0]), (np.array([-1.0, 1.0]),)
    if isinstance(elem, Tri):      # ref_elem_data.jl:20-27
        r1, w1 = quadrature_line(facet_rule)
        e, z = np.ones_like(r1), np.zeros_like(r1)
        rf = np.concatenate([r1, -r1, -e])
        sf = np.concatenate([-e, r1, r1])
        return (rf, sf), np.concatenate([w1, w1, w1]), (np.concatenate([z, e, -e]),
                                                       np.concatenate([-e, e, z]))
    if isinstance(elem, Tet):      # ref_elem_data.jl:85-100
        r2, s2, w2 = quadrature(Tri(), tuple(facet_rule))
        e, z = np.ones_like(r2), np.zeros_like(r2)
        rf = np.concatenate([r2, -(e + r2 + s2), -e, r2])
        sf = np.concatenate([-e, r2, r2, s2])
        tf = np.concatenate([s2, s2, s2, -e])
        return ((rf, sf, tf), np.concatenate([w2] * 4),
                (np.concatenate([z, e, -e, z]), np.concatenate([-e, e, z, z]),
                 np.concatenate([z, e, z, -e])))
    if isinstance(elem, Quad):
        r1, w1 = quadrature_line(facet_rule)
        e, z = np.ones_like(r1), np.zeros_like(r1)
        rf = np.concatenate([-e, e, r1, r1])
        sf = np.concatenate([r1, r1, -e, e])
        return (rf, sf), np.concatenate([w1] * 4), (np.concatenate([-e, e, z, z]),
                                                    np.concatenate([z, z, -e, e]))
    if isinstance(elem, Hex):
        rq, sq, wq = quadrature(Quad(), facet_rule)
        e, z = np.ones_like(rq), np.zeros_like(rq)
        rf = np.concatenate([-e, e, rq, rq, rq, rq])
        sf = np.concatenate([rq, rq, -e, e, sq, sq])
        tf = np.concatenate([sq, sq, sq, sq, -e, e])
        return ((rf, sf, tf), np.concatenate([wq] * 6),
                (np.concatenate([-e, e, z, z, z, z]), np.concatenate([z, z, -e, e, z, z]),
                 np.concatenate([z, z, z, z, -e, e])))
    raise TypeError(elem)


def make_ref_elem_data(elem, N: int, volume_quadrature, facet_rule) -> RefElemData:
    """``RefElemData(elem, approx_type, N; ...)`` (ref_elem_data.jl) for any element."""
    rst = mapping_nodes(elem, N)
    out = _basis(elem, N, *rst, grad=True)
    VDM, grads = out[0], out[1:]
    Drst = tuple(np.linalg.solve(VDM.T, g.T).T for g in grads)
    V1 = vertex_interpolation(elem, *rst)
    *rstq, wq = volume_quadrature
    rstq = tuple(rstq)
    Vq = np.linalg.solve(VDM.T, vandermonde(elem, N, *rstq).T).T
    rstf, wf, nrstJ = _facet_nodes(elem, facet_rule)
    Vf = np.linalg.solve(VDM.T, vandermonde(elem, N, *rstf).T).T
    return RefElemData(elem, N, rst, VDM, Drst, V1, rstq, wq, Vq, rstf, wf, Vf, nrstJ,
                       face_vertices(elem))


# ------------------------------------------------------- collapsed reference mapping
@dataclass
class NoMapping:
    pass


@dataclass
class ReferenceMapping:
    J_ref: np.ndarray            # (N_q,)
    Lambda_ref: np.ndarray       # (N_q, d, d)  J dη_l/dξ_m


def reference_geometric_factors(elem, rules):
    """tensor_simplex.jl:13-82."""
    if isinstance(elem, Tri):
        eta = quadrature(Quad(), rules)
        N = len(eta[0])
        L = np.zeros((N, 2, 2))
        ab = (rules[0].a, rules[0].b, rules[1].a, rules[1].b)
        if ab == (0, 0, 0, 0):
            J = 0.5 * (1.0 - eta[1])
            L[:, 0, 0] = 1.0
            L[:, 0, 1] = 0.5 * (1.0 + eta[0])
            L[:, 1, 1] = 0.5 * (1.0 - eta[1])
        elif ab == (0, 0, 1, 0):
            J = 0.5 * np.ones(N)
            L[:, 0, 0] = 1.0 / (1.0 - eta[1])
            L[:, 0, 1] = 0.5 * (1.0 + eta[0]) / (1.0 - eta[1])
            L[:, 1, 1] = 0.5
        else:
            raise ValueError("Chosen Jacobi weight not supported")
        return J, L
    if isinstance(elem, Tet):
        eta = quadrature(Hex(), rules)
        N = len(eta[0])
        L = np.zeros((N, 3, 3))
        ab = tuple(x for r in rules for x in (r.a, r.b))
        e1, e2, e3 = eta[0], eta[1], eta[2]
        if ab == (0, 0, 0, 0, 0, 0):
            J = 0.5 * (1 - e2) * (0.5 * (1 - e3)) ** 2
            L[:, 0, 0] = 0.5 * (1 - e3)
            L[:, 0, 1] = 0.5 * (1 + e1) * 0.5 * (1 - e3)
            L[:, 0, 2] = 0.5 * (1 + e1) * 0.5 * (1 - e3)
            L[:, 1, 1] = 0.5 * (1 - e2) * 0.5 * (1 - e3)
            L[:, 1, 2] = 0.5 * (1 + e2) * 0.5 * (1 - e2) * 0.5 * (1 - e3)
            L[:, 2, 2] = 0.5 * (1 - e2) * (0.5 * (1 - e3)) ** 2
        elif ab == (0, 0, 0, 0, 1, 0):
            J = 0.125 * (1 - e2) * (1 - e3)
            L[:, 0, 0] = 0.5
            L[:, 0, 1] = 0.25 * (1 + e1)
            L[:, 0, 2] = 0.25 * (1 + e1)
            L[:, 1, 1] = 0.25 * (1 - e2)
            L[:, 1, 2] = 0.125 * (1 + e2) * (1 - e2)
            L[:, 2, 2] = 0.125 * (1 - e2) * (1 - e3)
        else:
            raise ValueError("Chosen Jacobi weight not supported")
        return J, L
    raise TypeError(elem)


def warped_product(elem, p: int, eta1d):
    """tensor_simplex.jl:84-140 (0-based index tables; unused slots of sigma_i are -1)."""
    n = p + 1
    if isinstance(elem, Tri):
        M1, M2 = len(eta1d[0]), len(eta1d[1])
        sigma_o = np.arange(M1 * M2).reshape(M1, M2)
        sigma_i = -np.ones((n, n), dtype=np.int64)
        A = np.zeros((M1, n))
        B = np.zeros((M2, n, n))
        k = 0
        for i in range(n):
            A[:, i] = np.sqrt(2.0) * poly.jacobiP(eta1d[0], 0, 0, i)
            for j in range(n - i):
                sigma_i[i, j] = k
                k += 1
                B[:, i, j] = (1 - eta1d[1]) ** i * poly.jacobiP(eta1d[1], 2 * i + 1, 0, j)
        return WarpedTensorProductMap2D(A, B, sigma_i, sigma_o)
    if isinstance(elem, Tet):
        M1, M2, M3 = (len(e) for e in eta1d)
        sigma_o = np.arange(M1 * M2 * M3).reshape(M1, M2, M3)
        sigma_i = -np.ones((n, n, n), dtype=np.int64)
        A = np.zeros((M1, n))
        B = np.zeros((M2, n, n))
        C = np.zeros((M3, n, n, n))
        l = 0
        for i in range(n):
            A[:, i] = np.sqrt(2.0) * poly.jacobiP(eta1d[0], 0, 0, i)
            for j in range(n - i):
                B[:, i, j] = (1 - eta1d[1]) ** i * poly.jacobiP(eta1d[1], 2 * i + 1, 0, j)
                for k in range(n - i - j):
                    sigma_i[i, j, k] = l
                    l += 1
                    C[:, i, j, k] = (2 * (1 - eta1d[2]) ** (i + j)
                                     * poly.jacobiP(eta1d[2], 2 * i + 2 * j + 2, 0, k))
        return WarpedTensorProductMap3D(A, B, C, sigma_i, sigma_o)
    raise TypeError(elem)


def operators_1d(rules):
    """tensor_simplex.jl:142-156."""
    eta, q, V1, D1, I1, RL, RR = [], [], [], [], [], [], []
    for rule in rules:
        e, _ = quadrature_line(rule)
        eta.append(e)
        qm = len(e) - 1
        q.append(qm)
        Vm = poly.vandermonde_1d(qm, e)
        V1.append(Vm)
        D1.append(DenseMap(np.linalg.solve(Vm.T, poly.grad_vandermonde_1d(qm, e).T).T))
        I1.append(IdentityMap(qm + 1))
        RL.append(DenseMap(np.linalg.solve(Vm.T, poly.vandermonde_1d(qm, [-1.0]).T).T))
        RR.append(DenseMap(np.linalg.solve(Vm.T, poly.vandermonde_1d(qm, [1.0]).T).T))
    return eta, q, V1, D1, I1, RL, RR


def _interp_1d(vol_rule, fac_rule, q, V1d):
    if vol_rule == fac_rule:
        return IdentityMap(q + 1)
    ef, _ = quadrature_line(fac_rule)
    return DenseMap(np.linalg.solve(V1d.T, poly.vandermonde_1d(q, ef).T).T)


# ------------------------------------------------------------ ReferenceApproximation
@dataclass
class ReferenceApproximation:
    """SpatialDiscretizations.jl:186-246."""
    approx_type: object
    reference_element: RefElemData
    D: Tuple[LinearMap, ...]
    V: LinearMap
    Vf: LinearMap
    R: LinearMap
    reference_mapping: object = field(default_factory=NoMapping)
    # tensor-product structure of D (1-D matrices per direction), if any
    D_1D: Optional[Tuple[np.ndarray, ...]] = None

    def __post_init__(self):
        self.N_p = self.V.shape[1]
        self.N_q = self.V.shape[0]
        self.N_f = self.R.shape[0]
        self.W = np.asarray(self.reference_element.wq, dtype=np.float64)
        self.B = np.asarray(self.reference_element.wf, dtype=np.float64)

    @property
    def element_type(self):
        return self.reference_element.element_type

    @property
    def dim(self):
        return self.reference_element.element_type.dim


def make_reference_approximation(approx_type, elem, *, mapping_degree: int = 1,
                                 volume_quadrature_rule=None, facet_quadrature_rule=None,
                                 sum_factorize_vandermonde: bool = True
                                 ) -> ReferenceApproximation:
    """``ReferenceApproximation(approx_type, elem; kwargs...)`` for the supported pairs."""
    p = approx_type.p
    tensor = isinstance(approx_type, (NodalTensor, ModalTensor))

    # ---- Line -----------------------------------------------------------------
    if isinstance(elem, Line):
        if tensor:
            if isinstance(approx_type, ModalTensor):
                raise ValueError("ModalTensor is only defined on Tri and Tet")
            rule = volume_quadrature_rule or LGLQuadrature(p)       # tensor_cartesian.jl:1-34
            rq, wq = quadrature_line(rule)
            q = len(rq) - 1
            re = make_ref_elem_data(elem, mapping_degree, (rq, wq), None)
            VDM = poly.vandermonde_1d(q, rq)
            D = DenseMap(np.linalg.solve(VDM.T, poly.grad_vandermonde_1d(q, rq).T).T)
            if isinstance(rule, GaussLobattoQuadrature):
                R = SelectionMap([0, q], q + 1)
            else:
                R = DenseMap(np.linalg.solve(VDM.T, poly.vandermonde_1d(q, re.rstf[0]).T).T)
            return ReferenceApproximation(NodalTensor(q), re, (D,), IdentityMap(q + 1), R, R,
                                          D_1D=(D.A,))
        # ModalMulti / NodalMulti on the line (multidimensional.jl:1-80)
        rule = volume_quadrature_rule or DefaultQuadrature(2 * p)
        rq, wq = quadrature_line(rule)
        re = make_ref_elem_data(elem, mapping_degree, (rq, wq), None)
        VDM = poly.vandermonde_1d(p, rq)
        dVDM = poly.grad_vandermonde_1d(p, rq)
        Vf = poly.vandermonde_1d(p, re.rstf[0])
        P = np.linalg.solve(VDM.T @ (wq[:, None] * VDM), VDM.T * wq[None, :])
        if isinstance(approx_type, ModalMulti):
            return ReferenceApproximation(approx_type, re, (DenseMap(dVDM @ P),),
                                          DenseMap(VDM), DenseMap(Vf), DenseMap(Vf @ P))
        return ReferenceApproximation(approx_type, re, (DenseMap(dVDM @ P),),
                                      IdentityMap(len(wq)), DenseMap(Vf @ P), DenseMap(Vf @ P))

    if not tensor:
        raise NotImplementedError(
            "Multidimensional (non-tensor-product) rules on Tri/Tet need quadrature tables "
            "from un-vendored packages (NodesAndModes/StartUpDG); only Line is provided.")

    # ---- Quad / Hex (tensor_cartesian.jl:36-147) -----------------------------------
    if isinstance(elem, (Quad, Hex)):
        if isinstance(approx_type, ModalTensor):
            raise ValueError("ModalTensor is only defined on Tri and Tet")
        d = elem.dim
        vrule = volume_quadrature_rule or LGLQuadrature(p)
        frule = facet_quadrature_rule or LGLQuadrature(p)
        x1, _ = quadrature_line(vrule)
        q = len(x1) - 1
        VDM1 = poly.vandermonde_1d(q, x1)
        D1 = DenseMap(np.linalg.solve(VDM1.T, poly.grad_vandermonde_1d(q, x1).T).T)
        I1 = IdentityMap(q + 1)
        RL = DenseMap(np.linalg.solve(VDM1.T, poly.vandermonde_1d(q, [-1.0]).T).T)
        RR = DenseMap(np.linalg.solve(VDM1.T, poly.vandermonde_1d(q, [1.0]).T).T)
        re = make_ref_elem_data(elem, mapping_degree, quadrature(elem, vrule), frule)
        Xq = np.stack(re.rstq, axis=1)
        Xf = np.stack(re.rstf, axis=1)
        if vrule == frule and isinstance(vrule, GaussLobattoQuadrature):
            ids = [int(np.argmin(np.linalg.norm(Xq - xf, axis=1))) for xf in Xf]
            R = SelectionMap(ids, (q + 1) ** d)
        elif vrule == frule and d == 2:
            R = BlockMap([KroneckerMap(RL, I1), KroneckerMap(RR, I1), KroneckerMap(I1, RL),
                          KroneckerMap(I1, RR)])
        else:
            Vq_ = _basis(elem, q, *re.rstq)
            R = DenseMap(np.linalg.solve(Vq_.T, _basis(elem, q, *re.rstf).T).T)
        if d == 2:
            D = (KroneckerMap(D1, I1), KroneckerMap(I1, D1))
        else:
            D = (KroneckerMap(D1, I1, I1), KroneckerMap(I1, D1, I1), KroneckerMap(I1, I1, D1))
        return ReferenceApproximation(NodalTensor(q), re, D, IdentityMap((q + 1) ** d), R, R,
                                      D_1D=tuple(D1.A for _ in range(d)))

    # ---- Tri (tensor_simplex.jl:158-219) -----------------------------------------
    if isinstance(elem, Tri):
        vrule = tuple(volume_quadrature_rule or (LGQuadrature(p), LGQuadrature(p)))
        frule = facet_quadrature_rule or LGQuadrature(p)
        eta, q, V1, D1, I1, RL, RR = operators_1d(vrule)
        J_ref, L_ref = reference_geometric_factors(elem, vrule)
        e1f = _interp_1d(vrule[0], frule, q[0], V1[0])
        e2f = _interp_1d(vrule[1], frule, q[1], V1[1])
        R = BlockMap([KroneckerMap(e1f, RL[1]), KroneckerMap(RR[0], e2f),
                      KroneckerMap(RL[0], e2f)])
        re = make_ref_elem_data(elem, mapping_degree, quadrature(elem, vrule), frule)
        if isinstance(approx_type, ModalTensor):
            if sum_factorize_vandermonde:
                V = warped_product(elem, p, eta)
            else:
                V = DenseMap(vandermonde(elem, p, *re.rstq))
        else:
            V = IdentityMap((q[0] + 1) * (q[1] + 1))
        Vf = DenseMap(R.to_dense() @ V.to_dense())
        return ReferenceApproximation(approx_type, re,
                                      (KroneckerMap(D1[0], I1[1]), KroneckerMap(I1[0], D1[1])),
                                      V, Vf, R, ReferenceMapping(J_ref, L_ref),
                                      D_1D=(D1[0].A, D1[1].A))

    # ---- Tet (tensor_simplex.jl:221-306) -----------------------------------------
    if isinstance(elem, Tet):
        vrule = tuple(volume_quadrature_rule or (LGQuadrature(p), LGQuadrature(p),
                                                 GaussQuadrature(p, 1, 0)))
        frule = tuple(facet_quadrature_rule or (LGQuadrature(p), GaussQuadrature(p, 1, 0)))
        eta, q, V1, D1, I1, RL, RR = operators_1d(vrule)
        J_ref, L_ref = reference_geometric_factors(elem, vrule)
        e1f1 = _interp_1d(vrule[0], frule[0], q[0], V1[0])
        e2f1 = _interp_1d(vrule[1], frule[0], q[1], V1[1])
        e2f2 = _interp_1d(vrule[1], frule[1], q[1], V1[1])
        e3f2 = _interp_1d(vrule[2], frule[1], q[2], V1[2])
        R = BlockMap([KroneckerMap(e1f1, RL[1], e3f2), KroneckerMap(RR[0], e2f1, e3f2),
                      KroneckerMap(RL[0], e2f1, e3f2), KroneckerMap(e1f1, e2f2, RL[2])])
        re = make_ref_elem_data(elem, mapping_degree, quadrature(elem, vrule), frule)
        if isinstance(approx_type, ModalTensor):
            if sum_factorize_vandermonde:
                V = warped_product(elem, p, eta)
            else:
                V = DenseMap(vandermonde(elem, p, *re.rstq))
        else:
            V = IdentityMap((q[0] + 1) * (q[1] + 1) * (q[2] + 1))
        Vf = DenseMap(R.to_dense() @ V.to_dense())
        return ReferenceApproximation(
            approx_type, re,
            (KroneckerMap(D1[0], I1[1], I1[2]), KroneckerMap(I1[0], D1[1], I1[2]),
             KroneckerMap(I1[0], I1[1], D1[2])),
            V, Vf, R, ReferenceMapping(J_ref, L_ref), D_1D=(D1[0].A, D1[1].A, D1[2].A))

    raise TypeError(f"unsupported element {elem!r}")


def reference_derivative_operators(D_eta, reference_mapping):
    """SpatialDiscretizations.jl:414-423 (dense): D_ξm = Σ_l diag(Λ_ref[:,l,m]/J_ref) D_ηl."""
    Dd = [D.to_dense() for D in D_eta]
    if isinstance(reference_mapping, NoMapping):
        return Dd
    d = len(Dd)
    L, J = reference_mapping.Lambda_ref, reference_mapping.J_ref
    return [sum((L[:, l, m] / J)[:, None] * Dd[l] for l in range(d)) for m in range(d)]


def check_sbp_property(ra: ReferenceApproximation):
    """SpatialDiscretizations.jl:441-457."""
    D_xi = reference_derivative_operators(ra.D, ra.reference_mapping)
    R = ra.R.to_dense()
    out = []
    for m in range(ra.dim):
        Q = ra.W[:, None] * D_xi[m]
        E = R.T @ ((ra.B * ra.reference_element.nrstJ[m])[:, None] * R)
        out.append(np.max(np.abs(Q + Q.T - E)))
    return tuple(out)
