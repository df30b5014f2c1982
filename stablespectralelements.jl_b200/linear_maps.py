"""Host-side mirrors of the reference's matrix-free operator types.

Mirrors /root/reference/src/MatrixFreeOperators/ (``WarpedTensorProductMap2D/3D``,
``SelectionMap``, dense ``OctavianMap``/``GenericMatrixMap``) and the LinearMaps.jl types the
reference composes them with (``UniformScalingMap``, ``KroneckerMap`` built with ``⊗``,
``BlockMap`` built with ``vcat``).  On the host these objects only *describe* an operator:
``to_dense()`` feeds the oracle/tests, ``to_csr()`` and the warped-product tables feed the
C-ABI (``sse_operators``), where the actual application happens in CUDA.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


class LinearMap:
    shape: tuple

    def to_dense(self) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError

    def to_csr(self, tol: float = 0.0):
        """CSR triplet (rowptr int32, col int32, val f64); entries with |a| <= tol dropped."""
        A = self.to_dense()
        mask = np.abs(A) > tol
        rowptr = np.concatenate(([0], np.cumsum(mask.sum(axis=1)))).astype(np.int32)
        rows, cols = np.nonzero(mask)
        return rowptr, cols.astype(np.int32), A[rows, cols].astype(np.float64)

    def __matmul__(self, x):
        return self.to_dense() @ x

    @property
    def T(self):
        return DenseMap(self.to_dense().T)


class IdentityMap(LinearMap):
    """LinearMaps.UniformScalingMap (nodal schemes: V = I)."""

    def __init__(self, n: int):
        self.shape = (n, n)

    def to_dense(self):
        return np.eye(self.shape[0])


class DenseMap(LinearMap):
    """OctavianMap / GenericMatrixMap / WrappedMap: a plain dense matrix."""

    def __init__(self, A):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.shape = self.A.shape

    def to_dense(self):
        return self.A


class KroneckerMap(LinearMap):
    """``A ⊗ B [⊗ C]`` (first factor slowest index), as LinearMaps.KroneckerMap."""

    def __init__(self, *maps: LinearMap):
        self.maps = maps
        r = c = 1
        for m in maps:
            r *= m.shape[0]
            c *= m.shape[1]
        self.shape = (r, c)

    def to_dense(self):
        out = np.ones((1, 1))
        for m in self.maps:
            out = np.kron(out, m.to_dense())
        return out


class BlockMap(LinearMap):
    """Vertical concatenation ``[A; B; ...]`` (LinearMaps.BlockMap via vcat)."""

    def __init__(self, blocks: Sequence[LinearMap]):
        self.blocks = list(blocks)
        self.shape = (sum(b.shape[0] for b in self.blocks), self.blocks[0].shape[1])

    def to_dense(self):
        return np.vstack([b.to_dense() for b in self.blocks])


class SelectionMap(LinearMap):
    """selection.jl:3-34 -- gathers volume nodes ``facet_ids`` (0-based here)."""

    def __init__(self, facet_ids, n_vol: int):
        self.facet_ids = np.asarray(facet_ids, dtype=np.int64)
        self.shape = (len(self.facet_ids), n_vol)

    def to_dense(self):
        A = np.zeros(self.shape)
        A[np.arange(self.shape[0]), self.facet_ids] = 1.0
        return A


class WarpedTensorProductMap2D(LinearMap):
    """warped_product_2d.jl:2-25; tables built by tensor_simplex.jl:84-108.

    ``A[a1, b1]``, ``B[a2, b1, b2]``; input index ``sigma_i[b1, b2]`` (-1 where unused),
    output index ``sigma_o[a1, a2]`` (all 0-based).
    """

    def __init__(self, A, B, sigma_i, sigma_o):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.B = np.ascontiguousarray(B, dtype=np.float64)
        self.sigma_i = np.asarray(sigma_i, dtype=np.int64)
        self.sigma_o = np.asarray(sigma_o, dtype=np.int64)
        self.N2 = (self.sigma_i >= 0).sum(axis=1)
        self.shape = (self.A.shape[0] * self.B.shape[0], int(self.N2.sum()))

    def to_dense(self):
        M1, M2 = self.sigma_o.shape
        V = np.zeros(self.shape)
        for b1 in range(self.sigma_i.shape[0]):
            for b2 in range(self.N2[b1]):
                col = np.einsum("a,b->ab", self.A[:, b1], self.B[:, b1, b2])
                V[self.sigma_o.ravel(), self.sigma_i[b1, b2]] = col.ravel()
        return V


class WarpedTensorProductMap3D(LinearMap):
    """warped_product_3d.jl:2-43; tables built by tensor_simplex.jl:110-140."""

    def __init__(self, A, B, C, sigma_i, sigma_o):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.B = np.ascontiguousarray(B, dtype=np.float64)
        self.C = np.ascontiguousarray(C, dtype=np.float64)
        self.sigma_i = np.asarray(sigma_i, dtype=np.int64)
        self.sigma_o = np.asarray(sigma_o, dtype=np.int64)
        self.N2 = (self.sigma_i[:, 0, :] >= 0).sum(axis=1)
        self.N3 = (self.sigma_i >= 0).sum(axis=2)
        self.shape = (self.A.shape[0] * self.B.shape[0] * self.C.shape[0],
                      int((self.sigma_i >= 0).sum()))

    def to_dense(self):
        V = np.zeros(self.shape)
        n = self.sigma_i.shape[0]
        for b1 in range(n):
            for b2 in range(self.N2[b1]):
                for b3 in range(self.N3[b1, b2]):
                    col = np.einsum("a,b,c->abc", self.A[:, b1], self.B[:, b1, b2],
                                    self.C[:, b1, b2, b3])
                    V[self.sigma_o.ravel(), self.sigma_i[b1, b2, b3]] = col.ravel()
        return V
