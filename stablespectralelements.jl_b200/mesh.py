"""Periodic structured meshes, warping and facet connectivity (host-side setup).

Mirrors /root/reference/src/SpatialDiscretizations/mesh.jl:1-211,511-565
(``uniform_periodic_mesh``, ``warp_mesh`` with ``DelReyWarping``/``ChanWarping``/
``UniformWarping``) and the pieces of StartUpDG's ``MeshData``/``make_periodic`` the residual
needs: mapping nodes ``xyz``, volume/facet quadrature node coordinates ``xyzq``/``xyzf`` and
the facet-node connectivity ``mapP``.

StartUpDG is un-vendored, so the generator here is our own: 2 triangles per square split along
the lower-left/upper-right diagonal, 6 tetrahedra per cube (Kuhn split around the cell diagonal
(0,0,1)-(1,1,0), identified by matching the reference's golden Tet L2 errors) followed by the reference's collapsed-orientation vertex sort (mesh.jl:150-169),
Quad/Hex/Line cells as is.  ``mapP`` is 0-based and indexes the flattened (N_f, N_e) facet-node
array in column-major order (``j + N_f * k``), i.e. the reference's ``mesh.mapP`` minus one.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .reference_approximation import (Hex, Line, Quad, ReferenceApproximation, RefElemData,
                                      Tet, Tri)


def run_chunks(work, n: int, chunk: int):
    """``work(start, stop)`` over [0, n) in chunks on a thread pool.  The host-side setup is
    element-wise NumPy on large arrays, which releases the GIL; SSE_B200_SETUP_THREADS overrides
    the thread count (1 = serial)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    spans = [(s, min(s + chunk, n)) for s in range(0, n, chunk)]
    nthr = int(os.environ.get("SSE_B200_SETUP_THREADS", "0")) or min(16, os.cpu_count() or 1)
    if nthr <= 1 or len(spans) <= 1:
        for s, e in spans:
            work(s, e)
        return
    with ThreadPoolExecutor(max_workers=nthr) as pool:
        for f in [pool.submit(work, s, e) for s, e in spans]:
            f.result()                                   # re-raises a worker's exception


# ------------------------------------------------------------------------ warpings
@dataclass(frozen=True)
class DelReyWarping:
    factor: float
    L: Tuple[float, ...]


@dataclass(frozen=True)
class ChanWarping:
    factor: float
    L: Tuple[float, ...]


@dataclass(frozen=True)
class UniformWarping:
    factor: float
    L: Tuple[float, ...]


@dataclass
class MeshData:
    """Subset of StartUpDG.MeshData used by the residual setup."""
    VXYZ: Tuple[np.ndarray, ...]       # vertex coordinates
    EToV: np.ndarray                   # (N_e, n_vertices) 0-based
    xyz: Tuple[np.ndarray, ...]        # mapping nodes, each (N_map, N_e)
    xyzq: Tuple[np.ndarray, ...]       # (N_q, N_e)
    xyzf: Tuple[np.ndarray, ...]       # (N_f, N_e)
    mapP: np.ndarray                   # (N_f, N_e) int64, 0-based linear index j + N_f*k
    limits: Tuple[Tuple[float, float], ...]
    FToF: Optional[np.ndarray] = None  # (num_faces, N_e) linear face index f + num_faces*k
    # element shard (multi-GPU setup): this object holds elements [elem_start, elem_start+N_e)
    # of a larger mesh whose full connectivity is kept in ``mapP`` (N_f, N_e_global)
    elem_start: Optional[int] = None

    @property
    def N_e(self):
        return self.EToV.shape[0]

    @property
    def dim(self):
        return len(self.VXYZ)


# ------------------------------------------------------------- structured meshes
def cartesian_vertices(M, limits):
    """Lexicographic vertex grid, x fastest."""
    d = len(M)
    axes = [np.linspace(limits[m][0], limits[m][1], M[m] + 1) for m in range(d)]
    if d == 1:
        return (axes[0],)
    if d == 2:
        X, Y = np.meshgrid(axes[0], axes[1], indexing="ij")
        return X.ravel(order="F"), Y.ravel(order="F")
    X, Y, Z = np.meshgrid(*axes, indexing="ij")
    return X.ravel(order="F"), Y.ravel(order="F"), Z.ravel(order="F")


def _cell_corner_ids(M):
    d = len(M)
    if d == 1:
        ex = np.arange(M[0])
        return [ex, ex + 1]
    if d == 2:
        ex, ey = np.meshgrid(np.arange(M[0]), np.arange(M[1]), indexing="ij")
        ex, ey = ex.ravel(order="F"), ey.ravel(order="F")
        nx = M[0] + 1
        vid = lambda a, b: (ex + a) + (ey + b) * nx
        return [vid(0, 0), vid(1, 0), vid(0, 1), vid(1, 1)]
    ex, ey, ez = np.meshgrid(np.arange(M[0]), np.arange(M[1]), np.arange(M[2]), indexing="ij")
    ex, ey, ez = (a.ravel(order="F") for a in (ex, ey, ez))
    nx, ny = M[0] + 1, M[1] + 1
    vid = lambda a, b, c: (ex + a) + (ey + b) * nx + (ez + c) * nx * ny
    return [vid(a, b, c) for c in (0, 1) for b in (0, 1) for a in (0, 1)]


def cartesian_mesh(elem, M, limits):
    """Vertices + element-to-vertex table (cells in lexicographic order, x fastest)."""
    V = cartesian_vertices(M, limits)
    c = _cell_corner_ids(M)
    if isinstance(elem, Line):
        EToV = np.stack([c[0], c[1]], axis=1)
    elif isinstance(elem, Quad):
        EToV = np.stack(c, axis=1)
    elif isinstance(elem, Hex):
        EToV = np.stack(c, axis=1)
    elif isinstance(elem, Tri):
        # corners: c0=(0,0) c1=(1,0) c2=(0,1) c3=(1,1); diagonal c0-c3, both CCW.  This vertex
        # order (which fixes where the collapsed vertex of each triangle lies) reproduces the
        # reference's golden L2 errors (tests/test_oracle_golden.py), i.e. it is StartUpDG's.
        t1 = np.stack([c[3], c[2], c[0]], axis=1)
        t2 = np.stack([c[0], c[1], c[3]], axis=1)
        EToV = np.stack([t1, t2], axis=1).reshape(-1, 3)
    elif isinstance(elem, Tet):
        # Kuhn split: 6 tets around the cell diagonal (0,0,1)-(1,1,0), one per monotone edge path
        # between its ends; corner index = a + 2b + 4c.  StartUpDG's uniform_mesh(Tet(), ...) is
        # un-vendored; of the 4 possible diagonals x 6 vertex numberings only this one (with the
        # x-fastest, z-slowest vertex numbering used here) reproduces the reference's Tet golden
        # L2 errors (tests/test_oracle_golden.py; the others are off by 3e-3 .. 0.17 relative).
        start = 4
        tets = []
        for perm in itertools.permutations((1, 2, 4)):
            ids, cur = [start], start
            for bit in perm:
                cur ^= bit
                ids.append(cur)
            tets.append(np.stack([c[j] for j in ids], axis=1))
        EToV = np.stack(tets, axis=1).reshape(-1, 4)
    else:
        raise TypeError(elem)
    return V, np.ascontiguousarray(EToV, dtype=np.int64)


def _collapsed_orientation(VXYZ, EToV):
    """mesh.jl:150-169: sort vertex ids descending, swap the first two if det < 0."""
    E = -np.sort(-EToV, axis=1)
    P = np.stack([v[E] for v in VXYZ], axis=2)            # (N_e, 4, 3)
    X = P[:, 1:, :] - P[:, :1, :]                         # rows = edge vectors
    det = np.linalg.det(X)
    neg = det < 0
    E[neg, 0], E[neg, 1] = E[neg, 1].copy(), E[neg, 0].copy()
    return E


def _fix_orientation(elem, VXYZ, EToV):
    if isinstance(elem, Tri):
        P = np.stack([v[EToV] for v in VXYZ], axis=2)
        a, b = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]
        det = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
        assert np.all(det > 0), "negatively oriented triangle"
    return EToV


# ----------------------------------------------------------------- connectivity
def _periodic_diff(a, b, L):
    dlt = a - b
    for m, Lm in enumerate(L):
        dlt[..., m] = (dlt[..., m] + 0.5 * Lm) % Lm - 0.5 * Lm
    return dlt


def build_face_connectivity(elem, VXYZ, EToV, limits, fv):
    """Match element faces of the (straight, periodic) mesh by face centroid.

    Returns FToF (num_faces, N_e): linear index ``f' + num_faces*k'`` of the face glued to
    face f of element k.
    """
    d = len(VXYZ)
    N_e = EToV.shape[0]
    nfaces = len(fv)
    L = np.array([limits[m][1] - limits[m][0] for m in range(d)])
    lo = np.array([limits[m][0] for m in range(d)])
    P = np.stack([v[EToV] for v in VXYZ], axis=2)          # (N_e, nv, d)
    cent = np.stack([P[:, list(f), :].mean(axis=1) for f in fv], axis=0)  # (nfaces, N_e, d)
    # integer key on a fine lattice: face centroids of a structured mesh sit on multiples
    # of h_m/6 in each direction, so h_m/12 resolves them exactly (also across the period)
    dP = np.abs(P[:, 1:, :] - P[:, :1, :])
    step = np.array([np.min(dP[..., m][dP[..., m] > 1e-12 * L[m]]) for m in range(d)]) / 12.0
    nper = np.rint(L / step).astype(np.int64)
    key3 = np.rint((cent - lo) / step).astype(np.int64) % nper
    key = np.zeros(key3.shape[:2], dtype=np.int64)
    for m in range(d):
        key = key * nper[m] + key3[..., m]
    flat = key.ravel()                                     # index = f*N_e + k
    order = np.argsort(flat, kind="stable")
    sk = flat[order]
    if len(sk) % 2 or not np.all(sk[0::2] == sk[1::2]) or np.any(sk[1:-1:2] == sk[2::2]):
        raise RuntimeError("face matching failed (mesh too coarse for periodic matching?)")
    a, b = order[0::2], order[1::2]
    partner = np.empty_like(order)
    partner[a], partner[b] = b, a
    pf, pk = partner // N_e, partner % N_e
    FToF = (pf + nfaces * pk).reshape(nfaces, N_e)
    return FToF


def build_mapP(xyzf, FToF, nfaces, limits, tol=1e-7, chunk=65536):
    """Facet-node permutation between glued faces by (periodic) coordinate matching."""
    d = len(xyzf)
    N_f, N_e = xyzf[0].shape
    npf = N_f // nfaces
    L = np.array([limits[m][1] - limits[m][0] for m in range(d)])
    X = np.stack(xyzf, axis=2).reshape(nfaces, npf, N_e, d)   # [f, a, k, :]
    X = np.ascontiguousarray(X.transpose(2, 0, 1, 3)).reshape(N_e * nfaces, npf, d)
    partner = FToF.T.reshape(-1)                              # index k*nfaces+f -> f'+nfaces*k'
    pk, pf = partner // nfaces, partner % nfaces
    pidx = pk * nfaces + pf
    mapP = np.empty((N_e * nfaces, npf), dtype=np.int64)
    scale = float(np.max(L))

    def work(s, e):
        A = X[s:e]                                            # (C, npf, d)
        Bn = X[pidx[s:e]]
        # shift the partner face by whole periods so the two faces coincide
        shift = np.rint((A.mean(axis=1) - Bn.mean(axis=1)) / L) * L
        Bn = Bn + shift[:, None, :]
        # |a-b|^2 = |a|^2 + |b|^2 - 2 a.b  (centred to keep round-off small)
        c0 = A.mean(axis=1, keepdims=True)
        Ac, Bc = A - c0, Bn - c0
        dist = (np.einsum("cam,cam->ca", Ac, Ac)[:, :, None]
                + np.einsum("cbm,cbm->cb", Bc, Bc)[:, None, :]
                - 2.0 * np.matmul(Ac, Bc.transpose(0, 2, 1)))
        arg = np.argmin(dist, axis=2)
        best = np.take_along_axis(dist, arg[:, :, None], axis=2)[:, :, 0]
        if np.max(best) > (tol * scale) ** 2:
            raise RuntimeError(f"facet nodes do not conform (max dist {np.sqrt(np.max(best)):.3e})")
        if npf > 1 and np.any(np.sort(arg, axis=1) != np.arange(npf)[None, :]):
            raise RuntimeError("facet node matching is not a permutation")
        mapP[s:e] = arg + (pf[s:e] * npf)[:, None] + (pk[s:e] * N_f)[:, None]

    run_chunks(work, N_e * nfaces, min(chunk, 16384))
    # back to (N_f, N_e)
    return np.ascontiguousarray(mapP.reshape(N_e, nfaces * npf).T)


# ------------------------------------------------------------------- constructors
def _make_mesh(re: RefElemData, VXYZ, EToV, limits, xyz=None, FToF=None, mapP=None,
               straight_xyzf=None):
    elem = re.element_type
    d = elem.dim
    if xyz is None:
        xyz = tuple(re.V1 @ v[EToV].T for v in VXYZ)
    xyzq = tuple(re.Vq @ x for x in xyz)
    xyzf = tuple(re.Vf @ x for x in xyz)
    if FToF is False:
        FToF = None
    elif FToF is None:
        FToF = build_face_connectivity(elem, VXYZ, EToV, limits, re.fv)
    if mapP is None:
        mapP = build_mapP(straight_xyzf if straight_xyzf is not None else xyzf, FToF,
                          elem.num_faces, limits)
    return MeshData(tuple(VXYZ), EToV, xyz, xyzq, xyzf, mapP, tuple(limits), FToF)


def uniform_periodic_mesh(reference, limits, M, collapsed_orientation=None) -> MeshData:
    """mesh.jl:122-181,511-565.  ``reference`` is a ReferenceApproximation or RefElemData."""
    re = reference.reference_element if isinstance(reference, ReferenceApproximation) \
        else reference
    elem = re.element_type
    if elem.dim == 1:
        if isinstance(limits[0], (int, float)):
            limits = (tuple(limits),)
        M = (int(M),) if np.isscalar(M) else tuple(M)
    limits = tuple(tuple(float(x) for x in lim) for lim in limits)
    M = tuple(int(m) for m in M)
    if min(M) < 2:
        raise ValueError("periodic matching needs at least 2 cells per direction")
    VXYZ, EToV = cartesian_mesh(elem, M, limits)
    if isinstance(elem, Tet):
        tensor = isinstance(reference, ReferenceApproximation) and \
            type(reference.approx_type).__name__ in ("NodalTensor", "ModalTensor")
        if collapsed_orientation if collapsed_orientation is not None else tensor:
            EToV = _collapsed_orientation(VXYZ, EToV)
    EToV = _fix_orientation(elem, VXYZ, EToV)
    return _make_mesh(re, VXYZ, EToV, limits)


def _warp_coordinates(x, w):
    d = len(x)
    L, f = w.L, w.factor
    pi = np.pi
    if isinstance(w, DelReyWarping):
        if d == 2:
            X, Y = x
            xn = X + L[0] * f * np.sin(pi * X / L[0]) * np.sin(pi * Y / L[1])
            yn = Y + L[1] * f * np.exp(1.0 - Y / L[1]) * np.sin(pi * X / L[0]) * np.sin(pi * Y / L[1])
            return xn, yn
        X, Y, Z = x
        xn = X + L[0] * f * np.sin(pi * X / L[0]) * np.sin(pi * Y / L[1])
        yn = Y + L[1] * f * np.exp((1.0 - Y) / L[1]) * np.sin(pi * X / L[0]) * np.sin(pi * Y / L[1])
        zn = Z + 0.25 * L[2] * f * (np.sin(2 * pi * X / L[0]) * np.sin(2 * pi * Y / L[1])) \
            * np.sin(2 * pi * Z / L[2])
        return xn, yn, zn
    if isinstance(w, ChanWarping):
        if d == 2:
            X, Y = x
            xn = X + L[0] * f * np.cos(pi / L[0] * (X - 0.5 * L[0])) \
                * np.cos(3 * pi / L[1] * (Y - 0.5 * L[1]))
            yn = Y + L[1] * f * np.sin(4 * pi / L[0] * (xn - 0.5 * L[0])) \
                * np.cos(pi / L[1] * (Y - 0.5 * L[1]))
            return xn, yn
        X, Y, Z = x
        yn = Y + L[1] * f * np.cos(3 * pi / L[0] * (X - 0.5 * L[0])) \
            * np.cos(pi / L[1] * (Y - 0.5 * L[1])) * np.cos(pi / L[2] * (Z - 0.5 * L[2]))
        xn = X + L[0] * f * np.cos(pi / L[0] * (X - 0.5 * L[0])) \
            * np.sin(4 * pi / L[1] * (yn - 0.5 * L[1])) * np.cos(pi / L[2] * (Z - 0.5 * L[2]))
        zn = Z + L[2] * f * np.cos(pi / L[0] * (xn - 0.5 * L[0])) \
            * np.cos(2 * pi / L[1] * (yn - 0.5 * L[1])) * np.cos(pi / L[2] * (Z - 0.5 * L[2]))
        return xn, yn, zn
    if isinstance(w, UniformWarping):
        eps = f * np.ones_like(x[0])
        for m in range(d):
            eps = eps * np.sin(2 * pi * (x[m] - L[m] / 2) / L[m])
        return tuple(x[m] + L[m] * eps for m in range(d))
    raise TypeError(w)


def mesh_subset(mesh: MeshData, start: int, stop: int) -> MeshData:
    """Shard [start, stop) of the elements: node coordinates are sliced, the connectivity
    ``mapP`` stays global (the partitioner needs it), ``elem_start`` records the offset."""
    sl = slice(start, stop)
    cut = lambda t: tuple(np.ascontiguousarray(x[:, sl]) for x in t)
    return MeshData(mesh.VXYZ, mesh.EToV[sl], cut(mesh.xyz), cut(mesh.xyzq), cut(mesh.xyzf),
                    mesh.mapP, mesh.limits, None, elem_start=start)


def warp_mesh(mesh: MeshData, reference, warping=0.2, L: float = 1.0) -> MeshData:
    """mesh.jl:23-120: apply the warp to the mapping nodes and rebuild node coordinates.

    The warps vanish on the domain boundary planes they need to, so the straight mesh's
    connectivity (``mapP``) stays valid and is re-used.
    """
    re = reference.reference_element if isinstance(reference, ReferenceApproximation) \
        else reference
    d = mesh.dim
    if isinstance(warping, (int, float)):
        warping = DelReyWarping(float(warping), tuple(float(L) for _ in range(d)))
    xyz_new = _warp_coordinates(mesh.xyz, warping)
    out = _make_mesh(re, mesh.VXYZ, mesh.EToV, mesh.limits, xyz=tuple(xyz_new),
                     FToF=mesh.FToF if mesh.FToF is not None else False, mapP=mesh.mapP)
    out.elem_start = mesh.elem_start
    return out
