"""Geometric factors and the ``SpatialDiscretization`` bundle (host-side setup).

Mirrors /root/reference/src/SpatialDiscretizations/mesh.jl:213-509 (``metrics``,
``GeometricFactors`` for ``ExactMetrics`` and ``ConservativeCurlMetrics``) and
SpatialDiscretizations.jl:248-423 (``GeometricFactors`` struct, ``SpatialDiscretization``
constructors with Jacobian projection, ``apply_reference_mapping``, self-checks).  Everything
is vectorised over elements; arrays keep the reference's index order with the element index
last, stored as C-contiguous NumPy arrays of shape (N_e, ...) *reversed* -- see ``layout``
below -- so that a flat view equals Julia's column-major memory.

Layout convention used throughout the host package: a Julia array ``A[i1, i2, ..., k]`` is a
NumPy array ``A[k, ..., i2, i1]`` (C order), i.e. identical bytes.  E.g. ``J_q`` is (N_e, N_q),
``Lambda_q`` is (N_e, d_n, d_m, N_q) for Julia's ``Λ_q[i, m, n, k]``, ``nJf`` is (N_e, N_f, d).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import nodes as nd
from . import polynomials as poly
from .mesh import MeshData, run_chunks
from .reference_approximation import (Hex, NoMapping, ReferenceApproximation, ReferenceMapping,
                                      RefElemData, Tet, Tri, reference_vertices, vandermonde)


class ExactMetrics:
    pass


class ConservativeCurlMetrics:
    pass


ChanWilcoxMetrics = ConservativeCurlMetrics


@dataclass
class GeometricFactors:
    """SpatialDiscretizations.jl:248-283 (memory-identical layouts, see module docstring)."""
    J_q: np.ndarray        # (N_e, N_q)
    Lambda_q: np.ndarray   # (N_e, d[n], d[m], N_q)   == Julia Λ_q[i, m, n, k]
    J_f: np.ndarray        # (N_e, N_f)
    nJf: np.ndarray        # (N_e, N_f, d)            == Julia nJf[m, i, k]
    n_ref: np.ndarray      # (num_faces, d) reference normal of each face (nrstJ at first node)

    def nJq(self):
        """Julia nJq[n, f, i, k] = Σ_m Λ_q[i,m,n,k] n_ref[m,f] (mesh.jl:266-271) as
        (N_e, N_q, num_faces, d)."""
        return np.einsum("knmi,fm->kifn", self.Lambda_q, self.n_ref)


def _metrics_from_dxdr(dxdr):
    """mesh.jl:213-229.  dxdr[..., m, n] = ∂x_m/∂ξ_n ; returns J, Λ[..., l, m] = J ∂ξ_l/∂x_m."""
    d = dxdr.shape[-1]
    if d == 1:
        return dxdr[..., 0, 0].copy(), np.ones_like(dxdr)
    if d == 2:
        J = dxdr[..., 0, 0] * dxdr[..., 1, 1] - dxdr[..., 0, 1] * dxdr[..., 1, 0]
        L = np.empty_like(dxdr)
        L[..., 0, 0] = dxdr[..., 1, 1]
        L[..., 0, 1] = -dxdr[..., 0, 1]
        L[..., 1, 0] = -dxdr[..., 1, 0]
        L[..., 1, 1] = dxdr[..., 0, 0]
        return J, L
    a = dxdr
    L = np.empty_like(a)
    # adjugate (= J * inverse)
    L[..., 0, 0] = a[..., 1, 1] * a[..., 2, 2] - a[..., 1, 2] * a[..., 2, 1]
    L[..., 0, 1] = a[..., 0, 2] * a[..., 2, 1] - a[..., 0, 1] * a[..., 2, 2]
    L[..., 0, 2] = a[..., 0, 1] * a[..., 1, 2] - a[..., 0, 2] * a[..., 1, 1]
    L[..., 1, 0] = a[..., 1, 2] * a[..., 2, 0] - a[..., 1, 0] * a[..., 2, 2]
    L[..., 1, 1] = a[..., 0, 0] * a[..., 2, 2] - a[..., 0, 2] * a[..., 2, 0]
    L[..., 1, 2] = a[..., 0, 2] * a[..., 1, 0] - a[..., 0, 0] * a[..., 1, 2]
    L[..., 2, 0] = a[..., 1, 0] * a[..., 2, 1] - a[..., 1, 1] * a[..., 2, 0]
    L[..., 2, 1] = a[..., 0, 1] * a[..., 2, 0] - a[..., 0, 0] * a[..., 2, 1]
    L[..., 2, 2] = a[..., 0, 0] * a[..., 1, 1] - a[..., 0, 1] * a[..., 1, 0]
    J = a[..., 0, 0] * L[..., 0, 0] + a[..., 0, 1] * L[..., 1, 0] + a[..., 0, 2] * L[..., 2, 0]
    return J, L


def _n_ref(re: RefElemData):
    nfaces = re.element_type.num_faces
    npf = len(re.wf) // nfaces
    return np.array([[re.nrstJ[m][npf * f] for m in range(re.dim)] for f in range(nfaces)])


def _facet_normals(Lambda_f, re: RefElemData):
    """nJf[m,i] = Σ_n Λ_f[i,n,m] nrstJ[n][i]; J_f = |nJf| (mesh.jl:273-281).
    Lambda_f: (N_e, N_f, d[l], d[m])."""
    nr = np.stack(re.nrstJ, axis=1)                      # (N_f, d)
    N_e, N_f, d = Lambda_f.shape[:3]
    nJf = np.empty((N_e, N_f, d))
    J_f = np.empty((N_e, N_f))

    def work(s, e):
        nJf[s:e] = np.einsum("kinm,in->kim", Lambda_f[s:e], nr)
        J_f[s:e] = np.sqrt(np.sum(nJf[s:e] ** 2, axis=2))

    run_chunks(work, N_e, 8192)
    return nJf, J_f


def geometric_factors_exact(mesh: MeshData, re: RefElemData) -> GeometricFactors:
    """mesh.jl:231-284."""
    d = re.dim
    N_e = mesh.N_e
    N_q, N_f = re.Vq.shape[0], re.Vf.shape[0]
    dq = np.empty((N_e, N_q, d, d))
    df = np.empty((N_e, N_f, d, d))
    for m in range(d):
        for n in range(d):
            dxdr = re.Drst[n] @ mesh.xyz[m]              # (N_map, N_e)
            dq[:, :, m, n] = (re.Vq @ dxdr).T
            df[:, :, m, n] = (re.Vf @ dxdr).T
    J_q, Lq = _metrics_from_dxdr(dq)                      # Lq (N_e, N_q, l, m)
    _, Lf = _metrics_from_dxdr(df)
    nJf, J_f = _facet_normals(Lf, re)
    Lambda_q = np.ascontiguousarray(Lq.transpose(0, 3, 2, 1))   # (N_e, n=m_idx2, m=l, N_q)
    return GeometricFactors(np.ascontiguousarray(J_q), Lambda_q, J_f, nJf, _n_ref(re))


def _curl_metrics_3d(x, y, z, Dr, Ds, Dt):
    """StartUpDG ``geometric_factors(x,y,z,Dr,Ds,Dt)`` (un-vendored): conservative curl form of
    Kopriva (2006); returns the 9 scaled metric terms (as Λ[l][m] = J ∂ξ_l/∂x_m) and J."""
    xr, xs, xt = Dr @ x, Ds @ x, Dt @ x
    yr, ys, yt = Dr @ y, Ds @ y, Dt @ y
    zr, zs, zt = Dr @ z, Ds @ z, Dt @ z

    def curl(Fr, Fs, Ft):
        return (Dt @ Fs - Ds @ Ft, Dr @ Ft - Dt @ Fr, Ds @ Fr - Dr @ Fs)

    rxJ, sxJ, txJ = curl(yr * z, ys * z, yt * z)
    ryJ, syJ, tyJ = (-c for c in curl(xr * z, xs * z, xt * z))
    rzJ, szJ, tzJ = (-c for c in curl(yr * x, ys * x, yt * x))
    J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
    return ((rxJ, ryJ, rzJ), (sxJ, syJ, szJ), (txJ, tyJ, tzJ)), J


def geometric_factors_curl(mesh: MeshData, re: RefElemData) -> GeometricFactors:
    """mesh.jl:286-509 (2-D, Hex and Tet variants).  Elements are independent: chunks of them
    are evaluated on a thread pool (mesh.run_chunks)."""
    d = re.dim
    elem = re.element_type
    N_e = mesh.N_e
    if d == 1:
        return geometric_factors_exact(mesh, re)
    tet = isinstance(elem, Tet)
    if d == 3 and not tet and not isinstance(elem, Hex):
        raise TypeError(elem)
    if tet:
        # the argument of the curl is a degree N+1 polynomial: evaluate it on a degree N+1
        # nodal set, bring the (degree N) result back to the degree N nodes (mesh.jl:413-470)
        N = re.N
        r1, s1, t1 = nd.nodes_tet(N + 1)
        V1, Vr1, Vs1, Vt1 = poly.simplex_basis_3d(N + 1, r1, s1, t1, grad=True)
        N_to_Np1 = np.linalg.solve(re.VDM.T, vandermonde(elem, N, r1, s1, t1).T).T
        Np1_to_N = np.linalg.solve(V1.T, vandermonde(elem, N + 1, *re.rst).T).T
        Vq, Vf = re.Vq @ Np1_to_N, re.Vf @ Np1_to_N
        D1 = tuple(np.linalg.solve(V1.T, g.T).T for g in (Vr1, Vs1, Vt1))
    else:
        Vq, Vf = re.Vq, re.Vf
    N_q, N_f = Vq.shape[0], Vf.shape[0]
    J_q = np.empty((N_e, re.Vq.shape[0]))
    Lambda_q = np.empty((N_e, d, d, N_q))         # [k, n, m, i]  (m = ξ index, n = x index)
    Lf = np.empty((N_e, N_f, d, d))               # [k, i, l, m]

    def work(s, e):
        xyz = tuple(c[:, s:e] for c in mesh.xyz)
        if d == 2:
            x, y = xyz
            Dr, Ds = re.Drst
            xr, xs, yr, ys = Dr @ x, Ds @ x, Dr @ y, Ds @ y
            J = -xs * yr + xr * ys
            L = ((ys, -xs), (-yr, xr))               # L[l][m]: rxJ ryJ / sxJ syJ
        elif tet:
            _, J = _curl_metrics_3d(*xyz, *re.Drst)
            L, _ = _curl_metrics_3d(*(N_to_Np1 @ c for c in xyz), *D1)
        else:
            L, J = _curl_metrics_3d(*xyz, *re.Drst)
        J_q[s:e] = (re.Vq @ J).T
        for l in range(d):
            for m in range(d):
                Lambda_q[s:e, m, l, :] = (Vq @ L[l][m]).T
                Lf[s:e, :, l, m] = (Vf @ L[l][m]).T

    run_chunks(work, N_e, 8192)
    nJf, J_f = _facet_normals(Lf, re)
    return GeometricFactors(J_q, Lambda_q, J_f, nJf, _n_ref(re))


class DeviceGeometricFactors:
    """GeometricFactors evaluated and kept on the device (``sse_geometry_build``, SURVEY.md §8f
    item 2; mesh.jl:213-509).  ``sse_create`` takes the device pointers directly; host copies of
    the arrays (same layouts as ``GeometricFactors``) are downloaded lazily, only if some host-side
    consumer (initial-data projection, the test oracle) asks for them."""
    on_device = True

    def __init__(self, lib, geo, shapes, n_ref):
        self._lib, self._geo, self._shapes, self.n_ref = lib, geo, shapes, n_ref
        self._host = {}

    def device_pointers(self):
        g = self._geo
        return g.J_q, g.Lambda_q, g.J_f, g.nJf

    def _fetch(self, name):
        if name not in self._host:
            import ctypes as C
            out = np.empty(self._shapes[name])
            ptr = C.cast(getattr(self._geo, name), C.c_void_p)
            if self._lib.sse_copy_to_host(out.ctypes.data, ptr, out.nbytes) != 0:
                raise RuntimeError("sse_copy_to_host failed: " + self._lib.sse_last_error().decode())
            self._host[name] = out
        return self._host[name]

    J_q = property(lambda self: self._fetch("J_q"))
    Lambda_q = property(lambda self: self._fetch("Lambda_q"))
    J_f = property(lambda self: self._fetch("J_f"))
    nJf = property(lambda self: self._fetch("nJf"))

    def nJq(self):
        return np.einsum("knmi,fm->kifn", self.Lambda_q, self.n_ref)

    def free(self):
        if self._geo is not None:
            self._lib.sse_geometry_free(self._geo)
            self._geo = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def jacobian_projection_matrix(V, W):
    """SpatialDiscretizations.jl:311-318 as an (N_q x N_q) matrix acting on J_q."""
    VDM = V.to_dense()
    return VDM @ np.linalg.solve(VDM.T @ (W[:, None] * VDM), VDM.T * W[None, :])


def geometric_factors_device(mesh: MeshData, re: RefElemData, metric_type=None, Jproj=None,
                             device: int = 0) -> DeviceGeometricFactors:
    """Same quantities as ``make_geometric_factors`` but evaluated by the CUDA library from the
    mapping-node coordinates (no CPU fallback: raises if the library or a GPU is missing)."""
    import ctypes as C
    from . import device as dev
    lib = dev.load_library()
    d = re.dim
    elem = re.element_type
    exact = metric_type is None or isinstance(metric_type, ExactMetrics)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    keep = []

    def put(a):
        a = f64(a)
        keep.append(a)
        return dp(a)

    m = dev.SseMapping()
    m.dim, m.device = d, device
    m.N_q, m.N_f = re.Vq.shape[0], re.Vf.shape[0]
    m.N_e = mesh.N_e
    m.N_map = m.N_map1 = re.Vq.shape[1]
    for n in range(d):
        m.D[n] = put(re.Drst[n])
        m.xyz[n] = put(np.asarray(mesh.xyz[n]).T)          # (N_e, N_map) == Julia (N_map, N_e)
    m.Vq, m.Vf = put(re.Vq), put(re.Vf)
    m.nrstJ = put(np.stack(re.nrstJ, axis=1))
    if exact or d == 1:
        m.metric = 0
        if Jproj is not None:
            m.Jproj = put(Jproj)
    else:
        if d != 3:
            raise NotImplementedError("device curl-form metrics are implemented for d = 3")
        m.metric = 1
        if isinstance(elem, Tet):
            N = re.N
            r1, s1, t1 = nd.nodes_tet(N + 1)
            V1, Vr1, Vs1, Vt1 = poly.simplex_basis_3d(N + 1, r1, s1, t1, grad=True)
            N_to_Np1 = np.linalg.solve(re.VDM.T, vandermonde(elem, N, r1, s1, t1).T).T
            Np1_to_N = np.linalg.solve(V1.T, vandermonde(elem, N + 1, *re.rst).T).T
            m.N_map1 = V1.shape[0]
            m.P = put(N_to_Np1)
            for n, g in enumerate((Vr1, Vs1, Vt1)):
                m.D1[n] = put(np.linalg.solve(V1.T, g.T).T)
            m.Vq1, m.Vf1 = put(re.Vq @ Np1_to_N), put(re.Vf @ Np1_to_N)
        elif isinstance(elem, Hex):
            for n in range(3):
                m.D1[n] = m.D[n]
            m.Vq1, m.Vf1 = m.Vq, m.Vf
        else:
            raise TypeError(elem)
    geo = dev.SseGeometry()
    if lib.sse_geometry_build(C.byref(m), C.byref(geo)) != 0:
        raise RuntimeError("sse_geometry_build failed: " + lib.sse_last_error().decode())
    N_e, N_q, N_f = mesh.N_e, int(m.N_q), int(m.N_f)
    shapes = {"J_q": (N_e, N_q), "Lambda_q": (N_e, d, d, N_q), "J_f": (N_e, N_f),
              "nJf": (N_e, N_f, d)}
    return DeviceGeometricFactors(lib, geo, shapes, _n_ref(re))


def make_geometric_factors(mesh, re, metric_type=None) -> GeometricFactors:
    if metric_type is None or isinstance(metric_type, ExactMetrics):
        return geometric_factors_exact(mesh, re)
    return geometric_factors_curl(mesh, re)


def project_jacobian(J_q, V, W):
    """SpatialDiscretizations.jl:311-318: L2 projection of J onto the solution space."""
    VDM = V.to_dense()
    proj = VDM @ np.linalg.solve(VDM.T @ (W[:, None] * VDM), VDM.T * W[None, :])
    return np.ascontiguousarray(J_q @ proj.T)


@dataclass
class SpatialDiscretization:
    """SpatialDiscretizations.jl:292-393 (without plotting nodes / dense mass matrices)."""
    mesh: MeshData
    reference_approximation: ReferenceApproximation
    geometric_factors: GeometricFactors

    @property
    def N_e(self):
        return self.mesh.N_e

    @property
    def dim(self):
        return self.reference_approximation.dim


def make_spatial_discretization(mesh: MeshData, ra: ReferenceApproximation, metric_type=None,
                                project_jacobian_flag: Optional[bool] = None,
                                device_geometry: Optional[int] = None
                                ) -> SpatialDiscretization:
    """``SpatialDiscretization(mesh, ra[, metric_type]; project_jacobian=true)``.

    ExactMetrics projects the Jacobian by default; the ChanWilcox constructor never does
    (SpatialDiscretizations.jl:334-393)."""
    exact = metric_type is None or isinstance(metric_type, ExactMetrics)
    if project_jacobian_flag is None:
        project_jacobian_flag = exact
    if device_geometry is not None:
        # ``device_geometry`` = CUDA device index: evaluate the geometric factors there
        Jproj = (jacobian_projection_matrix(ra.V, ra.W) if exact and project_jacobian_flag
                 else None)
        gf = geometric_factors_device(mesh, ra.reference_element, metric_type, Jproj,
                                      device_geometry)
        return SpatialDiscretization(mesh, ra, gf)
    gf = make_geometric_factors(mesh, ra.reference_element, metric_type)
    if exact and project_jacobian_flag:
        gf.J_q = project_jacobian(gf.J_q, ra.V, ra.W)
    return SpatialDiscretization(mesh, ra, gf)


def apply_reference_mapping(gf: GeometricFactors, reference_mapping) -> np.ndarray:
    """SpatialDiscretizations.jl:396-412: Λ_η[i,m,n,k] = Σ_l Λ_ref[i,m,l] Λ_q[i,l,n,k]/J_ref[i];
    returns the (N_e, d[n], d[m], N_q) array."""
    if isinstance(reference_mapping, NoMapping):
        return gf.Lambda_q
    Lr, Jr = reference_mapping.Lambda_ref, reference_mapping.J_ref
    return np.einsum("iml,knli->knmi", Lr / Jr[:, None, None], gf.Lambda_q)


def check_normals(sd: SpatialDiscretization):
    """SpatialDiscretizations.jl:426-432."""
    nJf = sd.geometric_factors.nJf
    flat = nJf.reshape(-1, nJf.shape[2])
    mp = sd.mesh.mapP.T.reshape(-1)                       # [k, j] -> j + N_f k
    return np.max(np.abs(flat + flat[mp]))


def check_facet_nodes(sd: SpatialDiscretization):
    """SpatialDiscretizations.jl:434-439 (modulo the period)."""
    mesh = sd.mesh
    err = 0.0
    mp = mesh.mapP.ravel(order="F")
    for m in range(mesh.dim):
        L = mesh.limits[m][1] - mesh.limits[m][0]
        xf = mesh.xyzf[m].ravel(order="F")
        dlt = (xf - xf[mp] + 0.5 * L) % L - 0.5 * L
        err = max(err, float(np.max(np.abs(dlt))))
    return err


def check_metric_identities(sd: SpatialDiscretization):
    """Discrete free-stream preservation: max |Σ_l D_ξl Λ_q[:, l, m]| over elements."""
    from .reference_approximation import reference_derivative_operators
    ra = sd.reference_approximation
    D_xi = reference_derivative_operators(ra.D, ra.reference_mapping)
    L = sd.geometric_factors.Lambda_q                      # (k, n, m, i)
    out = 0.0
    for n in range(sd.dim):
        acc = sum(np.einsum("ij,kj->ki", D_xi[l], L[:, n, l, :]) for l in range(sd.dim))
        out = max(out, float(np.max(np.abs(acc))))
    return out
