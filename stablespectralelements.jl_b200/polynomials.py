"""Orthonormal Jacobi polynomials, 1-D quadrature rules and simplex / tensor bases.

Host-side setup only (runs once, NumPy).  These restate the pieces of the un-vendored
third-party packages the reference calls when it builds a ``ReferenceApproximation``:

* ``jacobiP`` / ``grad_jacobiP``  -- NodesAndModes.jl ``jacobiP`` (orthonormal w.r.t. the
  weight (1-x)^a (1+x)^b; the Hesthaven-Warburton recurrence), used at
  /root/reference/src/SpatialDiscretizations/tensor_simplex.jl:84-140.
* ``gauss_jacobi`` / ``gauss_lobatto`` / ``gauss_radau``  -- Jacobi.jl ``zgj/wgj``, ``zglj/wglj``,
  ``zgrjm/wgrjm`` (quadrature_rules.jl:76-92).
* ``simplex_basis_2d/3d``  -- StartUpDG/NodesAndModes ``basis(Tri()/Tet(), N, ...)``: the
  Dubiner/Koornwinder orthonormal basis (ordering i, j[, k] with i slowest), the same
  functions the warped tensor product of tensor_simplex.jl:84-140 factorises.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import roots_jacobi


# ----------------------------------------------------------------------------- Jacobi
def jacobiP(x, alpha: float, beta: float, N: int) -> np.ndarray:
    """Orthonormal Jacobi polynomial P_N^{(alpha,beta)}(x) (Hesthaven-Warburton form)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    PL = np.zeros((N + 1, x.size))
    gamma0 = (2.0 ** (alpha + beta + 1) / (alpha + beta + 1)
              * math.gamma(alpha + 1) * math.gamma(beta + 1) / math.gamma(alpha + beta + 1))
    PL[0] = 1.0 / math.sqrt(gamma0)
    if N == 0:
        return PL[0].copy()
    gamma1 = (alpha + 1) * (beta + 1) / (alpha + beta + 3) * gamma0
    PL[1] = ((alpha + beta + 2) * x / 2 + (alpha - beta) / 2) / math.sqrt(gamma1)
    if N == 1:
        return PL[1].copy()
    aold = 2.0 / (2 + alpha + beta) * math.sqrt((alpha + 1) * (beta + 1) / (alpha + beta + 3))
    for i in range(1, N):
        h1 = 2 * i + alpha + beta
        anew = 2.0 / (h1 + 2) * math.sqrt((i + 1) * (i + 1 + alpha + beta) * (i + 1 + alpha)
                                          * (i + 1 + beta) / (h1 + 1) / (h1 + 3))
        bnew = -(alpha ** 2 - beta ** 2) / h1 / (h1 + 2)
        PL[i + 1] = 1.0 / anew * (-aold * PL[i - 1] + (x - bnew) * PL[i])
        aold = anew
    return PL[N].copy()


def grad_jacobiP(x, alpha: float, beta: float, N: int) -> np.ndarray:
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if N == 0:
        return np.zeros_like(x)
    return math.sqrt(N * (N + alpha + beta + 1)) * jacobiP(x, alpha + 1, beta + 1, N - 1)


# ------------------------------------------------------------------ 1-D quadrature
def gauss_jacobi(n: int, a: float = 0.0, b: float = 0.0):
    """n-point Gauss-Jacobi rule, nodes ascending (Jacobi.jl zgj/wgj)."""
    z, w = roots_jacobi(n, a, b)
    return np.asarray(z, dtype=np.float64), np.asarray(w, dtype=np.float64)


def gauss_lobatto(n: int, a: float = 0.0, b: float = 0.0):
    """n-point Gauss-Lobatto-Jacobi rule (Jacobi.jl zglj/wglj); n >= 2."""
    if n < 2:
        raise ValueError("Gauss-Lobatto needs at least 2 nodes")
    if n == 2:
        zi = np.zeros(0)
    else:
        zi, _ = roots_jacobi(n - 2, a + 1, b + 1)
    z = np.concatenate(([-1.0], zi, [1.0]))
    # weights from exactness: solve the moment equations in an orthonormal basis
    # (degree 2n-3 exactness fixes them uniquely).
    V = np.stack([jacobiP(z, a, b, j) for j in range(n)], axis=1)  # n x n
    # int P_j w(x) dx = delta_{j0} * sqrt(gamma0)
    gamma0 = (2.0 ** (a + b + 1) / (a + b + 1) * math.gamma(a + 1) * math.gamma(b + 1)
              / math.gamma(a + b + 1))
    rhs = np.zeros(n)
    rhs[0] = math.sqrt(gamma0)
    w = np.linalg.solve(V.T, rhs)
    return z, w


def gauss_radau(n: int, a: float = 0.0, b: float = 0.0):
    """n-point Gauss-Radau-Jacobi rule including x=-1 (Jacobi.jl zgrjm/wgrjm)."""
    if n == 1:
        zi = np.zeros(0)
    else:
        zi, _ = roots_jacobi(n - 1, a, b + 1)
    z = np.concatenate(([-1.0], zi))
    V = np.stack([jacobiP(z, a, b, j) for j in range(n)], axis=1)
    gamma0 = (2.0 ** (a + b + 1) / (a + b + 1) * math.gamma(a + 1) * math.gamma(b + 1)
              / math.gamma(a + b + 1))
    rhs = np.zeros(n)
    rhs[0] = math.sqrt(gamma0)
    w = np.linalg.solve(V.T, rhs)
    return z, w


# ------------------------------------------------------------------------- Line
def vandermonde_1d(N: int, r) -> np.ndarray:
    r = np.atleast_1d(np.asarray(r, dtype=np.float64))
    return np.stack([jacobiP(r, 0, 0, j) for j in range(N + 1)], axis=1)


def grad_vandermonde_1d(N: int, r) -> np.ndarray:
    r = np.atleast_1d(np.asarray(r, dtype=np.float64))
    return np.stack([grad_jacobiP(r, 0, 0, j) for j in range(N + 1)], axis=1)


# -------------------------------------------------------------------------- Tri
def rs_to_ab(r, s):
    r = np.asarray(r, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    a = np.where(np.abs(s - 1.0) > 1e-14, 2.0 * (1.0 + r) / np.where(np.abs(s - 1.0) > 1e-14, 1.0 - s, 1.0) - 1.0, -1.0)
    return a, s.copy()


def simplex_2d(a, b, i: int, j: int) -> np.ndarray:
    return math.sqrt(2.0) * jacobiP(a, 0, 0, i) * jacobiP(b, 2 * i + 1, 0, j) * (1 - b) ** i


def grad_simplex_2d(a, b, i: int, j: int):
    fa = jacobiP(a, 0, 0, i)
    dfa = grad_jacobiP(a, 0, 0, i)
    gb = jacobiP(b, 2 * i + 1, 0, j)
    dgb = grad_jacobiP(b, 2 * i + 1, 0, j)
    dmodedr = dfa * gb
    if i > 0:
        dmodedr = dmodedr * ((0.5 * (1 - b)) ** (i - 1))
    dmodeds = dfa * (gb * (0.5 * (1 + a)))
    if i > 0:
        dmodeds = dmodeds * ((0.5 * (1 - b)) ** (i - 1))
    tmp = dgb * ((0.5 * (1 - b)) ** i)
    if i > 0:
        tmp = tmp - 0.5 * i * gb * ((0.5 * (1 - b)) ** (i - 1))
    dmodeds = dmodeds + fa * tmp
    c = 2.0 ** (i + 0.5)
    return c * dmodedr, c * dmodeds


def simplex_basis_2d(N: int, r, s, grad: bool = False):
    a, b = rs_to_ab(r, s)
    cols, cr, cs = [], [], []
    for i in range(N + 1):
        for j in range(N - i + 1):
            cols.append(simplex_2d(a, b, i, j))
            if grad:
                dr, ds = grad_simplex_2d(a, b, i, j)
                cr.append(dr)
                cs.append(ds)
    V = np.stack(cols, axis=1)
    if grad:
        return V, np.stack(cr, axis=1), np.stack(cs, axis=1)
    return V


# -------------------------------------------------------------------------- Tet
def rst_to_abc(r, s, t):
    r = np.asarray(r, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    den1 = s + t
    ok1 = np.abs(den1) > 1e-14
    a = np.where(ok1, 2.0 * (1.0 + r) / np.where(ok1, -den1, 1.0) - 1.0, -1.0)
    ok2 = np.abs(t - 1.0) > 1e-14
    b = np.where(ok2, 2.0 * (1.0 + s) / np.where(ok2, 1.0 - t, 1.0) - 1.0, -1.0)
    return a, b, t.copy()


def simplex_3d(a, b, c, i: int, j: int, k: int) -> np.ndarray:
    return (2.0 * math.sqrt(2.0) * jacobiP(a, 0, 0, i) * jacobiP(b, 2 * i + 1, 0, j)
            * (1 - b) ** i * jacobiP(c, 2 * (i + j) + 2, 0, k) * (1 - c) ** (i + j))


def grad_simplex_3d(a, b, c, i: int, j: int, k: int):
    fa = jacobiP(a, 0, 0, i)
    dfa = grad_jacobiP(a, 0, 0, i)
    gb = jacobiP(b, 2 * i + 1, 0, j)
    dgb = grad_jacobiP(b, 2 * i + 1, 0, j)
    hc = jacobiP(c, 2 * (i + j) + 2, 0, k)
    dhc = grad_jacobiP(c, 2 * (i + j) + 2, 0, k)

    Vr = dfa * (gb * hc)
    if i > 0:
        Vr = Vr * ((0.5 * (1 - b)) ** (i - 1))
    if i + j > 0:
        Vr = Vr * ((0.5 * (1 - c)) ** (i + j - 1))

    Vs = 0.5 * (1 + a) * Vr
    tmp = dgb * ((0.5 * (1 - b)) ** i)
    if i > 0:
        tmp = tmp + (-0.5 * i) * (gb * (0.5 * (1 - b)) ** (i - 1))
    if i + j > 0:
        tmp = tmp * ((0.5 * (1 - c)) ** (i + j - 1))
    tmp = fa * (tmp * hc)
    Vs = Vs + tmp

    Vt = 0.5 * (1 + a) * Vr + 0.5 * (1 + b) * tmp
    tmp = dhc * ((0.5 * (1 - c)) ** (i + j))
    if i + j > 0:
        tmp = tmp - 0.5 * (i + j) * (hc * ((0.5 * (1 - c)) ** (i + j - 1)))
    tmp = fa * (gb * tmp)
    tmp = tmp * ((0.5 * (1 - b)) ** i)
    Vt = Vt + tmp

    cst = 2.0 ** (2 * i + j + 1.5)
    return cst * Vr, cst * Vs, cst * Vt


def simplex_basis_3d(N: int, r, s, t, grad: bool = False):
    a, b, c = rst_to_abc(r, s, t)
    cols, cr, cs, ct = [], [], [], []
    for i in range(N + 1):
        for j in range(N - i + 1):
            for k in range(N - i - j + 1):
                cols.append(simplex_3d(a, b, c, i, j, k))
                if grad:
                    dr, ds, dt = grad_simplex_3d(a, b, c, i, j, k)
                    cr.append(dr)
                    cs.append(ds)
                    ct.append(dt)
    V = np.stack(cols, axis=1)
    if grad:
        return V, np.stack(cr, axis=1), np.stack(cs, axis=1), np.stack(ct, axis=1)
    return V


# ---------------------------------------------------------------- Quad / Hex
def tensor_basis_2d(N: int, r, s, grad: bool = False):
    """Tensor-product orthonormal Legendre basis on the square (span Q_N)."""
    Vr, Vs = vandermonde_1d(N, r), vandermonde_1d(N, s)
    V = np.einsum("ni,nj->nij", Vr, Vs).reshape(len(Vr), -1)
    if not grad:
        return V
    dVr, dVs = grad_vandermonde_1d(N, r), grad_vandermonde_1d(N, s)
    return (V, np.einsum("ni,nj->nij", dVr, Vs).reshape(len(Vr), -1),
            np.einsum("ni,nj->nij", Vr, dVs).reshape(len(Vr), -1))


def tensor_basis_3d(N: int, r, s, t, grad: bool = False):
    Vr, Vs, Vt = vandermonde_1d(N, r), vandermonde_1d(N, s), vandermonde_1d(N, t)
    V = np.einsum("ni,nj,nk->nijk", Vr, Vs, Vt).reshape(len(Vr), -1)
    if not grad:
        return V
    dVr, dVs, dVt = (grad_vandermonde_1d(N, r), grad_vandermonde_1d(N, s),
                     grad_vandermonde_1d(N, t))
    return (V,
            np.einsum("ni,nj,nk->nijk", dVr, Vs, Vt).reshape(len(Vr), -1),
            np.einsum("ni,nj,nk->nijk", Vr, dVs, Vt).reshape(len(Vr), -1),
            np.einsum("ni,nj,nk->nijk", Vr, Vs, dVt).reshape(len(Vr), -1))
