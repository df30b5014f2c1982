"""Interpolation (mapping) nodes on the reference elements.

StartUpDG's ``nodes(elem, N)`` is un-vendored; the reference only uses these nodes to
*represent the element mapping* (``mesh.xyz``; /root/reference/src/SpatialDiscretizations/
mesh.jl:231-284), so any unisolvent degree-N node set spans the same space.  We use Legendre-
Gauss-Lobatto nodes on the line (and their tensor products on Quad/Hex) and the
Hesthaven-Warburton warp-and-blend nodes on Tri/Tet with alpha = 0, whose edge nodes are
the 1-D LGL nodes; on the triangle this reproduces the reference's golden numbers exactly.
"""
from __future__ import annotations

import math

import numpy as np

from .polynomials import gauss_lobatto, jacobiP, vandermonde_1d


def nodes_line(N: int) -> np.ndarray:
    return gauss_lobatto(N + 1)[0]


def nodes_quad(N: int):
    r1 = nodes_line(N)
    # first coordinate slowest (matches the quadrature meshgrid convention)
    r = np.repeat(r1, N + 1)
    s = np.tile(r1, N + 1)
    return r, s


def nodes_hex(N: int):
    r1 = nodes_line(N)
    n = N + 1
    r = np.repeat(r1, n * n)
    s = np.tile(np.repeat(r1, n), n)
    t = np.tile(r1, n * n)
    return r, s, t


# ----------------------------------------------------------------------------- Tri
_ALPOPT_2D = [0.0000, 0.0000, 1.4152, 0.1001, 0.2751, 0.9800, 1.0999, 1.2832, 1.3648,
              1.4773, 1.4959, 1.5743, 1.5770, 1.6223, 1.6258]


def _warpfactor(N: int, rout: np.ndarray) -> np.ndarray:
    LGLr = gauss_lobatto(N + 1)[0]
    req = np.linspace(-1.0, 1.0, N + 1)
    Veq = vandermonde_1d(N, req)
    Pmat = np.stack([jacobiP(rout, 0, 0, i) for i in range(N + 1)], axis=0)
    Lmat = np.linalg.solve(Veq.T, Pmat)
    warp = Lmat.T @ (LGLr - req)
    zerof = (np.abs(rout) < 1.0 - 1.0e-10).astype(np.float64)
    sf = 1.0 - (zerof * rout) ** 2
    return warp / sf + warp * (zerof - 1.0)


def nodes_tri(N: int):
    if N == 0:
        return np.array([-1.0 / 3.0]), np.array([-1.0 / 3.0])
    # alpha = 0 (no alpha-optimisation): this is what reproduces the reference's golden L2
    # error for the mapping_degree = 4 triangle case to 4e-16 (runtests.jl:38-60), i.e. what
    # NodesAndModes' ``nodes(Tri(), N)`` yields; the H-W optimised table (_ALPOPT_2D) does not.
    alpha = 0.0
    Np = (N + 1) * (N + 2) // 2
    L1 = np.zeros(Np)
    L3 = np.zeros(Np)
    sk = 0
    for n in range(1, N + 2):
        for m in range(1, N + 3 - n):
            L1[sk] = (n - 1) / N
            L3[sk] = (m - 1) / N
            sk += 1
    L2 = 1.0 - L1 - L3
    x = -L2 + L3
    y = (-L2 - L3 + 2 * L1) / math.sqrt(3.0)
    blend1, blend2, blend3 = 4 * L2 * L3, 4 * L1 * L3, 4 * L1 * L2
    warpf1 = _warpfactor(N, L3 - L2)
    warpf2 = _warpfactor(N, L1 - L3)
    warpf3 = _warpfactor(N, L2 - L1)
    warp1 = blend1 * warpf1 * (1 + (alpha * L1) ** 2)
    warp2 = blend2 * warpf2 * (1 + (alpha * L2) ** 2)
    warp3 = blend3 * warpf3 * (1 + (alpha * L3) ** 2)
    x = x + warp1 + math.cos(2 * math.pi / 3) * warp2 + math.cos(4 * math.pi / 3) * warp3
    y = y + math.sin(2 * math.pi / 3) * warp2 + math.sin(4 * math.pi / 3) * warp3
    # equilateral -> reference right triangle
    L1 = (math.sqrt(3.0) * y + 1.0) / 3.0
    L2 = (-3.0 * x - math.sqrt(3.0) * y + 2.0) / 6.0
    L3 = (3.0 * x - math.sqrt(3.0) * y + 2.0) / 6.0
    r = -L2 + L3 - L1
    s = -L2 - L3 + L1
    return r, s


# ----------------------------------------------------------------------------- Tet
_ALPOPT_3D = [0.0, 0.0, 0.0, 0.1002, 1.1332, 1.5608, 1.3413, 1.2577, 1.1603, 1.10153,
              0.6080, 0.4523, 0.8856, 0.8717, 0.9655]


def equi_nodes_tet(N: int):
    Np = (N + 1) * (N + 2) * (N + 3) // 6
    X, Y, Z = np.zeros(Np), np.zeros(Np), np.zeros(Np)
    sk = 0
    for n in range(1, N + 2):
        for m in range(1, N + 3 - n):
            for q in range(1, N + 4 - n - m):
                X[sk] = -1 + (q - 1) * 2.0 / N
                Y[sk] = -1 + (m - 1) * 2.0 / N
                Z[sk] = -1 + (n - 1) * 2.0 / N
                sk += 1
    return X, Y, Z


def _evalwarp(p: int, xnodes: np.ndarray, xout: np.ndarray) -> np.ndarray:
    warp = np.zeros_like(xout)
    xeq = np.array([-1.0 + 2.0 * (p - i) / p for i in range(p + 1)])
    for i in range(p + 1):
        d = (xnodes[i] - xeq[i]) * np.ones_like(xout)
        for j in range(1, p):
            if i != j:
                d = d * (xout - xeq[j]) / (xeq[i] - xeq[j])
        if i != 0:
            d = -d / (xeq[i] - xeq[0])
        if i != p:
            d = d / (xeq[i] - xeq[p])
        warp = warp + d
    return warp


def _evalshift(p: int, pval: float, L1, L2, L3):
    gaussX = -gauss_lobatto(p + 1)[0]
    blend1, blend2, blend3 = L2 * L3, L1 * L3, L1 * L2
    wf1 = 4 * _evalwarp(p, gaussX, L3 - L2)
    wf2 = 4 * _evalwarp(p, gaussX, L1 - L3)
    wf3 = 4 * _evalwarp(p, gaussX, L2 - L1)
    warp1 = blend1 * wf1 * (1 + (pval * L1) ** 2)
    warp2 = blend2 * wf2 * (1 + (pval * L2) ** 2)
    warp3 = blend3 * wf3 * (1 + (pval * L3) ** 2)
    dx = warp1 + math.cos(2 * math.pi / 3) * warp2 + math.cos(4 * math.pi / 3) * warp3
    dy = math.sin(2 * math.pi / 3) * warp2 + math.sin(4 * math.pi / 3) * warp3
    return dx, dy


def nodes_tet(N: int):
    if N == 0:
        return np.array([-0.5]), np.array([-0.5]), np.array([-0.5])
    alpha = 0.0  # consistent with the triangle (see nodes_tri)
    tol = 1e-10
    r, s, t = equi_nodes_tet(N)
    L1 = (1 + t) / 2
    L2 = (1 + s) / 2
    L3 = -(1 + r + s + t) / 2
    L4 = (1 + r) / 2
    v1 = np.array([-1.0, -1 / math.sqrt(3), -1 / math.sqrt(6)])
    v2 = np.array([1.0, -1 / math.sqrt(3), -1 / math.sqrt(6)])
    v3 = np.array([0.0, 2 / math.sqrt(3), -1 / math.sqrt(6)])
    v4 = np.array([0.0, 0.0, 3 / math.sqrt(6)])
    t1 = np.stack([v2 - v1, v2 - v1, v3 - v2, v3 - v1])
    t2 = np.stack([v3 - 0.5 * (v1 + v2), v4 - 0.5 * (v1 + v2), v4 - 0.5 * (v2 + v3),
                   v4 - 0.5 * (v1 + v3)])
    t1 = t1 / np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = t2 / np.linalg.norm(t2, axis=1, keepdims=True)
    XYZ = np.outer(L3, v1) + np.outer(L4, v2) + np.outer(L2, v3) + np.outer(L1, v4)
    shift = np.zeros_like(XYZ)
    for face in range(4):
        if face == 0:
            La, Lb, Lc, Ld = L1, L2, L3, L4
        elif face == 1:
            La, Lb, Lc, Ld = L2, L1, L3, L4
        elif face == 2:
            La, Lb, Lc, Ld = L3, L1, L4, L2
        else:
            La, Lb, Lc, Ld = L4, L1, L3, L2
        warp1, warp2 = _evalshift(N, alpha, Lb, Lc, Ld)
        blend = Lb * Lc * Ld
        denom = (Lb + 0.5 * La) * (Lc + 0.5 * La) * (Ld + 0.5 * La)
        ids = denom > tol
        blend = np.where(ids, (1 + (alpha * La) ** 2) * blend / np.where(ids, denom, 1.0), blend)
        shift = shift + np.outer(blend * warp1, t1[face]) + np.outer(blend * warp2, t2[face])
        fix = (La < tol) & (((Lb > tol).astype(int) + (Lc > tol).astype(int)
                             + (Ld > tol).astype(int)) < 3)
        shift[fix] = np.outer(warp1[fix], t1[face]) + np.outer(warp2[fix], t2[face])
    XYZ = XYZ + shift
    # equilateral tet -> reference tet
    A = np.stack([0.5 * (v2 - v1), 0.5 * (v3 - v1), 0.5 * (v4 - v1)], axis=1)
    rhs = XYZ.T - 0.5 * (v2 + v3 + v4 - v1)[:, None]
    rst = np.linalg.solve(A, rhs)
    return rst[0].copy(), rst[1].copy(), rst[2].copy()
