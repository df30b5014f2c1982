"""Host emulation of the sum-factorised 3-D warped-product stages (csrc/vmap3.cuh): the same
source the kernels compile is run as loops over the thread index (tests/emu/vmap3_emu.cpp) and
compared with the dense Vandermonde matrix of the reference's WarpedTensorProductMap3D
(warped_product_3d.jl:47-136).  No GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from sse_b200.reference_approximation import ModalTensor, Tet, make_reference_approximation

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "libvmap3_emu.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                    os.path.join(HERE, "emu", "vmap3_emu.cpp")], check=True)
    return ctypes.CDLL(str(out))


def _ptr(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


@pytest.mark.parametrize("p,nc,e", [(4, 5, 1), (4, 1, 1), (4, 4, 1), (3, 5, 2), (3, 1, 2),
                                    (2, 5, 4), (2, 1, 4)])
@pytest.mark.parametrize("nthr", [128, 64])
def test_vmap3_stages_match_dense_V(emu, p, nc, e, nthr):
    ra = make_reference_approximation(ModalTensor(p), Tet())
    V = ra.V
    n = p + 1
    Vd = V.to_dense()
    Nq, Np = Vd.shape
    sig = np.ascontiguousarray(V.sigma_i, dtype=np.int32)
    A = np.ascontiguousarray(V.A); B = np.ascontiguousarray(V.B); C = np.ascontiguousarray(V.C)
    rng = np.random.default_rng(p * 10 + nc)
    ZS = n * n * (n + 1) // 2
    Z = np.zeros(2 * e * nc * ZS)
    # V
    x = rng.standard_normal((e * nc, Np))
    src = x.copy().ravel(); dst = np.zeros(e * nc * Nq)
    rc = emu.vmap3_emu(n, nc, e, 0, _ptr(A), _ptr(B), _ptr(C), _ptr(sig, ctypes.c_int),
                       _ptr(src), _ptr(dst), _ptr(Z), nthr)
    assert rc == 0
    ref = x @ Vd.T
    assert np.max(np.abs(dst.reshape(e * nc, Nq) - ref)) < 1e-13 * np.max(np.abs(ref))
    # V^T (destroys its source)
    y = rng.standard_normal((e * nc, Nq))
    src = y.copy().ravel(); dst = np.zeros(e * nc * Np)
    rc = emu.vmap3_emu(n, nc, e, 1, _ptr(A), _ptr(B), _ptr(C), _ptr(sig, ctypes.c_int),
                       _ptr(src), _ptr(dst), _ptr(Z), nthr)
    assert rc == 0
    ref = y @ Vd
    assert np.max(np.abs(dst.reshape(e * nc, Np) - ref)) < 1e-13 * np.max(np.abs(ref))
    # V V^T in place (fused middle stage)
    y2 = rng.standard_normal((e * nc, Nq))
    src = y2.copy().ravel(); dst = np.zeros(1)
    rc = emu.vmap3_emu(n, nc, e, 2, _ptr(A), _ptr(B), _ptr(C), _ptr(sig, ctypes.c_int),
                       _ptr(src), _ptr(dst), _ptr(Z), nthr)
    assert rc == 0
    ref = y2 @ Vd @ Vd.T
    assert np.max(np.abs(src.reshape(e * nc, Nq) - ref)) < 1e-13 * np.max(np.abs(ref))


@pytest.mark.parametrize("p,ncol,g", [(4, 5, 4), (4, 5, 5), (4, 5, 1), (3, 5, 6), (2, 5, 8),
                                      (4, 4, 3), (3, 2, 5)])
@pytest.mark.parametrize("nthr", [128, 96])
def test_batched_engine_matches_dense_V(emu, p, ncol, g, nthr):
    """csrc/vmap3b.cuh (all columns of a tensor line / b1 group / pair per work item, several
    elements per CTA): V, V^T and the fused V V^T against the dense Vandermonde matrix."""
    ra = make_reference_approximation(ModalTensor(p), Tet())
    V = ra.V
    n = p + 1
    Vd = V.to_dense()
    Nq, Np = Vd.shape
    sig = np.ascontiguousarray(V.sigma_i, dtype=np.int32)
    A = np.ascontiguousarray(V.A); B = np.ascontiguousarray(V.B); C = np.ascontiguousarray(V.C)
    rng = np.random.default_rng(p * 100 + ncol * 10 + g)
    ZS = n * n * (n + 1) // 2
    cols = g * ncol

    def call(mode, X, M):
        Z = np.full(g * (ncol * ZS + 16), np.nan)
        rc = emu.vmap3b_emu(n, ncol, g, mode, _ptr(A), _ptr(B), _ptr(C), _ptr(sig, ctypes.c_int),
                            _ptr(X), _ptr(M), _ptr(Z), nthr)
        assert rc == 0

    m = rng.standard_normal((cols, Np))
    X = np.full(cols * Nq, np.nan); M = m.copy().ravel()
    call(0, X, M)
    ref = m @ Vd.T
    assert np.max(np.abs(X.reshape(cols, Nq) - ref)) < 1e-13 * np.max(np.abs(ref))
    x = rng.standard_normal((cols, Nq))
    X = x.copy().ravel(); M = np.full(cols * Np, np.nan)
    call(1, X, M)
    ref = x @ Vd
    assert np.max(np.abs(M.reshape(cols, Np) - ref)) < 1e-13 * np.max(np.abs(ref))
    X = x.copy().ravel(); M = np.zeros(1)
    call(2, X, M)
    ref = x @ Vd @ Vd.T
    assert np.max(np.abs(X.reshape(cols, Nq) - ref)) < 1e-13 * np.max(np.abs(ref))
