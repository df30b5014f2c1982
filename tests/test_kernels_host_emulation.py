"""The product's CUDA kernel SOURCES, executed on the CPU.

tests/emu/cuda_emu.h runs every CUDA thread of a block as a fiber (``__syncthreads`` = yield to a
round-robin scheduler) and maps the runtime API onto host memory; tests/emu/build_emu.py compiles
csrc/sse_b200.cu against it with g++.  The emulated library goes through the same C ABI, the
same ``sse_create`` table construction and the same kernels (specialised and generic) as the GPU
build, so index logic, shared-memory carve-ups, barrier placement and operator tables are checked
against the oracle in the CPU suite -- shared and "device" memory are NaN-poisoned, so a read of
an unwritten slot fails the comparison.  It is test infrastructure: never built by
``__graft_entry__.build()``, refused by ``device.load_library`` unless asked for, and no
statement about sm_100a code generation or speed (that is ``-m gpu``)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import cases
import golden_cases as gc
import sse_oracle as oc
from bridge import oracle_problem
from sse_b200 import device as dev

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    lib = dev.load_library(build_emu.build(), allow_emulation=True)
    assert lib.sse_version() < 0
    lib.emu_launch_log.restype = C.c_char_p
    saved = dev._LIB
    dev._LIB = lib                      # DeviceResidual picks the library up from here
    try:
        yield lib
    finally:
        dev._LIB = saved


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


# name -> (builder, kernels that must have run)
CASES = {
    # north-star path: specialised loop A / loop B kernels on collapsed tetrahedra
    "euler3d_tet_p4_warp_lf": (lambda: cases.euler_tet_case(p=4, M=2, lazy=True, warp=True,
                                                            ic="periodic"),
                               ["k_nodal_tensorILi3ELi5E", "k_fluxdiff_tensorILi3ELi5E"]),
    "euler3d_tet_p3_warp_ec": (lambda: cases.euler_tet_case(p=3, M=2, lazy=True, warp=True,
                                                            interface="ec", ic="periodic"),
                               ["k_nodal_tensorILi3ELi4E", "k_fluxdiff_tensorILi3ELi4E"]),
    "euler3d_tet_p2_nodal": (lambda: cases.euler_tet_case(p=2, M=2, lazy=True, approx="nodal"),
                             ["k_nodal_", "k_fluxdiff"]),
    "euler2d_tri_p4_lf": (lambda: cases.euler_tri_case(p=4, M=3, lazy=True),
                          ["k_nodal_tensorILi2ELi5E", "k_fluxdiff_tensorILi2ELi5E"]),
    # standard form: specialised scalar kernels (elements as components) and generic ones
    "adv3d_tet_p4": (lambda: cases.advection_tet_case(p=4, M=2, lazy=True),
                     ["k_nodal_batchedILi3ELi5E", "k_standard_tensorILi3ELi5E"]),
    "adv2d_tri_p4": (lambda: cases.advection_tri_case(p=4, M=3, lazy=True), ["k_standard"]),
    "burgers2d_tri_p3_ec": (lambda: cases.burgers_tri_case(p=3, M=3, lazy=True), ["k_fluxdiff"]),
    "euler3d_hex_nodal_p3_ec": (lambda: cases.euler_hex_case(p=3, M=2, lazy=True),
                                ["k_nodal_values", "k_fluxdiffILi3E"]),
    # physical operators, BR1 (two k_physical launches: auxiliary_variable!, time_derivative!)
    "advdiff1d_p4": (lambda: cases.advection_diffusion_case(d=1, p=4, M=4, lazy=True),
                     ["k_physicalILi1E"]),
    "advdiff2d_p3": (lambda: cases.advection_diffusion_case(d=2, p=3, M=3, lazy=True),
                     ["k_physicalILi2E"]),
    "golden_euler1d_gauss": (lambda: gc.euler_1d_gauss(lazy=True)[:2], ["k_fluxdiffILi1E"]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_emulated_kernels_match_oracle(emu_lib, name):
    build, expected = CASES[name]
    solver, u0 = build()
    u = cases.rough_state(solver, u0, seed=1)
    d = dev.DeviceResidual(solver)
    try:
        emu_lib.emu_launch_log()
        dudt = np.full_like(u, np.nan)
        d.residual_host(u, dudt)
        launched = emu_lib.emu_launch_log().decode()
        for k in expected:
            assert k in launched, (k, launched)
        ref = oc.semi_discrete_residual(oracle_problem(solver), u)
        assert np.all(np.isfinite(dudt))
        assert _rel(dudt, ref) < 1e-12
    finally:
        d.close()


def test_product_loader_refuses_emulation_build(emu_lib):
    import build_emu
    saved, dev._LIB = dev._LIB, None
    try:
        with pytest.raises(RuntimeError, match="host-emulation test build"):
            dev.load_library(build_emu.build())
    finally:
        dev._LIB = saved
