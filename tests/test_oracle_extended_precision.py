"""The NumPy oracle evaluated in 80-bit extended precision (np.longdouble) as the accuracy reference
of the FP64 residuals -- the justification of the one relaxed parity bound (parity
caveat v).

For every state but one the FP64 oracle is within 3e-13 of the 80-bit evaluation of the same
algorithm on the same FP64 inputs.  The smooth low-Mach Taylor-Green state (p ~ 71, |V| <= 1) is the
exception: round-off is amplified to ~5e-12 inside the algorithm (differences of nearly equal
fluxes), although the mathematical conditioning -- a 1-ulp input perturbation evaluated in 80-bit
arithmetic -- is only ~7e-14.  No two FP64 implementations can agree to 1e-12 there; the parity
tests require the CUDA kernels to be as close to the 80-bit residual as the FP64 oracle is."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))

import sse_oracle as oc  # noqa: E402
import cases  # noqa: E402
from bridge import oracle_problem  # noqa: E402
from sse_b200 import device as dev  # noqa: E402

pytestmark = pytest.mark.skipif(np.finfo(np.longdouble).eps > 2e-19,
                                reason="np.longdouble is not 80-bit extended precision here")


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


BUILD = {
    "tet_p4_lf": lambda: cases.euler_tet_case(p=4, M=2, lazy=True),
    "tet_p4_warp_ec": lambda: cases.euler_tet_case(p=4, M=2, lazy=True, warp=True, interface="ec"),
    "tri_p4_lf": lambda: cases.euler_tri_case(p=4, M=3, lazy=True),
}


@pytest.mark.parametrize("name,state,lo,hi", [
    ("tet_p4_lf", "smooth", 1e-12, 2e-11),       # the ill-conditioned-in-FP64 state: ~5e-12
    ("tet_p4_lf", "rough", 0.0, 1e-12),
    ("tet_p4_warp_ec", "smooth", 0.0, 1e-12),
    ("tet_p4_warp_ec", "rough", 0.0, 1e-12),
    ("tri_p4_lf", "rough", 0.0, 1e-12),
])
def test_fp64_oracle_against_80_bit_oracle(name, state, lo, hi):
    solver, u0 = BUILD[name]()
    u = u0 if state == "smooth" else cases.rough_state(solver, u0, seed=1)
    prob = oracle_problem(solver)
    ref = oc.semi_discrete_residual(prob, u)
    ref_x = oc.semi_discrete_residual(prob, u.astype(np.longdouble))
    assert ref_x.dtype == np.longdouble
    err = _rel(ref.astype(np.longdouble), ref_x)
    assert lo <= err < hi, err
    # mathematical conditioning: a +-1 ulp perturbation of the input, evaluated in 80-bit arithmetic
    rng = np.random.default_rng(123)
    up = u * (1.0 + rng.integers(-1, 2, size=u.shape) * 1.1e-16)
    cond = _rel(oc.semi_discrete_residual(prob, up.astype(np.longdouble)), ref_x)
    assert cond < 2e-13, cond


def test_kernels_are_as_close_to_the_80_bit_residual_as_the_fp64_oracle():
    """The product kernels (host emulation build: same source, same FP64 operation order up to FMA
    contraction) on the smooth Taylor-Green state."""
    import build_emu
    lib = dev.load_library(build_emu.build(), allow_emulation=True)
    solver, u0 = BUILD["tet_p4_lf"]()
    prob = oracle_problem(solver)
    ref = oc.semi_discrete_residual(prob, u0)
    ref_x = oc.semi_discrete_residual(prob, u0.astype(np.longdouble))
    floor = _rel(ref.astype(np.longdouble), ref_x)
    saved, dev._LIB = dev._LIB, lib       # DeviceResidual picks the library up from here
    try:
        d = dev.DeviceResidual(solver)
        try:
            dudt = np.full_like(u0, np.nan)
            d.residual_host(u0, dudt)
        finally:
            d.close()
    finally:
        dev._LIB = saved
    err_x = _rel(dudt.astype(np.longdouble), ref_x)
    assert err_x < max(1e-12, 1.5 * floor), (err_x, floor)
