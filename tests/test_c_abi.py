"""The C-ABI library loads, exports every symbol include/sse_b200.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
from sse_b200 import device as dev

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    return dev.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sse_b200.h")).read()
    # function declarations only: "<type> sse_xxx(" at the start of a line
    declared = set(re.findall(r"^[a-z_0-9\* ]+?\b(sse_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sse_b200.h but not exported"
    assert declared == set(dev.EXPORTS)
    assert lib.sse_version() >= 100


def test_struct_layout_matches_header():
    # sizes follow the C struct with natural alignment (checked against the compiled library by
    # every GPU test; here: the ctypes mirror is self-consistent)
    assert C.sizeof(dev.SseConfig) % 8 == 0
    assert dev.SseConfig.N_e.offset == 24 and dev.SseConfig.a.offset == 48


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    solver, u0 = cases.advection_tri_case(p=2, M=2, lazy=True)
    with pytest.raises(RuntimeError, match="no CUDA device|sse_create failed"):
        solver.handle
    with pytest.raises(RuntimeError):
        from sse_b200.solvers import semi_discrete_residual
        semi_discrete_residual(np.empty_like(u0), u0, solver, 0.0)
