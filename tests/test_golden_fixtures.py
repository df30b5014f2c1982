"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU: the oracle still reproduces them (guards the checker against drift).
GPU: the CUDA residual behind the C-ABI reproduces them to 1e-12."""
import os
import sys

import numpy as np
import pytest

import sse_oracle as oc
from bridge import oracle_problem

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402


def _load(name):
    d = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return d["u"], d["dudt"]


@pytest.mark.parametrize("name", sorted(mg.FIXTURES))
def test_oracle_reproduces_fixture(name):
    u, dudt = _load(name)
    solver, u2 = mg.build(name)
    assert np.array_equal(u, u2)
    r = oc.semi_discrete_residual(oracle_problem(solver), u)
    assert np.max(np.abs(r - dudt)) <= 1e-13 * np.max(np.abs(dudt))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.FIXTURES))
def test_cuda_reproduces_fixture(name):
    from sse_b200.solvers import semi_discrete_residual
    u, dudt = _load(name)
    solver, _ = mg.build(name)
    try:
        out = np.empty_like(u)
        semi_discrete_residual(out, u, solver, 0.0)
        assert np.max(np.abs(out - dudt)) / np.max(np.abs(dudt)) < 1e-12
    finally:
        solver.close()
