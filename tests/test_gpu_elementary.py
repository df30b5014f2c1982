"""Accuracy of the device log / exp that the entropy-variable maps use (physics.cuh: flog, fexp;
the reference evaluates euler_navierstokes.jl:100-131 with Base.log / Base.exp), through the C ABI
(sse_probe_elementary).  Tolerance: 2 ulp relative for exp; for log, 2 ulp of max(1, |log x|) --
what the 1e-12 residual parity needs is absolute accuracy of log p - gamma log rho."""
import numpy as np
import pytest

from sse_b200 import device as dev


def elementary_inputs():
    rng = np.random.default_rng(11)
    pos = np.concatenate([np.exp(rng.uniform(-700, 700, 40000)), rng.uniform(0.5, 2.0, 40000),
                          1.0 + rng.uniform(-1e-6, 1e-6, 4000), rng.uniform(1e-3, 1e3, 40000),
                          [1.0, 2.0, 0.5, np.sqrt(0.5), np.sqrt(2.0), np.nextafter(1.0, 2), np.nextafter(1.0, 0),
                           2.2250738585072014e-308, 1.7976931348623157e308]])
    special_log = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310])
    ex = np.concatenate([rng.uniform(-699, 699, 40000), rng.uniform(-1, 1, 40000), rng.uniform(-40, 40, 40000),
                         [0.0, -0.0, 1e-300, np.log(2) / 2, -np.log(2) / 2, 699.999, -699.999]])
    special_exp = np.array([709.0, 710.0, -708.0, -800.0, np.inf, -np.inf, np.nan, 700.0, -700.0])
    return pos, special_log, ex, special_exp


def check_elementary(probe):
    pos, special_log, ex, special_exp = elementary_inputs()
    ulp = np.finfo(np.float64).eps
    ref = np.log(pos.astype(np.longdouble))
    got = probe("log", pos)
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    assert float(err.max()) < 2 * ulp, float(err.max())
    near1 = np.abs(pos - 1.0) < 0.3          # relative accuracy where the result is small
    rel = np.abs(got[near1] - ref[near1]) / np.maximum(np.abs(ref[near1]), np.finfo(np.float64).tiny)
    assert float(rel.max()) < 2 * ulp, float(rel.max())
    with np.errstate(all="ignore"):
        want = np.log(special_log)
    got = probe("log", special_log)
    np.testing.assert_allclose(got, want, rtol=2 * ulp, equal_nan=True)
    ref = np.exp(ex.astype(np.longdouble))
    got = probe("exp", ex)
    rel = np.abs(got - ref) / ref
    assert float(rel.max()) < 2 * ulp, float(rel.max())
    with np.errstate(all="ignore"):
        want = np.exp(special_exp)
    got = probe("exp", special_exp)
    np.testing.assert_allclose(got, want, rtol=4 * ulp, equal_nan=True)


@pytest.mark.gpu
def test_device_log_exp_accuracy():
    check_elementary(lambda which, x: dev.probe_elementary(which, x))
