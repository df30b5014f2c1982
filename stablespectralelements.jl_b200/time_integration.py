"""Device-resident explicit time integration.

The reference hands the residual to OrdinaryDiffEq (``solve(ode, CarpenterKennedy2N54(); dt,
adaptive=false, callback=save_callback(...))``, /root/reference/test/test_driver.jl:77-83,
/root/reference/src/File/save.jl:79-90).  OrdinaryDiffEq calls ``f(du, u, p, t)`` with host arrays
every stage; here the state stays on the device and each 2N Runge-Kutta stage
``k <- a k + dt R(u); u <- u + b k`` is applied in the epilogue of the loop-B kernel
(``sse_rk_stage``), so a time step is 10 kernel launches and no host traffic.  General explicit
tableaus (``RK4``, ``SSPRK33``, ``DP5``, ``DP8`` -- the reference's 3-D Euler test integrates
with DP8) run through ``sse_erk_step`` with device-resident stage buffers.  Snapshots are copied
back only when the callback asks for them.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

# Carpenter & Kennedy (1994) 5-stage 4th-order 2N coefficients (same values as the C library)
CK54_A = (0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238,
          -3550918686646 / 2091501179385, -1275806237668 / 842570457699)
CK54_B = (1432997174477 / 9575080441755, 5161836677717 / 13612068292357,
          1720146321549 / 2090206949498, 3134564353537 / 4481467310338,
          2277821191437 / 14882151754819)


class CarpenterKennedy2N54:
    """Algorithm tag, named like OrdinaryDiffEq's."""
    a, b = CK54_A, CK54_B


class ExplicitRK:
    """A general explicit Runge-Kutta scheme (Butcher tableau ``A``, weights ``b``, nodes ``c``)
    stepped on the device by ``sse_erk_step``: OrdinaryDiffEq's non-low-storage algorithms."""

    def __init__(self, A, b, c=None):
        self.A = np.asarray(A, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64)
        self.c = self.A.sum(axis=1) if c is None else np.asarray(c, dtype=np.float64)
        if self.A.shape != (len(self.b),) * 2 or np.any(np.triu(self.A) != 0.0):
            raise ValueError("explicit tableau expected: A (s, s) strictly lower triangular")


class RK4(ExplicitRK):
    """The classical fourth-order scheme (OrdinaryDiffEq ``RK4`` with ``adaptive=false``)."""

    def __init__(self):
        super().__init__([[0, 0, 0, 0], [0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 1, 0]],
                         [1 / 6, 1 / 3, 1 / 3, 1 / 6])


class SSPRK33(ExplicitRK):
    """Shu-Osher three-stage third-order SSP scheme."""

    def __init__(self):
        super().__init__([[0, 0, 0], [1, 0, 0], [0.25, 0.25, 0]], [1 / 6, 1 / 6, 2 / 3])


class DP5(ExplicitRK):
    """Dormand-Prince 5(4), fixed step (the 5th-order weights; the FSAL stage has weight 0)."""

    def __init__(self):
        A = np.zeros((6, 6))
        A[1, :1] = [1 / 5]
        A[2, :2] = [3 / 40, 9 / 40]
        A[3, :3] = [44 / 45, -56 / 15, 32 / 9]
        A[4, :4] = [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729]
        A[5, :5] = [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]
        super().__init__(A, [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])


class DP8(ExplicitRK):
    """Dormand-Prince 8(5,3) with a fixed step -- ``DP8()`` of the reference's 3-D Euler test
    (/root/reference/test/euler_3d.jl:44-51).  OrdinaryDiffEq's DP8 is Hairer's DOP853; the
    published 12-stage tableau is taken from SciPy's copy of it."""

    def __init__(self):
        from scipy.integrate._ivp import dop853_coefficients as dc
        super().__init__(dc.A[:12, :12], dc.B, dc.C[:12])


def solve(ode, alg=None, dt: float = None, save_every: Optional[int] = None,
          callback: Optional[Callable[[np.ndarray, float, int], None]] = None) -> np.ndarray:
    """Integrate ``ode`` (from ``semidiscretize``) over ``ode.tspan`` with fixed ``dt`` (the last
    step is clipped to land on the end time, as OrdinaryDiffEq does with ``adaptive=false``).
    ``callback(u, t, step)`` receives a host copy of the state every ``save_every`` steps (and at
    the end), mirroring ``save_callback(results_path, tspan, interval)``.  Returns the final state."""
    alg = alg or CarpenterKennedy2N54()
    if dt is None or dt <= 0:
        raise ValueError("a positive fixed time step dt is required")
    h = ode.p.handle
    h.set_state(np.ascontiguousarray(ode.u0, dtype=np.float64))
    t, t_end = ode.tspan
    step = 0
    while t < t_end - 1e-12 * max(1.0, abs(t_end)):
        hstep = min(dt, t_end - t)
        if isinstance(alg, ExplicitRK):
            h.erk_step(alg.A, alg.b, hstep)
        else:                                     # 2N low-storage: fused into loop B's epilogue
            for a, b in zip(alg.a, alg.b):
                h.rk_stage(a, b, hstep)
        t += hstep
        step += 1
        if callback is not None and save_every and step % save_every == 0:
            callback(h.get_state(), t, step)
    u = h.get_state()
    if callback is not None:
        callback(u, t, step)
    return u
