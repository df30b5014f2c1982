"""Device-resident explicit time integration.

The reference hands the residual to OrdinaryDiffEq (``solve(ode, CarpenterKennedy2N54(); dt,
adaptive=false, callback=save_callback(...))``, /root/reference/test/test_driver.jl:77-83,
/root/reference/src/File/save.jl:79-90).  OrdinaryDiffEq calls ``f(du, u, p, t)`` with host arrays
every stage; here the state stays on the device and each 2N Runge-Kutta stage
``k <- a k + dt R(u); u <- u + b k`` is applied in the epilogue of the loop-B kernel
(``sse_rk_stage``), so a time step is 10 kernel launches and no host traffic.  Snapshots are
copied back only when the callback asks for them.
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
