"""Analysis functionals evaluated on the device.

Mirror of /root/reference/src/Analysis/conservation.jl (``PrimaryConservationAnalysis``,
``EnergyConservationAnalysis``, ``EntropyConservationAnalysis``; ``evaluate_conservation`` :113-143,
``evaluate_conservation_residual`` :145-190) and of ``ErrorAnalysis`` / ``analyze``
(/root/reference/src/Analysis/error.jl:1-91, default error quadrature).  The reference walks the
elements on the host with dense per-element mass matrices; here ``sse_functional`` reduces over
the device-resident arrays of the solver's handle, so a million-element state never leaves HBM.
File output (JLD2) of the reference's ``analyze`` drivers is out of scope.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import grid_functions as gfn


class _DeviceAnalysis:
    def __init__(self, solver):
        self.solver = solver
        self.N_p, self.N_c, self.N_e = solver.size()

    def _load(self, u: Optional[np.ndarray], need_dudt: bool = False):
        """``u=None`` means "use the state resident on the device" (e.g. the time integrator's).
        The residual, when needed, is evaluated on the device from that state."""
        h = self.solver.handle
        if u is not None:
            h.set_state(np.ascontiguousarray(u, dtype=np.float64))
        if need_dudt:
            h.nodal_values()
            h.time_derivative()
        return h


class PrimaryConservationAnalysis(_DeviceAnalysis):
    """∫ u dx per conserved variable (conservation.jl:4-11, 113-117, 145-152)."""

    def evaluate_conservation(self, u: Optional[np.ndarray] = None) -> np.ndarray:
        return self._load(u).functional("conservation", "state")

    def evaluate_conservation_residual(self, u: Optional[np.ndarray] = None) -> np.ndarray:
        return self._load(u, need_dudt=True).functional("conservation", "dudt")


class EnergyConservationAnalysis(_DeviceAnalysis):
    """∫ ½u² dx = ½ uᵀ M u with the mass solver's mass matrix (conservation.jl:14-21, 130-143,
    154-167)."""

    def evaluate_conservation(self, u: Optional[np.ndarray] = None) -> np.ndarray:
        return self._load(u).functional("energy")

    def evaluate_conservation_residual(self, u: Optional[np.ndarray] = None) -> np.ndarray:
        return self._load(u, need_dudt=True).functional("energy_residual")


class EntropyConservationAnalysis(_DeviceAnalysis):
    """∫ S(u) dx and (P w)ᵀ M dudt (conservation.jl:23-34, 119-128, 169-190)."""

    def evaluate_conservation(self, u: Optional[np.ndarray] = None) -> np.ndarray:
        return self._load(u).functional("entropy")

    def evaluate_conservation_residual(self, u: Optional[np.ndarray] = None) -> np.ndarray:
        return self._load(u, need_dudt=True).functional("entropy_residual")


class ErrorAnalysis(_DeviceAnalysis):
    """L2 error against an exact solution sampled at the volume quadrature nodes
    (error.jl:14-56 with ``error_quadrature_rule = nothing``, ``analyze`` :58-91)."""

    def __init__(self, solver):
        super().__init__(solver)
        sd = solver.spatial_discretization
        self.xyzq = tuple(x.T for x in sd.mesh.xyzq)            # (N_e, N_q) each
        ra = sd.reference_approximation
        self.total_volume = float(np.sum(ra.W[None, :] * sd.geometric_factors.J_q))

    def analyze(self, sol: Optional[np.ndarray], exact_solution, t: float = 0.0,
                normalize: bool = False) -> np.ndarray:
        u_exact = gfn.evaluate(exact_solution, self.xyzq, t)     # (N_c, N_e, N_q)
        exact_q = np.ascontiguousarray(u_exact.transpose(1, 0, 2))
        err = self._load(sol).functional("l2_error", exact_q=exact_q)
        return err / np.sqrt(self.total_volume) if normalize else err
