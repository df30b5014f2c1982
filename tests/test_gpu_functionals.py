"""GPU parity of the device-side analysis functionals against the oracle's
restatement of Analysis/conservation.jl:113-190 and Analysis/error.jl:58-91."""
import numpy as np
import pytest

import cases
import sse_oracle as oc
from bridge import oracle_problem
from sse_b200 import analysis
from sse_b200.grid_functions import InitialDataSine

pytestmark = pytest.mark.gpu

CASES = {
    "adv2d_tri_p4_central": lambda: cases.advection_tri_case(p=4, M=3, lazy=False, lam=0.0),
    "euler3d_tet_p3_warp_ec": lambda: cases.euler_tet_case(p=3, M=2, lazy=False, warp=True,
                                                           interface="ec", ic="periodic"),
    "euler2d_tri_p3_lf": lambda: cases.euler_tri_case(p=3, M=3, lazy=False),
    "euler3d_hex_p3_ec": lambda: cases.euler_hex_case(p=3, M=2, lazy=False),
    "advdiff1d_p4": lambda: cases.advection_diffusion_case(d=1, p=4, M=6, lazy=False),
    "burgers2d_tri_p3": lambda: cases.burgers_tri_case(p=3, M=3, lazy=False),
}


def _close(a, b, scale, tol=1e-12):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) <= tol * scale


@pytest.mark.parametrize("name", sorted(CASES))
def test_functionals_match_oracle(name):
    solver, u0 = CASES[name]()
    try:
        u = cases.rough_state(solver, u0, seed=3)
        prob = oracle_problem(solver)
        law = prob["law"]
        du = oc.semi_discrete_residual(prob, u)
        WJ = prob["W"][None, :] * prob["J_q"]
        V = prob["V"]
        u_q = np.einsum("qp,kep->kqe", V, u)
        scale_u = float(np.sum(WJ) * np.max(np.abs(u)))
        scale_du = float(np.sum(WJ) * np.max(np.abs(du)) * max(1.0, np.max(np.abs(u))))

        pc = analysis.PrimaryConservationAnalysis(solver)
        assert _close(pc.evaluate_conservation(u), np.einsum("kq,kqe->e", WJ, u_q), scale_u)
        assert _close(pc.evaluate_conservation_residual(u), oc.conservation_residual(prob, du),
                      scale_du)

        ec = analysis.EntropyConservationAnalysis(solver)
        S_ref = float(np.sum(WJ * oc.entropy(law, u_q)))
        assert _close(ec.evaluate_conservation(u), [S_ref],
                      scale_u * max(1.0, np.max(np.abs(oc.entropy(law, u_q)))))
        w_scale = max(1.0, float(np.max(np.abs(oc.conservative_to_entropy(law, u_q)))))
        assert _close(ec.evaluate_conservation_residual(u), [oc.entropy_residual(prob, u, du)],
                      scale_du * w_scale)

        en = analysis.EnergyConservationAnalysis(solver)
        E_ref = np.zeros(u.shape[1])
        for k in range(u.shape[0]):
            E_ref += 0.5 * np.einsum("ep,pq,eq->e", u[k], oc.mass_matrix(prob, k), u[k])
        assert _close(en.evaluate_conservation(u), E_ref, scale_u * np.max(np.abs(u)), 1e-11)
        assert _close(en.evaluate_conservation_residual(u), oc.energy_residual(prob, u, du),
                      scale_du, 1e-11)

        # L2 distance between the state and a sampled "exact" field (error.jl:58-91)
        rng = np.random.default_rng(5)
        exact_q = u_q.transpose(0, 2, 1) + 0.1 * rng.standard_normal((u.shape[0], u.shape[1],
                                                                       V.shape[0]))
        err = solver.handle.functional("l2_error", exact_q=np.ascontiguousarray(exact_q))
        assert _close(err, oc.l2_error(prob, u, exact_q.transpose(0, 2, 1)), np.max(err))
    finally:
        solver.close()


def test_entropy_conservation_and_error_analysis_api():
    """EC interface flux: |(P w)^T M dudt| ~ round-off on the device (runtests.jl:95,108,142), and
    ErrorAnalysis.analyze of the projected initial data against the initial data itself."""
    solver, u0 = cases.euler_tet_case(p=3, M=2, lazy=False, warp=True, interface="ec",
                                      ic="periodic")
    try:
        ec = analysis.EntropyConservationAnalysis(solver)
        pc = analysis.PrimaryConservationAnalysis(solver)
        dS = ec.evaluate_conservation_residual(u0)[0]
        dU = pc.evaluate_conservation_residual()          # state already resident
        solver.handle.nodal_values(); solver.handle.time_derivative()
        assert abs(dS) < 1e-11 and np.max(np.abs(dU)) < 1e-11
    finally:
        solver.close()
    solver, u0 = cases.advection_tri_case(p=4, M=4, lazy=False)
    try:
        ea = analysis.ErrorAnalysis(solver)
        import math
        e = ea.analyze(u0, InitialDataSine(1.0, (2 * math.pi,) * 2), 0.0)
        assert 0.0 < e[0] < 1e-2         # projection error of a smooth field at p = 4
        en = ea.analyze(None, InitialDataSine(1.0, (2 * math.pi,) * 2), 0.0, normalize=True)
        assert abs(en[0] - e[0] / math.sqrt(ea.total_volume)) < 1e-15
    finally:
        solver.close()
