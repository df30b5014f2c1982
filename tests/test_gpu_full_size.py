"""Full-size property checks of the north-star configuration on the GPU (BASELINE configs[3]:
3-D Euler, p=4 tetrahedra, M=44 -> 511 104 curved elements, 89.4 M DOF), where the oracle is too
slow to serve as a checker: size-independent invariants of the scheme instead
(/root/reference/src/Analysis/conservation.jl:145-190).

  * free-stream preservation: a constant state has zero residual (metric identities + watertight
    connectivity over all 51 M facet nodes);
  * conservation: sum_k 1^T W J_k V dudt_k = 0 per variable;
  * entropy conservation with the EC interface flux: sum_k (P_k w)^T M_k dudt_k = 0;
  * determinism: two evaluations agree bitwise; the input is not modified.
"""
import numpy as np
import pytest
import torch

import cases
from sse_b200.solvers import semi_discrete_residual

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
M = 44


def _entropy_vars(u_q, g=1.4):
    rho, E = u_q[..., 0], u_q[..., 4]
    k = 0.5 * (u_q[..., 1] ** 2 + u_q[..., 2] ** 2 + u_q[..., 3] ** 2) / rho
    p = (g - 1) * (E - k)
    w = np.empty_like(u_q)
    w[..., 0] = (g - np.log(p / rho ** g)) / (g - 1) - k / p
    w[..., 1:4] = u_q[..., 1:4] / p[..., None]
    w[..., 4] = -rho / p
    return w


def test_full_size_invariants():
    solver, u0 = cases.euler_tet_case(p=4, M=M, lazy=False, warp=True, interface="ec", ic="tgv")
    try:
        sd = solver.spatial_discretization
        ra, gf = sd.reference_approximation, sd.geometric_factors
        N_e = sd.N_e
        assert N_e == 6 * M ** 3 == 511104
        V, W = ra.V.to_dense(), ra.W
        dudt = np.empty_like(u0)

        # ---- free stream: project a constant state, expect a vanishing residual
        const = np.array([1.2, 0.3, -0.2, 0.5, 3.0])
        c_modal = V.T @ (W[:, None] * np.ones((ra.N_q, 1)))            # modes of the constant 1
        # exact L2 projection of a constant on each element is the same coefficient vector
        # scaled per variable only if J is in the space; use the weight-adjusted projection the
        # scheme itself uses: u = V^T W 1 * const (V^T W V = I)
        u_const = np.ascontiguousarray(
            np.broadcast_to(const[None, :, None] * c_modal[None, None, :, 0], u0.shape))
        semi_discrete_residual(dudt, u_const, solver, 0.0)
        flux_scale = 3.0 * np.max(np.abs(gf.Lambda_q)) / np.min(gf.J_q)
        assert np.max(np.abs(dudt)) < 1e-9 * flux_scale, np.max(np.abs(dudt))

        # ---- a rough admissible state
        rng = np.random.default_rng(5)
        u = np.ascontiguousarray(u0 * (1.0 + 0.02 * (rng.random(u0.shape) - 0.5)))
        u_copy = u.copy()
        semi_discrete_residual(dudt, u, solver, 0.0)
        assert np.array_equal(u, u_copy)
        d2 = np.empty_like(u)
        semi_discrete_residual(d2, u, solver, 0.0)
        assert np.array_equal(d2, dudt)
        assert np.all(np.isfinite(dudt))

        # ---- conservation: g_k = V^T (W J_k); sum_k g_k . dudt_k[c]
        g = gf.J_q @ (W[:, None] * V)                                   # (N_e, N_p)
        cons = np.einsum("kp,kcp->c", g, dudt)
        scale = np.einsum("kp,kcp->c", np.abs(g), np.abs(dudt))
        assert np.all(np.abs(cons) < 1e-11 * scale), (cons, scale)

        # ---- entropy conservation (EC volume + EC interface flux), in element chunks
        total, tscale = 0.0, 0.0
        for s in range(0, N_e, 32768):
            e = min(s + 32768, N_e)
            J = gf.J_q[s:e]
            u_q = np.einsum("qp,kcp->kqc", V, u[s:e])
            w_q = _entropy_vars(u_q)
            A = np.einsum("qa,kq,qb->kab", V, W[None, :] / J, V)        # M_k^-1 (weight-adjusted)
            rhs = np.einsum("qp,kq,kqc->kpc", V, W[None, :] * J, w_q)   # V^T WJ w_q
            Pw = A @ rhs                                                 # projected entropy vars
            # M_k dudt_k: batched 35x35 solves -- checker arithmetic, done with torch on the GPU
            # because 511 104 LAPACK calls on the host take minutes
            At = torch.from_numpy(A).cuda()
            Bt = torch.from_numpy(np.ascontiguousarray(dudt[s:e].transpose(0, 2, 1))).cuda()
            Mdu = torch.linalg.solve(At, Bt).cpu().numpy()
            total += float(np.sum(Pw * Mdu))
            tscale += float(np.sum(np.abs(Pw * Mdu)))
        assert abs(total) < 1e-11 * tscale, (total, tscale)
    finally:
        solver.close()
