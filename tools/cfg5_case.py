"""Config 5 (advection-diffusion, BR1, PhysicalOperator) on one GPU: residual time and the share of
the measured HBM peak the per-element operator stream reaches.
usage: cfg5_case.py d p M [reps]"""
import sys
sys.path.insert(0, '.')
from sse_b200 import problems

d, p, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
solver, u0 = problems.advection_diffusion_case(d=d, p=p, M=M, lazy=False)
ra = solver.spatial_discretization.reference_approximation
Np, Nq, Nf = ra.N_p, ra.N_q, ra.N_f
h = solver.handle
h.set_state(u0)
h.time_residual(3)
ms, ta, tb = h.time_residual(reps, split=True)
B = 2 * 8 * (d * Np * Nq + Np * Nf)
gbs = B * u0.shape[0] / (ms / reps * 1e-3) / 1e9
print("cfg5 d=%d p=%d N_e=%d: %.4f ms/residual (A %.4f, B %.4f) -> %.0f GB/s algorithmic (%.3f of 6550)"
      % (d, p, u0.shape[0], ms / reps, ta / reps, tb / reps, gbs, gbs / 6550), flush=True)
solver.close()
