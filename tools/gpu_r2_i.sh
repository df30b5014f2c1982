#!/bin/bash
# Round 2, session I: config 5 sweep (k_physical with the operators staged through shared memory)
mkdir -p gpurun_out
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from sse_b200 import problems
for d, p, M in [(2, 2, 128), (2, 3, 128), (2, 4, 128), (2, 5, 128), (2, 6, 64), (2, 7, 64), (2, 8, 64), (1, 4, 65536)]:
    solver, u0 = problems.advection_diffusion_case(d=d, p=p, M=M, lazy=False)
    ra = solver.spatial_discretization.reference_approximation
    Np, Nq, Nf = ra.N_p, ra.N_q, ra.N_f
    h = solver.handle
    h.set_state(u0); h.time_residual(3)
    ms, ta, tb = h.time_residual(10, split=True)
    B = 2 * 8 * (d * Np * Nq + Np * Nf)
    print("cfg5 d=%d p=%d N_e=%d: %.4f ms/residual -> %.0f GB/s algorithmic (%.3f of 6550)" % (d, p, u0.shape[0], ms / 10, B * u0.shape[0] / (ms / 10 * 1e-3) / 1e9, B * u0.shape[0] / (ms / 10 * 1e-3) / 1e9 / 6550), flush=True)
    solver.close()
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "physical or advdiff or viscous" 2>&1 | tail -2
