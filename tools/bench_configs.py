"""Secondary BASELINE.json configurations (parity-test cases, not the headline bench line):
device-resident residual time, DOF/s and algorithmic HBM GB/s for configs 1, 2, 3, 5."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from sse_b200 import problems as cases  # noqa: E402


def run(name, builder, bytes_per_elt, reps=10):
    t0 = time.time()
    solver, u0 = builder()
    h = solver.handle
    h.set_state(u0)
    h.time_residual(3)
    ms, ta, tb = h.time_residual(reps, split=True)
    N_e = u0.shape[0]
    out = dict(config=name, N_e=N_e, dof=int(u0.size), ms_per_residual=ms / reps,
               loop_a_ms=ta / reps, loop_b_ms=tb / reps, dof_per_s=u0.size / (ms / reps * 1e-3),
               algorithmic_GBps=bytes_per_elt * N_e / (ms / reps * 1e-3) / 1e9,
               bytes_per_element=bytes_per_elt, setup_s=round(time.time() - t0, 1))
    print(json.dumps(out), flush=True)
    solver.close()


which = sys.argv[1:] or ["1", "2", "3", "5"]
if "1" in which:   # 2-D advection, Tri p=4, 32x32 (the reference's CPU-runnable case)
    run("cfg1 adv2d tri p4 M=32", lambda: cases.advection_tri_case(p=4, M=32, lazy=False),
        8 * (2 * 15 + 4 * 25 + 2 * 15 + 25 + 3 * 15) + 4 * 15)
if "2" in which:   # 2-D Euler vortex, Tri p=4, M=256 -> 131072 elements
    run("cfg2 euler2d tri p4 M=256", lambda: cases.euler_tri_case(p=4, M=256, lazy=False),
        8 * (2 * 60 + 60 + 2 * 25 + 100 + 30 + 60 + 120) + 4 * 15)
if "3" in which:   # 3-D advection, Tet p=4 (M=55 is the full config: 998 250 elements)
    M = int(os.environ.get("CFG3_M", "40"))
    run(f"cfg3 adv3d tet p4 M={M}",
        lambda: cases.advection_tet_case(p=4, M=M, lazy=False, mapping_degree=2), 15760)
if "5" in which:   # advection-diffusion BR1, PhysicalOperator, 2-D Tri p=4
    Np, Nq, Nf = 15, 25, 15
    run("cfg5 advdiff2d tri p4 M=128",
        lambda: cases.advection_diffusion_case(d=2, p=4, M=128, lazy=False),
        2 * 8 * (2 * Np * Nq + Np * Nf))
