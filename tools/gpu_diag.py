"""Diagnostic: run every parity case on the GPU and print the relative errors (no early exit)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import sse_oracle as oc
from bridge import oracle_problem
import cases
from test_gpu_parity import CASES
from sse_b200.solvers import semi_discrete_residual

for name in sorted(CASES):
    try:
        t0 = time.time()
        solver, u0 = CASES[name]()
        prob = oracle_problem(solver)
        out = []
        for u in (u0, cases.rough_state(solver, u0, seed=1)):
            dudt = np.full_like(u, np.nan)
            semi_discrete_residual(dudt, u, solver, 0.0)
            ref = oc.semi_discrete_residual(prob, u)
            out.append(np.max(np.abs(dudt - ref)) / np.max(np.abs(ref)))
        print(f"{name:28s} rel err smooth {out[0]:.3e} rough {out[1]:.3e}  ({time.time()-t0:.1f}s)", flush=True)
        solver.close()
    except Exception as e:
        print(f"{name:28s} FAILED: {e}", flush=True)
        traceback.print_exc()
