"""CPU baseline (test/bench infrastructure only): times the restated reference residual on the
host cores for a bounded sample of the north-star workload.

kind = "port": the reference itself is Julia and cannot run here (no julia binary in the image
or on the GPU box), so the baseline is the oracle restatement -- the C/OpenMP loops of
oracle/c/sse_oracle.c when built (threads over elements, like the reference's
``Threads.@threads for k`` in Solvers.jl:509-515), else the NumPy oracle.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(M=6, warp=True, steps=3, warmup=1):
    import cases
    import sse_oracle as oc
    from bridge import oracle_problem
    solver, u0 = cases.euler_tet_case(p=4, M=M, lazy=True, warp=warp)
    prob = oracle_problem(solver)
    dof = u0.size
    try:
        import c_oracle
        V = solver.spatial_discretization.reference_approximation.V
        warped = (V.A, V.B, getattr(V, "C", None), V.sigma_i) if hasattr(V, "sigma_i") else None
        fn, cores, impl = (c_oracle.make_residual(prob, warped), c_oracle.num_threads(),
                           "C/OpenMP")
    except Exception:
        fn, cores, impl = (lambda u: oc.semi_discrete_residual(prob, u)), 1, "NumPy"
    for _ in range(warmup):
        fn(u0)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        fn(u0)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    return {"value": dof / t, "ms_per_step": t * 1e3, "cores": cores, "kind": "port",
            "N_e": u0.shape[0],
            "sample": f"{impl} restatement of the reference loops, Tet p=4 Euler flux "
                      f"differencing, M={M} ({u0.shape[0]} elements, {dof} DOF), median of "
                      f"{max(1, steps)} residuals"}
