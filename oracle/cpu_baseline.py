"""CPU baseline (test/bench infrastructure only): times the restated reference residual on the
host cores for the north-star workload -- the SAME mesh and state as the GPU arm, each timed step
a bounded sample of it (a contiguous slab of elements; the whole mesh when the host is fast
enough).

kind = "port": the reference itself is Julia and cannot run here (no julia binary in the image
or on the GPU box), so the baseline is the oracle restatement -- the C/OpenMP loops of
oracle/c/sse_oracle.c (threads over elements, like the reference's ``Threads.@threads for k`` in
Solvers.jl:509-515).  Timing protocol of the reference's own benchmark driver
(src/Analysis/benchmark.jl:157-171): warm-up calls, then min and median of the samples.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_on(solver, u0, steps=3, warmup=1, budget_s=20.0, label=""):
    """Time the C/OpenMP oracle on ``solver``'s mesh and state ``u0``.

    Every step evaluates loops A and B on the elements [k0, k1) of the mesh; the range is sized
    from a calibration call so that ``warmup + steps`` steps take about ``budget_s`` seconds, and
    it is the whole mesh when that fits.  All host cores are used regardless of an inherited
    OMP_NUM_THREADS (torch.distributed.run exports OMP_NUM_THREADS=1)."""
    import c_oracle
    from bridge import oracle_problem
    prob = oracle_problem(solver)
    V = solver.spatial_discretization.reference_approximation.V
    warped = (V.A, V.B, getattr(V, "C", None), V.sigma_i) if hasattr(V, "sigma_i") else None
    c_oracle.set_threads(host_cores())
    fn = c_oracle.make_residual(prob, warped)
    cores = c_oracle.num_threads()
    N_e, N_c, N_p = u0.shape
    N_f = prob["N_f"]
    u = np.ascontiguousarray(u0)
    out = np.zeros_like(u)

    def span(k0, k1):
        """Element range whose traces loop B of [k0, k1) reads."""
        kk = prob["mapP"][:, k0:k1] // N_f
        return min(k0, int(kk.min())), max(k1, int(kk.max()) + 1)

    # calibration on a slab in the middle of the mesh (its neighbours' traces made valid first)
    n_cal = min(N_e, 16384)
    c0 = (N_e - n_cal) // 2
    fn.range(u, out, *span(c0, c0 + n_cal))
    t0 = time.perf_counter()
    fn.range(u, out, c0, c0 + n_cal)
    t_cal = time.perf_counter() - t0
    per_elt = t_cal / n_cal
    n_s = int(budget_s / max(1, steps + warmup) / per_elt)
    n_s = max(min(N_e, 8192), min(N_e, n_s))
    if n_s >= 0.9 * N_e:
        n_s = N_e
    k0 = (N_e - n_s) // 2
    k1 = k0 + n_s
    if n_s < N_e:
        s0, s1 = span(k0, k1)
        if (s1 - s0) > n_s + n_cal:      # traces outside what the calibration made valid
            fn.range(u, out, s0, s1)
    for _ in range(warmup):
        fn.range(u, out, k0, k1)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        fn.range(u, out, k0, k1)
        times.append(time.perf_counter() - t0)
    dof = n_s * N_c * N_p
    t_med, t_min = float(np.median(times)), float(np.min(times))
    t_mean = float(np.mean(times))
    whole = "the whole mesh" if n_s == N_e else f"elements [{k0}, {k1}) of the same mesh"
    return {"value": dof / t_mean, "value_median": dof / t_med, "value_best": dof / t_min,
            "ms_per_step": t_mean * 1e3, "cores": cores, "kind": "port", "N_e": N_e,
            "sample_elements": n_s, "sample_dof": dof,
            "sample": f"C/OpenMP restatement of the reference's two threaded element loops "
                      f"({cores} threads) on {label or 'the workload'}: each step = loops A+B on "
                      f"{whole} ({n_s} of {N_e} elements, {dof} DOF); "
                      f"{max(1, steps)} timed steps after {warmup} warm-up (mean; median "
                      f"{dof / t_med:.4g}, best {dof / t_min:.4g} DOF/s)"}


def run(M=6, warp=True, steps=3, warmup=1, budget_s=20.0):
    """Stand-alone: build the Tet p=4 Euler mesh of size M and time the oracle on it."""
    from sse_b200 import problems
    solver, u0 = problems.euler_tet_case(p=4, M=M, lazy=True, warp=warp)
    return run_on(solver, u0, steps=steps, warmup=warmup, budget_s=budget_s,
                  label=f"Tet p=4 Euler flux differencing, M={M}")
