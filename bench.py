#!/usr/bin/env python
"""bench.py -- residual throughput of the north-star configuration on N B200s.

Metric (BASELINE.json): 3-D Euler, entropy-stable flux differencing, tetrahedra, p = 4:
residual evaluations per second x degrees of freedom (DOF = N_p * N_c * N_e).  A "step" is one
semi-discrete residual evaluation over the whole mesh (configs[3]: M = 44 -> 511 104 elements,
89.4 M DOF) with the state already resident in HBM; `e2e` is the same call made through the
reference-facing `semi_discrete_residual(dudt, u, solver, t)` with pinned HOST buffers
(H2D of u and D2H of dudt inside the timed region).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # CPU restatement of the reference on the host cores

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# algorithmic work per element, Tet p=4 Euler flux differencing
FLOP_PER_ELT = {"loop_b": 133.5e3 + 221.0e3, "loop_a": 75.0e3, "residual": 430.0e3}
BYTES_PER_ELT_LOOP_B = 8 * (625 + 1125 + 125 + 300 + 100 + 500 + 500 + 175) + 4 * 100
# dram__bytes_read.sum + dram__bytes_write.sum of the loop-B kernel per element, from the
# `ncu --set full` capture at M=16 (profiles/r1_fluxdiff_tensor_v12.md): 590.6 MB / 24 576
TRAFFIC_PER_ELT_LOOP_B = 24032.0


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                 str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        sm.sort()
        # median over the samples taken under load (upper half of the clock samples)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(M, warp, lazy=True, shard=None, device_geometry=None):
    import cases
    return cases.euler_tet_case(p=4, M=M, lazy=lazy, warp=warp, interface="lf", ic="tgv",
                                shard=shard, device_geometry=device_geometry)


def cpu_reference_arm(args):
    """--impl reference: the restated reference CPU path on the host cores (bounded sample)."""
    import cpu_baseline
    res = cpu_baseline.run(M=args.cpu_m, warp=not args.straight, steps=args.steps,
                           warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "3D Euler ES tet p=4 residual DOF/s", "unit": "DOF/s",
        "value": res["value"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D Euler TGV, Tet p=4 flux differencing (EC volume + LF facets); "
                               f"bounded sample M={args.cpu_m} -> {res['N_e']} elements"},
        "cpu_baseline": {"value": res["value"], "unit": "DOF/s", "cores": res["cores"],
                         "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--M", type=int, default=44, help="cubes per direction (6 M^3 tetrahedra)")
    ap.add_argument("--straight", action="store_true", help="straight-sided mesh (default: warped)")
    ap.add_argument("--cpu-m", type=int, default=16, help="mesh size of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--device-geometry", action="store_true",
                    help="evaluate the geometric factors on the GPU (sse_geometry_build) at setup")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            cpu_reference_arm(args)
        return

    # native libraries (NCCL prints a version banner) must not write to stdout: the driver reads
    # ONE JSON line from it.  Route fd 1 to stderr until the line is printed.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from sse_b200 import device as dev
    from sse_b200.distributed import DistributedResidual

    t_setup = time.time()
    # every rank builds the (cheap) global connectivity but only its own shard's geometry
    solver, u0 = build_problem(args.M, warp=not args.straight, lazy=True,
                               shard=(rank, world) if world > 1 else None,
                               device_geometry=local_rank if args.device_geometry else None)
    N_e = solver.spatial_discretization.mesh.mapP.shape[1]
    N_c, N_p = u0.shape[1], u0.shape[2]
    dof = N_e * N_c * N_p
    dres = DistributedResidual(solver, rank=rank, world=world, device=local_rank)
    u0_local = u0 if world > 1 else u0[dres.elements]
    dres.set_state(u0_local)
    t_setup = time.time() - t_setup

    def barrier():
        dres.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K timed steps
    for _ in range(args.warmup):
        dres.residual()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = dres.kernel_launches()
    ms_total = dres.timed_residuals(args.steps)          # CUDA events on the launching stream
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = dres.kernel_launches() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = dof / (ms_step * 1e-3)

    # ---- per-kernel split (second pass; loop A vs loop B) for the roofline of the dominant kernel
    split = dres.split_times(max(3, min(args.steps, 10)))
    fp64_peak = dev.measure_fp64_peak(local_rank)

    # ---- end to end through the public API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        n_loc = dres.n_local_state
        u_host = torch.empty(n_loc, dtype=torch.float64, pin_memory=True)
        du_host = torch.empty(n_loc, dtype=torch.float64, pin_memory=True)
        u_np = u_host.numpy().reshape(dres.local_shape)
        du_np = du_host.numpy().reshape(dres.local_shape)
        u_np[...] = u0_local
        ke = max(3, min(args.steps, 10))
        for _ in range(2):
            dres.residual_host(u_np, du_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            dres.residual_host(u_np, du_np)
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / ke], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": dof / float(te.item()), "unit": "DOF/s",
               "h2d_bytes_per_step": 8 * dof, "d2h_bytes_per_step": 8 * dof,
               "ms_per_step": float(te.item()) * 1e3}

    if rank == 0:
        peaks, peak_src = read_peaks()
        n_loc_e = len(dres.elements)
        tb = split["loop_b_ms"] * 1e-3
        ta = split["loop_a_ms"] * 1e-3
        hbm_ach = BYTES_PER_ELT_LOOP_B * n_loc_e / tb / 1e9
        fp64_ach = FLOP_PER_ELT["loop_b"] * n_loc_e / tb / 1e12
        line = {
            "metric": "3D Euler ES tet p=4 residual DOF/s", "value": value, "unit": "DOF/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": (f"3D Euler Taylor-Green vortex Ma=0.1 on (0,2pi)^3, ModalTensor(4) "
                             f"tetrahedra, M={args.M} -> {N_e} elements, {dof} DOF, "
                             f"{'straight' if args.straight else 'ChanWarping(1/16) curved'} mesh, "
                             f"FluxDifferencingForm (EC two-point flux, Lax-Friedrichs facets), "
                             f"weight-adjusted mass solver"),
                "parallelism": f"element-sharded x{world}, facet-trace halo over NCCL",
                "l2": "inputs (6.5 GB geometry + state) far exceed the 126 MB L2; no flush needed",
                "setup_s": round(t_setup, 1),
                "geometry": "device (sse_geometry_build)" if args.device_geometry else "host",
            },
            "roofline": {
                "kernel": "k_fluxdiff_tensor<3,5,Euler,collapsed,8> (loop B: interface flux + "
                          "volume flux differencing + facet correction + lift + mass solve)",
                "bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": hbm_ach / peaks["hbm_gbs"],
                "traffic": TRAFFIC_PER_ELT_LOOP_B * n_loc_e, "traffic_unit": "bytes/launch",
                "traffic_source": "ncu --set full at M=16, scaled per element",
                "peak_source": peak_src,
                "note": "this kernel is FP64-pipe bound by design (AI ~ 12 flop/B); see roofline_fp64",
            },
            "roofline_fp64": {
                "bound": "fp64", "achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": fp64_ach / fp64_peak, "flop_per_element": FLOP_PER_ELT["loop_b"],
                "peak_source": "measured in this run (sse_measure_fp64_peak, DFMA chains)",
                "whole_residual_tflops": FLOP_PER_ELT["residual"] * n_loc_e / (ta + tb) / 1e12,
            },
            "kernel_ms": split, "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
        }
        if not args.no_cpu_baseline:
            import cpu_baseline
            res = cpu_baseline.run(M=args.cpu_m, warp=not args.straight, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": res["value"], "unit": "DOF/s", "cores": res["cores"],
                                    "kind": res["kind"], "sample": res["sample"]}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dres.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
