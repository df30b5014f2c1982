#!/usr/bin/env python
"""bench.py -- residual throughput of the BASELINE.json configurations on N B200s.

Headline (BASELINE.json `metric`, configs[3]): 3-D Euler, entropy-stable flux differencing,
tetrahedra, p = 4: residual evaluations per second x degrees of freedom (DOF = N_p N_c N_e).  A
"step" is one semi-discrete residual over the whole mesh (M = 44 -> 511 104 elements, 89.4 M DOF)
with the state already resident in HBM; `e2e` is the same call through the reference-facing
`semi_discrete_residual(dudt, u, solver, t)` with pinned HOST buffers (H2D of u and D2H of dudt
inside the timed region).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # CPU restatement of the reference on the host cores
    python bench.py --config 3              # another BASELINE config as the headline of the line

Prints ONE JSON line on rank 0.  Besides the contract keys it carries
  roofline      FP64 view of the dominant kernel (loop B is FP64-pipe bound), roofline_hbm next to it;
  check         driver-visible correctness of the run: conservation / entropy residuals of the
                computed dudt, a partition-independent digest of dudt, and (N > 1) an in-run
                comparison of the sharded residual with the single-GPU one on a small mesh;
  secondary     the HBM-bound BASELINE configs (3: sharded over the same N GPUs; N = 1 also
                2, 5 (p sweep) and 1) with their achieved fraction of the measured HBM peak.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# Algorithmic work per element.  Tet p=4 Euler flux differencing:
FLOP_PER_ELT = {"loop_b": 133.5e3 + 221.0e3, "loop_a": 75.0e3, "residual": 430.0e3,
                "volume": 133.5e3}
BYTES_PER_ELT_LOOP_B = 8 * (625 + 1125 + 125 + 300 + 100 + 500 + 500 + 175) + 4 * 100
# dram__bytes_read.sum + dram__bytes_write.sum of the loop-B kernel per element from the
# `ncu --set full` capture of this round (profiles/): not measurable inside this process
TRAFFIC_PER_ELT_LOOP_B = 24032.0
TRAFFIC_SOURCE = "ncu --set full capture (profiles/), scaled per element"
# Tet p=4 advection, standard form (config 3): u 280 + u_q 1000 rw + u_f 800 rw + exterior 800 +
# Lambda_q 9000 + J_q 1000 + J_f 800 + nJf 2400 + offsets 400 + dudt 280
BYTES_PER_ELT_CFG3 = 15760

METRIC = "3D Euler ES tet p=4 residual DOF/s"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


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
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- workloads
def workload(config, M, straight=False):
    """(description, builder(shard, device_geometry) -> (solver, u0)) of a BASELINE config."""
    from sse_b200 import problems
    if config == 4:
        M = M or 44
        desc = (f"3D Euler Taylor-Green vortex Ma=0.1 on (0,2pi)^3, ModalTensor(4) tetrahedra, "
                f"M={M} -> {6 * M ** 3} elements, {6 * M ** 3 * 35 * 5} DOF, "
                f"{'straight' if straight else 'ChanWarping(1/16) curved'} mesh, "
                f"FluxDifferencingForm (EC two-point flux, Lax-Friedrichs facets), "
                f"weight-adjusted mass solver")
        return desc, lambda shard=None, device_geometry=None: problems.euler_tet_case(
            p=4, M=M, lazy=True, warp=not straight, interface="lf", ic="tgv", shard=shard,
            device_geometry=device_geometry)
    if config == 3:
        M = M or 55
        desc = (f"3D linear advection on curved tetrahedra, ModalTensor(4), M={M} -> "
                f"{6 * M ** 3} elements, {6 * M ** 3 * 35} DOF, StandardForm (skew-symmetric) + "
                f"ReferenceOperator, Lax-Friedrichs facets")
        return desc, lambda shard=None, device_geometry=None: problems.advection_tet_case(
            p=4, M=M, lazy=True, mapping_degree=2, shard=shard)
    raise SystemExit(f"--config {config}: only 3 and 4 run as the headline line "
                     f"(1, 2 and 5 are reported in the `secondary` block)")


def digest(a, start):
    """Partition-independent checksum of a float64 array holding the global entries
    [start, start + a.size): sum over entries of (IEEE bit pattern x odd multiplier of the global
    index) modulo 2^64.  Integer arithmetic, so the sum over shards is exact and order-free:
    the value is the same for every element partition iff every entry is bitwise the same."""
    bits = np.ascontiguousarray(a).reshape(-1).view(np.uint64)
    idx = np.arange(start, start + bits.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        w = idx * np.uint64(0x9E3779B97F4A7C15) | np.uint64(1)
        return int(np.sum(bits * w, dtype=np.uint64))


class Dist:
    """torch.distributed helpers that degrade to no-ops at world = 1."""

    def __init__(self, world, rank):
        self.world, self.rank = world, rank

    def barrier(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out


def measure(builder, D, local_rank, steps, warmup, device_geometry=False, e2e=True,
            functionals=True):
    """Build this rank's shard, time `steps` device-resident residuals (CUDA events on the
    launching stream, max over ranks), the per-kernel split, the end-to-end call with pinned host
    buffers, and the check quantities of the computed dudt."""
    import torch
    from sse_b200.distributed import DistributedResidual
    rank, world = D.rank, D.world
    t_setup = time.time()
    solver, u0 = builder(shard=(rank, world) if world > 1 else None,
                         device_geometry=local_rank if device_geometry else None)
    N_e = solver.spatial_discretization.mesh.mapP.shape[1]
    N_c, N_p = u0.shape[1], u0.shape[2]
    dof = N_e * N_c * N_p
    dres = DistributedResidual(solver, rank=rank, world=world, device=local_rank)
    dres.set_state(u0)
    t_setup = time.time() - t_setup

    def barrier():
        dres.sync()
        D.barrier()

    for _ in range(warmup):
        dres.residual()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = dres.kernel_launches()
    ms_total = dres.timed_residuals(steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = dres.kernel_launches() - launches0
    ms_step = D.max(ms_total) / steps
    out = {"dof": dof, "N_e": N_e, "n_local": len(dres.elements), "ms_per_step": ms_step,
           "value": dof / (ms_step * 1e-3), "clocks": clocks, "gpu_launches": launches,
           "setup_s": round(t_setup, 1), "solver": solver, "u0": u0,
           "split": dres.split_times(max(3, min(steps, 10)))}

    # ---- check: functionals and digest of the dudt of the LAST residual (device-resident path)
    if functionals:
        dres.residual()
        dres.sync()
        du = dres.get_dudt()
        start = int(dres.elements[0]) * N_c * N_p
        cons = dres.dev.functional("conservation", "dudt")
        ent = dres.dev.functional("entropy_residual") if N_c > 1 else \
            dres.dev.functional("energy_residual")
        parts = D.gather({"digest": digest(du, start), "cons": [float(c) for c in cons],
                          "ent": float(np.sum(ent)), "l1": float(np.sum(np.abs(du))),
                          "finite": bool(np.all(np.isfinite(du)))})
        l1 = sum(p["l1"] for p in parts)
        out["check"] = {
            "dudt_digest_u64": f"{sum(p['digest'] for p in parts) % (1 << 64):016x}",
            "dudt_l1": l1,
            "conservation_residual": [sum(p["cons"][c] for p in parts) for c in range(N_c)],
            "entropy_residual" if N_c > 1 else "energy_residual": sum(p["ent"] for p in parts),
            "finite": all(p["finite"] for p in parts),
            "note": "digest = sum(bits(dudt[g]) * odd(g)) mod 2^64 over global DOF g: equal for "
                    "every element partition iff dudt is bitwise equal; conservation residual "
                    "is round-off of dudt_l1; the Lax-Friedrichs entropy residual is <= 0",
        }
        del du

    # ---- end to end through the public API with pinned host buffers
    if e2e:
        n_loc = dres.n_local_state
        u_host = torch.empty(n_loc, dtype=torch.float64, pin_memory=True)
        du_host = torch.empty(n_loc, dtype=torch.float64, pin_memory=True)
        u_np = u_host.numpy().reshape(dres.local_shape)
        du_np = du_host.numpy().reshape(dres.local_shape)
        u_np[...] = u0
        ke = max(3, min(steps, 10))
        for _ in range(2):
            dres.residual_host(u_np, du_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            dres.residual_host(u_np, du_np)
        barrier()
        te = D.max((time.perf_counter() - t0) / ke)
        # the floor of the host-buffer call: the same bytes over PCIe, both directions at once, all
        # ranks together, no kernels (e2e >= max(this, the device-resident step))
        dv = torch.device("cuda", local_rank)
        d_in = torch.empty(n_loc, dtype=torch.float64, device=dv)
        d_out = torch.zeros(n_loc, dtype=torch.float64, device=dv)
        s_up, s_dn = torch.cuda.Stream(dv), torch.cuda.Stream(dv)
        tp = []
        for rep in range(4):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                d_in.copy_(u_host, non_blocking=True)
            with torch.cuda.stream(s_dn):
                du_host.copy_(d_out, non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
            tp.append(D.max(time.perf_counter() - t0))
        du_np[...] = 0.0
        dres.residual_host(u_np, du_np)          # restore dudt for the digest check below
        t_pcie = min(tp[1:])
        del d_in, d_out
        out["e2e"] = {"value": dof / te, "unit": "DOF/s", "h2d_bytes_per_step": 8 * dof,
                      "d2h_bytes_per_step": 8 * dof, "ms_per_step": te * 1e3,
                      "flow": dres.host_flow_name(),
                      "pcie_floor_ms": t_pcie * 1e3,
                      "pcie_GBps_each_way_per_gpu": 8 * n_loc / t_pcie / 1e9}
        if functionals:   # the host-buffer path must give the very same dudt
            start = int(dres.elements[0]) * N_c * N_p
            parts = D.gather(digest(du_np, start))
            out["check"]["e2e_digest_matches"] = (
                f"{sum(parts) % (1 << 64):016x}" == out["check"]["dudt_digest_u64"])
    out["dres"] = dres
    return out


def small_mesh_parity(D, local_rank, M=8):
    """N > 1: the sharded residual (halo over NCCL) against the single-GPU residual of the same
    small mesh, inside this run: rank 0 evaluates the whole mesh on its GPU, all ranks their
    shards; the digests must be equal (bitwise-equal dudt)."""
    from sse_b200 import problems
    from sse_b200.distributed import DistributedResidual
    rank, world = D.rank, D.world
    out = {"mesh": f"Tet p=4 Euler, M={M} ({6 * M ** 3} elements), rough state"}
    solver, u0 = problems.euler_tet_case(p=4, M=M, lazy=True, warp=True, interface="lf", ic="tgv")
    u = problems.rough_state(solver, u0, seed=7)
    if rank == 0:
        one = DistributedResidual(solver, rank=0, world=1, device=local_rank)
        one.set_state(u)
        one.residual()
        one.sync()
        out["digest_single_gpu"] = f"{digest(one.get_dudt(), 0):016x}"
        one.close()
    sh = DistributedResidual(solver, rank=rank, world=world, device=local_rank)
    sh.set_state(u[sh.elements])
    sh.residual()
    sh.sync()
    start = int(sh.elements[0]) * u.shape[1] * u.shape[2]
    parts = D.gather(digest(sh.get_dudt(), start))
    sh.close()
    out["digest_sharded"] = f"{sum(parts) % (1 << 64):016x}"
    if rank == 0:
        out["match"] = out["digest_sharded"] == out["digest_single_gpu"]
    return out


def secondary_single_gpu(reps=10):
    """BASELINE configs 2, 5 (p = 2..8) and 1 on one GPU: device-resident residual time and the
    algorithmic HBM GB/s (bytes per element from DESIGN.md §5) -- the HBM-bound half of the
    north star.  Setup is host-side NumPy, so the meshes are kept moderate."""
    from sse_b200 import problems
    peaks, _ = read_peaks()
    rows = []

    def run(name, builder, bytes_per_elt, kernel):
        t0 = time.time()
        solver, u0 = builder()
        h = solver.handle
        h.set_state(u0)
        h.time_residual(3)
        ms, ta, tb = h.time_residual(reps, split=True)
        N_e = u0.shape[0]
        gbs = bytes_per_elt * N_e / (ms / reps * 1e-3) / 1e9
        rows.append({"config": name, "N_e": N_e, "dof": int(u0.size),
                     "ms_per_residual": ms / reps, "loop_a_ms": ta / reps, "loop_b_ms": tb / reps,
                     "dof_per_s": u0.size / (ms / reps * 1e-3), "algorithmic_GBps": gbs,
                     "hbm_frac": gbs / peaks["hbm_gbs"], "bytes_per_element": bytes_per_elt,
                     "kernel": kernel, "setup_s": round(time.time() - t0, 1)})
        solver.close()

    run("cfg2 euler2d tri p4 flux differencing M=256",
        lambda: problems.euler_tri_case(p=4, M=256, lazy=False),
        8 * (2 * 60 + 60 + 2 * 25 + 100 + 30 + 60 + 120) + 4 * 15,
        "k_nodal_tensor<2,5,Euler> + k_fluxdiff_tensor<2,5,Euler>")
    # the reference's own 3-D Euler test (NodalTensor Hex p=4, EC interface flux,
    # test/euler_3d.jl, M = 2) scaled up to 64^3 / 8 = 32 768 hexahedra; FP64-bound like config 4:
    # 750 pair fluxes + 150 interface fluxes per element, ~155 kflop, 2 x 5000 + 11 000 B
    run("hex: euler3d NodalTensor Hex p4 flux differencing (diagonal-E) M=32",
        lambda: problems.euler_hex_case(p=4, M=32, lazy=False),
        8 * (2 * 625 + 625 + 1125 + 125 + 150 + 450 + 2 * 750) + 4 * 150,
        "k_nodal_values<3,Euler> + k_fluxdiff_tensor<3,5,Euler,non-collapsed>")
    for p in range(2, 9):
        n1 = p + 1
        Np, Nq, Nf = n1 * (n1 + 1) // 2, n1 * n1, 3 * n1
        M = 128 if p <= 5 else 64
        # VOL (d N_p N_q) + FAC (N_p N_f) streamed by both BR1 launches
        run(f"cfg5 advdiff2d tri p{p} BR1 PhysicalOperator M={M}",
            lambda p=p, M=M: problems.advection_diffusion_case(d=2, p=p, M=M, lazy=False),
            2 * 8 * (2 * Np * Nq + Np * Nf), "k_physical<2,adv> x2")
    run("cfg1 adv2d tri p4 M=32 (the reference's CPU-runnable case)",
        lambda: problems.advection_tri_case(p=4, M=32, lazy=False),
        8 * (2 * 15 + 4 * 25 + 2 * 15 + 25 + 3 * 15) + 4 * 15, "k_nodal_tensor<2,5,adv> + k_standard_ref")
    return rows


def reference_arm(args):
    """--impl reference: the restated reference CPU path on all host cores, on the SAME mesh and
    state as the GPU arm; each step a bounded sample of it (see oracle/cpu_baseline.py)."""
    import __graft_entry__ as ge
    ge.build_oracle()
    import cpu_baseline
    desc, builder = workload(args.config, args.M, args.straight)
    if args.config != 4:
        raise SystemExit("the C/OpenMP restatement covers the flux-differencing path (config 4)")
    t0 = time.time()
    solver, u0 = builder()
    t_setup = time.time() - t0
    res = cpu_baseline.run_on(solver, u0, steps=args.steps, warmup=args.warmup,
                              budget_s=args.cpu_budget, label=desc)
    line = {
        "impl": "reference", "metric": METRIC, "unit": "DOF/s", "value": res["value"],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc},
        "cpu_baseline": {"value": res["value"], "unit": "DOF/s", "cores": res["cores"],
                         "kind": res["kind"], "sample": res["sample"],
                         "value_median": res["value_median"], "value_best": res["value_best"]},
        "e2e": {"value": res["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "run": {"setup_s": round(t_setup, 1),
                                   "sample_elements": res["sample_elements"]},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, help="BASELINE.json config (1-based): 3 or 4")
    ap.add_argument("--M", type=int, default=0, help="cubes per direction (6 M^3 tetrahedra)")
    ap.add_argument("--straight", action="store_true", help="straight-sided mesh (default: warped)")
    ap.add_argument("--cpu-budget", type=float, default=None,
                    help="seconds of CPU work for the cpu_baseline leg / the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--device-geometry", action="store_true",
                    help="evaluate the geometric factors on the GPU (sse_geometry_build) at setup")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3          # timing rule: at least three warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.cpu_budget is None:
            args.cpu_budget = 150.0
        if rank == 0:
            reference_arm(args)
        return
    if args.cpu_budget is None:
        args.cpu_budget = 20.0

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
    # host threads and pinned buffers of this rank on the GPU's own NUMA node (end-to-end path)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sse_b200.device import bind_host_to_gpu_numa_node
    numa = bind_host_to_gpu_numa_node(local_rank) if os.environ.get("SSE_B200_NUMA_BIND", "1") != "0" else {"bound": False}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    D = Dist(world, rank)
    D.barrier()
    from sse_b200 import device as dev

    desc, builder = workload(args.config, args.M, args.straight)
    main_r = measure(builder, D, local_rank, args.steps, args.warmup,
                     device_geometry=args.device_geometry, e2e=not args.no_e2e,
                     functionals=not args.no_check)
    fp64_peak = dev.measure_fp64_peak(local_rank)
    peaks, peak_src = read_peaks()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == 4:
        import cpu_baseline
        cpu = cpu_baseline.run_on(main_r["solver"], main_r["u0"], steps=3, warmup=1,
                                  budget_s=args.cpu_budget, label="the same mesh")
    main_r.pop("dres").close()
    main_r.pop("solver")
    main_r.pop("u0")

    parity = None
    if world > 1 and not args.no_check and args.config == 4:
        parity = small_mesh_parity(D, local_rank)

    secondary = None
    if not args.no_secondary and args.config == 4:
        secondary = {}
        # config 3 sharded over the same GPUs (HBM-bound operator kernels)
        d3, b3 = workload(3, 0)
        r3 = measure(b3, D, local_rank, max(5, min(args.steps, 20)), 3, e2e=False,
                     functionals=not args.no_check)
        r3.pop("dres").close()
        r3.pop("solver")
        r3.pop("u0")
        n3 = r3["n_local"]
        t3 = (r3["split"]["loop_a_ms"] + r3["split"]["loop_b_ms"]) * 1e-3
        secondary["cfg3"] = {
            "workload": d3, "n_gpus": world, "ms_per_step": r3["ms_per_step"],
            "dof_per_s": r3["value"], "kernel_ms": r3["split"],
            "algorithmic_GBps_per_gpu": BYTES_PER_ELT_CFG3 * n3 / t3 / 1e9,
            "hbm_frac": BYTES_PER_ELT_CFG3 * n3 / t3 / 1e9 / peaks["hbm_gbs"],
            "bytes_per_element": BYTES_PER_ELT_CFG3, "setup_s": r3["setup_s"],
            "kernels": "k_nodal_batched<3,5,adv> + k_standard_tensor<3,5,adv>",
            "check": r3.get("check"),
        }
        if rank == 0 and world == 1:
            secondary["single_gpu"] = secondary_single_gpu()

    if rank == 0:
        split = main_r["split"]
        n_loc_e = main_r["n_local"]
        tb, ta = split["loop_b_ms"] * 1e-3, split["loop_a_ms"] * 1e-3
        if args.config == 4:
            hbm_ach = BYTES_PER_ELT_LOOP_B * n_loc_e / tb / 1e9
            fp64_ach = FLOP_PER_ELT["loop_b"] * n_loc_e / tb / 1e12
            roofline = {
                "kernel": "k_fluxdiff_tensor<3,5,Euler,collapsed,8> (loop B: interface flux + "
                          "volume flux differencing + facet correction + lift + mass solve)",
                "bound": "fp64", "achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": fp64_ach / fp64_peak, "flop_per_element": FLOP_PER_ELT["loop_b"],
                "traffic": TRAFFIC_PER_ELT_LOOP_B * n_loc_e, "traffic_unit": "bytes/launch",
                "traffic_source": TRAFFIC_SOURCE,
                "peak_source": "FP64 DFMA-chain peak measured in this run "
                               "(sse_measure_fp64_peak; MEASURED_PEAKS.json has no FP64 figure)",
                "loop_a": {"kernel": "k_nodal_tensor<3,5,Euler,PROJ_CT=2>", "ms": split["loop_a_ms"],
                           "fp64_frac": FLOP_PER_ELT["loop_a"] * n_loc_e / ta / 1e12 / fp64_peak},
                "whole_residual_tflops": FLOP_PER_ELT["residual"] * n_loc_e / (ta + tb) / 1e12,
            }
            roofline_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"],
                            "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"],
                            "peak_source": peak_src,
                            "note": "secondary view: loop B is FP64-pipe bound (AI ~ 12 flop/B)"}
        else:
            gbs = BYTES_PER_ELT_CFG3 * n_loc_e / (ta + tb) / 1e9
            roofline = {"kernel": "k_nodal_batched<3,5,adv> + k_standard_tensor<3,5,adv>",
                        "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src}
            roofline_hbm = None
        line = {
            "metric": METRIC if args.config == 4 else "3D advection tet p=4 residual DOF/s",
            "value": main_r["value"], "unit": "DOF/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc},
            "run": {"parallelism": f"element-sharded x{world}, facet-trace halo over NCCL",
                    "l2": "inputs (geometry + state, > 6 GB per GPU) far exceed the 126 MB L2; "
                          "no flush needed",
                    "setup_s": main_r["setup_s"],
                    "geometry": "device (sse_geometry_build)" if args.device_geometry else "host",
                    "numa_rank0": numa},
            "roofline": roofline, "kernel_ms": split, "clocks": main_r["clocks"],
            "gpu_launches": main_r["gpu_launches"], "e2e": main_r.get("e2e"),
        }
        if roofline_hbm:
            line["roofline_hbm"] = roofline_hbm
        if "check" in main_r:
            line["check"] = main_r["check"]
            if parity is not None:
                line["check"]["sharded_vs_single_gpu"] = parity
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": "DOF/s", "cores": cpu["cores"],
                                    "kind": cpu["kind"], "sample": cpu["sample"]}
        if secondary is not None:
            line["secondary"] = secondary
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
        if parity is not None and not parity.get("match", False):
            raise SystemExit("sharded residual differs from the single-GPU residual")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
