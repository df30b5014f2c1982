"""Multi-GPU parity check (run under torchrun): the element-sharded residual with the NCCL halo
exchange must equal the single-domain oracle residual on every shard."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist
from sse_b200 import problems as cases
import sse_oracle as oc
from bridge import oracle_problem
from sse_b200.distributed import DistributedResidual

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
worst = 0.0
for name, (solver, u0) in {
        "euler_tet_p4_warp": cases.euler_tet_case(p=4, M=4, lazy=True, warp=True, ic="periodic"),
        "euler_tri_p4": cases.euler_tri_case(p=4, M=8, lazy=True),
        # config 3: scalar standard form on tetrahedra (k_standard_tensor + the batched projection)
        "adv_tet_p4": cases.advection_tet_case(p=4, M=4, lazy=True),
        # second-order (BR1): two halo exchanges, u_f then q_f
        "advdiff1d_p4_br1": cases.advection_diffusion_case(d=1, p=4, M=16, lazy=True),
        "advdiff2d_p3_br1": cases.advection_diffusion_case(d=2, p=3, M=8, lazy=True)}.items():
    u = cases.rough_state(solver, u0, seed=3)
    ref = oc.semi_discrete_residual(oracle_problem(solver), u)
    # "library": sse_shard_create / sse_shard_residual (partition + NCCL inside the C ABI);
    # "python": the torch.distributed P2P flow of distributed.py
    for backend in ("library", "python"):
        d = DistributedResidual(solver, rank=rank, world=world, device=lr, backend=backend)
        d.set_state(u[d.elements])
        for _ in range(2):
            d.residual()
        d.sync()
        out = d.get_dudt()
        err = float(np.max(np.abs(out - ref[d.elements])) / np.max(np.abs(ref)))
        u_h = np.ascontiguousarray(u[d.elements]); du_h = np.full_like(u_h, np.nan)
        for _ in range(2):
            d.residual_host(u_h, du_h)
        err2 = float(np.max(np.abs(du_h - ref[d.elements])) / np.max(np.abs(ref)))
        same = bool(np.array_equal(du_h, out))
        t = torch.tensor([max(err, err2)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name} [{backend}]: world={world} max rel err over ranks = {t.item():.3e}, host path "
                  f"bitwise = {same} (interior {d.part.interior}, halo {d.part.n_halo})", flush=True)
        worst = max(worst, t.item())
        # low-storage RK step on the shards (sse_shard_rk_step_ck54: the update fused into the
        # epilogue of every loop-B range of the flow) against the single-handle step
        if backend == "library" and name in ("euler_tet_p4_warp", "adv_tet_p4", "euler_tri_p4"):
            from sse_b200.device import DeviceResidual
            dt = 1e-4
            d.set_state(u[d.elements])
            for _ in range(2):
                d.dev.shard_rk_step_ck54(dt)
            d.sync()
            got = d.dev.get_state()
            one = DeviceResidual(solver, device=lr)
            one.set_state(u)
            for _ in range(2):
                one.rk_step_ck54(dt)
            one.sync()
            want = one.get_state()[d.elements]
            one.close()
            e3 = torch.tensor([float(np.max(np.abs(got - want)) / np.max(np.abs(want)))],
                              device="cuda", dtype=torch.float64)
            dist.all_reduce(e3, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"{name} [library]: two sharded CK54 steps vs single handle: max rel diff "
                      f"{e3.item():.3e}", flush=True)
            worst = max(worst, e3.item())
        d.close()
dist.barrier()
dist.destroy_process_group()
assert worst < 1e-12, worst
