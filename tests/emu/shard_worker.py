"""One rank of the CPU run of the library's sharded flow (tests/test_sharded_library_cpu.py).

TEST INFRASTRUCTURE: the product's CUDA sources compiled for the host (tests/emu) plus the file
transport of tests/emu/fake_nccl.c standing in for NCCL, one process per rank.

usage: shard_worker.py <rank> <world> <id dir> <case> <out.npz>"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def build_fake_nccl():
    import subprocess
    out_dir = os.path.join(HERE, "_build", "fake_nccl")
    out = os.path.join(out_dir, "libnccl.so.2")
    src = os.path.join(HERE, "fake_nccl.c")
    os.makedirs(out_dir, exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        tmp = out + ".%d.tmp" % os.getpid()
        subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-Wl,-soname,libnccl.so.2", "-o", tmp, src],
                       check=True)
        os.replace(tmp, out)
    return out


def problem(case):
    from sse_b200 import problems as cases
    if case == "euler_tet_p2":          # loop B = k_fluxdiff_nodal + k_project_tet
        return cases.euler_tet_case(p=2, M=6, lazy=True, warp=True, ic="periodic")
    if case == "adv_tet_p2":            # config 3 kernels: k_standard_tensor + k_project_tet
        return cases.advection_tet_case(p=2, M=6, lazy=True)
    if case == "euler_tri_p4":          # 2-D: one loop-B kernel
        return cases.euler_tri_case(p=4, M=8, lazy=True)
    if case == "advdiff2d_p3_br1":      # second order: two exchanges, main-stream flow
        return cases.advection_diffusion_case(d=2, p=3, M=8, lazy=True)
    raise ValueError(case)


def main():
    rank, world, id_dir, case, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    # the fake transport must be in the process before the library dlopen()s "libnccl.so.2"
    ctypes.CDLL(build_fake_nccl(), mode=ctypes.RTLD_GLOBAL)
    import build_emu
    from sse_b200 import device as dev
    from sse_b200 import problems as cases
    from sse_b200.distributed import DistributedResidual
    # SSE_EMU_ASAN=1: the AddressSanitizer build (every "device" buffer its own heap block)
    lib = dev.load_library(build_emu.build(asan=os.environ.get("SSE_EMU_ASAN") == "1"), allow_emulation=True)
    assert lib.sse_version() < 0
    dev._LIB = lib
    solver, u0 = problem(case)
    u = cases.rough_state(solver, u0, seed=5)
    nccl_id = id_dir.encode().ljust(128, b"\0")
    d = DistributedResidual(solver, rank=rank, world=world, device=0, backend="library", nccl_id=nccl_id)
    res = {"elements": d.elements, "interior": np.array(d.part.interior), "n_halo": d.part.n_halo}
    d.set_state(np.ascontiguousarray(u[d.elements]))
    for _ in range(2):            # twice: message sequence numbers, halo slots overwritten
        d.residual()
    d.sync()
    res["dudt"] = d.get_dudt()
    u_h = np.ascontiguousarray(u[d.elements])
    du_h = np.full_like(u_h, np.nan)
    d.residual_host(u_h, du_h)
    res["dudt_host"] = du_h
    if not d.second_order:
        d.set_state(np.ascontiguousarray(u[d.elements]))
        for _ in range(2):
            d.dev.shard_rk_step_ck54(1e-4)
        d.sync()
        res["state_ck54"] = d.dev.get_state()
    d.close()
    np.savez(out, **res)


if __name__ == "__main__":
    main()
