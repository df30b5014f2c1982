"""The library's element-sharded flow (csrc/tu_shard.cu: sse_shard_plan_build / sse_shard_create /
sse_shard_residual / sse_shard_rk_step_ck54) on the CPU, one PROCESS per rank: the product's CUDA
sources compiled for the host (tests/emu) and the file transport of tests/emu/fake_nccl.c in place
of NCCL.  Covers what otherwise only runs on >= 2 GPUs (tools/dist_check.py): the partition, pack /
exchange / unpack, the interior range overlapping the exchange, the two boundary ranges on their
own streams in every SSE_B200_SHARD_STREAMS mode, the host-buffer flow and the fused Runge-Kutta
update on shards -- against the oracle (1e-12) and bitwise against the single-handle run."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import cases
import sse_oracle as oc
from bridge import oracle_problem

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "emu", "shard_worker.py")
sys.path.insert(0, os.path.join(HERE, "emu"))


def _run_world(case, world, mode, asan_rt=None):
    import shard_worker
    shard_worker.build_fake_nccl()          # once, before the ranks race for it
    import build_emu
    build_emu.build(asan=asan_rt is not None)
    tmp = tempfile.mkdtemp(prefix="sse_shard_")
    try:
        env = dict(os.environ, SSE_B200_SHARD_STREAMS=str(mode), FAKE_NCCL_TIMEOUT_S="400")
        env.pop("SSE_B200_LIB", None)
        if asan_rt:
            env.update(SSE_EMU_ASAN="1", LD_PRELOAD=asan_rt,
                       ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=0")
        procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), tmp, case,
                                   os.path.join(tmp, f"out{r}.npz")], env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
                 for r in range(world)]
        outs = []
        for p in procs:
            try:
                log, _ = p.communicate(timeout=600)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            assert "ERROR: AddressSanitizer" not in log.decode(), log.decode()[-3000:]
            assert p.returncode == 0, log.decode()[-3000:]
        for r in range(world):
            outs.append(dict(np.load(os.path.join(tmp, f"out{r}.npz"))))
        return outs
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


@pytest.fixture(scope="module")
def single_handle():
    """Single-handle results of the same problems on the emulator (cached per case)."""
    import build_emu
    from sse_b200 import device as dev
    import shard_worker
    lib = dev.load_library(build_emu.build(), allow_emulation=True)
    saved = dev._LIB
    dev._LIB = lib
    cache = {}

    def get(case):
        if case not in cache:
            solver, u0 = shard_worker.problem(case)
            u = cases.rough_state(solver, u0, seed=5)
            d = dev.DeviceResidual(solver)
            try:
                dudt = np.full_like(u, np.nan)
                d.residual_host(u, dudt)
                state = None
                if solver.law_desc["kind"] not in ("advection_diffusion", "viscous_burgers"):
                    d.set_state(u)
                    for _ in range(2):
                        d.rk_step_ck54(1e-4)
                    d.sync()
                    state = d.get_state()
            finally:
                d.close()
            ref = oc.semi_discrete_residual(oracle_problem(solver), u)
            cache[case] = (dudt, state, ref)
        return cache[case]

    try:
        yield get
    finally:
        dev._LIB = saved


@pytest.mark.parametrize("case,world,mode", [
    ("euler_tet_p2", 2, 1), ("euler_tet_p2", 3, 2),
    ("adv_tet_p2", 2, 1),
    ("euler_tri_p4", 2, 0), ("euler_tri_p4", 2, 1), ("euler_tri_p4", 4, 2),
    ("advdiff2d_p3_br1", 2, 1),
])
def test_library_sharded_flow_on_cpu(single_handle, case, world, mode):
    dudt1, state1, ref = single_handle(case)
    outs = _run_world(case, world, mode)
    seen = np.zeros(len(dudt1), dtype=bool)
    for o in outs:
        el = o["elements"]
        assert not seen[el].any()
        seen[el] = True
        # same kernels, same arithmetic per element: bitwise the single-handle residual
        assert np.array_equal(o["dudt"], dudt1[el])
        assert np.array_equal(o["dudt_host"], dudt1[el])
        assert np.max(np.abs(o["dudt"] - ref[el])) <= 1e-12 * np.max(np.abs(ref))
        if state1 is not None:
            assert np.array_equal(o["state_ck54"], state1[el])
    assert seen.all()
    if case != "advdiff2d_p3_br1" and world == 2:
        # the flow under test needs an interior AND both boundary ranges on at least one rank
        assert any(0 < o["interior"][0] < o["interior"][1] < len(o["elements"]) for o in outs)


@pytest.mark.parametrize("case,world,mode", [("euler_tri_p4", 2, 1), ("adv_tet_p2", 2, 2)])
def test_library_sharded_flow_address_sanitizer(single_handle, case, world, mode):
    """The same run on the AddressSanitizer build of the emulated library: every "device" buffer
    (state, traces with their halo slots, send / receive buffers) is its own heap block, so an
    out-of-range halo slot, send index or element range of the sharded flow is trapped."""
    asan_rt = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True,
                             text=True).stdout.strip()
    if not os.path.isabs(asan_rt) or not os.path.exists(asan_rt):
        pytest.skip("libasan not available")
    dudt1, _, _ = single_handle(case)
    for o in _run_world(case, world, mode, asan_rt=asan_rt):
        assert np.array_equal(o["dudt"], dudt1[o["elements"]])
        assert np.array_equal(o["dudt_host"], dudt1[o["elements"]])
