"""Element-sharded residual on ONE GPU: W shards (W handles on cuda:0) run the production
control flow of ``DistributedResidual._flow`` -- pack, exchange, interior ranges, unpack,
boundary ranges; twice for second-order (BR1) equations -- with device copies between the
shards' send/recv buffers standing in for the NCCL transport.  Each shard's result must be
bitwise the single-handle residual of its elements (same kernels, same arithmetic) and match the
oracle to 1e-12.  This is what covers sse_halo_pack/unpack(_aux), sse_auxiliary_variable_range
and sse_time_derivative_only_range when only one GPU is available; the NCCL path itself is the
same flow driven by ``_exchange_and_time_derivative`` (tools/dist_check.py, bench.py --gpus N)."""
import numpy as np
import pytest

import cases
import sse_oracle as oc
from bridge import oracle_problem

pytestmark = pytest.mark.gpu


def _local_exchange(shards, width):
    """send buffer segments -> the peers' recv buffer segments (same order as _p2p_ops)."""
    for dst in shards:
        for peer, (_, rcv) in dst.halo_segments(width).items():
            if rcv.stop == rcv.start:
                continue
            snd = shards[peer].halo_segments(width)[dst.rank][0]
            assert snd.stop - snd.start == rcv.stop - rcv.start
            dst.recv_t[rcv] = shards[peer].send_t[snd]


def _run_sharded(solver, u, world, shard_cls=None):
    from sse_b200.distributed import DistributedResidual
    shard_cls = shard_cls or DistributedResidual
    shards = [shard_cls(solver, rank=r, world=world, device=0, backend="python")
              for r in range(world)]
    try:
        for sh in shards:
            sh.set_state(np.ascontiguousarray(u[sh.elements]))
            sh.dev.nodal_values()
        flows = [sh._flow() for sh in shards]
        widths = [next(f) for f in flows]
        n_exchanges = 0
        while widths:
            assert len(set(widths)) == 1 and len(widths) == world
            _local_exchange(shards, widths[0])          # every shard has packed by now
            n_exchanges += 1
            nxt = []
            for f in flows:
                try:
                    nxt.append(f.send(lambda: None))
                except StopIteration:
                    pass
            widths = nxt
        out = [sh.get_dudt() for sh in shards]
        return np.concatenate(out, axis=0), n_exchanges
    finally:
        for sh in shards:
            sh.close()


CASES = {
    "euler3d_tet_p2": (lambda: cases.euler_tet_case(p=2, M=2, lazy=True, warp=True,
                                                    ic="periodic"), 1),
    "euler2d_tri_p3": (lambda: cases.euler_tri_case(p=3, M=4, lazy=True), 1),
    "adv3d_tet_p2": (lambda: cases.advection_tet_case(p=2, M=2, lazy=True), 1),
    "advdiff1d_p4_br1": (lambda: cases.advection_diffusion_case(d=1, p=4, M=7, lazy=True), 2),
    "advdiff2d_p3_br1": (lambda: cases.advection_diffusion_case(d=2, p=3, M=4, lazy=True), 2),
}


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", sorted(CASES))
def test_sharded_flow_matches_single_domain(name, world, shard_cls=None):
    from sse_b200.distributed import DistributedResidual
    build, n_exch = CASES[name]
    solver, u0 = build()
    u = cases.rough_state(solver, u0, seed=5)
    whole = DistributedResidual(solver, rank=0, world=1, device=0)
    try:
        whole.set_state(u)
        whole.residual()
        ref_gpu = whole.get_dudt()
    finally:
        whole.close()
    got, n = _run_sharded(solver, u, world, shard_cls)
    assert n == n_exch
    assert np.array_equal(got, ref_gpu)
    ref = oc.semi_discrete_residual(oracle_problem(solver), u)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12
