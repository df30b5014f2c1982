"""N > 1 path on CPU: world_size-2 (and 3) gloo processes exercise the element partition and the
facet-trace halo exchange of sse_b200.distributed with the oracle as the local compute, and
must reproduce the single-domain residual exactly (same arithmetic, different ownership)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import sse_oracle as oc
from bridge import oracle_problem
from sse_b200.distributed import element_ranges, partition


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _restrict(prob, part):
    sl = slice(part.start, part.stop)
    loc = dict(prob)
    loc.pop("_SC", None)
    loc.pop("_PHYS", None)
    for key in ("J_q", "Lambda_q", "J_f", "nJf"):
        loc[key] = prob[key][sl]
    loc["N_e"] = part.stop - part.start
    loc["mapP"] = part.mapP_local
    return loc


def _worker(rank, world, port, builder, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        solver, u0 = getattr(cases, builder[0])(**builder[1])
        u = cases.rough_state(solver, u0, seed=7)
        prob = oracle_problem(solver)
        part = partition(prob["mapP"], rank, world)
        loc = _restrict(prob, part)
        N_f, N_c = prob["N_f"], prob["N_c"]
        # loop A on the local elements
        u_q, u_f = oc.nodal_values_fluxdiff(loc, u[part.start:part.stop])
        flat = u_f.reshape(-1, N_c)                          # index j + N_f * k_local
        send = torch.from_numpy(np.ascontiguousarray(flat[part.send_idx]))
        recv = torch.empty((part.n_halo, N_c), dtype=torch.float64)
        ops, so, ro = [], 0, 0
        for peer in sorted(set(part.send_counts) | set(part.recv_counts)):
            ns, nr = part.send_counts.get(peer, 0), part.recv_counts.get(peer, 0)
            if nr:
                ops.append(dist.P2POp(dist.irecv, recv[ro:ro + nr], peer))
            if ns:
                ops.append(dist.P2POp(dist.isend, send[so:so + ns], peer))
            so += ns
            ro += nr
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        ext = np.concatenate([flat, recv.numpy()], axis=0)
        u_out = ext[part.mapP_local.T.reshape(-1)].reshape(u_f.shape)
        dudt = oc.fluxdiff_loop_b(loc, u_q, u_f, u_out)
        ref = oc.semi_discrete_residual(prob, u)[part.start:part.stop]
        err = float(np.max(np.abs(dudt - ref)))
        k_lo, k_hi = part.interior
        interior_ok = bool(np.all(part.mapP_local[:, k_lo:k_hi] < N_f * loc["N_e"]))
        res = torch.tensor([err, float(interior_ok), float(k_hi - k_lo)], dtype=torch.float64)
        gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, res)
        if rank == 0:
            out.put([g.tolist() for g in gathered])
    finally:
        dist.destroy_process_group()


def _exchange(part, send_rows, width):
    """One halo exchange over gloo: rows of ``send_rows`` (n_send, width) go to the peers in the
    partition's order; returns the (n_halo, width) halo rows."""
    send = torch.from_numpy(np.ascontiguousarray(send_rows))
    recv = torch.empty((part.n_halo, width), dtype=torch.float64)
    ops, so, ro = [], 0, 0
    for peer in sorted(set(part.send_counts) | set(part.recv_counts)):
        ns, nr = part.send_counts.get(peer, 0), part.recv_counts.get(peer, 0)
        if nr:
            ops.append(dist.P2POp(dist.irecv, recv[ro:ro + nr], peer))
        if ns:
            ops.append(dist.P2POp(dist.isend, send[so:so + ns], peer))
        so += ns
        ro += nr
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return recv.numpy()


def _worker_second_order(rank, world, port, builder, out):
    """BR1: two exchanges -- u_f before auxiliary_variable!, q_f before time_derivative!."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        solver, u0 = getattr(cases, builder[0])(**builder[1])
        u = cases.rough_state(solver, u0, seed=7)
        prob = oracle_problem(solver)
        part = partition(prob["mapP"], rank, world)
        loc = _restrict(prob, part)
        N_f, N_c, d = prob["N_f"], prob["N_c"], prob["d"]
        gather = part.mapP_local.T.reshape(-1)
        u_q = np.einsum("qp,kep->kqe", prob["V"], u[part.start:part.stop])
        u_f = np.einsum("fq,kqe->kfe", prob["R"], u_q)
        flat = u_f.reshape(-1, N_c)
        halo = _exchange(part, flat[part.send_idx], N_c)
        u_out = np.concatenate([flat, halo], axis=0)[gather].reshape(u_f.shape)
        q_q, q_f = oc.second_order_auxiliary_variable(loc, u_q, u_f, u_out)
        flat_q = q_f.reshape(-1, N_c * d)
        halo_q = _exchange(part, flat_q[part.send_idx], N_c * d)
        q_out = np.concatenate([flat_q, halo_q], axis=0)[gather].reshape(q_f.shape)
        dudt = oc.second_order_time_derivative(loc, u_q, u_f, u_out, q_q, q_f, q_out)
        ref = oc.semi_discrete_residual(prob, u)[part.start:part.stop]
        res = torch.tensor([float(np.max(np.abs(dudt - ref))), 1.0, 0.0], dtype=torch.float64)
        gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, res)
        if rank == 0:
            out.put([g.tolist() for g in gathered])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,builder", [
    (2, ("advection_diffusion_case", dict(d=1, p=4, M=6))),
    (3, ("advection_diffusion_case", dict(d=2, p=3, M=4))),
])
def test_partitioned_second_order_residual_matches_global(world, builder):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_second_order, args=(r, world, port, builder, out))
             for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for err, _, _ in res:
        assert err < 1e-13, err


@pytest.mark.parametrize("world,builder", [
    (2, ("euler_tri_case", dict(p=3, M=4))),
    (2, ("euler_tet_case", dict(p=2, M=2, warp=True))),
    (3, ("euler_tri_case", dict(p=2, M=5, interface="ec"))),
])
def test_partitioned_residual_matches_global(world, builder):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, builder, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for err, interior_ok, n_int in res:
        assert err < 1e-13, err          # identical arithmetic: only ownership differs
        assert interior_ok == 1.0


def test_partition_properties():
    solver, _ = cases.euler_tet_case(p=2, M=3, lazy=True)
    mapP = solver.spatial_discretization.mesh.mapP
    N_f, N_e = mapP.shape
    world = 4
    ranges = element_ranges(N_e, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == N_e
    assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1
    parts = [partition(mapP, r, world) for r in range(world)]
    for p in parts:
        # what I send to a peer is what that peer expects from me
        for peer, cnt in p.send_counts.items():
            assert parts[peer].recv_counts[p.rank] == cnt
        assert p.n_halo == sum(p.recv_counts.values())
        assert p.mapP_local.max() < N_f * (p.stop - p.start) + p.n_halo
        halo = p.mapP_local[p.mapP_local >= N_f * (p.stop - p.start)]
        assert len(np.unique(halo)) == len(halo) == p.n_halo


@pytest.mark.parametrize("builder", [("euler_tet_case", dict(p=2, M=3, lazy=True)),
                                     ("advection_diffusion_case", dict(d=1, p=2, M=7, lazy=True)),
                                     ("euler_tri_case", dict(p=2, M=4, lazy=True))])
@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_library_partition_matches_the_python_partitioner(builder, world):
    """sse_shard_plan_build (C ABI, pure host arithmetic -- what a Julia rank calls through
    sse_shard_create) against the NumPy ``partition`` every sharded test above is built on: same
    ranges, halo numbering, send lists, peer counts and interior range."""
    from sse_b200 import device as dev
    import __graft_entry__ as ge
    ge.build()
    solver, _ = getattr(cases, builder[0])(**builder[1])
    mapP = solver.spatial_discretization.mesh.mapP
    N_f, N_e = mapP.shape
    for rank in range(world):
        part = partition(mapP, rank, world)
        plan, mp_local, send = dev.shard_plan(mapP[:, part.start:part.stop], N_e, rank, world)
        assert (plan.start, plan.stop) == (part.start, part.stop)
        assert plan.n_halo == part.n_halo and plan.n_send == len(part.send_idx)
        assert np.array_equal(mp_local, part.mapP_local)
        assert np.array_equal(send, part.send_idx)
        peers = [int(plan.peers[q]) for q in range(plan.n_peers)]
        assert peers == sorted(part.send_counts)
        assert [int(plan.send_counts[q]) for q in range(plan.n_peers)] == \
            [part.send_counts[p] for p in peers]
        assert [int(plan.recv_counts[q]) for q in range(plan.n_peers)] == \
            [part.recv_counts[p] for p in peers]
        assert (plan.k_lo, plan.k_hi) == tuple(part.interior)
