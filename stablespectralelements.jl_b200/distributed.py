"""Element-sharded multi-GPU residual: one process per GPU, facet-trace halo over NCCL.

The reference has no distributed path at all (its only parallelism is ``Threads.@threads`` over
elements, /root/reference/src/Solvers/Solvers.jl:498-518).  The only inter-element data the
residual reads is the exterior trace ``u_f[CI[mapP[:, k]], :]``
(flux_differencing_form.jl:312-313, standard_form_first_order.jl:33-34), so sharding the
elements needs exactly one neighbour exchange of trace values between loop A and loop B:

    loop A (all local elements)  ->  pack boundary traces  ->  isend/irecv (NCCL, NVLink)
    loop B on interior elements (overlaps the transfer)    ->  unpack halo  ->  loop B on the
    boundary elements.

Second-order (BR1) equations read the neighbour's auxiliary-variable trace ``q_f`` as well
(standard_form_second_order.jl:63-64), so they exchange twice: u_f before ``auxiliary_variable!``
and q_f before ``time_derivative!`` (``DistributedResidual._flow``).

``partition`` is pure NumPy (tested on CPU with gloo, tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np


@dataclass
class Partition:
    rank: int
    world: int
    start: int                   # first global element owned
    stop: int                    # one past the last
    mapP_local: np.ndarray       # (N_f, N_loc) local/halo linear indices
    n_halo: int
    send_idx: np.ndarray         # local linear indices (j + N_f*k_loc) to pack, grouped by peer
    send_counts: Dict[int, int]  # peer -> number of trace nodes sent (ordered by peer rank)
    recv_counts: Dict[int, int]  # peer -> number of trace nodes received (halo slots, in order)
    interior: tuple              # (k_lo, k_hi): local elements that read no halo value

    @property
    def elements(self):
        return np.arange(self.start, self.stop)


def element_ranges(N_e: int, world: int):
    """Equal-count contiguous chunks of the lexicographic element ordering."""
    bounds = [(N_e * r) // world for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def partition(mapP: np.ndarray, rank: int, world: int) -> Partition:
    """Split the global connectivity ``mapP`` (N_f, N_e) for ``rank``.

    Halo slots are numbered peer by peer (ascending peer rank) and, within a peer, by the
    owner's global trace index; the owner packs its send list in the same order, so no index
    lists ever have to be communicated."""
    N_f, N_e = mapP.shape
    ranges = element_ranges(N_e, world)
    start, stop = ranges[rank]
    n_loc = stop - start
    starts = np.array([r[0] for r in ranges] + [N_e])
    tgt = mapP[:, start:stop]                                # global linear index j' + N_f*k'
    tk = tgt // N_f
    owner = np.searchsorted(starts, tk, side="right") - 1
    local = owner == rank
    mp = np.empty_like(tgt)
    mp[local] = tgt[local] - N_f * start
    send_idx: List[np.ndarray] = []
    send_counts: Dict[int, int] = {}
    recv_counts: Dict[int, int] = {}
    halo_off = 0
    gl = (np.arange(N_f)[:, None] + N_f * (start + np.arange(n_loc))[None, :])   # own global idx
    ll = (np.arange(N_f)[:, None] + N_f * np.arange(n_loc)[None, :])            # own local idx
    for peer in range(world):
        if peer == rank:
            continue
        sel = owner == peer
        cnt = int(sel.sum())
        if cnt == 0:
            continue
        # receive: order halo slots by the owner's global index
        want = tgt[sel]
        order = np.argsort(want, kind="stable")
        slot = np.empty(cnt, dtype=np.int64)
        slot[order] = halo_off + np.arange(cnt)
        mp[sel] = N_f * n_loc + slot
        recv_counts[peer] = cnt
        halo_off += cnt
        # send: my nodes whose partner lives on `peer`, ordered by my global index
        mine_g = gl[sel]
        o2 = np.argsort(mine_g, kind="stable")
        send_idx.append(ll[sel][o2])
        send_counts[peer] = cnt
    touches_halo = np.any(~local, axis=0)
    inner = np.nonzero(~touches_halo)[0]
    if len(inner) == 0:
        interior = (0, 0)
    else:
        # largest contiguous run of elements that read no halo value
        brk = np.nonzero(np.diff(inner) > 1)[0]
        seg_s = np.concatenate(([0], brk + 1))
        seg_e = np.concatenate((brk + 1, [len(inner)]))
        best = int(np.argmax(seg_e - seg_s))
        interior = (int(inner[seg_s[best]]), int(inner[seg_e[best] - 1]) + 1)
    return Partition(rank, world, start, stop, mp, halo_off,
                     np.concatenate(send_idx) if send_idx else np.zeros(0, dtype=np.int64),
                     send_counts, recv_counts, interior)


class _DevBuf:
    """Expose a raw device pointer through __cuda_array_interface__ (for torch.as_tensor)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 3, "strides": None}


class DistributedResidual:
    """Residual of one element shard on one GPU (world = 1: the whole mesh, no exchange)."""

    def __init__(self, solver, rank: int = 0, world: int = 1, device: int = 0,
                 backend: Optional[str] = None, nccl_id: Optional[bytes] = None):
        """``backend``: "library" (default at world > 1): partition, halo and the NCCL exchange
        live behind the C ABI (``sse_shard_create`` / ``sse_shard_residual``; the 128-byte NCCL id
        is taken from ``nccl_id`` or broadcast from rank 0 over torch.distributed);
        "python": the flow generator below drives the exchange with torch.distributed P2P (what
        the sharded-emulation tests run, with device copies as the transport)."""
        import os
        from .device import DeviceResidual
        self.solver, self.rank, self.world = solver, rank, world
        sd = solver.spatial_discretization
        self.dim = sd.reference_approximation.dim
        self.second_order = solver.law_desc["kind"] in ("advection_diffusion", "viscous_burgers")
        self.backend = backend or os.environ.get("SSE_B200_SHARD_BACKEND", "library")
        if world == 1:
            self.part = None
            self.backend = "single"
            self.elements = np.arange(sd.N_e)
            self.dev = DeviceResidual(solver, device=device)
        elif self.backend == "library":
            from .device import nccl_unique_id
            N_e_global = sd.mesh.mapP.shape[1]
            start, stop = element_ranges(N_e_global, world)[rank]
            self.elements = np.arange(start, stop)
            if nccl_id is None:
                import torch.distributed as dist
                box = [nccl_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(box, src=0)
                nccl_id = box[0]
            cols = sd.mesh.mapP[:, start:stop]
            local = sd.mesh.elem_start is not None
            if local and (sd.mesh.elem_start != start or sd.N_e != stop - start):
                raise ValueError("mesh shard does not match this rank's element range")
            self.dev = DeviceResidual(solver, device=device, mapP=cols,
                                      elements=None if local else self.elements,
                                      shard=(rank, world, nccl_id, N_e_global))
            pl = self.dev.shard_plan()
            self.part = Partition(rank, world, start, stop, None, int(pl.n_halo), None,
                                  {int(pl.peers[q]): int(pl.send_counts[q]) for q in range(pl.n_peers)},
                                  {int(pl.peers[q]): int(pl.recv_counts[q]) for q in range(pl.n_peers)},
                                  (int(pl.k_lo), int(pl.k_hi)))
        else:
            import torch
            self.torch = torch
            self.part = partition(sd.mesh.mapP, rank, world)
            self.elements = self.part.elements
            if sd.mesh.elem_start is not None:
                # the discretization already holds only this rank's shard (setup memory ~1/world)
                if sd.mesh.elem_start != self.part.start or sd.N_e != len(self.elements):
                    raise ValueError("mesh shard does not match this rank's element range")
                self.dev = DeviceResidual(solver, device=device, mapP=self.part.mapP_local,
                                          n_halo=self.part.n_halo)
            else:
                self.dev = DeviceResidual(solver, device=device, mapP=self.part.mapP_local,
                                          n_halo=self.part.n_halo, elements=self.elements)
            self.dev.halo_setup(self.part.send_idx)
            self._attach_buffers(device)
            self._ops = None
        self.local_shape = self.dev.shape
        self.n_local_state = int(np.prod(self.local_shape))

    def _attach_buffers(self, device: int):
        """Wrap the library's packed send / receive buffers as torch CUDA tensors (what NCCL
        sends from / receives into) and run the library on torch's current stream so that the
        NCCL ordering is stream-ordered."""
        torch = self.torch
        s_ptr, r_ptr, n_s, n_r = self.dev.halo_buffers()
        # doubles per trace node: N_c for u_f, dim * N_c for the BR1 auxiliary traces q_f
        width = self.dev.N_c * (self.dim if self.second_order else 1)
        dv = torch.device("cuda", device)
        self.send_t = torch.as_tensor(_DevBuf(s_ptr, n_s * width), device=dv)
        self.recv_t = torch.as_tensor(_DevBuf(r_ptr, n_r * width), device=dv)
        self.dev.set_stream(torch.cuda.current_stream(dv).cuda_stream)

    # ------------------------------------------------------------------ plumbing
    def halo_segments(self, width: int):
        """peer -> (send slice, recv slice) of the packed buffers, ``width`` doubles per node."""
        seg, so, ro = {}, 0, 0
        for peer in sorted(set(self.part.send_counts) | set(self.part.recv_counts)):
            ns = self.part.send_counts.get(peer, 0) * width
            nr = self.part.recv_counts.get(peer, 0) * width
            seg[peer] = (slice(so, so + ns), slice(ro, ro + nr))
            so += ns
            ro += nr
        return seg

    def _p2p_ops(self, width: Optional[int] = None):
        """isend/irecv pairs of one halo exchange, ``width`` doubles per trace node."""
        import torch.distributed as dist
        ops = []
        width = self.dev.N_c if width is None else width
        for peer, (snd, rcv) in self.halo_segments(width).items():
            if rcv.stop > rcv.start:
                ops.append(dist.P2POp(dist.irecv, self.recv_t[rcv], peer))
            if snd.stop > snd.start:
                ops.append(dist.P2POp(dist.isend, self.send_t[snd], peer))
        return ops

    def _flow(self, dudt_host=None):
        """Everything after loop A, as a generator: it yields the number of doubles per trace
        node each time a halo exchange of the packed send buffer has to be STARTED and is sent
        back a ``wait()`` callable, which it calls right before it needs the received values --
        the element ranges that read no halo value run in between and overlap the transfer.
        (The caller owns the transport: NCCL in production, device copies between two shards on
        one GPU in tests/test_gpu_sharded_emulation.py.)

        First order: loop B.  Second order (BR1; Solvers.jl:520-570): loop A2
        (``auxiliary_variable!``) needs the neighbour's u_f, loop B (``time_derivative!``) the
        neighbour's q_f -- two exchanges.  With ``dudt_host`` the result is copied back range by
        range while later ranges still compute (the interior is cut into a few pieces for that).
        """
        d = self.dev
        k_lo, k_hi = self.part.interior
        boundary = [(a, b) for a, b in ((0, k_lo), (max(k_hi, k_lo), d.N_e)) if b > a]
        npieces = 6 if (dudt_host is not None and k_hi - k_lo >= 6 * 4096) else 1
        cuts = [k_lo + ((k_hi - k_lo) * q) // npieces for q in range(npieces + 1)]
        interior = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

        def loop_b(ranges, fn):
            for a, b in ranges:
                fn(a, b)
                if dudt_host is not None:
                    d.download_dudt_range(dudt_host, a, b)

        d.halo_pack()
        wait = yield d.N_c
        if not self.second_order:
            loop_b(interior, d.time_derivative_range)          # overlaps the NVLink transfer
            wait()
            d.halo_unpack()
            loop_b(boundary, d.time_derivative_range)
            return
        if k_hi > k_lo:
            d.auxiliary_variable_range(k_lo, k_hi)
        wait()
        d.halo_unpack()
        for a, b in boundary:
            d.auxiliary_variable_range(a, b)
        d.halo_pack_aux()
        wait = yield d.N_c * self.dim
        loop_b(interior, d.time_derivative_only_range)
        wait()
        d.halo_unpack_aux()
        loop_b(boundary, d.time_derivative_only_range)

    def _neighbour_span(self):
        """Per local element: lowest / highest LOCAL element one of its facet nodes reads from
        (halo reads excluded) -- what has to have run loop A before loop B of the element may."""
        if getattr(self, "_span", None) is None:
            N_f, n_loc = self.part.mapP_local.shape
            kk = self.part.mapP_local // N_f
            own = np.arange(n_loc)[None, :]
            local = kk < n_loc
            lo = np.where(local, kk, own).min(axis=0)
            hi = np.where(local, kk, own).max(axis=0)
            self._span = (np.minimum(lo, own[0]), np.maximum(hi, own[0]))
        return self._span

    def _flow_host_interleaved(self, u_host, dudt_host, n_pieces: int = 12,
                               min_piece: int = 2048):
        """Host-buffer residual of a first-order equation with the upload interleaved with BOTH
        loops (opt-in, SSE_B200_SHARD_PIPELINE=1; same generator protocol as ``_flow``).

        The default host path uploads everything (overlapped with loop A only) before loop B
        starts, so it costs H2D + loop B.  Here the boundary elements go first -- their traces
        are packed and the halo exchange starts while the interior is still being uploaded --
        then the interior arrives piece by piece, and loop B of a piece is launched as soon as
        loop A has been queued for every element it reads a trace from; results stream back on
        the second copy stream.  All kernels still run on ONE stream in program order; the only
        asynchrony is the existing H2D-event -> loop A and loop B -> event -> D2H pattern."""
        d = self.dev
        N = d.N_e
        k_lo, k_hi = self.part.interior
        boundary = [(a, b) for a, b in ((0, k_lo), (max(k_hi, k_lo), N)) if b > a]
        n_pieces = max(1, min(n_pieces, (k_hi - k_lo) // min_piece)) if k_hi > k_lo else 0
        cuts = [k_lo + ((k_hi - k_lo) * q) // n_pieces for q in range(n_pieces + 1)] if n_pieces else []
        interior = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        span_lo, span_hi = self._neighbour_span()
        need = [(int(span_lo[a:b].min()), int(span_hi[a:b].max()) + 1) for a, b in interior]
        have = np.zeros(N, dtype=bool)                 # loop A queued for these elements

        first = True
        for a, b in boundary:
            d.upload_range_and_nodal_values(u_host, a, b, first=first)
            first = False
            have[a:b] = True
        d.halo_pack()
        wait = yield d.N_c
        done = [False] * len(interior)
        for i, (a, b) in enumerate(interior):
            d.upload_range_and_nodal_values(u_host, a, b, first=first)
            first = False
            have[a:b] = True
            for j, (c, e) in enumerate(interior[:i + 1]):
                if not done[j] and have[need[j][0]:need[j][1]].all():
                    d.time_derivative_range(c, e)
                    d.download_dudt_range(dudt_host, c, e)
                    done[j] = True
        for j, (c, e) in enumerate(interior):          # whatever is still open (normally none)
            if not done[j]:
                assert have[need[j][0]:need[j][1]].all()
                d.time_derivative_range(c, e)
                d.download_dudt_range(dudt_host, c, e)
        wait()
        d.halo_unpack()
        for a, b in boundary:
            d.time_derivative_range(a, b)
            d.download_dudt_range(dudt_host, a, b)

    def _exchange_and_time_derivative(self, dudt_host=None):
        """Halo exchange(s) over NCCL + the loops that follow loop A."""
        self._drive(self._flow(dudt_host))

    def _drive(self, flow):
        """Run a flow generator with NCCL as the transport of its halo exchanges."""
        import torch.distributed as dist
        try:
            width = next(flow)
            while True:
                works = dist.batch_isend_irecv(self._p2p_ops(width))
                width = flow.send(lambda works=works: [w.wait() for w in works])
        except StopIteration:
            pass

    def residual(self):
        """One residual of the device-resident state (dudt stays on the device)."""
        if self.backend == "library":
            self.dev.shard_residual()
        elif self.world == 1:
            self.dev.nodal_values()
            self.dev.time_derivative()
        else:
            self.dev.nodal_values()
            self._exchange_and_time_derivative()

    def residual_host(self, u: np.ndarray, dudt: np.ndarray):
        """Public-API path with host buffers for this shard (H2D + residual + D2H)."""
        if self.world == 1:
            self.dev.residual_host(u, dudt)
            return
        if self.backend == "library":
            self.dev.shard_residual(u, dudt)
            return
        import os
        if os.environ.get("SSE_B200_SHARD_PIPELINE") == "1" and not self.second_order:
            # opt-in: upload interleaved with both loops (see _flow_host_interleaved)
            if not getattr(self, "_split_streams", False):
                self.dev.set_copy_streams(True)
                self._split_streams = True
            self._drive(self._flow_host_interleaved(u, dudt))
        else:
            self.dev.upload_and_nodal_values(u)          # chunked H2D overlapped with loop A
            self._exchange_and_time_derivative(dudt)     # loop B overlapped with the D2H copies
        self.dev.sync_copies()

    def host_flow_name(self) -> str:
        """Which host-buffer flow residual_host runs (reported by bench.py)."""
        import os
        if self.world == 1:
            return "sse_residual(where=HOST): chunked H2D / loop A / loop B / D2H pipeline"
        if self.backend == "library":
            return ("sse_shard_residual(where=HOST): chunked H2D + loop A, NCCL exchange || interior "
                    "loop B, boundary loop B, D2H by ranges")
        if os.environ.get("SSE_B200_SHARD_PIPELINE") == "1" and not self.second_order:
            return "interleaved: boundary upload -> exchange || interior upload + loops A, B + D2H"
        return "upload + loop A, exchange || interior loop B, boundary loop B, D2H by ranges"

    def timed_residuals(self, steps: int) -> float:
        """Milliseconds for ``steps`` residuals, CUDA events on the launching stream."""
        if self.world == 1:
            return self.dev.time_residual(steps, split=False)[0]
        if self.backend == "library":
            return self.dev.shard_time_residual(steps)
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.residual()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    def split_times(self, reps: int):
        _, ta, tb = self.dev.time_residual(reps, split=True)
        return {"loop_a_ms": ta / reps, "loop_b_ms": tb / reps, "note":
                "per-kernel CUDA-event times from a separate pass (halo exchange excluded)"}

    def set_state(self, u_local: np.ndarray):
        self.dev.set_state(u_local)

    def get_dudt(self) -> np.ndarray:
        out = np.empty(self.local_shape)
        self.dev.download_dudt(out)
        return out

    def kernel_launches(self) -> int:
        return self.dev.kernel_launches()

    def sync(self):
        if self.backend == "library":
            self.dev.shard_sync()
        else:
            self.dev.sync()

    def close(self):
        self.dev.close()
