"""Non-default paths of the library on real hardware (each also runs on the host emulator).

* SSE_B200_SPLIT_B=1: loop B as k_fluxdiff_volume + k_fluxdiff_facet -- must reproduce the fused
  kernel to round-off (same per-node arithmetic, the nodal residual merely travels through
  global memory).
* DistributedResidual._flow_host_interleaved: host-buffer residual of a shard with the upload
  interleaved with both loops; exercised here with two shards on one GPU (device copies as the
  halo transport), which is what checks its stream / event dependencies on real hardware."""
import math
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", [4, 3])
def test_split_loop_b_matches_the_fused_kernel(p, monkeypatch):
    from sse_b200 import device as dev
    solver, u0 = cases.euler_tet_case(p=p, M=4, lazy=True, warp=True, ic="periodic")
    u = cases.rough_state(solver, u0, seed=3)
    outs = {}
    for split in ("0", "1"):
        monkeypatch.setenv("SSE_B200_SPLIT_B", split)
        d = dev.DeviceResidual(solver)
        try:
            dudt = np.full_like(u, np.nan)
            for _ in range(2):
                d.residual_host(u, dudt)
            outs[split] = dudt
            launches = d.kernel_launches()
        finally:
            d.close()
    assert np.all(np.isfinite(outs["1"]))
    # same per-node arithmetic in the same order; nvcc may still contract multiply-adds
    # differently in the two instantiations, hence round-off rather than bitwise here (the
    # emulator build, which contracts nothing, is bitwise: test_kernels_host_emulation.py)
    assert np.max(np.abs(outs["0"] - outs["1"])) <= 1e-13 * np.max(np.abs(outs["0"]))


def _long_mesh_case(n_layers):
    from sse_b200.conservation_laws import EulerEquations, LaxFriedrichsNumericalFlux
    from sse_b200.geometric_factors import ChanWilcoxMetrics, make_spatial_discretization
    from sse_b200.grid_functions import EulerPeriodicTest
    from sse_b200.mesh import ChanWarping, uniform_periodic_mesh, warp_mesh
    from sse_b200.reference_approximation import ModalTensor, Tet, make_reference_approximation
    from sse_b200.solvers import (FluxDifferencingForm, ReferenceOperator, Solver,
                                  project_function)
    L = 2 * math.pi
    ra = make_reference_approximation(ModalTensor(3), Tet(), mapping_degree=2)
    mesh = warp_mesh(uniform_periodic_mesh(ra, ((0.0, L),) * 3, (6, 6, n_layers)), ra,
                     ChanWarping(1 / 16, (L, L, L)))
    sd = make_spatial_discretization(mesh, ra, ChanWilcoxMetrics())
    solver = Solver(EulerEquations(3, 1.4), sd,
                    FluxDifferencingForm(inviscid_numerical_flux=LaxFriedrichsNumericalFlux()),
                    ReferenceOperator(), lazy=True)
    return solver, project_function(EulerPeriodicTest(3, 1.4, 0.2, L), sd)


def test_interleaved_host_flow_two_shards_on_one_gpu():
    import torch
    from sse_b200 import device as dev
    from sse_b200.distributed import DistributedResidual
    from test_gpu_sharded_emulation import _local_exchange
    solver, u0 = _long_mesh_case(24)                      # 5184 elements, 2592 per shard
    u = cases.rough_state(solver, u0, seed=6)
    whole = dev.DeviceResidual(solver)
    try:
        ref = np.empty_like(u)
        whole.residual_host(u, ref)
    finally:
        whole.close()
    shards = [DistributedResidual(solver, rank=r, world=2, device=0, backend="python")
              for r in range(2)]
    try:
        for sh in shards:
            sh.dev.set_copy_streams(True)
        pin = lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()
        us = [pin(np.ascontiguousarray(u[sh.elements])) for sh in shards]
        outs = [pin(np.full_like(x, np.nan)) for x in us]
        for rep in range(3):                              # back-to-back calls reuse the buffers
            for o in outs:
                o[...] = np.nan
            flows = [sh._flow_host_interleaved(x, o, n_pieces=6, min_piece=200)
                     for sh, x, o in zip(shards, us, outs)]
            widths = [next(f) for f in flows]
            _local_exchange(shards, widths[0])
            for f in flows:
                with pytest.raises(StopIteration):
                    f.send(lambda: None)
            for sh in shards:
                sh.dev.sync_copies()
            assert np.array_equal(np.concatenate(outs, axis=0), ref), rep
    finally:
        for sh in shards:
            sh.close()


@pytest.mark.parametrize("staged", ["0", "1"])
@pytest.mark.parametrize("case", ["advdiff2d_p3", "advdiff2d_p4", "advdiff1d_p4", "adv2d_physical", "burgers2d_physical"])
def test_both_forms_of_the_physical_operator_kernel(case, staged, monkeypatch):
    """k_physical with the per-element operators staged in shared memory by cp.async.bulk
    (mbarrier completion; cp.async fallback for misaligned blocks) and streamed from global memory
    by a half-warp per row: both against the oracle at 1e-12, whatever sse_create would choose."""
    import sse_oracle as oc
    from bridge import oracle_problem
    from sse_b200 import device as dev
    monkeypatch.setenv("SSE_B200_PHYS_STAGED", staged)
    build = {"advdiff2d_p3": lambda: cases.advection_diffusion_case(d=2, p=3, M=3, lazy=True),
             "advdiff2d_p4": lambda: cases.advection_diffusion_case(d=2, p=4, M=5, lazy=True),
             "advdiff1d_p4": lambda: cases.advection_diffusion_case(d=1, p=4, M=7, lazy=True),
             "adv2d_physical": lambda: cases.advection_physical_case(d=2, p=3, M=3, lazy=True),
             "burgers2d_physical": lambda: cases.burgers_physical_case(p=3, M=3, lazy=True)}[case]
    solver, u0 = build()
    u = cases.rough_state(solver, u0, seed=2)
    d = dev.DeviceResidual(solver)
    try:
        dudt = np.full_like(u, np.nan)
        d.residual_host(u, dudt)
    finally:
        d.close()
    ref = oc.semi_discrete_residual(oracle_problem(solver), u)
    assert np.all(np.isfinite(dudt))
    assert np.max(np.abs(dudt - ref)) < 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("env", [{"SSE_B200_HOST_CHUNKS": "9"}, {"SSE_B200_HOST_CHUNKS": "9", "SSE_B200_HOST_TAPER": "0"},
                                 {"SSE_B200_HOST_CHUNKS": "32", "SSE_B200_HOST_ASTREAM": "0"},
                                 {"SSE_B200_HOST_CHUNKS": "5", "SSE_B200_HOST_TRACE": "1"}])
def test_host_buffer_pipeline_knobs_are_bitwise_neutral(env, monkeypatch):
    """sse_residual(where=HOST) cut into tapered / equal chunks, loop A on its own stream or not:
    bitwise equal to the single-chunk residual of the same mesh (real streams and events here)."""
    from sse_b200 import device as dev
    solver, u0 = cases.euler_tet_case(p=4, M=4, lazy=True, warp=True, ic="periodic")
    u = cases.rough_state(solver, u0, seed=4)
    outs = []
    for e in (env, {"SSE_B200_HOST_CHUNKS": "1"}):
        for k in ("SSE_B200_HOST_CHUNKS", "SSE_B200_HOST_TAPER", "SSE_B200_HOST_ASTREAM", "SSE_B200_HOST_TRACE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in e.items():
            monkeypatch.setenv(k, v)
        d = dev.DeviceResidual(solver)
        try:
            dudt = np.full_like(u, np.nan)
            for _ in range(3):          # repeated calls reuse the chunk events
                d.residual_host(u, dudt)
            outs.append(dudt)
        finally:
            d.close()
    assert np.all(np.isfinite(outs[0]))
    assert np.array_equal(outs[0], outs[1])


def test_warp_private_projection_matches_the_batched_kernel(monkeypatch):
    """SSE_B200_PROJ_WARP=1: k_project_tet_w (one element per warp, __syncwarp between the stages)
    against the CTA-level k_project_tet: the same work items with the same arithmetic."""
    from sse_b200 import device as dev
    solver, u0 = cases.euler_tet_case(p=4, M=4, lazy=True, warp=True, ic="periodic")
    u = cases.rough_state(solver, u0, seed=7)
    outs = []
    for pw in ("0", "1"):
        monkeypatch.setenv("SSE_B200_PROJ_WARP", pw)
        d = dev.DeviceResidual(solver)
        try:
            dudt = np.full_like(u, np.nan)
            d.residual_host(u, dudt)
            outs.append(dudt)
        finally:
            d.close()
    assert np.all(np.isfinite(outs[1]))
    assert np.max(np.abs(outs[0] - outs[1])) <= 1e-13 * np.max(np.abs(outs[0]))


@pytest.mark.parametrize("case", ["tet_p4", "tet_p3", "tri_p4"])
def test_loop_a_projection_instantiation_matches_the_run_time_mode(case, monkeypatch):
    """Loop A of the modal Euler schemes runs k_nodal_tensor<..., PROJ_CT = 2> (the entropy
    projection compiled in, the dense-row forms of R left out); SSE_B200_NODAL_RT_PROJ=1 keeps the
    kernel that takes the mode at run time.  Same stages, same arithmetic: equal up to the FMA
    contraction choices of two instantiations (the emulator, which has none, sees them bitwise)."""
    from sse_b200 import device as dev
    if case == "tri_p4":
        solver, u0 = cases.euler_tri_case(p=4, M=6, lazy=True)
    else:
        solver, u0 = cases.euler_tet_case(p=int(case[-1]), M=3, lazy=True, warp=True, ic="periodic")
    u = cases.rough_state(solver, u0, seed=11)
    outs = []
    for rt in ("0", "1"):
        monkeypatch.setenv("SSE_B200_NODAL_RT_PROJ", rt)
        d = dev.DeviceResidual(solver)
        try:
            dudt = np.full_like(u, np.nan)
            d.residual_host(u, dudt)
            outs.append(dudt)
        finally:
            d.close()
    assert np.all(np.isfinite(outs[0]))
    assert np.max(np.abs(outs[0] - outs[1])) <= 1e-13 * np.max(np.abs(outs[0]))
    return outs
