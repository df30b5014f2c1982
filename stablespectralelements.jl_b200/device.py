"""ctypes binding of ``libsse_b200.so`` (include/sse_b200.h) and the packer that turns a host
``Solver`` into the ``sse_config`` / ``sse_operators`` / ``sse_geometry`` structs.

This is the Python counterpart of the Julia ``ccall`` shim shown in INTEGRATION.md.  The
library must be present (built by ``__graft_entry__.build()``); there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .linear_maps import (IdentityMap, SelectionMap, WarpedTensorProductMap2D,
                          WarpedTensorProductMap3D)
from .reference_approximation import NoMapping

_LIB = None
LIB_NAME = "libsse_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

LAW = dict(advection=0, burgers=1, euler=2, advection_diffusion=3, viscous_burgers=4)
INVISCID = dict(lf=0, central=1, ec=2)
MASS = dict(diagonal=0, weight_adjusted=1, cholesky=2)

c_d_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)


class SseConfig(C.Structure):
    _fields_ = [("dim", C.c_int32), ("N_p", C.c_int32), ("N_q", C.c_int32), ("N_f", C.c_int32),
                ("N_c", C.c_int32), ("num_faces", C.c_int32), ("N_e", C.c_int64),
                ("N_halo", C.c_int64), ("law", C.c_int32), ("a", C.c_double * 3),
                ("b", C.c_double), ("gamma", C.c_double), ("form", C.c_int32),
                ("strategy", C.c_int32), ("inviscid_flux", C.c_int32),
                ("half_lambda", C.c_double), ("two_point_flux", C.c_int32),
                ("v_kind", C.c_int32), ("r_is_selection", C.c_int32),
                ("mass_solver", C.c_int32), ("device", C.c_int32)]


class SseOperators(C.Structure):
    _fields_ = [("V_dense", c_d_p), ("n1d", C.c_int32), ("warp_A", c_d_p), ("warp_B", c_d_p),
                ("warp_C", c_d_p), ("sigma_i", c_i32_p),
                ("R_rowptr", c_i32_p), ("R_col", c_i32_p), ("R_val", c_d_p),
                ("D_rowptr", c_i32_p * 3), ("D_col", c_i32_p * 3), ("D_val", c_d_p * 3),
                ("W", c_d_p), ("B", c_d_p), ("Lambda_ref", c_d_p), ("J_ref", c_d_p),
                ("n_ref", c_d_p), ("Minv", c_d_p)]


class SseGeometry(C.Structure):
    _fields_ = [("J_q", c_d_p), ("Lambda_q", c_d_p), ("J_f", c_d_p), ("nJf", c_d_p),
                ("VOL", c_d_p), ("FAC", c_d_p), ("Minv_elem", c_d_p)]


class SseMapping(C.Structure):
    """sse_mapping (include/sse_b200.h): inputs of the on-device GeometricFactors."""
    _fields_ = [("dim", C.c_int32), ("metric", C.c_int32), ("N_map", C.c_int32),
                ("N_map1", C.c_int32), ("N_q", C.c_int32), ("N_f", C.c_int32),
                ("N_e", C.c_int64), ("D", c_d_p * 3), ("Vq", c_d_p), ("Vf", c_d_p), ("P", c_d_p),
                ("D1", c_d_p * 3), ("Vq1", c_d_p), ("Vf1", c_d_p), ("nrstJ", c_d_p),
                ("Jproj", c_d_p), ("xyz", c_d_p * 3), ("device", C.c_int32)]


EXPORTS = ["sse_last_error", "sse_version", "sse_create", "sse_destroy", "sse_residual",
           "sse_nodal_values", "sse_time_derivative", "sse_set_state", "sse_get_state",
           "sse_state_ptr", "sse_rk_stage", "sse_rk_step_ck54", "sse_halo_setup",
           "sse_halo_buffers", "sse_halo_pack", "sse_halo_unpack", "sse_sync", "sse_stream",
           "sse_time_residual", "sse_kernel_launches", "sse_device_bytes",
           "sse_measure_fp64_peak", "sse_time_derivative_range", "sse_set_stream",
           "sse_upload_state", "sse_download_dudt", "sse_upload_and_nodal_values",
           "sse_download_dudt_range", "sse_sync_copies", "sse_functional",
           "sse_geometry_build", "sse_geometry_free", "sse_copy_to_host",
           "sse_auxiliary_variable_range", "sse_time_derivative_only_range",
           "sse_halo_pack_aux", "sse_halo_unpack_aux", "sse_erk_step",
           "sse_upload_range_and_nodal_values", "sse_set_copy_streams",
           "sse_measure_dmma_peak", "sse_shard_range", "sse_shard_plan_build",
           "sse_nccl_unique_id", "sse_shard_create", "sse_shard_destroy", "sse_shard_handle",
           "sse_shard_get_plan", "sse_shard_residual", "sse_shard_rk_stage",
           "sse_shard_rk_step_ck54", "sse_shard_time_residual", "sse_shard_sync",
           "sse_probe_elementary"]


SSE_MAX_PEERS = 64


class SseShardPlan(C.Structure):
    """sse_shard_plan of include/sse_b200.h."""
    _fields_ = [("start", C.c_int64), ("stop", C.c_int64), ("n_halo", C.c_int64),
                ("n_send", C.c_int64), ("k_lo", C.c_int64), ("k_hi", C.c_int64),
                ("n_peers", C.c_int32), ("peers", C.c_int32 * SSE_MAX_PEERS),
                ("send_counts", C.c_int64 * SSE_MAX_PEERS),
                ("recv_counts", C.c_int64 * SSE_MAX_PEERS)]


def load_library(path: Optional[str] = None, allow_emulation: bool = False):
    """Load libsse_b200.so; raises if it has not been built (no fallback).

    A host-emulation build of the kernels (tests/emu, ``sse_version() < 0``) is test
    infrastructure: it is refused unless the caller -- a test -- asks for it explicitly."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or os.environ.get("SSE_B200_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: run `python -c 'import __graft_entry__ as g; "
                           f"g.build()'` first -- the residual has no CPU fallback")
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.sse_last_error.restype = C.c_char_p
    lib.sse_version.restype = C.c_int
    if lib.sse_version() < 0 and not allow_emulation:
        raise RuntimeError(f"{p} is a host-emulation test build of the kernels, not the CUDA "
                           f"library: the residual has no CPU fallback")
    lib.sse_create.argtypes = [C.POINTER(SseConfig), C.POINTER(SseOperators),
                               C.POINTER(SseGeometry), c_i64_p, C.POINTER(vp)]
    lib.sse_destroy.argtypes = [vp]
    lib.sse_residual.argtypes = [vp, vp, vp, C.c_double, C.c_int]
    lib.sse_nodal_values.argtypes = [vp, vp]
    lib.sse_time_derivative.argtypes = [vp, vp]
    lib.sse_set_state.argtypes = [vp, vp]
    lib.sse_get_state.argtypes = [vp, vp]
    lib.sse_state_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.sse_rk_stage.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    lib.sse_rk_step_ck54.argtypes = [vp, C.c_double]
    lib.sse_halo_setup.argtypes = [vp, c_i64_p, C.c_int64]
    lib.sse_halo_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), c_i64_p, c_i64_p]
    lib.sse_halo_pack.argtypes = [vp]
    lib.sse_halo_unpack.argtypes = [vp]
    lib.sse_sync.argtypes = [vp]
    lib.sse_stream.argtypes = [vp]
    lib.sse_stream.restype = vp
    lib.sse_time_residual.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    lib.sse_kernel_launches.argtypes = [vp]
    lib.sse_kernel_launches.restype = C.c_int64
    lib.sse_device_bytes.argtypes = [vp]
    lib.sse_device_bytes.restype = C.c_int64
    lib.sse_measure_fp64_peak.argtypes = [C.c_int, c_d_p]
    lib.sse_measure_dmma_peak.argtypes = [C.c_int, c_d_p]
    lib.sse_probe_elementary.argtypes = [C.c_int, C.c_int, c_d_p, c_d_p, C.c_int64]
    lib.sse_time_derivative_range.argtypes = [vp, vp, C.c_int64, C.c_int64]
    lib.sse_set_stream.argtypes = [vp, vp]
    lib.sse_upload_state.argtypes = [vp, vp]
    lib.sse_download_dudt.argtypes = [vp, vp]
    lib.sse_upload_and_nodal_values.argtypes = [vp, vp]
    lib.sse_download_dudt_range.argtypes = [vp, vp, C.c_int64, C.c_int64]
    lib.sse_sync_copies.argtypes = [vp]
    lib.sse_functional.argtypes = [vp, C.c_int, C.c_int, vp, c_d_p]
    lib.sse_geometry_build.argtypes = [C.POINTER(SseMapping), C.POINTER(SseGeometry)]
    lib.sse_geometry_free.argtypes = [C.POINTER(SseGeometry)]
    lib.sse_copy_to_host.argtypes = [vp, vp, C.c_int64]
    lib.sse_auxiliary_variable_range.argtypes = [vp, C.c_int64, C.c_int64]
    lib.sse_time_derivative_only_range.argtypes = [vp, vp, C.c_int64, C.c_int64]
    lib.sse_halo_pack_aux.argtypes = [vp]
    lib.sse_halo_unpack_aux.argtypes = [vp]
    lib.sse_erk_step.argtypes = [vp, C.c_int, c_d_p, c_d_p, C.c_double]
    lib.sse_upload_range_and_nodal_values.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int]
    lib.sse_set_copy_streams.argtypes = [vp, C.c_int]
    lib.sse_shard_range.argtypes = [C.c_int64, C.c_int, C.c_int, c_i64_p, c_i64_p]
    lib.sse_shard_plan_build.argtypes = [c_i64_p, C.c_int32, C.c_int64, C.c_int, C.c_int,
                                         C.POINTER(SseShardPlan), c_i64_p, c_i64_p]
    lib.sse_nccl_unique_id.argtypes = [vp]
    lib.sse_shard_create.argtypes = [C.POINTER(SseConfig), C.POINTER(SseOperators),
                                     C.POINTER(SseGeometry), c_i64_p, C.c_int, C.c_int, vp,
                                     C.POINTER(vp)]
    lib.sse_shard_destroy.argtypes = [vp]
    lib.sse_shard_handle.argtypes = [vp]
    lib.sse_shard_handle.restype = vp
    lib.sse_shard_get_plan.argtypes = [vp, C.POINTER(SseShardPlan)]
    lib.sse_shard_residual.argtypes = [vp, vp, vp, C.c_double, C.c_int]
    lib.sse_shard_rk_stage.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    lib.sse_shard_rk_step_ck54.argtypes = [vp, C.c_double]
    lib.sse_shard_time_residual.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    lib.sse_shard_sync.argtypes = [vp]
    if path is None:
        _LIB = lib
    return lib


def measure_fp64_peak(device: int = 0) -> float:
    """Measured FP64 FMA throughput of the device in TFLOP/s (library microbenchmark)."""
    lib = load_library()
    out = C.c_double()
    if lib.sse_measure_fp64_peak(device, C.byref(out)) != 0:
        raise RuntimeError("sse_measure_fp64_peak failed: " + lib.sse_last_error().decode())
    return float(out.value)


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId for sse_shard_create (call on ONE rank, broadcast to the others)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.sse_nccl_unique_id(buf) != 0:
        raise RuntimeError("sse_nccl_unique_id failed: " + lib.sse_last_error().decode())
    return buf.raw


def shard_plan(mapP_cols: np.ndarray, N_e_global: int, rank: int, world: int):
    """sse_shard_plan_build (pure host): (plan, mapP_local (N_f, n_loc), send_idx) of ``rank`` from
    the columns [start, stop) of the global connectivity ``mapP_cols`` (N_f, n_loc)."""
    lib = load_library()
    N_f, n_loc = mapP_cols.shape
    cols = np.ascontiguousarray(np.asarray(mapP_cols).T, dtype=np.int64)      # [k][j]
    plan = SseShardPlan()
    mp = np.empty(N_f * n_loc, dtype=np.int64)
    send = np.empty(N_f * n_loc, dtype=np.int64)
    rc = lib.sse_shard_plan_build(cols.ctypes.data_as(c_i64_p), N_f, N_e_global, rank, world,
                                  C.byref(plan), mp.ctypes.data_as(c_i64_p),
                                  send.ctypes.data_as(c_i64_p))
    if rc != 0:
        raise RuntimeError("sse_shard_plan_build failed: " + lib.sse_last_error().decode())
    return plan, mp.reshape(n_loc, N_f).T, send[:plan.n_send].copy()


def bind_host_to_gpu_numa_node(device: int = 0) -> dict:
    """Best effort: restrict this process to the CPUs of the NUMA node the GPU hangs off, BEFORE
    pinned host buffers are allocated (first touch then places them on that node), so that the
    host<->device copies of ``sse_residual(where=HOST)`` do not cross the socket interconnect.
    Returns what was found and done; never raises (containers often hide the topology)."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        sysdev = "/sys/bus/pci/devices/%s:%s" % (dom[-4:].lower(), rest.lower())
        with open(sysdev + "/numa_node") as f:
            node = int(f.read().strip())
        info["gpu_numa_node"] = node
        allowed = os.sched_getaffinity(0)
        info["cpus_allowed"] = len(allowed)
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        local = allowed & cpus
        info["cpus_on_node"] = len(local)
        if local and local != allowed:
            os.sched_setaffinity(0, local)
            info["bound"] = True
    except Exception as e:  # noqa: BLE001 -- topology files / NVML may be absent
        info["error"] = type(e).__name__
    return info


def probe_elementary(which: str, x: np.ndarray, device: int = 0, lib=None) -> np.ndarray:
    """The device ``log`` / ``exp`` of the entropy-variable maps applied to a host array."""
    lib = lib or load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    rc = lib.sse_probe_elementary(device, {"log": 0, "exp": 1}[which], x.ctypes.data_as(c_d_p),
                                  y.ctypes.data_as(c_d_p), x.size)
    if rc != 0:
        raise RuntimeError("sse_probe_elementary failed: " + lib.sse_last_error().decode())
    return y


def measure_dmma_peak(device: int = 0) -> float:
    """Measured FP64 tensor-core (mma.sync m8n8k4) throughput in TFLOP/s."""
    lib = load_library()
    out = C.c_double()
    if lib.sse_measure_dmma_peak(device, C.byref(out)) != 0:
        raise RuntimeError("sse_measure_dmma_peak failed: " + lib.sse_last_error().decode())
    return float(out.value)


def _dp(a):
    return a.ctypes.data_as(c_d_p) if a is not None else c_d_p()


def _ip(a):
    return a.ctypes.data_as(c_i32_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def physical_operators(solver):
    """Per-element dense VOL/FAC (Solvers/operators.jl:85-164), built once on the host.

    Returns VOL (N_e, d, N_p, N_q) and FAC (N_e, N_p, N_f), C-contiguous."""
    sd = solver.spatial_discretization
    ra = sd.reference_approximation
    gf = sd.geometric_factors
    d, N_e = ra.dim, sd.N_e
    V, R = ra.V.to_dense(), ra.R.to_dense()
    D = [Dm.to_dense() for Dm in ra.D]
    W, B = ra.W, ra.B
    skew = solver.form_desc["mapping_form"] == "skew"
    from .geometric_factors import apply_reference_mapping
    Lq = apply_reference_mapping(gf, ra.reference_mapping)       # (N_e, n, m, N_q)
    J = gf.J_q
    if solver.mass_kind == "diagonal":
        Minv = np.zeros((N_e, ra.N_p, ra.N_p))
        idx = np.arange(ra.N_p)
        Minv[:, idx, idx] = 1.0 / (W[None, :] * J)
    elif solver.mass_kind == "cholesky":
        Minv = np.linalg.inv(np.einsum("qa,kq,qb->kab", V, W[None, :] * J, V))
    else:
        Minv = np.einsum("qa,kq,qb->kab", V, W[None, :] / J, V)
        if solver.Minv is not None:
            Minv = solver.Minv[None] @ Minv @ solver.Minv[None]
    VOL = np.empty((N_e, d, ra.N_p, ra.N_q))
    for n in range(d):
        if d == 1 and not skew:
            A = np.broadcast_to(D[0].T * W[None, :], (N_e,) + D[0].shape)
        elif not skew:
            A = sum(D[m].T[None] * (W[None, :] * Lq[:, n, m, :])[:, None, :] for m in range(d))
        else:
            A = sum(D[m].T[None] * (0.5 * W[None, :] * Lq[:, n, m, :])[:, None, :]
                    - (0.5 * W[None, :] * Lq[:, n, m, :])[:, :, None] * D[m][None]
                    for m in range(d))
            A = A + np.einsum("fq,kf,fr->kqr", R, 0.5 * B[None, :] * gf.nJf[:, :, n], R)
        VOL[:, n] = Minv @ (V.T[None] @ A)
    if d == 1 and not skew:
        FAC = -Minv @ np.broadcast_to(V.T @ (R.T * B[None, :]), (N_e, ra.N_p, ra.N_f))
    else:
        FAC = -Minv @ (V.T[None] @ (R.T[None] * (B[None, :] * gf.J_f)[:, None, :]))
    return np.ascontiguousarray(VOL), np.ascontiguousarray(FAC)


class DeviceResidual:
    """Owns an ``sse_handle``: device-resident operators, geometry, state and scratch."""

    def __init__(self, solver, device: int = 0, mapP: Optional[np.ndarray] = None,
                 n_halo: int = 0, elements: Optional[np.ndarray] = None, shard=None):
        """``elements``: optional index array selecting a shard of the mesh (multi-GPU);
        ``mapP``: (N_f, N_e_local) local connectivity override with halo slots.
        ``shard`` = (rank, world, nccl_id, N_e_global): create the handle through
        ``sse_shard_create`` -- ``mapP`` then holds the columns of the GLOBAL connectivity of this
        rank's elements, and the partition, the halo and the NCCL exchange live in the library."""
        self.lib = load_library()
        sd = solver.spatial_discretization
        ra = sd.reference_approximation
        gf = sd.geometric_factors
        law, form = solver.law_desc, solver.form_desc
        sel = slice(None) if elements is None else np.asarray(elements)
        N_e = sd.N_e if elements is None else len(sel)
        self.N_e, self.N_c, self.N_p = N_e, law["N_c"], ra.N_p
        self.N_f, self.N_q = ra.N_f, ra.N_q
        self._keep = []

        cfg = SseConfig()
        cfg.dim, cfg.N_p, cfg.N_q, cfg.N_f = ra.dim, ra.N_p, ra.N_q, ra.N_f
        cfg.N_c, cfg.num_faces = law["N_c"], ra.element_type.num_faces
        cfg.N_e, cfg.N_halo = N_e, n_halo
        cfg.law = LAW[law["kind"]]
        for m, am in enumerate(law.get("a", ())):
            cfg.a[m] = am
        cfg.b = law.get("b", 0.0)
        cfg.gamma = law.get("gamma", 1.4)
        cfg.form = 1 if form["kind"] == "flux_differencing" else 0
        cfg.strategy = 1 if form["strategy"] == "physical" else 0
        cfg.inviscid_flux = INVISCID[form["inviscid"][0]]
        cfg.half_lambda = form["inviscid"][1] if form["inviscid"][0] == "lf" else 0.0
        cfg.two_point_flux = 1 if form["two_point"] == "ec" else 0
        cfg.r_is_selection = int(isinstance(ra.R, SelectionMap))
        cfg.mass_solver = MASS[solver.mass_kind]
        cfg.device = device

        ops = SseOperators()
        V = ra.V
        if isinstance(V, IdentityMap):
            cfg.v_kind = 0
        elif isinstance(V, (WarpedTensorProductMap2D, WarpedTensorProductMap3D)):
            cfg.v_kind = 1
            ops.n1d = V.A.shape[0]
            A, Bt = _f64(V.A), _f64(V.B)
            sig = np.ascontiguousarray(V.sigma_i, dtype=np.int32)
            self._keep += [A, Bt, sig]
            ops.warp_A, ops.warp_B, ops.sigma_i = _dp(A), _dp(Bt), _ip(sig)
            if isinstance(V, WarpedTensorProductMap3D):
                Ct = _f64(V.C)
                self._keep.append(Ct)
                ops.warp_C = _dp(Ct)
        else:
            cfg.v_kind = 2
            Vd = _f64(V.to_dense())
            self._keep.append(Vd)
            ops.V_dense = _dp(Vd)
        if ops.n1d == 0 and ra.D_1D is not None and len({D1.shape[0] for D1 in ra.D_1D}) == 1:
            ops.n1d = ra.D_1D[0].shape[0]          # tensor-product element: nodes per direction
        rp, ci, val = ra.R.to_csr()
        self._keep += [rp, ci, val]
        ops.R_rowptr, ops.R_col, ops.R_val = _ip(rp), _ip(ci), _dp(val)
        for m, Dm in enumerate(ra.D):
            rp, ci, val = Dm.to_csr()
            self._keep += [rp, ci, val]
            ops.D_rowptr[m], ops.D_col[m], ops.D_val[m] = _ip(rp), _ip(ci), _dp(val)
        W, B = _f64(ra.W), _f64(ra.B)
        n_ref = _f64(gf.n_ref)
        self._keep += [W, B, n_ref]
        ops.W, ops.B, ops.n_ref = _dp(W), _dp(B), _dp(n_ref)
        if not isinstance(ra.reference_mapping, NoMapping):
            Lr = _f64(ra.reference_mapping.Lambda_ref)        # C-ordered [i][m][l]
            Jr = _f64(ra.reference_mapping.J_ref)
            self._keep += [Lr, Jr]
            ops.Lambda_ref, ops.J_ref = _dp(Lr), _dp(Jr)
        if solver.Minv is not None:
            Mi = _f64(solver.Minv)
            self._keep.append(Mi)
            ops.Minv = _dp(Mi)

        geo = SseGeometry()
        on_device = getattr(gf, "on_device", False) and elements is None
        if on_device:
            # geometric factors built on the device (sse_geometry_build): hand the device
            # pointers over, nothing crosses PCIe
            geo.J_q, geo.Lambda_q, geo.J_f, geo.nJf = gf.device_pointers()
            Jq = None
        else:
            Jq, Lq = _f64(gf.J_q[sel]), _f64(gf.Lambda_q[sel])
            Jf, nJf = _f64(gf.J_f[sel]), _f64(gf.nJf[sel])
            self._keep += [Jq, Lq, Jf, nJf]
            geo.J_q, geo.Lambda_q, geo.J_f, geo.nJf = _dp(Jq), _dp(Lq), _dp(Jf), _dp(nJf)
        if Jq is None and (form["strategy"] == "physical" or solver.mass_kind == "cholesky"):
            Jq = _f64(gf.J_q[sel])       # these host-side builders need J_q (downloads it once)
        if form["strategy"] == "physical":
            VOL, FAC = physical_operators(solver)
            VOL, FAC = _f64(VOL[sel]), _f64(FAC[sel])
            self._keep += [VOL, FAC]
            geo.VOL, geo.FAC = _dp(VOL), _dp(FAC)
        elif solver.mass_kind == "cholesky":
            Vd = ra.V.to_dense()
            Mi = _f64(np.linalg.inv(np.einsum("qa,kq,qb->kab", Vd, W[None, :] * Jq, Vd)))
            self._keep.append(Mi)
            geo.Minv_elem = _dp(Mi)

        if mapP is None:
            if elements is not None:
                raise ValueError("a shard needs its local mapP")
            mapP = sd.mesh.mapP
        mp = np.ascontiguousarray(np.asarray(mapP).T, dtype=np.int64)      # [k][j]
        h = C.c_void_p()
        self.shard = None
        if shard is not None:
            rank, world, nccl_id, N_e_global = shard
            cfg.N_e, cfg.N_halo = N_e_global, 0
            sh = C.c_void_p()
            idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
            rc = self.lib.sse_shard_create(C.byref(cfg), C.byref(ops), C.byref(geo),
                                           mp.ctypes.data_as(c_i64_p), rank, world, idbuf,
                                           C.byref(sh))
            if rc != 0:
                raise RuntimeError("sse_shard_create failed: " +
                                   self.lib.sse_last_error().decode())
            self.shard = sh
            h = C.c_void_p(self.lib.sse_shard_handle(sh))
        else:
            rc = self.lib.sse_create(C.byref(cfg), C.byref(ops), C.byref(geo),
                                     mp.ctypes.data_as(c_i64_p), C.byref(h))
            if rc != 0:
                raise RuntimeError("sse_create failed: " + self.lib.sse_last_error().decode())
        self.h = h
        self._keep = []   # everything was copied to the device

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: " + self.lib.sse_last_error().decode())

    def close(self):
        if getattr(self, "shard", None):
            self.lib.sse_shard_destroy(self.shard)      # destroys the handle as well
            self.shard = None
            self.h = None
        if getattr(self, "h", None):
            self.lib.sse_destroy(self.h)
            self.h = None

    # ------------------------------------------------------------------ sharded (sse_shard_*)
    def shard_plan(self):
        plan = SseShardPlan()
        self._check(self.lib.sse_shard_get_plan(self.shard, C.byref(plan)), "sse_shard_get_plan")
        return plan

    def shard_residual(self, u: Optional[np.ndarray] = None, dudt: Optional[np.ndarray] = None):
        """Device-resident (no arguments) or host-buffer residual of the shard."""
        if u is None:
            self._check(self.lib.sse_shard_residual(self.shard, None, None, 0.0, 1),
                        "sse_shard_residual")
            return
        assert u.shape == self.shape and dudt.shape == self.shape
        assert u.flags.c_contiguous and dudt.flags.c_contiguous
        self._check(self.lib.sse_shard_residual(self.shard, u.ctypes.data, dudt.ctypes.data, 0.0,
                                                0), "sse_shard_residual")

    def shard_time_residual(self, reps: int) -> float:
        ms = C.c_float()
        self._check(self.lib.sse_shard_time_residual(self.shard, reps, C.byref(ms)),
                    "sse_shard_time_residual")
        return float(ms.value)

    def shard_rk_step_ck54(self, dt: float):
        self._check(self.lib.sse_shard_rk_step_ck54(self.shard, dt), "sse_shard_rk_step_ck54")

    def shard_sync(self):
        self._check(self.lib.sse_shard_sync(self.shard), "sse_shard_sync")

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def shape(self):
        return (self.N_e, self.N_c, self.N_p)

    # ------------------------------------------------------------------ residual
    def residual_host(self, u: np.ndarray, dudt: np.ndarray, t: float = 0.0):
        """sse_residual with host buffers (H2D/D2H copies inside the call)."""
        if u.dtype != np.float64 or dudt.dtype != np.float64:
            raise TypeError("u and dudt must be Float64")
        if u.shape != self.shape or dudt.shape != self.shape:
            raise ValueError(f"expected arrays of shape {self.shape}")
        if not (u.flags.c_contiguous and dudt.flags.c_contiguous):
            raise ValueError("u and dudt must be contiguous")
        self._check(self.lib.sse_residual(self.h, u.ctypes.data, dudt.ctypes.data, t, 0),
                    "sse_residual")
        return dudt

    def residual_device(self, u_ptr: int, dudt_ptr: int, t: float = 0.0):
        self._check(self.lib.sse_residual(self.h, u_ptr, dudt_ptr, t, 1), "sse_residual")

    def nodal_values(self, u_ptr: int = 0):
        self._check(self.lib.sse_nodal_values(self.h, u_ptr or None), "sse_nodal_values")

    def time_derivative(self, dudt_ptr: int = 0):
        self._check(self.lib.sse_time_derivative(self.h, dudt_ptr or None),
                    "sse_time_derivative")

    def time_derivative_range(self, k_begin: int, k_end: int, dudt_ptr: int = 0):
        self._check(self.lib.sse_time_derivative_range(self.h, dudt_ptr or None, k_begin, k_end),
                    "sse_time_derivative_range")

    def set_stream(self, stream: int):
        self._check(self.lib.sse_set_stream(self.h, stream), "sse_set_stream")

    def upload_state(self, u: np.ndarray):
        assert u.shape == self.shape and u.dtype == np.float64 and u.flags.c_contiguous
        self._check(self.lib.sse_upload_state(self.h, u.ctypes.data), "sse_upload_state")

    def download_dudt(self, dudt: np.ndarray):
        assert dudt.shape == self.shape and dudt.dtype == np.float64 and dudt.flags.c_contiguous
        self._check(self.lib.sse_download_dudt(self.h, dudt.ctypes.data), "sse_download_dudt")

    def upload_and_nodal_values(self, u: np.ndarray):
        assert u.shape == self.shape and u.dtype == np.float64 and u.flags.c_contiguous
        self._check(self.lib.sse_upload_and_nodal_values(self.h, u.ctypes.data),
                    "sse_upload_and_nodal_values")

    def upload_range_and_nodal_values(self, u: np.ndarray, k_begin: int, k_end: int,
                                      first: bool = False):
        assert u.shape == self.shape and u.dtype == np.float64 and u.flags.c_contiguous
        self._check(self.lib.sse_upload_range_and_nodal_values(self.h, u.ctypes.data, k_begin,
                                                               k_end, int(first)),
                    "sse_upload_range_and_nodal_values")

    def set_copy_streams(self, split: bool):
        self._check(self.lib.sse_set_copy_streams(self.h, int(split)), "sse_set_copy_streams")

    def download_dudt_range(self, dudt: np.ndarray, k_begin: int, k_end: int):
        assert dudt.shape == self.shape and dudt.dtype == np.float64 and dudt.flags.c_contiguous
        self._check(self.lib.sse_download_dudt_range(self.h, dudt.ctypes.data, k_begin, k_end),
                    "sse_download_dudt_range")

    def sync_copies(self):
        self._check(self.lib.sse_sync_copies(self.h), "sse_sync_copies")

    def set_state(self, u: np.ndarray):
        u = _f64(u)
        assert u.shape == self.shape
        self._check(self.lib.sse_set_state(self.h, u.ctypes.data), "sse_set_state")

    def get_state(self) -> np.ndarray:
        u = np.empty(self.shape)
        self._check(self.lib.sse_get_state(self.h, u.ctypes.data), "sse_get_state")
        return u

    def state_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self.lib.sse_state_ptr(self.h, C.byref(a), C.byref(b)), "sse_state_ptr")
        return a.value, b.value

    def rk_stage(self, a: float, b: float, dt: float):
        self._check(self.lib.sse_rk_stage(self.h, a, b, dt), "sse_rk_stage")

    def erk_step(self, A: np.ndarray, b: np.ndarray, dt: float):
        """One general explicit Runge-Kutta step (tableau A, weights b) on the resident state."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        if A.ndim != 2 or A.shape[0] != A.shape[1] or b.shape != (A.shape[0],):
            raise ValueError("A must be (s, s) and b (s,)")
        self._check(self.lib.sse_erk_step(self.h, A.shape[0], _dp(A), _dp(b), dt), "sse_erk_step")

    def rk_step_ck54(self, dt: float):
        self._check(self.lib.sse_rk_step_ck54(self.h, dt), "sse_rk_step_ck54")

    def sync(self):
        self._check(self.lib.sse_sync(self.h), "sse_sync")

    def stream(self) -> int:
        return self.lib.sse_stream(self.h)

    def time_residual(self, reps: int, split: bool = False):
        ms = (C.c_float * 3)()
        self._check(self.lib.sse_time_residual(self.h, reps, int(split), ms),
                    "sse_time_residual")
        return float(ms[0]), float(ms[1]), float(ms[2])

    def kernel_launches(self) -> int:
        return int(self.lib.sse_kernel_launches(self.h))

    def device_bytes(self) -> int:
        return int(self.lib.sse_device_bytes(self.h))

    # ------------------------------------------------------------------ analysis functionals
    FUNCTIONALS = {"conservation": 0, "entropy": 1, "energy": 2, "energy_residual": 3,
                   "entropy_residual": 4, "l2_error": 5}

    def functional(self, which: str, arg: str = "state",
                   exact_q: Optional[np.ndarray] = None) -> np.ndarray:
        """sse_functional on the device-resident state / last residual (Analysis/conservation.jl
        :113-190, error.jl:58-91).  ``exact_q``: (N_e, N_c, N_q) host array for ``l2_error``."""
        kind = self.FUNCTIONALS[which]
        n_out = 1 if which in ("entropy", "entropy_residual") else self.N_c
        out = np.zeros(n_out)
        ptr = None
        if exact_q is not None:
            exact_q = _f64(exact_q)
            assert exact_q.shape == (self.N_e, self.N_c, self.N_q)
            ptr = exact_q.ctypes.data
        self._check(self.lib.sse_functional(self.h, kind, 1 if arg == "dudt" else 0, ptr,
                                            out.ctypes.data_as(c_d_p)), "sse_functional")
        return out

    # ------------------------------------------------------------------ halo
    def halo_setup(self, send_idx: np.ndarray):
        idx = np.ascontiguousarray(send_idx, dtype=np.int64)
        self._check(self.lib.sse_halo_setup(self.h, idx.ctypes.data_as(c_i64_p), len(idx)),
                    "sse_halo_setup")

    def halo_buffers(self):
        s, r = C.c_void_p(), C.c_void_p()
        ns, nr = C.c_int64(), C.c_int64()
        self._check(self.lib.sse_halo_buffers(self.h, C.byref(s), C.byref(r), C.byref(ns),
                                              C.byref(nr)), "sse_halo_buffers")
        return s.value, r.value, ns.value, nr.value

    def halo_pack(self):
        self._check(self.lib.sse_halo_pack(self.h), "sse_halo_pack")

    def halo_unpack(self):
        self._check(self.lib.sse_halo_unpack(self.h), "sse_halo_unpack")

    # second-order (BR1) equations on shards: the two loops of sse_time_derivative separately,
    # and the halo of the auxiliary-variable traces q_f exchanged in between
    def auxiliary_variable_range(self, k_begin: int, k_end: int):
        self._check(self.lib.sse_auxiliary_variable_range(self.h, k_begin, k_end),
                    "sse_auxiliary_variable_range")

    def time_derivative_only_range(self, k_begin: int, k_end: int, dudt_ptr: int = 0):
        self._check(self.lib.sse_time_derivative_only_range(self.h, dudt_ptr or None, k_begin,
                                                            k_end),
                    "sse_time_derivative_only_range")

    def halo_pack_aux(self):
        self._check(self.lib.sse_halo_pack_aux(self.h), "sse_halo_pack_aux")

    def halo_unpack_aux(self):
        self._check(self.lib.sse_halo_unpack_aux(self.h), "sse_halo_unpack_aux")
