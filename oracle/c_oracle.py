"""ctypes wrapper of the C/OpenMP oracle (oracle/c/sse_oracle.c) -- checker / CPU baseline only."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

import sse_oracle as oc

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsse_oracle.so")
dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)


class Problem(C.Structure):
    _fields_ = [("d", C.c_int), ("N_p", C.c_int), ("N_q", C.c_int), ("N_f", C.c_int),
                ("N_c", C.c_int), ("num_faces", C.c_int), ("n1", C.c_int), ("N_e", C.c_int64),
                ("law", C.c_int), ("a", C.c_double * 3), ("gamma", C.c_double),
                ("inviscid", C.c_int), ("half_lambda", C.c_double), ("two_point", C.c_int),
                ("proj", C.c_int), ("mass", C.c_int), ("has_C", C.c_int),
                ("V", dp), ("wA", dp), ("wB", dp), ("wC", dp), ("sig", ip),
                ("R_rp", ip), ("R_ci", ip), ("R_v", dp),
                ("S_cp", ip * 3), ("S_ri", ip * 3), ("S_v", dp * 3),
                ("C_cp", ip), ("C_ri", ip), ("C_v", dp),
                ("W", dp), ("B", dp), ("n_ref", dp),
                ("J_q", dp), ("L_q", dp), ("J_f", dp), ("nJf", dp), ("mapP", lp)]


def _lib():
    if not os.path.exists(LIB):
        raise RuntimeError("libsse_oracle.so not built (run __graft_entry__.build())")
    lib = C.CDLL(LIB)
    lib.oracle_threads.restype = C.c_int
    lib.oracle_residual_fluxdiff.argtypes = [C.POINTER(Problem), dp, dp, dp, dp]
    lib.oracle_residual_fluxdiff_range.argtypes = [C.POINTER(Problem), dp, dp, dp, dp, C.c_int64,
                                                   C.c_int64]
    lib.oracle_set_threads.argtypes = [C.c_int]
    lib.oracle_set_threads.restype = None
    return lib


def num_threads():
    return _lib().oracle_threads()


def set_threads(n):
    """Use n OpenMP threads from now on (overrides an inherited OMP_NUM_THREADS)."""
    _lib().oracle_set_threads(int(n))


def make_residual(prob, warped=None):
    """Returns fn(u) -> dudt for a flux-differencing oracle problem dict (tests/bridge.py).
    ``warped``: optional (A, B, C|None, sigma_i) tables to use the sum-factorised V."""
    lib = _lib()
    if prob["form"]["kind"] != "flux_differencing":
        raise ValueError("the C oracle restates the flux-differencing path only")
    if prob["mass_solver"] == "cholesky" or prob.get("Minv") is not None:
        raise ValueError("unsupported mass solver in the C oracle")
    keep = []

    def f64(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return a.ctypes.data_as(dp)

    def i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return a.ctypes.data_as(ip)

    P = Problem()
    law = prob["law"]
    P.d, P.N_p, P.N_q, P.N_f = prob["d"], prob["N_p"], prob["N_q"], prob["N_f"]
    P.N_c, P.num_faces, P.N_e = prob["N_c"], prob["num_faces"], prob["N_e"]
    P.law = dict(advection=0, burgers=1, euler=2)[law["kind"]]
    for m, am in enumerate(law.get("a", ())):
        P.a[m] = am
    P.gamma = law.get("gamma", 1.4)
    inv = prob["form"]["inviscid"]
    P.inviscid = dict(lf=0, central=1, ec=2)[inv[0]]
    P.half_lambda = inv[1] if inv[0] == "lf" else 0.0
    P.two_point = 1 if prob["form"]["two_point"] == "ec" else 0
    if law["N_c"] == 1 or (prob.get("V_is_identity") and prob.get("R_is_selection")):
        P.proj = 0
    elif prob.get("V_is_identity"):
        P.proj = 1
    else:
        P.proj = 2
    P.mass = 0 if prob["mass_solver"] == "diagonal" else 1
    P.V = f64(prob["V"])
    if warped is not None:
        A, B, Ct, sig = warped
        P.n1 = A.shape[0]
        P.wA, P.wB, P.sig = f64(A), f64(B), i32(sig)
        if Ct is not None:
            P.wC = f64(Ct)
    R = sp.csr_matrix(prob["R"])
    P.R_rp, P.R_ci, P.R_v = i32(R.indptr), i32(R.indices), f64(R.data)
    S, Cm = oc.flux_differencing_operators(prob)
    for m in range(prob["d"]):
        Sm = sp.csc_matrix(S[m])
        P.S_cp[m], P.S_ri[m], P.S_v[m] = i32(Sm.indptr), i32(Sm.indices), f64(Sm.data)
    P.has_C = int(Cm is not None)
    if Cm is not None:
        Cc = sp.csc_matrix(Cm)
        P.C_cp, P.C_ri, P.C_v = i32(Cc.indptr), i32(Cc.indices), f64(Cc.data)
    P.W, P.B, P.n_ref = f64(prob["W"]), f64(prob["B"]), f64(prob["n_ref"])
    P.J_q, P.L_q = f64(prob["J_q"]), f64(prob["Lambda_q"])
    P.J_f, P.nJf = f64(prob["J_f"]), f64(prob["nJf"])
    mp = np.ascontiguousarray(prob["mapP"].T, dtype=np.int64)
    keep.append(mp)
    P.mapP = mp.ctypes.data_as(lp)
    N_e, N_c, N_p, N_q, N_f = prob["N_e"], prob["N_c"], prob["N_p"], prob["N_q"], prob["N_f"]
    u_q = np.empty((N_e, N_c, N_q))
    u_f = np.empty((N_c, N_e, N_f))

    def fn(u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty_like(u)
        rc = lib.oracle_residual_fluxdiff(C.byref(P), u.ctypes.data_as(dp), out.ctypes.data_as(dp),
                                          u_q.ctypes.data_as(dp), u_f.ctypes.data_as(dp))
        assert rc == 0
        return out

    def fn_range(u, out, k0, k1):
        """Loops A and B on the elements [k0, k1) only, into the caller's ``out`` (the traces of
        neighbours outside the range must be valid from an earlier full call)."""
        rc = lib.oracle_residual_fluxdiff_range(C.byref(P), u.ctypes.data_as(dp),
                                                out.ctypes.data_as(dp), u_q.ctypes.data_as(dp),
                                                u_f.ctypes.data_as(dp), int(k0), int(k1))
        assert rc == 0

    fn.range = fn_range
    fn._keep = (keep, P)
    return fn
