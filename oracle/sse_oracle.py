"""CPU ORACLE (test infrastructure only -- never imported by the product path).

A NumPy restatement of the reference's semi-discrete residual ``semi_discrete_residual!`` and
of everything it calls, used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg as the *checker*.  Each function cites the reference file:line it follows
(paths relative to /root/reference/src/).

PINNING STATUS.  The reference (100 % Julia) cannot run in this environment and its repository
holds no residual-level golden vectors.  What pins this oracle instead (tests/test_oracle_golden.py):
  * the end-to-end L2 errors hard-coded in the reference's own test suite
    (/root/reference/test/runtests.jl:34-143) for the cases whose meshes are unambiguous without
    StartUpDG (1-D advection-diffusion ModalMulti/BR1/PhysicalOperator; 1-D Euler Gauss
    collocation flux differencing with entropy projection and facet correction; 2-D Euler
    vortex ModalTensor Tri flux differencing with Lax-Friedrichs; 3-D Euler NodalTensor Hex
    flux differencing with the EC interface flux and conservative-curl metrics), reproduced by
    integrating this oracle with the same scheme (CK54 / DP8) and time step;
  * the reference's conservation / energy / entropy invariants (Analysis/conservation.jl:145-190).
For the north-star configuration itself (flux differencing on tetrahedra) the reference has no
test of any kind: at that boundary parity is "unpinned by golden vectors" and rests on the
shared code path with the pinned cases -- the 3-D Euler physics (two-point flux, entropy maps,
3-D conservative-curl metrics) through the Hex case, the collapsed-coordinate modal operators,
entropy projection and facet correction through the 2-D Tri case -- plus the invariants.

Array convention: a Julia array A[i1, ..., k] is the C-contiguous NumPy array A[k, ..., i1]
(identical bytes).  ``u``/``dudt`` are (N_e, N_c, N_p).
"""
from __future__ import annotations

import numpy as np


# =============================================================================== physics
def _real(x):
    """Floating-point view of x: FP64 unless the caller works in extended precision
    (np.longdouble input runs the whole oracle in 80-bit arithmetic: the accuracy reference of
    tests/test_oracle_extended_precision.py)."""
    x = np.asarray(x)
    return x if x.dtype == np.longdouble else np.asarray(x, dtype=np.float64)


def logmean(x, y):
    """ConservationLaws/ConservationLaws.jl:132-145."""
    x = _real(x)
    y = _real(y)
    f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y)
    with np.errstate(divide="ignore", invalid="ignore"):
        taylor = (x + y) * 105 / (210 + f2 * (70 + f2 * (42 + f2 * 30)))
        full = (y - x) / np.log(y / x)
    return np.where(f2 < 1.0e-4, taylor, full)


def inv_logmean(x, y):
    """ConservationLaws/ConservationLaws.jl:147-156."""
    x = _real(x)
    y = _real(y)
    f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y)
    with np.errstate(divide="ignore", invalid="ignore"):
        taylor = (210 + f2 * (70 + f2 * (42 + f2 * 30))) / ((x + y) * 105)
        full = np.log(y / x) / (y - x)
    return np.where(f2 < 1.0e-4, taylor, full)


def physical_flux(law, u, q=None):
    """(…, N_c) -> (…, N_c, d).  linear_advection_diffusion.jl:53-75, burgers.jl:51-72,
    euler_navierstokes.jl:58-69."""
    kind = law["kind"]
    d = law["d"]
    if kind in ("advection", "advection_diffusion"):
        f = np.stack([law["a"][m] * u for m in range(d)], axis=-1)
        if kind == "advection_diffusion":
            f = f - law["b"] * q
        return f
    if kind in ("burgers", "viscous_burgers"):
        f = np.stack([0.5 * law["a"][m] * u ** 2 for m in range(d)], axis=-1)
        if kind == "viscous_burgers":
            f = f - law["b"] * q
        return f
    if kind == "euler":
        gm1 = law["gamma"] - 1.0
        rho = u[..., 0]
        V = [u[..., m + 1] / rho for m in range(d)]
        p = gm1 * (u[..., -1] - 0.5 * sum(u[..., m + 1] * V[m] for m in range(d)))
        h_t = u[..., -1] + p
        f = np.empty(u.shape + (d,), dtype=u.dtype)
        for n in range(d):
            f[..., 0, n] = u[..., n + 1]
            for m in range(d):
                f[..., m + 1, n] = u[..., m + 1] * V[n] + (p if m == n else 0.0)
            f[..., d + 1, n] = h_t * V[n]
        return f
    raise ValueError(kind)


def two_point_flux(law, flux_kind, uL, uR):
    """(…, N_c) x2 -> (…, N_c, d).  linear_advection_diffusion.jl:113-119, burgers.jl:110-133,
    euler_navierstokes.jl:152-195."""
    kind = law["kind"]
    d = law["d"]
    if kind in ("advection", "advection_diffusion"):
        f1 = 0.5 * (uL + uR)
        return np.stack([law["a"][m] * f1 for m in range(d)], axis=-1)
    if kind in ("burgers", "viscous_burgers"):
        if flux_kind == "ec":
            f1 = (uL ** 2 + uL * uR + uR ** 2) / 6
        else:
            f1 = (uL ** 2 + uR ** 2) * 0.25
        return np.stack([law["a"][m] * f1 for m in range(d)], axis=-1)
    if kind == "euler":
        if flux_kind == "conservative":
            return 0.5 * (physical_flux(law, uL) + physical_flux(law, uR))
        gm1 = law["gamma"] - 1.0
        inv_gm1 = 1.0 / gm1
        V_L = [uL[..., m + 1] / uL[..., 0] for m in range(d)]
        V_R = [uR[..., m + 1] / uR[..., 0] for m in range(d)]
        p_L = gm1 * (uL[..., -1] - 0.5 * uL[..., 0] * sum(V_L[m] ** 2 for m in range(d)))
        p_R = gm1 * (uR[..., -1] - 0.5 * uR[..., 0] * sum(V_R[m] ** 2 for m in range(d)))
        rho_avg = logmean(uL[..., 0], uR[..., 0])
        V_avg = [0.5 * (V_L[m] + V_R[m]) for m in range(d)]
        p_avg = 0.5 * (p_L + p_R)
        C = 0.5 * sum(V_L[m] * V_R[m] for m in range(d)) + inv_gm1 * inv_logmean(
            uL[..., 0] / p_L, uR[..., 0] / p_R)
        shape = np.broadcast(uL[..., 0], uR[..., 0]).shape
        f = np.empty(shape + (d + 2, d), dtype=np.result_type(uL, uR))
        for n in range(d):
            f_rho = rho_avg * V_avg[n]
            f[..., 0, n] = f_rho
            for m in range(d):
                # f_ρV[m, n] = f_ρ[m] V_avg[n] + δ_mn p_avg  (euler_navierstokes.jl:192)
                f[..., m + 1, n] = rho_avg * V_avg[m] * V_avg[n] + (p_avg if m == n else 0.0)
            f[..., d + 1, n] = f_rho * C + 0.5 * (p_L * V_R[n] + p_R * V_L[n])
        return f
    raise ValueError(kind)


def wave_speed(law, u_in, u_out, n):
    """linear_advection_diffusion.jl:107-111, burgers.jl:102-108, euler_navierstokes.jl:133-150.
    n: (…, d)."""
    kind = law["kind"]
    d = law["d"]
    if kind in ("advection", "advection_diffusion"):
        return np.abs(sum(law["a"][m] * n[..., m] for m in range(d)))
    if kind in ("burgers", "viscous_burgers"):
        a_n = sum(law["a"][m] * n[..., m] for m in range(d))
        return np.maximum(np.abs(a_n * u_in[..., 0]), np.abs(a_n * u_out[..., 0]))
    g = law["gamma"]
    gm1 = g - 1.0

    def prim(u):
        V = [u[..., m + 1] / u[..., 0] for m in range(d)]
        p = gm1 * (u[..., -1] - (0.5 / u[..., 0]) * sum(u[..., m + 1] ** 2 for m in range(d)))
        Vn = sum(V[m] * n[..., m] for m in range(d))
        return Vn, np.sqrt(g * p / u[..., 0])

    Vn_in, c_in = prim(u_in)
    Vn_out, c_out = prim(u_out)
    return np.maximum(np.abs(Vn_in), np.abs(Vn_out)) + np.maximum(c_in, c_out)


def numerical_flux(law, inviscid, u_in, u_out, n_f, flux_kind="conservative"):
    """ConservationLaws.jl:75-128.  inviscid = ("lf", halfλ) | ("central",) | ("ec",)."""
    f_s = two_point_flux(law, flux_kind, u_in, u_out)
    f_n = np.einsum("...em,...m->...e", f_s, n_f)
    if inviscid[0] == "lf":
        a = inviscid[1] * wave_speed(law, u_in, u_out, n_f)
        return f_n + a[..., None] * (u_in - u_out)
    return f_n


def conservative_to_entropy(law, u):
    """euler_navierstokes.jl:100-114 (identity for scalar laws: nodal_values! skips it)."""
    if law["kind"] != "euler":
        return u.copy()
    d = law["d"]
    g = law["gamma"]
    gm1 = g - 1.0
    k = (0.5 / u[..., 0]) * sum(u[..., m + 1] ** 2 for m in range(d))
    p = gm1 * (u[..., -1] - k)
    inv_p = 1.0 / p
    w = np.empty_like(u)
    w[..., 0] = (1.0 / gm1) * (g - np.log(p / (u[..., 0] ** g))) - k * inv_p
    for m in range(d):
        w[..., m + 1] = u[..., m + 1] * inv_p
    w[..., d + 1] = -u[..., 0] * inv_p
    return w


def entropy_to_conservative(law, w):
    """euler_navierstokes.jl:116-131."""
    if law["kind"] != "euler":
        return w.copy()
    d = law["d"]
    g = law["gamma"]
    gm1 = g - 1.0
    inv_gm1 = 1.0 / gm1
    w = w * gm1
    k = sum(w[..., m + 1] ** 2 for m in range(d)) / (2 * w[..., d + 1])
    s = g - w[..., 0] + k
    rho_e = (gm1 / ((-w[..., d + 1]) ** g)) ** inv_gm1 * np.exp(-s * inv_gm1)
    u = np.empty_like(w)
    u[..., 0] = -w[..., d + 1] * rho_e
    for m in range(d):
        u[..., m + 1] = w[..., m + 1] * rho_e
    u[..., d + 1] = rho_e * (1 - k)
    return u


def entropy(law, u):
    """euler_navierstokes.jl:93-98; ½u² for scalar laws."""
    if law["kind"] != "euler":
        return 0.5 * u[..., 0] ** 2
    d = law["d"]
    g = law["gamma"]
    p = (g - 1.0) * (u[..., -1] - (0.5 / u[..., 0]) * sum(u[..., m + 1] ** 2 for m in range(d)))
    return -u[..., 0] * np.log(p / (u[..., 0] ** g)) / (g - 1.0)


# ========================================================================== mass solvers
def mass_matrix_inverse(prob, k):
    """Solvers/mass_matrix.jl:153-167."""
    V, W = prob["V"], prob["W"]
    J = prob["J_q"][k]
    ms = prob["mass_solver"]
    if ms == "diagonal":
        return np.diag(1.0 / (W * J))
    if ms == "cholesky":
        return np.linalg.inv(V.T @ ((W * J)[:, None] * V))
    Minv = prob.get("Minv")
    A = V.T @ ((W / J)[:, None] * V)
    return A if Minv is None else Minv @ A @ Minv


def mass_matrix(prob, k):
    """Solvers/mass_matrix.jl:138-151."""
    ms = prob["mass_solver"]
    if ms == "cholesky":
        return prob["V"].T @ ((prob["W"] * prob["J_q"][k])[:, None] * prob["V"])
    return np.linalg.inv(mass_matrix_inverse(prob, k))


def mass_matrix_solve(prob, rhs):
    """Solvers/mass_matrix.jl:169-196.  rhs: (N_e, N_c, N_p)."""
    V, W, J = prob["V"], prob["W"], prob["J_q"]
    ms = prob["mass_solver"]
    if ms == "diagonal":
        return rhs / (W[None, :] * J)[:, None, :]
    if ms == "cholesky":
        M = np.einsum("qa,kq,qb->kab", V, W[None, :] * J, V)
        return np.linalg.solve(M, rhs.transpose(0, 2, 1)).transpose(0, 2, 1)
    Minv = prob.get("Minv")
    if Minv is not None:
        rhs = rhs @ Minv.T
    tmp = rhs @ V.T                                     # V rhs: (N_e, N_c, N_q)
    tmp = tmp * (W[None, :] / J)[:, None, :]
    out = tmp @ V
    if Minv is not None:
        out = out @ Minv.T
    return out


# ============================================================== derived operator bundles
def flux_differencing_operators(prob):
    """Solvers/operators.jl:40-83,166-225 and SpatialDiscretizations.jl:414-419:
    S_m = ½(W D_ξm − D_ξmᵀ W), C = Rᵀ B (None when R is a selection, i.e. diag-E)."""
    d = prob["d"]
    W, B, R = prob["W"], prob["B"], prob["R"]
    D = prob["D"]
    if prob.get("Lambda_ref") is not None:
        L, Jr = prob["Lambda_ref"], prob["J_ref"]
        D_xi = [sum((L[:, l, m] / Jr)[:, None] * D[l] for l in range(d)) for m in range(d)]
    else:
        D_xi = D
    S = [0.5 * (W[:, None] * D_xi[m] - D_xi[m].T * W[None, :]) for m in range(d)]
    C = None if prob.get("R_is_selection", False) else R.T * B[None, :]
    return S, C


def _facet_geometry(prob):
    nJf, J_f = prob["nJf"], prob["J_f"]
    n_f = nJf / J_f[:, :, None]                        # (N_e, N_f, d)
    BJf = prob["B"][None, :] * J_f
    return n_f, BJf


def _gather_exterior(u_f, mapP):
    """u_f[CI[mapP[:, k]], :] (flux_differencing_form.jl:312-313). u_f: (N_e, N_f, N_c)."""
    N_e, N_f, N_c = u_f.shape
    return u_f.reshape(N_e * N_f, N_c)[mapP.T.reshape(-1)].reshape(N_e, N_f, N_c)


# ================================================================ flux-differencing form
def nodal_values_fluxdiff(prob, u):
    """flux_differencing_form.jl:171-292.  Returns u_q (N_e,N_q,N_c), u_f (N_e,N_f,N_c)."""
    law = prob["law"]
    V, R = prob["V"], prob["R"]
    u_q = np.einsum("qp,kep->kqe", V, u)
    if law["N_c"] == 1:                                # :253-265
        return u_q, np.einsum("fq,kqe->kfe", R, u_q)
    if prob.get("V_is_identity", False) and prob.get("R_is_selection", False):   # :171-187
        return u_q, np.einsum("fq,kqe->kfe", R, u_q)
    if prob.get("V_is_identity", False):               # :190-211
        w_q = conservative_to_entropy(law, u_q)
        w_f = np.einsum("fq,kqe->kfe", R, w_q)
        return u_q, entropy_to_conservative(law, w_f)
    # general (modal) :214-250
    w_q = conservative_to_entropy(law, u_q)
    w_q = w_q * (prob["W"][None, :] * prob["J_q"])[:, :, None]
    w = np.einsum("qp,kqe->kep", V, w_q)
    w = mass_matrix_solve(prob, w)
    w_q = np.einsum("qp,kep->kqe", V, w)
    w_f = np.einsum("fq,kqe->kfe", R, w_q)
    return entropy_to_conservative(law, w_q), entropy_to_conservative(law, w_f)


def flux_difference(prob, S, u_q):
    """flux_differencing_form.jl:1-75: r[i] -= Σ_m S_m[i,j] Σ_n (Λ_i[m,n]+Λ_j[m,n]) F_n(u_i,u_j),
    r[j] += the same, over pairs i<j of the union sparsity of S."""
    law = prob["law"]
    d = prob["d"]
    flux_kind = prob["form"]["two_point"]
    Lq = prob["Lambda_q"]                              # (N_e, n, m, N_q)
    N_q = S[0].shape[0]
    mask = np.zeros((N_q, N_q), dtype=bool)
    for m in range(d):
        mask |= S[m] != 0.0
    mask = np.triu(mask | mask.T, k=1)
    I, Jx = np.nonzero(mask)
    r_q = np.zeros_like(u_q)
    F = two_point_flux(law, flux_kind, u_q[:, I, :], u_q[:, Jx, :])     # (N_e, P, N_c, d)
    Lsum = Lq[:, :, :, I] + Lq[:, :, :, Jx]                               # (N_e, n, m, P)
    Svals = np.stack([S[m][I, Jx] for m in range(d)], axis=0)             # (m, P)
    coef = np.einsum("mp,knmp->kpn", Svals, Lsum)                         # Σ_m S_m (Λ_i+Λ_j)[m,n]
    diff = np.einsum("kpn,kpen->kpe", coef, F)
    np.subtract.at(r_q, (slice(None), I), diff)
    np.add.at(r_q, (slice(None), Jx), diff)
    return r_q


def facet_correction(prob, C, r_q, f_f, u_q, u_f):
    """flux_differencing_form.jl:78-168: for (i,j) in nz(C):
    diff = C[i,j] Σ_m (halfnJf[m,j] + halfnJq[m,f(j),i]) F_m(u_q[i], u_f[j]);
    r_q[i] -= diff; f_f[j] -= diff."""
    if C is None:
        return r_q, f_f
    law = prob["law"]
    flux_kind = prob["form"]["two_point"]
    nfaces = prob["num_faces"]
    N_f = C.shape[1]
    npf = N_f // nfaces
    I, Jx = np.nonzero(C)
    face = Jx // npf
    halfnJf = 0.5 * prob["nJf"]                                           # (N_e, N_f, d)
    # nJq[n,f,i,k] = Σ_m Λ_q[i,m,n,k] n_ref[m,f]  (SpatialDiscretizations/mesh.jl:266-271)
    halfnJq = 0.5 * np.einsum("knmi,fm->kifn", prob["Lambda_q"], prob["n_ref"])
    F = two_point_flux(law, flux_kind, u_q[:, I, :], u_f[:, Jx, :])      # (N_e, P, N_c, d)
    nJ = halfnJf[:, Jx, :] + halfnJq[:, I, face, :]                       # (N_e, P, d)
    diff = C[I, Jx][None, :, None] * np.einsum("kpm,kpem->kpe", nJ, F)
    r_q = r_q.copy()
    f_f = f_f.copy()
    np.subtract.at(r_q, (slice(None), I), diff)
    np.subtract.at(f_f, (slice(None), Jx), diff)
    return r_q, f_f


def fluxdiff_loop_b(prob, u_q, u_f, u_out):
    """time_derivative! (flux_differencing_form.jl:294-347) given the exterior traces u_out."""
    law, form = prob["law"], prob["form"]
    V, R = prob["V"], prob["R"]
    S, C = prob.get("_SC") or flux_differencing_operators(prob)
    prob["_SC"] = (S, C)
    n_f, BJf = _facet_geometry(prob)
    f_f = numerical_flux(law, form["inviscid"], u_f, u_out, n_f, form["two_point"])
    f_f = f_f * BJf[:, :, None]
    r_q = flux_difference(prob, S, u_q)
    r_q, f_f = facet_correction(prob, C, r_q, f_f, u_q, u_f)
    r_q = r_q - np.einsum("fq,kfe->kqe", R, f_f)
    dudt = np.einsum("qp,kqe->kep", V, r_q)
    return mass_matrix_solve(prob, dudt)


def residual_fluxdiff(prob, u):
    """Solvers.jl:476-518 with flux_differencing_form.jl:268-347: loop A, (barrier), loop B."""
    u_q, u_f = nodal_values_fluxdiff(prob, u)
    u_out = _gather_exterior(u_f, prob["mapP"])
    return fluxdiff_loop_b(prob, u_q, u_f, u_out)


# ================================================= standard form, reference operators
def residual_standard_reference(prob, u):
    """Solvers.jl:476-518 with standard_form_first_order.jl:1-63 (always skew-symmetric)."""
    law, form = prob["law"], prob["form"]
    d = prob["d"]
    V, R, D, W = prob["V"], prob["R"], prob["D"], prob["W"]
    n_f, BJf = _facet_geometry(prob)
    # Λ_η (SpatialDiscretizations.jl:399-412) and halfWΛ (operators.jl:19-21)
    Lq = prob["Lambda_q"]
    if prob.get("Lambda_ref") is not None:
        Lq = np.einsum("iml,knli->knmi", prob["Lambda_ref"] / prob["J_ref"][:, None, None], Lq)
    halfWL = 0.5 * W[None, None, None, :] * Lq          # (N_e, n, m, N_q)
    u_q = np.einsum("qp,kep->kqe", V, u)
    u_f = np.einsum("fq,kqe->kfe", R, u_q)
    f_q = physical_flux(law, u_q)                        # (N_e, N_q, N_c, d)
    u_out = _gather_exterior(u_f, prob["mapP"])
    f_f = numerical_flux(law, form["inviscid"], u_f, u_out, n_f, "conservative")
    r_q = np.zeros_like(u_q)
    for n in range(d):
        for m in range(d):
            tmp = halfWL[:, n, m, :, None] * f_q[..., n]
            r_q += np.einsum("ji,kje->kie", D[m], tmp)                   # D_mᵀ
            r_q -= halfWL[:, n, m, :, None] * np.einsum("ij,kje->kie", D[m], f_q[..., n])
        f_n = np.einsum("fq,kqe->kfe", R, f_q[..., n])
        f_f = f_f - 0.5 * n_f[:, :, n, None] * f_n
    f_f = f_f * BJf[:, :, None]
    r_q = r_q - np.einsum("fq,kfe->kqe", R, f_f)
    dudt = np.einsum("qp,kqe->kep", V, r_q)
    return mass_matrix_solve(prob, dudt)


# ================================================== standard form, physical operators
def physical_operators(prob):
    """Solvers/operators.jl:85-164: per-element dense VOL[k][m] (N_p x N_q), FAC[k] (N_p x N_f)."""
    if "_PHYS" in prob:
        return prob["_PHYS"]
    d, N_e = prob["d"], prob["N_e"]
    V, R, D, W, B = prob["V"], prob["R"], prob["D"], prob["W"], prob["B"]
    nJf, J_f = prob["nJf"], prob["J_f"]
    skew = prob["form"].get("mapping_form", "skew") == "skew"
    Lq = prob["Lambda_q"]
    if prob.get("Lambda_ref") is not None:
        Lq = np.einsum("iml,knli->knmi", prob["Lambda_ref"] / prob["J_ref"][:, None, None], Lq)
    VOL = np.empty((N_e, d, V.shape[1], V.shape[0]))
    FAC = np.empty((N_e, V.shape[1], R.shape[0]))
    for k in range(N_e):
        Minv = mass_matrix_inverse(prob, k)
        for n in range(d):
            if d == 1 and not skew:                     # operators.jl:85-102
                A = D[0].T * W[None, :]
            elif not skew:                              # :104-133
                A = sum(D[m].T * (W * Lq[k, n, m])[None, :] for m in range(d))
            else:                                       # :135-164
                A = sum(D[m].T * (0.5 * W * Lq[k, n, m])[None, :]
                        - (0.5 * W * Lq[k, n, m])[:, None] * D[m] for m in range(d)) \
                    + R.T @ ((0.5 * B * nJf[k, :, n])[:, None] * R)
            VOL[k, n] = Minv @ (V.T @ A)
        if d == 1 and not skew:
            FAC[k] = -Minv @ (V.T @ (R.T * B[None, :]))
        else:
            FAC[k] = -Minv @ (V.T @ (R.T * (B * J_f[k])[None, :]))
    prob["_PHYS"] = (VOL, FAC)
    return VOL, FAC


def _n_f_physical(prob):
    # 1-D StandardMapping stores nJf un-normalised (operators.jl:101); |nJf| = 1 there anyway.
    return prob["nJf"] / prob["J_f"][:, :, None]


def residual_standard_physical_first_order(prob, u):
    """standard_form_first_order.jl:1-14,65-94."""
    law, form = prob["law"], prob["form"]
    V, R = prob["V"], prob["R"]
    VOL, FAC = physical_operators(prob)
    n_f = _n_f_physical(prob)
    u_q = np.einsum("qp,kep->kqe", V, u)
    u_f = np.einsum("fq,kqe->kfe", R, u_q)
    f_q = physical_flux(law, u_q)
    u_out = _gather_exterior(u_f, prob["mapP"])
    f_f = numerical_flux(law, form["inviscid"], u_f, u_out, n_f, "conservative")
    return np.einsum("kmpq,kqem->kep", VOL, f_q) + np.einsum("kpf,kfe->kep", FAC, f_f)


def second_order_auxiliary_variable(prob, u_q, u_f, u_out):
    """Loop A2, ``auxiliary_variable!`` (standard_form_second_order.jl:3-40): BR1 gradient
    q = -(VOL u_q + FAC u* n), u* n = ½(u⁻+u⁺) n (linear_advection_diffusion.jl:77-89);
    returns q_q (N_e,N_q,N_c,d) and q_f (N_e,N_f,N_c,d)."""
    V, R = prob["V"], prob["R"]
    VOL, FAC = physical_operators(prob)
    n_f = _n_f_physical(prob)
    u_n = 0.5 * (u_f + u_out)[..., None] * n_f[:, :, None, :]          # (N_e,N_f,N_c,d)
    q = -(np.einsum("kmpq,kqe->kepm", VOL, u_q) + np.einsum("kpf,kfem->kepm", FAC, u_n))
    q_q = np.einsum("qp,kepm->kqem", V, q)
    q_f = np.einsum("fq,kqem->kfem", R, q_q)
    return q_q, q_f


def second_order_time_derivative(prob, u_q, u_f, u_out, q_q, q_f, q_out):
    """Loop B, ``time_derivative!`` (standard_form_second_order.jl:42-75)."""
    law, form = prob["law"], prob["form"]
    VOL, FAC = physical_operators(prob)
    n_f = _n_f_physical(prob)
    f_q = physical_flux(law, u_q, q_q)
    f_f = numerical_flux(law, form["inviscid"], u_f, u_out, n_f, "conservative")
    # BR1 viscous flux: f* += Σ_m b (−½(q⁻+q⁺))_m n_m  (linear_advection_diffusion.jl:91-105)
    f_f = f_f + np.einsum("kfem,kfm->kfe", law["b"] * (-0.5) * (q_f + q_out), n_f)
    return np.einsum("kmpq,kqem->kep", VOL, f_q) + np.einsum("kpf,kfe->kep", FAC, f_f)


def residual_standard_physical_second_order(prob, u):
    """Solvers.jl:520-570 with standard_form_second_order.jl:3-75 (BR1): three element loops."""
    d = prob["d"]
    mapP = prob["mapP"]
    u_q = np.einsum("qp,kep->kqe", prob["V"], u)
    u_f = np.einsum("fq,kqe->kfe", prob["R"], u_q)
    u_out = _gather_exterior(u_f, mapP)
    q_q, q_f = second_order_auxiliary_variable(prob, u_q, u_f, u_out)
    q_out = np.stack([_gather_exterior(q_f[..., m], mapP) for m in range(d)], axis=-1)
    return second_order_time_derivative(prob, u_q, u_f, u_out, q_q, q_f, q_out)


def semi_discrete_residual(prob, u, t=0.0):
    """Dispatch of Solvers.jl:287-377 / 476-570 (``t`` is unused by the reference too)."""
    form = prob["form"]
    second = prob["law"]["kind"] in ("advection_diffusion", "viscous_burgers")
    if second:
        return residual_standard_physical_second_order(prob, u)
    if form["kind"] == "flux_differencing":
        return residual_fluxdiff(prob, u)
    if form.get("strategy", "reference") == "physical":
        return residual_standard_physical_first_order(prob, u)
    return residual_standard_reference(prob, u)


# ====================================================================== time integration
CK54_A = [0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238,
          -3550918686646 / 2091501179385, -1275806237668 / 842570457699]
CK54_B = [1432997174477 / 9575080441755, 5161836677717 / 13612068292357,
          1720146321549 / 2090206949498, 3134564353537 / 4481467310338,
          2277821191437 / 14882151754819]
CK54_C = [0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896,
          2006345519317 / 3224310063776, 2802321613138 / 2924317926251]


def ck54_integrate(rhs, u0, tspan, dt, callback=None):
    """Carpenter-Kennedy (5,4) 2N low-storage RK, fixed dt with the last step clipped to hit
    tspan[1] (what OrdinaryDiffEq's CarpenterKennedy2N54(adaptive=false) does in the
    reference's tests, test/test_driver.jl:77-83)."""
    u = u0.copy()
    t = tspan[0]
    k = np.zeros_like(u)
    step = 0
    while t < tspan[1] - 1e-12 * max(1.0, abs(tspan[1])):
        h = min(dt, tspan[1] - t)
        for s in range(5):
            k = CK54_A[s] * k + h * rhs(u, t + CK54_C[s] * h)
            u = u + CK54_B[s] * k
        t += h
        step += 1
        if callback is not None:
            callback(u, t, step)
    return u


def dp8_integrate(rhs, u0, tspan, n_steps):
    """Dormand-Prince 8(5,3) with a fixed step -- OrdinaryDiffEq's ``DP8(adaptive=false)`` as
    used by test/euler_3d.jl:44-51.  OrdinaryDiffEq is not vendored; DP8 there is Hairer's DOP853
    tableau, whose published coefficients ship with SciPy (scipy.integrate DOP853): 12 stages,
    u_{n+1} = u_n + h Σ b_s k_s."""
    from scipy.integrate._ivp import dop853_coefficients as dc
    A, B, C = dc.A[:12, :12], dc.B, dc.C[:12]
    u = u0.copy()
    h = (tspan[1] - tspan[0]) / n_steps
    for n in range(n_steps):
        t = tspan[0] + n * h
        K = []
        for s in range(12):
            us = u.copy()
            for j in range(s):
                if A[s, j] != 0.0:
                    us += (h * A[s, j]) * K[j]
            K.append(rhs(us, t + C[s] * h))
        for s in range(12):
            if B[s] != 0.0:
                u = u + (h * B[s]) * K[s]
    return u


# ========================================================================== functionals
def conservation_residual(prob, dudt):
    """Analysis/conservation.jl:145-152: Σ_k 1ᵀ W J_k V dudt_k per variable."""
    WJ = prob["W"][None, :] * prob["J_q"]
    return np.einsum("kq,qp,kep->e", WJ, prob["V"], dudt)


def energy_residual(prob, u, dudt):
    """Analysis/conservation.jl:154-167: Σ_k u_kᵀ M_k dudt_k per variable."""
    out = np.zeros(u.shape[1])
    for k in range(u.shape[0]):
        M = mass_matrix(prob, k)
        out += np.einsum("ep,pq,eq->e", u[k], M, dudt[k])
    return out


def entropy_residual(prob, u, dudt):
    """Analysis/conservation.jl:169-190: Σ_k (P_k w(V u_k))ᵀ M_k dudt_k."""
    law = prob["law"]
    V = prob["V"]
    tot = 0.0
    for k in range(u.shape[0]):
        M = mass_matrix(prob, k)
        Minv = mass_matrix_inverse(prob, k)
        u_q = np.einsum("qp,ep->qe", V, u[k])
        w_q = conservative_to_entropy(law, u_q)
        P = Minv @ (V.T * (prob["W"] * prob["J_q"][k])[None, :])
        tot += np.einsum("pe,pq,eq->", P @ w_q, M, dudt[k])
    return tot


def l2_error(prob, u, exact_q):
    """Analysis/error.jl:58-91 with the default (volume) quadrature.  exact_q: (N_e,N_q,N_c)."""
    u_q = np.einsum("qp,kep->kqe", prob["V"], u)
    WJ = prob["W"][None, :] * prob["J_q"]
    return np.sqrt(np.einsum("kq,kqe->e", WJ, (exact_q - u_q) ** 2))


def l2_error_quadrature(prob, u, exact, xyzq, t, volq_to_err, w_err):
    """Analysis/error.jl:13-91 with a separate error quadrature: ``volq_to_err`` (N_err, N_q) =
    V_modes_to_errq P_volq_to_modes interpolates volume-node values (coordinates, Jacobian and,
    through V, the solution) to the error-quadrature nodes with weights ``w_err``.
    ``xyzq``: d arrays (N_e, N_q); ``exact(x..., t)`` -> tuple of N_c arrays."""
    V_err = volq_to_err @ prob["V"]
    err = np.zeros(u.shape[1])
    for k in range(u.shape[0]):
        x_err = [volq_to_err @ x[k] for x in xyzq]
        ue = np.stack(exact(*x_err, t), axis=-1)
        ua = np.einsum("qp,ep->qe", V_err, u[k])
        wj = w_err * (volq_to_err @ prob["J_q"][k])
        err += np.einsum("q,qe->e", wj, (ue - ua) ** 2)
    return np.sqrt(err)


def project_initial_data(prob, u_q):
    """Solvers.jl:389-428: nodal -> copy; modal -> per-element L2 projection
    (VᵀWJV) \\ Vᵀ WJ u_q.  u_q: (N_e, N_q, N_c)."""
    if prob.get("V_is_identity", False):
        return np.ascontiguousarray(u_q.transpose(0, 2, 1))
    V = prob["V"]
    WJ = prob["W"][None, :] * prob["J_q"]
    M = np.einsum("qa,kq,qb->kab", V, WJ, V)
    rhs = np.einsum("qp,kq,kqe->kpe", V, WJ, u_q)
    return np.ascontiguousarray(np.linalg.solve(M, rhs).transpose(0, 2, 1))
