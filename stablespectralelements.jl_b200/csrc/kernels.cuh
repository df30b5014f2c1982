// Element kernels for the semi-discrete residual (FP64, sm_100a).
//
// Design: one CTA owns E whole elements; the (p+1)^d nodal block of each element, its
// geometric factors and all intermediate element-local vectors are staged in shared memory,
// and every reference-element operator is applied from small read-only tables that stay
// L1/L2 resident (they are identical for all elements):
//   * V / V^T  : sum-factorised warped tensor product (warped_product_{2d,3d}.jl), identity,
//                or dense;
//   * R, R^T, D_eta, D_eta^T, S (flux-differencing adjacency), C = R^T B : CSR tables.  The
//                Kronecker structure of the tensor-product operators *is* their sparsity, so
//                a row of D_eta^m touches the (p+1) nodes of one tensor line, a row of R the
//                (p+1) [(p+1)^2 on the collapsed face] nodes it interpolates from.
// Global arrays keep the reference's (Julia, column-major) layout: each element's block is
// contiguous, so the per-element loads below are fully coalesced.
#pragma once
#include "physics.cuh"

namespace sse {

struct Tables {
  int dim, N_p, N_q, N_f, N_c, nfaces, npf, n1;
  int v_kind, mass_kind, has_Minv, r_is_selection;
  int nnzRt;
  const double *Vd, *VdT;              // dense V [N_q][N_p] and its transpose [N_p][N_q]
  const double *wA, *wB, *wC;          // warped-product tables
  const double *wCt;                   // 3-D: C re-ordered [mode][a3]; pair / mode tables (vmap3.cuh)
  const int *pairtab, *modetab;
  const double *wK;                    // 3-D: fused middle stage of V V^T (vmap3.cuh)
  const int *sig;                      // sigma_i (n1^d), -1 where unused
  const int *R_rp, *R_ci;  const double *R_v;  const int *R_slot;   // R rows -> Rt entry ids
  // arithmetic-progression descriptor of R row j (tensor-product elements; else NULL):
  // columns start + q*stride, q < count; ELL slot k of this facet node in the rows of R^T.
  // packed start | stride << 10 | count << 20 | k << 27
  const int *R_desc;
  // separable collapsed-face rows (tensor-product simplices): a row of R with N1^2 entries is
  // rank one, R[j][a2*N1+a3] = R_E[j][a2] * R_r3[a3]; rows with the same start share the inner
  // contraction.  R_ng = number of distinct starts (0: not separable), R_gstart[g] = start,
  // R_grp[j] = group of row j.
  int R_ng;
  const int *R_gstart, *R_grp;
  const double *R_E;
  double R_r3[8];
  const int *Rt_rp, *Rt_ci; const double *Rt_v; const double *C_v;  // R^T rows; C = R^T B
  const int *S_rp, *S_ci;  const double *S_v;                       // S_v[e*dim + m]
  const int *D_rp[3], *D_ci[3];  const double *D_v[3];
  const int *Dt_rp[3], *Dt_ci[3]; const double *Dt_v[3];
  const double *W, *B, *n_ref, *Gref, *Minv;   // Gref[i][m][l] = Lambda_ref[i,m,l]/J_ref[i]
};

struct Geo {
  long long N_e;      // one past the last element this launch may touch
  long long k_begin;  // first element of this launch (element-range launches)
  const double *J_q, *L_q, *J_f, *nJf;
  const int *toff;      // trace offset of the exterior node: (k'*N_c)*N_f + j'
  const int *mapP;      // raw linear index j' + N_f*k'
  const double *VOL, *FAC, *Minv_e;
  int pf_dist;          // elements ahead whose inputs a CTA prefetches into L2 (0 = off)
};

struct RK {
  int mode;             // 0: write dudt; 1: k = a*k + dt*R, u += b*k
  double a, b, dt;
  double *k, *u;
};

enum { V_IDENTITY = 0, V_WARPED = 1, V_DENSE = 2 };
enum { MASS_DIAGONAL = 0, MASS_WEIGHT_ADJUSTED = 1, MASS_CHOLESKY = 2 };

#define SSE_LOOP(idx, total) for (int idx = threadIdx.x; idx < (total); idx += blockDim.x)

// sum_i a[i] x[i], both in shared memory: four independent partial sums, unrolled by four
// (k_physical with the operators staged in shared memory: a thread per output row)
__device__ __forceinline__ double smem_dot(const double* __restrict__ a, const double* __restrict__ x,
                                           int n) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    s0 = fma(a[i], x[i], s0);
    s1 = fma(a[i + 1], x[i + 1], s1);
    s2 = fma(a[i + 2], x[i + 2], s2);
    s3 = fma(a[i + 3], x[i + 3], s3);
  }
  for (; i < n; ++i) s0 = fma(a[i], x[i], s0);
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ double halfwarp_sum(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// One output row of k_physical streamed from GLOBAL memory by a half-warp: NSEG row segments
// a[s] (N_p x n row-major operators, n doubles each, contiguous) times vectors x[s] in shared
// memory.  The 16 lanes read 16 consecutive doubles -- one full 128-byte line per segment step --
// and the NSEG loads of a step are independent, so a lane keeps NSEG (x 2 rows, see the caller)
// 8-byte loads in flight.  ok = false: a padding row (no loads, contributes 0).
template <int NSEG>
__device__ __forceinline__ double stream_row(const double* const* a, const double* const* x,
                                             int n, int l16, bool ok, double acc) {
  for (int i = l16; i < n; i += 16) {
    double v[NSEG];
#pragma unroll
    for (int s_ = 0; s_ < NSEG; ++s_) v[s_] = ok ? __ldcs(a[s_] + i) : 0.0;
#pragma unroll
    for (int s_ = 0; s_ < NSEG; ++s_) acc = fma(v[s_], x[s_][i], acc);
  }
  return acc;
}

// ------------------------------------------------------------------ V and V^T
// src: [E][NC][N_p] -> dst: [E][NC][N_q]; tmp holds 2*E*NC*n1^DIM doubles (warped only).
template <int DIM>
__device__ void apply_V(const Tables& T, int E, int NC, const double* __restrict__ src,
                        double* __restrict__ dst, double* __restrict__ tmp) {
  const int Np = T.N_p, Nq = T.N_q, n = T.n1;
  if (T.v_kind == V_IDENTITY) {
    SSE_LOOP(idx, E * NC * Nq) dst[idx] = src[idx];
    __syncthreads();
    return;
  }
  if (T.v_kind == V_DENSE) {
    SSE_LOOP(idx, E * NC * Nq) {
      int i = idx % Nq, ec = idx / Nq;
      const double* s = src + ec * Np;
      const double* v = T.Vd + i * Np;
      double acc = 0.0;
      for (int p = 0; p < Np; ++p) acc = fma(v[p], s[p], acc);
      dst[idx] = acc;
    }
    __syncthreads();
    return;
  }
  if constexpr (DIM == 2) {
    double* Z = tmp;  // [E*NC][b1][a2]
    SSE_LOOP(idx, E * NC * n * n) {
      int a2 = idx % n, b1 = (idx / n) % n, ec = idx / (n * n);
      const double* s = src + ec * Np;
      double acc = 0.0;
      for (int b2 = 0; b2 < n - b1; ++b2)
        acc = fma(T.wB[(a2 * n + b1) * n + b2], s[T.sig[b1 * n + b2]], acc);
      Z[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * n * n) {
      int a2 = idx % n, a1 = (idx / n) % n, ec = idx / (n * n);
      const double* z = Z + ec * n * n;
      double acc = 0.0;
      for (int b1 = 0; b1 < n; ++b1) acc = fma(T.wA[a1 * n + b1], z[b1 * n + a2], acc);
      dst[ec * Nq + a1 * n + a2] = acc;
    }
    __syncthreads();
  } else if constexpr (DIM == 3) {
    const int n3 = n * n * n;
    double* Z = tmp;               // [E*NC][b1][b2][a3]
    double* Wt = tmp + E * NC * n3;  // [E*NC][b1][a2][a3]
    SSE_LOOP(idx, E * NC * n3) {
      int a3 = idx % n, b2 = (idx / n) % n, b1 = (idx / (n * n)) % n, ec = idx / n3;
      double acc = 0.0;
      if (b2 < n - b1) {
        const double* s = src + ec * Np;
        const double* c = T.wC + ((a3 * n + b1) * n + b2) * n;
        const int* sg = T.sig + (b1 * n + b2) * n;
        for (int b3 = 0; b3 < n - b1 - b2; ++b3) acc = fma(c[b3], s[sg[b3]], acc);
      }
      Z[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * n3) {
      int a3 = idx % n, a2 = (idx / n) % n, b1 = (idx / (n * n)) % n, ec = idx / n3;
      const double* z = Z + ec * n3 + b1 * n * n + a3;
      const double* b = T.wB + (a2 * n + b1) * n;
      double acc = 0.0;
      for (int b2 = 0; b2 < n - b1; ++b2) acc = fma(b[b2], z[b2 * n], acc);
      Wt[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * n3) {
      int a23 = idx % (n * n), a1 = (idx / (n * n)) % n, ec = idx / n3;
      const double* w = Wt + ec * n3 + a23;
      double acc = 0.0;
      for (int b1 = 0; b1 < n; ++b1) acc = fma(T.wA[a1 * n + b1], w[b1 * n * n], acc);
      dst[ec * Nq + a1 * n * n + a23] = acc;
    }
    __syncthreads();
  }
}

// src: [E][NC][N_q] -> dst: [E][NC][N_p]
template <int DIM>
__device__ void apply_Vt(const Tables& T, int E, int NC, const double* __restrict__ src,
                         double* __restrict__ dst, double* __restrict__ tmp) {
  const int Np = T.N_p, Nq = T.N_q, n = T.n1;
  if (T.v_kind == V_IDENTITY) {
    SSE_LOOP(idx, E * NC * Nq) dst[idx] = src[idx];
    __syncthreads();
    return;
  }
  if (T.v_kind == V_DENSE) {
    SSE_LOOP(idx, E * NC * Np) {
      int p = idx % Np, ec = idx / Np;
      const double* s = src + ec * Nq;
      const double* v = T.VdT + p * Nq;
      double acc = 0.0;
      for (int i = 0; i < Nq; ++i) acc = fma(v[i], s[i], acc);
      dst[idx] = acc;
    }
    __syncthreads();
    return;
  }
  if constexpr (DIM == 2) {
    double* Z = tmp;  // [E*NC][b1][a2]
    SSE_LOOP(idx, E * NC * n * n) {
      int a2 = idx % n, b1 = (idx / n) % n, ec = idx / (n * n);
      const double* s = src + ec * Nq + a2;
      double acc = 0.0;
      for (int a1 = 0; a1 < n; ++a1) acc = fma(T.wA[a1 * n + b1], s[a1 * n], acc);
      Z[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * n * n) {
      int b2 = idx % n, b1 = (idx / n) % n, ec = idx / (n * n);
      if (b2 < n - b1) {
        const double* z = Z + ec * n * n + b1 * n;
        double acc = 0.0;
        for (int a2 = 0; a2 < n; ++a2) acc = fma(T.wB[(a2 * n + b1) * n + b2], z[a2], acc);
        dst[ec * Np + T.sig[b1 * n + b2]] = acc;
      }
    }
    __syncthreads();
  } else if constexpr (DIM == 3) {
    const int n3 = n * n * n;
    double* Wt = tmp;                // [E*NC][b1][a2][a3]
    double* Z = tmp + E * NC * n3;   // [E*NC][b1][b2][a3]
    SSE_LOOP(idx, E * NC * n3) {
      int a23 = idx % (n * n), b1 = (idx / (n * n)) % n, ec = idx / n3;
      const double* s = src + ec * Nq + a23;
      double acc = 0.0;
      for (int a1 = 0; a1 < n; ++a1) acc = fma(T.wA[a1 * n + b1], s[a1 * n * n], acc);
      Wt[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * n3) {
      int a3 = idx % n, b2 = (idx / n) % n, b1 = (idx / (n * n)) % n, ec = idx / n3;
      double acc = 0.0;
      if (b2 < n - b1) {
        const double* w = Wt + ec * n3 + b1 * n * n + a3;
        for (int a2 = 0; a2 < n; ++a2) acc = fma(T.wB[(a2 * n + b1) * n + b2], w[a2 * n], acc);
      }
      Z[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * n3) {
      int b3 = idx % n, b2 = (idx / n) % n, b1 = (idx / (n * n)) % n, ec = idx / n3;
      if (b2 < n - b1 && b3 < n - b1 - b2) {
        const double* z = Z + ec * n3 + (b1 * n + b2) * n;
        double acc = 0.0;
        for (int a3 = 0; a3 < n; ++a3)
          acc = fma(T.wC[((a3 * n + b1) * n + b2) * n + b3], z[a3], acc);
        dst[ec * Np + T.sig[(b1 * n + b2) * n + b3]] = acc;
      }
    }
    __syncthreads();
  }
}

// dst[E][NC][N_f] = R src[E][NC][N_q]
__device__ inline void apply_R(const Tables& T, int E, int NC, const double* __restrict__ src,
                               double* __restrict__ dst) {
  const int Nq = T.N_q, Nf = T.N_f;
  SSE_LOOP(idx, E * NC * Nf) {
    int j = idx % Nf, ec = idx / Nf;
    const double* s = src + ec * Nq;
    double acc = 0.0;
    for (int e = T.R_rp[j]; e < T.R_rp[j + 1]; ++e) acc = fma(T.R_v[e], s[T.R_ci[e]], acc);
    dst[idx] = acc;
  }
  __syncthreads();
}

// In-place mass-matrix solve on rhs[E][NC][N_p] (mass_matrix.jl:169-196).
// q: scratch [E][NC][N_q]; tmp: warped scratch.
template <int DIM>
__device__ void mass_solve(const Tables& T, const Geo& G, long long k0, int E, int NC,
                           double* __restrict__ rhs, double* __restrict__ q,
                           double* __restrict__ tmp) {
  const int Np = T.N_p, Nq = T.N_q;
  if (T.mass_kind == MASS_DIAGONAL) {
    SSE_LOOP(idx, E * NC * Np) {
      int i = idx % Np, e = idx / (Np * NC);
      long long k = k0 + e;
      if (k < G.N_e) rhs[idx] = rhs[idx] / (T.W[i] * G.J_q[k * Nq + i]);
    }
    __syncthreads();
    return;
  }
  if (T.mass_kind == MASS_CHOLESKY) {
    SSE_LOOP(idx, E * NC * Np) {
      int p = idx % Np, ec = idx / Np, e = ec / NC;
      long long k = k0 + e;
      double acc = 0.0;
      if (k < G.N_e) {
        const double* Mi = G.Minv_e + (k * Np + p) * (long long)Np;  // symmetric
        const double* s = rhs + ec * Np;
        for (int r = 0; r < Np; ++r) acc = fma(Mi[r], s[r], acc);
      }
      q[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * Np) rhs[idx] = q[idx];
    __syncthreads();
    return;
  }
  // weight-adjusted: rhs <- M^-1 V^T (W/J) V M^-1 rhs
  if (T.has_Minv) {
    SSE_LOOP(idx, E * NC * Np) {
      int p = idx % Np, ec = idx / Np;
      const double* s = rhs + ec * Np;
      double acc = 0.0;
      for (int r = 0; r < Np; ++r) acc = fma(T.Minv[p * Np + r], s[r], acc);
      q[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * Np) rhs[idx] = q[idx];
    __syncthreads();
  }
  apply_V<DIM>(T, E, NC, rhs, q, tmp);
  SSE_LOOP(idx, E * NC * Nq) {
    int i = idx % Nq, e = idx / (Nq * NC);
    long long k = k0 + e;
    if (k < G.N_e) q[idx] *= T.W[i] / G.J_q[k * Nq + i];
  }
  __syncthreads();
  apply_Vt<DIM>(T, E, NC, q, rhs, tmp);
  if (T.has_Minv) {
    SSE_LOOP(idx, E * NC * Np) {
      int p = idx % Np, ec = idx / Np;
      const double* s = rhs + ec * Np;
      double acc = 0.0;
      for (int r = 0; r < Np; ++r) acc = fma(T.Minv[p * Np + r], s[r], acc);
      q[idx] = acc;
    }
    __syncthreads();
    SSE_LOOP(idx, E * NC * Np) rhs[idx] = q[idx];
    __syncthreads();
  }
}

// Epilogue: write dudt or apply the fused low-storage RK update.
__device__ inline void store_result(const Tables& T, const Geo& G, const RK& rk, long long k0,
                                    int E, int NC, const double* __restrict__ res,
                                    double* __restrict__ dudt) {
  const int blk = NC * T.N_p;
  SSE_LOOP(idx, E * blk) {
    long long k = k0 + idx / blk;
    if (k < G.N_e) {
      long long g = k0 * blk + idx;
      if (rk.mode == 0) {
        dudt[g] = res[idx];
      } else {
        double kk = rk.a * rk.k[g] + rk.dt * res[idx];
        rk.k[g] = kk;
        rk.u[g] += rk.b * kk;
      }
    }
  }
}

// =========================================================================== loop A
// nodal_values! / entropy_projection! (standard_form_first_order.jl:1-14,
// flux_differencing_form.jl:171-292).  proj: 0 = u_q = V u, u_f = R u_q;
// 1 = nodal scheme with general R (entropy variables extrapolated); 2 = modal projection.
// shared: bufP[E*NC*N_p] | bufQ[E*NC*N_q] | bufQ2[E*NC*N_q] | bufF[E*NC*N_f] | tmp
template <int DIM, int LAW>
__global__ void __launch_bounds__(256)
k_nodal_values(Tables T, Geo G, Phys P, const double* __restrict__ u, double* __restrict__ u_q,
               double* __restrict__ u_f, int E, int proj) {
  constexpr int NC = LawTraits<DIM, LAW>::NC;
  SSE_SHARED(sm);
  const int Np = T.N_p, Nq = T.N_q, Nf = T.N_f;
  double* bufP = sm;
  double* bufQ = bufP + E * NC * Np;
  double* bufQ2 = bufQ + E * NC * Nq;
  double* bufF = bufQ2 + E * NC * Nq;
  double* tmp = bufF + E * NC * Nf;
  const long long k0 = G.k_begin + (long long)blockIdx.x * E;
  const int Ev = (int)min((long long)E, G.N_e - k0);   // valid elements in this CTA

  SSE_LOOP(idx, E * NC * Np) bufP[idx] = (idx < Ev * NC * Np) ? u[k0 * NC * Np + idx] : 1.0;
  __syncthreads();
  apply_V<DIM>(T, E, NC, bufP, bufQ, tmp);

  if (proj == 0) {
    apply_R(T, E, NC, bufQ, bufF);
    SSE_LOOP(idx, Ev * NC * Nq) u_q[k0 * NC * Nq + idx] = bufQ[idx];
    SSE_LOOP(idx, Ev * NC * Nf) u_f[k0 * NC * Nf + idx] = bufF[idx];
    return;
  }
  // entropy variables at the volume nodes
  SSE_LOOP(idx, E * Nq) {
    int i = idx % Nq, e = idx / Nq;
    double uu[NC], w[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) uu[c] = bufQ[(e * NC + c) * Nq + i];
    cons_to_entropy<DIM, LAW>(P, uu, w);
    double sc = 1.0;
    if (proj == 2) {
      long long k = min(k0 + e, G.N_e - 1);
      sc = T.W[i] * G.J_q[k * Nq + i];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) bufQ2[(e * NC + c) * Nq + i] = w[c] * sc;
  }
  __syncthreads();
  if (proj == 2) {
    apply_Vt<DIM>(T, E, NC, bufQ2, bufP, tmp);
    mass_solve<DIM>(T, G, k0, E, NC, bufP, bufQ2, tmp);
    apply_V<DIM>(T, E, NC, bufP, bufQ2, tmp);
  }
  apply_R(T, E, NC, bufQ2, bufF);
  if (proj == 2) {
    SSE_LOOP(idx, Ev * Nq) {
      int i = idx % Nq, e = idx / Nq;
      double w[NC], uu[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) w[c] = bufQ2[(e * NC + c) * Nq + i];
      entropy_to_cons<DIM, LAW>(P, w, uu);
#pragma unroll
      for (int c = 0; c < NC; ++c) u_q[((k0 + e) * NC + c) * Nq + i] = uu[c];
    }
  } else {
    SSE_LOOP(idx, Ev * NC * Nq) u_q[k0 * NC * Nq + idx] = bufQ[idx];
  }
  SSE_LOOP(idx, Ev * Nf) {
    int j = idx % Nf, e = idx / Nf;
    double w[NC], uu[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) w[c] = bufF[(e * NC + c) * Nf + j];
    entropy_to_cons<DIM, LAW>(P, w, uu);
#pragma unroll
    for (int c = 0; c < NC; ++c) u_f[((k0 + e) * NC + c) * Nf + j] = uu[c];
  }
}

// Interface flux at one facet node: f* = F(u-,u+).n [+ halfλ a (u- - u+)]
// (ConservationLaws.jl:75-128).  Returns the exterior-trace-based flux in fs, the interior
// state in sl.
// interface flux from already-loaded interior/exterior conservative traces
template <int DIM, int LAW>
__device__ __forceinline__ void interface_flux_vals(const Phys& P, int two_point,
                                                    const double* um, const double* up,
                                                    const double* nf, double* sl, double* fs) {
  constexpr int NC = LawTraits<DIM, LAW>::NC;
  constexpr int NS = LawTraits<DIM, LAW>::NS;
  double sr[NS];
  cons_to_state<DIM, LAW>(P, um, sl);
  cons_to_state<DIM, LAW>(P, up, sr);
  two_point_flux_c<DIM, LAW>(P, two_point, sl, sr, nf, fs);
  if (P.inviscid == 0) {
    double a = P.half_lambda * wave_speed<DIM, LAW>(P, sl, sr, nf);
#pragma unroll
    for (int c = 0; c < NC; ++c) fs[c] = fma(a, um[c] - up[c], fs[c]);
  }
}

template <int DIM, int LAW>
__device__ __forceinline__ void interface_flux(const Phys& P, int two_point,
                                               const double* __restrict__ u_f, long long own,
                                               long long ext, int stride, const double* nf,
                                               double* sl, double* fs) {
  constexpr int NC = LawTraits<DIM, LAW>::NC;
  double um[NC], up[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    um[c] = __ldcg(u_f + own + (long long)c * stride);
    up[c] = __ldcg(u_f + ext + (long long)c * stride);
  }
  interface_flux_vals<DIM, LAW>(P, two_point, um, up, nf, sl, fs);
}

// ==================================================== loop B, flux-differencing form
// time_derivative! of flux_differencing_form.jl:294-347 with flux_difference! (:1-75) and
// facet_correction! (:78-168).
// shared (per CTA): sP[E][NS][N_q] | sL[E][D*D][N_q] | sPf[E][NS][N_f] | sFf[E][NC][N_f] |
//                   sNf[E][N_f][D] | sR[E][NC][N_q] | sM[E][NC][N_p] | sD[...]
template <int DIM, int LAW>
__global__ void __launch_bounds__(256)
k_fluxdiff(Tables T, Geo G, Phys P, RK rk, const double* __restrict__ u_q,
           const double* __restrict__ u_f, double* __restrict__ dudt, int E) {
  constexpr int NC = LawTraits<DIM, LAW>::NC;
  constexpr int NS = LawTraits<DIM, LAW>::NS;
  constexpr int DD = DIM * DIM;
  SSE_SHARED(sm);
  const int Np = T.N_p, Nq = T.N_q, Nf = T.N_f;
  double* sP = sm;
  double* sL = sP + E * NS * Nq;
  double* sPf = sL + E * DD * Nq;
  double* sFf = sPf + E * NS * Nf;
  double* sNf = sFf + E * NC * Nf;
  double* sR = sNf + E * Nf * DIM;
  double* sM = sR + E * NC * Nq;
  double* sD = sM + E * NC * Np;
  const long long k0 = G.k_begin + (long long)blockIdx.x * E;
  const int Ev = (int)min((long long)E, G.N_e - k0);

  // ---- phase 0: stage nodal states and metric terms
  SSE_LOOP(idx, E * Nq) {
    int i = idx % Nq, e = idx / Nq;
    long long k = min(k0 + e, G.N_e - 1);
    double uu[NC], s[NS];
#pragma unroll
    for (int c = 0; c < NC; ++c) uu[c] = u_q[(k * NC + c) * Nq + i];
    cons_to_state<DIM, LAW>(P, uu, s);
#pragma unroll
    for (int c = 0; c < NS; ++c) sP[(e * NS + c) * Nq + i] = s[c];
  }
  SSE_LOOP(idx, E * DD * Nq) {
    int e = idx / (DD * Nq);
    long long k = min(k0 + e, G.N_e - 1);
    sL[idx] = G.L_q[k * DD * Nq + (idx - e * DD * Nq)];
  }
  // ---- phase 1: interface numerical flux at the facet nodes
  SSE_LOOP(idx, E * Nf) {
    int j = idx % Nf, e = idx / Nf;
    long long k = min(k0 + e, G.N_e - 1);
    long long gj = k * Nf + j;
    double nJ[DIM], nf[DIM], sl[NS], fs[NC];
    double Jf = G.J_f[gj];
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      nJ[m] = G.nJf[gj * DIM + m];
      nf[m] = nJ[m] / Jf;
    }
    interface_flux<DIM, LAW>(P, P.two_point, u_f, k * NC * Nf + j, G.toff[gj], Nf, nf, sl, fs);
    double bj = T.B[j] * Jf;
#pragma unroll
    for (int c = 0; c < NC; ++c) sFf[(e * NC + c) * Nf + j] = bj * fs[c];
#pragma unroll
    for (int c = 0; c < NS; ++c) sPf[(e * NS + c) * Nf + j] = sl[c];
#pragma unroll
    for (int m = 0; m < DIM; ++m) sNf[(e * Nf + j) * DIM + m] = 0.5 * nJ[m];
  }
  __syncthreads();

  // ---- phase 2: volume flux differencing + facet correction, one thread per volume node
  SSE_LOOP(idx, E * Nq) {
    int i = idx % Nq, e = idx / Nq;
    const double* Pe = sP + e * NS * Nq;
    const double* Le = sL + e * DD * Nq;
    double si[NS], Li[DD], r[NC];
#pragma unroll
    for (int c = 0; c < NS; ++c) si[c] = Pe[c * Nq + i];
#pragma unroll
    for (int c = 0; c < DD; ++c) Li[c] = Le[c * Nq + i];   // Li[m + DIM*n] = Λ[i,m,n]
#pragma unroll
    for (int c = 0; c < NC; ++c) r[c] = 0.0;
    for (int en = T.S_rp[i]; en < T.S_rp[i + 1]; ++en) {
      int j = T.S_ci[en];
      double cv[DIM], sj[NS], f[NC];
#pragma unroll
      for (int n = 0; n < DIM; ++n) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < DIM; ++m)
          acc = fma(T.S_v[en * DIM + m], Li[m + DIM * n] + Le[(m + DIM * n) * Nq + j], acc);
        cv[n] = acc;
      }
#pragma unroll
      for (int c = 0; c < NS; ++c) sj[c] = Pe[c * Nq + j];
      two_point_flux_c<DIM, LAW>(P, P.two_point, si, sj, cv, f);
#pragma unroll
      for (int c = 0; c < NC; ++c) r[c] -= f[c];
    }
    if (!T.r_is_selection) {
      const double* Pfe = sPf + e * NS * Nf;
      for (int en = T.Rt_rp[i]; en < T.Rt_rp[i + 1]; ++en) {
        int j = T.Rt_ci[en];
        int fc = j / T.npf;
        double nJ[DIM], sj[NS], f[NC];
#pragma unroll
        for (int n = 0; n < DIM; ++n) {
          double acc = 0.0;
#pragma unroll
          for (int m = 0; m < DIM; ++m) acc = fma(Li[m + DIM * n], T.n_ref[fc * DIM + m], acc);
          nJ[n] = sNf[(e * Nf + j) * DIM + n] + 0.5 * acc;
        }
#pragma unroll
        for (int c = 0; c < NS; ++c) sj[c] = Pfe[c * Nf + j];
        two_point_flux_c<DIM, LAW>(P, P.two_point, si, sj, nJ, f);
        double cij = T.C_v[en];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double dlt = cij * f[c];
          r[c] -= dlt;
          sD[(e * T.nnzRt + en) * NC + c] = dlt;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) sR[(e * NC + c) * Nq + i] = r[c];
  }
  __syncthreads();

  // ---- phase 3: f_f -= column sums of the facet-correction terms
  if (!T.r_is_selection) {
    SSE_LOOP(idx, E * NC * Nf) {
      int j = idx % Nf, c = (idx / Nf) % NC, e = idx / (Nf * NC);
      double acc = 0.0;
      for (int en = T.R_rp[j]; en < T.R_rp[j + 1]; ++en)
        acc += sD[(e * T.nnzRt + T.R_slot[en]) * NC + c];
      sFf[idx] -= acc;
    }
    __syncthreads();
  }
  // ---- phase 4: r_q -= R^T f_f
  SSE_LOOP(idx, E * NC * Nq) {
    int i = idx % Nq, ec = idx / Nq;
    const double* ff = sFf + ec * Nf;
    double acc = 0.0;
    for (int en = T.Rt_rp[i]; en < T.Rt_rp[i + 1]; ++en) acc = fma(T.Rt_v[en], ff[T.Rt_ci[en]], acc);
    sR[idx] -= acc;
  }
  __syncthreads();
  // ---- phase 5: dudt = M^-1 V^T r_q   (sD is free now: V scratch; sP region: q scratch)
  apply_Vt<DIM>(T, E, NC, sR, sM, sD);
  mass_solve<DIM>(T, G, k0, E, NC, sM, sR, sD);
  store_result(T, G, rk, k0, E, NC, sM, dudt);
  (void)Ev;
}

// ============================================ loop B, standard form, reference operators
// time_derivative! of standard_form_first_order.jl:16-63 (skew-symmetric split form):
//   r = Σ_m [ D_m^T g_m − Σ_n hWΛ[m,n] ∘ (D_m f_n) ],  g_m = Σ_n hWΛ[m,n] ∘ f_n,
//   hWΛ[m,n] = ½ W Λ_η[m,n],  Λ_η[i,m,n] = Σ_l Λ_ref[i,m,l] Λ_q[i,l,n] / J_ref[i].
// shared: sU[E][NS][N_q] | sH[E][D*D][N_q] | sFq[E][D][NC][N_q] | sG[E][D][NC][N_q] |
//         sFf[E][NC][N_f] | sR[E][NC][N_q] | sM[E][NC][N_p] | tmp
template <int DIM, int LAW>
__global__ void __launch_bounds__(256)
k_standard_ref(Tables T, Geo G, Phys P, RK rk, const double* __restrict__ u_q,
               const double* __restrict__ u_f, double* __restrict__ dudt, int E) {
  constexpr int NC = LawTraits<DIM, LAW>::NC;
  constexpr int NS = LawTraits<DIM, LAW>::NS;
  constexpr int DD = DIM * DIM;
  SSE_SHARED(sm);
  const int Np = T.N_p, Nq = T.N_q, Nf = T.N_f;
  double* sH = sm;                              // hWΛ, index (m + DIM*n)
  double* sFq = sH + E * DD * Nq;
  double* sG = sFq + E * DIM * NC * Nq;
  double* sFf = sG + E * DIM * NC * Nq;
  double* sR = sFf + E * NC * Nf;
  double* sM = sR + E * NC * Nq;
  double* tmp = sM + E * NC * Np;
  const long long k0 = G.k_begin + (long long)blockIdx.x * E;

  // ---- phase 0: physical flux at volume nodes, collapsed metrics, g_m
  SSE_LOOP(idx, E * Nq) {
    int i = idx % Nq, e = idx / Nq;
    long long k = min(k0 + e, G.N_e - 1);
    double uu[NC], s[NS], Lq[DD], H[DD];
#pragma unroll
    for (int c = 0; c < NC; ++c) uu[c] = u_q[(k * NC + c) * Nq + i];
    cons_to_state<DIM, LAW>(P, uu, s);
#pragma unroll
    for (int c = 0; c < DD; ++c) Lq[c] = G.L_q[(k * DD + c) * Nq + i];
    double hw = 0.5 * T.W[i];
#pragma unroll
    for (int m = 0; m < DIM; ++m)
#pragma unroll
      for (int n = 0; n < DIM; ++n) {
        double v;
        if (T.Gref) {
          v = 0.0;
#pragma unroll
          for (int l = 0; l < DIM; ++l) v += T.Gref[(i * DIM + m) * DIM + l] * Lq[l + DIM * n];
        } else {
          v = Lq[m + DIM * n];
        }
        H[m + DIM * n] = hw * v;
        sH[(e * DD + m + DIM * n) * Nq + i] = H[m + DIM * n];
      }
    double fq[DIM][NC];
#pragma unroll
    for (int n = 0; n < DIM; ++n) {
      double cdir[DIM];
#pragma unroll
      for (int m = 0; m < DIM; ++m) cdir[m] = (m == n) ? 1.0 : 0.0;
      physical_flux_c<DIM, LAW>(P, s, cdir, fq[n]);
#pragma unroll
      for (int c = 0; c < NC; ++c) sFq[((e * DIM + n) * NC + c) * Nq + i] = fq[n][c];
    }
#pragma unroll
    for (int m = 0; m < DIM; ++m)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double g = 0.0;
#pragma unroll
        for (int n = 0; n < DIM; ++n) g += H[m + DIM * n] * fq[n][c];
        sG[((e * DIM + m) * NC + c) * Nq + i] = g;
      }
  }
  __syncthreads();
  // ---- phase 1: facet nodes: f_f = BJf (f* − Σ_n ½ n_n (R f_n))
  SSE_LOOP(idx, E * Nf) {
    int j = idx % Nf, e = idx / Nf;
    long long k = min(k0 + e, G.N_e - 1);
    long long gj = k * Nf + j;
    double nf[DIM], sl[NS], fs[NC];
    double Jf = G.J_f[gj];
#pragma unroll
    for (int m = 0; m < DIM; ++m) nf[m] = G.nJf[gj * DIM + m] / Jf;
    interface_flux<DIM, LAW>(P, 0, u_f, k * NC * Nf + j, G.toff[gj], Nf, nf, sl, fs);
    for (int en = T.R_rp[j]; en < T.R_rp[j + 1]; ++en) {
      int i = T.R_ci[en];
      double rv = T.R_v[en];
#pragma unroll
      for (int n = 0; n < DIM; ++n) {
        double hn = 0.5 * nf[n] * rv;
#pragma unroll
        for (int c = 0; c < NC; ++c) fs[c] -= hn * sFq[((e * DIM + n) * NC + c) * Nq + i];
      }
    }
    double bj = T.B[j] * Jf;
#pragma unroll
    for (int c = 0; c < NC; ++c) sFf[(e * NC + c) * Nf + j] = bj * fs[c];
  }
  __syncthreads();
  // ---- phase 2: volume + lifting
  SSE_LOOP(idx, E * NC * Nq) {
    int i = idx % Nq, c = (idx / Nq) % NC, e = idx / (Nq * NC);
    double r = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      const double* g = sG + ((e * DIM + m) * NC + c) * Nq;
      for (int en = T.Dt_rp[m][i]; en < T.Dt_rp[m][i + 1]; ++en)
        r = fma(T.Dt_v[m][en], g[T.Dt_ci[m][en]], r);
      for (int en = T.D_rp[m][i]; en < T.D_rp[m][i + 1]; ++en) {
        int j = T.D_ci[m][en];
        double acc = 0.0;
#pragma unroll
        for (int n = 0; n < DIM; ++n)
          acc = fma(sH[(e * DD + m + DIM * n) * Nq + i], sFq[((e * DIM + n) * NC + c) * Nq + j], acc);
        r = fma(-T.D_v[m][en], acc, r);
      }
    }
    const double* ff = sFf + (e * NC + c) * Nf;
    for (int en = T.Rt_rp[i]; en < T.Rt_rp[i + 1]; ++en) r = fma(-T.Rt_v[en], ff[T.Rt_ci[en]], r);
    sR[idx] = r;
  }
  __syncthreads();
  apply_Vt<DIM>(T, E, NC, sR, sM, tmp);
  mass_solve<DIM>(T, G, k0, E, NC, sM, sR, tmp);
  store_result(T, G, rk, k0, E, NC, sM, dudt);
}

// ============================================== physical-operator form (dense per element)
// First-order: time_derivative! of standard_form_first_order.jl:65-94.
// Second-order (BR1): auxiliary_variable! and time_derivative! of
// standard_form_second_order.jl:3-75.  VOL[k][m] is (N_p x N_q) row-major, FAC[k] (N_p x N_f).
// stage 0: q = -(VOL u_q + FAC u*n), q_q = V q, q_f = R q_q     (second order only)
// stage 1: dudt = Σ_m VOL_m f_m + FAC f*
// shared: sU[E][NC][N_q] | sFq[E][D][NC][N_q] | sFn[E][D][NC][N_f] | sP[E][D][NC][N_p] | tmp
template <int DIM, int LAW>
__global__ void __launch_bounds__(256)
k_physical(Tables T, Geo G, Phys P, RK rk, const double* __restrict__ u_q,
           const double* __restrict__ u_f, double* __restrict__ q_q, double* __restrict__ q_f,
           double* __restrict__ dudt, int E, int stage, int second_order) {
  constexpr int NC = LawTraits<DIM, LAW>::NC;
  constexpr int NS = LawTraits<DIM, LAW>::NS;
  SSE_SHARED(sm);
  const int Np = T.N_p, Nq = T.N_q, Nf = T.N_f;
  double* sFq = sm;                              // [E][D][NC][N_q]
  double* sFn = sFq + E * DIM * NC * Nq;         // [E][D][NC][N_f] (stage 0) / [E][NC][N_f]
  double* sP = sFn + E * DIM * NC * Nf;          // [E][D][NC][N_p]
  double* sQ = sP + E * DIM * NC * Np;           // [E][D][NC][N_q]
  double* tmp = sQ + E * DIM * NC * Nq;
  const long long k0 = G.k_begin + (long long)blockIdx.x * E;
  const int Ev = (int)min((long long)E, G.N_e - k0);
  // The per-element operators VOL[k][m] (N_p x N_q, d of them) and FAC[k] (N_p x N_f) are pure
  // streaming data -- every entry is used once -- and 90 % of the kernel's bytes.
  //  * default: they are never staged.  After the fluxes are in shared memory, a half-warp per
  //    output row streams the row from global memory (stream_row): the CTA holds ~1 KB of shared
  //    memory per element, ten CTAs are resident per SM, and the latency-bound phases of one CTA
  //    (small gathers, flux evaluation, barriers) overlap the streaming phases of the others.
  //  * staged (SSE_B200_PHYS_STAGED=1, the measured alternative): the blocks of the CTA's E
  //    consecutive elements are contiguous in global memory, so ONE thread moves each with one
  //    bulk asynchronous copy (cp.async.bulk, the TMA engine) issued first, completing on an
  //    mbarrier every thread polls before the row-times-vector products read shared memory.
  //    Bytes in flight are then bounded by shared memory held for a whole CTA lifetime
  //    (~190 KB per SM / ~11 us = 37 % of the HBM peak, profiles/r2_ab_log.md).  A block whose
  //    source, destination or size is not a multiple of 16 bytes falls back to cp.async.
  const bool staged = (stage & 16) != 0;
  stage &= 15;
  int nd = 1;
#pragma unroll
  for (int m = 0; m < DIM; ++m) nd *= T.n1;
  const int nvol = DIM * Np * Nq, nfac = Np * Nf;
  double* sOp = tmp + E * DIM * NC * (T.v_kind == V_WARPED ? 2 * nd : 0);
  sOp += ((size_t)sOp >> 3) & 1;            // 16-byte aligned: [E][nvol] | [E][nfac]
  double* sVol = sOp;
  double* sFac = sOp + (E * nvol + ((E * nvol) & 1));
  __shared__ unsigned long long op_bar;
  const long long nv = (long long)Ev * nvol, nfc = (long long)Ev * nfac;
  const double* vsrc = G.VOL + k0 * nvol;
  const double* fsrc = G.FAC + k0 * nfac;
  const bool bulk_v = staged && (((size_t)vsrc & 15) == 0) && ((nv & 1) == 0) && SSE_SMEM_ALIGNED16(sVol);
  const bool bulk_f = staged && (((size_t)fsrc & 15) == 0) && ((nfc & 1) == 0) && SSE_SMEM_ALIGNED16(sFac);
  if (bulk_v || bulk_f) {
    if (threadIdx.x == 0) {
      SSE_MBAR_INIT(&op_bar, 1);
      SSE_MBAR_INIT_FENCE();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      SSE_MBAR_EXPECT_TX(&op_bar, (bulk_v ? nv * 8 : 0) + (bulk_f ? nfc * 8 : 0));
      if (bulk_v) SSE_BULK_G2S(sVol, vsrc, nv * 8, &op_bar);
      if (bulk_f) SSE_BULK_G2S(sFac, fsrc, nfc * 8, &op_bar);
    }
  }
  if (staged && !bulk_v)
    for (long long o = threadIdx.x; o < nv; o += blockDim.x) SSE_CP_ASYNC8(sVol + o, vsrc + o);
  if (staged && !bulk_f)
    for (long long o = threadIdx.x; o < nfc; o += blockDim.x) SSE_CP_ASYNC8(sFac + o, fsrc + o);
  // half-warp roles of the streaming row products
  const int hw = threadIdx.x >> 4, nhw = blockDim.x >> 4, l16 = threadIdx.x & 15;

  if (stage == 0) {
    // u_q as the "flux" in every direction; u* n at the facets (BR1: ½(u⁻+u⁺) n)
    SSE_LOOP(idx, E * NC * Nq) {
      int e = idx / (NC * Nq);
      long long k = min(k0 + e, G.N_e - 1);
      sFq[idx] = u_q[k * NC * Nq + (idx - e * NC * Nq)];
    }
    SSE_LOOP(idx, E * Nf) {
      int j = idx % Nf, e = idx / Nf;
      long long k = min(k0 + e, G.N_e - 1);
      long long gj = k * Nf + j;
      double Jf = G.J_f[gj];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double avg = 0.5 * (u_f[(k * NC + c) * Nf + j] + u_f[G.toff[gj] + (long long)c * Nf]);
#pragma unroll
        for (int m = 0; m < DIM; ++m)
          sFn[((e * DIM + m) * NC + c) * Nf + j] = avg * (G.nJf[gj * DIM + m] / Jf);
      }
    }
    SSE_CP_ASYNC_WAIT_ALL();
    __syncthreads();
    if (bulk_v || bulk_f) SSE_MBAR_WAIT(&op_bar, 0);
    const int nrows = Ev * DIM * NC * Np;   // row idx = ((e * DIM + m) * NC + c) * Np + p
    if (staged) {
      SSE_LOOP(idx, nrows) {
        int p = idx % Np, c = (idx / Np) % NC, m = (idx / (Np * NC)) % DIM, e = idx / (Np * NC * DIM);
        sP[idx] = -(smem_dot(sVol + e * nvol + (m * Np + p) * Nq, sFq + (e * NC + c) * Nq, Nq) +
                    smem_dot(sFac + e * nfac + p * Nf, sFn + ((e * DIM + m) * NC + c) * Nf, Nf));
      }
    } else {
      // two rows per half-warp and step (block-uniform trip count: shuffles inside)
      int p0 = hw % Np, g0 = hw / Np, p1 = (hw + nhw) % Np, g1 = (hw + nhw) / Np;   // g = idx / Np
      for (int base = 0; base < nrows; base += 2 * nhw) {
        double acc[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int pp = r ? p1 : p0, g = r ? g1 : g0;
          const bool ok = base + r * nhw + hw < nrows;
          const int c = g % NC, em = g / NC, m = em % DIM, e = ok ? em / DIM : 0;
          const double* const av[1] = {G.VOL + ((k0 + e) * DIM + m) * (long long)Np * Nq + pp * Nq};
          const double* const xv[1] = {sFq + (e * NC + c) * Nq};
          const double* const af[1] = {G.FAC + (k0 + e) * (long long)nfac + pp * Nf};
          const double* const xf[1] = {sFn + ((e * DIM + m) * NC + c) * Nf};
          acc[r] = stream_row<1>(av, xv, Nq, l16, ok, 0.0);
          acc[r] = stream_row<1>(af, xf, Nf, l16, ok, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const double v = halfwarp_sum(acc[r]);
          const int idx = base + r * nhw + hw;
          if (idx < nrows && l16 == 0) sP[idx] = -v;
        }
        p0 += 2 * nhw; while (p0 >= Np) { p0 -= Np; ++g0; }
        p1 += 2 * nhw; while (p1 >= Np) { p1 -= Np; ++g1; }
      }
    }
    __syncthreads();
    apply_V<DIM>(T, E, DIM * NC, sP, sQ, tmp);
    apply_R(T, E, DIM * NC, sQ, sFn);
    SSE_LOOP(idx, Ev * DIM * NC * Nq) q_q[k0 * DIM * NC * Nq + idx] = sQ[idx];
    SSE_LOOP(idx, Ev * DIM * NC * Nf) q_f[k0 * DIM * NC * Nf + idx] = sFn[idx];
    return;
  }

  // stage 1: physical flux (with the viscous part when second order)
  SSE_LOOP(idx, E * Nq) {
    int i = idx % Nq, e = idx / Nq;
    long long k = min(k0 + e, G.N_e - 1);
    double uu[NC], s[NS];
#pragma unroll
    for (int c = 0; c < NC; ++c) uu[c] = u_q[(k * NC + c) * Nq + i];
    cons_to_state<DIM, LAW>(P, uu, s);
#pragma unroll
    for (int n = 0; n < DIM; ++n) {
      double cdir[DIM], f[NC];
#pragma unroll
      for (int m = 0; m < DIM; ++m) cdir[m] = (m == n) ? 1.0 : 0.0;
      physical_flux_c<DIM, LAW>(P, s, cdir, f);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (second_order) f[c] -= P.b * q_q[((k * DIM + n) * NC + c) * Nq + i];
        sFq[((e * DIM + n) * NC + c) * Nq + i] = f[c];
      }
    }
  }
  SSE_LOOP(idx, E * Nf) {
    int j = idx % Nf, e = idx / Nf;
    long long k = min(k0 + e, G.N_e - 1);
    long long gj = k * Nf + j;
    double nf[DIM], sl[NS], fs[NC];
    double Jf = G.J_f[gj];
#pragma unroll
    for (int m = 0; m < DIM; ++m) nf[m] = G.nJf[gj * DIM + m] / Jf;
    interface_flux<DIM, LAW>(P, 0, u_f, k * NC * Nf + j, G.toff[gj], Nf, nf, sl, fs);
    if (second_order) {
      long long kp = G.mapP[gj] / Nf;
      int jp = G.mapP[gj] % Nf;
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int m = 0; m < DIM; ++m) {
          double qa = q_f[((k * DIM + m) * NC + c) * Nf + j] +
                      q_f[((kp * DIM + m) * NC + c) * Nf + jp];
          fs[c] += P.b * (-0.5 * qa) * nf[m];
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) sFn[(e * NC + c) * Nf + j] = fs[c];
  }
  SSE_CP_ASYNC_WAIT_ALL();
  __syncthreads();
  if (bulk_v || bulk_f) SSE_MBAR_WAIT(&op_bar, 0);
  const int nrows = Ev * NC * Np;   // row idx = (e * NC + c) * Np + p
  if (staged) {
    SSE_LOOP(idx, nrows) {
      int p = idx % Np, c = (idx / Np) % NC, e = idx / (Np * NC);
      double acc = smem_dot(sFac + e * nfac + p * Nf, sFn + (e * NC + c) * Nf, Nf);
#pragma unroll
      for (int m = 0; m < DIM; ++m)
        acc += smem_dot(sVol + e * nvol + (m * Np + p) * Nq, sFq + ((e * DIM + m) * NC + c) * Nq, Nq);
      sP[idx] = acc;
    }
  } else {
    int p0 = hw % Np, g0 = hw / Np, p1 = (hw + nhw) % Np, g1 = (hw + nhw) / Np;   // g = e * NC + c
    for (int base = 0; base < nrows; base += 2 * nhw) {
      double acc[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int pp = r ? p1 : p0, g = r ? g1 : g0;
        const bool ok = base + r * nhw + hw < nrows;
        const int c = g % NC, e = ok ? g / NC : 0;
        const double* av[DIM];
        const double* xv[DIM];
#pragma unroll
        for (int m = 0; m < DIM; ++m) {
          av[m] = G.VOL + ((k0 + e) * DIM + m) * (long long)Np * Nq + pp * Nq;
          xv[m] = sFq + ((e * DIM + m) * NC + c) * Nq;
        }
        const double* const af[1] = {G.FAC + (k0 + e) * (long long)nfac + pp * Nf};
        const double* const xf[1] = {sFn + (e * NC + c) * Nf};
        acc[r] = stream_row<DIM>(av, xv, Nq, l16, ok, 0.0);
        acc[r] = stream_row<1>(af, xf, Nf, l16, ok, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const double v = halfwarp_sum(acc[r]);
        const int idx = base + r * nhw + hw;
        if (idx < nrows && l16 == 0) sP[idx] = v;
      }
      p0 += 2 * nhw; while (p0 >= Np) { p0 -= Np; ++g0; }
      p1 += 2 * nhw; while (p1 >= Np) { p1 -= Np; ++g1; }
    }
  }
  __syncthreads();
  store_result(T, G, rk, k0, E, NC, sP, dudt);
}

// ------------------------------------------------------------------------- halo
// send[s*NC + c] = u_f[off(idx[s]) + c*N_f]  (node-major: per-peer segments are contiguous)
static __global__ void k_halo_pack(const double* __restrict__ u_f, const int* __restrict__ off, int n,
                            int NC, int Nf, double* __restrict__ send) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n)
    for (int c = 0; c < NC; ++c) send[(long long)s * NC + c] = u_f[off[s] + (long long)c * Nf];
}

// halo slot h lives at pseudo-element N_e + h / N_f, node h % N_f
static __global__ void k_halo_unpack(double* __restrict__ u_f, const double* __restrict__ recv, int n,
                              int NC, int Nf, long long N_e) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < n) {
    long long kk = N_e + h / Nf;
    int j = h % Nf;
    for (int c = 0; c < NC; ++c) u_f[(kk * NC + c) * Nf + j] = recv[(long long)h * NC + c];
  }
}

// the same for the BR1 auxiliary-variable traces q_f [k][m][c][j]: DIM * NC doubles per node,
// ordered [m][c].  off[s] = k * NC * Nf + j is the u_f offset of the node (sse_halo_setup).
static __global__ void k_halo_pack_aux(const double* __restrict__ q_f, const int* __restrict__ off, int n,
                                int NC, int D, int Nf, double* __restrict__ send) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) {
    const long long k = off[s] / (NC * Nf);
    const int j = off[s] % (NC * Nf);
    for (int mc = 0; mc < D * NC; ++mc)
      send[(long long)s * D * NC + mc] = q_f[(k * D * NC + mc) * Nf + j];
  }
}

static __global__ void k_halo_unpack_aux(double* __restrict__ q_f, const double* __restrict__ recv, int n,
                                  int NC, int D, int Nf, long long N_e) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < n) {
    long long kk = N_e + h / Nf;
    int j = h % Nf;
    for (int mc = 0; mc < D * NC; ++mc)
      q_f[(kk * D * NC + mc) * Nf + j] = recv[(long long)h * D * NC + mc];
  }
}

// out = base + sum_t c[t] * x[t]  (stage states and the final update of a general explicit
// Runge-Kutta step, sse_erk_step); out may alias base
#define SSE_ERK_MAX_TERMS 16
struct LinComb {
  int n;
  double c[SSE_ERK_MAX_TERMS];
  const double* x[SSE_ERK_MAX_TERMS];
};
static __global__ void k_lincomb(double* out, const double* base, LinComb L, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double acc = base[i];
    for (int t = 0; t < L.n; ++t) acc = fma(L.c[t], L.x[t][i], acc);
    out[i] = acc;
  }
}

static __global__ void k_axpy_rk(double* __restrict__ u, double* __restrict__ k, const double* r,
                          double a, double b, double dt, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double kk = a * k[i] + dt * r[i];
    k[i] = kk;
    u[i] += b * kk;
  }
}

}  // namespace sse
