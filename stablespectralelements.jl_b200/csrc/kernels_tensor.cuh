// Compile-time specialised kernels for tensor-product elements (Tri/Tet in collapsed
// coordinates, Quad/Hex): N1 = p + 1 nodes per direction, N_q = N1^DIM, known at compile time so
// that all index arithmetic folds to constants and the contraction loops unroll.
//
// Loop B (flux-differencing form) differs from the generic kernel in two ways:
//  * volume term: the S_m couple only nodes on the same tensor line (the Kronecker structure of
//    D_eta), so each thread (= volume node) evaluates the pairs with its cyclic successors at
//    offsets 1..N1/2 along each direction and hands the result to the partner through shared
//    memory.  Every unordered pair is evaluated exactly once -- d*N1^d*(N1-1)/2 two-point
//    fluxes instead of twice that (flux_differencing_form.jl:10-34 does the same pairwise).
//  * facet correction: C = R^T B has the same number KC of entries in every row (one facet node
//    per non-collapsed face, N1 on the collapsed one), stored in ELL format.
#pragma once
#include "kernels.cuh"
#include "vmap3.cuh"
#include "vmap3b.cuh"

namespace sse {

struct FastTables {
  const double* Sp;    // [DIM][H][DIM][NQ]: S_m[i, partner(i; l, o)]
  const int* Cj;       // [KC][NQ] facet node of ELL slot k | (face index << 16)
  const double* Cv;    // [KC][NQ] C[i, j] = R[j, i] B[j]
  const double* Rv;    // [KC][NQ] R[j, i]
  const double* D1;    // [DIM][N1][N1] 1-D derivative matrices of D_eta (row-major)
  // by value (constant bank, uniform indexing): the reference normals [face][DIM], times one half.
  // The specialised kernels assume the CANONICAL facet layout of the collapsed tensor-product
  // simplices (verified by sse_create, else the generic kernels run): volume node
  // i = (a1, a2[, a3]) meets, in ELL slot k of its row of C = R^T B,
  //   3-D: k=0: face 0 node (a1,a3) | k=1,2: face k node (a2,a3) | k=3+f: face 3 node (a1,f)
  //   2-D: k=0: face 0 node a1      | k=1,2: face k node a2
  // and a facet node of a line face collects the N1 nodes of one tensor line (direction 2 for
  // face 0, direction 1 for faces 1 and 2), a node (a1,f) of the collapsed face the N1^2 nodes
  // (a1,*,*) through slot 3+f.
  double nref[12];
};

__host__ __device__ constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

// facet node of ELL slot kk of volume node (a1, a2, a3) in the canonical layout (see FastTables)
template <int DIM, int N1>
__device__ __forceinline__ int canon_facet_node(int kk, int a1, int a2, int a3) {
  constexpr int NPF = ipow(N1, DIM - 1);
  if constexpr (DIM == 3) {
    if (kk >= 3) return 3 * NPF + a1 * N1 + (kk - 3);
    return kk * NPF + (kk == 0 ? a1 : a2) * N1 + a3;
  } else {
    return kk * NPF + (kk == 0 ? a1 : a2);
  }
}

// the few table pointers the V / V^T applies need, passed BY VALUE to the out-of-line functions
// (a reference to the kernel-parameter struct would force a local-memory copy of all of it).
// c_wA / c_wB (constant-bank copies of the warped-product A and B tables) live in vmap3.cuh.
struct VTab {
  const double *wA, *wB, *wC, *Vd, *VdT;
  const int* sig;
  int N_p, v_kind;
  V3Tab v3;
};
__device__ __forceinline__ VTab vtab(const Tables& T) {
  return VTab{T.wA, T.wB, T.wC, T.Vd, T.VdT, T.sig, T.N_p, T.v_kind,
              V3Tab{T.wC, T.wCt, T.pairtab, T.modetab, T.wK}};
}

// scratch doubles the V / V^T applies need for E elements of NC components
template <int DIM, int N1>
__host__ __device__ constexpr int vtmp_per_column() {
  return DIM == 3 ? V3Dims<N1>::ZS : ipow(N1, DIM);
}

// ---------------------------------------------------------------- sum-factorised V, V^T
// src [E][NC][N_p] -> dst [E][NC][NQ]; every thread carries all NC components of one output.
template <int DIM, int N1, int NC, int E, bool OOL = false>
__device__ __noinline__ void apply_V_t(const VTab T, const double* __restrict__ src,
                                          double* __restrict__ dst, double* __restrict__ tmp) {
  constexpr int NQ = ipow(N1, DIM);
  const int Np = T.N_p;
  if (T.v_kind == V_IDENTITY) {
    SSE_LOOP(idx, E * NC * NQ) dst[idx] = src[idx];
    __syncthreads();
    return;
  }
  if (T.v_kind == V_DENSE) {
    SSE_LOOP(idx, E * NQ) {
      int i = idx % NQ, e = idx / NQ;
      double acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = 0.0;
      for (int p = 0; p < Np; ++p) {
        double v = __ldg(T.Vd + i * Np + p);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, src[(e * NC + c) * Np + p], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) dst[(e * NC + c) * NQ + i] = acc[c];
    }
    __syncthreads();
    return;
  }
  if constexpr (DIM == 2) {
    constexpr int N2 = N1 * N1;
    double* Z = tmp;  // [E][NC][b1][a2]
    SSE_LOOP(idx, E * N2) {
      int a2 = idx % N1, b1 = (idx / N1) % N1, e = idx / N2;
      double acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = 0.0;
#pragma unroll
      for (int b2 = 0; b2 < N1; ++b2)
        if (b2 < N1 - b1) {
          double v = __ldg(T.wB + (a2 * N1 + b1) * N1 + b2);
          int si = __ldg(T.sig + b1 * N1 + b2);
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[c] = fma(v, src[(e * NC + c) * Np + si], acc[c]);
        }
#pragma unroll
      for (int c = 0; c < NC; ++c) Z[(e * NC + c) * N2 + b1 * N1 + a2] = acc[c];
    }
    __syncthreads();
    SSE_LOOP(idx, E * N2) {
      int a2 = idx % N1, a1 = (idx / N1) % N1, e = idx / N2;
      double acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = 0.0;
#pragma unroll
      for (int b1 = 0; b1 < N1; ++b1) {
        double v = __ldg(T.wA + a1 * N1 + b1);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, Z[(e * NC + c) * N2 + b1 * N1 + a2], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) dst[(e * NC + c) * NQ + a1 * N1 + a2] = acc[c];
    }
    __syncthreads();
  } else if constexpr (DIM == 3) {
    // vmap3.cuh: compact Z in tmp, stage B writes dst, stage A runs in place on dst
    v3_stageC<N1, NC, E>(threadIdx.x, 128, T.v3, src, tmp);
    __syncthreads();
    v3_stageB<N1, E * NC, OOL>(threadIdx.x, 128, tmp, dst);
    __syncthreads();
    v3_stageA<N1, E * NC>(threadIdx.x, 128, dst);
    __syncthreads();
  }
}

// The 3-D warped-product V alone (no run-time choice of the V kind, five table pointers instead of
// the whole VTab through the call): loop A's entropy-projection instantiation
template <int N1, int NC, int E, bool OOL>
__device__ __noinline__ void apply_V3_t(const V3Tab T, const double* __restrict__ src,
                                        double* __restrict__ dst, double* __restrict__ tmp) {
  v3_stageC<N1, NC, E>(threadIdx.x, 128, T, src, tmp);
  __syncthreads();
  v3_stageB<N1, E * NC, OOL>(threadIdx.x, 128, tmp, dst);
  __syncthreads();
  v3_stageA<N1, E * NC>(threadIdx.x, 128, dst);
  __syncthreads();
}

// src [E][NC][NQ] -> dst [E][NC][N_p].  The 3-D warped product works in place on src (destroyed).
template <int DIM, int N1, int NC, int E, bool OOL = false>
__device__ __noinline__ void apply_Vt_t(const VTab T, double* __restrict__ src,
                                           double* __restrict__ dst, double* __restrict__ tmp) {
  constexpr int NQ = ipow(N1, DIM);
  const int Np = T.N_p;
  if (T.v_kind == V_IDENTITY) {
    SSE_LOOP(idx, E * NC * NQ) dst[idx] = src[idx];
    __syncthreads();
    return;
  }
  if (T.v_kind == V_DENSE) {
    SSE_LOOP(idx, E * Np) {
      int p = idx % Np, e = idx / Np;
      double acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = 0.0;
      for (int i = 0; i < NQ; ++i) {
        double v = __ldg(T.VdT + p * NQ + i);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, src[(e * NC + c) * NQ + i], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) dst[(e * NC + c) * Np + p] = acc[c];
    }
    __syncthreads();
    return;
  }
  if constexpr (DIM == 2) {
    constexpr int N2 = N1 * N1;
    double* Z = tmp;  // [E][NC][b1][a2]
    SSE_LOOP(idx, E * N2) {
      int a2 = idx % N1, b1 = (idx / N1) % N1, e = idx / N2;
      double acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = 0.0;
#pragma unroll
      for (int a1 = 0; a1 < N1; ++a1) {
        double v = __ldg(T.wA + a1 * N1 + b1);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, src[(e * NC + c) * NQ + a1 * N1 + a2], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) Z[(e * NC + c) * N2 + b1 * N1 + a2] = acc[c];
    }
    __syncthreads();
    SSE_LOOP(idx, E * N2) {
      int b2 = idx % N1, b1 = (idx / N1) % N1, e = idx / N2;
      if (b2 < N1 - b1) {
        double acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = 0.0;
#pragma unroll
        for (int a2 = 0; a2 < N1; ++a2) {
          double v = __ldg(T.wB + (a2 * N1 + b1) * N1 + b2);
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[c] = fma(v, Z[(e * NC + c) * N2 + b1 * N1 + a2], acc[c]);
        }
        int si = __ldg(T.sig + b1 * N1 + b2);
#pragma unroll
        for (int c = 0; c < NC; ++c) dst[(e * NC + c) * Np + si] = acc[c];
      }
    }
    __syncthreads();
  } else if constexpr (DIM == 3) {
    vt3_stageA<N1, E * NC>(threadIdx.x, 128, src);
    __syncthreads();
    vt3_stageB<N1, E * NC, OOL>(threadIdx.x, 128, src, tmp);
    __syncthreads();
    vt3_stageC<N1, E * NC>(threadIdx.x, 128, T.v3, tmp, dst);
    __syncthreads();
  }
}

// x <- V V^T x in place on the nodal block x [E][NC][NQ] (3-D warped product only): the two
// ragged b3-contractions around the modal coefficients are fused into one 5-stage pass
// (A^T, B^T, K, B, A; vmap3.cuh), so the modal intermediate is never formed.
// tmp: 2 * E * NC * ZS doubles.
template <int N1, int NC, int E, bool OOL = false>
__device__ __noinline__ void apply_VtV_t(const VTab T, double* __restrict__ x,
                                         double* __restrict__ tmp) {
  double* Z2 = tmp + E * NC * V3Dims<N1>::ZS;
  vt3_stageA<N1, E * NC>(threadIdx.x, 128, x);
  __syncthreads();
  vt3_stageB<N1, E * NC, OOL>(threadIdx.x, 128, x, tmp);
  __syncthreads();
  vtv3_stageK<N1, NC, E>(threadIdx.x, 128, T.v3, tmp, Z2);
  __syncthreads();
  v3_stageB<N1, E * NC, OOL>(threadIdx.x, 128, Z2, x);
  __syncthreads();
  v3_stageA<N1, E * NC>(threadIdx.x, 128, x);
  __syncthreads();
}

// x <- V V^T diag(sc) V V^T x in place (3-D warped product): the entropy projection with the
// weight-adjusted mass inverse, V M^-1 V^T with M^-1 = V^T (W/J) V.  Nine stages instead of the
// eleven of two apply_VtV_t around a scaling pass: A, the scaling and A^T run on the a1-line in
// the registers of one thread (v3_stageA_scale_At).  sc: [E][NQ].
// the inner three stages B^T, K, B (in place on x [b1][a2][a3]) are one out-of-line body shared
// by both halves: loop A is sensitive to its instruction footprint
#ifndef SSE_NODAL_BTKB_INLINE
#define SSE_NODAL_BTKB_INLINE 0
#endif
#if SSE_NODAL_BTKB_INLINE
#define SSE_BTKB_ATTR __forceinline__
#else
#define SSE_BTKB_ATTR __noinline__
#endif
#ifndef SSE_NODAL_OOL
#define SSE_NODAL_OOL true   // loop A: stage-B bodies out of line (one per b1)
#endif
template <int N1, int NC, int E, bool OOL>
__device__ SSE_BTKB_ATTR void apply_BtKB_t(const V3Tab T, double* __restrict__ x,
                                          double* __restrict__ tmp) {
  double* Z2 = tmp + E * NC * V3Dims<N1>::ZS;
  vt3_stageB<N1, E * NC, OOL>(threadIdx.x, 128, x, tmp);
  __syncthreads();
  vtv3_stageK<N1, NC, E>(threadIdx.x, 128, T, tmp, Z2);
  __syncthreads();
  v3_stageB<N1, E * NC, OOL>(threadIdx.x, 128, Z2, x);
  __syncthreads();
}
template <int N1, int NC, int E, bool OOL = false>
__device__ __forceinline__ void apply_VtV_scaled_VtV_t(const VTab T, double* __restrict__ x,
                                                       double* __restrict__ tmp,
                                                       const double* __restrict__ sc) {
  vt3_stageA<N1, E * NC>(threadIdx.x, 128, x);
  __syncthreads();
  apply_BtKB_t<N1, NC, E, OOL>(T.v3, x, tmp);
  v3_stageA_scale_At<N1, E * NC, NC>(threadIdx.x, 128, x, sc);
  __syncthreads();
  apply_BtKB_t<N1, NC, E, OOL>(T.v3, x, tmp);
  v3_stageA<N1, E * NC>(threadIdx.x, 128, x);   // (one shared out-of-line copy of this stage: +0.8 %)
  __syncthreads();
}

// dst [E][NC][N_f] = R src [E][NC][NQ]; all components per thread.  Rows of R on tensor-product
// elements touch an arithmetic progression of volume nodes (one tensor line, or the N1 x N1
// block behind a node of the collapsed face), so no column indices are loaded.
// scratch doubles apply_R_t needs per column for the separable collapsed-face rows
template <int N1> __host__ __device__ constexpr int rsep_per_column() { return N1 * N1; }

// SEP_ONLY: every row of R is either a tensor line (N1 terms) or a separable collapsed-face row
// (T.R_ng > 0, checked by the launcher): the dense-row forms are not compiled
template <int NQ, int NC, int E, int Nf, int N1, bool SEP_ONLY = false>
__device__ __forceinline__ void apply_R_t(const Tables& T, const double* __restrict__ src,
                                          double* __restrict__ dst, double* __restrict__ tsc) {
  // Collapsed face: R[(f1,f2)][a1=f1][a2][a3] = E[f2][a2] r3[a3] (Kronecker factors of
  // tensor_simplex.jl:221-306), so the a3-contraction is shared by the N1 rows of a group:
  // 2 * N1 instead of N1^2 terms per row, and every row of R costs the same.
  const int ng = SEP_ONLY ? N1 : T.R_ng;   // SEP_ONLY: one group per a1 (launcher: T.R_ng == N1)
  if (SEP_ONLY || ng > 0) {
    SSE_LOOP(idx, E * NC * ng * N1) {
      const int a2 = idx % N1, g = (idx / N1) % ng, ec = idx / (N1 * ng);
      const double* s0 = src + ec * NQ + __ldg(T.R_gstart + g) + a2 * N1;
      double acc = 0.0;
#pragma unroll
      for (int a3 = 0; a3 < N1; ++a3) acc = fma(T.R_r3[a3], s0[a3], acc);
      tsc[idx] = acc;
    }
    __syncthreads();
  }
  SSE_LOOP(idx, E * Nf) {
    int j = idx % Nf, e = idx / Nf;
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    const int b = __ldg(T.R_rp + j);
    const int desc = __ldg(T.R_desc + j);
    const int start = desc & 1023, stride = (desc >> 10) & 1023, cnt = (desc >> 20) & 127;
    const double* s0 = src + e * NC * NQ + start;
    if (cnt == N1) {
#pragma unroll
      for (int q = 0; q < N1; ++q) {
        double v = __ldg(T.R_v + b + q);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, s0[c * NQ + q * stride], acc[c]);
      }
    } else if (SEP_ONLY || (ng > 0 && cnt == N1 * N1 && stride == 1)) {
      const double* t0 = tsc + (e * NC * ng + __ldg(T.R_grp + j)) * N1;
#pragma unroll
      for (int a2 = 0; a2 < N1; ++a2) {
        double v = __ldg(T.R_E + j * N1 + a2);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, t0[c * ng * N1 + a2], acc[c]);
      }
    } else if (cnt == N1 * N1 && stride == 1) {
#pragma unroll
      for (int q = 0; q < N1 * N1; ++q) {
        double v = __ldg(T.R_v + b + q);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, s0[c * NQ + q], acc[c]);
      }
    } else {
      for (int q = 0; q < cnt; ++q) {
        double v = __ldg(T.R_v + b + q);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, s0[c * NQ + q * stride], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) dst[(e * NC + c) * Nf + j] = acc[c];
  }
  __syncthreads();
}

// weight-adjusted (M^-1 = I) or diagonal mass solve, in place on rhs [E][NC][N_p]
template <int DIM, int N1, int NC, int E>
__device__ __forceinline__ void mass_solve_t(const Tables& T, const Geo& G, long long k0,
                                             double* __restrict__ rhs, double* __restrict__ q,
                                             double* __restrict__ tmp) {
  constexpr int NQ = ipow(N1, DIM);
  if (T.mass_kind == MASS_DIAGONAL) {
    SSE_LOOP(idx, E * NC * NQ) {
      int i = idx % NQ, e = idx / (NQ * NC);
      long long k = min(k0 + e, G.N_e - 1);
      rhs[idx] = fdiv(rhs[idx], T.W[i] * G.J_q[k * NQ + i]);
    }
    __syncthreads();
    return;
  }
  apply_V_t<DIM, N1, NC, E>(vtab(T), rhs, q, tmp);
  SSE_LOOP(idx, E * NQ) {
    int i = idx % NQ, e = idx / NQ;
    long long k = min(k0 + e, G.N_e - 1);
    double sc = fdiv(__ldg(T.W + i), G.J_q[k * NQ + i]);
#pragma unroll
    for (int c = 0; c < NC; ++c) q[(e * NC + c) * NQ + i] *= sc;
  }
  __syncthreads();
  apply_Vt_t<DIM, N1, NC, E>(vtab(T), q, rhs, tmp);
}

// dudt-side epilogue of the loop-B kernels: modal = M^-1 V^T r for the nodal residual r
// [E][NC][NQ] (destroyed).  With the 3-D warped product and the weight-adjusted solver,
// M^-1 V^T r = V^T (W/J) (V V^T r): one fused V V^T pass, the scaling, one V^T.
template <int DIM, int N1, int NC, int E>
__device__ __forceinline__ void project_and_solve_t(const Tables& T, const Geo& G, long long k0,
                                                    double* __restrict__ r,
                                                    double* __restrict__ modal,
                                                    double* __restrict__ tmp) {
  constexpr int NQ = ipow(N1, DIM);
  if constexpr (DIM == 3) {
    if (T.v_kind == V_WARPED && T.mass_kind == MASS_WEIGHT_ADJUSTED && !T.has_Minv) {
      apply_VtV_t<N1, NC, E>(vtab(T), r, tmp);
      SSE_LOOP(idx, E * NQ) {
        int i = idx % NQ, e = idx / NQ;
        long long k = min(k0 + e, G.N_e - 1);
        double sc = fdiv(__ldg(T.W + i), G.J_q[k * NQ + i]);
#pragma unroll
        for (int c = 0; c < NC; ++c) r[(e * NC + c) * NQ + i] *= sc;
      }
      __syncthreads();
      apply_Vt_t<DIM, N1, NC, E>(vtab(T), r, modal, tmp);
      return;
    }
  }
  apply_Vt_t<DIM, N1, NC, E>(vtab(T), r, modal, tmp);
  mass_solve_t<DIM, N1, NC, E>(T, G, k0, modal, r, tmp);
}

// =========================================================================== loop A
// facet nodes of a tensor-product element
template <int DIM, int N1, bool COLLAPSED>
struct TensorNF {
  static constexpr int value = COLLAPSED ? (DIM == 3 ? 4 * N1 * N1 : 3 * N1)
                                         : 2 * DIM * ipow(N1, DIM - 1);
};

// Loop A: one CTA (128 threads) owns E = floor(128 / N_q) whole elements.  (Batching more
// elements per CTA as extra "components" of the applies was measured 10 % slower for systems on
// B200 -- the larger shared-memory footprint costs resident CTAs -- and is not used.)
//
// Shared memory (doubles):  bufQ [E*NC*NQ] | region, where the region holds the V-apply scratch
// followed by the modal block bufP [E*NC*N_p], and is reused as bufF [E*NC*N_f] once the last
// V / V^T apply is done.  The entropy variables overwrite the nodal values in place.  Tet p=4
// Euler: 625 + 550 doubles = 9.4 KB per CTA, so residency is bounded by registers, not smem.
template <int DIM, int N1>
struct NodalCfg {
  static constexpr int NQ = ipow(N1, DIM);
  static constexpr int E = (128 / NQ) > 0 ? 128 / NQ : 1;
  static __host__ __device__ constexpr int tmp(int NC) { return E * NC * vtmp_per_column<DIM, N1>(); }
  static __host__ __device__ constexpr int mx(int a, int b) { return a > b ? a : b; }
  // the fused V V^T pass needs two Z buffers; the second one overlays bufP, dead by then
  static __host__ __device__ constexpr int region(int NC, int Np, int Nf) {
    return mx(mx(tmp(NC) + E * NC * Np, E * NC * Nf), DIM == 3 ? 2 * tmp(NC) : 0);
  }
  static __host__ __device__ constexpr size_t bytes(int NC, int Np, int Nf) {
    return sizeof(double) * (size_t)(E * NC * NQ + region(NC, Np, Nf) + E * NC * rsep_per_column<N1>() +
                                     (DIM == 3 ? E * NQ : 0));
  }
};

// Resident CTAs per SM the loop-A kernel is compiled for (register cap 65536 / (128 * MINB)).  The
// kernel is latency-bound: Tet p=4 gains 2 % at 12 CTAs (40 registers, 8 bytes of spill) over 10
// and loses 4 % at 16 (32 registers, 56 bytes); the other instantiations spill more at 40
// registers and stay at 10 (profiles/r2_ab_log.md).
#ifndef SSE_NODAL_MINB_T5
#define SSE_NODAL_MINB_T5 12
#endif
#ifndef SSE_NODAL_MINB
#define SSE_NODAL_MINB(DIM, N1) (((DIM) == 3 && (N1) == 5) ? SSE_NODAL_MINB_T5 : 10)
#endif

// PROJ_CT: -1 = the projection mode is the run-time argument; 2 = compiled for the entropy
// projection only (3-D: warped-product V with the weight-adjusted mass solver, checked by the
// launcher) -- the other modes' code and their tests leave the kernel (Tet p=4 Euler: 4096 -> 3496
// SASS instructions, loop A -5.2 % on B200, profiles/r2_ab_log.md session AA)
template <int DIM, int N1, int LAW, bool COLLAPSED, int PROJ_CT = -1>
__global__ void __launch_bounds__(128, SSE_NODAL_MINB(DIM, N1))
k_nodal_tensor(Tables T, Geo G, Phys P, const double* __restrict__ u, double* __restrict__ u_q,
               double* __restrict__ u_f, int proj_rt) {
  const int proj = PROJ_CT >= 0 ? PROJ_CT : proj_rt;
  constexpr int NC = LawTraits<DIM, LAW>::NC;
  constexpr int NQ = ipow(N1, DIM);
  using Cf = NodalCfg<DIM, N1>;
  constexpr int E = Cf::E;
  constexpr int Nf = TensorNF<DIM, N1, COLLAPSED>::value;
  SSE_SHARED16(sm);
  const int Np = T.N_p;
  double* bufQ = sm;
  double* tmp = bufQ + E * NC * NQ;
  double* bufP = tmp + Cf::tmp(NC);
  double* bufF = tmp;                      // aliases tmp/bufP; live only after the last apply
  double* rsc = tmp + Cf::region(NC, Np, Nf);   // scratch of the separable rows of R
  double* scq = rsc + E * NC * rsep_per_column<N1>();   // W/J at the volume nodes (3-D only)
  const long long k0 = G.k_begin + (long long)blockIdx.x * E;
  const int Ev = (int)min((long long)E, G.N_e - k0);

  if (G.pf_dist > 0) {   // L2 prefetch for the CTA one wave ahead (see k_fluxdiff_tensor)
    const long long kp = k0 + G.pf_dist;
    if (kp + E <= G.N_e) {
      for (int o = threadIdx.x * 128; o < E * NC * Np * 8; o += 128 * 128)
        SSE_PREFETCH_L2((const char*)(u + kp * NC * Np) + o);
      if (proj == 2)
        for (int o = threadIdx.x * 128; o < E * NQ * 8; o += 128 * 128)
          SSE_PREFETCH_L2((const char*)(G.J_q + kp * NQ) + o);
    }
  }
  // own node's Jacobian for the two weightings of the projection (thread = volume node)
  double jq = 1.0;
  if (proj == 2 && threadIdx.x < E * NQ) {
    const int i = threadIdx.x % NQ, e = threadIdx.x / NQ;
    jq = __ldcg(G.J_q + min(k0 + e, G.N_e - 1) * NQ + i);
  }
  SSE_LOOP(idx, E * NC * Np) bufP[idx] = (idx < Ev * NC * Np) ? __ldcg(u + k0 * NC * Np + idx) : 1.0;
  if constexpr (DIM == 3) {   // W/J of the weight-adjusted mass inverse, read by line in the fused stage
    if (proj == 2 && threadIdx.x < E * NQ) scq[threadIdx.x] = fdiv(__ldg(T.W + threadIdx.x % NQ), jq);
  }
  __syncthreads();
  if constexpr (PROJ_CT == 2 && DIM == 3)
    apply_V3_t<N1, NC, E, SSE_NODAL_OOL>(V3Tab{T.wC, T.wCt, T.pairtab, T.modetab, T.wK}, bufP, bufQ, tmp);
  else
    apply_V_t<DIM, N1, NC, E, true>(vtab(T), bufP, bufQ, tmp);
  if (proj == 0) {
    apply_R_t<NQ, NC, E, Nf, N1>(T, bufQ, bufF, rsc);
    SSE_LOOP(idx, Ev * NC * NQ) u_q[k0 * NC * NQ + idx] = bufQ[idx];
    SSE_LOOP(idx, Ev * NC * Nf) u_f[k0 * NC * Nf + idx] = bufF[idx];
    return;
  }
  if (proj != 2) {   // nodal schemes keep the nodal values themselves as u_q
    SSE_LOOP(idx, Ev * NC * NQ) u_q[k0 * NC * NQ + idx] = bufQ[idx];
    __syncthreads();
  }
  // entropy variables, in place (each thread reads and rewrites its own node only)
  if (threadIdx.x < E * NQ) {
    const int i = threadIdx.x % NQ, e = threadIdx.x / NQ;
    double uu[NC], w[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) uu[c] = bufQ[(e * NC + c) * NQ + i];
    cons_to_entropy<DIM, LAW>(P, uu, w);
    const double sc = (proj == 2) ? __ldg(T.W + i) * jq : 1.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) bufQ[(e * NC + c) * NQ + i] = w[c] * sc;
  }
  __syncthreads();
  if (DIM == 3 && (PROJ_CT == 2 || (proj == 2 && T.v_kind == V_WARPED && T.mass_kind == MASS_WEIGHT_ADJUSTED))) {
    // projected entropy variables at the volume nodes, V M^-1 V^T (W J w) with
    // M^-1 = V^T (W/J) V:  two fused V V^T passes around the W/J scaling, no modal intermediate
#ifdef SSE_NO_FUSED_SCALE   // A/B knob: two V V^T passes around a scaling pass
    if constexpr (DIM == 3) apply_VtV_t<N1, NC, E, true>(vtab(T), bufQ, tmp);
    if (threadIdx.x < E * NQ) {
      const int i = threadIdx.x % NQ, e = threadIdx.x / NQ;
#pragma unroll
      for (int c = 0; c < NC; ++c) bufQ[(e * NC + c) * NQ + i] *= scq[threadIdx.x];
    }
    __syncthreads();
    if constexpr (DIM == 3) apply_VtV_t<N1, NC, E, true>(vtab(T), bufQ, tmp);
#else
    if constexpr (DIM == 3) apply_VtV_scaled_VtV_t<N1, NC, E, SSE_NODAL_OOL>(vtab(T), bufQ, tmp, scq);
#endif
  } else if (proj == 2) {
    apply_Vt_t<DIM, N1, NC, E, true>(vtab(T), bufQ, bufP, tmp);
    // mass solve (weight-adjusted, M^-1 = I): V, W/J, V^T -- or the diagonal scaling
    if (T.mass_kind == MASS_DIAGONAL) {
      SSE_LOOP(idx, E * NC * NQ) {
        int i = idx % NQ, e = idx / (NQ * NC);
        long long k = min(k0 + e, G.N_e - 1);
        bufP[idx] = fdiv(bufP[idx], __ldg(T.W + i) * __ldcg(G.J_q + k * NQ + i));
      }
      __syncthreads();
    } else {
      apply_V_t<DIM, N1, NC, E, true>(vtab(T), bufP, bufQ, tmp);
      if (threadIdx.x < E * NQ) {
        const int i = threadIdx.x % NQ, e = threadIdx.x / NQ;
        const double sc = fdiv(__ldg(T.W + i), jq);
#pragma unroll
        for (int c = 0; c < NC; ++c) bufQ[(e * NC + c) * NQ + i] *= sc;
      }
      __syncthreads();
      apply_Vt_t<DIM, N1, NC, E, true>(vtab(T), bufQ, bufP, tmp);
    }
    apply_V_t<DIM, N1, NC, E, true>(vtab(T), bufP, bufQ, tmp);
  }
  apply_R_t<NQ, NC, E, Nf, N1, (PROJ_CT == 2 && DIM == 3)>(T, bufQ, bufF, rsc);
  // entropy -> conservative variables at the volume nodes (modal case) and the facet nodes,
  // one loop so the log/exp sequence is instantiated once
  const int nvol = (proj == 2) ? Ev * NQ : 0;
  SSE_LOOP(idx, nvol + Ev * Nf) {
    const bool vol = idx < nvol;
    const int ii = vol ? idx : idx - nvol;
    const int npt = vol ? NQ : Nf;
    const int pt = E == 1 ? ii : ii % npt, e = E == 1 ? 0 : ii / npt;
    const double* srcw = vol ? bufQ : bufF;
    double w[NC], uu[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) w[c] = srcw[(e * NC + c) * npt + pt];
    entropy_to_cons<DIM, LAW>(P, w, uu);
    double* dstu = vol ? u_q : u_f;
#pragma unroll
    for (int c = 0; c < NC; ++c) dstu[((k0 + e) * NC + c) * npt + pt] = uu[c];
  }
}

// Loop A of scalar conservation laws without entropy projection (standard form, proj == 0):
// u_q = V u, u_f = R u_q.  One element is a single column for the V / R applies -- far too little
// work for a CTA -- so NB consecutive elements ride along as "components" (as in
// k_standard_tensor): table reads, index arithmetic and barriers are amortised over NB elements
// and the stage-B / stage-A work items fill the warps.
template <int DIM, int N1, int NB>
struct NodalBatchCfg {
  static constexpr int NQ = ipow(N1, DIM);
  static __host__ __device__ constexpr int region(int Np, int Nf) {
    return NB * (vtmp_per_column<DIM, N1>() + Np) > NB * Nf ? NB * (vtmp_per_column<DIM, N1>() + Np)
                                                          : NB * Nf;
  }
  static __host__ __device__ constexpr size_t bytes(int Np, int Nf) {
    return sizeof(double) * (size_t)(NB * NQ + region(Np, Nf) + NB * rsep_per_column<N1>());
  }
};

#ifndef SSE_NODALB_MINB
#define SSE_NODALB_MINB 0
#endif
template <int DIM, int N1, int LAW, bool COLLAPSED, int NB>
#if SSE_NODALB_MINB > 0
__global__ void __launch_bounds__(128, SSE_NODALB_MINB)
#else
__global__ void __launch_bounds__(128)
#endif
k_nodal_batched(Tables T, Geo G, const double* __restrict__ u, double* __restrict__ u_q,
                double* __restrict__ u_f) {
  static_assert(LawTraits<DIM, LAW>::NC == 1, "scalar conservation laws only");
  constexpr int NQ = ipow(N1, DIM);
  constexpr int Nf = TensorNF<DIM, N1, COLLAPSED>::value;
  SSE_SHARED16(sm);
  const int Np = T.N_p;
  double* bufQ = sm;
  double* tmp = bufQ + NB * NQ;
  double* bufP = tmp + NB * vtmp_per_column<DIM, N1>();
  double* bufF = tmp;                      // aliases tmp/bufP once the V apply is done
  const long long k0 = G.k_begin + (long long)blockIdx.x * NB;
  const int Ev = (int)min((long long)NB, G.N_e - k0);
  SSE_LOOP(idx, NB * Np) bufP[idx] = (idx < Ev * Np) ? __ldcg(u + k0 * Np + idx) : 0.0;
  __syncthreads();
  apply_V_t<DIM, N1, NB, 1>(vtab(T), bufP, bufQ, tmp);
  apply_R_t<NQ, NB, 1, Nf, N1>(T, bufQ, bufF, tmp + NodalBatchCfg<DIM, N1, NB>::region(Np, Nf));
  SSE_LOOP(idx, Ev * NQ) u_q[k0 * NQ + idx] = bufQ[idx];
  SSE_LOOP(idx, Ev * Nf) u_f[k0 * Nf + idx] = bufF[idx];
}

// ============================================================ projection on the batched engine
// (A loop-A kernel on the same engine, E = 4 / 5 / 8 elements per CTA, was measured and deleted: it
// executes 25 % fewer instructions than k_nodal_tensor but, at 10 KB of shared memory per element,
// keeps 16-20 warps per SM instead of 40 and was 18-50 % slower: profiles/r2_engine_ab.md.)
// k_project_tet: dudt = M^-1 V^T r for the nodal residual r [k][NC][NQ] that the loop-B kernel
// leaves in global memory, with the weight-adjusted solver: V^T (W/J) (V V^T r), followed by the
// dudt store or the fused low-storage Runge-Kutta update.  E elements per CTA on the batched
// engine -- as the tail of the loop-B kernel (one element per CTA) the same algebra ran its ragged
// stages at a quarter of the lanes and was 27 % of that kernel's time.
// Columns: NC components of E = G * NCOL / NC consecutive elements, NCOL columns per work item of
// the engine (systems: NCOL = NC, one element per item group; scalar laws: NCOL consecutive
// ELEMENTS ride along as the columns of a group).
template <int N1, int NC, int NCOL, int G>
struct ProjectTetCfg {
  using D = V3Dims<N1>;
  static_assert((G * NCOL) % NC == 0, "whole elements per CTA");
  static constexpr int NQ = D::N3, COLS = G * NCOL, E = COLS / NC;
  static constexpr int oX = 0;
  static constexpr int oZ = oX + COLS * NQ;
  static constexpr int oM = oX;   // the modal result overlays X, dead once the last B^T has run
  // Systems (one element per column group): W/J at the nodes of the E elements is kept in shared
  // memory and A, W/J, A^T run as one pass on the a1-line.  Scalar laws (every column another
  // element) would double their footprint that way and keep the three passes.
#ifdef SSE_NO_FUSED_SCALE
  static constexpr bool FUSE_SCALE = false;
#else
  static constexpr bool FUSE_SCALE = (NCOL == NC);
#endif
  static constexpr int oS = oZ + G * VBLayout<N1, NCOL>::ZG;
  static constexpr size_t bytes = sizeof(double) * (size_t)(oS + (FUSE_SCALE ? E * NQ : 0));
  static constexpr int NR = (E * NQ + 127) / 128;
};

#ifndef SSE_PROJECT_TET_E
#define SSE_PROJECT_TET_E 4
#endif
#ifndef SSE_PROJECT_TET_MINB
#define SSE_PROJECT_TET_MINB 5
#endif
// two work items per column set in the ragged stages (all four warps busy, critical path of the
// stage 3/5): measured 2.8 % SLOWER on loop B (the coefficient blocks are fetched twice) -- off
#ifndef SSE_PROJECT_TET_REGSCALE
#define SSE_PROJECT_TET_REGSCALE 1   // scalar laws: A (W/J) A^T fused with the scaling in registers
#endif
#ifndef SSE_PROJECT_TET_SPLIT
#define SSE_PROJECT_TET_SPLIT 1
#endif

template <int N1, int NC, int NCOL, int G>
__global__ void __launch_bounds__(128, SSE_PROJECT_TET_MINB)
k_project_tet(Tables T, Geo G_, RK rk, const double* __restrict__ r_q, double* __restrict__ dudt) {
  using Cf = ProjectTetCfg<N1, NC, NCOL, G>;
  using D = V3Dims<N1>;
  constexpr int NQ = Cf::NQ, NR = Cf::NR, E = Cf::E;
  const Geo& Gm = G_;
  SSE_SHARED16(sm);
  double* X = sm + Cf::oX;
  double* Z = sm + Cf::oZ;
  double* M = sm + Cf::oM;
  const int tid = threadIdx.x;
  const long long k0 = Gm.k_begin + (long long)blockIdx.x * E;
  const V3Tab v3{T.wC, T.wCt, T.pairtab, T.modetab, T.wK};
  for (int idx = tid; idx < E * NC * NQ; idx += 128) {   // asynchronous copies, all in flight
    const int e = idx / (NC * NQ);
    const long long k = min(k0 + e, Gm.N_e - 1);
    SSE_CP_ASYNC8(X + idx, r_q + k * NC * NQ + (idx - e * NC * NQ));
  }
  double* SC = sm + Cf::oS;
  // W/J.  Systems: in shared memory, read by line in the fused stage.  Scalar laws with one line
  // item per thread (REG_SCALE): the 5 x NCOL values of the thread's own line in registers.
  // Otherwise: the thread's NR nodes in registers for a separate scaling pass.
  constexpr bool REG_SCALE = !Cf::FUSE_SCALE && NC == 1 && G * D::N2 <= 128 && SSE_PROJECT_TET_REGSCALE;
  constexpr int NJ = REG_SCALE ? NCOL * N1 : NR;
  double jq[NJ];
  if constexpr (REG_SCALE) {
    const int a23 = tid % D::N2, g = min(tid / D::N2, G - 1);
#pragma unroll
    for (int c = 0; c < NCOL; ++c)
#pragma unroll
      for (int a1 = 0; a1 < N1; ++a1)
        jq[c * N1 + a1] = __ldcg(Gm.J_q + min(k0 + g * NCOL + c, Gm.N_e - 1) * NQ + a1 * D::N2 + a23);
#pragma unroll
    for (int c = 0; c < NCOL; ++c)
#pragma unroll
      for (int a1 = 0; a1 < N1; ++a1)
        jq[c * N1 + a1] = fdiv(__ldg(T.W + a1 * D::N2 + a23), jq[c * N1 + a1]);
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int idx = tid + r * 128;
      jq[r] = 1.0;
      if (idx < E * NQ) jq[r] = __ldcg(Gm.J_q + min(k0 + idx / NQ, Gm.N_e - 1) * NQ + idx % NQ);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int idx = tid + r * 128;
      if (idx < E * NQ) {
        jq[r] = fdiv(__ldg(T.W + idx % NQ), jq[r]);
        if constexpr (Cf::FUSE_SCALE) SC[idx] = jq[r];
      }
    }
  }
  SSE_CP_ASYNC_WAIT_ALL();
  __syncthreads();
  // work items per column set in the ragged stages: two where both halves still fit one round
  constexpr int SP = (SSE_PROJECT_TET_SPLIT == 2 && 2 * D::T2 * G <= 128) ? 2 : 1;
  vb_stageA<N1, NCOL, G, true>(tid, 128, X);
  __syncthreads();
  vb_stageB<N1, NCOL, G, true, SP>(tid, 128, Z, X);
  __syncthreads();
  vb_stageK<N1, NCOL, G, SP>(tid, 128, v3, Z);
  __syncthreads();
  vb_stageB<N1, NCOL, G, false, SP>(tid, 128, Z, X);
  __syncthreads();
  if constexpr (Cf::FUSE_SCALE) {
    vb_stageA_scale_At<N1, NCOL, G, NC>(tid, 128, X, SC);   // A, W/J, A^T on the line in registers
    __syncthreads();
  } else if constexpr (REG_SCALE) {
    vb_stageA_scale_At_regs<N1, NCOL, G>(tid, X, jq);
    __syncthreads();
  } else {
    vb_stageA<N1, NCOL, G, false>(tid, 128, X);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int idx = tid + r * 128;
      if (idx < E * NQ) {
        const int i = idx % NQ, e = idx / NQ;
        double* xb = X + e * NC * NQ + i;
#pragma unroll
        for (int c = 0; c < NC; ++c) xb[c * NQ] *= jq[r];
      }
    }
    __syncthreads();
    vb_stageA<N1, NCOL, G, true>(tid, 128, X);
    __syncthreads();
  }
  vb_stageB<N1, NCOL, G, true, SP>(tid, 128, Z, X);
  __syncthreads();
  vb_stageC<N1, NCOL, G, true, SP>(tid, 128, v3, M, Z);
  __syncthreads();
  store_result(T, Gm, rk, k0, E, NC, M, dudt);
}

// Warp-private edition of the projection for systems (NCOL = NC): every WARP owns one element --
// its nodal block X, intermediate Z and W/J in its own slice of shared memory -- and runs the seven
// stages on its own, separated by __syncwarp() only.  The work items of the batched engine carry 25
// independent FMA chains each (all columns of a line / b1-group / pair), so a warp needs no other
// warp to hide latency, and no warp ever waits at a block barrier for the two warps that hold the
// 60 items of a ragged stage (k_project_tet: 3.7 stalled warps per issue at barriers).  The price
// is lane utilisation (25 line items, 15 group / pair items per element on 32 lanes).
#ifndef SSE_PROJECT_TET_W_EW
#define SSE_PROJECT_TET_W_EW 1      // elements per warp
#endif
#ifndef SSE_PROJECT_TET_W_MINB
#define SSE_PROJECT_TET_W_MINB 5
#endif
template <int N1, int NC>
struct ProjectTetWarpCfg {
  using D = V3Dims<N1>;
  static constexpr int NQ = D::N3, WPB = 4, EW = SSE_PROJECT_TET_W_EW;   // warps per CTA, elements per warp
  static constexpr int ZG = VBLayout<N1, NC>::ZG;
  static constexpr int per_warp = ((EW * (NC * NQ + ZG + NQ)) + 1) & ~1;  // doubles, even
  static constexpr size_t bytes = sizeof(double) * (size_t)(WPB * per_warp);
};
template <int N1, int NC>
__global__ void __launch_bounds__(128, SSE_PROJECT_TET_W_MINB)
k_project_tet_w(Tables T, Geo G_, RK rk, const double* __restrict__ r_q, double* __restrict__ dudt) {
  using Cf = ProjectTetWarpCfg<N1, NC>;
  using D = V3Dims<N1>;
  constexpr int NQ = Cf::NQ, EW = Cf::EW;
  const Geo& Gm = G_;
  SSE_SHARED16(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* X = sm + warp * Cf::per_warp;          // [EW][NC][NQ]
  double* Z = X + EW * NC * NQ;                  // [EW][ZG]
  double* SC = Z + EW * Cf::ZG;                  // [EW][NQ]
  double* M = X;                       // the modal result overlays X, dead once the last B^T has run
  const long long k0 = Gm.k_begin + ((long long)blockIdx.x * Cf::WPB + warp) * EW;
  const V3Tab v3{T.wC, T.wCt, T.pairtab, T.modetab, T.wK};
  // (elements past the end compute on the last element and store nothing)
  for (int idx = lane; idx < EW * NC * NQ; idx += 32) {
    const int e = idx / (NC * NQ);
    SSE_CP_ASYNC8(X + idx, r_q + min(k0 + e, Gm.N_e - 1) * NC * NQ + (idx - e * NC * NQ));
  }
  for (int i = lane; i < EW * NQ; i += 32)
    SC[i] = fdiv(__ldg(T.W + i % NQ), __ldcg(Gm.J_q + min(k0 + i / NQ, Gm.N_e - 1) * NQ + i % NQ));
  SSE_CP_ASYNC_WAIT_ALL();
  __syncwarp();
  vb_stageA<N1, NC, EW, true>(lane, 32, X);
  __syncwarp();
  vb_stageB<N1, NC, EW, true>(lane, 32, Z, X);
  __syncwarp();
  vb_stageK<N1, NC, EW>(lane, 32, v3, Z);
  __syncwarp();
  vb_stageB<N1, NC, EW, false>(lane, 32, Z, X);
  __syncwarp();
  vb_stageA_scale_At<N1, NC, EW, NC>(lane, 32, X, SC);
  __syncwarp();
  vb_stageB<N1, NC, EW, true>(lane, 32, Z, X);
  __syncwarp();
  vb_stageC<N1, NC, EW, true>(lane, 32, v3, M, Z);
  __syncwarp();
  const int blk = NC * T.N_p;
  for (int idx = lane; idx < EW * blk; idx += 32) {
    const long long k = k0 + idx / blk;
    if (k < Gm.N_e) {
      const long long g = k0 * blk + idx;
      if (rk.mode == 0) {
        dudt[g] = M[idx];
      } else {
        const double kk = rk.a * rk.k[g] + rk.dt * M[idx];
        rk.k[g] = kk;
        rk.u[g] += rk.b * kk;
      }
    }
  }
}

// ==================================================== loop B, flux-differencing form
// Compile-time geometry of the specialised loop-B kernel: EL elements per 128-thread CTA,
// NF facet nodes, and the shared-memory carve-up (in doubles; regions holding double2 start
// on even offsets).
// Occupancy knobs of the specialised loop-B kernel: minimum resident CTAs per SM (register
// cap = 65536 / (128 * SSE_FD_MINB)), ELL slots exchanged per facet-correction part, and
// single- vs double-buffered pair exchange (1 = single buffer + one more barrier per direction).
#ifndef SSE_FD_MINB
#define SSE_FD_MINB 4
#endif
#ifndef SSE_FD_KQ
#define SSE_FD_KQ 0      // 0: two halves (ceil(KC/2) slots per part)
#endif
#ifndef SSE_FD_FACET_UNROLL
#define SSE_FD_FACET_UNROLL 1
#endif
#ifndef SSE_FD_SINGLE_BUF
#define SSE_FD_SINGLE_BUF 0
#endif

template <int DIM, int N1, int LAW, bool COLLAPSED, int KC>
struct FDCfg {
  static constexpr int NC = LawTraits<DIM, LAW>::NC;
  static constexpr int NS = LawTraits<DIM, LAW>::NS;
  static constexpr int NS2 = (NS + 1) / 2;
  static constexpr int NQ = ipow(N1, DIM);
  static constexpr int EL = (128 / NQ) > 0 ? 128 / NQ : 1;
  static constexpr int NF = COLLAPSED ? (DIM == 3 ? 4 * N1 * N1 : 3 * N1)
                                      : 2 * DIM * ipow(N1, DIM - 1);
  static constexpr int H = N1 / 2;
  static constexpr int KH = SSE_FD_KQ > 0 ? SSE_FD_KQ : (KC + 1) / 2;   // slots per part
  static constexpr int NPART = (KC + KH - 1) / KH;
  static constexpr int NBUF = SSE_FD_SINGLE_BUF ? 1 : 2;
  static constexpr int nq = EL * NQ, nf = EL * NF;
  static __host__ __device__ constexpr int ev(int n) { return (n + 1) & ~1; }
  static constexpr int oS = 0;                                   // double2 [NS2][nq]
  static constexpr int oLa = oS + 2 * NS2 * nq;                  // double2 [DIM][nq]
  static constexpr int oLb = oLa + 2 * DIM * nq;                 // double  [DIM][nq] (3-D)
  static constexpr int oSf = oLb + ev(DIM == 3 ? DIM * nq : 0);  // double2 [NS2][nf]
  static constexpr int oNf = oSf + 2 * NS2 * nf;                 // double  [DIM][nf]
  static constexpr int oFf = oNf + ev(DIM * nf);                 // double  [EL][NC][NF]
  // r_q and the modal result alias the state/metric staging area, which is dead by then
  static constexpr int oR = 0;                                   // double  [EL][NC][NQ]
  static constexpr int oM = oR + ev(NC * nq);                    // double  [EL][NC][Np]
  static constexpr int oEnd = oFf + ev(NC * nf);
  static __host__ __device__ constexpr int xmax(int a, int b) { return a > b ? a : b; }
  static constexpr int sX = xmax(xmax(NBUF * H * NC, KH * NC), 2 * NC) * nq;
  static __host__ __device__ constexpr int oX(int Np) {
    return xmax(oEnd, oM + ev(EL * NC * Np));
  }
  static __host__ __device__ constexpr size_t bytes(int Np) { return sizeof(double) * (size_t)(oX(Np) + sX); }
  // volume-only kernel: state / metric staging + the pair-exchange buffers
  static __host__ __device__ constexpr size_t bytes_volume() {
    return sizeof(double) * (size_t)(oSf + NBUF * H * NC * nq);
  }
};

// PART selects what the body does: 0 = the whole of loop B in one kernel; 3 = everything up to the
// nodal residual r_q, which k_project_tet then projects (the default on tetrahedra);
// 1 = volume flux differencing only, nodal residual written to r_q; 2 = everything else (interface
// flux, facet correction, lift, projection, mass solve, epilogue), nodal residual read from r_q.
// The split pair exists so that the volume kernel can run under its own register / shared-memory
// budget and be measured against the FP64 roofline on its own (opt-in, SSE_B200_SPLIT_B=1).
template <int DIM, int N1, int LAW, bool COLLAPSED, int KC, int PART>
__device__ __forceinline__ void fluxdiff_tensor_body(
    const FastTables& F, const Tables& T, const Geo& G, const Phys& P, const RK& rk,
    const double* __restrict__ u_q, const double* __restrict__ u_f, double* __restrict__ dudt,
    double* __restrict__ r_q) {
  using Cf = FDCfg<DIM, N1, LAW, COLLAPSED, KC>;
  constexpr int NC = Cf::NC, NS2 = Cf::NS2, NQ = Cf::NQ, NF = Cf::NF;
  constexpr int DD = DIM * DIM, H = Cf::H, KH = Cf::KH, EL = Cf::EL, nq = Cf::nq, nf = Cf::nf;
  SSE_SHARED16(sm);
  const int Np = T.N_p;
  double2* sS2 = reinterpret_cast<double2*>(sm + Cf::oS);
  double2* sLa = reinterpret_cast<double2*>(sm + Cf::oLa);
  double* sLb = sm + Cf::oLb;
  double2* sSf2 = reinterpret_cast<double2*>(sm + Cf::oSf);
  double* sNf = sm + Cf::oNf;
  double* sFf = sm + Cf::oFf;
  double* sR = sm + Cf::oR;
  double* sM = sm + Cf::oM;
  double* sX = sm + (PART == 1 ? Cf::oSf : Cf::oX(Np));
  const long long k0 = G.k_begin + (long long)blockIdx.x * EL;
  const int tid = threadIdx.x;
  const bool active = tid < nq;
  const int e = active ? tid / NQ : 0;
  const int i = active ? tid % NQ : 0;
  // tensor coordinates of the own node, slowest first
  const int ia1 = i / ipow(N1, DIM - 1);
  const int ia2 = (i / ipow(N1, DIM >= 2 ? DIM - 2 : 0)) % N1;
  const int ia3 = DIM == 3 ? i % N1 : 0;

  // L2 prefetch of the inputs of the CTA that will run in this slot one wave later (CTAs are
  // scheduled in blockIdx order, pf_dist = SMs x resident CTAs x EL elements ahead): its
  // prologue then waits for L2 instead of DRAM.
  if (G.pf_dist > 0) {
    const long long kp = k0 + G.pf_dist;
    if (kp + EL <= G.N_e) {
      auto pf = [&](const void* base, int bytes) {
        for (int o = tid * 128; o < bytes; o += 128 * 128)
          SSE_PREFETCH_L2((const char*)base + o);
      };
      pf(u_q + kp * NC * NQ, EL * NC * NQ * 8);
      pf(G.L_q + kp * DD * NQ, EL * DD * NQ * 8);
      if constexpr (PART != 1) {
        pf(G.nJf + kp * NF * DIM, EL * NF * DIM * 8);
        pf(u_f + kp * NC * NF, EL * NC * NF * 8);
        pf(G.J_f + kp * NF, EL * NF * 8);
        pf(G.toff + kp * NF, EL * NF * 4);
        pf(G.J_q + kp * NQ, EL * NQ * 8);
      }
      if constexpr (PART == 2) pf(r_q + kp * NC * NQ, EL * NC * NQ * 8);
    }
  }
  double si[2 * NS2], Li[DD], r[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) r[c] = 0.0;
#pragma unroll
  for (int c = 0; c < 2 * NS2; ++c) si[c] = 0.0;

  // ---- phases 0/1.  All global loads of the CTA's first round are issued before any
  // arithmetic so the DRAM round trips of the volume data (u_q, Λ_q) and of the facet data
  // (J_f, nJf, own trace, exterior offset -> exterior trace) overlap.
  const bool hasf = (PART != 1) && tid < nf;
  const int fj = hasf ? tid % NF : 0, fe = hasf ? tid / NF : 0;
  const long long fk = min(k0 + fe, G.N_e - 1);
  const long long fgj = fk * NF + fj;
  double fJf = 1.0, fnJ[DIM], fum[NC], fup[NC];
  int fext = 0;
  if (hasf) {
    fJf = __ldcg(G.J_f + fgj);
    fext = __ldcg(G.toff + fgj);
#pragma unroll
    for (int m = 0; m < DIM; ++m) fnJ[m] = __ldcg(G.nJf + fgj * DIM + m);
#pragma unroll
    for (int c = 0; c < NC; ++c) fum[c] = __ldcg(u_f + fk * NC * NF + fj + (long long)c * NF);
  }
  double uu0[NC];
  if (active) {
    long long k = min(k0 + e, G.N_e - 1);
#pragma unroll
    for (int c = 0; c < NC; ++c) uu0[c] = __ldcg(u_q + (k * NC + c) * NQ + i);
#pragma unroll
    for (int c = 0; c < DD; ++c) Li[c] = __ldcg(G.L_q + (k * DD + c) * NQ + i);
  }
  if (hasf) {
#pragma unroll
    for (int c = 0; c < NC; ++c) fup[c] = __ldcg(u_f + fext + (long long)c * NF);
  }
  if constexpr (PART == 2) {   // the volume kernel's nodal residual
    if (active) {
      long long k = min(k0 + e, G.N_e - 1);
#pragma unroll
      for (int c = 0; c < NC; ++c) r[c] = __ldcg(r_q + (k * NC + c) * NQ + i);
    }
  }
  // ---- phase 0: stage nodal states and metric terms (kept in registers for the own node)
  if (active) {
    cons_to_state<DIM, LAW>(P, uu0, si);
    if constexpr (LAW == LAW_EULER) euler_state_to_half<DIM>(si);
    if constexpr (PART != 2) {   // other nodes' states are only read by the volume term
#pragma unroll
      for (int c = 0; c < NS2; ++c) sS2[c * nq + tid] = make_double2(si[2 * c], si[2 * c + 1]);
#pragma unroll
      for (int m = 0; m < DIM; ++m) {
        sLa[m * nq + tid] = make_double2(Li[m], Li[m + DIM]);
        if constexpr (DIM == 3) sLb[m * nq + tid] = Li[m + 2 * DIM];
      }
    }
  }
  // ---- phase 1: interface numerical flux at the facet nodes
  for (int idx = tid; idx < (PART == 1 ? 0 : nf); idx += 128) {
    const int j = idx % NF, ee = idx / NF;
    double nJ[DIM], nfv[DIM], sl[2 * NS2], fs[NC];
#pragma unroll
    for (int c = 0; c < 2 * NS2; ++c) sl[c] = 0.0;
    double Jf;
    if (idx == tid) {            // first round: operands are already in registers
      Jf = fJf;
#pragma unroll
      for (int m = 0; m < DIM; ++m) nJ[m] = fnJ[m];
      const double iJf = frcp(Jf);
#pragma unroll
      for (int m = 0; m < DIM; ++m) nfv[m] = nJ[m] * iJf;
      interface_flux_vals<DIM, LAW>(P, P.two_point, fum, fup, nfv, sl, fs);
    } else {
      long long k = min(k0 + ee, G.N_e - 1);
      long long gj = k * NF + j;
      Jf = __ldcg(G.J_f + gj);
      const int ext = __ldcg(G.toff + gj);
#pragma unroll
      for (int m = 0; m < DIM; ++m) nJ[m] = __ldcg(G.nJf + gj * DIM + m);
      const double iJf = frcp(Jf);
#pragma unroll
      for (int m = 0; m < DIM; ++m) nfv[m] = nJ[m] * iJf;
      interface_flux<DIM, LAW>(P, P.two_point, u_f, k * NC * NF + j, ext, NF, nfv, sl, fs);
    }
    if constexpr (LAW == LAW_EULER) euler_state_to_half<DIM>(sl);
    const double bj = __ldg(T.B + j) * Jf;
#pragma unroll
    for (int c = 0; c < NC; ++c) sFf[(ee * NC + c) * NF + j] = bj * fs[c];
#pragma unroll
    for (int c = 0; c < NS2; ++c) sSf2[c * nf + idx] = make_double2(sl[2 * c], sl[2 * c + 1]);
#pragma unroll
    for (int m = 0; m < DIM; ++m) sNf[m * nf + idx] = 0.5 * nJ[m];
  }
  __syncthreads();

  // ---- phase 2: volume flux differencing, every pair on a tensor line evaluated once
#pragma unroll
  for (int l = 0; l < (PART == 2 ? 0 : DIM); ++l) {
    constexpr int s0 = ipow(N1, DIM - 1), s1 = ipow(N1, DIM >= 2 ? DIM - 2 : 0);
    const int stride = (l == 0) ? s0 : (l == 1 ? s1 : 1);
    const int al = (i / stride) % N1;
    double* buf = sX + (Cf::NBUF == 2 ? (l & 1) : 0) * (H * NC * nq);
    if (Cf::NBUF == 1 && l > 0) __syncthreads();   // previous direction's reads are done
    if (active) {
#pragma unroll
      for (int o = 1; o <= H; ++o) {
        const bool mine = (N1 % 2 == 1) || (2 * o < N1) || (al < N1 / 2);
        if (mine) {
          int ap = al + o;
          if (ap >= N1) ap -= N1;
          const int jt = tid + (ap - al) * stride;
          const double* Sp = F.Sp + ((l * H + (o - 1)) * DIM) * NQ + i;
          double sv[DIM], cv[DIM], sj[2 * NS2], f[NC];
#pragma unroll
          for (int m = 0; m < DIM; ++m) {
            const bool used = COLLAPSED ? (m >= l) : (m == l);
            sv[m] = used ? __ldg(Sp + m * NQ) : 0.0;
          }
#pragma unroll
          for (int c = 0; c < NS2; ++c) {
            const double2 v = sS2[c * nq + jt];
            sj[2 * c] = v.x;
            sj[2 * c + 1] = v.y;
          }
#pragma unroll
          for (int n = 0; n < DIM; ++n) cv[n] = 0.0;
#pragma unroll
          for (int m = 0; m < DIM; ++m) {
            const bool used = COLLAPSED ? (m >= l) : (m == l);
            if (used) {
              const double2 la = sLa[m * nq + jt];
              cv[0] = fma(sv[m], Li[m] + la.x, cv[0]);
              cv[1] = fma(sv[m], Li[m + DIM] + la.y, cv[1]);
              if constexpr (DIM == 3) cv[2] = fma(sv[m], Li[m + 2 * DIM] + sLb[m * nq + jt], cv[2]);
            }
          }
          if constexpr (LAW == LAW_EULER) ec_flux_half_c<DIM>(P, si, sj, cv, f);
          else two_point_flux_c<DIM, LAW>(P, P.two_point, si, sj, cv, f);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            r[c] -= f[c];
            buf[((o - 1) * NC + c) * nq + tid] = f[c];
          }
        }
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int o = 1; o <= H; ++o) {
        const bool theirs = (N1 % 2 == 1) || (2 * o < N1) || (al >= N1 / 2);
        if (theirs) {
          int as = al - o;
          if (as < 0) as += N1;
          const int jt = tid + (as - al) * stride;
#pragma unroll
          for (int c = 0; c < NC; ++c) r[c] += buf[((o - 1) * NC + c) * nq + jt];
        }
      }
    }
  }

  if constexpr (PART == 1) {
    if (active && k0 + e < G.N_e) {
#pragma unroll
      for (int c = 0; c < NC; ++c) r_q[((k0 + e) * NC + c) * NQ + i] = r[c];
    }
    return;
  }
  // ---- phases 3/4: facet correction (ELL rows of C = R^T B), exchanged in two halves.  Only the
  // collapsed simplices have one; the non-collapsed instantiations are the diagonal-E schemes on
  // quadrilaterals / hexahedra (LGL collocation, R a selection: flux_differencing_form.jl:171-187)
  if constexpr (COLLAPSED) if (!T.r_is_selection) {
#pragma unroll
    for (int half = 0; half < Cf::NPART; ++half) {
      __syncthreads();   // previous users of sX (pair buffers / previous part) are done
      if (active) {
        const int kend = (half + 1) * KH < KC ? (half + 1) * KH : KC;
        // software-pipelined table read: C[i, j] of slot kk+1 is fetched while slot kk is evaluated
        double cij = __ldg(F.Cv + (half * KH) * NQ + i);
        constexpr int FACET_UNROLL = SSE_FD_FACET_UNROLL;
        int fc_prev = -1;
        double hq[DIM];
#pragma unroll
        for (int n = 0; n < DIM; ++n) hq[n] = 0.0;
#pragma unroll FACET_UNROLL
        for (int kk = half * KH; kk < kend; ++kk) {
          const int kn = (kk + 1 < kend) ? kk + 1 : kk;
          const double cij_next = __ldg(F.Cv + kn * NQ + i);
          const int j = canon_facet_node<DIM, N1>(kk, ia1, ia2, ia3);
          const int fc = kk < DIM ? kk : DIM;   // canonical layout: slots >= DIM lie on the last face
          const int jj = e * NF + j;
          double nJ[DIM], sj[2 * NS2], f[NC];
#pragma unroll
          for (int c = 0; c < NS2; ++c) {
            const double2 v = sSf2[c * nf + jj];
            sj[2 * c] = v.x;
            sj[2 * c + 1] = v.y;
          }
          // ½ nJq of this slot's face (mesh.jl:266-271): consecutive slots mostly belong to the
          // same face (the N1 nodes of the collapsed one), so it is kept across slots
          if (fc != fc_prev) {
            fc_prev = fc;
#pragma unroll
            for (int n = 0; n < DIM; ++n) {
              double acc = 0.0;
#pragma unroll
              for (int m = 0; m < DIM; ++m) acc = fma(Li[m + DIM * n], F.nref[fc * DIM + m], acc);
              hq[n] = acc;                       // F.nref holds ½ n_ref
            }
          }
          // the flux is linear in its direction vector: C[i,j] scales the direction (DIM
          // products) instead of the NC flux components
#pragma unroll
          for (int n = 0; n < DIM; ++n) nJ[n] = cij * (hq[n] + sNf[n * nf + jj]);
          if constexpr (LAW == LAW_EULER) ec_flux_half_c<DIM>(P, si, sj, nJ, f);
          else two_point_flux_c<DIM, LAW>(P, P.two_point, si, sj, nJ, f);
          double* dst = sX + (kk - half * KH) * NC * nq + tid;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            r[c] -= f[c];
            dst[c * nq] = f[c];
          }
          cij = cij_next;
        }
      }
      __syncthreads();
      // f_f -= column sums of this part's terms (canonical layout, see FastTables): per line
      // slot one lane per (element, component, facet node) adds the N1 terms of its tensor line;
      // per collapsed-face slot one lane per (element, component, a1) the N1^2 terms behind it.
      {
        constexpr int NPF = ipow(N1, DIM - 1);
        const int k_lo = half * KH, k_hi = (half + 1) * KH < KC ? (half + 1) * KH : KC;
        const int n_line = (k_hi < 3 ? k_hi : 3) - (k_lo < 3 ? k_lo : 3);   // line slots here
        for (int it = tid; it < n_line * EL * NC * NPF; it += 128) {
          const int g = it % NPF, c = (it / NPF) % NC, ee = (it / (NPF * NC)) % EL;
          const int kk = k_lo + it / (NPF * NC * EL);
          int start, stride;
          if constexpr (DIM == 3) {
            const int g1 = g / N1, g2 = g % N1;
            start = kk == 0 ? g1 * N1 * N1 + g2 : g1 * N1 + g2;
            stride = kk == 0 ? N1 : N1 * N1;
          } else {
            start = kk == 0 ? g * N1 : g;
            stride = kk == 0 ? 1 : N1;
          }
          const double* bc = sX + ((kk - k_lo) * NC + c) * nq + ee * NQ + start;
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < N1; ++q) acc += bc[q * stride];
          sFf[(ee * NC + c) * NF + kk * NPF + g] -= acc;
        }
        if constexpr (DIM == 3) {
          const int c_lo = k_lo > 3 ? k_lo : 3;                 // collapsed-face slots of this part
          const int n_col = k_hi > c_lo ? k_hi - c_lo : 0;
          for (int it = tid; it < n_col * EL * NC * N1; it += 128) {
            const int g1 = it % N1, c = (it / N1) % NC, ee = (it / (N1 * NC)) % EL;
            const int kk = c_lo + it / (N1 * NC * EL);
            const double* bc = sX + ((kk - k_lo) * NC + c) * nq + ee * NQ + g1 * N1 * N1;
            double part[N1];
#pragma unroll
            for (int q2 = 0; q2 < N1; ++q2) part[q2] = bc[q2];
#pragma unroll
            for (int q = N1; q < N1 * N1; ++q) part[q % N1] += bc[q];
            double acc = 0.0;
#pragma unroll
            for (int q2 = 0; q2 < N1; ++q2) acc += part[q2];
            sFf[(ee * NC + c) * NF + 3 * NPF + g1 * N1 + (kk - 3)] -= acc;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- phase 5: r_q -= R^T f_f (ELL), then hand r_q to the modal projection
  if (active) {
    if constexpr (COLLAPSED) {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const int j = canon_facet_node<DIM, N1>(kk, ia1, ia2, ia3);
        const double rv = __ldg(F.Rv + kk * NQ + i);
        const double* ff = sFf + e * NC * NF + j;
#pragma unroll
        for (int c = 0; c < NC; ++c) r[c] = fma(-rv, ff[c * NF], r[c]);
      }
    } else {   // selection R: the node lies on 0..DIM faces (rows of R^T, CSR)
      for (int en = __ldg(T.Rt_rp + i); en < __ldg(T.Rt_rp + i + 1); ++en) {
        const double rv = __ldg(T.Rt_v + en);
        const double* ff = sFf + e * NC * NF + __ldg(T.Rt_ci + en);
#pragma unroll
        for (int c = 0; c < NC; ++c) r[c] = fma(-rv, ff[c * NF], r[c]);
      }
    }
    if constexpr (PART == 3) {   // the projection runs as its own batched kernel (k_project_tet)
      if (k0 + e < G.N_e) {
#pragma unroll
        for (int c = 0; c < NC; ++c) r_q[((k0 + e) * NC + c) * NQ + i] = r[c];
      }
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) sR[(e * NC + c) * NQ + i] = r[c];
    }
  }
  if constexpr (PART == 3) return;
  __syncthreads();
  // ---- phase 6: dudt = M^-1 V^T r_q
  project_and_solve_t<DIM, N1, NC, EL>(T, G, k0, sR, sM, sX);
  store_result(T, G, rk, k0, EL, NC, sM, dudt);
}

template <int DIM, int N1, int LAW, bool COLLAPSED, int KC>
__global__ void __launch_bounds__(128, SSE_FD_MINB)
k_fluxdiff_tensor(FastTables F, Tables T, Geo G, Phys P, RK rk, const double* __restrict__ u_q,
                  const double* __restrict__ u_f, double* __restrict__ dudt) {
  fluxdiff_tensor_body<DIM, N1, LAW, COLLAPSED, KC, 0>(F, T, G, P, rk, u_q, u_f, dudt, nullptr);
}

// Loop B up to the nodal residual r_q (PART 3); k_project_tet finishes dudt = M^-1 V^T r_q.
template <int DIM, int N1, int LAW, bool COLLAPSED, int KC>
__global__ void __launch_bounds__(128, SSE_FD_MINB)
k_fluxdiff_nodal(FastTables F, Tables T, Geo G, Phys P, const double* __restrict__ u_q,
                 const double* __restrict__ u_f, double* __restrict__ r_q) {
  fluxdiff_tensor_body<DIM, N1, LAW, COLLAPSED, KC, 3>(F, T, G, P, RK{}, u_q, u_f, nullptr, r_q);
}

// The split pair (measurement mode): volume term under its own occupancy target, then the rest.
#ifndef SSE_FD_VOL_MINB
#define SSE_FD_VOL_MINB 6
#endif
template <int DIM, int N1, int LAW, bool COLLAPSED, int KC>
__global__ void __launch_bounds__(128, SSE_FD_VOL_MINB)
k_fluxdiff_volume(FastTables F, Tables T, Geo G, Phys P, const double* __restrict__ u_q,
                  double* __restrict__ r_q) {
  fluxdiff_tensor_body<DIM, N1, LAW, COLLAPSED, KC, 1>(F, T, G, P, RK{}, u_q, nullptr, nullptr,
                                                       r_q);
}
template <int DIM, int N1, int LAW, bool COLLAPSED, int KC>
__global__ void __launch_bounds__(128, SSE_FD_MINB)
k_fluxdiff_facet(FastTables F, Tables T, Geo G, Phys P, RK rk, const double* __restrict__ u_q,
                 const double* __restrict__ u_f, double* __restrict__ dudt,
                 double* __restrict__ r_q) {
  fluxdiff_tensor_body<DIM, N1, LAW, COLLAPSED, KC, 2>(F, T, G, P, rk, u_q, u_f, dudt, r_q);
}

// ============================================ loop B, standard form, reference operators
// time_derivative! of standard_form_first_order.jl:16-63 on tensor-product simplices, scalar
// conservation laws (linear advection, Burgers): f_n(u) = a_n φ(u), so the skew-symmetric split
//   r = Σ_m [ D_m^T (h_m ∘ φ) − h_m ∘ (D_m φ) ],   h_m = ½ W Σ_n Λ_η[m,n] a_n
// needs one nodal scalar φ and d metric scalars h_m per node.  HBM-bound: each thread owns node
// i of NB consecutive elements ("elements as components"), so every operator-table read, index
// computation and barrier is amortised over NB elements while 11*NB independent global loads per
// thread are in flight.
// shared (doubles): sPhi[NB][NQ] | sG[DIM][NB][NQ] | sFf[NB][NF] | sD[DIM][N1][N1] |
//                   sR[NB][NQ] | sM[NB][Np] | sX[2*NB*NQ]
template <int DIM, int N1, int LAW, int KC, int NB>
struct STCfg {
  static constexpr int NQ = ipow(N1, DIM);
  static constexpr int NF = TensorNF<DIM, N1, true>::value;
  static constexpr int oPhi = 0;
  static constexpr int oG = oPhi + NB * NQ;
  static constexpr int oFf = oG + DIM * NB * NQ;
  static constexpr int oD = oFf + NB * NF;
  static constexpr int oR = oD + DIM * N1 * N1;
  static __host__ __device__ constexpr int oM() { return oR + NB * NQ; }
  static __host__ __device__ constexpr int oX(int Np) { return oM() + NB * Np; }
  static __host__ __device__ constexpr size_t bytes(int Np) {
    return sizeof(double) * (size_t)(oX(Np) + 2 * NB * NQ);
  }
  // without the projection tail: sR only serves as the scratch of the separable rows of R
  static __host__ __device__ constexpr size_t bytes_nodal() {
    return sizeof(double) * (size_t)(oM());
  }
};

// PROJ = true: the projection / mass solve runs as the tail of this kernel; PROJ = false: the
// nodal residual goes to global memory (dudt = r_q here) and k_project_tet finishes it, 25 elements
// per CTA on the batched engine.
#ifndef SSE_STD_MINB
#define SSE_STD_MINB 12    // resident CTAs per SM the config-3 operator kernel is compiled for (40 registers; 10: +2.6 %, 16: spills, +60 %)
#endif
template <int DIM, int N1, int LAW, int KC, int NB, bool PROJ>
__global__ void __launch_bounds__(128, PROJ ? 9 : SSE_STD_MINB)   // (the fused-projection form spills at 40 registers)
k_standard_tensor(FastTables F, Tables T, Geo G, Phys P, RK rk, const double* __restrict__ u_q,
                  const double* __restrict__ u_f, double* __restrict__ dudt) {
  static_assert(LAW != LAW_EULER, "scalar conservation laws only");
  using Cf = STCfg<DIM, N1, LAW, KC, NB>;
  constexpr int NQ = Cf::NQ, NF = Cf::NF;
  constexpr int DD = DIM * DIM;
  SSE_SHARED16(sm);
  const int Np = T.N_p;
  double* sPhi = sm + Cf::oPhi;
  double* sG = sm + Cf::oG;
  double* sFf = sm + Cf::oFf;
  double* sD = sm + Cf::oD;
  double* sR = sm + Cf::oR;
  double* sM = sm + Cf::oM();
  double* sX = sm + Cf::oX(Np);
  const long long k0 = G.k_begin + (long long)blockIdx.x * NB;
  const int tid = threadIdx.x;
  const bool active = tid < NQ;
  const int i = active ? tid : 0;

  // (An L2 prefetch of the next wave's inputs, as in k_fluxdiff_tensor, was measured on this
  // kernel: 1.283 vs 1.228 ms at 196 608 elements, i.e. 4.5 % SLOWER -- the kernel already keeps
  // 44 independent loads per thread in flight and the extra requests only compete with them.)
  for (int idx = tid; idx < DIM * N1 * N1; idx += 128) sD[idx] = __ldg(F.D1 + idx);


  double ha[NB][DIM];
  // ---- phase 0: φ(u) and the metric scalars h_m at the volume nodes
  if (active) {
    double uu[NB], Lq[NB][DD];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const long long k = min(k0 + b, G.N_e - 1);
      uu[b] = __ldcg(u_q + k * NQ + i);
#pragma unroll
      for (int c = 0; c < DD; ++c) Lq[b][c] = __ldcg(G.L_q + (k * DD + c) * NQ + i);
    }
    const double hw = 0.5 * __ldg(T.W + i);
    double gref[DD];
#pragma unroll
    for (int c = 0; c < DD; ++c) gref[c] = __ldg(T.Gref + i * DD + c);   // [m][l]
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      double la[DIM];   // Σ_n Λ_q[l,n] a_n
#pragma unroll
      for (int l = 0; l < DIM; ++l) {
        double v = 0.0;
#pragma unroll
        for (int n = 0; n < DIM; ++n) v = fma(Lq[b][l + DIM * n], P.a[n], v);
        la[l] = v;
      }
      const double phi = (LAW == LAW_ADV) ? uu[b] : 0.5 * uu[b] * uu[b];
      sPhi[b * NQ + i] = phi;
#pragma unroll
      for (int m = 0; m < DIM; ++m) {
        double v = 0.0;
#pragma unroll
        for (int l = m; l < DIM; ++l) v = fma(gref[m * DIM + l], la[l], v);   // upper triangular
        ha[b][m] = hw * v;
        sG[(m * NB + b) * NQ + i] = ha[b][m] * phi;
      }
    }
  }
  __syncthreads();
  // separable collapsed-face rows of R (see apply_R_t): shared a3-contraction of φ, kept in sR
  // (free until phase 2)
  const int ng = T.R_ng;
  if (ng > 0) {
    SSE_LOOP(idx, NB * ng * N1) {
      const int a2 = idx % N1, g = (idx / N1) % ng, b = idx / (N1 * ng);
      const double* s0 = sPhi + b * NQ + __ldg(T.R_gstart + g) + a2 * N1;
      double acc = 0.0;
#pragma unroll
      for (int a3 = 0; a3 < N1; ++a3) acc = fma(T.R_r3[a3], s0[a3], acc);
      sR[idx] = acc;
    }
    __syncthreads();
  }
  // ---- phase 1: facet nodes: f_f = B J_f (f* − ½ (a·n) (R φ))
  for (int idx = tid; idx < NB * NF; idx += 128) {
    const int j = idx % NF, b = idx / NF;
    const long long k = min(k0 + b, G.N_e - 1);
    const long long gj = k * NF + j;
    double nfv[DIM], sl[2], fs[1];
    const double Jf = __ldcg(G.J_f + gj);
    const int ext = __ldcg(G.toff + gj);
    const double iJf = frcp(Jf);
    double an = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      nfv[m] = __ldcg(G.nJf + gj * DIM + m) * iJf;
      an = fma(P.a[m], nfv[m], an);
    }
    interface_flux<DIM, LAW>(P, 0, u_f, k * NF + j, ext, NF, nfv, sl, fs);
    const int rb = __ldg(T.R_rp + j);
    const int desc = __ldg(T.R_desc + j);
    const int start = desc & 1023, stride = (desc >> 10) & 1023, cnt = (desc >> 20) & 127;
    const double* ph = sPhi + b * NQ + start;
    double rphi = 0.0;
    if (cnt == N1) {
#pragma unroll
      for (int q = 0; q < N1; ++q) rphi = fma(__ldg(T.R_v + rb + q), ph[q * stride], rphi);
    } else if (ng > 0 && cnt == N1 * N1 && stride == 1) {
      const double* t0 = sR + (b * ng + __ldg(T.R_grp + j)) * N1;
#pragma unroll
      for (int a2 = 0; a2 < N1; ++a2) rphi = fma(__ldg(T.R_E + j * N1 + a2), t0[a2], rphi);
    } else if (cnt == N1 * N1 && stride == 1) {
      double part[N1];
#pragma unroll
      for (int q2 = 0; q2 < N1; ++q2) {
        part[q2] = 0.0;
#pragma unroll
        for (int q = 0; q < N1; ++q)
          part[q2] = fma(__ldg(T.R_v + rb + q2 * N1 + q), ph[q2 * N1 + q], part[q2]);
      }
#pragma unroll
      for (int q2 = 0; q2 < N1; ++q2) rphi += part[q2];
    } else {
      for (int q = 0; q < cnt; ++q) rphi = fma(__ldg(T.R_v + rb + q), ph[q * stride], rphi);
    }
    sFf[b * NF + j] = __ldg(T.B + j) * Jf * fma(-0.5 * an, rphi, fs[0]);
  }
  __syncthreads();
  // ---- phase 2: volume terms along the tensor lines + lifting
  if (active) {
    double r[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) r[b] = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      constexpr int s0 = ipow(N1, DIM - 1), s1 = ipow(N1, DIM >= 2 ? DIM - 2 : 0);
      const int stride = (m == 0) ? s0 : (m == 1 ? s1 : 1);
      const int am = (i / stride) % N1;
      const int line0 = i - am * stride;
      const double* Dm = sD + m * N1 * N1;
#pragma unroll
      for (int q = 0; q < N1; ++q) {
        const int jt = line0 + q * stride;
        const double dt = Dm[q * N1 + am];   // D_m[q, a]  (transpose apply)
        const double dd = Dm[am * N1 + q];   // D_m[a, q]
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          r[b] = fma(dt, sG[(m * NB + b) * NQ + jt], r[b]);
          r[b] = fma(-dd * ha[b][m], sPhi[b * NQ + jt], r[b]);
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const int j = __ldg(F.Cj + kk * NQ + i) & 0xffff;
      const double rv = __ldg(F.Rv + kk * NQ + i);
#pragma unroll
      for (int b = 0; b < NB; ++b) r[b] = fma(-rv, sFf[b * NF + j], r[b]);
    }
    if constexpr (!PROJ) {
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (k0 + b < G.N_e) dudt[(k0 + b) * NQ + i] = r[b];
    } else {
#pragma unroll
      for (int b = 0; b < NB; ++b) sR[b * NQ + i] = r[b];
    }
  }
  if constexpr (!PROJ) return;
  __syncthreads();
  // ---- phase 3: dudt = M^-1 V^T r  (the NB elements ride along as NB "components")
  if (DIM == 3 && T.v_kind == V_WARPED && T.mass_kind == MASS_WEIGHT_ADJUSTED) {
    // M^-1 V^T r = V^T (W/J) (V V^T r): fused V V^T pass, scaling, one V^T
    if constexpr (DIM == 3) apply_VtV_t<N1, NB, 1>(vtab(T), sR, sX);
    SSE_LOOP(idx, NB * NQ) {
      const int ii = idx % NQ, b = idx / NQ;
      const long long k = min(k0 + b, G.N_e - 1);
      sR[idx] *= fdiv(__ldg(T.W + ii), __ldcg(G.J_q + k * NQ + ii));
    }
    __syncthreads();
    apply_Vt_t<DIM, N1, NB, 1>(vtab(T), sR, sM, sX);
    store_result(T, G, rk, k0, NB, 1, sM, dudt);
    return;
  }
  apply_Vt_t<DIM, N1, NB, 1>(vtab(T), sR, sM, sX);
  if (T.mass_kind == MASS_DIAGONAL) {
    SSE_LOOP(idx, NB * NQ) {
      const int ii = idx % NQ, b = idx / NQ;
      const long long k = min(k0 + b, G.N_e - 1);
      sM[idx] = fdiv(sM[idx], __ldg(T.W + ii) * __ldcg(G.J_q + k * NQ + ii));
    }
    __syncthreads();
  } else {
    apply_V_t<DIM, N1, NB, 1>(vtab(T), sM, sR, sX);
    SSE_LOOP(idx, NB * NQ) {
      const int ii = idx % NQ, b = idx / NQ;
      const long long k = min(k0 + b, G.N_e - 1);
      sR[idx] *= fdiv(__ldg(T.W + ii), __ldcg(G.J_q + k * NQ + ii));
    }
    __syncthreads();
    apply_Vt_t<DIM, N1, NB, 1>(vtab(T), sR, sM, sX);
  }
  store_result(T, G, rk, k0, NB, 1, sM, dudt);
}

}  // namespace sse
