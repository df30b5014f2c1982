// Batched, register-tiled edition of the sum-factorised warped tensor-product Vandermonde map on
// the collapsed tetrahedron (WarpedTensorProductMap3D, /root/reference/src/MatrixFreeOperators/
// warped_product_3d.jl:47-136) for G column groups (= elements) of NCOL columns (= conservative
// variables) each, all data in shared memory:  X [G][NCOL][n^3] nodal,  Z [G][NCOL][ZS] compact
// intermediate ((b1,b2) pairs x a3),  M [G][NCOL][NP] modal.
//
// What differs from vmap3.cuh (one column per work item): a work item owns one tensor line /
// (b1-group, a3) / (b1,b2) pair of ONE group and runs ALL its columns, so that
//  * the n x n coefficient block of the item is fetched once (uniform registers from the constant
//    bank for the stages whose coefficients are the same in every lane; 25 read-only loads for the
//    pair-dependent ones) and reused by NCOL x n x n FMAs -- sm_100 FP64 instructions take no
//    constant-bank operands, so the one-column form pays one LDCU per DFMA;
//  * every operand read from shared memory feeds n FMAs (register tile n inputs -> n outputs);
//  * the ragged b2-contraction is load-balanced by pairing b1 with n - b1 (every item of stage B
//    costs n x n FMAs per column), and the pair stages run in place on Z (K) or between M and Z.
// Several groups per CTA fill the lanes (G x n^2 line items, G x 3 x n group items, G x T2 pairs).
// Every stage is a function of the thread index: the same source runs as a host loop in
// tests/test_vmap3_host.py.
#pragma once
#include "vmap3.cuh"

namespace sse {

// Group stride of the compact intermediate Z in shared memory, padded to 11 (mod 16) doubles: the
// pair stages (lanes = pairs, 5 doubles apart, then groups) and stage B (lanes = a3, then groups)
// then spread a half-warp over all 16 eight-byte bank pairs.
template <int N1, int NCOL> struct VBLayout {
  static constexpr int raw = NCOL * V3Dims<N1>::ZS;
  static constexpr int ZG = raw + ((11 - raw % 16) + 16) % 16;
};

// ---- stage A (V) / A^T (V^T), in place on X: [b1|a1][a2][a3] along the slowest index
template <int N1, int NCOL, int G, bool TRANSPOSE>
SSE_HD void vb_stageA(int tid, int nthr, double* X) {
  using D = V3Dims<N1>;
  for (int it = tid; it < G * D::N2; it += nthr) {
    const int a23 = it % D::N2, g = it / D::N2;
    double* col0 = X + g * NCOL * D::N3 + a23;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      double* col = col0 + c * D::N3;
      double x[N1];
#pragma unroll
      for (int q = 0; q < N1; ++q) x[q] = col[q * D::N2];
#pragma unroll
      for (int o = 0; o < N1; ++o) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < N1; ++q)
          acc = fma(TRANSPOSE ? c_wA[N1 - 3][q * N1 + o] : c_wA[N1 - 3][o * N1 + q], x[q], acc);
        col[o * D::N2] = acc;
      }
    }
  }
}

// ---- A, a nodal scaling, A^T in one pass on the a1-line (see v3_stageA_scale_At in vmap3.cuh).
// sc: [elements of the CTA][n^3]; column c of group g belongs to element (g * NCOL + c) / NC.
template <int N1, int NCOL, int G, int NC>
SSE_HD void vb_stageA_scale_At(int tid, int nthr, double* X, const double* sc) {
  using D = V3Dims<N1>;
  for (int it = tid; it < G * D::N2; it += nthr) {
    const int a23 = it % D::N2, g = it / D::N2;
    double* col0 = X + g * NCOL * D::N3 + a23;
    const double* s0 = sc + ((g * NCOL) / NC) * D::N3 + a23;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      double* col = col0 + c * D::N3;
      // NCOL % NC == 0 or NC % NCOL == 0 with whole elements per CTA: (g*NCOL + c)/NC - (g*NCOL)/NC == c/NC
      const double* s = s0 + (c / NC) * D::N3;
      double x[N1], y[N1];
#pragma unroll
      for (int q = 0; q < N1; ++q) x[q] = col[q * D::N2];
#pragma unroll
      for (int o = 0; o < N1; ++o) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < N1; ++q) acc = fma(c_wA[N1 - 3][o * N1 + q], x[q], acc);
        y[o] = acc * s[o * D::N2];
      }
#pragma unroll
      for (int o = 0; o < N1; ++o) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < N1; ++q) acc = fma(c_wA[N1 - 3][q * N1 + o], y[q], acc);
        col[o * D::N2] = acc;
      }
    }
  }
}

// The same pass with the scaling of the item's own line in REGISTERS (scalar laws: every column is
// another element, so each W/J value is used exactly once; one item per thread, G * n^2 <= nthr).
// sc[c * N1 + a1]: scaling of column c at the a1-th node of the thread's line.
template <int N1, int NCOL, int G>
SSE_HD void vb_stageA_scale_At_regs(int tid, double* X, const double (&sc)[NCOL * N1]) {
  using D = V3Dims<N1>;
  if (tid < G * D::N2) {
    const int a23 = tid % D::N2, g = tid / D::N2;
    double* col0 = X + g * NCOL * D::N3 + a23;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      double* col = col0 + c * D::N3;
      double x[N1], y[N1];
#pragma unroll
      for (int q = 0; q < N1; ++q) x[q] = col[q * D::N2];
#pragma unroll
      for (int o = 0; o < N1; ++o) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < N1; ++q) acc = fma(c_wA[N1 - 3][o * N1 + q], x[q], acc);
        y[o] = acc * sc[c * N1 + o];
      }
#pragma unroll
      for (int o = 0; o < N1; ++o) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < N1; ++q) acc = fma(c_wA[N1 - 3][q * N1 + o], y[q], acc);
        col[o * D::N2] = acc;
      }
    }
  }
}

// ---- stage B: the b2-contraction, ragged in b1 (b2 < n - b1).  Group gi of b1 values:
// gi = 0 -> {0};  gi > 0 -> {gi, n - gi} (one value when they coincide): n x n FMAs per column each
// (n even: the middle group has n x n / 2).
template <int N1> struct VBGroups {
  static constexpr int NG = 1 + N1 / 2;
  static SSE_CX constexpr int second(int gi) { return (gi > 0 && N1 - gi != gi) ? N1 - gi : -1; }
};

// V: Z [pair][a3] -> X [b1][a2][a3] for one compile-time b1, one lane = one (group, a3), one column
template <int N1, int B1>
SSE_HD void vb_B_fwd(const double* zc, double* xc) {
  using D = V3Dims<N1>;
  constexpr int CNT = N1 - B1;
  double z[CNT];
#pragma unroll
  for (int b2 = 0; b2 < CNT; ++b2) z[b2] = zc[(D::off2(B1) + b2) * N1];
#pragma unroll
  for (int a2 = 0; a2 < N1; ++a2) {
    double acc = 0.0;
#pragma unroll
    for (int b2 = 0; b2 < CNT; ++b2) acc = fma(c_wB[N1 - 3][(a2 * N1 + B1) * N1 + b2], z[b2], acc);
    xc[B1 * D::N2 + a2 * N1] = acc;
  }
}
// V^T: X [b1][a2][a3] -> Z [pair][a3]
template <int N1, int B1>
SSE_HD void vb_B_bwd(const double* xc, double* zc) {
  using D = V3Dims<N1>;
  constexpr int CNT = N1 - B1;
  double w[N1];
#pragma unroll
  for (int a2 = 0; a2 < N1; ++a2) w[a2] = xc[B1 * D::N2 + a2 * N1];
#pragma unroll
  for (int b2 = 0; b2 < CNT; ++b2) {
    double acc = 0.0;
#pragma unroll
    for (int a2 = 0; a2 < N1; ++a2) acc = fma(c_wB[N1 - 3][(a2 * N1 + B1) * N1 + b2], w[a2], acc);
    zc[(D::off2(B1) + b2) * N1] = acc;
  }
}
// one out-of-line body per b1 group (inlined into the switch the compiler hoists the constant
// loads of ALL groups in front of it)
template <int N1, int NCOL, int GI, bool TRANSPOSE, int C0 = 0, int C1 = NCOL>
SSE_HD_NOINLINE void vb_stageB_group(double* zg, double* xg) {
  using D = V3Dims<N1>;
  constexpr int B2 = VBGroups<N1>::second(GI);
#pragma unroll
  for (int c = C0; c < C1; ++c) {
    if constexpr (TRANSPOSE) {
      vb_B_bwd<N1, GI>(xg + c * D::N3, zg + c * D::ZS);
      if constexpr (B2 >= 0) vb_B_bwd<N1, B2>(xg + c * D::N3, zg + c * D::ZS);
    } else {
      vb_B_fwd<N1, GI>(zg + c * D::ZS, xg + c * D::N3);
      if constexpr (B2 >= 0) vb_B_fwd<N1, B2>(zg + c * D::ZS, xg + c * D::N3);
    }
  }
}
// SPLIT = 2: the columns of an item are shared by two work items (columns [0, H) and [H, NCOL),
// H = ceil(NCOL / 2); the second halves are the upper half of the index range, so warps stay
// uniform).  The ragged stages have G * 3 * n resp. G * T2 items -- 60 for four Tet p=4 elements,
// two of a CTA's four warps -- and the kernels that use them wait at barriers more than anything
// else (k_project_tet: 3.7 stalled warps per issue): halving the longest item shortens the
// critical path of the stage although the coefficient block is fetched twice.
template <int NCOL> struct VBSplit { static constexpr int H = (NCOL + 1) / 2; };

// items (a3 fastest, then group-of-columns g, then b1 group gi): warps are (nearly) uniform in gi
template <int N1, int NCOL, int G, bool TRANSPOSE, int SPLIT = 1>
SSE_HD void vb_stageB(int tid, int nthr, double* Z, double* X) {
  using D = V3Dims<N1>;
  constexpr int NG = VBGroups<N1>::NG, NI = NG * G * N1, H = VBSplit<NCOL>::H;
  static_assert(SPLIT == 1 || SPLIT == 2, "one or two work items per column set");
  for (int it0 = tid; it0 < SPLIT * NI; it0 += nthr) {
    const int sub = it0 / NI, it = it0 - sub * NI;
    const int a3 = it % N1, g = (it / N1) % G, gi = it / (N1 * G);
    double* zg = Z + g * VBLayout<N1, NCOL>::ZG + a3;
    double* xg = X + g * NCOL * D::N3 + a3;
    if (SPLIT == 1) {
      switch (gi) {
        case 0: vb_stageB_group<N1, NCOL, 0, TRANSPOSE>(zg, xg); break;
        case 1: vb_stageB_group<N1, NCOL, 1, TRANSPOSE>(zg, xg); break;
        case 2: if constexpr (NG > 2) vb_stageB_group<N1, NCOL, 2, TRANSPOSE>(zg, xg); break;
        default: break;
      }
    } else if (sub == 0) {
      switch (gi) {
        case 0: vb_stageB_group<N1, NCOL, 0, TRANSPOSE, 0, H>(zg, xg); break;
        case 1: vb_stageB_group<N1, NCOL, 1, TRANSPOSE, 0, H>(zg, xg); break;
        case 2: if constexpr (NG > 2) vb_stageB_group<N1, NCOL, 2, TRANSPOSE, 0, H>(zg, xg); break;
        default: break;
      }
    } else {
      switch (gi) {
        case 0: vb_stageB_group<N1, NCOL, 0, TRANSPOSE, H, NCOL>(zg, xg); break;
        case 1: vb_stageB_group<N1, NCOL, 1, TRANSPOSE, H, NCOL>(zg, xg); break;
        case 2: if constexpr (NG > 2) vb_stageB_group<N1, NCOL, 2, TRANSPOSE, H, NCOL>(zg, xg); break;
        default: break;
      }
    }
  }
}

// ---- pair stages: one item per (group, (b1,b2) pair), coefficients of the pair in registers
// K: Z <- K[pair] Z in place (the two ragged b3-contractions of V V^T fused, vmap3.cuh)
template <int N1, int NCOL, int G, int SPLIT = 1>
SSE_HD void vb_stageK(int tid, int nthr, V3Tab T, double* Z) {
  using D = V3Dims<N1>;
  constexpr int NI = D::T2 * G, H = VBSplit<NCOL>::H;
  for (int it0 = tid; it0 < SPLIT * NI; it0 += nthr) {
    const int sub = it0 / NI, it = it0 - sub * NI;
    const int pr = it % D::T2, g = it / D::T2;
    double k[N1 * N1];
#pragma unroll
    for (int q = 0; q < N1 * N1; ++q) k[q] = SSE_LDG(T.wK + pr * N1 * N1 + q);
    double* zb = Z + g * VBLayout<N1, NCOL>::ZG + pr * N1;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      if (SPLIT == 2 && ((c < H) != (sub == 0))) continue;
      double z[N1];
#pragma unroll
      for (int a = 0; a < N1; ++a) z[a] = zb[c * D::ZS + a];
#pragma unroll
      for (int o = 0; o < N1; ++o) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < N1; ++a) acc = fma(k[o * N1 + a], z[a], acc);
        zb[c * D::ZS + o] = acc;
      }
    }
  }
}
// C (V): M [mode] -> Z [pair][a3];  C^T (V^T): Z -> M.  cnt = n - b1 - b2 modes per pair.
template <int N1, int NCOL, int G, bool TRANSPOSE, int SPLIT = 1>
SSE_HD void vb_stageC(int tid, int nthr, V3Tab T, double* M, double* Z) {
  using D = V3Dims<N1>;
  constexpr int NI = D::T2 * G, H = VBSplit<NCOL>::H;
  for (int it0 = tid; it0 < SPLIT * NI; it0 += nthr) {
    const int sub = it0 / NI, it = it0 - sub * NI;
    const int pr = it % D::T2, g = it / D::T2;
    const int pt = SSE_LDG(T.pairtab + pr);
    const int b1 = pt & 15, b2 = (pt >> 4) & 15, s0 = pt >> 8;
    const int cnt = N1 - b1 - b2;
    double cf[N1][N1];   // [b3][a3] = C[a3][b1][b2][b3]
#pragma unroll
    for (int b3 = 0; b3 < N1; ++b3)
#pragma unroll
      for (int a3 = 0; a3 < N1; ++a3)
        cf[b3][a3] = (b3 < cnt) ? SSE_LDG(T.wCt + (s0 + b3) * N1 + a3) : 0.0;
    double* mb = M + g * NCOL * D::NP + s0;
    double* zb = Z + g * VBLayout<N1, NCOL>::ZG + pr * N1;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      if (SPLIT == 2 && ((c < H) != (sub == 0))) continue;
      if constexpr (!TRANSPOSE) {
        double m[N1];
#pragma unroll
        for (int b3 = 0; b3 < N1; ++b3) m[b3] = (b3 < cnt) ? mb[c * D::NP + b3] : 0.0;
#pragma unroll
        for (int a3 = 0; a3 < N1; ++a3) {
          double acc = 0.0;
#pragma unroll
          for (int b3 = 0; b3 < N1; ++b3)
            if (b3 < cnt) acc = fma(cf[b3][a3], m[b3], acc);
          zb[c * D::ZS + a3] = acc;
        }
      } else {
        double z[N1];
#pragma unroll
        for (int a3 = 0; a3 < N1; ++a3) z[a3] = zb[c * D::ZS + a3];
#pragma unroll
        for (int b3 = 0; b3 < N1; ++b3)
          if (b3 < cnt) {
            double acc = 0.0;
#pragma unroll
            for (int a3 = 0; a3 < N1; ++a3) acc = fma(cf[b3][a3], z[a3], acc);
            mb[c * D::NP + b3] = acc;
          }
      }
    }
  }
}

}  // namespace sse
