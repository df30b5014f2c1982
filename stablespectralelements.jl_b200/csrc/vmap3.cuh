// Sum-factorised warped tensor-product Vandermonde map on the collapsed tetrahedron and its
// transpose (WarpedTensorProductMap3D, /root/reference/src/MatrixFreeOperators/
// warped_product_3d.jl:47-136; tables built by tensor_simplex.jl:84-140):
//   (V x)[a1,a2,a3] = sum_b1 A[a1,b1] sum_{b2<n-b1} B[a2,b1,b2] sum_{b3<n-b1-b2} C[a3,b1,b2,b3]
//                     x[sigma(b1,b2,b3)]
// for EC = (elements x components) columns at once, one CTA, all data in shared memory.
//
// Every stage is written as a function of the thread index so that the same source runs as a
// host loop over tid (tests/test_vmap3_host.py compiles it with g++ and checks it against the
// dense V) and as the device code of the kernels.
//
//  * stage C (ragged in b3): one work item per (pair (b1,b2), a3), all components in registers;
//    sigma(b1,b2,b3) = s0(b1,b2) + b3 (modes are ordered b3-fastest, verified at create), so the
//    modal operands and the C coefficients sit at immediate offsets from one per-thread base.
//    The intermediate Z is stored compactly, [ec][pair][a3] (n^2(n+1)/2 per column).
//  * stage B (ragged in b2): b1 is made WARP-UNIFORM (warp w owns b1 = w, w+4, ..), lanes own
//    (column, a3) and produce all n outputs along a2 from <= n loaded operands; B[a2,b1,b2] are
//    then compile-time indexed constant-bank operands of the DFMAs: 1 LDS per n FMAs.
//  * stage A: one thread per (column, a2, a3) produces all n outputs along a1 IN PLACE
//    (thread-private column), A from the constant bank.
// V^T runs the transposed stages in the opposite order; its stage A runs in place on the source,
// which is therefore destroyed.
#pragma once

#if defined(SSE_HOST_EMU)
#define SSE_HD inline
#define SSE_HD_NOINLINE inline
#define SSE_CX
#define SSE_LDG(p) (*(p))
static double c_wA[3][25];
static double c_wB[3][125];
#else
#define SSE_HD __device__ __forceinline__
#define SSE_HD_NOINLINE __device__ __noinline__
#define SSE_CX __host__ __device__
#define SSE_LDG(p) __ldg(p)
// A[a1][b1] and B[a2][b1][b2] of the warped product, one slot per n = N1 in {3,4,5}
__constant__ double c_wA[3][25];
__constant__ double c_wB[3][125];
#endif

// ---- host side: derived tables (shared by sse_create and the host emulation test)
#include <vector>
struct V3HostTables {
  std::vector<double> wCt;
  std::vector<double> wK;   // [pair][a3'][a3] = sum_b3 C[a3',pair,b3] C[a3,pair,b3]  (V V^T, middle)
  std::vector<int> pairtab, modetab;
};
// sigma [n][n][n] (-1 where unused), wC [a3][b1][b2][b3].  Returns false when the modes are not
// ordered b3-fastest inside each (b1,b2) pair (then the specialised kernels are not used).
inline bool v3_build_tables(int n, const int* sigma, const double* wC, V3HostTables& out) {
  const int T2 = n * (n + 1) / 2, NP = n * (n + 1) * (n + 2) / 6;
  out.pairtab.assign(T2, 0);
  out.modetab.assign(NP, -1);
  out.wCt.assign((size_t)NP * n, 0.0);
  int pr = 0;
  for (int b1 = 0; b1 < n; ++b1)
    for (int b2 = 0; b2 < n - b1; ++b2, ++pr) {
      const int s0 = sigma[(b1 * n + b2) * n];
      if (s0 < 0 || b1 > 15 || b2 > 15) return false;
      out.pairtab[pr] = b1 | (b2 << 4) | (s0 << 8);
      for (int b3 = 0; b3 < n - b1 - b2; ++b3) {
        const int m = sigma[(b1 * n + b2) * n + b3];
        if (m != s0 + b3 || m >= NP) return false;
        out.modetab[m] = pr;
        for (int a3 = 0; a3 < n; ++a3)
          out.wCt[(size_t)m * n + a3] = wC[((a3 * n + b1) * n + b2) * n + b3];
      }
    }
  for (int m = 0; m < NP; ++m)
    if (out.modetab[m] < 0) return false;
  out.wK.assign((size_t)T2 * n * n, 0.0);
  pr = 0;
  for (int b1 = 0; b1 < n; ++b1)
    for (int b2 = 0; b2 < n - b1; ++b2, ++pr)
      for (int a = 0; a < n; ++a)
        for (int a3 = 0; a3 < n; ++a3) {
          double acc = 0.0;
          for (int b3 = 0; b3 < n - b1 - b2; ++b3)
            acc += wC[((a * n + b1) * n + b2) * n + b3] * wC[((a3 * n + b1) * n + b2) * n + b3];
          out.wK[((size_t)pr * n + a) * n + a3] = acc;
        }
  return true;
}

namespace sse {

struct V3Tab {
  const double* wC;    // [a3][b1][b2][b3]
  const double* wCt;   // [mode][a3]  (= C[a3][b1][b2][b3] of that mode)
  const int* pairtab;  // [pair] b1 | b2 << 4 | s0 << 8
  const int* modetab;  // [mode] pair
  const double* wK;    // [pair][a3'][a3]: stage C of V^T followed by stage C of V, fused
};

template <int N1> struct V3Dims {
  static constexpr int N2 = N1 * N1, N3 = N1 * N1 * N1;
  static constexpr int T2 = N1 * (N1 + 1) / 2;          // pairs (b1,b2), b1 + b2 < N1
  static constexpr int ZS = T2 * N1;                    // compact Z per column
  static constexpr int NP = N1 * (N1 + 1) * (N1 + 2) / 6;
  static SSE_CX constexpr int off2(int b1) { return b1 * N1 - b1 * (b1 - 1) / 2; }   // pairs before b1
};

// ---- V, stage C: src [EC][NP] -> Z [EC][ZS].  E elements of NC components each.
template <int N1, int NC, int E>
SSE_HD void v3_stageC(int tid, int nthr, V3Tab T, const double* src, double* Z) {
  using D = V3Dims<N1>;
  for (int idx = tid; idx < E * D::ZS; idx += nthr) {
    const int a3 = idx % N1, pr = (idx / N1) % D::T2, e = idx / D::ZS;
    const int pt = SSE_LDG(T.pairtab + pr);
    const int b1 = pt & 15, b2 = (pt >> 4) & 15, s0 = pt >> 8;
    const int cnt = N1 - b1 - b2;
    const double* cb = T.wC + ((a3 * N1 + b1) * N1 + b2) * N1;
    const double* sb = src + e * NC * D::NP + s0;
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
#pragma unroll
    for (int b3 = 0; b3 < N1; ++b3)
      if (b3 < cnt) {
        const double v = SSE_LDG(cb + b3);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(v, sb[c * D::NP + b3], acc[c]);
      }
    double* zb = Z + e * NC * D::ZS + pr * N1 + a3;
#pragma unroll
    for (int c = 0; c < NC; ++c) zb[c * D::ZS] = acc[c];
  }
}

// ---- V, stage B for one (compile-time) b1: Z -> W [EC][b1][a2][a3] (stored in dst)
template <int N1, int EC, int B1>
SSE_HD void v3_stageB_b1(int lane, const double* Z, double* dst) {
  using D = V3Dims<N1>;
  constexpr int CNT = N1 - B1;
  for (int it = lane; it < EC * N1; it += 32) {
    const int a3 = it % N1, ec = it / N1;
    const double* zb = Z + ec * D::ZS + D::off2(B1) * N1 + a3;
    double z[CNT];
#pragma unroll
    for (int b2 = 0; b2 < CNT; ++b2) z[b2] = zb[b2 * N1];
    double* db = dst + ec * D::N3 + B1 * D::N2 + a3;
#pragma unroll
    for (int a2 = 0; a2 < N1; ++a2) {
      double acc = 0.0;
#pragma unroll
      for (int b2 = 0; b2 < CNT; ++b2) acc = fma(c_wB[N1 - 3][(a2 * N1 + B1) * N1 + b2], z[b2], acc);
      db[a2 * N1] = acc;
    }
  }
}
// OOL = true calls one out-of-line body per b1.  Inlined into the warp switch, the compiler
// hoists the constant loads of ALL cases in front of it (75 LDCU + uniform-register spills per
// warp); measured on B200 the out-of-line form is 1.4 % faster in loop A and 2 % slower in loop B
// (whose 128-register callers pay more for the calls), so each kernel picks its own.
template <int N1, int EC, int B1>
SSE_HD_NOINLINE void v3_stageB_b1_ool(int lane, const double* Z, double* dst) {
  v3_stageB_b1<N1, EC, B1>(lane, Z, dst);
}
template <int N1, int EC, int B1, bool OOL>
SSE_HD void v3_stageB_case(int lane, const double* Z, double* dst) {
  if constexpr (OOL) v3_stageB_b1_ool<N1, EC, B1>(lane, Z, dst);
  else v3_stageB_b1<N1, EC, B1>(lane, Z, dst);
}
template <int N1, int EC, bool OOL = false>
SSE_HD void v3_stageB(int tid, int nthr, const double* Z, double* dst) {
  const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
  // (dealing the b1 values boustrophedon -- {0}, {1}, {2}, {3, 4} for n = 5 and four warps, 25 / 20 /
  // 15 / 15 FMAs per lane instead of 30 / 20 / 15 / 10 -- was measured: no difference in loop A)
  for (int b1 = warp; b1 < N1; b1 += nw) {
    switch (b1) {
      case 0: v3_stageB_case<N1, EC, 0, OOL>(lane, Z, dst); break;
      case 1: v3_stageB_case<N1, EC, 1, OOL>(lane, Z, dst); break;
      case 2: v3_stageB_case<N1, EC, 2, OOL>(lane, Z, dst); break;
      case 3: if constexpr (N1 > 3) v3_stageB_case<N1, EC, 3, OOL>(lane, Z, dst); break;
      case 4: if constexpr (N1 > 4) v3_stageB_case<N1, EC, 4, OOL>(lane, Z, dst); break;
      default: break;
    }
  }
}

// ---- V, stage A, in place on dst: [EC][b1][a2][a3] -> [EC][a1][a2][a3]
template <int N1, int EC>
SSE_HD void v3_stageA(int tid, int nthr, double* dst) {
  using D = V3Dims<N1>;
  for (int idx = tid; idx < EC * D::N2; idx += nthr) {
    const int a23 = idx % D::N2, ec = idx / D::N2;
    double* col = dst + ec * D::N3 + a23;
    double w[N1];
#pragma unroll
    for (int b1 = 0; b1 < N1; ++b1) w[b1] = col[b1 * D::N2];
#pragma unroll
    for (int a1 = 0; a1 < N1; ++a1) {
      double acc = 0.0;
#pragma unroll
      for (int b1 = 0; b1 < N1; ++b1) acc = fma(c_wA[N1 - 3][a1 * N1 + b1], w[b1], acc);
      col[a1 * D::N2] = acc;
    }
  }
}

// ---- V^T, stage A, in place on src: [EC][a1][a2][a3] -> [EC][b1][a2][a3]
template <int N1, int EC>
SSE_HD void vt3_stageA(int tid, int nthr, double* src) {
  using D = V3Dims<N1>;
  for (int idx = tid; idx < EC * D::N2; idx += nthr) {
    const int a23 = idx % D::N2, ec = idx / D::N2;
    double* col = src + ec * D::N3 + a23;
    double x[N1];
#pragma unroll
    for (int a1 = 0; a1 < N1; ++a1) x[a1] = col[a1 * D::N2];
#pragma unroll
    for (int b1 = 0; b1 < N1; ++b1) {
      double acc = 0.0;
#pragma unroll
      for (int a1 = 0; a1 < N1; ++a1) acc = fma(c_wA[N1 - 3][a1 * N1 + b1], x[a1], acc);
      col[b1 * D::N2] = acc;
    }
  }
}

// ---- stage A of V, a nodal scaling, stage A of V^T in one pass, in place on x
// ([EC][b1][a2][a3] -> [EC][b1][a2][a3]): where V^T follows V with only a diagonal matrix between
// them (the W/J of the weight-adjusted mass inverse between the two V V^T passes) the a1-line
// stays in the registers of the thread that owns it -- one shared-memory round trip and two
// barriers instead of three and four.  sc: [E][n^3] scaling at the volume nodes, NC columns per
// element.
template <int N1, int EC, int NC>
SSE_HD void v3_stageA_scale_At(int tid, int nthr, double* x, const double* sc) {
  using D = V3Dims<N1>;
  for (int idx = tid; idx < EC * D::N2; idx += nthr) {
    const int a23 = idx % D::N2, ec = idx / D::N2;
    double* col = x + ec * D::N3 + a23;
    const double* s = sc + (ec / NC) * D::N3 + a23;
    double w[N1], y[N1];
#pragma unroll
    for (int b1 = 0; b1 < N1; ++b1) w[b1] = col[b1 * D::N2];
#pragma unroll
    for (int a1 = 0; a1 < N1; ++a1) {
      double acc = 0.0;
#pragma unroll
      for (int b1 = 0; b1 < N1; ++b1) acc = fma(c_wA[N1 - 3][a1 * N1 + b1], w[b1], acc);
      y[a1] = acc * s[a1 * D::N2];
    }
#pragma unroll
    for (int b1 = 0; b1 < N1; ++b1) {
      double acc = 0.0;
#pragma unroll
      for (int a1 = 0; a1 < N1; ++a1) acc = fma(c_wA[N1 - 3][a1 * N1 + b1], y[a1], acc);
      col[b1 * D::N2] = acc;
    }
  }
}

// ---- V^T, stage B for one b1: W [EC][b1][a2][a3] -> Z [EC][pair][a3]
template <int N1, int EC, int B1>
SSE_HD void vt3_stageB_b1(int lane, const double* W, double* Z) {
  using D = V3Dims<N1>;
  constexpr int CNT = N1 - B1;
  for (int it = lane; it < EC * N1; it += 32) {
    const int a3 = it % N1, ec = it / N1;
    const double* wb = W + ec * D::N3 + B1 * D::N2 + a3;
    double w[N1];
#pragma unroll
    for (int a2 = 0; a2 < N1; ++a2) w[a2] = wb[a2 * N1];
    double* zb = Z + ec * D::ZS + D::off2(B1) * N1 + a3;
#pragma unroll
    for (int b2 = 0; b2 < CNT; ++b2) {
      double acc = 0.0;
#pragma unroll
      for (int a2 = 0; a2 < N1; ++a2) acc = fma(c_wB[N1 - 3][(a2 * N1 + B1) * N1 + b2], w[a2], acc);
      zb[b2 * N1] = acc;
    }
  }
}
template <int N1, int EC, int B1>
SSE_HD_NOINLINE void vt3_stageB_b1_ool(int lane, const double* W, double* Z) {
  vt3_stageB_b1<N1, EC, B1>(lane, W, Z);
}
template <int N1, int EC, int B1, bool OOL>
SSE_HD void vt3_stageB_case(int lane, const double* W, double* Z) {
  if constexpr (OOL) vt3_stageB_b1_ool<N1, EC, B1>(lane, W, Z);
  else vt3_stageB_b1<N1, EC, B1>(lane, W, Z);
}
template <int N1, int EC, bool OOL = false>
SSE_HD void vt3_stageB(int tid, int nthr, const double* W, double* Z) {
  const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
  // (dealing the b1 values boustrophedon -- {0}, {1}, {2}, {3, 4} for n = 5 and four warps, 25 / 20 /
  // 15 / 15 FMAs per lane instead of 30 / 20 / 15 / 10 -- was measured: no difference in loop A)
  for (int b1 = warp; b1 < N1; b1 += nw) {
    switch (b1) {
      case 0: vt3_stageB_case<N1, EC, 0, OOL>(lane, W, Z); break;
      case 1: vt3_stageB_case<N1, EC, 1, OOL>(lane, W, Z); break;
      case 2: vt3_stageB_case<N1, EC, 2, OOL>(lane, W, Z); break;
      case 3: if constexpr (N1 > 3) vt3_stageB_case<N1, EC, 3, OOL>(lane, W, Z); break;
      case 4: if constexpr (N1 > 4) vt3_stageB_case<N1, EC, 4, OOL>(lane, W, Z); break;
      default: break;
    }
  }
}

// ---- V^T, stage C: Z -> dst [EC][NP], one work item per (column, mode)
template <int N1, int EC>
SSE_HD void vt3_stageC(int tid, int nthr, V3Tab T, const double* Z, double* dst) {
  using D = V3Dims<N1>;
  for (int idx = tid; idx < EC * D::NP; idx += nthr) {
    const int m = idx % D::NP, ec = idx / D::NP;
    const int pr = SSE_LDG(T.modetab + m);
    const double* cb = T.wCt + m * N1;
    const double* zb = Z + ec * D::ZS + pr * N1;
    double acc = 0.0;
#pragma unroll
    for (int a3 = 0; a3 < N1; ++a3) acc = fma(SSE_LDG(cb + a3), zb[a3], acc);
    dst[idx] = acc;
  }
}

// ---- V V^T, fused middle: the two ragged b3-contractions around the modal coefficients collapse
// to one n x n matrix per pair, Z2[ec][pair][a3'] = sum_a3 K[pair][a3'][a3] Z[ec][pair][a3].
// One work item per (pair, a3'), all components in registers (as in stage C of V).
template <int N1, int NC, int E>
SSE_HD void vtv3_stageK(int tid, int nthr, V3Tab T, const double* Z, double* Z2) {
  using D = V3Dims<N1>;
  for (int idx = tid; idx < E * D::ZS; idx += nthr) {
    const int a = idx % N1, pr = (idx / N1) % D::T2, e = idx / D::ZS;
    const double* kb = T.wK + (pr * N1 + a) * N1;
    const double* zb = Z + e * NC * D::ZS + pr * N1;
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
#pragma unroll
    for (int a3 = 0; a3 < N1; ++a3) {
      const double v = SSE_LDG(kb + a3);
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = fma(v, zb[c * D::ZS + a3], acc[c]);
    }
    double* ob = Z2 + e * NC * D::ZS + pr * N1 + a;
#pragma unroll
    for (int c = 0; c < NC; ++c) ob[c * D::ZS] = acc[c];
  }
}

}  // namespace sse
