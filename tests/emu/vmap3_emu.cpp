// Host emulation of the device stage functions in csrc/vmap3.cuh: every stage is run as a loop
// over the thread index with the same barriers the kernels place between stages.  Test
// infrastructure only (tests/test_vmap3_host.py); validates the index logic without a GPU.
#define SSE_HOST_EMU 1
#include <cmath>
#include <cstring>
#include "../../stablespectralelements.jl_b200/csrc/vmap3b.cuh"

using namespace sse;

template <int N1, int NC, int E>
static void run(int transpose, V3Tab T, double* src, double* dst, double* Z, int nthr) {
  constexpr int EC = E * NC;
  if (transpose == 2) {          // x <- V V^T x in place, Z and Z2 = Z + EC*ZS as scratch
    double* Z2 = Z + EC * V3Dims<N1>::ZS;
    for (int t = 0; t < nthr; ++t) vt3_stageA<N1, EC>(t, nthr, src);
    for (int t = 0; t < nthr; ++t) vt3_stageB<N1, EC>(t, nthr, src, Z);
    for (int t = 0; t < nthr; ++t) vtv3_stageK<N1, NC, E>(t, nthr, T, Z, Z2);
    for (int t = 0; t < nthr; ++t) v3_stageB<N1, EC>(t, nthr, Z2, src);
    for (int t = 0; t < nthr; ++t) v3_stageA<N1, EC>(t, nthr, src);
  } else if (!transpose) {
    for (int t = 0; t < nthr; ++t) v3_stageC<N1, NC, E>(t, nthr, T, src, Z);
    for (int t = 0; t < nthr; ++t) v3_stageB<N1, EC>(t, nthr, Z, dst);
    for (int t = 0; t < nthr; ++t) v3_stageA<N1, EC>(t, nthr, dst);
  } else {
    for (int t = 0; t < nthr; ++t) vt3_stageA<N1, EC>(t, nthr, src);
    for (int t = 0; t < nthr; ++t) vt3_stageB<N1, EC>(t, nthr, src, Z);
    for (int t = 0; t < nthr; ++t) vt3_stageC<N1, EC>(t, nthr, T, Z, dst);
  }
}

extern "C" int vmap3_emu(int n1, int nc, int e, int transpose, const double* wA, const double* wB,
                         const double* wC, const int* sigma, double* src, double* dst, double* Z,
                         int nthr) {
  std::memcpy(c_wA[n1 - 3], wA, sizeof(double) * n1 * n1);
  std::memcpy(c_wB[n1 - 3], wB, sizeof(double) * n1 * n1 * n1);
  V3HostTables ht;
  if (!v3_build_tables(n1, sigma, wC, ht)) return -2;
  V3Tab T{wC, ht.wCt.data(), ht.pairtab.data(), ht.modetab.data(), ht.wK.data()};
#define CASE(N, C, EE) if (n1 == N && nc == C && e == EE) { run<N, C, EE>(transpose, T, src, dst, Z, nthr); return 0; }
  CASE(5, 5, 1) CASE(5, 1, 1) CASE(5, 4, 1) CASE(4, 5, 2) CASE(4, 1, 2) CASE(3, 5, 4) CASE(3, 1, 4)
  CASE(4, 4, 1) CASE(3, 4, 1)
  return -1;
}


// ---- batched, register-tiled engine (csrc/vmap3b.cuh): X [G][NCOL][n^3], M [G][NCOL][NP], Z scratch
template <int N1, int NCOL, int G>
static void run_b(int mode, V3Tab T, double* X, double* M, double* Z, int nthr) {
#define STAGE(call) for (int t = 0; t < nthr; ++t) call
  if (mode == 0) {            // X = V M
    STAGE((vb_stageC<N1, NCOL, G, false>(t, nthr, T, M, Z)));
    STAGE((vb_stageB<N1, NCOL, G, false>(t, nthr, Z, X)));
    STAGE((vb_stageA<N1, NCOL, G, false>(t, nthr, X)));
  } else if (mode == 1) {     // M = V^T X (X destroyed)
    STAGE((vb_stageA<N1, NCOL, G, true>(t, nthr, X)));
    STAGE((vb_stageB<N1, NCOL, G, true>(t, nthr, Z, X)));
    STAGE((vb_stageC<N1, NCOL, G, true>(t, nthr, T, M, Z)));
  } else {                    // X <- V V^T X in place
    STAGE((vb_stageA<N1, NCOL, G, true>(t, nthr, X)));
    STAGE((vb_stageB<N1, NCOL, G, true>(t, nthr, Z, X)));
    STAGE((vb_stageK<N1, NCOL, G>(t, nthr, T, Z)));
    STAGE((vb_stageB<N1, NCOL, G, false>(t, nthr, Z, X)));
    STAGE((vb_stageA<N1, NCOL, G, false>(t, nthr, X)));
  }
#undef STAGE
}

extern "C" int vmap3b_emu(int n1, int ncol, int g, int mode, const double* wA, const double* wB,
                          const double* wC, const int* sigma, double* X, double* M, double* Z,
                          int nthr) {
  std::memcpy(c_wA[n1 - 3], wA, sizeof(double) * n1 * n1);
  std::memcpy(c_wB[n1 - 3], wB, sizeof(double) * n1 * n1 * n1);
  V3HostTables ht;
  if (!v3_build_tables(n1, sigma, wC, ht)) return -2;
  V3Tab T{wC, ht.wCt.data(), ht.pairtab.data(), ht.modetab.data(), ht.wK.data()};
#define CASEB(N, C, GG) if (n1 == N && ncol == C && g == GG) { run_b<N, C, GG>(mode, T, X, M, Z, nthr); return 0; }
  CASEB(5, 5, 4) CASEB(5, 5, 5) CASEB(5, 5, 1) CASEB(4, 5, 6) CASEB(3, 5, 8) CASEB(5, 4, 3) CASEB(4, 2, 5)
  return -1;
}
