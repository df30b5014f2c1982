// Shared by the translation units of libsse_b200.so: the handle, the error helpers and the entry
// points of the compile-time specialised launchers.  The library is built from several .cu files
// (sse_b200.cu: C ABI + generic kernels; tu_nodal.cu / tu_fluxdiff.cu / tu_standard.cu: the
// specialised tensor-product kernels, the first two once per dimension) so that they compile in
// parallel; the host-emulation build of the test suite includes them into one unit instead.
#pragma once
#ifdef SSE_HOST_EMU
#include "cuda_emu.h"   // tests/emu: host emulation of the execution model, test builds only
#else
#include <cuda_runtime.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sse_b200.h"
#include "kernels.cuh"
#include "kernels_tensor.cuh"

using namespace sse;

// records the message returned by sse_last_error() (thread-local) and returns -1
int sse_fail(const char* fmt, ...);
#define fail sse_fail

#define CU(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess)                                                            \
      return fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                  cudaGetErrorString(e_));                                            \
  } while (0)

#define SSE_MAX_CHUNKS 64       // bit masks of chunks are 64 bits wide
#define SSE_DEFAULT_CHUNKS 32   // host-buffer pipeline (SSE_B200_HOST_CHUNKS overrides, <= SSE_MAX_CHUNKS)

// Tuning knobs of the scalar standard-form kernels (elements per CTA riding along as components);
// the defaults are the measured optimum, tools/gpu_variants.sh sweeps -D overrides.
#ifndef SSE_STD_NB
#define SSE_STD_NB 2      // k_standard_tensor (loop B); measured 2 / 4 / 6: 1.014 / 1.232 / 1.169 ms
#endif                    // at 196 608 elements (profiles/r2_sweep.md)
#ifndef SSE_NODAL_NB
#define SSE_NODAL_NB 8    // k_nodal_batched (loop A)
#endif

struct sse_handle {
  sse_config cfg{};
  Tables T{};
  Geo G{};
  Phys P{};
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t d2h_stream = nullptr;          // second copy stream: D2H of finished chunks while
                                              // later chunks are still being uploaded
  // host-buffer pipeline: chunk_need[c] = bit mask of the element chunks that hold a neighbour
  // of chunk c (loop B of c may start once loop A of those chunks is enqueued)
  int n_chunk = 1;
  uint64_t chunk_need[SSE_MAX_CHUNKS] = {};
  int64_t chunk_lo[SSE_MAX_CHUNKS + 1] = {};   // chunk c = elements [chunk_lo[c], chunk_lo[c+1])
  cudaEvent_t ev_chunk[SSE_MAX_CHUNKS] = {};
  // host-buffer pipeline: loop A of the chunks runs on its own stream (ev_a[c] = loop A of chunk c
  // done), so that its CTAs fill the tails of the loop-B kernels of earlier chunks
  cudaStream_t a_stream = nullptr;
  cudaEvent_t ev_a[SSE_MAX_CHUNKS] = {};
  int host_a_stream = 1;
  unsigned next_ev = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<void*> allocs;
  int64_t bytes = 0;
  int64_t launches = 0;
  // state / scratch
  double *u = nullptr, *dudt = nullptr, *rk_k = nullptr;
  double *u_q = nullptr, *u_f = nullptr, *q_q = nullptr, *q_f = nullptr;
  int64_t n_state = 0, halo_elems = 0;
  // halo
  int* send_off = nullptr;
  double *send_buf = nullptr, *recv_buf = nullptr;
  int64_t n_send = 0;
  // launch configuration
  int second_order = 0, proj = 0, law_t = 0;
  double* erk_k = nullptr;      // sse_erk_step: stage derivatives [erk_stages][n_state]
  double* erk_u = nullptr;      //               stage state
  int erk_stages = 0;
  int sm_count = 0;             // SMs of cfg.device (grid / prefetch-distance sizing)
  int prefetch = 1;             // L2 prefetch one wave ahead in the specialised kernels
  int proj_warp = 0;            // projection kernel: one element per warp (k_project_tet_w)
  int r_sep_only = 0;           // rows of R: tensor lines and separable collapsed-face blocks only
  int nodal_rt_proj = 0;        // loop A: keep the run-time projection mode (SSE_B200_NODAL_RT_PROJ=1)
  int phys_staged = 0;          // k_physical: operators staged in shared memory by bulk copies (A/B)
  RK rk_override{};             // sse_shard_rk_stage: the RK epilogue of the range launches
  int use_rk_override = 0;
  int proj_split = 0;           // 1: loop B as k_fluxdiff_nodal + k_project_tet
  int split_b = 0;              // 1: loop B as k_fluxdiff_volume + k_fluxdiff_facet (measurement
                                //    mode of the volume term on its own, SSE_B200_SPLIT_B=1)
  double* r_q = nullptr;        //    nodal residual handed from the volume to the facet kernel
  int split_copy_streams = 0;   // 1: sse_download_dudt_range copies on d2h_stream (sse_set_copy_streams)
  int b_stages = 3;   // second order: bit 0 = auxiliary_variable! (A2), bit 1 = time_derivative!
  int E_a = 1, E_b = 1, thr_a = 128, thr_b = 128;
  size_t smem_a = 0, smem_b = 0;
  // compile-time specialised tensor-product path
  FastTables F{};
  int fast_a = 0, fast_b = 0, n1 = 0, kc = 0, collapsed = 0;
  int const_conflict = 0;
  bool r_ap = false;
  int fast_std = 0;
  std::vector<std::vector<double>> S_dense;
};


// ---- specialised launchers (tu_nodal.cu, tu_fluxdiff.cu: one object per dimension; tu_standard.cu)
int sse_launch_nodal_fast_2d(sse_handle* h, const double* u_dev);
int sse_launch_nodal_fast_3d(sse_handle* h, const double* u_dev);
int sse_launch_fluxdiff_fast_2d(sse_handle* h, double* dudt_dev, const sse::RK& rk);
int sse_launch_fluxdiff_fast_3d(sse_handle* h, double* dudt_dev, const sse::RK& rk);
int sse_launch_standard_fast(sse_handle* h, double* dudt_dev, const sse::RK& rk);
// Without relocatable device code every translation unit owns its copy of the __constant__ tables
// c_wA / c_wB (vmap3.cuh): sse_create uploads the warped-product A and B tables into each of them.
int sse_tu_nodal2_set_constants(const double* A, const double* B, int n);
int sse_tu_nodal3_set_constants(const double* A, const double* B, int n);
int sse_tu_fluxdiff2_set_constants(const double* A, const double* B, int n);
int sse_tu_fluxdiff3_set_constants(const double* A, const double* B, int n);
int sse_tu_standard_set_constants(const double* A, const double* B, int n);
#define SSE_UPLOAD_WARP_CONSTANTS(A, B, n)                                                      \
  CU(cudaMemcpyToSymbol(c_wA, (A), sizeof(double) * (n) * (n), sizeof(double) * 25 * ((n)-3))); \
  CU(cudaMemcpyToSymbol(c_wB, (B), sizeof(double) * (n) * (n) * (n),                            \
                        sizeof(double) * 125 * ((n)-3)));                                       \
  return 0
