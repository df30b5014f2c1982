// Translation unit of the specialised flux-differencing loop-B kernels (k_fluxdiff_tensor and the
// measurement-only split pair), compiled once per dimension (-DSSE_TU_DIM=2 / 3).
#include "handle.h"

#ifndef SSE_TU_DIM
#error "compile with -DSSE_TU_DIM=2 or 3"
#endif

#ifndef SSE_TU_FLUXDIFF_TEMPLATE
#define SSE_TU_FLUXDIFF_TEMPLATE
template <int DIM, int N1, int LAW, bool COLLAPSED, int KC>
static int launch_b_fast(sse_handle* h, double* dudt_dev, const RK& rk) {
  using Cf = FDCfg<DIM, N1, LAW, COLLAPSED, KC>;
  if (Cf::NF != h->cfg.N_f) return fail("facet-node count does not match the specialised kernel");
  const size_t smem = Cf::bytes(h->cfg.N_p);
  CU(cudaFuncSetAttribute(k_fluxdiff_tensor<DIM, N1, LAW, COLLAPSED, KC>,
                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)((h->G.N_e - h->G.k_begin + Cf::EL - 1) / Cf::EL);
  h->G.pf_dist = h->prefetch ? h->sm_count * SSE_FD_MINB * Cf::EL : 0;
  if constexpr (DIM == 3 && LAW == LAW_EULER) {
    if (h->proj_split && !h->split_b) {   // loop B up to r_q, then the batched projection kernel
      constexpr int EP = SSE_PROJECT_TET_E;
      using Pf = ProjectTetCfg<N1, Cf::NC, Cf::NC, EP>;
      CU(cudaFuncSetAttribute(k_fluxdiff_nodal<DIM, N1, LAW, COLLAPSED, KC>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CU(cudaFuncSetAttribute(k_project_tet<N1, Cf::NC, Cf::NC, EP>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Pf::bytes));
      k_fluxdiff_nodal<DIM, N1, LAW, COLLAPSED, KC> SSE_LAUNCH(grid, 128, smem, h->stream)(
          h->F, h->T, h->G, h->P, h->u_q, h->u_f, h->r_q);
      const int pgrid = (int)((h->G.N_e - h->G.k_begin + EP - 1) / EP);
      if (h->proj_warp) {   // one element per warp, no block barriers
        using Wf = ProjectTetWarpCfg<N1, Cf::NC>;
        CU(cudaFuncSetAttribute(k_project_tet_w<N1, Cf::NC>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Wf::bytes));
        const int wgrid = (int)((h->G.N_e - h->G.k_begin + Wf::WPB * Wf::EW - 1) / (Wf::WPB * Wf::EW));
        k_project_tet_w<N1, Cf::NC> SSE_LAUNCH(wgrid, 128, Wf::bytes, h->stream)(
            h->T, h->G, rk, h->r_q, dudt_dev);
      } else {
        k_project_tet<N1, Cf::NC, Cf::NC, EP> SSE_LAUNCH(pgrid, 128, Pf::bytes, h->stream)(
            h->T, h->G, rk, h->r_q, dudt_dev);
      }
      h->launches += 2;
      CU(cudaGetLastError());
      return 0;
    }
  }
  if constexpr (COLLAPSED) if (h->split_b) {
    const size_t smem_v = Cf::bytes_volume();
    CU(cudaFuncSetAttribute(k_fluxdiff_volume<DIM, N1, LAW, COLLAPSED, KC>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v));
    CU(cudaFuncSetAttribute(k_fluxdiff_facet<DIM, N1, LAW, COLLAPSED, KC>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Geo Gv = h->G;
    if (Gv.pf_dist) Gv.pf_dist = Gv.pf_dist / SSE_FD_MINB * SSE_FD_VOL_MINB;
    k_fluxdiff_volume<DIM, N1, LAW, COLLAPSED, KC> SSE_LAUNCH(grid, 128, smem_v, h->stream)(
        h->F, h->T, Gv, h->P, h->u_q, h->r_q);
    k_fluxdiff_facet<DIM, N1, LAW, COLLAPSED, KC> SSE_LAUNCH(grid, 128, smem, h->stream)(
        h->F, h->T, h->G, h->P, rk, h->u_q, h->u_f, dudt_dev, h->r_q);
    h->launches += 2;
    CU(cudaGetLastError());
    return 0;
  }
  k_fluxdiff_tensor<DIM, N1, LAW, COLLAPSED, KC> SSE_LAUNCH(grid, 128, smem, h->stream)(
      h->F, h->T, h->G, h->P, rk, h->u_q, h->u_f, dudt_dev);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}
#endif

#if SSE_TU_DIM == 2
int sse_launch_fluxdiff_fast_2d(sse_handle* h, double* dudt_dev, const RK& rk) {
  switch (h->fast_b) {
    case 203: return launch_b_fast<2, 3, LAW_EULER, true, 3>(h, dudt_dev, rk);
    case 204: return launch_b_fast<2, 4, LAW_EULER, true, 3>(h, dudt_dev, rk);
    case 205: return launch_b_fast<2, 5, LAW_EULER, true, 3>(h, dudt_dev, rk);
    // diagonal-E collocation on quadrilaterals (NodalTensor LGL, R a selection)
    case 1204: return launch_b_fast<2, 4, LAW_EULER, false, 1>(h, dudt_dev, rk);
    case 1205: return launch_b_fast<2, 5, LAW_EULER, false, 1>(h, dudt_dev, rk);
    default: return fail("no specialised 2-D loop-B kernel for key %d", h->fast_b);
  }
}
int sse_tu_fluxdiff2_set_constants(const double* A, const double* B, int n) {
  SSE_UPLOAD_WARP_CONSTANTS(A, B, n);
}
#else
int sse_launch_fluxdiff_fast_3d(sse_handle* h, double* dudt_dev, const RK& rk) {
  switch (h->fast_b) {
    case 303: return launch_b_fast<3, 3, LAW_EULER, true, 6>(h, dudt_dev, rk);
    case 304: return launch_b_fast<3, 4, LAW_EULER, true, 7>(h, dudt_dev, rk);
    case 305: return launch_b_fast<3, 5, LAW_EULER, true, 8>(h, dudt_dev, rk);
    // diagonal-E collocation on hexahedra (NodalTensor LGL, R a selection)
    case 1304: return launch_b_fast<3, 4, LAW_EULER, false, 1>(h, dudt_dev, rk);
    case 1305: return launch_b_fast<3, 5, LAW_EULER, false, 1>(h, dudt_dev, rk);
    default: return fail("no specialised 3-D loop-B kernel for key %d", h->fast_b);
  }
}
int sse_tu_fluxdiff3_set_constants(const double* A, const double* B, int n) {
  SSE_UPLOAD_WARP_CONSTANTS(A, B, n);
}
#endif
