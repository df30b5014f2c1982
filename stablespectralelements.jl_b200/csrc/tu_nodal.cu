// Translation unit of the specialised loop-A kernels (k_nodal_tensor / k_nodal_batched), compiled
// once per dimension (-DSSE_TU_DIM=2 / 3) so that the instantiations build in parallel.
#include "handle.h"

#ifndef SSE_TU_DIM
#error "compile with -DSSE_TU_DIM=2 or 3"
#endif

#ifndef SSE_TU_NODAL_TEMPLATE
#define SSE_TU_NODAL_TEMPLATE
template <int DIM, int N1, int LAW>
static int launch_a_fast(sse_handle* h, const double* u_dev) {
  constexpr int EL = NodalCfg<DIM, N1>::E;
  if (TensorNF<DIM, N1, true>::value != h->cfg.N_f)
    return fail("facet-node count does not match the specialised kernel");
  if constexpr (DIM == 3 && LawTraits<DIM, LAW>::NC == 1) {
    if (h->proj == 0) {   // scalar law, no entropy projection: NB elements per CTA as components
      constexpr int NB = SSE_NODAL_NB;
      const size_t smem = NodalBatchCfg<DIM, N1, NB>::bytes(h->cfg.N_p, h->cfg.N_f);
      CU(cudaFuncSetAttribute(k_nodal_batched<DIM, N1, LAW, true, NB>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int grid = (int)((h->G.N_e - h->G.k_begin + NB - 1) / NB);
      k_nodal_batched<DIM, N1, LAW, true, NB> SSE_LAUNCH(grid, 128, smem, h->stream)(h->T, h->G, u_dev,
                                                                             h->u_q, h->u_f);
      h->launches++;
      CU(cudaGetLastError());
      return 0;
    }
  }
  const size_t smem = NodalCfg<DIM, N1>::bytes(h->cfg.N_c, h->cfg.N_p, h->cfg.N_f);
  int grid = (int)((h->G.N_e - h->G.k_begin + EL - 1) / EL);
  h->G.pf_dist = h->prefetch ? h->sm_count * SSE_NODAL_MINB(DIM, N1) * EL : 0;
  if constexpr (LAW == LAW_EULER) {
    // the entropy-projection path of the modal schemes as its own instantiation
    // (3-D: warped-product V, weight-adjusted mass solver, separable collapsed-face rows of R)
    const bool warped_wa = h->T.v_kind == V_WARPED && h->T.mass_kind == MASS_WEIGHT_ADJUSTED &&
                           h->T.R_ng == N1 && h->r_sep_only;
    if (h->proj == 2 && (DIM == 2 || warped_wa) && !h->nodal_rt_proj) {
      CU(cudaFuncSetAttribute(k_nodal_tensor<DIM, N1, LAW, true, 2>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_nodal_tensor<DIM, N1, LAW, true, 2> SSE_LAUNCH(grid, 128, smem, h->stream)(
          h->T, h->G, h->P, u_dev, h->u_q, h->u_f, 2);
      h->launches++;
      CU(cudaGetLastError());
      return 0;
    }
  }
  CU(cudaFuncSetAttribute(k_nodal_tensor<DIM, N1, LAW, true>,
                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_nodal_tensor<DIM, N1, LAW, true> SSE_LAUNCH(grid, 128, smem, h->stream)(
      h->T, h->G, h->P, u_dev, h->u_q, h->u_f, h->proj);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}
#endif

#if SSE_TU_DIM == 2
int sse_launch_nodal_fast_2d(sse_handle* h, const double* u_dev) {
  switch (h->fast_a) {
    case 232: return launch_a_fast<2, 3, LAW_EULER>(h, u_dev);
    case 242: return launch_a_fast<2, 4, LAW_EULER>(h, u_dev);
    case 252: return launch_a_fast<2, 5, LAW_EULER>(h, u_dev);
    case 230: return launch_a_fast<2, 3, LAW_ADV>(h, u_dev);
    case 240: return launch_a_fast<2, 4, LAW_ADV>(h, u_dev);
    case 250: return launch_a_fast<2, 5, LAW_ADV>(h, u_dev);
    default: return fail("no specialised 2-D loop-A kernel for key %d", h->fast_a);
  }
}
int sse_tu_nodal2_set_constants(const double* A, const double* B, int n) {
  SSE_UPLOAD_WARP_CONSTANTS(A, B, n);
}
#else
int sse_launch_nodal_fast_3d(sse_handle* h, const double* u_dev) {
  switch (h->fast_a) {
    case 332: return launch_a_fast<3, 3, LAW_EULER>(h, u_dev);
    case 342: return launch_a_fast<3, 4, LAW_EULER>(h, u_dev);
    case 352: return launch_a_fast<3, 5, LAW_EULER>(h, u_dev);
    case 330: return launch_a_fast<3, 3, LAW_ADV>(h, u_dev);
    case 340: return launch_a_fast<3, 4, LAW_ADV>(h, u_dev);
    case 350: return launch_a_fast<3, 5, LAW_ADV>(h, u_dev);
    default: return fail("no specialised 3-D loop-A kernel for key %d", h->fast_a);
  }
}
int sse_tu_nodal3_set_constants(const double* A, const double* B, int n) {
  SSE_UPLOAD_WARP_CONSTANTS(A, B, n);
}
#endif
