// Translation unit of the specialised standard-form loop-B kernel on collapsed tetrahedra
// (k_standard_tensor: scalar laws, reference operators).
#include "handle.h"

template <int DIM, int N1, int LAW, int KC, int NB>
static int launch_std_fast(sse_handle* h, double* dudt_dev, const RK& rk) {
  using Cf = STCfg<DIM, N1, LAW, KC, NB>;
  if (Cf::NF != h->cfg.N_f) return fail("facet-node count does not match the specialised kernel");
  int grid = (int)((h->G.N_e - h->G.k_begin + NB - 1) / NB);
  if (h->proj_split) {   // nodal residual to r_q, projection on the batched engine
    constexpr int NCOL = 5, GP = SSE_PROJECT_TET_E;          // 25 elements per CTA
    using Pf = ProjectTetCfg<N1, 1, NCOL, GP>;
    const size_t smem = Cf::bytes_nodal();
    CU(cudaFuncSetAttribute(k_standard_tensor<DIM, N1, LAW, KC, NB, false>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaFuncSetAttribute(k_project_tet<N1, 1, NCOL, GP>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Pf::bytes));
    k_standard_tensor<DIM, N1, LAW, KC, NB, false> SSE_LAUNCH(grid, 128, smem, h->stream)(
        h->F, h->T, h->G, h->P, RK{}, h->u_q, h->u_f, h->r_q);
    const int pgrid = (int)((h->G.N_e - h->G.k_begin + Pf::E - 1) / Pf::E);
    k_project_tet<N1, 1, NCOL, GP> SSE_LAUNCH(pgrid, 128, Pf::bytes, h->stream)(h->T, h->G, rk, h->r_q,
                                                                          dudt_dev);
    h->launches += 2;
    CU(cudaGetLastError());
    return 0;
  }
  const size_t smem = Cf::bytes(h->cfg.N_p);
  CU(cudaFuncSetAttribute(k_standard_tensor<DIM, N1, LAW, KC, NB, true>,
                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_standard_tensor<DIM, N1, LAW, KC, NB, true> SSE_LAUNCH(grid, 128, smem, h->stream)(
      h->F, h->T, h->G, h->P, rk, h->u_q, h->u_f, dudt_dev);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int sse_launch_standard_fast(sse_handle* h, double* dudt_dev, const RK& rk) {
  switch (h->fast_std) {   // standard form, advection on collapsed tetrahedra
    case 303: return launch_std_fast<3, 3, LAW_ADV, 6, SSE_STD_NB>(h, dudt_dev, rk);
    case 304: return launch_std_fast<3, 4, LAW_ADV, 7, SSE_STD_NB>(h, dudt_dev, rk);
    case 305: return launch_std_fast<3, 5, LAW_ADV, 8, SSE_STD_NB>(h, dudt_dev, rk);
    default: return fail("no specialised standard-form kernel for key %d", h->fast_std);
  }
}
int sse_tu_standard_set_constants(const double* A, const double* B, int n) {
  SSE_UPLOAD_WARP_CONSTANTS(A, B, n);
}
