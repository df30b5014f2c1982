/*
 * sse_b200.h -- C ABI of libsse_b200.so, the B200 (sm_100a) residual engine for
 * StableSpectralElements.jl.
 *
 * One entry point per thing the reference's hot path does (file:line cite the reference,
 * relative to /root/reference/src/):
 *
 *   sse_create            <- Solver(...) constructors, Solvers/Solvers.jl:287-377, including the
 *                            operator bundles of Solvers/operators.jl:1-225 (halfWΛ, n_f, BJf, WJ,
 *                            S, C are derived on the device side from the arrays passed here) and
 *                            the mass solvers of Solvers/mass_matrix.jl:1-136
 *   sse_residual          <- semi_discrete_residual!(dudt, u, solver, t), Solvers/Solvers.jl:455-570
 *   sse_nodal_values      <- loop A: nodal_values!/entropy_projection!,
 *                            Solvers/standard_form_first_order.jl:1-14,
 *                            Solvers/flux_differencing_form.jl:171-292
 *   sse_time_derivative   <- loop B (and A2 for second-order PDEs): time_derivative!,
 *                            auxiliary_variable!, standard_form_first_order.jl:16-94,
 *                            standard_form_second_order.jl:3-75, flux_differencing_form.jl:294-347
 *   sse_rk_stage/_step    <- the low-storage 2N Runge-Kutta update OrdinaryDiffEq applies around
 *                            the residual (test/test_driver.jl:77-83), fused into the residual
 *                            epilogue so the state never leaves the device
 *   sse_halo_*            <- (new) facet-trace halo for element-sharded multi-GPU runs; the only
 *                            inter-element read of the reference is u_f[CI[mapP[:,k]],:]
 *                            (flux_differencing_form.jl:312-313)
 *
 * Conventions: all arrays are Float64 in the reference's (Julia, column-major) memory layout;
 * integer indices are 0-based; every function returns 0 on success and a negative code on
 * failure with a message available from sse_last_error().  Nothing throws across the ABI.
 * One in-flight call per handle (like the reference's Solver, whose scratch is shared).
 */
#ifndef SSE_B200_H
#define SSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sse_handle sse_handle;

enum sse_law { SSE_LAW_ADVECTION = 0, SSE_LAW_BURGERS = 1, SSE_LAW_EULER = 2,
               SSE_LAW_ADVECTION_DIFFUSION = 3, SSE_LAW_VISCOUS_BURGERS = 4 };
enum sse_form { SSE_FORM_STANDARD = 0, SSE_FORM_FLUX_DIFFERENCING = 1 };
enum sse_strategy { SSE_REFERENCE_OPERATOR = 0, SSE_PHYSICAL_OPERATOR = 1 };
enum sse_inviscid_flux { SSE_FLUX_LAX_FRIEDRICHS = 0, SSE_FLUX_CENTRAL = 1,
                         SSE_FLUX_ENTROPY_CONSERVATIVE = 2 };
enum sse_two_point_flux { SSE_TWO_POINT_CONSERVATIVE = 0, SSE_TWO_POINT_ENTROPY_CONSERVATIVE = 1 };
enum sse_v_kind { SSE_V_IDENTITY = 0, SSE_V_WARPED = 1, SSE_V_DENSE = 2 };
enum sse_mass_solver { SSE_MASS_DIAGONAL = 0, SSE_MASS_WEIGHT_ADJUSTED = 1, SSE_MASS_CHOLESKY = 2 };
enum sse_where { SSE_HOST = 0, SSE_DEVICE = 1 };

/* Mirrors the type parameters of Solver{...} (Solvers.jl:259-272) as plain enums. */
typedef struct {
  int32_t dim;            /* d                                                         */
  int32_t N_p, N_q, N_f;  /* modes / volume nodes / facet nodes per element            */
  int32_t N_c;            /* conservative variables                                    */
  int32_t num_faces;      /* faces per element (N_f / num_faces nodes on each)         */
  int64_t N_e;            /* local elements                                            */
  int64_t N_halo;         /* extra trace slots (facet nodes) owned by other ranks      */
  int32_t law;            /* enum sse_law                                              */
  double  a[3];           /* advection velocity / Burgers direction                    */
  double  b;              /* diffusion coefficient                                     */
  double  gamma;          /* Euler specific-heat ratio                                 */
  int32_t form;           /* enum sse_form                                             */
  int32_t strategy;       /* enum sse_strategy                                         */
  int32_t inviscid_flux;  /* enum sse_inviscid_flux                                    */
  double  half_lambda;    /* LaxFriedrichsNumericalFlux.halfλ (ConservationLaws.jl:52) */
  int32_t two_point_flux; /* enum sse_two_point_flux (flux-differencing form only)     */
  int32_t v_kind;         /* enum sse_v_kind                                           */
  int32_t r_is_selection; /* R is a SelectionMap (diag-E): no facet correction         */
  int32_t mass_solver;    /* enum sse_mass_solver                                      */
  int32_t device;         /* CUDA device ordinal                                       */
} sse_config;

/* Reference-element operators (ReferenceApproximation, SpatialDiscretizations.jl:186-246).
 * Sparse operators are CSR (rowptr has rows+1 entries).  Unused pointers may be NULL. */
typedef struct {
  /* V: (N_q x N_p).  DENSE: row-major matrix.  WARPED: tables of
   * MatrixFreeOperators/warped_product_{2d,3d}.jl built by tensor_simplex.jl:84-140:
   * A[a1][b1] (n x n), B[a2][b1][b2] (n^3), C[a3][b1][b2][b3] (n^4, 3-D only),
   * sigma_i[b1][b2]([b3]) (-1 where unused), all C-ordered, n = p + 1 nodes per direction. */
  const double*  V_dense;
  int32_t        n1d;
  const double*  warp_A;
  const double*  warp_B;
  const double*  warp_C;
  const int32_t* sigma_i;
  /* R: (N_f x N_q) interpolation/extrapolation to facet nodes */
  const int32_t* R_rowptr; const int32_t* R_col; const double* R_val;
  /* D_eta[m]: (N_q x N_q) derivative operators in (collapsed) reference coordinates */
  const int32_t* D_rowptr[3]; const int32_t* D_col[3]; const double* D_val[3];
  const double*  W;            /* N_q volume quadrature weights   */
  const double*  B;            /* N_f facet quadrature weights    */
  const double*  Lambda_ref;   /* Λ_ref[i][l][m] = J dη_l/dξ_m, C-ordered (N_q,d,d), or NULL (NoMapping) */
  const double*  J_ref;        /* N_q or NULL                                           */
  const double*  n_ref;        /* (num_faces x d) row-major reference face normals      */
  const double*  Minv;         /* (N_p x N_p) row-major M^-1 of WeightAdjustedSolver, or NULL = I */
} sse_operators;

/* GeometricFactors (SpatialDiscretizations.jl:248-283), Julia memory layout. */
typedef struct {
  const double* J_q;       /* (N_q, N_e)                                                */
  const double* Lambda_q;  /* (N_q, d, d, N_e):  J dξ_l/dx_m at [i, l, m, k]            */
  const double* J_f;       /* (N_f, N_e)                                                */
  const double* nJf;       /* (d, N_f, N_e)                                             */
  /* PhysicalOperators (operators.jl:85-164), only for SSE_PHYSICAL_OPERATOR:             */
  const double* VOL;       /* (N_q, N_p, d, N_e) i.e. VOL[k][m] stored transposed (row-major N_p x N_q) */
  const double* FAC;       /* (N_f, N_p, N_e)    i.e. FAC[k] row-major N_p x N_f        */
  const double* Minv_elem; /* (N_p, N_p, N_e) per-element inverse mass matrix, SSE_MASS_CHOLESKY */
} sse_geometry;

const char* sse_last_error(void);
int  sse_version(void);

/* mapP: (N_f, N_e) 0-based linear index j' + N_f*k' into the (N_f, N_e [+halo]) trace array;
 * indices >= N_f*N_e address halo slots filled by sse_halo_unpack / the exchange. */
int sse_create(const sse_config* cfg, const sse_operators* ops, const sse_geometry* geo,
               const int64_t* mapP, sse_handle** out);
int sse_destroy(sse_handle* h);

/* semi_discrete_residual!(dudt, u, solver, t): u, dudt are (N_p, N_c, N_e).
 * where = SSE_HOST: pageable/pinned host pointers, copies included (u is uploaded into the
 *   handle's device-resident state, i.e. it REPLACES whatever sse_set_state / the RK entry points
 *   left there; the upload is ordered after work still queued on the handle's stream);
 * where = SSE_DEVICE: device pointers on cfg.device (no copies). */
int sse_residual(sse_handle* h, const double* u, double* dudt, double t, int where);

/* The two loops separately (element-sharded runs exchange the halo in between). */
int sse_nodal_values(sse_handle* h, const double* u_dev);
int sse_time_derivative(sse_handle* h, double* dudt_dev);

/* Loop B on the local element range [k_begin, k_end) only (interior/boundary split that
 * overlaps the halo exchange).  First-order equations; for a second-order (BR1) equation a proper
 * sub-range is an error here (time_derivative! would read stale q_f of the elements outside it):
 * use sse_auxiliary_variable_range + sse_time_derivative_only_range below. */
int sse_time_derivative_range(sse_handle* h, double* dudt_dev, int64_t k_begin, int64_t k_end);

/* Device-resident state and fused low-storage (2N) Runge-Kutta:
 *   k <- a*k + dt*R(u);  u <- u + b*k   (evaluated in the residual kernel's epilogue). */
int sse_set_state(sse_handle* h, const double* u_host);
int sse_get_state(sse_handle* h, double* u_host);
int sse_state_ptr(sse_handle* h, double** u_dev, double** dudt_dev);
int sse_rk_stage(sse_handle* h, double a, double b, double dt);
int sse_rk_step_ck54(sse_handle* h, double dt);   /* Carpenter-Kennedy (5,4), 5 fused stages */
/* One step of a general explicit Runge-Kutta scheme on the device-resident state -- what
 * OrdinaryDiffEq's non-low-storage algorithms do around f(du,u,p,t), e.g. DP8 in the reference's
 * 3-D Euler test (test/euler_3d.jl:44-51).  A: n_stages x n_stages row-major, strictly lower
 * triangular; b: n_stages weights; at most 16 stages.  k_s = R(u + dt sum_j A[s][j] k_j),
 * u <- u + dt sum_s b[s] k_s; the stage buffers are allocated on first use. */
int sse_erk_step(sse_handle* h, int n_stages, const double* A, const double* b, double dt);

/* Halo of facet traces for element-sharded runs.  send_idx: n_send linear indices (j + N_f*k)
 * of local trace nodes to pack; the packed buffer holds N_c doubles per index, node-major
 * ([n][c]).  Received values are unpacked into halo slots [0, N_halo). */
int sse_halo_setup(sse_handle* h, const int64_t* send_idx, int64_t n_send);
int sse_halo_buffers(sse_handle* h, double** send_dev, double** recv_dev, int64_t* n_send,
                     int64_t* n_recv);
int sse_halo_pack(sse_handle* h);      /* traces -> send buffer (after sse_nodal_values)     */
int sse_halo_unpack(sse_handle* h);    /* recv buffer -> halo trace slots                     */

/* Second-order (BR1) equations on element shards.  The reference runs three element loops with a
 * barrier after each (Solvers/Solvers.jl:520-570): nodal_values!, auxiliary_variable!
 * (standard_form_second_order.jl:3-40) and time_derivative! (:42-75); the last one reads the
 * neighbour's auxiliary-variable trace q_f, so a sharded run exchanges a second halo in between:
 *   sse_nodal_values -> pack/exchange/unpack u_f -> sse_auxiliary_variable_range (all elements)
 *   -> sse_halo_pack_aux / exchange / sse_halo_unpack_aux -> sse_time_derivative_only_range.
 * The aux halo uses the same send list and buffers, dim * N_c doubles per node ([node][m][c]).
 * sse_time_derivative(_range) keeps running both loops back to back (single-GPU use). */
int sse_auxiliary_variable_range(sse_handle* h, int64_t k_begin, int64_t k_end);
int sse_time_derivative_only_range(sse_handle* h, double* dudt_dev, int64_t k_begin,
                                   int64_t k_end);
int sse_halo_pack_aux(sse_handle* h);
int sse_halo_unpack_aux(sse_handle* h);

/* Streams / timing / introspection. */
/* Host-buffer building blocks for element-sharded runs: chunked H2D of u overlapped with loop A
 * (returns without synchronising); D2H of dudt for an element range, ordered after the work
 * already queued, on the handle's copy stream; sse_sync_copies waits for the copies. */
int sse_upload_and_nodal_values(sse_handle* h, const double* u_host);
int sse_download_dudt_range(sse_handle* h, double* dudt_host, int64_t k_begin, int64_t k_end);
/* H2D of the element range [k_begin, k_end) of u_host (base of the whole local array) on the copy
 * stream + loop A of that range once it has landed; first != 0 on the first range of a residual
 * (orders the copies after the previous residual's readers of u). */
int sse_upload_range_and_nodal_values(sse_handle* h, const double* u_host, int64_t k_begin,
                                      int64_t k_end, int first);
int sse_sync_copies(sse_handle* h);
/* split != 0: sse_download_dudt_range copies on a second copy stream (full-duplex with uploads) */
int sse_set_copy_streams(sse_handle* h, int split);
/* Asynchronous (stream-ordered) H2D of the state / D2H of dudt (the latter synchronises). */
int sse_upload_state(sse_handle* h, const double* u_host);
int sse_download_dudt(sse_handle* h, double* dudt_host);
/* Run all kernels/copies of this handle on an external cudaStream_t (e.g. the stream the host
 * framework orders its NCCL operations against) instead of the handle's own stream. */
int sse_set_stream(sse_handle* h, void* stream);
int sse_sync(sse_handle* h);
void* sse_stream(sse_handle* h);                         /* cudaStream_t the kernels run on */
/* Runs `reps` residuals on the device-resident state, timed with CUDA events on the
 * library's stream; ms[0] = total, ms[1] = loop-A kernels, ms[2] = loop-B kernels (the
 * split is measured in a second pass when split != 0). */
int sse_time_residual(sse_handle* h, int reps, int split, float* ms);
/* FP64 FMA-chain microbenchmark (TFLOP/s, FMA = 2 flops): roofline denominator for the
 * FP64-bound flux-differencing kernel. */
int sse_measure_fp64_peak(int device, double* tflops);
/* The same for FP64 tensor-core instructions (mma.sync m8n8k4, dense 8x8x4 tiles): the measured
 * basis of the decision NOT to use DMMA for the (p+1)-wide contractions. */
int sse_measure_dmma_peak(int device, double* tflops);
/* y[i] = log(x[i]) (which = 0) or exp(x[i]) (which = 1) through the device functions the
 * entropy-variable maps use (euler_navierstokes.jl:100-131 evaluates them with Base.log / exp);
 * host arrays; a test hook for their accuracy. */
int sse_probe_elementary(int device, int which, const double* x, double* y, int64_t n);
int64_t sse_kernel_launches(sse_handle* h);             /* kernels launched so far          */
int64_t sse_device_bytes(sse_handle* h);                /* device memory owned by the handle */

/* ---- on-device geometric factors ----------------------------------------------
 * GeometricFactors(mesh, reference_element, metric_type) of SpatialDiscretizations/mesh.jl:213-509
 * evaluated on the device from the mapping-node coordinates: ExactMetrics (d = 1, 2, 3; optional
 * Jacobian projection, SpatialDiscretizations.jl:311-318) and ConservativeCurlMetrics in 3-D
 * (Tet: curl form on the degree-(N+1) nodes, mesh.jl:413-509; Hex: P = NULL).  The result holds
 * DEVICE pointers; sse_create accepts them in place of host arrays (it copies device to device);
 * release with sse_geometry_free.  All matrices row-major, host or device memory.             */
enum sse_metric_type { SSE_METRIC_EXACT = 0, SSE_METRIC_CONSERVATIVE_CURL = 1 };
typedef struct {
  int32_t dim, metric;
  int32_t N_map, N_map1;     /* mapping nodes of degree N and (curl form, Tet) N+1              */
  int32_t N_q, N_f;
  int64_t N_e;
  const double* D[3];        /* (N_map x N_map) d/dr_n on the mapping nodes (RefElemData.Drst)  */
  const double* Vq;          /* (N_q x N_map) mapping nodes -> volume quadrature nodes          */
  const double* Vf;          /* (N_f x N_map) mapping nodes -> facet quadrature nodes           */
  const double* P;           /* curl: (N_map1 x N_map) degree N -> N+1, NULL = identity         */
  const double* D1[3];       /* curl: (N_map1 x N_map1)                                         */
  const double* Vq1;         /* curl: (N_q x N_map1)                                            */
  const double* Vf1;         /* curl: (N_f x N_map1)                                            */
  const double* nrstJ;       /* (N_f x d) scaled reference normals at the facet nodes           */
  const double* Jproj;       /* exact: (N_q x N_q) L2 projection of J, or NULL                  */
  const double* xyz[3];      /* (N_map, N_e) mapping-node coordinates (mesh.xyz), Julia layout  */
  int32_t device;
} sse_mapping;
int sse_geometry_build(const sse_mapping* m, sse_geometry* out_device_pointers);
int sse_geometry_free(sse_geometry* g);
int sse_copy_to_host(void* dst_host, const void* src_dev, int64_t bytes);

/* ---- analysis functionals on the device --------------------------------------
 * Analysis/conservation.jl:113-190 (evaluate_conservation, evaluate_conservation_residual for the
 * PrimaryConservation / EnergyConservation / EntropyConservation analyses) and
 * Analysis/error.jl:58-91 (L2 error, default error quadrature = volume quadrature), evaluated on
 * the handle's device-resident state u and last residual dudt.  Sums run in a fixed order.
 *   SSE_FN_CONSERVATION      out[e] = sum_k 1^T W J_k V x_k[:,e]; x = u or dudt (arg)   N_c values
 *   SSE_FN_ENTROPY           out[0] = sum_k sum_i W_i J_ki S((V u_k)_i)                 1 value
 *   SSE_FN_ENERGY            out[e] = 1/2 sum_k u_k^T M_k u_k                           N_c values
 *   SSE_FN_ENERGY_RESIDUAL   out[e] = sum_k u_k^T M_k dudt_k                            N_c values
 *   SSE_FN_ENTROPY_RESIDUAL  out[0] = sum_k (P_k w(V u_k))^T M_k dudt_k                 1 value
 *   SSE_FN_L2_ERROR          out[e] = sqrt(sum_k sum_i W_i J_ki (exact - V u)^2);       N_c values
 *                            exact_q_host = exact solution at the volume nodes, (N_q, N_c, N_e)
 * M_k is the mass matrix of the handle's mass solver (mass_matrix.jl:138-151).                */
enum sse_functional_kind {
  SSE_FN_CONSERVATION = 0, SSE_FN_ENTROPY = 1, SSE_FN_ENERGY = 2, SSE_FN_ENERGY_RESIDUAL = 3,
  SSE_FN_ENTROPY_RESIDUAL = 4, SSE_FN_L2_ERROR = 5
};
enum sse_functional_arg { SSE_ARG_STATE = 0, SSE_ARG_DUDT = 1 };
int sse_functional(sse_handle* h, int which, int arg, const double* exact_q_host, double* out);

/* ---- element-sharded multi-GPU residual (one process per GPU, NCCL over NVLink) -------------
 * New relative to the reference (no distributed path there); what a multi-GPU
 * semi_discrete_residual! needs, behind the same ABI so that a host rank makes ONE call per
 * residual.  Rank r of W owns the contiguous element range [N_e r / W, N_e (r+1) / W) of the
 * mesh's element ordering (sse_shard_range) and passes
 *   cfg        with N_e = the GLOBAL element count (N_halo is ignored),
 *   geo_local  the geometric factors of its OWN elements only,
 *   mapP_cols  the columns [start, stop) of the global connectivity: (N_f, n_loc) GLOBAL 0-based
 *              linear indices j' + N_f k' (the reference's mapP, SpatialDiscretizations/mesh.jl),
 *   nccl_id128 the 128-byte ncclUniqueId obtained by ONE rank from sse_nccl_unique_id and
 *              distributed by whatever the host launcher has (MPI_Bcast, a file, torch.distributed).
 * NCCL is loaded at run time (libnccl.so.2); world = 1 needs neither NCCL nor an id.
 * Per residual: loop A -> pack boundary traces -> grouped ncclSend/ncclRecv on a communication
 * stream || loop B on the interior elements -> unpack -> loop B on the boundary elements
 * (second-order equations exchange twice: u_f, then the BR1 auxiliary traces q_f).             */
#define SSE_MAX_PEERS 64
typedef struct sse_shard sse_shard;
typedef struct {
  int64_t start, stop;       /* global element range of the rank                              */
  int64_t n_halo, n_send;    /* trace nodes received into halo slots / packed for sending      */
  int64_t k_lo, k_hi;        /* local elements [k_lo, k_hi) read no halo value (the interior)  */
  int32_t n_peers;
  int32_t peers[SSE_MAX_PEERS];          /* ascending ranks this rank exchanges with           */
  int64_t send_counts[SSE_MAX_PEERS];    /* trace nodes per peer, in packing order             */
  int64_t recv_counts[SSE_MAX_PEERS];
} sse_shard_plan;
int sse_shard_range(int64_t N_e_global, int rank, int world, int64_t* start, int64_t* stop);
/* Pure host arithmetic (no GPU, no communication): the partition of rank `rank`.  mapP_local
 * (N_f * n_loc, out): local connectivity with halo slots numbered from N_f * n_loc, peer by peer
 * and by the owner's global trace index; send_idx (capacity N_f * n_loc, out): the n_send local
 * trace nodes to pack, in the order the receiving rank numbers its halo slots.                  */
int sse_shard_plan_build(const int64_t* mapP_cols, int32_t N_f, int64_t N_e_global, int rank,
                         int world, sse_shard_plan* plan, int64_t* mapP_local, int64_t* send_idx);
int sse_nccl_unique_id(void* id128);
int sse_shard_create(const sse_config* cfg, const sse_operators* ops, const sse_geometry* geo_local,
                     const int64_t* mapP_cols, int rank, int world, const void* nccl_id128,
                     sse_shard** out);
int sse_shard_destroy(sse_shard* s);
/* The rank's own handle: sse_set_state / sse_get_state / sse_functional / sse_state_ptr ... act
 * on the shard (functionals return the rank's partial sums).                                   */
sse_handle* sse_shard_handle(sse_shard* s);
int sse_shard_get_plan(sse_shard* s, sse_shard_plan* plan);
/* semi_discrete_residual! of the shard.  where = SSE_DEVICE: u / dudt are device pointers or
 * NULL (the handle's resident state / dudt buffer), asynchronous; where = SSE_HOST: host buffers
 * of the rank's (N_p, N_c, n_loc) slices, copies inside the call, synchronous.                  */
int sse_shard_residual(sse_shard* s, const double* u, double* dudt, double t, int where);
/* 2N Runge-Kutta on the sharded device-resident state (update fused into loop B's epilogue). */
int sse_shard_rk_stage(sse_shard* s, double a, double b, double dt);
int sse_shard_rk_step_ck54(sse_shard* s, double dt);
/* `reps` device-resident residuals timed with CUDA events on the handle's stream (ms total). */
int sse_shard_time_residual(sse_shard* s, int reps, float* ms);
int sse_shard_sync(sse_shard* s);

#ifdef __cplusplus
}
#endif
#endif /* SSE_B200_H */
