// libsse_b200.so -- C ABI (include/sse_b200.h) over the sm_100a residual kernels.
//
// Host side of the library: packs the reference-element operators and geometric factors into
// device-resident buffers once (north-star part (d)), derives the operator bundles the
// reference builds in Solvers/operators.jl (S, C = R^T B, transposes), and launches the
// element kernels of kernels.cuh.  No CPU fallback: every entry point needs a CUDA device.
#include <map>
#include <mutex>

#include "handle.h"
#include "functionals.cuh"
#include "geometry.cuh"

static thread_local std::string g_err;

int sse_fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return -1;
}

template <typename Tp>
static int dev_upload(sse_handle* h, const Tp* src, size_t n, Tp** out) {
  *out = nullptr;
  if (n == 0) return 0;
  void* p = nullptr;
  CU(cudaMalloc(&p, n * sizeof(Tp)));
  h->allocs.push_back(p);
  h->bytes += (int64_t)(n * sizeof(Tp));
  // cudaMemcpyDefault: the source may be host memory or (sse_geometry_build) device memory
  if (src) CU(cudaMemcpy(p, src, n * sizeof(Tp), cudaMemcpyDefault));
  else CU(cudaMemset(p, 0, n * sizeof(Tp)));
  *out = (Tp*)p;
  return 0;
}

template <typename Tp>
static int dev_upload_vec(sse_handle* h, const std::vector<Tp>& v, const Tp** out) {
  Tp* p = nullptr;
  // keep at least one element so table pointers are never NULL
  if (v.empty()) {
    Tp z{};
    int rc = dev_upload(h, &z, 1, &p);
    *out = p;
    return rc;
  }
  int rc = dev_upload(h, v.data(), v.size(), &p);
  *out = p;
  return rc;
}

struct Csr {
  std::vector<int> rp, ci;
  std::vector<double> v;
};

static Csr csr_from(const int32_t* rp, const int32_t* ci, const double* v, int rows) {
  Csr c;
  c.rp.assign(rp, rp + rows + 1);
  c.ci.assign(ci, ci + rp[rows]);
  c.v.assign(v, v + rp[rows]);
  return c;
}

static Csr csr_transpose(const Csr& a, int rows, int cols) {
  Csr t;
  t.rp.assign(cols + 1, 0);
  for (int c : a.ci) t.rp[c + 1]++;
  for (int i = 0; i < cols; ++i) t.rp[i + 1] += t.rp[i];
  t.ci.resize(a.ci.size());
  t.v.resize(a.v.size());
  std::vector<int> pos(t.rp.begin(), t.rp.end() - 1);
  for (int r = 0; r < rows; ++r)
    for (int e = a.rp[r]; e < a.rp[r + 1]; ++e) {
      int q = pos[a.ci[e]]++;
      t.ci[q] = r;
      t.v[q] = a.v[e];
    }
  return t;
}

static std::vector<double> csr_dense(const Csr& a, int rows, int cols) {
  std::vector<double> d((size_t)rows * cols, 0.0);
  for (int r = 0; r < rows; ++r)
    for (int e = a.rp[r]; e < a.rp[r + 1]; ++e) d[(size_t)r * cols + a.ci[e]] += a.v[e];
  return d;
}

// --------------------------------------------------------------------------- dispatch
template <int DIM, int LAW>
static int launch_a(sse_handle* h, const double* u_dev) {
  CU(cudaFuncSetAttribute(k_nodal_values<DIM, LAW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)h->smem_a));
  int grid = (int)((h->G.N_e - h->G.k_begin + h->E_a - 1) / h->E_a);
  k_nodal_values<DIM, LAW> SSE_LAUNCH(grid, h->thr_a, h->smem_a, h->stream)(h->T, h->G, h->P, u_dev,
                                                                     h->u_q, h->u_f, h->E_a,
                                                                     h->proj);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <int DIM, int LAW>
static int launch_b(sse_handle* h, double* dudt_dev, const RK& rk) {
  int grid = (int)((h->G.N_e - h->G.k_begin + h->E_b - 1) / h->E_b);
  if (h->cfg.strategy == SSE_PHYSICAL_OPERATOR) {
    CU(cudaFuncSetAttribute(k_physical<DIM, LAW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)h->smem_b));
    if (h->second_order && (h->b_stages & 1)) {
      RK none{};
      k_physical<DIM, LAW> SSE_LAUNCH(grid, h->thr_b, h->smem_b, h->stream)(
          h->T, h->G, h->P, none, h->u_q, h->u_f, h->q_q, h->q_f, dudt_dev, h->E_b,
          0 | (h->phys_staged ? 16 : 0), 1);
      h->launches++;
      if (!(h->b_stages & 2)) {
        CU(cudaGetLastError());
        return 0;
      }
    }
    k_physical<DIM, LAW> SSE_LAUNCH(grid, h->thr_b, h->smem_b, h->stream)(
        h->T, h->G, h->P, rk, h->u_q, h->u_f, h->q_q, h->q_f, dudt_dev, h->E_b,
        1 | (h->phys_staged ? 16 : 0), h->second_order);
  } else if (h->cfg.form == SSE_FORM_FLUX_DIFFERENCING) {
    CU(cudaFuncSetAttribute(k_fluxdiff<DIM, LAW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)h->smem_b));
    k_fluxdiff<DIM, LAW> SSE_LAUNCH(grid, h->thr_b, h->smem_b, h->stream)(h->T, h->G, h->P, rk, h->u_q,
                                                                   h->u_f, dudt_dev, h->E_b);
  } else {
    CU(cudaFuncSetAttribute(k_standard_ref<DIM, LAW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)h->smem_b));
    k_standard_ref<DIM, LAW> SSE_LAUNCH(grid, h->thr_b, h->smem_b, h->stream)(
        h->T, h->G, h->P, rk, h->u_q, h->u_f, dudt_dev, h->E_b);
  }
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

#define SSE_DISPATCH(fn, ...)                                                        \
  switch (h->cfg.dim * 10 + h->law_t) {                                              \
    case 10: return fn<1, LAW_ADV>(__VA_ARGS__);                                     \
    case 11: return fn<1, LAW_BURGERS>(__VA_ARGS__);                                 \
    case 12: return fn<1, LAW_EULER>(__VA_ARGS__);                                   \
    case 20: return fn<2, LAW_ADV>(__VA_ARGS__);                                     \
    case 21: return fn<2, LAW_BURGERS>(__VA_ARGS__);                                 \
    case 22: return fn<2, LAW_EULER>(__VA_ARGS__);                                   \
    case 30: return fn<3, LAW_ADV>(__VA_ARGS__);                                     \
    case 31: return fn<3, LAW_BURGERS>(__VA_ARGS__);                                 \
    case 32: return fn<3, LAW_EULER>(__VA_ARGS__);                                   \
    default: return fail("unsupported dim/law combination");                         \
  }

// instantiated (dim, n1, law) combinations of the specialised kernels
static int fast_a_key(int dim, int n1, int law) {
  if ((dim == 2 || dim == 3) && n1 >= 3 && n1 <= 5 && (law == LAW_EULER || law == LAW_ADV))
    return dim * 100 + n1 * 10 + law;
  return 0;
}
static int fast_b_key(int dim, int n1, int law, int collapsed, int kc) {
  if (law != LAW_EULER || n1 < 3 || n1 > 5) return 0;
  if (!collapsed)   // kc = -1 marks the diagonal-E selection path (quadrilaterals / hexahedra)
    return (kc == -1 && n1 >= 4 && (dim == 2 || dim == 3)) ? 1000 + dim * 100 + n1 : 0;
  if (dim == 3 && kc == 3 + n1) return 300 + n1;
  if (dim == 2 && kc == 3) return 200 + n1;
  return 0;
}

static int run_a(sse_handle* h, const double* u_dev) {
  if (h->fast_a) return h->cfg.dim == 3 ? sse_launch_nodal_fast_3d(h, u_dev)
                                        : sse_launch_nodal_fast_2d(h, u_dev);
  SSE_DISPATCH(launch_a, h, u_dev);
}
static int run_b(sse_handle* h, double* dudt_dev, const RK& rk) {
  if (h->fast_std) return sse_launch_standard_fast(h, dudt_dev, rk);
  if (h->fast_b) return h->cfg.dim == 3 ? sse_launch_fluxdiff_fast_3d(h, dudt_dev, rk)
                                        : sse_launch_fluxdiff_fast_2d(h, dudt_dev, rk);
  SSE_DISPATCH(launch_b, h, dudt_dev, rk);
}

// ------------------------------------------------------------------------------- API
// ---- analysis functionals (functionals.cuh)
template <int DIM, int LAW>
static int launch_functional(sse_handle* h, int which, const double* xa, const double* xb,
                             double* partial, int n_out) {
  const int Nc = h->cfg.N_c, Np = h->cfg.N_p, Nq = h->cfg.N_q;
  size_t wt = 1;
  for (int m = 0; m < DIM; ++m) wt *= (size_t)std::max(h->T.n1, 1);
  const size_t smem = sizeof(double) * ((size_t)Nc * (6 * (size_t)Np + 2 * (size_t)Nq + 2 * wt) + 32);
  if (smem > 200 * 1024) return fail("element too large for the functional kernel (%zu B)", smem);
  CU(cudaFuncSetAttribute(k_functional<DIM, LAW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)smem));
  Geo G = h->G;
  G.k_begin = 0;
  k_functional<DIM, LAW> SSE_LAUNCH((unsigned)h->cfg.N_e, 128, smem, h->stream)(h->T, G, h->P, which, xa,
                                                                         xb, partial, n_out);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}
static int run_functional(sse_handle* h, int which, const double* xa, const double* xb,
                          double* partial, int n_out) {
  SSE_DISPATCH(launch_functional, h, which, xa, xb, partial, n_out);
}


extern "C" {

const char* sse_last_error(void) { return g_err.c_str(); }
// (negative: a host-emulation test build, which device.load_library refuses to use)
#ifdef SSE_HOST_EMU
int sse_version(void) { return -100; }
#else
int sse_version(void) { return 100; }
#endif

int sse_destroy(sse_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  if (h->a_stream) cudaStreamDestroy(h->a_stream);
  for (auto& e : h->ev_chunk)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_a)
    if (e) cudaEventDestroy(e);
  delete h;
  return 0;
}

static int create_impl(sse_handle* h, const sse_config* cfg, const sse_operators* ops,
                       const sse_geometry* geo, const int64_t* mapP) {
  const int d = cfg->dim, Np = cfg->N_p, Nq = cfg->N_q, Nf = cfg->N_f, Nc = cfg->N_c;
  const int64_t Ne = cfg->N_e;
  if (d < 1 || d > 3) return fail("dim must be 1, 2 or 3");
  if (Ne < 1 || Np < 1 || Nq < 1 || Nf < 1) return fail("empty discretization");
  if (cfg->num_faces < 1 || Nf % cfg->num_faces) return fail("N_f must be a multiple of num_faces");
  int law_t;
  h->second_order = 0;
  switch (cfg->law) {
    case SSE_LAW_ADVECTION: law_t = LAW_ADV; break;
    case SSE_LAW_BURGERS: law_t = LAW_BURGERS; break;
    case SSE_LAW_EULER: law_t = LAW_EULER; break;
    case SSE_LAW_ADVECTION_DIFFUSION: law_t = LAW_ADV; h->second_order = 1; break;
    case SSE_LAW_VISCOUS_BURGERS: law_t = LAW_BURGERS; h->second_order = 1; break;
    default: return fail("unknown conservation law %d", cfg->law);
  }
  h->law_t = law_t;
  const int nc_expected = (law_t == LAW_EULER) ? d + 2 : 1;
  if (Nc != nc_expected) return fail("N_c = %d does not match the conservation law (%d)", Nc, nc_expected);
  if (h->second_order && (cfg->strategy != SSE_PHYSICAL_OPERATOR || cfg->form != SSE_FORM_STANDARD))
    return fail("second-order equations need StandardForm with PhysicalOperators");
  if (cfg->form == SSE_FORM_FLUX_DIFFERENCING && cfg->strategy == SSE_PHYSICAL_OPERATOR)
    return fail("no physical-operator formulation for the flux-differencing form");
  if ((int64_t)Nf * (Ne + (cfg->N_halo + Nf - 1) / Nf) * Nc >= (1LL << 31))
    return fail("trace array too large for 32-bit offsets");
  if (!ops->R_rowptr || !ops->W || !ops->B || !geo->J_q || !geo->Lambda_q || !geo->J_f ||
      !geo->nJf || !mapP)
    return fail("missing operator/geometry arrays");

  CU(cudaSetDevice(cfg->device));
  CU(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
  {
    const char* e = getenv("SSE_B200_PREFETCH");
    h->prefetch = (e && atoi(e) == 0) ? 0 : 1;
    // k_physical: stage the per-element operators in shared memory with bulk copies, or stream
    // their rows from global memory.  Measured on B200 (profiles/r2_ab_log.md): staging wins while
    // an element's operators are small and N_q is not a multiple of 16 (a thread per staged row
    // strides by N_q: 16-way bank conflicts for N_q = 16, 64); streaming wins otherwise.
    if (const char* pw = getenv("SSE_B200_PROJ_WARP")) h->proj_warp = atoi(pw);
    if (const char* rp = getenv("SSE_B200_NODAL_RT_PROJ")) h->nodal_rt_proj = atoi(rp);
    const char* ps = getenv("SSE_B200_PHYS_STAGED");
    const size_t op_bytes = sizeof(double) * ((size_t)cfg->dim * cfg->N_p * cfg->N_q + (size_t)cfg->N_p * cfg->N_f);
    h->phys_staged = ps ? (atoi(ps) == 1) : (cfg->N_q % 16 != 0 && op_bytes <= (size_t)48 * 1024);
  }
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&h->a_stream, cudaStreamNonBlocking));
  for (auto& e : h->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : h->ev_a) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  {
    const char* e = getenv("SSE_B200_HOST_ASTREAM");
    h->host_a_stream = (e && atoi(e) == 0) ? 0 : 1;
  }
  for (auto& e : h->ev) CU(cudaEventCreate(&e));

  Tables& T = h->T;
  T.dim = d; T.N_p = Np; T.N_q = Nq; T.N_f = Nf; T.N_c = Nc;
  T.nfaces = cfg->num_faces; T.npf = Nf / cfg->num_faces; T.n1 = ops->n1d;
  T.v_kind = cfg->v_kind; T.mass_kind = cfg->mass_solver; T.r_is_selection = cfg->r_is_selection;
  T.has_Minv = ops->Minv != nullptr;

  // ---- V
  if (cfg->v_kind == SSE_V_DENSE) {
    if (!ops->V_dense) return fail("V_dense missing");
    std::vector<double> Vt((size_t)Np * Nq);
    for (int i = 0; i < Nq; ++i)
      for (int p = 0; p < Np; ++p) Vt[(size_t)p * Nq + i] = ops->V_dense[(size_t)i * Np + p];
    double* p1; double* p2;
    if (dev_upload(h, ops->V_dense, (size_t)Nq * Np, &p1)) return -1;
    if (dev_upload(h, Vt.data(), Vt.size(), &p2)) return -1;
    T.Vd = p1; T.VdT = p2;
  } else if (cfg->v_kind == SSE_V_WARPED) {
    int n = ops->n1d;
    if (d < 2 || n < 1 || !ops->warp_A || !ops->warp_B || !ops->sigma_i || (d == 3 && !ops->warp_C))
      return fail("warped-product tables missing");
    size_t nd = 1;
    for (int m = 0; m < d; ++m) nd *= n;
    if ((int)nd != Nq) return fail("warped product needs N_q = n1d^dim");
    // the kernels assume the simplex index pattern N2[b1] = n-b1, N3[b1,b2] = n-b1-b2
    for (int b1 = 0; b1 < n; ++b1)
      for (int b2 = 0; b2 < n; ++b2) {
        if (d == 2) {
          bool used = ops->sigma_i[b1 * n + b2] >= 0;
          if (used != (b2 < n - b1)) return fail("sigma_i is not a simplex pattern");
        } else {
          for (int b3 = 0; b3 < n; ++b3) {
            bool used = ops->sigma_i[(b1 * n + b2) * n + b3] >= 0;
            if (used != (b2 < n - b1 && b3 < n - b1 - b2)) return fail("sigma_i is not a simplex pattern");
          }
        }
      }
    double *pA, *pB, *pC = nullptr; int* ps;
    if (dev_upload(h, ops->warp_A, (size_t)n * n, &pA)) return -1;
    if (dev_upload(h, ops->warp_B, (size_t)n * n * n, &pB)) return -1;
    if (d == 3 && dev_upload(h, ops->warp_C, (size_t)n * n * n * n, &pC)) return -1;
    if (dev_upload(h, ops->sigma_i, nd, &ps)) return -1;
    T.wA = pA; T.wB = pB; T.wC = pC; T.sig = ps;
    if (d == 3) {
      // derived tables of the specialised 3-D applies (vmap3.cuh)
      V3HostTables ht;
      if (v3_build_tables(n, ops->sigma_i, ops->warp_C, ht)) {
        if (dev_upload_vec(h, ht.wCt, &T.wCt) || dev_upload_vec(h, ht.pairtab, &T.pairtab) ||
            dev_upload_vec(h, ht.modetab, &T.modetab) || dev_upload_vec(h, ht.wK, &T.wK))
          return -1;
      } else {
        h->const_conflict = 1;   // modes not ordered b3-fastest: generic kernels only
      }
    }
    if (n >= 3 && n <= 5) {
      // constant-bank copies of A and B for the specialised kernels (one slot per n, per device);
      // two handles with different tables for the same n on one device cannot share them
      static std::mutex mtx;
      static std::map<std::pair<int, int>, std::vector<double>> seen;   // (device, n) -> A | B
      std::lock_guard<std::mutex> lock(mtx);
      std::vector<double> cur(ops->warp_A, ops->warp_A + n * n);
      cur.insert(cur.end(), ops->warp_B, ops->warp_B + n * n * n);
      auto it = seen.find({cfg->device, n});
      if (it != seen.end() && it->second != cur) {
        h->const_conflict = 1;
      } else if (it == seen.end()) {
        if (sse_tu_nodal2_set_constants(ops->warp_A, ops->warp_B, n) ||
            sse_tu_nodal3_set_constants(ops->warp_A, ops->warp_B, n) ||
            sse_tu_fluxdiff2_set_constants(ops->warp_A, ops->warp_B, n) ||
            sse_tu_fluxdiff3_set_constants(ops->warp_A, ops->warp_B, n) ||
            sse_tu_standard_set_constants(ops->warp_A, ops->warp_B, n))
          return -1;
        seen[{cfg->device, n}] = cur;
      }
    }
  } else if (cfg->v_kind == SSE_V_IDENTITY) {
    if (Np != Nq) return fail("identity V needs N_p = N_q");
  } else {
    return fail("unknown v_kind");
  }
  if (cfg->mass_solver == SSE_MASS_DIAGONAL && cfg->v_kind != SSE_V_IDENTITY)
    return fail("diagonal mass solver needs a nodal scheme");
  if (T.n1 < 1) T.n1 = 1;

  // ---- sparse reference operators
  Csr R = csr_from(ops->R_rowptr, ops->R_col, ops->R_val, Nf);
  Csr Rt = csr_transpose(R, Nf, Nq);
  std::vector<double> Cv(Rt.v.size());
  for (int i = 0; i < Nq; ++i)
    for (int e = Rt.rp[i]; e < Rt.rp[i + 1]; ++e) Cv[e] = Rt.v[e] * ops->B[Rt.ci[e]];
  std::vector<int> Rslot(R.ci.size());
  for (int j = 0; j < Nf; ++j)
    for (int e = R.rp[j]; e < R.rp[j + 1]; ++e) {
      int i = R.ci[e], slot = -1;
      for (int q = Rt.rp[i]; q < Rt.rp[i + 1]; ++q)
        if (Rt.ci[q] == j) { slot = q; break; }
      Rslot[e] = slot;
    }
  T.nnzRt = (int)Rt.ci.size();
  {
    // arithmetic-progression descriptors of the rows of R (tensor-product elements)
    std::vector<int> desc(Nf, 0);
    bool ap = Nq < 1024;
    for (int j = 0; j < Nf && ap; ++j) {
      int b = R.rp[j], en = R.rp[j + 1], cnt = en - b;
      if (cnt < 1 || cnt > 127) { ap = false; break; }
      int start = R.ci[b], stride = cnt > 1 ? R.ci[b + 1] - R.ci[b] : 0;
      int k = Rslot[b] - Rt.rp[R.ci[b]];
      for (int q = 0; q < cnt; ++q) {
        if (R.ci[b + q] != start + q * stride) ap = false;
        if (Rslot[b + q] - Rt.rp[R.ci[b + q]] != k) ap = false;
      }
      if (stride < 0 || stride > 1023 || k > 31) ap = false;
      desc[j] = start | (stride << 10) | (cnt << 20) | (k << 27);
    }
    h->r_ap = ap;
    if (ap && dev_upload_vec(h, desc, &T.R_desc)) return -1;
    // separable rows: N1^2 contiguous entries that factor as E_j[a2] * r3[a3] with a common r3
    T.R_ng = 0;
    const int n1r = ops->n1d;
    if (ap && n1r >= 2 && n1r <= 8) {
      std::vector<int> blocks, gstart, grp(Nf, 0);
      for (int j = 0; j < Nf; ++j) {
        const int cnt = R.rp[j + 1] - R.rp[j];
        if (cnt == n1r * n1r && cnt > n1r && ((desc[j] >> 10) & 1023) == 1) blocks.push_back(j);
      }
      bool sep = !blocks.empty();
      std::vector<double> r3(8, 0.0), RE((size_t)Nf * n1r, 0.0);
      if (sep) {
        const double* M0 = R.v.data() + R.rp[blocks[0]];
        int piv = 0;
        for (int q = 1; q < n1r * n1r; ++q)
          if (std::fabs(M0[q]) > std::fabs(M0[piv])) piv = q;
        const int a2p = piv / n1r, a3p = piv % n1r;
        for (int a3 = 0; a3 < n1r; ++a3) r3[a3] = M0[a2p * n1r + a3] / M0[piv];
        for (int j : blocks) {
          const double* M = R.v.data() + R.rp[j];
          double mx = 0.0;
          for (int q = 0; q < n1r * n1r; ++q) mx = std::max(mx, std::fabs(M[q]));
          for (int a2 = 0; a2 < n1r; ++a2) {
            const double e = M[a2 * n1r + a3p];
            RE[(size_t)j * n1r + a2] = e;
            for (int a3 = 0; a3 < n1r; ++a3)
              if (std::fabs(M[a2 * n1r + a3] - e * r3[a3]) > 1e-14 * mx) sep = false;
          }
          const int st = desc[j] & 1023;
          int g = -1;
          for (size_t q = 0; q < gstart.size(); ++q)
            if (gstart[q] == st) g = (int)q;
          if (g < 0) { g = (int)gstart.size(); gstart.push_back(st); }
          grp[j] = g;
        }
        if ((int)gstart.size() > n1r) sep = false;
      }
      if (sep && !getenv("SSE_B200_NO_RSEP")) {
        T.R_ng = (int)gstart.size();
        // every row a tensor line of n1 terms or one of the separable blocks?  (then loop A's
        // entropy-projection instantiation compiles only those two row forms)
        h->r_sep_only = 1;
        for (int j = 0; j < Nf; ++j) {
          const int cnt = R.rp[j + 1] - R.rp[j];
          const bool block = cnt == n1r * n1r && cnt > n1r && ((desc[j] >> 10) & 1023) == 1;
          if (cnt != n1r && !block) h->r_sep_only = 0;
        }
        for (int a3 = 0; a3 < 8; ++a3) T.R_r3[a3] = r3[a3];
        if (dev_upload_vec(h, gstart, &T.R_gstart) || dev_upload_vec(h, grp, &T.R_grp) ||
            dev_upload_vec(h, RE, &T.R_E))
          return -1;
      }
    }
  }
  if (dev_upload_vec(h, R.rp, &T.R_rp) || dev_upload_vec(h, R.ci, &T.R_ci) ||
      dev_upload_vec(h, R.v, &T.R_v) || dev_upload_vec(h, Rslot, &T.R_slot) ||
      dev_upload_vec(h, Rt.rp, &T.Rt_rp) || dev_upload_vec(h, Rt.ci, &T.Rt_ci) ||
      dev_upload_vec(h, Rt.v, &T.Rt_v) || dev_upload_vec(h, Cv, &T.C_v))
    return -1;

  std::vector<double> Gref;
  if (ops->Lambda_ref) {
    if (!ops->J_ref) return fail("J_ref missing");
    Gref.resize((size_t)Nq * d * d);
    for (int i = 0; i < Nq; ++i)
      for (int q = 0; q < d * d; ++q)
        Gref[(size_t)i * d * d + q] = ops->Lambda_ref[(size_t)i * d * d + q] / ops->J_ref[i];
    if (dev_upload_vec(h, Gref, &T.Gref)) return -1;
  }
  const bool need_D = cfg->strategy == SSE_REFERENCE_OPERATOR;
  if (need_D) {
    std::vector<std::vector<double>> Dd(d);
    for (int m = 0; m < d; ++m) {
      if (!ops->D_rowptr[m]) return fail("D operators missing");
      Csr D = csr_from(ops->D_rowptr[m], ops->D_col[m], ops->D_val[m], Nq);
      Csr Dt = csr_transpose(D, Nq, Nq);
      if (dev_upload_vec(h, D.rp, &T.D_rp[m]) || dev_upload_vec(h, D.ci, &T.D_ci[m]) ||
          dev_upload_vec(h, D.v, &T.D_v[m]) || dev_upload_vec(h, Dt.rp, &T.Dt_rp[m]) ||
          dev_upload_vec(h, Dt.ci, &T.Dt_ci[m]) || dev_upload_vec(h, Dt.v, &T.Dt_v[m]))
        return -1;
      Dd[m] = csr_dense(D, Nq, Nq);
    }
    if (cfg->form == SSE_FORM_FLUX_DIFFERENCING) {
      // S_m = ½ (W D_ξm − D_ξm^T W),  D_ξm = Σ_l diag(Λ_ref[:,l,m]/J_ref) D_ηl
      // (operators.jl:186-204, SpatialDiscretizations.jl:414-419)
      std::vector<std::vector<double>> S(d, std::vector<double>((size_t)Nq * Nq, 0.0));
      for (int m = 0; m < d; ++m) {
        std::vector<double> Dxi((size_t)Nq * Nq, 0.0);
        for (int i = 0; i < Nq; ++i)
          for (int j = 0; j < Nq; ++j) {
            double v = 0.0;
            if (Gref.empty()) v = Dd[m][(size_t)i * Nq + j];
            else
              for (int l = 0; l < d; ++l)
                v += Gref[((size_t)i * d + l) * d + m] * Dd[l][(size_t)i * Nq + j];
            Dxi[(size_t)i * Nq + j] = v;
          }
        for (int i = 0; i < Nq; ++i)
          for (int j = 0; j < Nq; ++j)
            S[m][(size_t)i * Nq + j] = 0.5 * (ops->W[i] * Dxi[(size_t)i * Nq + j] -
                                              Dxi[(size_t)j * Nq + i] * ops->W[j]);
      }
      Csr A;
      A.rp.assign(Nq + 1, 0);
      for (int i = 0; i < Nq; ++i) {
        for (int j = 0; j < Nq; ++j) {
          if (i == j) continue;
          bool nz = false;
          for (int m = 0; m < d; ++m)
            nz = nz || S[m][(size_t)i * Nq + j] != 0.0 || S[m][(size_t)j * Nq + i] != 0.0;
          if (nz) {
            A.ci.push_back(j);
            for (int m = 0; m < d; ++m) A.v.push_back(S[m][(size_t)i * Nq + j]);
          }
        }
        A.rp[i + 1] = (int)A.ci.size();
      }
      if (dev_upload_vec(h, A.rp, &T.S_rp) || dev_upload_vec(h, A.ci, &T.S_ci) ||
          dev_upload_vec(h, A.v, &T.S_v))
        return -1;

      h->S_dense = S;   // consumed by the tensor-product fast-path setup below
    }
    // ---- tensor-product fast paths: ELL tables for C = R^T B / R^T, 1-D derivative matrices,
    //      and (flux differencing) the line-pair tables of S
    {
      const int n1 = ops->n1d;
      size_t nd1 = 1;
      for (int m = 0; m < d; ++m) nd1 *= (size_t)std::max(n1, 1);
      int kc = Nq > 0 ? Rt.rp[1] - Rt.rp[0] : 0;
      bool uniform = true;
      for (int i = 0; i < Nq; ++i) uniform = uniform && (Rt.rp[i + 1] - Rt.rp[i] == kc);
      const bool collapsed = !Gref.empty();
      // diagonal-E collocation on quadrilaterals / hexahedra: identity V, selection R (no facet
      // correction, flux_differencing_form.jl:171-187), diagonal mass solve
      const bool sel_path = cfg->r_is_selection && !collapsed && cfg->v_kind == SSE_V_IDENTITY &&
                            cfg->mass_solver == SSE_MASS_DIAGONAL &&
                            cfg->form == SSE_FORM_FLUX_DIFFERENCING;
      size_t npf_sel = 1;
      for (int m = 0; m + 1 < d; ++m) npf_sel *= (size_t)std::max(n1, 1);
      bool ok = d >= 2 && n1 >= 2 && (int)nd1 == Nq &&
                (sel_path ? Nf == (int)(2 * d * npf_sel)
                          : (uniform && kc > 0 && !cfg->r_is_selection && h->r_ap)) &&
                !ops->Minv && Nf < 65536 &&
                (cfg->mass_solver == SSE_MASS_WEIGHT_ADJUSTED || cfg->mass_solver == SSE_MASS_DIAGONAL);
      if (sel_path) kc = 0;
      std::vector<int> stride(d, 1);
      for (int l = 0; l < d; ++l)
        for (int q = l + 1; q < d; ++q) stride[l] *= std::max(n1, 1);
      // D_eta^m must be I (x) D1_m (x) I
      std::vector<double> D1((size_t)d * n1 * n1, 0.0);
      for (int m = 0; m < d && ok; ++m) {
        for (int a = 0; a < n1; ++a)
          for (int bb = 0; bb < n1; ++bb)
            D1[((size_t)m * n1 + a) * n1 + bb] = Dd[m][(size_t)(a * stride[m]) * Nq + bb * stride[m]];
        for (int i = 0; i < Nq && ok; ++i)
          for (int j = 0; j < Nq && ok; ++j) {
            int ndiff = 0;
            for (int l = 0; l < d; ++l)
              if (l != m && (i / stride[l]) % n1 != (j / stride[l]) % n1) ++ndiff;
            double expect = ndiff ? 0.0
                                  : D1[((size_t)m * n1 + (i / stride[m]) % n1) * n1 + (j / stride[m]) % n1];
            if (Dd[m][(size_t)i * Nq + j] != expect) ok = false;
          }
      }
      if (ok) {
        std::vector<int> Cj((size_t)kc * Nq);
        std::vector<double> Cvv((size_t)kc * Nq), Rvv((size_t)kc * Nq);
        for (int i = 0; i < Nq; ++i)
          for (int q = 0; q < kc; ++q) {
            int en = Rt.rp[i] + q;
            Cj[(size_t)q * Nq + i] = Rt.ci[en] | ((Rt.ci[en] / T.npf) << 16);
            Cvv[(size_t)q * Nq + i] = Cv[en];
            Rvv[(size_t)q * Nq + i] = Rt.v[en];
          }
        if (dev_upload_vec(h, Cj, &h->F.Cj) || dev_upload_vec(h, Cvv, &h->F.Cv) ||
            dev_upload_vec(h, Rvv, &h->F.Rv) || dev_upload_vec(h, D1, &h->F.D1))
          return -1;
        // canonical facet layout of the collapsed tensor-product simplices (FastTables): ELL slot
        // -> facet node by formula, and the rows of R as the tensor lines / collapsed-face blocks
        // the specialised kernels sum over
        bool slots_ok = sel_path;
        if (!sel_path) slots_ok = collapsed && (d == 2 || d == 3) && cfg->num_faces == d + 1 &&
                        cfg->num_faces * d <= 12 && ops->n_ref != nullptr &&
                        kc == (d == 3 ? 3 + n1 : 3);
        if (!sel_path) {
          const int npf = d == 3 ? n1 * n1 : n1;
          slots_ok = slots_ok && Nf == (d == 3 ? 4 : 3) * npf;
          for (int i = 0; i < Nq && slots_ok; ++i) {
            const int a1 = i / npf, a2 = (d == 3 ? (i / n1) % n1 : i % n1), a3 = d == 3 ? i % n1 : 0;
            for (int q = 0; q < kc; ++q) {
              int want;
              if (d == 3) want = q >= 3 ? 3 * npf + a1 * n1 + (q - 3)
                                        : q * npf + (q == 0 ? a1 : a2) * n1 + a3;
              else want = q * npf + (q == 0 ? a1 : a2);
              slots_ok = slots_ok && Rt.ci[Rt.rp[i] + q] == want;
            }
          }
          // every row of R: the N1 nodes of its tensor line, or the N1^2 nodes behind a node of
          // the collapsed face -- given the slot check above, the row lengths pin the rest
          for (int j = 0; j < Nf && slots_ok; ++j) {
            const int cnt = R.rp[j + 1] - R.rp[j];
            const bool block_face = d == 3 && j >= 3 * npf;
            slots_ok = slots_ok && cnt == (block_face ? n1 * n1 : n1);
          }
        }
        if (slots_ok && !sel_path)
          for (int q = 0; q < cfg->num_faces * d; ++q) h->F.nref[q] = 0.5 * ops->n_ref[q];   // ½ n_ref
        ok = ok && (slots_ok || cfg->form != SSE_FORM_FLUX_DIFFERENCING);
        h->n1 = n1; h->kc = sel_path ? -1 : kc; h->collapsed = collapsed;
        h->fast_std = (cfg->form == SSE_FORM_STANDARD) ? 1 : 0;
      }
      if (ok && cfg->form == SSE_FORM_FLUX_DIFFERENCING) {
        const std::vector<std::vector<double>>& S = h->S_dense;
        const int H = n1 / 2;
        // every non-zero of S_m must couple two nodes of one tensor line l with an allowed m
        for (int m = 0; m < d && ok; ++m)
          for (int i = 0; i < Nq && ok; ++i)
            for (int j = 0; j < Nq && ok; ++j) {
              if (S[m][(size_t)i * Nq + j] == 0.0) continue;
              int line = -1, ndiff = 0;
              for (int l = 0; l < d; ++l)
                if ((i / stride[l]) % n1 != (j / stride[l]) % n1) { line = l; ++ndiff; }
              if (ndiff != 1 || !(collapsed ? m >= line : m == line)) ok = false;
            }
        if (ok) {
          std::vector<double> Sp((size_t)d * H * d * Nq, 0.0);
          for (int l = 0; l < d; ++l)
            for (int o = 1; o <= H; ++o)
              for (int i = 0; i < Nq; ++i) {
                int al = (i / stride[l]) % n1;
                int ap = (al + o) % n1;
                int j = i + (ap - al) * stride[l];
                for (int m = 0; m < d; ++m)
                  Sp[(((size_t)l * H + (o - 1)) * d + m) * Nq + i] = S[m][(size_t)i * Nq + j];
              }
          if (dev_upload_vec(h, Sp, &h->F.Sp)) return -1;
          h->fast_b = 1;
        }
      }
      h->S_dense.clear();
    }
  }
  {
    double *pW, *pB, *pn;
    if (dev_upload(h, ops->W, Nq, &pW) || dev_upload(h, ops->B, Nf, &pB)) return -1;
    T.W = pW; T.B = pB;
    if (!ops->n_ref) return fail("n_ref missing");
    if (dev_upload(h, ops->n_ref, (size_t)cfg->num_faces * d, &pn)) return -1;
    T.n_ref = pn;
    if (ops->Minv) {
      double* pm;
      if (dev_upload(h, ops->Minv, (size_t)Np * Np, &pm)) return -1;
      T.Minv = pm;
    }
  }

  // ---- geometry (kept in the reference's layout: every element block is contiguous)
  Geo& G = h->G;
  G.N_e = Ne;
  G.k_begin = 0;
  {
    double *p1, *p2, *p3, *p4;
    if (dev_upload(h, geo->J_q, (size_t)Nq * Ne, &p1) ||
        dev_upload(h, geo->Lambda_q, (size_t)Nq * d * d * Ne, &p2) ||
        dev_upload(h, geo->J_f, (size_t)Nf * Ne, &p3) ||
        dev_upload(h, geo->nJf, (size_t)d * Nf * Ne, &p4))
      return -1;
    G.J_q = p1; G.L_q = p2; G.J_f = p3; G.nJf = p4;
  }
  if (cfg->strategy == SSE_PHYSICAL_OPERATOR) {
    if (!geo->VOL || !geo->FAC) return fail("VOL/FAC missing for PhysicalOperator");
    double *p1, *p2;
    if (dev_upload(h, geo->VOL, (size_t)Nq * Np * d * Ne, &p1) ||
        dev_upload(h, geo->FAC, (size_t)Nf * Np * Ne, &p2))
      return -1;
    G.VOL = p1; G.FAC = p2;
  }
  if (cfg->mass_solver == SSE_MASS_CHOLESKY && cfg->strategy != SSE_PHYSICAL_OPERATOR) {
    if (!geo->Minv_elem) return fail("Minv_elem missing for the Cholesky mass solver");
    double* p1;
    if (dev_upload(h, geo->Minv_elem, (size_t)Np * Np * Ne, &p1)) return -1;
    G.Minv_e = p1;
  }
  {
    h->halo_elems = (cfg->N_halo + Nf - 1) / Nf;
    const int64_t ntr = (int64_t)Nf * (Ne + h->halo_elems);
    std::vector<int> toff((size_t)Nf * Ne), mp((size_t)Nf * Ne);
    h->n_chunk = (int)std::min<int64_t>(SSE_DEFAULT_CHUNKS, std::max<int64_t>(1, Ne / 4096));
    if (const char* ce = getenv("SSE_B200_HOST_CHUNKS"))
      h->n_chunk = (int)std::min<int64_t>(std::min<int64_t>(SSE_MAX_CHUNKS, Ne), std::max(1, atoi(ce)));
    if (h->second_order || cfg->N_halo > 0) h->n_chunk = 1;
    const int nch = h->n_chunk;
    {
      // Chunk boundaries of the host-buffer pipeline.  The first chunks are small so that the
      // kernels start after a short upload, the last ones so that little is left to compute and
      // to copy back once the last slice of u has landed: weights 1, 2, 4, 8, 8, ..., 8, 4, 2, 1
      // (SSE_B200_HOST_TAPER=0: equal chunks).
      const char* te = getenv("SSE_B200_HOST_TAPER");
      const bool taper = !(te && atoi(te) == 0) && nch >= 8;
      std::vector<double> w((size_t)nch, 8.0);
      if (taper)
        for (int i = 0; i < 3; ++i) w[i] = w[nch - 1 - i] = (double)(1 << i);
      double tot = 0.0, run = 0.0;
      for (double x : w) tot += x;
      h->chunk_lo[0] = 0;
      for (int c = 0; c < nch; ++c) {
        run += w[c];
        h->chunk_lo[c + 1] = std::max<int64_t>(h->chunk_lo[c] + 1, (int64_t)((double)Ne * (run / tot)));
      }
      h->chunk_lo[nch] = Ne;
      for (int c = nch - 1; c > 0; --c)   // keep the boundaries strictly increasing
        h->chunk_lo[c] = std::min(h->chunk_lo[c], h->chunk_lo[c + 1] - 1);
    }
    auto chunk_of = [&](int64_t k) {
      return (int)(std::upper_bound(h->chunk_lo, h->chunk_lo + nch + 1, k) - h->chunk_lo) - 1;
    };
    for (int64_t g = 0; g < (int64_t)Nf * Ne; ++g) {
      int64_t t = mapP[g];
      if (t < 0 || t >= ntr) return fail("mapP[%lld] = %lld out of range", (long long)g, (long long)t);
      int64_t kp = t / Nf, jp = t % Nf;
      if (kp < Ne) h->chunk_need[chunk_of(g / Nf)] |= 1ull << chunk_of(kp);
      toff[g] = (int)(kp * Nc * Nf + jp);
      mp[g] = (int)t;
    }
    int *p1, *p2;
    if (dev_upload(h, toff.data(), toff.size(), &p1) || dev_upload(h, mp.data(), mp.size(), &p2))
      return -1;
    G.toff = p1; G.mapP = p2;
  }

  // ---- physics
  Phys& P = h->P;
  for (int m = 0; m < 3; ++m) P.a[m] = cfg->a[m];
  P.b = cfg->b; P.gamma = cfg->gamma; P.half_lambda = cfg->half_lambda;
  P.inv_gm1 = 1.0 / (cfg->gamma - 1.0); P.log_gm1 = std::log(cfg->gamma - 1.0);
  P.inviscid = cfg->inviscid_flux;
  P.two_point = (cfg->form == SSE_FORM_FLUX_DIFFERENCING) ? cfg->two_point_flux : 0;

  // ---- state / scratch
  h->n_state = (int64_t)Np * Nc * Ne;
  if (dev_upload<double>(h, nullptr, h->n_state, &h->u) ||
      dev_upload<double>(h, nullptr, h->n_state, &h->dudt) ||
      dev_upload<double>(h, nullptr, h->n_state, &h->rk_k) ||
      dev_upload<double>(h, nullptr, (size_t)Nq * Nc * Ne, &h->u_q) ||
      dev_upload<double>(h, nullptr, (size_t)Nf * Nc * (Ne + h->halo_elems), &h->u_f))
    return -1;
  if (h->second_order) {
    // q_f carries the same halo pseudo-elements as u_f (BR1 needs the neighbour's q trace)
    if (dev_upload<double>(h, nullptr, (size_t)Nq * Nc * d * Ne, &h->q_q) ||
        dev_upload<double>(h, nullptr, (size_t)Nf * Nc * d * (Ne + h->halo_elems), &h->q_f))
      return -1;
  }
  if (cfg->N_halo) {
    // second order: the buffers also carry q_f (d * N_c doubles per node)
    if (dev_upload<double>(h, nullptr, (size_t)cfg->N_halo * Nc * (h->second_order ? d : 1),
                           &h->recv_buf))
      return -1;
  }

  // ---- projection mode of loop A (flux_differencing_form.jl:171-292)
  h->proj = 0;
  if (cfg->form == SSE_FORM_FLUX_DIFFERENCING && Nc > 1) {
    if (cfg->v_kind == SSE_V_IDENTITY) h->proj = cfg->r_is_selection ? 0 : 1;
    else h->proj = 2;
  }

  // ---- launch configuration: E elements per CTA, shared memory per kernel
  int dev_smem = 0;
  CU(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
  const size_t budget = (size_t)std::min(dev_smem, 200 * 1024);
  size_t nd = 1;
  for (int m = 0; m < d; ++m) nd *= T.n1;
  const size_t wtmp = (cfg->v_kind == SSE_V_WARPED) ? 2 * nd : 0;   // per (element, component)
  const int NS = (law_t == LAW_EULER) ? d + 3 : 1;
  auto smem_a = [&](int E) {
    return sizeof(double) * (size_t)E * ((size_t)Nc * Np + 2 * (size_t)Nc * Nq + (size_t)Nc * Nf + wtmp * Nc);
  };
  auto smem_b = [&](int E) -> size_t {
    if (cfg->strategy == SSE_PHYSICAL_OPERATOR)   // (+ the operators VOL[k], FAC[k] when staged)
      return sizeof(double) * ((size_t)E * (d * Nc * ((size_t)2 * Nq + Nf + Np + wtmp) +
                                            (h->phys_staged ? (size_t)d * Np * Nq + (size_t)Np * Nf : 0)) + 4);
    if (cfg->form == SSE_FORM_FLUX_DIFFERENCING) {
      size_t sD = std::max((size_t)T.nnzRt * Nc, wtmp * Nc);
      return sizeof(double) * (size_t)E * ((size_t)NS * Nq + (size_t)d * d * Nq + (size_t)NS * Nf +
                                           (size_t)Nc * Nf + (size_t)Nf * d + (size_t)Nc * Nq +
                                           (size_t)Nc * Np + sD);
    }
    return sizeof(double) * (size_t)E * ((size_t)d * d * Nq + 2 * (size_t)d * Nc * Nq + (size_t)Nc * Nf +
                                         (size_t)Nc * Nq + (size_t)Nc * Np + wtmp * Nc);
  };
  auto pick = [&](auto fn, int* E, int* thr, size_t* sm) -> int {
    int e = std::max(1, std::min(32, 128 / Nq));
    e = (int)std::min<int64_t>(e, Ne);
    while (e > 1 && fn(e) > budget) --e;
    if (fn(e) > budget) return fail("element does not fit in shared memory (%zu B)", fn(e));
    *E = e; *sm = fn(e);
    int t = ((e * Nq + 31) / 32) * 32;
    *thr = std::max(64, std::min(256, t));
    return 0;
  };
  // specialised kernels, when this (dim, n1, law) combination is instantiated
  {
    const int n1 = ops->n1d;
    size_t nd1 = 1;
    for (int m = 0; m < d; ++m) nd1 *= (size_t)std::max(n1, 1);
    const bool mass_ok = !ops->Minv && (cfg->mass_solver == SSE_MASS_WEIGHT_ADJUSTED ||
                                         cfg->mass_solver == SSE_MASS_DIAGONAL);
    const bool force_generic = getenv("SSE_B200_GENERIC") != nullptr || h->const_conflict;
    if (!force_generic && !h->second_order && cfg->strategy == SSE_REFERENCE_OPERATOR &&
        (int)nd1 == Nq && mass_ok && h->r_ap && ops->Lambda_ref != nullptr &&
        Nf == (d == 3 ? 4 * n1 * n1 : 3 * n1))
      h->fast_a = fast_a_key(d, n1, law_t);
    // (the specialised loop B hard-wires the entropy-conservative two-point flux)
    if (h->fast_b && !force_generic && cfg->two_point_flux == SSE_TWO_POINT_ENTROPY_CONSERVATIVE)
      h->fast_b = fast_b_key(d, h->n1, law_t, h->collapsed, h->kc);
    else
      h->fast_b = 0;
    if (h->fast_std && !force_generic && law_t == LAW_ADV && h->collapsed && h->n1 >= 3 &&
        h->n1 <= 5 && cfg->strategy == SSE_REFERENCE_OPERATOR &&
        d == 3 && h->kc == 3 + h->n1)
      h->fast_std = d * 100 + h->n1;
    else
      h->fast_std = 0;
    {   // Collapsed tetrahedra, warped-product V, weight-adjusted M^-1 = I: the projection
        // M^-1 V^T r of loop B runs as its own kernel on the batched engine (k_project_tet);
        // SSE_B200_TET_ENGINE=0 keeps it as the tail of the loop-B kernel (the A/B reference)
      const char* e = getenv("SSE_B200_TET_ENGINE");
      const int want = e ? atoi(e) : 2;
      const bool tet_ok = d == 3 && law_t == LAW_EULER && cfg->v_kind == SSE_V_WARPED &&
                          cfg->mass_solver == SSE_MASS_WEIGHT_ADJUSTED && !ops->Minv &&
                          !h->const_conflict;
      h->proj_split = (tet_ok && h->fast_b && (want & 2)) ? 1 : 0;
      // the scalar standard-form kernel on tetrahedra (config 3) hands its nodal residual to the
      // same projection kernel
      const bool std_ok = d == 3 && cfg->v_kind == SSE_V_WARPED && !ops->Minv && !h->const_conflict &&
                          cfg->mass_solver == SSE_MASS_WEIGHT_ADJUSTED && h->fast_std > 0;
      if (std_ok && (want & 2)) h->proj_split = 1;
      if (h->proj_split && dev_upload<double>(h, nullptr, (size_t)Nq * Nc * Ne, &h->r_q)) return -1;
    }
    {   // measurement mode: loop B as a volume kernel + a facet kernel (see fluxdiff_tensor_body)
      const char* e = getenv("SSE_B200_SPLIT_B");
      if (h->fast_b && e && atoi(e) == 1) {
        if (!h->r_q && dev_upload<double>(h, nullptr, (size_t)Nq * Nc * Ne, &h->r_q)) return -1;
        h->split_b = 1;
      }
    }
  }
  auto smem_a_fast = [&](int E) {
    // upper bound of NodalCfg::bytes (the launch computes the exact figure)
    return sizeof(double) * (size_t)E * ((size_t)Nc * Np + 4 * (size_t)Nc * Nq + (size_t)Nc * Nf);
  };
  auto smem_b_fast = [&](int E) {
    const size_t H = (size_t)h->n1 / 2;
    size_t sX = std::max(std::max(2 * H * Nc, (size_t)((h->kc + 1) / 2) * Nc), (size_t)2 * Nc) * Nq;
    return sizeof(double) * (size_t)E * ((size_t)NS * Nq + (size_t)d * d * Nq + (size_t)NS * Nf +
                                         (size_t)d * Nf + (size_t)Nc * Nf + (size_t)Nc * Nq +
                                         (size_t)Nc * Np + sX);
  };
  if (h->fast_a ? pick(smem_a_fast, &h->E_a, &h->thr_a, &h->smem_a)
                : pick(smem_a, &h->E_a, &h->thr_a, &h->smem_a))
    return -1;
  if (h->fast_b ? pick(smem_b_fast, &h->E_b, &h->thr_b, &h->smem_b)
                : pick(smem_b, &h->E_b, &h->thr_b, &h->smem_b))
    return -1;
  if (cfg->strategy == SSE_PHYSICAL_OPERATOR) {
    int e = h->E_b;
    if (h->phys_staged) {
      // operators staged with bulk copies: CTAs kept small (<= 40 KB: five or more per SM) and E
      // even, which keeps every batch's blocks 16-byte aligned for odd N_p N_q / N_p N_f
      while (e > 1 && (smem_b(e) > (size_t)40 * 1024 || (e & 1))) --e;
    }
    h->E_b = e;
    h->smem_b = smem_b(e);
    h->thr_b = std::max(128, std::min(256, ((e * Nq + 31) / 32) * 32));   // a half-warp per output row
  }
  if (h->fast_b) h->E_b = std::max(1, 128 / Nq);     // FDCfg::EL
  if (h->fast_std) h->E_b = SSE_STD_NB;               // STCfg NB
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int sse_create(const sse_config* cfg, const sse_operators* ops, const sse_geometry* geo,
               const int64_t* mapP, sse_handle** out) {
  if (!cfg || !ops || !geo || !out) return fail("null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device available (libsse_b200 has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("invalid device ordinal %d", cfg->device);
  sse_handle* h = new sse_handle();
  h->cfg = *cfg;
  if (create_impl(h, cfg, ops, geo, mapP)) {
    std::string keep = g_err;
    sse_destroy(h);
    g_err = keep;
    return -1;
  }
  *out = h;
  return 0;
}

int sse_nodal_values(sse_handle* h, const double* u_dev) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  return run_a(h, u_dev ? u_dev : h->u);
}

int sse_time_derivative(sse_handle* h, double* dudt_dev) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  RK rk{};
  return run_b(h, dudt_dev ? dudt_dev : h->dudt, rk);
}

int sse_residual(sse_handle* h, const double* u, double* dudt, double t, int where) {
  (void)t;  // the reference's residual ignores t as well (no source terms are applied)
  if (!h || !u || !dudt) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  RK rk{};
  if (where == SSE_DEVICE) {
    if (run_a(h, u)) return -1;
    return run_b(h, dudt, rk);
  }
  // Host buffers: pipeline the copies against the kernels in element chunks.  Loop A of chunk c
  // starts as soon as its slice of u has landed.  Loop B of chunk b reads the traces of b's
  // neighbours only, so it is enqueued as soon as loop A of every chunk holding such a
  // neighbour (chunk_need, from mapP) has been enqueued -- on a slab-ordered mesh that is one
  // chunk behind the upload front -- and its slice of dudt is copied back on a second copy
  // stream while later chunks are still uploading.  The residual then costs about
  // max(H2D, compute) instead of H2D + loop B.
  // (Asynchronous only for pinned host memory; pageable memory degrades to staged copies.)
  const int64_t Ne = h->cfg.N_e;
  const int64_t blk = (int64_t)h->cfg.N_p * h->cfg.N_c;
  const int nchunk = h->n_chunk;
  auto lo = [&](int c) { return h->chunk_lo[c]; };
  // every exit path leaves the handle's element range whole again
  struct RangeGuard {
    sse_handle* h;
    cudaStream_t main;
    ~RangeGuard() { h->G.k_begin = 0; h->G.N_e = h->cfg.N_e; h->stream = main; }
  } guard{h, h->stream};
  cudaStream_t const main_stream = h->stream;
  cudaStream_t const a_stream = h->host_a_stream ? h->a_stream : main_stream;
  // the upload overwrites the resident state h->u: order it after whatever asynchronous work
  // (sse_rk_stage, sse_erk_step, sse_nodal_values, ...) is still queued on the main stream
  CU(cudaEventRecord(h->ev[3], h->stream));
  CU(cudaStreamWaitEvent(h->copy_stream, h->ev[3], 0));
  // SSE_B200_HOST_TRACE=1: timeline of the pipeline on stderr (timing events per chunk)
  static const bool trace = [] { const char* e = getenv("SSE_B200_HOST_TRACE"); return e && atoi(e) == 1; }();
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tev.push_back(e);
  };
  // SSE_B200_HOST_NOCOPY (diagnostic): the chunked kernel schedule without the copies (1), without
  // the downloads only (2), without the uploads only (3)
  static const int nocopy_mode = [] { const char* e = getenv("SSE_B200_HOST_NOCOPY"); return e ? atoi(e) : 0; }();
  const bool no_h2d = nocopy_mode == 1 || nocopy_mode == 3, no_d2h = nocopy_mode == 1 || nocopy_mode == 2;
  mark(h->copy_stream);                       // [0] start
  for (int c = 0; c < nchunk; ++c) {
    if (!no_h2d)
    CU(cudaMemcpyAsync(h->u + lo(c) * blk, u + lo(c) * blk, (lo(c + 1) - lo(c)) * blk * sizeof(double),
                       cudaMemcpyHostToDevice, h->copy_stream));
    CU(cudaEventRecord(h->ev_chunk[c], h->copy_stream));
    mark(h->copy_stream);                     // [1 + c] H2D of chunk c done
  }
  std::vector<int> b_order;
  int rc = 0;
  uint64_t a_done = 0, b_done = 0, a_seen = 0;   // a_seen: loop-A events the main stream waited for
  for (int c = 0; c < nchunk && !rc; ++c) {
    CU(cudaStreamWaitEvent(a_stream, h->ev_chunk[c], 0));
    h->G.k_begin = lo(c);
    h->G.N_e = lo(c + 1);
    h->stream = a_stream;
    rc = run_a(h, h->u);
    h->stream = main_stream;
    if (a_stream != main_stream) CU(cudaEventRecord(h->ev_a[c], a_stream));
    a_done |= 1ull << c;
    for (int b = 0; b < nchunk && !rc; ++b) {
      if ((b_done >> b) & 1ull) continue;
      const uint64_t need = h->chunk_need[b] | (1ull << b);
      if (need & ~a_done) continue;   // a neighbour chunk is missing
      if (a_stream != main_stream)
        for (int n = 0; n < nchunk; ++n)
          if ((need & ~a_seen) & (1ull << n)) {
            CU(cudaStreamWaitEvent(main_stream, h->ev_a[n], 0));
            a_seen |= 1ull << n;
          }
      h->G.k_begin = lo(b);
      h->G.N_e = lo(b + 1);
      rc = run_b(h, h->dudt, rk);
      if (rc) break;
      b_done |= 1ull << b;
      // ev_chunk[b] was consumed by loop A of chunk b (b <= c), so it can be reused
      CU(cudaEventRecord(h->ev_chunk[b], h->stream));
      CU(cudaStreamWaitEvent(h->d2h_stream, h->ev_chunk[b], 0));
      mark(h->stream);                        // loop B of chunk b done
      if (!no_d2h)
      CU(cudaMemcpyAsync(dudt + lo(b) * blk, h->dudt + lo(b) * blk,
                         (lo(b + 1) - lo(b)) * blk * sizeof(double), cudaMemcpyDeviceToHost,
                         h->d2h_stream));
      mark(h->d2h_stream);                    // D2H of chunk b done
      b_order.push_back(b);
    }
  }
  h->G.k_begin = 0;
  h->G.N_e = Ne;
  if (rc) return -1;
  CU(cudaStreamSynchronize(h->copy_stream));
  CU(cudaStreamSynchronize(h->d2h_stream));
  if (a_stream != main_stream) CU(cudaStreamSynchronize(a_stream));
  CU(cudaStreamSynchronize(h->stream));
  if (trace) {
    auto at = [&](size_t i) { float ms = 0.f; cudaEventElapsedTime(&ms, tev[0], tev[i]); return ms; };
    fprintf(stderr, "[sse host trace] %d chunks; H2D done at:", nchunk);
    for (int c = 0; c < nchunk; ++c) fprintf(stderr, " %.2f", at(1 + c));
    fprintf(stderr, "\n[sse host trace] chunk: loop B done / D2H done at:");
    for (size_t i = 0; i < b_order.size(); ++i)
      fprintf(stderr, " %d: %.2f / %.2f;", b_order[i], at(1 + nchunk + 2 * i), at(2 + nchunk + 2 * i));
    fprintf(stderr, "\n");
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
  }
  return 0;
}

// Host-buffer building blocks for element-sharded runs (the host framework issues the halo
// exchange between them): chunked H2D of u overlapped with loop A ...
int sse_upload_and_nodal_values(sse_handle* h, const double* u_host) {
  if (!h || !u_host) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  const int64_t Ne = h->cfg.N_e;
  const int64_t blk = (int64_t)h->cfg.N_p * h->cfg.N_c;
  const int nchunk = (int)std::min<int64_t>(SSE_DEFAULT_CHUNKS, std::max<int64_t>(1, Ne / 4096));
  auto lo = [&](int c) { return (Ne * c) / nchunk; };
  // the copy stream must not overwrite u while earlier work of the main stream still reads it
  CU(cudaEventRecord(h->ev[3], h->stream));
  CU(cudaStreamWaitEvent(h->copy_stream, h->ev[3], 0));
  for (int c = 0; c < nchunk; ++c) {
    CU(cudaMemcpyAsync(h->u + lo(c) * blk, u_host + lo(c) * blk,
                       (lo(c + 1) - lo(c)) * blk * sizeof(double), cudaMemcpyHostToDevice,
                       h->copy_stream));
    CU(cudaEventRecord(h->ev_chunk[c], h->copy_stream));
  }
  int rc = 0;
  for (int c = 0; c < nchunk && !rc; ++c) {
    CU(cudaStreamWaitEvent(h->stream, h->ev_chunk[c], 0));
    h->G.k_begin = lo(c);
    h->G.N_e = lo(c + 1);
    rc = run_a(h, h->u);
  }
  h->G.k_begin = 0;
  h->G.N_e = Ne;
  return rc;
}

// The same for one element range [k_begin, k_end) of the shard (u_host is the base of the whole
// local array): H2D on the copy stream, loop A of the range on the main stream once it has
// landed.  Lets the host framework choose the upload order (boundary elements first, so that the
// halo exchange starts early) and interleave loop B of ranges whose neighbours are in place.
int sse_upload_range_and_nodal_values(sse_handle* h, const double* u_host, int64_t k_begin,
                                      int64_t k_end, int first) {
  if (!h || !u_host) return fail("null argument");
  if (k_begin < 0 || k_end > h->cfg.N_e || k_begin > k_end) return fail("bad element range");
  if (k_begin == k_end) return 0;
  CU(cudaSetDevice(h->cfg.device));
  const int64_t blk = (int64_t)h->cfg.N_p * h->cfg.N_c;
  if (first) {   // the copy stream must not overwrite u while earlier main-stream work reads it
    CU(cudaEventRecord(h->ev[3], h->stream));
    CU(cudaStreamWaitEvent(h->copy_stream, h->ev[3], 0));
  }
  cudaEvent_t ev = h->ev_chunk[h->next_ev++ % SSE_MAX_CHUNKS];
  CU(cudaMemcpyAsync(h->u + k_begin * blk, u_host + k_begin * blk,
                     (k_end - k_begin) * blk * sizeof(double), cudaMemcpyHostToDevice,
                     h->copy_stream));
  CU(cudaEventRecord(ev, h->copy_stream));
  CU(cudaStreamWaitEvent(h->stream, ev, 0));
  h->G.k_begin = k_begin;
  h->G.N_e = k_end;
  int rc = run_a(h, h->u);
  h->G.k_begin = 0;
  h->G.N_e = h->cfg.N_e;
  return rc;
}

// ... and D2H of dudt for the element range [k_begin, k_end), ordered after the work already
// queued on the main stream, on the copy stream (sse_sync_copies waits for all of them).
int sse_download_dudt_range(sse_handle* h, double* dudt_host, int64_t k_begin, int64_t k_end) {
  if (!h || !dudt_host) return fail("null argument");
  if (k_begin < 0 || k_end > h->cfg.N_e || k_begin > k_end) return fail("bad element range");
  if (k_begin == k_end) return 0;
  CU(cudaSetDevice(h->cfg.device));
  const int64_t blk = (int64_t)h->cfg.N_p * h->cfg.N_c;
  cudaEvent_t ev = h->ev_chunk[h->next_ev++ % SSE_MAX_CHUNKS];
  cudaStream_t cs = h->split_copy_streams ? h->d2h_stream : h->copy_stream;
  CU(cudaEventRecord(ev, h->stream));
  CU(cudaStreamWaitEvent(cs, ev, 0));
  CU(cudaMemcpyAsync(dudt_host + k_begin * blk, h->dudt + k_begin * blk,
                     (k_end - k_begin) * blk * sizeof(double), cudaMemcpyDeviceToHost, cs));
  return 0;
}

int sse_sync_copies(sse_handle* h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->copy_stream));
  if (h->split_copy_streams) CU(cudaStreamSynchronize(h->d2h_stream));
  return 0;
}

// 1: range downloads use the second copy stream, so that a D2H waiting for its loop B does not
// hold back the H2D of later ranges queued behind it (the interleaved host-buffer flow)
int sse_set_copy_streams(sse_handle* h, int split) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->copy_stream));
  CU(cudaStreamSynchronize(h->d2h_stream));
  h->split_copy_streams = split ? 1 : 0;
  return 0;
}

int sse_set_state(sse_handle* h, const double* u_host) {
  if (!h || !u_host) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(h->u, u_host, h->n_state * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(h->rk_k, 0, h->n_state * sizeof(double), h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int sse_get_state(sse_handle* h, double* u_host) {
  if (!h || !u_host) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(u_host, h->u, h->n_state * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int sse_state_ptr(sse_handle* h, double** u_dev, double** dudt_dev) {
  if (!h) return fail("null handle");
  if (u_dev) *u_dev = h->u;
  if (dudt_dev) *dudt_dev = h->dudt;
  return 0;
}

int sse_rk_stage(sse_handle* h, double a, double b, double dt) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  RK rk{1, a, b, dt, h->rk_k, h->u};
  if (run_a(h, h->u)) return -1;
  return run_b(h, h->dudt, rk);
}

int sse_rk_step_ck54(sse_handle* h, double dt) {
  static const double A[5] = {0.0, -567301805773.0 / 1357537059087.0,
                              -2404267990393.0 / 2016746695238.0,
                              -3550918686646.0 / 2091501179385.0,
                              -1275806237668.0 / 842570457699.0};
  static const double B[5] = {1432997174477.0 / 9575080441755.0,
                              5161836677717.0 / 13612068292357.0,
                              1720146321549.0 / 2090206949498.0,
                              3134564353537.0 / 4481467310338.0,
                              2277821191437.0 / 14882151754819.0};
  for (int s = 0; s < 5; ++s)
    if (sse_rk_stage(h, A[s], B[s], dt)) return -1;
  return 0;
}

// General explicit Runge-Kutta step on the device-resident state (the reference integrates with
// OrdinaryDiffEq; its 3-D Euler test uses DP8, test/euler_3d.jl:44-51):
//   k_s = R(u + dt sum_{j<s} A[s][j] k_j),   u <- u + dt sum_s b[s] k_s.
int sse_erk_step(sse_handle* h, int n_stages, const double* A, const double* b, double dt) {
  if (!h || !A || !b) return fail("null argument");
  if (n_stages < 1 || n_stages > SSE_ERK_MAX_TERMS)
    return fail("sse_erk_step: 1..%d stages supported", SSE_ERK_MAX_TERMS);
  if (h->cfg.N_halo) return fail("sse_erk_step: element shards are stepped by the host framework");
  for (int s = 0; s < n_stages; ++s)
    for (int j = s; j < n_stages; ++j)
      if (A[s * n_stages + j] != 0.0) return fail("sse_erk_step: the tableau must be explicit");
  CU(cudaSetDevice(h->cfg.device));
  if (h->erk_stages < n_stages) {
    if (h->erk_k) {                    // release the smaller stage buffer of an earlier scheme
      CU(cudaStreamSynchronize(h->stream));
      h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), (void*)h->erk_k),
                      h->allocs.end());
      h->bytes -= (int64_t)h->erk_stages * h->n_state * (int64_t)sizeof(double);
      cudaFree(h->erk_k);
      h->erk_k = nullptr;
      h->erk_stages = 0;
    }
    if (dev_upload<double>(h, nullptr, (size_t)n_stages * h->n_state, &h->erk_k)) return -1;
    if (!h->erk_u && dev_upload<double>(h, nullptr, (size_t)h->n_state, &h->erk_u)) return -1;
    h->erk_stages = n_stages;
  }
  const long long n = h->n_state;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148LL * 16);
  RK none{};
  for (int s = 0; s <= n_stages; ++s) {
    // s < n_stages: stage state into erk_u;  s == n_stages: the update of u itself
    LinComb L{};
    const double* coef = (s < n_stages) ? A + (size_t)s * n_stages : b;
    const int terms = (s < n_stages) ? s : n_stages;
    for (int j = 0; j < terms; ++j)
      if (coef[j] != 0.0) {
        L.c[L.n] = dt * coef[j];
        L.x[L.n] = h->erk_k + (size_t)j * n;
        ++L.n;
      }
    const double* u_stage = h->u;
    if (s == n_stages || L.n > 0) {
      double* out = (s < n_stages) ? h->erk_u : h->u;
      k_lincomb SSE_LAUNCH(blocks, 256, 0, h->stream)(out, h->u, L, n);
      h->launches++;
      CU(cudaGetLastError());
      u_stage = out;
    }
    if (s == n_stages) break;
    if (run_a(h, u_stage)) return -1;
    if (run_b(h, h->erk_k + (size_t)s * n, none)) return -1;
  }
  return 0;
}

int sse_halo_setup(sse_handle* h, const int64_t* send_idx, int64_t n_send) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  const int Nf = h->cfg.N_f, Nc = h->cfg.N_c;
  std::vector<int> off((size_t)n_send);
  for (int64_t s = 0; s < n_send; ++s) {
    int64_t g = send_idx[s];
    if (g < 0 || g >= (int64_t)Nf * h->cfg.N_e) return fail("send index out of range");
    off[s] = (int)((g / Nf) * Nc * Nf + g % Nf);
  }
  if (dev_upload(h, off.data(), off.size(), &h->send_off)) return -1;
  if (dev_upload<double>(h, nullptr, (size_t)n_send * Nc * (h->second_order ? h->cfg.dim : 1),
                         &h->send_buf))
    return -1;
  h->n_send = n_send;
  return 0;
}

int sse_halo_buffers(sse_handle* h, double** send_dev, double** recv_dev, int64_t* n_send,
                     int64_t* n_recv) {
  if (!h) return fail("null handle");
  if (send_dev) *send_dev = h->send_buf;
  if (recv_dev) *recv_dev = h->recv_buf;
  if (n_send) *n_send = h->n_send;
  if (n_recv) *n_recv = h->cfg.N_halo;
  return 0;
}

int sse_halo_pack(sse_handle* h) {
  if (!h) return fail("null handle");
  if (h->n_send == 0) return 0;
  CU(cudaSetDevice(h->cfg.device));
  int n = (int)h->n_send;
  k_halo_pack SSE_LAUNCH((n + 255) / 256, 256, 0, h->stream)(h->u_f, h->send_off, n, h->cfg.N_c,
                                                      h->cfg.N_f, h->send_buf);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int sse_halo_unpack(sse_handle* h) {
  if (!h) return fail("null handle");
  if (h->cfg.N_halo == 0) return 0;
  CU(cudaSetDevice(h->cfg.device));
  int n = (int)h->cfg.N_halo;
  k_halo_unpack SSE_LAUNCH((n + 255) / 256, 256, 0, h->stream)(h->u_f, h->recv_buf, n, h->cfg.N_c,
                                                        h->cfg.N_f, h->cfg.N_e);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

static int time_derivative_range(sse_handle* h, double* dudt_dev, int64_t k_begin, int64_t k_end,
                                 int stages) {
  if (!h) return fail("null handle");
  if (k_begin < 0 || k_end > h->cfg.N_e || k_begin > k_end) return fail("bad element range");
  if (k_begin == k_end) return 0;
  CU(cudaSetDevice(h->cfg.device));
  RK rk = h->use_rk_override ? h->rk_override : RK{};
  h->G.k_begin = k_begin;
  h->G.N_e = k_end;
  h->b_stages = stages;
  int rc = run_b(h, dudt_dev ? dudt_dev : h->dudt, rk);
  h->b_stages = 3;
  h->G.k_begin = 0;
  h->G.N_e = h->cfg.N_e;
  return rc;
}

int sse_time_derivative_range(sse_handle* h, double* dudt_dev, int64_t k_begin, int64_t k_end) {
  // Second order (BR1): time_derivative! reads the neighbours' auxiliary traces q_f, so running
  // auxiliary_variable! on a sub-range only would consume stale q_f of the elements outside it.
  if (h && h->second_order && !(k_begin == 0 && k_end == h->cfg.N_e))
    return fail("sse_time_derivative_range: a second-order (BR1) equation needs the whole element "
                "range here; for sub-ranges call sse_auxiliary_variable_range on all elements, "
                "then sse_time_derivative_only_range");
  return time_derivative_range(h, dudt_dev, k_begin, k_end, 3);
}

int sse_auxiliary_variable_range(sse_handle* h, int64_t k_begin, int64_t k_end) {
  if (!h) return fail("null handle");
  if (!h->second_order) return 0;   // first-order equations have no auxiliary variable
  return time_derivative_range(h, nullptr, k_begin, k_end, 1);
}

int sse_time_derivative_only_range(sse_handle* h, double* dudt_dev, int64_t k_begin,
                                   int64_t k_end) {
  return time_derivative_range(h, dudt_dev, k_begin, k_end, 2);
}

int sse_halo_pack_aux(sse_handle* h) {
  if (!h) return fail("null handle");
  if (!h->second_order) return fail("sse_halo_pack_aux: not a second-order equation");
  if (h->n_send == 0) return 0;
  CU(cudaSetDevice(h->cfg.device));
  int n = (int)h->n_send;
  k_halo_pack_aux SSE_LAUNCH((n + 255) / 256, 256, 0, h->stream)(h->q_f, h->send_off, n, h->cfg.N_c,
                                                          h->cfg.dim, h->cfg.N_f, h->send_buf);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int sse_halo_unpack_aux(sse_handle* h) {
  if (!h) return fail("null handle");
  if (!h->second_order) return fail("sse_halo_unpack_aux: not a second-order equation");
  if (h->cfg.N_halo == 0) return 0;
  CU(cudaSetDevice(h->cfg.device));
  int n = (int)h->cfg.N_halo;
  k_halo_unpack_aux SSE_LAUNCH((n + 255) / 256, 256, 0, h->stream)(h->q_f, h->recv_buf, n, h->cfg.N_c,
                                                            h->cfg.dim, h->cfg.N_f, h->cfg.N_e);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int sse_set_stream(sse_handle* h, void* stream) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  if (h->own_stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)stream;
  h->own_stream = false;
  return 0;
}

int sse_upload_state(sse_handle* h, const double* u_host) {
  if (!h || !u_host) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(h->u, u_host, h->n_state * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int sse_download_dudt(sse_handle* h, double* dudt_host) {
  if (!h || !dudt_host) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(dudt_host, h->dudt, h->n_state * sizeof(double), cudaMemcpyDeviceToHost,
                     h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int sse_sync(sse_handle* h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int sse_functional(sse_handle* h, int which, int arg, const double* exact_q_host, double* out) {
  if (!h || !out) return fail("null argument");
  if (which < SSE_FN_CONSERVATION || which > SSE_FN_L2_ERROR) return fail("unknown functional %d", which);
  CU(cudaSetDevice(h->cfg.device));
  const int Nc = h->cfg.N_c, Nq = h->cfg.N_q;
  const int64_t Ne = h->cfg.N_e;
  const int n_out = (which == SSE_FN_ENTROPY || which == SSE_FN_ENTROPY_RESIDUAL) ? 1 : Nc;
  const double* xa = h->u;
  const double* xb = h->dudt;
  if (which == SSE_FN_CONSERVATION && arg == SSE_ARG_DUDT) xa = h->dudt;
  double* exact_dev = nullptr;
  if (which == SSE_FN_L2_ERROR) {
    if (!exact_q_host) return fail("the L2 error needs the exact solution at the volume nodes");
    const size_t n = (size_t)Nq * Nc * Ne;
    CU(cudaMalloc(&exact_dev, n * sizeof(double)));
    if (cudaMemcpyAsync(exact_dev, exact_q_host, n * sizeof(double), cudaMemcpyHostToDevice,
                        h->stream) != cudaSuccess) {
      cudaFree(exact_dev);
      return fail("functional: upload of the exact solution failed");
    }
    xb = exact_dev;
  }
  const int nblk = (int)std::min<int64_t>(256, Ne);
  const int64_t chunk = (Ne + nblk - 1) / nblk;
  double *partial = nullptr, *blocks = nullptr;
  int rc = 0;
  if (cudaMalloc(&partial, (size_t)Ne * n_out * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&blocks, (size_t)nblk * n_out * sizeof(double)) != cudaSuccess)
    rc = fail("out of device memory for the functional partial sums");
  std::vector<double> hb((size_t)nblk * n_out);
  if (!rc) rc = run_functional(h, which, xa, xb, partial, n_out);
  if (!rc) {
    k_reduce_partials SSE_LAUNCH(nblk, 256, 0, h->stream)(partial, Ne, n_out, chunk, blocks);
    h->launches++;
    if (cudaMemcpyAsync(hb.data(), blocks, hb.size() * sizeof(double), cudaMemcpyDeviceToHost,
                        h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess)
      rc = fail("functional: %s", cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(partial); cudaFree(blocks); cudaFree(exact_dev);
  if (rc) return rc;
  for (int c = 0; c < n_out; ++c) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += hb[(size_t)b * n_out + c];
    out[c] = (which == SSE_FN_L2_ERROR) ? std::sqrt(s) : s;
  }
  return 0;
}

// ---- on-device geometric factors (geometry.cuh)
int sse_geometry_build(const sse_mapping* m, sse_geometry* out) {
  if (!m || !out) return fail("null argument");
  std::memset(out, 0, sizeof(*out));
  const int d = m->dim, Nm = m->N_map, Nq = m->N_q, Nf = m->N_f;
  const int64_t Ne = m->N_e;
  if (d < 1 || d > 3 || Nm < 1 || Nq < 1 || Nf < 1 || Ne < 1) return fail("bad mapping sizes");
  const bool curl = m->metric == SSE_METRIC_CONSERVATIVE_CURL;
  if (curl && d != 3) return fail("the conservative curl form is implemented on the device for d = 3 only");
  const int Nm1 = curl ? (m->P ? m->N_map1 : Nm) : Nm;
  if (!m->Vq || !m->Vf || !m->nrstJ) return fail("mapping operators missing");
  for (int n = 0; n < d; ++n)
    if (!m->D[n] || !m->xyz[n]) return fail("mapping operators missing");
  if (curl && (!m->D1[0] || !m->D1[1] || !m->D1[2] || !m->Vq1 || !m->Vf1))
    return fail("degree-(N+1) operators missing for the curl form");
  CU(cudaSetDevice(m->device));
  std::vector<void*> tmp;
  auto up = [&](const double* src, size_t n, const double** dst) -> int {
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(double)) != cudaSuccess) return fail("out of device memory");
    tmp.push_back(p);
    if (cudaMemcpy(p, src, n * sizeof(double), cudaMemcpyDefault) != cudaSuccess)
      return fail("copy of the mapping data failed");
    *dst = (const double*)p;
    return 0;
  };
  MapOps M{};
  M.Nm = Nm; M.Nm1 = Nm1; M.Nq = Nq; M.Nf = Nf;
  int rc = 0;
  for (int n = 0; n < d && !rc; ++n) {
    rc = up(m->D[n], (size_t)Nm * Nm, &M.D[n]);
    if (!rc) rc = up(m->xyz[n], (size_t)Nm * Ne, &M.xyz[n]);
    if (!rc && curl) rc = up(m->D1[n], (size_t)Nm1 * Nm1, &M.D1[n]);
  }
  if (!rc) rc = up(m->Vq, (size_t)Nq * Nm, &M.Vq);
  if (!rc) rc = up(m->Vf, (size_t)Nf * Nm, &M.Vf);
  if (!rc) rc = up(m->nrstJ, (size_t)Nf * d, &M.nrstJ);
  if (!rc && curl && m->P) rc = up(m->P, (size_t)Nm1 * Nm, &M.P);
  if (!rc && curl) rc = up(m->Vq1, (size_t)Nq * Nm1, &M.Vq1);
  if (!rc && curl) rc = up(m->Vf1, (size_t)Nf * Nm1, &M.Vf1);
  if (!rc && !curl && m->Jproj) rc = up(m->Jproj, (size_t)Nq * Nq, &M.Jproj);
  double *Jq = nullptr, *Lq = nullptr, *Jf = nullptr, *nJ = nullptr;
  if (!rc && (cudaMalloc(&Jq, (size_t)Nq * Ne * sizeof(double)) != cudaSuccess ||
              cudaMalloc(&Lq, (size_t)Nq * d * d * Ne * sizeof(double)) != cudaSuccess ||
              cudaMalloc(&Jf, (size_t)Nf * Ne * sizeof(double)) != cudaSuccess ||
              cudaMalloc(&nJ, (size_t)d * Nf * Ne * sizeof(double)) != cudaSuccess))
    rc = fail("out of device memory for the geometric factors");
  if (!rc) {
    if (curl) {
      const size_t smem = sizeof(double) * ((size_t)13 * Nm + (size_t)24 * Nm1);
      cudaFuncSetAttribute(k_geometry_curl3d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_geometry_curl3d SSE_LAUNCH((unsigned)Ne, 128, smem)(M, Ne, Jq, Lq, Jf, nJ);
    } else {
      const size_t smem = sizeof(double) * ((size_t)(d + d * d) * Nm + Nq);
      if (d == 1) {
        cudaFuncSetAttribute(k_geometry_exact<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_geometry_exact<1> SSE_LAUNCH((unsigned)Ne, 128, smem)(M, Ne, Jq, Lq, Jf, nJ);
      } else if (d == 2) {
        cudaFuncSetAttribute(k_geometry_exact<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_geometry_exact<2> SSE_LAUNCH((unsigned)Ne, 128, smem)(M, Ne, Jq, Lq, Jf, nJ);
      } else {
        cudaFuncSetAttribute(k_geometry_exact<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_geometry_exact<3> SSE_LAUNCH((unsigned)Ne, 128, smem)(M, Ne, Jq, Lq, Jf, nJ);
      }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail("geometry kernel: %s", cudaGetErrorString(e));
  }
  for (void* p : tmp) cudaFree(p);
  if (rc) {
    cudaFree(Jq); cudaFree(Lq); cudaFree(Jf); cudaFree(nJ);
    return rc;
  }
  out->J_q = Jq; out->Lambda_q = Lq; out->J_f = Jf; out->nJf = nJ;
  return 0;
}

int sse_geometry_free(sse_geometry* g) {
  if (!g) return 0;
  cudaFree((void*)g->J_q); cudaFree((void*)g->Lambda_q);
  cudaFree((void*)g->J_f); cudaFree((void*)g->nJf);
  g->J_q = g->Lambda_q = g->J_f = g->nJf = nullptr;
  return 0;
}

int sse_copy_to_host(void* dst_host, const void* src_dev, int64_t bytes) {
  if (!dst_host || !src_dev || bytes < 0) return fail("bad argument");
  CU(cudaMemcpy(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost));
  return 0;
}

void* sse_stream(sse_handle* h) { return h ? (void*)h->stream : nullptr; }

int sse_time_residual(sse_handle* h, int reps, int split, float* ms) {
  if (!h || !ms || reps < 1) return fail("bad argument");
  CU(cudaSetDevice(h->cfg.device));
  RK rk{};
  ms[0] = ms[1] = ms[2] = 0.f;
  CU(cudaEventRecord(h->ev[0], h->stream));
  for (int r = 0; r < reps; ++r) {
    if (run_a(h, h->u)) return -1;
    if (run_b(h, h->dudt, rk)) return -1;
  }
  CU(cudaEventRecord(h->ev[1], h->stream));
  CU(cudaEventSynchronize(h->ev[1]));
  CU(cudaEventElapsedTime(&ms[0], h->ev[0], h->ev[1]));
  if (split) {
    for (int r = 0; r < reps; ++r) {
      float ta = 0.f, tb = 0.f;
      CU(cudaEventRecord(h->ev[0], h->stream));
      if (run_a(h, h->u)) return -1;
      CU(cudaEventRecord(h->ev[1], h->stream));
      if (run_b(h, h->dudt, rk)) return -1;
      CU(cudaEventRecord(h->ev[2], h->stream));
      CU(cudaEventSynchronize(h->ev[2]));
      CU(cudaEventElapsedTime(&ta, h->ev[0], h->ev[1]));
      CU(cudaEventElapsedTime(&tb, h->ev[1], h->ev[2]));
      ms[1] += ta;
      ms[2] += tb;
    }
  }
  return 0;
}

// FP64 FMA-chain microbenchmark: the measured denominator for the FP64-bound kernels
// (MEASURED_PEAKS.json carries no FP64 figure).
__global__ void k_fp64_peak(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int sse_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail("null argument");
  CU(cudaSetDevice(device));
  int sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  CU(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CU(cudaEventRecord(e0));
    k_fp64_peak SSE_LAUNCH(blocks, threads)(out, iters, 1.0);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return 0;
}

// The device log / exp of the entropy-variable maps (physics.cuh: flog, fexp) on host arrays:
// test hook for their accuracy (tests/test_gpu_elementary.py; CPU: the emulation build).
__global__ void k_probe_elementary(int which, const double* __restrict__ x, double* __restrict__ y,
                                   long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = which == 0 ? flog(x[i]) : fexp(x[i]);
}

int sse_probe_elementary(int device, int which, const double* x, double* y, int64_t n) {
  if (!x || !y || n < 0 || (which != 0 && which != 1)) return fail("bad argument");
  if (n == 0) return 0;
  CU(cudaSetDevice(device));
  double *dx = nullptr, *dy = nullptr;
  CU(cudaMalloc(&dx, (size_t)n * sizeof(double)));
  if (cudaMalloc(&dy, (size_t)n * sizeof(double)) != cudaSuccess) {
    cudaFree(dx);
    return fail("cudaMalloc failed");
  }
  cudaError_t e = cudaMemcpy(dx, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const int blocks = (int)std::min<int64_t>((n + 127) / 128, 4096);
    k_probe_elementary SSE_LAUNCH(blocks, 128)(which, dx, dy, (long long)n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(y, dy, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dy);
  if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  return 0;
}

// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) throughput with the same protocol: the evidence for
// the "tensor cores only if DMMA beats the FP64 CUDA-core path" clause of the north star.  Eight
// independent accumulator tiles per warp; FMA = 2 flops, 8*8*4 FMAs per instruction.
#ifndef SSE_HOST_EMU
__global__ void k_dmma_peak(double* out, int iters, double seed) {
  double a = seed + 1e-3 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = seed + t;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

int sse_measure_dmma_peak(int device, double* tflops) {
  if (!tflops) return fail("null argument");
#ifdef SSE_HOST_EMU
  (void)device;
  return fail("no tensor cores in the host emulation");
#else
  CU(cudaSetDevice(device));
  int sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, threads = 256, iters = 2048;
  double* out = nullptr;
  CU(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CU(cudaEventRecord(e0));
    k_dmma_peak SSE_LAUNCH(blocks, threads)(out, iters, 1e-30);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    // per warp and iteration: 8 mma x (8*8*4) FMAs
    double tf = 2.0 * 8.0 * 256.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return 0;
#endif
}

int64_t sse_kernel_launches(sse_handle* h) { return h ? h->launches : 0; }
int64_t sse_device_bytes(sse_handle* h) { return h ? h->bytes : 0; }

}  // extern "C"

#ifdef SSE_HOST_EMU
// host-emulation test build: one translation unit
#include "tu_shard.cu"
#define SSE_TU_DIM 2
#include "tu_nodal.cu"
#include "tu_fluxdiff.cu"
#undef SSE_TU_DIM
#define SSE_TU_DIM 3
#include "tu_nodal.cu"
#include "tu_fluxdiff.cu"
#undef SSE_TU_DIM
#include "tu_standard.cu"
#endif
