// Element-sharded multi-GPU residual behind the C ABI (sse_shard_*, include/sse_b200.h).
//
// The reference has no distributed path (its only parallelism is Threads.@threads over elements,
// /root/reference/src/Solvers/Solvers.jl:498-518).  The only inter-element read of the residual is
// the exterior trace u_f[CI[mapP[:, k]], :] (flux_differencing_form.jl:312-313,
// standard_form_first_order.jl:33-34; second-order equations also read q_f,
// standard_form_second_order.jl:63-64), so rank r of W owns the contiguous element range
// [N_e r / W, N_e (r+1) / W) and exchanges facet traces with the ranks its boundary elements touch:
//
//   loop A (all local elements) -> pack boundary traces -> grouped ncclSend/ncclRecv on the
//   communication stream (NVLink) || loop B on the interior elements -> unpack the halo -> loop B on
//   the boundary elements.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy
// torch already loaded, elsewhere the system library), so the single-GPU library has no link-time
// dependency on it.  The partition is pure host arithmetic (sse_shard_plan_build): halo slots are
// numbered peer by peer and, within a peer, by the owner's global trace index; the owner packs its
// send list in the same order, so no index lists ever travel.
#include <dlfcn.h>

#include <numeric>

#include "handle.h"

namespace {

struct NcclId { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat64 = 8;

int load_nccl(NcclApi** out) {
  static NcclApi api;
  static int state = 0;   // 0 untried, 1 ok, -1 failed
  if (state == 0) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    state = -1;
    if (api.lib) {
      api.GetUniqueId = (int (*)(NcclId*))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
      api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
      api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
      api.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclSend");
      api.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclRecv");
      api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd &&
          api.Send && api.Recv && api.GetErrorString)
        state = 1;
    }
  }
  if (state != 1) return fail("NCCL not available (dlopen of libnccl.so.2 failed): %s",
                              dlerror() ? dlerror() : "symbols missing");
  *out = &api;
  return 0;
}

#define NC(api, call)                                                                  \
  do {                                                                                 \
    int r_ = (call);                                                                   \
    if (r_ != 0) return fail("%s failed: %s", #call, (api)->GetErrorString(r_));       \
  } while (0)

}  // namespace

struct sse_shard {
  sse_handle* h = nullptr;
  sse_shard_plan plan{};
  int rank = 0, world = 1;
  NcclApi* nccl = nullptr;
  void* comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_pack = nullptr, ev_xchg = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  // boundary element ranges of the device-resident flow on their own streams (SSE_B200_SHARD_STREAMS)
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaEvent_t ev_unpack = nullptr, ev_join[2] = {nullptr, nullptr};
  int stream_mode = 1;
  int width = 1;   // doubles per trace node of the widest exchange
  // host-buffer flow: interior pieces [cut[i], cut[i+1]) and, per piece, one past the highest
  // INTERIOR element one of its facet nodes reads a trace from
  std::vector<int64_t> cut, need_hi;
};

extern "C" {

int sse_shard_range(int64_t N_e_global, int rank, int world, int64_t* start, int64_t* stop) {
  if (world < 1 || rank < 0 || rank >= world || N_e_global < 1) return fail("bad rank / world");
  if (start) *start = (N_e_global * rank) / world;
  if (stop) *stop = (N_e_global * (rank + 1)) / world;
  return 0;
}

int sse_shard_plan_build(const int64_t* mapP_cols, int32_t N_f, int64_t N_e_global, int rank,
                         int world, sse_shard_plan* plan, int64_t* mapP_local, int64_t* send_idx) {
  if (!mapP_cols || !plan || !mapP_local || !send_idx) return fail("null argument");
  if (world > SSE_MAX_PEERS) return fail("at most %d ranks", SSE_MAX_PEERS);
  int64_t start, stop;
  if (sse_shard_range(N_e_global, rank, world, &start, &stop)) return -1;
  const int64_t n_loc = stop - start, n = n_loc * N_f;
  std::memset(plan, 0, sizeof(*plan));
  plan->start = start;
  plan->stop = stop;
  std::vector<int64_t> bounds(world + 1);
  for (int r = 0; r <= world; ++r) bounds[r] = (N_e_global * r) / world;
  auto owner_of = [&](int64_t k) {
    int r = (int)std::min<int64_t>(world - 1, (k * world) / N_e_global);
    while (r + 1 < world && bounds[r + 1] <= k) ++r;
    while (r > 0 && bounds[r] > k) --r;
    return r;
  };
  // entries (local linear index g = j + N_f k_loc) whose partner lives on another rank, per peer
  std::vector<std::vector<int64_t>> ext(world);
  std::vector<char> touches(n_loc, 0);
  for (int64_t g = 0; g < n; ++g) {
    const int64_t t = mapP_cols[g];
    if (t < 0 || t >= N_f * N_e_global) return fail("mapP entry %lld out of range", (long long)g);
    const int own = owner_of(t / N_f);
    if (own == rank) {
      mapP_local[g] = t - (int64_t)N_f * start;
    } else {
      ext[own].push_back(g);
      touches[g / N_f] = 1;
    }
  }
  int64_t halo = 0, nsend = 0;
  for (int peer = 0; peer < world; ++peer) {
    std::vector<int64_t>& e = ext[peer];
    if (e.empty()) continue;
    const int64_t cnt = (int64_t)e.size();
    // receive: halo slots ordered by the owner's global trace index (stable)
    std::vector<int64_t> ord(cnt);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(),
                     [&](int64_t a, int64_t b) { return mapP_cols[e[a]] < mapP_cols[e[b]]; });
    for (int64_t s = 0; s < cnt; ++s) mapP_local[e[ord[s]]] = (int64_t)N_f * n_loc + halo + s;
    // send: my nodes whose partner lives on `peer`, ordered by my own global index, i.e. by g
    // (e is already ascending in g)
    for (int64_t s = 0; s < cnt; ++s) send_idx[nsend + s] = e[s];
    const int q = plan->n_peers++;
    plan->peers[q] = peer;
    plan->send_counts[q] = cnt;
    plan->recv_counts[q] = cnt;
    halo += cnt;
    nsend += cnt;
  }
  plan->n_halo = halo;
  plan->n_send = nsend;
  // largest contiguous run of elements that read no halo value
  int64_t best_lo = 0, best_hi = 0, lo = -1;
  for (int64_t k = 0; k <= n_loc; ++k) {
    const bool in = k < n_loc && !touches[k];
    if (in && lo < 0) lo = k;
    if (!in && lo >= 0) {
      if (k - lo > best_hi - best_lo) { best_lo = lo; best_hi = k; }
      lo = -1;
    }
  }
  plan->k_lo = best_lo;
  plan->k_hi = best_hi;
  return 0;
}

int sse_nccl_unique_id(void* id128) {
  if (!id128) return fail("null argument");
  NcclApi* api;
  if (load_nccl(&api)) return -1;
  NcclId id;
  NC(api, api->GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return 0;
}

int sse_shard_destroy(sse_shard* s) {
  if (!s) return 0;
  if (s->h) cudaSetDevice(s->h->cfg.device);
  if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
  if (s->comm && s->nccl) s->nccl->CommDestroy(s->comm);
  for (cudaStream_t a : s->aux)
    if (a) cudaStreamSynchronize(a);
  for (cudaEvent_t e : {s->ev_pack, s->ev_xchg, s->ev_t0, s->ev_t1, s->ev_unpack, s->ev_join[0], s->ev_join[1]})
    if (e) cudaEventDestroy(e);
  for (cudaStream_t a : s->aux)
    if (a) cudaStreamDestroy(a);
  if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
  if (s->h) sse_destroy(s->h);
  delete s;
  return 0;
}

int sse_shard_create(const sse_config* cfg, const sse_operators* ops, const sse_geometry* geo_local,
                     const int64_t* mapP_cols, int rank, int world, const void* nccl_id128,
                     sse_shard** out) {
  if (!cfg || !ops || !geo_local || !mapP_cols || !out) return fail("null argument");
  *out = nullptr;
  if (world > 1 && !nccl_id128) return fail("world > 1 needs the NCCL unique id of rank 0");
  sse_shard* s = new sse_shard();
  s->rank = rank;
  s->world = world;
  int64_t start, stop;
  if (sse_shard_range(cfg->N_e, rank, world, &start, &stop)) { delete s; return -1; }
  const int64_t n_loc = stop - start;
  std::vector<int64_t> mp((size_t)cfg->N_f * n_loc), send((size_t)cfg->N_f * n_loc);
  if (sse_shard_plan_build(mapP_cols, cfg->N_f, cfg->N_e, rank, world, &s->plan, mp.data(),
                           send.data())) { delete s; return -1; }
  sse_config local = *cfg;
  local.N_e = n_loc;
  local.N_halo = s->plan.n_halo;
  if (sse_create(&local, ops, geo_local, mp.data(), &s->h)) { delete s; return -1; }
  auto bail = [&](void) { std::string keep = sse_last_error(); sse_shard_destroy(s);
                          sse_fail("%s", keep.c_str()); return -1; };
  if (s->plan.n_send && sse_halo_setup(s->h, send.data(), s->plan.n_send)) return bail();
  s->width = cfg->N_c * (s->h->second_order ? cfg->dim : 1);
  {
    const int64_t k_lo = s->plan.k_lo, k_hi = s->plan.k_hi;
    const int64_t want = 12, min_piece = 2048;
    const int64_t np = k_hi > k_lo ? std::max<int64_t>(1, std::min(want, (k_hi - k_lo) / min_piece)) : 0;
    for (int64_t q = 0; q <= np && np > 0; ++q) s->cut.push_back(k_lo + ((k_hi - k_lo) * q) / np);
    for (int64_t i = 0; i + 1 < (int64_t)s->cut.size(); ++i) {
      int64_t hi = s->cut[i + 1];
      for (int64_t k = s->cut[i]; k < s->cut[i + 1]; ++k)
        for (int j = 0; j < cfg->N_f; ++j) {
          const int64_t kk = mp[(size_t)k * cfg->N_f + j] / cfg->N_f;
          if (kk >= k_lo && kk < k_hi) hi = std::max(hi, kk + 1);
        }
      s->need_hi.push_back(hi);
    }
  }
  if (cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_pack, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_xchg, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&s->ev_t0) != cudaSuccess || cudaEventCreate(&s->ev_t1) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s->aux[0], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s->aux[1], cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_unpack, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_join[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_join[1], cudaEventDisableTiming) != cudaSuccess) {
    sse_fail("stream / event creation failed");
    return bail();
  }
  if (const char* e = getenv("SSE_B200_SHARD_STREAMS")) s->stream_mode = atoi(e);
  if (world > 1) {
    if (load_nccl(&s->nccl)) return bail();
    NcclId id;
    std::memcpy(&id, nccl_id128, sizeof(id));
    int r = s->nccl->CommInitRank(&s->comm, world, id, rank);
    if (r != 0) { sse_fail("ncclCommInitRank failed: %s", s->nccl->GetErrorString(r)); return bail(); }
  }
  *out = s;
  return 0;
}

sse_handle* sse_shard_handle(sse_shard* s) { return s ? s->h : nullptr; }

int sse_shard_get_plan(sse_shard* s, sse_shard_plan* plan) {
  if (!s || !plan) return fail("null argument");
  *plan = s->plan;
  return 0;
}

}  // extern "C"

// one halo exchange of the packed send buffer, `width` doubles per trace node: enqueued on the
// communication stream behind everything queued on the main stream so far
static int exchange_start(sse_shard* s, int width) {
  sse_handle* h = s->h;
  if (s->world == 1 || s->plan.n_peers == 0) return 0;
  CU(cudaEventRecord(s->ev_pack, h->stream));
  CU(cudaStreamWaitEvent(s->comm_stream, s->ev_pack, 0));
  NcclApi* api = s->nccl;
  NC(api, api->GroupStart());
  int64_t so = 0, ro = 0;
  for (int q = 0; q < s->plan.n_peers; ++q) {
    const int64_t ns = s->plan.send_counts[q] * width, nr = s->plan.recv_counts[q] * width;
    NC(api, api->Recv(h->recv_buf + ro, (size_t)nr, kNcclFloat64, s->plan.peers[q], s->comm,
                      s->comm_stream));
    NC(api, api->Send(h->send_buf + so, (size_t)ns, kNcclFloat64, s->plan.peers[q], s->comm,
                      s->comm_stream));
    so += ns;
    ro += nr;
  }
  NC(api, api->GroupEnd());
  CU(cudaEventRecord(s->ev_xchg, s->comm_stream));
  return 0;
}
// the main stream waits for the exchange started last
static int exchange_wait(sse_shard* s) {
  if (s->world == 1 || s->plan.n_peers == 0) return 0;
  CU(cudaStreamWaitEvent(s->h->stream, s->ev_xchg, 0));
  return 0;
}

// everything after loop A (device-resident data); dudt_host != nullptr: ranges are copied back as
// they finish (sse_download_dudt_range), the interior cut into a few pieces for that
static int shard_flow(sse_shard* s, double* dudt_dev, double* dudt_host) {
  sse_handle* h = s->h;
  const int64_t N = h->cfg.N_e, k_lo = s->plan.k_lo, k_hi = s->plan.k_hi;
  struct Range { int64_t a, b; };
  std::vector<Range> interior, boundary;
  if (k_lo > 0) boundary.push_back({0, k_lo});
  if (std::max(k_hi, k_lo) < N) boundary.push_back({std::max(k_hi, k_lo), N});
  const int pieces = (dudt_host && k_hi - k_lo >= 6 * 4096) ? 6 : 1;
  for (int q = 0; q < pieces; ++q) {
    const int64_t a = k_lo + ((k_hi - k_lo) * q) / pieces, b = k_lo + ((k_hi - k_lo) * (q + 1)) / pieces;
    if (b > a) interior.push_back({a, b});
  }
  auto loop_b = [&](const std::vector<Range>& rs, bool only) -> int {
    for (const Range& r : rs) {
      if (only ? sse_time_derivative_only_range(h, dudt_dev, r.a, r.b)
               : sse_time_derivative_range(h, dudt_dev, r.a, r.b))
        return -1;
      if (dudt_host && sse_download_dudt_range(h, dudt_host, r.a, r.b)) return -1;
    }
    return 0;
  };
  if (s->world == 1) {
    if (h->second_order) return sse_time_derivative(h, dudt_dev) ||
                                (dudt_host ? sse_download_dudt_range(h, dudt_host, 0, N) : 0);
    std::vector<Range> all{{0, N}};
    return loop_b(all, false);
  }
  if (sse_halo_pack(h) || exchange_start(s, h->cfg.N_c)) return -1;
  if (!h->second_order && !dudt_host && s->stream_mode > 0 && s->plan.n_peers > 0 && !boundary.empty()) {
    // Device-resident flow with the boundary ranges on their own streams.  A boundary range at
    // N = 8 is a handful of waves of each kernel (Tet p=4, 511 104 elements: 11 616 elements =
    // 19.6 waves of the flux kernel, 3.9 of the projection), so launched one after the other on
    // the main stream every one of the six loop-B kernels ends in a partly filled wave.  Mode 1
    // (default): the two ranges run side by side once the interior is queued and the halo is
    // unpacked.  Mode 2: their chain -- unpack, flux, projection -- is ordered behind the
    // EXCHANGE only, not behind the interior kernels, and fills the interior's tail waves.
    // The interior reads no halo slot and the ranges write disjoint elements (dudt or the fused
    // RK update), so any interleaving is safe; the main stream joins both at the end.
    // Measured on 2 B200 (profiles/r2_ab_log.md, sessions AC / AD): Euler Tet p=4, 65 856
    // elements per rank (the N = 8 shard size): 2.239 -> 2.216 (mode 1) / 2.208 ms (mode 2);
    // config 3, 65 856 elements per rank: 0.475 -> 0.437 ms (mode 1).  One projection launch
    // over the whole shard after the flux kernels of all ranges (mode "3") was slower (2.25 ms):
    // the short projection launches overlap with the other ranges' flux kernels, removed.
    cudaStream_t const main_stream = h->stream;
    struct Restore { sse_handle* h; cudaStream_t m; ~Restore() { h->stream = m; } } restore{h, main_stream};
    int rc = 0;
    if (s->stream_mode >= 2) {
      CU(cudaStreamWaitEvent(s->aux[0], s->ev_xchg, 0));
      h->stream = s->aux[0];
      rc = sse_halo_unpack(h);
      h->stream = main_stream;
      if (rc) return -1;
      CU(cudaEventRecord(s->ev_unpack, s->aux[0]));
      if (loop_b(interior, false)) return -1;            // overlaps the NVLink transfer
    } else {
      if (loop_b(interior, false)) return -1;
      if (exchange_wait(s) || sse_halo_unpack(h)) return -1;
      CU(cudaEventRecord(s->ev_unpack, main_stream));
      CU(cudaStreamWaitEvent(s->aux[0], s->ev_unpack, 0));
    }
    for (size_t q = 0; q < boundary.size() && q < 2; ++q) {
      if (q > 0) CU(cudaStreamWaitEvent(s->aux[q], s->ev_unpack, 0));
      h->stream = s->aux[q];
      rc = sse_time_derivative_range(h, dudt_dev, boundary[q].a, boundary[q].b);
      h->stream = main_stream;
      if (rc) return -1;
      CU(cudaEventRecord(s->ev_join[q], s->aux[q]));
      CU(cudaStreamWaitEvent(main_stream, s->ev_join[q], 0));
    }
    return 0;
  }
  if (!h->second_order) {
    if (loop_b(interior, false)) return -1;            // overlaps the NVLink transfer
    if (exchange_wait(s) || sse_halo_unpack(h)) return -1;
    return loop_b(boundary, false);
  }
  // BR1: u_f before auxiliary_variable!, q_f before time_derivative! (two exchanges)
  if (k_hi > k_lo && sse_auxiliary_variable_range(h, k_lo, k_hi)) return -1;
  if (exchange_wait(s) || sse_halo_unpack(h)) return -1;
  for (const Range& r : boundary)
    if (sse_auxiliary_variable_range(h, r.a, r.b)) return -1;
  if (sse_halo_pack_aux(h) || exchange_start(s, h->cfg.N_c * h->cfg.dim)) return -1;
  if (loop_b(interior, true)) return -1;
  if (exchange_wait(s) || sse_halo_unpack_aux(h)) return -1;
  return loop_b(boundary, true);
}


// Host-buffer residual of a first-order equation with the upload interleaved with BOTH loops: the
// boundary elements go first -- their traces are packed and the halo exchange starts while the
// interior is still being uploaded -- then the interior arrives piece by piece (H2D on the copy
// stream, loop A of the piece behind it), loop B of a piece is queued as soon as loop A has been
// queued for every element it reads a trace from, and results stream back on the second copy
// stream.  Costs about max(H2D, loops A + B) instead of H2D + loop B.  (Measured at N = 2 on
// B200: 15.1 ms against 18.1 ms for upload-everything-first, profiles/r2_multi_gpu.md.)
static int shard_flow_host_interleaved(sse_shard* s, const double* u_host, double* dudt_host) {
  sse_handle* h = s->h;
  const int64_t N = h->cfg.N_e, k_lo = s->plan.k_lo, k_hi = s->plan.k_hi;
  struct Range { int64_t a, b; };
  std::vector<Range> boundary;
  if (k_lo > 0) boundary.push_back({0, k_lo});
  if (std::max(k_hi, k_lo) < N) boundary.push_back({std::max(k_hi, k_lo), N});
  int first = 1;
  for (const Range& r : boundary) {
    if (sse_upload_range_and_nodal_values(h, u_host, r.a, r.b, first)) return -1;
    first = 0;
  }
  if (sse_halo_pack(h) || exchange_start(s, h->cfg.N_c)) return -1;
  const int64_t np = (int64_t)s->cut.size() - 1;
  std::vector<char> done(np > 0 ? np : 0, 0);
  for (int64_t i = 0; i < np; ++i) {
    if (sse_upload_range_and_nodal_values(h, u_host, s->cut[i], s->cut[i + 1], first)) return -1;
    first = 0;
    for (int64_t j = 0; j <= i; ++j)
      if (!done[j] && s->need_hi[j] <= s->cut[i + 1]) {
        if (sse_time_derivative_range(h, nullptr, s->cut[j], s->cut[j + 1]) ||
            sse_download_dudt_range(h, dudt_host, s->cut[j], s->cut[j + 1]))
          return -1;
        done[j] = 1;
      }
  }
  for (int64_t j = 0; j < np; ++j)
    if (!done[j]) return fail("interleaved flow: interior piece %lld left without its neighbours",
                              (long long)j);
  if (exchange_wait(s) || sse_halo_unpack(h)) return -1;
  for (const Range& r : boundary)
    if (sse_time_derivative_range(h, nullptr, r.a, r.b) ||
        sse_download_dudt_range(h, dudt_host, r.a, r.b))
      return -1;
  return 0;
}

extern "C" {

int sse_shard_residual(sse_shard* s, const double* u, double* dudt, double t, int where) {
  (void)t;
  if (!s) return fail("null shard");
  sse_handle* h = s->h;
  CU(cudaSetDevice(h->cfg.device));
  if (where == SSE_DEVICE) {
    if (sse_nodal_values(h, u)) return -1;     // u == NULL: the device-resident state
    return shard_flow(s, dudt, nullptr);       // dudt == NULL: the handle's own buffer
  }
  if (!u || !dudt) return fail("null argument");
  if (s->world == 1 && !h->second_order) return sse_residual(h, u, dudt, t, SSE_HOST);
  if (s->world > 1 && !h->second_order && s->plan.n_peers > 0) {
    if (!h->split_copy_streams && sse_set_copy_streams(h, 1)) return -1;   // D2H on its own stream
    if (shard_flow_host_interleaved(s, u, dudt)) return -1;
  } else {
    // chunked H2D overlapped with loop A, results copied back range by range
    if (sse_upload_and_nodal_values(h, u)) return -1;
    if (shard_flow(s, nullptr, dudt)) return -1;
  }
  if (sse_sync_copies(h)) return -1;
  return sse_sync(h);
}

// low-storage 2N Runge-Kutta on the sharded, device-resident state: the update of a stage is
// fused into the epilogue of every loop-B launch of the flow (interior and boundary ranges)
int sse_shard_rk_stage(sse_shard* s, double a, double b, double dt) {
  if (!s) return fail("null shard");
  sse_handle* h = s->h;
  if (s->world == 1) return sse_rk_stage(h, a, b, dt);
  if (h->second_order) return fail("sse_shard_rk_stage: first-order equations only");
  CU(cudaSetDevice(h->cfg.device));
  // loop A reads u of ALL local elements before any range updates it: enqueue it first, then run
  // the flow with the RK epilogue active
  if (sse_nodal_values(h, nullptr)) return -1;
  h->rk_override = RK{1, a, b, dt, h->rk_k, h->u};
  h->use_rk_override = 1;
  int rc = shard_flow(s, nullptr, nullptr);
  h->use_rk_override = 0;
  return rc;
}

int sse_shard_rk_step_ck54(sse_shard* s, double dt) {
  static const double A[5] = {0.0, -567301805773.0 / 1357537059087.0,
                              -2404267990393.0 / 2016746695238.0,
                              -3550918686646.0 / 2091501179385.0,
                              -1275806237668.0 / 842570457699.0};
  static const double B[5] = {1432997174477.0 / 9575080441755.0,
                              5161836677717.0 / 13612068292357.0,
                              1720146321549.0 / 2090206949498.0,
                              3134564353537.0 / 4481467310338.0,
                              2277821191437.0 / 14882151754819.0};
  for (int st = 0; st < 5; ++st)
    if (sse_shard_rk_stage(s, A[st], B[st], dt)) return -1;
  return 0;
}

int sse_shard_time_residual(sse_shard* s, int reps, float* ms) {
  if (!s || !ms || reps < 1) return fail("bad argument");
  sse_handle* h = s->h;
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaEventRecord(s->ev_t0, h->stream));
  for (int r = 0; r < reps; ++r)
    if (sse_shard_residual(s, nullptr, nullptr, 0.0, SSE_DEVICE)) return -1;
  CU(cudaEventRecord(s->ev_t1, h->stream));
  CU(cudaEventSynchronize(s->ev_t1));
  CU(cudaEventElapsedTime(ms, s->ev_t0, s->ev_t1));
  return 0;
}

int sse_shard_sync(sse_shard* s) {
  if (!s) return fail("null shard");
  CU(cudaSetDevice(s->h->cfg.device));
  CU(cudaStreamSynchronize(s->comm_stream));
  return sse_sync(s->h);
}

}  // extern "C"
