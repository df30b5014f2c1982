// Host emulation of the CUDA execution model -- TEST INFRASTRUCTURE ONLY.
//
// tests/emu/build_emu.py compiles csrc/sse_b200.cu with g++ -DSSE_HOST_EMU against this header
// (instead of <cuda_runtime.h>) into tests/emu/_build/libsse_b200_emu.so, so that the CPU test
// suite can execute the real kernel SOURCES -- index logic, shared-memory carve-ups, barrier
// placement, operator tables -- against the oracle without a GPU.  It is never built by
// __graft_entry__.build(), never shipped and never loaded by the package: device.load_library
// refuses an emulation build (sse_version() < 0) unless a test asks for it explicitly.  It says
// nothing about performance or about sm_100a code generation; GPU parity is tests -m gpu.
//
// Model: one OS thread; every CUDA thread of a block is a ucontext fiber; __syncthreads() yields
// to a round-robin scheduler that resumes the block's fibers in thread order until all of them
// are at the barrier (or have returned).  Blocks run one after another.  No warp-level
// primitives are emulated (the kernels use none); "device memory" is host memory.
#pragma once
#include <dlfcn.h>
#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

// ------------------------------------------------------------------ language extensions
#define __global__
#define __device__
#define __host__
#define __constant__ static
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static   // static __shared__ arrays inside kernels: blocks run one at a time

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace emu {
struct State {
  uint3 tIdx{0, 0, 0}, bIdx{0, 0, 0};
  dim3 bDim, gDim;
  double* smem = nullptr;
  ucontext_t sched;
  ucontext_t* fibers = nullptr;
  int current = -1;
};
inline State& st() {
  static State s;
  return s;
}
inline void barrier() {   // __syncthreads(): back to the scheduler, resumed after all arrive
  State& s = st();
  swapcontext(&s.fibers[s.current], &s.sched);
}
}  // namespace emu

static inline void __syncthreads() { emu::barrier(); }
// __syncwarp(): valid in the emulation where every warp of the block executes the same number of
// them (the warp-private kernels do): a block-wide barrier is then a (stricter) substitute
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier(); }

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }

template <class A, class B>
static inline typename std::common_type<A, B>::type min(A a, B b) {
  using C = typename std::common_type<A, B>::type;
  return (C)a < (C)b ? (C)a : (C)b;
}
template <class A, class B>
static inline typename std::common_type<A, B>::type max(A a, B b) {
  using C = typename std::common_type<A, B>::type;
  return (C)a > (C)b ? (C)a : (C)b;
}
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
// MUFU.RCP64H-like seed of the kernels' Newton reciprocals (rcp.approx.ftz.f64: ~2^-22)
static inline double emu_rcp_approx(double x) { return (double)(float)(1.0 / x); }
// bit-level views of a double (device intrinsics of the same names)
static inline int __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) {
  long long b = ((long long)hi << 32) | (unsigned)lo;
  double x; std::memcpy(&x, &b, 8); return x;
}
static inline double __int2double_rn(int v) { return (double)v; }

// ------------------------------------------------------------------ launches
namespace emu {
constexpr size_t kStack = 256 * 1024;

struct Pool {
  std::vector<char> stacks;
  std::vector<ucontext_t> ctx;
  std::vector<char> done;
  std::vector<double> smem;
};
inline Pool& pool() {
  static Pool p;
  return p;
}
inline int& order_mode() {   // 0: ascending thread index, 1: descending, 2: a scrambled order
  static int m = 0;
  return m;
}
inline std::function<void()>*& body_slot() {
  static std::function<void()>* b = nullptr;
  return b;
}
inline void fiber_main(int tid) {
  (*body_slot())();
  pool().done[tid] = 1;
  State& s = st();
  swapcontext(&s.fibers[tid], &s.sched);   // never resumed
}

// run `body` as a grid of blocks of `nthr` fibers each
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body) {
  Pool& P = pool();
  State& s = st();
  const int nthr = (int)(block.x * block.y * block.z);
  if (P.ctx.size() < (size_t)nthr) {
    P.ctx.resize(nthr);
    P.done.resize(nthr);
    P.stacks.resize((size_t)nthr * kStack);
  }
  // exactly the dynamic shared memory the launch asked for (a fresh heap block, so that an
  // AddressSanitizer build of the emulator traps accesses past its end), NaN-poisoned: reads of
  // never-written shared memory show up as NaN
  std::vector<double>(smem_bytes / sizeof(double), std::nan("")).swap(P.smem);
  s.smem = P.smem.data();
  s.bDim = block;
  s.gDim = grid;
  s.fibers = P.ctx.data();
  body_slot() = &body;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bIdx = uint3{bx, by, bz};
        for (int t = 0; t < nthr; ++t) {
          P.done[t] = 0;
          getcontext(&P.ctx[t]);
          P.ctx[t].uc_stack.ss_sp = P.stacks.data() + (size_t)t * kStack;
          P.ctx[t].uc_stack.ss_size = kStack;
          P.ctx[t].uc_link = &s.sched;
          makecontext(&P.ctx[t], (void (*)())fiber_main, 1, t);
        }
        int remaining = nthr;
        unsigned pass = 0;
        while (remaining > 0) {       // one pass = one barrier interval
          remaining = 0;
          ++pass;
          for (int q = 0; q < nthr; ++q) {
            // thread order inside a barrier interval: results must not depend on it, so running
            // the suite with another order (emu_set_order) is a shared-memory race check
            int t = q;
            if (order_mode() == 1) t = nthr - 1 - q;
            else if (order_mode() == 2) t = (int)(((unsigned long long)q * 61u + 17u * pass) % (unsigned)nthr);
            if (P.done[t]) continue;
            s.current = t;
            s.tIdx = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y,
                                (unsigned)t / (block.x * block.y)};
            swapcontext(&s.sched, &P.ctx[t]);
            if (!P.done[t]) ++remaining;
          }
        }
      }
  body_slot() = nullptr;
}

template <class... Args>
struct Bound {
  dim3 grid, block;
  size_t smem;
  std::tuple<typename std::decay<Args>::type...> args;
};
struct Launcher {
  dim3 grid, block;
  size_t smem;
  Launcher(dim3 g, dim3 b, size_t s = 0, void* = nullptr) : grid(g), block(b), smem(s) {}
  template <class... Args>
  Bound<Args...> operator()(Args&&... a) const {
    return Bound<Args...>{grid, block, smem, std::make_tuple(std::forward<Args>(a)...)};
  }
};
// kernel * Launcher(grid, block, smem, stream)(args...)  ==  kernel<<<grid, block, smem, stream>>>(args...)
inline std::string& launch_log() {
  static std::string s;
  return s;
}
template <class... KArgs, class... Args>
inline void operator*(void (*kernel)(KArgs...), const Bound<Args...>& b) {
  Dl_info info;   // mangled kernel name, so that tests can assert WHICH kernels ran
  if (dladdr((void*)kernel, &info) && info.dli_sname) launch_log() += std::string(info.dli_sname) + "\n";
  launch(b.grid, b.block, b.smem, [&]() {
    std::apply([&](const auto&... a) { kernel(a...); }, b.args);
  });
}
}  // namespace emu
using emu::operator*;
#define threadIdx (emu::st().tIdx)
#define blockIdx (emu::st().bIdx)
#define blockDim (emu::st().bDim)
#define gridDim (emu::st().gDim)

// Warp shuffle, valid where every thread of the BLOCK executes it in uniform control flow (the
// only use: block_sum in functionals.cuh): exchange through a buffer between two barriers.
static inline double __shfl_down_sync(unsigned, double v, int delta) {
  static std::vector<double> buf;
  emu::State& s = emu::st();
  const int nthr = (int)(s.bDim.x * s.bDim.y * s.bDim.z);
  const int tid = s.current, lane = tid & 31;
  if ((int)buf.size() < nthr) buf.resize(nthr);
  buf[tid] = v;
  emu::barrier();
  const double r = (lane + delta < 32 && tid + delta < nthr) ? buf[tid + delta] : v;
  emu::barrier();
  return r;
}

static inline double __shfl_xor_sync(unsigned, double v, int mask) {   // same contract as above
  static std::vector<double> buf;
  emu::State& s = emu::st();
  const int nthr = (int)(s.bDim.x * s.bDim.y * s.bDim.z);
  const int tid = s.current;
  if ((int)buf.size() < nthr) buf.resize(nthr);
  buf[tid] = v;
  emu::barrier();
  const int src = (tid & ~31) | ((tid & 31) ^ mask);
  const double r = src < nthr ? buf[src] : v;
  emu::barrier();
  return r;
}

#define SSE_LAUNCH(...) * emu::Launcher(__VA_ARGS__)
#define SSE_SHARED(name) double* name = emu::st().smem
#define SSE_SHARED16(name) double* name = emu::st().smem

// ------------------------------------------------------------------ runtime API
typedef int cudaError_t;
typedef void* cudaStream_t;
struct emu_event { std::chrono::steady_clock::time_point t; };
typedef emu_event* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost,
                      cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };

static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  *v = (a == cudaDevAttrMultiProcessorCount) ? 4 : 227 * 1024;
  return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) {
  *p = (T*)std::malloc(n ? n : 1);
  if (*p) std::memset((void*)*p, 0xFF, n);   // NaN-poison: catches reads of unwritten buffers
  return *p ? cudaSuccess : cudaErrorUnknown;
}
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind,
                                          cudaStream_t = nullptr) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
  std::memset(d, v, n);
  return cudaSuccess;
}
template <class S> static inline cudaError_t cudaMemcpyToSymbol(S& sym, const void* src, size_t n,
                                                                size_t off = 0) {
  std::memcpy((char*)&sym + off, src, n);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emu_event; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
  e->t = std::chrono::steady_clock::now();
  return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) {
  *n = 3;
  return cudaSuccess;
}

// names (mangled) of the kernels launched since the last call, one per line
extern "C" __attribute__((used)) const char* emu_launch_log() {
  static std::string out;
  out.swap(emu::launch_log());
  emu::launch_log().clear();
  return out.c_str();
}

extern "C" __attribute__((used)) void emu_set_order(int mode) { emu::order_mode() = mode; }
