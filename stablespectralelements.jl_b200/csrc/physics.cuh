// Pointwise physics as __device__ functions (FP64).
//
// Restates /root/reference/src/ConservationLaws/: logmean / inv_logmean
// (ConservationLaws.jl:132-156), Euler maps and fluxes (euler_navierstokes.jl:58-225), linear
// advection (linear_advection_diffusion.jl:53-119) and Burgers (burgers.jl:51-133).
//
// Two-point fluxes are evaluated already contracted with a direction vector c:
//   out[e] = sum_n c[n] * F[e][n](u_L, u_R)
// which is the only way the residual ever uses the flux tensor
// (flux_differencing_form.jl:20-33,108-121; ConservationLaws.jl:92-100).
#pragma once
#include <math.h>

// Kernel-launch and dynamic-shared-memory spellings.  The one place they differ is the host
// emulation build of the test suite (tests/emu/cuda_emu.h defines them before this header is
// seen and runs every CUDA thread as a fiber); under nvcc they are the plain CUDA syntax.
#ifndef SSE_HOST_EMU
#define SSE_LAUNCH(...) <<<__VA_ARGS__>>>
#define SSE_SHARED(name) extern __shared__ double name[]
#define SSE_SHARED16(name) extern __shared__ __align__(16) double name[]
#define SSE_RCP_APPROX(y, x) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x))
#define SSE_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
// asynchronous global -> shared copies (LDGSTS), 8 bytes each: the batched kernels stage the
// inputs of their E elements without a register round trip, every copy in flight at once
#define SSE_CP_ASYNC8(dst, src)                                                     \
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(                    \
                   (unsigned)__cvta_generic_to_shared(dst)),                        \
               "l"(src) : "memory")
#define SSE_CP_ASYNC_WAIT_ALL() asm volatile("cp.async.wait_all;" ::: "memory")
// Bulk asynchronous copies (the TMA engine's 1-D copy, SASS UBLKCP) that complete on a
// shared-memory mbarrier: one instruction moves a whole operator block; source, destination and
// size are multiples of 16 bytes.
#define SSE_SMEM_U32(p) ((unsigned)__cvta_generic_to_shared(p))
#define SSE_MBAR_INIT(bar, n)                                                                 \
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SSE_SMEM_U32(bar)), "r"(n) : "memory")
#define SSE_MBAR_INIT_FENCE() asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory")
#define SSE_MBAR_EXPECT_TX(bar, bytes)                                                        \
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SSE_SMEM_U32(bar)), \
               "r"((unsigned)(bytes)) : "memory")
#define SSE_BULK_G2S(dst, src, bytes, bar)                                                    \
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" \
               ::"r"(SSE_SMEM_U32(dst)), "l"(src), "r"((unsigned)(bytes)), "r"(SSE_SMEM_U32(bar)) : "memory")
#define SSE_SMEM_ALIGNED16(p) ((SSE_SMEM_U32(p) & 15u) == 0u)
// Every thread polls the phase.  A wait that outlives any plausible copy (2 s of wall clock on
// %globaltimer, looked at every 4096 polls -- a bound in time, not in polls, so that preemption,
// time slicing or a debugger cannot trip it) traps instead of hanging the device on a wrong byte
// count.
__device__ __forceinline__ void sse_mbar_wait(const void* bar, unsigned parity) {
  const unsigned a = SSE_SMEM_U32(bar);
  unsigned long long t0 = 0;
  for (unsigned spin = 1;; ++spin) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if ((spin & 4095u) == 0u) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) __trap();
    }
  }
}
#define SSE_MBAR_WAIT(bar, parity) sse_mbar_wait(bar, parity)
#else
// host emulation: the copy is done on the spot by the issuing fiber, the barrier ops are no-ops
// (the kernels place a __syncthreads between the issue and the first use)
#define SSE_MBAR_INIT(bar, n) ((void)(bar))
#define SSE_MBAR_INIT_FENCE() ((void)0)
#define SSE_MBAR_EXPECT_TX(bar, bytes) ((void)(bar))
#define SSE_BULK_G2S(dst, src, bytes, bar) memcpy((void*)(dst), (const void*)(src), (size_t)(bytes))
#define SSE_MBAR_WAIT(bar, parity) ((void)(bar))
#define SSE_SMEM_ALIGNED16(p) ((((size_t)(p)) & 15u) == 0u)
#define SSE_CP_ASYNC8(dst, src) (*(dst) = *(src))
#define SSE_CP_ASYNC_WAIT_ALL() ((void)0)
#define SSE_RCP_APPROX(y, x) y = emu_rcp_approx(x)
#define SSE_PREFETCH_L2(p) ((void)(p))
#endif

namespace sse {

enum { LAW_ADV = 0, LAW_BURGERS = 1, LAW_EULER = 2 };

struct Phys {
  double a[3];
  double b;
  double gamma;
  double inv_gm1;     // 1/(gamma-1), log(gamma-1): set once on the host
  double log_gm1;
  double half_lambda;
  int inviscid;    // sse_inviscid_flux
  int two_point;   // sse_two_point_flux used by the interface flux / volume terms
};

template <int DIM, int LAW> struct LawTraits {
  static constexpr int NC = (LAW == LAW_EULER) ? DIM + 2 : 1;
  // per-node state kept in shared memory: Euler primitives {rho, V[DIM], p, rho/p}; scalar {u}
  static constexpr int NS = (LAW == LAW_EULER) ? DIM + 3 : 1;
};

// a / b = a * frcp(b): ~1 ulp from IEEE division, far inside the 1e-12 parity budget.
// frcp: MUFU.RCP64H seed (rel. error < 2^-22) + two Newton steps; branch-free.  Valid for the
// normal, finite, non-zero arguments that occur here (densities, pressures, Jacobians).
__device__ __forceinline__ double frcp(double x) {
  double y;
  SSE_RCP_APPROX(y, x);
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double fdiv(double a, double b) { return a * frcp(b); }

// reciprocal good to ~2^-44 relative (seed + one Newton step): enough wherever the result only
// enters through f^2 < 1e-4 (see logmean below)
__device__ __forceinline__ double frcp1(double x) {
  double y;
  SSE_RCP_APPROX(y, x);
  double e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// log / exp of the entropy-variable maps (loop A evaluates 475 logs and 225 exps per Tet p=4
// element).  libdevice's versions carry their polynomial coefficients as 64-bit literals -- two
// move instructions each on sm_100, half of the instructions of a call -- and branches for the
// special cases.  These take the coefficients from constant memory and send anything that is not
// a positive normal finite number (log) / not in |x| < 700 (exp) to libdevice out of line, so the
// results for such arguments are libdevice's.
//   flog: fdlibm's e_log.c scheme (x = 2^k m, m in [sqrt(1/2), sqrt(2)), s = f/(2+f), degree-7
//         polynomial in s^2; < 1 ulp), the division through frcp;
//   fexp: x = n ln2 + r, |r| <= ln2/2, exp(r) = 1 + r + r^2 q(r), q of degree 9 (Chebyshev
//         interpolant, tools/fit_elementary.py: 0.66 ulp), 2^n added into the exponent field.
__constant__ double c_el[24] = {
    // [0..6] Lg1..Lg7, [7] pad
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
    2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
    1.479819860511658591e-01, 0.0,
    // [8] ln2_hi, [9] ln2_lo, [10] log2(e), [11] 1.5 * 2^52 (round-to-nearest-integer shift)
    6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.4426950408889634074, 6755399441055744.0,
    // [12..21] q(r) of exp
    5.00000000000000111e-01, 1.66666666666666685e-01, 4.16666666666241289e-02,
    8.33333333333006153e-03, 1.38888889172137167e-03, 1.98412698630536177e-04,
    2.48015212959543761e-05, 2.75572684599970641e-06, 2.76200884454097462e-07,
    2.51003854955103203e-08, 0.0, 0.0};
// one out-of-line copy each: loop A streams through ~34 KB of straight-line code per element and
// is sensitive to its instruction footprint (profiles/r2_ab_log.md)
#ifdef SSE_ELEM_INLINE
#define SSE_ELEM_FN __device__ __forceinline__
#else
#define SSE_ELEM_FN static __device__ __noinline__
#endif
static __device__ __noinline__ double log_libdevice(double x) { return log(x); }
static __device__ __noinline__ double exp_libdevice(double x) { return exp(x); }

SSE_ELEM_FN double flog(double x) {
#ifdef SSE_LIBDEVICE_ELEM   // A/B knob: libdevice's log / exp inline
  return log(x);
#endif
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log_libdevice(x);
  hi += 0x3ff00000 - 0x3fe6a09e;
  const double dk = (double)((hi >> 20) - 0x3ff);
  const double m = __hiloint2double((hi & 0x000fffff) + 0x3fe6a09e, lo);
  const double f = m - 1.0;
  const double s = f * frcp(2.0 + f);
  const double z = s * s, w = z * z;
  const double t1 = w * fma(w, fma(w, c_el[5], c_el[3]), c_el[1]);
  const double t2 = fma(w, fma(w, fma(w, c_el[6], c_el[4]), c_el[2]), c_el[0]);
  const double R = fma(z, t2, t1);
  const double hfsq = 0.5 * f * f;
  return fma(dk, c_el[8], -((hfsq - fma(s, hfsq + R, dk * c_el[9])) - f));
}

SSE_ELEM_FN double fexp(double x) {
#ifdef SSE_LIBDEVICE_ELEM
  return exp(x);
#endif
  if (!(fabs(x) < 700.0)) return exp_libdevice(x);
  const double t = fma(x, c_el[10], c_el[11]);
  const int n = __double2loint(t);
  const double fn = t - c_el[11];
  double r = fma(fn, -c_el[8], x);
  r = fma(fn, -c_el[9], r);
  double q = c_el[21];
#pragma unroll
  for (int i = 20; i >= 12; --i) q = fma(q, r, c_el[i]);
  const double e = fma(r * r, q, r) + 1.0;
  return __hiloint2double(__double2hiint(e) + (n << 20), __double2loint(e));
}

// logmean / inv_logmean (ConservationLaws.jl:132-156).  The reference's
//   f^2 = (x(x-2y)+y^2)/(x(x+2y)+y^2) = ((x-y)/(x+y))^2,
// and on its Taylor branch (f^2 < 1e-4), with z = f^2/3 + f^4/5 + f^6/7,
//   logmean     = (x+y) 105/(210 + f^2(70 + f^2(42 + 30 f^2))) = (x+y)/2 / (1 + z)
//   inv_logmean = 2/(x+y) (1 + z)                                  (a finite polynomial).
// 1/(1+z) is expanded as 1 - f^2/3 - 4/45 f^4 - 44/945 f^6 (+0.081 f^8 < 1e-17, dropped), so
// logmean needs no full-precision division: 1/(x+y) only enters through f^2, whose relative
// error e contributes e*f^2/3 < 4e-5 e to the result -- one Newton step (e ~ 6e-14) is ample.
// inv_logmean needs 1/(x+y) itself and reuses it for f.  Both agree with the reference's
// formula to ~2 ulp; a flipped branch decision at the threshold is harmless (the branches
// agree to ~2e-17 relative there).
// The polynomial coefficients and the threshold live in constant memory: sm_100 FP64
// instructions take no 64-bit immediates, so a literal costs two UMOVs per use where a
// constant-bank value costs one LDCU.
__constant__ double c_lm[8] = {-1.0 / 3.0, -4.0 / 45.0, -44.0 / 945.0,   // 1/(1+z) series
                               1.0 / 3.0,  1.0 / 5.0,   1.0 / 7.0,       // 1 + z
                               1.0e-4, 0.0};

// rare branch (|x-y|/(x+y) >= 1e-2): kept out of line so the common path stays small
static __device__ __noinline__ double logmean_full(double x, double y) { return (y - x) / log(y / x); }

// Taylor-branch values; f2 is returned so that callers can test both means with one branch
__device__ __forceinline__ double logmean_taylor(double x, double y, double& f2) {
  const double d = x - y, s = x + y;
  const double g = d * frcp1(s);
  f2 = g * g;
  const double pl = fma(fma(fma(f2, c_lm[2], c_lm[1]), f2, c_lm[0]), f2, 1.0);
  return (0.5 * s) * pl;
}
__device__ __forceinline__ double inv_logmean_taylor(double x, double y, double& f2) {
  const double d = x - y, s = x + y;
  const double is = frcp(s);
  const double g = d * is;
  f2 = g * g;
  const double pl = fma(fma(fma(f2, c_lm[5], c_lm[4]), f2, c_lm[3]), f2, 1.0);
  return (is + is) * pl;
}

__device__ __forceinline__ double logmean(double x, double y) {
  double f2;
  double r = logmean_taylor(x, y, f2);
  if (!(f2 < c_lm[6])) r = logmean_full(x, y);
  return r;
}

__device__ __forceinline__ double inv_logmean(double x, double y) {
  double f2;
  double r = inv_logmean_taylor(x, y, f2);
  if (!(f2 < c_lm[6])) r = frcp(logmean_full(x, y));
  return r;
}

// both means of the Ranocha flux when at least one of them left the Taylor branch (rare):
// .x = logmean(x0, y0), .y = inv_logmean(x1, y1)
static __device__ __noinline__ double2 logmeans_slow(double x0, double y0, double x1, double y1) {
  return make_double2(logmean(x0, y0), inv_logmean(x1, y1));
}

// conservative -> shared-memory state
template <int DIM, int LAW>
__device__ __forceinline__ void cons_to_state(const Phys& P, const double* u, double* s) {
  if constexpr (LAW == LAW_EULER) {
    double rho = u[0];
    double irho = frcp(rho);
    double k = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      s[1 + m] = u[1 + m] * irho;
      k += s[1 + m] * s[1 + m];
    }
    double p = (P.gamma - 1.0) * (u[DIM + 1] - 0.5 * rho * k);
    s[0] = rho;
    s[DIM + 1] = p;
    s[DIM + 2] = fdiv(rho, p);
  } else {
    s[0] = u[0];
  }
}

template <int DIM, int LAW>
__device__ __forceinline__ void state_to_cons(const Phys& P, const double* s, double* u) {
  if constexpr (LAW == LAW_EULER) {
    double k = 0.0;
    u[0] = s[0];
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      u[1 + m] = s[0] * s[1 + m];
      k += s[1 + m] * s[1 + m];
    }
    u[DIM + 1] = s[DIM + 1] * P.inv_gm1 + 0.5 * s[0] * k;
  } else {
    u[0] = s[0];
  }
}

// physical flux contracted with c, from the state
template <int DIM, int LAW>
__device__ __forceinline__ void physical_flux_c(const Phys& P, const double* s, const double* c,
                                                double* out) {
  if constexpr (LAW == LAW_EULER) {
    double vc = 0.0, k = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      vc += s[1 + m] * c[m];
      k += s[1 + m] * s[1 + m];
    }
    double rho = s[0], p = s[DIM + 1];
    double E = p * P.inv_gm1 + 0.5 * rho * k;
    out[0] = rho * vc;
#pragma unroll
    for (int m = 0; m < DIM; ++m) out[1 + m] = rho * s[1 + m] * vc + p * c[m];
    out[DIM + 1] = (E + p) * vc;
  } else {
    double ac = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) ac += P.a[m] * c[m];
    if constexpr (LAW == LAW_ADV) out[0] = ac * s[0];
    else out[0] = 0.5 * ac * s[0] * s[0];
  }
}

// two-point flux contracted with c
template <int DIM, int LAW>
__device__ __forceinline__ void two_point_flux_c(const Phys& P, int kind, const double* L,
                                                 const double* R, const double* c, double* out) {
  if constexpr (LAW == LAW_EULER) {
    if (kind == 1) {  // entropy-conservative (Ranocha), euler_navierstokes.jl:171-195
      double f2a, f2b;
      double rho_avg = logmean_taylor(L[0], R[0], f2a);
      double ilm = inv_logmean_taylor(L[DIM + 2], R[DIM + 2], f2b);
      if (!((f2a < c_lm[6]) && (f2b < c_lm[6]))) {   // one (rare) branch for both means
        const double2 m = logmeans_slow(L[0], R[0], L[DIM + 2], R[DIM + 2]);
        rho_avg = m.x;
        ilm = m.y;
      }
      double p_avg = 0.5 * (L[DIM + 1] + R[DIM + 1]);
      double vlvr = 0.0, vc = 0.0, vlc = 0.0, vrc = 0.0;
      double vavg[DIM];
#pragma unroll
      for (int m = 0; m < DIM; ++m) {
        vavg[m] = 0.5 * (L[1 + m] + R[1 + m]);
        vlvr += L[1 + m] * R[1 + m];
        vc += vavg[m] * c[m];
        vlc += L[1 + m] * c[m];
        vrc += R[1 + m] * c[m];
      }
      double C = fma(ilm, P.inv_gm1, 0.5 * vlvr);
      double f_rho = rho_avg * vc;
      out[0] = f_rho;
#pragma unroll
      for (int m = 0; m < DIM; ++m) out[1 + m] = f_rho * vavg[m] + p_avg * c[m];
      out[DIM + 1] = f_rho * C + 0.5 * (L[DIM + 1] * vrc + R[DIM + 1] * vlc);
    } else {  // arithmetic mean of physical fluxes, euler_navierstokes.jl:152-158
      double fl[DIM + 2], fr[DIM + 2];
      physical_flux_c<DIM, LAW>(P, L, c, fl);
      physical_flux_c<DIM, LAW>(P, R, c, fr);
#pragma unroll
      for (int e = 0; e < DIM + 2; ++e) out[e] = 0.5 * (fl[e] + fr[e]);
    }
  } else {
    double ac = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) ac += P.a[m] * c[m];
    if constexpr (LAW == LAW_ADV) {
      out[0] = ac * (0.5 * (L[0] + R[0]));
    } else {
      if (kind == 1) out[0] = ac * ((L[0] * L[0] + L[0] * R[0] + R[0] * R[0]) / 6.0);
      else out[0] = ac * ((L[0] * L[0] + R[0] * R[0]) * 0.25);
    }
  }
}

// ---- Ranocha's entropy-conservative flux in "half-velocity" form (specialised loop-B kernels).
// Node states are kept as {rho, V/2 (DIM), p, rho/p}: scaling by 1/2 is exact in binary floating
// point, so the averages  ½(V_L+V_R) = h_L + h_R,  ½ V_L.V_R = 2 h_L.h_R  and
// ½(p_L V_R + p_R V_L).c = p_L (h_R.c) + p_R (h_L.c)  lose their explicit multiplications by ½
// (6 of 64 FP64 instructions per flux) without changing a single rounding of the products.
template <int DIM>
__device__ __forceinline__ void euler_state_to_half(double* s) {
#pragma unroll
  for (int m = 0; m < DIM; ++m) s[1 + m] *= 0.5;
}

template <int DIM>
__device__ __forceinline__ void ec_flux_half_c(const Phys& P, const double* L, const double* R,
                                               const double* c, double* out) {
  double f2a, f2b;
  double rho_avg = logmean_taylor(L[0], R[0], f2a);
  double ilm = inv_logmean_taylor(L[DIM + 2], R[DIM + 2], f2b);
  if (!((f2a < c_lm[6]) && (f2b < c_lm[6]))) {   // one (rare) branch for both means
    const double2 m = logmeans_slow(L[0], R[0], L[DIM + 2], R[DIM + 2]);
    rho_avg = m.x;
    ilm = m.y;
  }
  double hh = 0.0, hlc = 0.0, hrc = 0.0;
  double vavg[DIM];
#pragma unroll
  for (int m = 0; m < DIM; ++m) {
    vavg[m] = L[1 + m] + R[1 + m];
    hh = fma(L[1 + m], R[1 + m], hh);
    hlc = fma(L[1 + m], c[m], hlc);
    hrc = fma(R[1 + m], c[m], hrc);
  }
  const double vc = hlc + hrc;                       // ½ (V_L + V_R) . c
  const double C = fma(ilm, P.inv_gm1, hh + hh);     // 1/((gamma-1) beta_lm) + ½ V_L.V_R
  const double p_avg = 0.5 * (L[DIM + 1] + R[DIM + 1]);
  const double f_rho = rho_avg * vc;
  out[0] = f_rho;
#pragma unroll
  for (int m = 0; m < DIM; ++m) out[1 + m] = fma(f_rho, vavg[m], p_avg * c[m]);
  out[DIM + 1] = fma(f_rho, C, fma(L[DIM + 1], hrc, R[DIM + 1] * hlc));
}

// wave speed for the Lax-Friedrichs flux (unit normal n)
template <int DIM, int LAW>
__device__ __forceinline__ double wave_speed(const Phys& P, const double* L, const double* R,
                                             const double* n) {
  if constexpr (LAW == LAW_EULER) {
    double vnl = 0.0, vnr = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) {
      vnl += L[1 + m] * n[m];
      vnr += R[1 + m] * n[m];
    }
    // c^2 = gamma p / rho = gamma / (rho/p)
    double cl = sqrt(fdiv(P.gamma, L[DIM + 2]));
    double cr = sqrt(fdiv(P.gamma, R[DIM + 2]));
    return fmax(fabs(vnl), fabs(vnr)) + fmax(cl, cr);
  } else {
    double an = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) an += P.a[m] * n[m];
    if constexpr (LAW == LAW_ADV) return fabs(an);
    else return fmax(fabs(an * L[0]), fabs(an * R[0]));
  }
}

// entropy variables (euler_navierstokes.jl:100-131); identity for scalar laws
template <int DIM, int LAW>
__device__ __forceinline__ void cons_to_entropy(const Phys& P, const double* u, double* w) {
  if constexpr (LAW == LAW_EULER) {
    double g = P.gamma, gm1 = g - 1.0;
    double k = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) k += u[1 + m] * u[1 + m];
    k *= 0.5 * frcp(u[0]);
    double p = gm1 * (u[DIM + 1] - k);
    double inv_p = frcp(p);
    // log(p / rho^gamma) = log p - gamma log rho (two logs instead of pow + log)
    w[0] = (g - (flog(p) - g * flog(u[0]))) * P.inv_gm1 - k * inv_p;
#pragma unroll
    for (int m = 0; m < DIM; ++m) w[1 + m] = u[1 + m] * inv_p;
    w[DIM + 1] = -u[0] * inv_p;
  } else {
    w[0] = u[0];
  }
}

template <int DIM, int LAW>
__device__ __forceinline__ void entropy_to_cons(const Phys& P, const double* w_in, double* u) {
  if constexpr (LAW == LAW_EULER) {
    double g = P.gamma, gm1 = g - 1.0, inv_gm1 = P.inv_gm1;
    double w[DIM + 2];
#pragma unroll
    for (int e = 0; e < DIM + 2; ++e) w[e] = w_in[e] * gm1;
    double k = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) k += w[1 + m] * w[1 + m];
    k = fdiv(k, 2.0 * w[DIM + 1]);
    double s = g - w[0] + k;
    // ((gm1 / (-w_last)^g)^(1/gm1)) exp(-s/gm1) = exp((log gm1 - g log(-w_last) - s) / gm1)
    double rho_e = fexp((P.log_gm1 - g * flog(-w[DIM + 1]) - s) * inv_gm1);
    u[0] = -w[DIM + 1] * rho_e;
#pragma unroll
    for (int m = 0; m < DIM; ++m) u[1 + m] = w[1 + m] * rho_e;
    u[DIM + 1] = rho_e * (1.0 - k);
  } else {
    u[0] = w_in[0];
  }
}

}  // namespace sse
