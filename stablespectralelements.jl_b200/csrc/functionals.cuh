// Analysis functionals on the device: the conservation / energy / entropy
// functionals and their time derivatives of /root/reference/src/Analysis/conservation.jl:113-190
// and the L2 error of Analysis/error.jl:58-91 (default error quadrature = volume quadrature),
// evaluated on the handle's device-resident arrays.
//
// One CTA per element writes N_out partial values; k_reduce_partials sums them in a fixed order
// (per-block contiguous chunks, then the host adds the <= 256 block results), so results are
// reproducible run to run.  Generic over element type / V kind / mass solver: these run once
// per analysis step, not per Runge-Kutta stage.
#pragma once
#include "kernels.cuh"

namespace sse {

enum { FN_CONSERVATION = 0, FN_ENTROPY = 1, FN_ENERGY = 2, FN_ENERGY_RESIDUAL = 3,
       FN_ENTROPY_RESIDUAL = 4, FN_L2_ERROR = 5 };

// mathematical entropy (euler_navierstokes.jl:93-98; 1/2 u^2 for the scalar laws)
template <int DIM, int LAW>
__device__ __forceinline__ double entropy_fn(const Phys& P, const double* u) {
  if constexpr (LAW == LAW_EULER) {
    double k = 0.0;
#pragma unroll
    for (int m = 0; m < DIM; ++m) k += u[1 + m] * u[1 + m];
    const double p = (P.gamma - 1.0) * (u[DIM + 1] - 0.5 * k / u[0]);
    return -u[0] * (log(p) - P.gamma * log(u[0])) * P.inv_gm1;
  } else {
    return 0.5 * u[0] * u[0];
  }
}

// sum of v over the CTA, returned to every thread (fixed order: lanes, then warps)
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int q = 0; q < nw; ++q) s += red[q];
  return s;
}

// shared: a[NC*Np] | b[NC*Np] | aq[NC*Nq] | bq[NC*Nq] | tmp[2*NC*n1^DIM] | cg[4*NC*Np] | red[32]
template <int DIM, int LAW>
__global__ void __launch_bounds__(128)
k_functional(Tables T, Geo G, Phys P, int which, const double* __restrict__ xa,
             const double* __restrict__ xb, double* __restrict__ partial, int n_out) {
  constexpr int NC = LawTraits<DIM, LAW>::NC;
  SSE_SHARED16(sm);
  const int Np = T.N_p, Nq = T.N_q;
  int wt = 1;
  for (int m = 0; m < DIM; ++m) wt *= (T.n1 > 0 ? T.n1 : 1);
  double* a = sm;
  double* b = a + NC * Np;
  double* aq = b + NC * Np;
  double* bq = aq + NC * Nq;
  double* tmp = bq + NC * Nq;
  double* cg = tmp + 2 * NC * wt;
  double* red = cg + 4 * NC * Np;
  const long long k = blockIdx.x;
  double out[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) out[c] = 0.0;

  SSE_LOOP(idx, NC * Np) a[idx] = xa[k * NC * Np + idx];
  const bool need_b = (which == FN_ENERGY_RESIDUAL || which == FN_ENTROPY_RESIDUAL);
  if (need_b) SSE_LOOP(idx, NC * Np) b[idx] = xb[k * NC * Np + idx];
  __syncthreads();

  if (which == FN_ENERGY || which == FN_ENERGY_RESIDUAL) {
    // out[c] = a_c^T M x_c with x = a (energy, times 1/2) or b (its time derivative)
    const double* x = (which == FN_ENERGY) ? a : b;
    if (T.mass_kind == MASS_DIAGONAL) {          // M = diag(W J), nodal scheme (N_p = N_q)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double v = 0.0;
        for (int i = threadIdx.x; i < Np; i += blockDim.x)
          v += T.W[i] * G.J_q[k * Nq + i] * a[c * Np + i] * x[c * Np + i];
        out[c] = block_sum(v, red);
      }
    } else if (T.mass_kind == MASS_CHOLESKY) {   // M = V^T W J V
      apply_V<DIM>(T, 1, NC, a, aq, tmp);
      apply_V<DIM>(T, 1, NC, x, bq, tmp);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double v = 0.0;
        for (int i = threadIdx.x; i < Nq; i += blockDim.x)
          v += T.W[i] * G.J_q[k * Nq + i] * aq[c * Nq + i] * bq[c * Nq + i];
        out[c] = block_sum(v, red);
      }
    } else {
      // weight-adjusted: M = (M^-1)^-1 where M^-1 is the operator mass_solve applies
      // (mass_matrix.jl:138-151 inverts it densely); y = M x by conjugate gradients on
      // M^-1 y = x, all components at once, each with its own step lengths.
      double* y = cg;
      double* r = y + NC * Np;
      double* p = r + NC * Np;
      double* Ap = p + NC * Np;
      double rs[NC], rs0[NC];
      SSE_LOOP(idx, NC * Np) { y[idx] = 0.0; r[idx] = x[idx]; p[idx] = x[idx]; }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double v = 0.0;
        for (int i = threadIdx.x; i < Np; i += blockDim.x) v += r[c * Np + i] * r[c * Np + i];
        rs[c] = rs0[c] = block_sum(v, red);
      }
      for (int it = 0; it < Np + 10; ++it) {
        bool done = true;
#pragma unroll
        for (int c = 0; c < NC; ++c) done = done && !(rs[c] > 1e-30 * rs0[c]);
        if (done) break;                       // uniform: rs comes from block_sum
        SSE_LOOP(idx, NC * Np) Ap[idx] = p[idx];
        __syncthreads();
        mass_solve<DIM>(T, G, k, 1, NC, Ap, aq, tmp);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double v = 0.0;
          for (int i = threadIdx.x; i < Np; i += blockDim.x) v += p[c * Np + i] * Ap[c * Np + i];
          const double pAp = block_sum(v, red);
          const double alpha = (rs[c] > 1e-30 * rs0[c] && pAp > 0.0) ? rs[c] / pAp : 0.0;
          double v2 = 0.0;
          for (int i = threadIdx.x; i < Np; i += blockDim.x) {
            y[c * Np + i] += alpha * p[c * Np + i];
            const double rr = r[c * Np + i] - alpha * Ap[c * Np + i];
            r[c * Np + i] = rr;
            v2 += rr * rr;
          }
          const double rsn = block_sum(v2, red);
          const double beta = rs[c] > 0.0 ? rsn / rs[c] : 0.0;
          for (int i = threadIdx.x; i < Np; i += blockDim.x)
            p[c * Np + i] = r[c * Np + i] + beta * p[c * Np + i];
          rs[c] = rsn;
        }
        __syncthreads();
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double v = 0.0;
        for (int i = threadIdx.x; i < Np; i += blockDim.x) v += a[c * Np + i] * y[c * Np + i];
        out[c] = block_sum(v, red);
      }
    }
    if (which == FN_ENERGY)
#pragma unroll
      for (int c = 0; c < NC; ++c) out[c] *= 0.5;
  } else {
    apply_V<DIM>(T, 1, NC, a, aq, tmp);
    if (need_b) apply_V<DIM>(T, 1, NC, b, bq, tmp);
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    for (int i = threadIdx.x; i < Nq; i += blockDim.x) {
      const double wj = T.W[i] * G.J_q[k * Nq + i];
      double uu[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) uu[c] = aq[c * Nq + i];
      if (which == FN_CONSERVATION) {
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] += wj * uu[c];
      } else if (which == FN_ENTROPY) {
        acc[0] += wj * entropy_fn<DIM, LAW>(P, uu);
      } else if (which == FN_ENTROPY_RESIDUAL) {
        // (P w)^T M dudt with P = M^-1 V^T W J and symmetric M:  w_q^T W J (V dudt)
        double w[NC];
        cons_to_entropy<DIM, LAW>(P, uu, w);
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) s += w[c] * bq[c * Nq + i];
        acc[0] += wj * s;
      } else {  // FN_L2_ERROR: xb = exact solution at the volume nodes, (N_q, N_c, N_e)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const double e = xb[(k * NC + c) * Nq + i] - uu[c];
          acc[c] += wj * e * e;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) out[c] = block_sum(acc[c], red);
  }
  if (threadIdx.x == 0)
    for (int c = 0; c < n_out; ++c) partial[k * n_out + c] = out[c < NC ? c : 0];
}

// block b sums elements [b*chunk, (b+1)*chunk) of partial[N_e][n_out] in index order
__global__ void __launch_bounds__(256)
k_reduce_partials(const double* __restrict__ partial, long long N_e, int n_out, long long chunk,
                  double* __restrict__ out) {
  __shared__ double red[8];
  const long long k0 = (long long)blockIdx.x * chunk;
  const long long k1 = min(N_e, k0 + chunk);
  for (int c = 0; c < n_out; ++c) {
    double v = 0.0;
    for (long long k = k0 + threadIdx.x; k < k1; k += blockDim.x) v += partial[k * n_out + c];
    const double s = block_sum(v, red);
    if (threadIdx.x == 0) out[blockIdx.x * n_out + c] = s;
  }
}

}  // namespace sse
