// On-device geometric factors: GeometricFactors(mesh, reference_element,
// metric_type) of /root/reference/src/SpatialDiscretizations/mesh.jl:213-509 evaluated from the
// mapping-node coordinates, one CTA per element, every per-element intermediate in shared memory.
// The reference-element matrices (a few tens of KB, identical for all elements) are read through
// the read-only path.  Runs once at setup; what it saves is the host-side metric computation and
// the multi-GB host-to-device upload of J_q / Lambda_q / J_f / nJf for million-element meshes.
//
// Output layouts are the reference's (Julia, column-major), see include/sse_b200.h:
//   J_q (N_q, N_e); Lambda_q (N_q, d, d, N_e) = J dxi_l/dx_m at [i, l, m, k];
//   J_f (N_f, N_e); nJf (d, N_f, N_e).
#pragma once
#ifndef SSE_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace sse {

struct MapOps {
  int Nm, Nm1, Nq, Nf;
  const double* D[3];     // (Nm x Nm) row-major
  const double *Vq, *Vf;  // (Nq x Nm), (Nf x Nm)
  const double* P;        // (Nm1 x Nm) or nullptr (identity)
  const double* D1[3];    // (Nm1 x Nm1)
  const double *Vq1, *Vf1;
  const double* nrstJ;    // (Nf x d)
  const double* Jproj;    // (Nq x Nq) or nullptr
  const double* xyz[3];   // device, (Nm, N_e)
};

// J and Lambda[l][m] = J dxi_l/dx_m from dxdr[m][n] = dx_m/dxi_n (mesh.jl:213-229)
template <int DIM>
__device__ __forceinline__ void metrics_from_dxdr(const double (&a)[DIM][DIM], double& J,
                                                  double (&L)[DIM][DIM]) {
  if constexpr (DIM == 1) {
    J = a[0][0];
    L[0][0] = 1.0;
  } else if constexpr (DIM == 2) {
    J = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    L[0][0] = a[1][1];  L[0][1] = -a[0][1];
    L[1][0] = -a[1][0]; L[1][1] = a[0][0];
  } else {
    L[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    L[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    L[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    L[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    L[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    L[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    L[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    L[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    L[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    J = a[0][0] * L[0][0] + a[0][1] * L[1][0] + a[0][2] * L[2][0];
  }
}

// facet normals from the metric terms at a facet node: nJf[m] = sum_n L[n][m] nrstJ[n]
// (mesh.jl:273-281)
template <int DIM>
__device__ __forceinline__ void store_facet(const double (&L)[DIM][DIM], const double* nr,
                                            long long k, int i, int Nf, double* J_f, double* nJf) {
  double s = 0.0;
#pragma unroll
  for (int m = 0; m < DIM; ++m) {
    double v = 0.0;
#pragma unroll
    for (int n = 0; n < DIM; ++n) v = fma(L[n][m], nr[n], v);
    nJf[(k * Nf + i) * DIM + m] = v;
    s = fma(v, v, s);
  }
  J_f[k * Nf + i] = sqrt(s);
}

// ExactMetrics (mesh.jl:231-284): derivatives of the mapping at the mapping nodes, interpolated
// to the volume / facet quadrature nodes, metrics evaluated there; optional L2 projection of J.
// shared: X[DIM][Nm] | dX[DIM*DIM][Nm] | Jq[Nq]
template <int DIM>
__global__ void __launch_bounds__(128)
k_geometry_exact(MapOps M, long long N_e, double* __restrict__ J_q, double* __restrict__ L_q,
                 double* __restrict__ J_f, double* __restrict__ nJf) {
  SSE_SHARED16(sm);
  const int Nm = M.Nm, Nq = M.Nq, Nf = M.Nf;
  double* X = sm;
  double* dX = X + DIM * Nm;
  double* Jq = dX + DIM * DIM * Nm;
  const long long k = blockIdx.x;
  for (int idx = threadIdx.x; idx < DIM * Nm; idx += blockDim.x)
    X[idx] = M.xyz[idx / Nm][k * Nm + idx % Nm];
  __syncthreads();
  for (int idx = threadIdx.x; idx < DIM * DIM * Nm; idx += blockDim.x) {
    const int a = idx % Nm, n = (idx / Nm) % DIM, m = idx / (Nm * DIM);
    const double* Dn = M.D[n] + a * Nm;
    double acc = 0.0;
    for (int b = 0; b < Nm; ++b) acc = fma(__ldg(Dn + b), X[m * Nm + b], acc);
    dX[(m * DIM + n) * Nm + a] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < Nq + Nf; idx += blockDim.x) {
    const bool vol = idx < Nq;
    const int i = vol ? idx : idx - Nq;
    const double* Vr = (vol ? M.Vq : M.Vf) + (long long)i * Nm;
    double g[DIM][DIM], L[DIM][DIM], J;
#pragma unroll
    for (int m = 0; m < DIM; ++m)
#pragma unroll
      for (int n = 0; n < DIM; ++n) g[m][n] = 0.0;
    for (int a = 0; a < Nm; ++a) {
      const double v = __ldg(Vr + a);
#pragma unroll
      for (int m = 0; m < DIM; ++m)
#pragma unroll
        for (int n = 0; n < DIM; ++n) g[m][n] = fma(v, dX[(m * DIM + n) * Nm + a], g[m][n]);
    }
    metrics_from_dxdr<DIM>(g, J, L);
    if (vol) {
      Jq[i] = J;
#pragma unroll
      for (int l = 0; l < DIM; ++l)
#pragma unroll
        for (int m = 0; m < DIM; ++m) L_q[((k * DIM + m) * DIM + l) * Nq + i] = L[l][m];
    } else {
      store_facet<DIM>(L, M.nrstJ + i * DIM, k, i, Nf, J_f, nJf);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Nq; i += blockDim.x) {
    double v = Jq[i];
    if (M.Jproj) {   // SpatialDiscretizations.jl:311-318
      v = 0.0;
      for (int j = 0; j < Nq; ++j) v = fma(__ldg(M.Jproj + (long long)i * Nq + j), Jq[j], v);
    }
    J_q[k * Nq + i] = v;
  }
}

// ConservativeCurlMetrics in 3-D (mesh.jl:286-509; Kopriva's conservative curl form as in
// StartUpDG's geometric_factors): J from the degree-N mapping, the nine metric terms from the
// curl form evaluated on the degree-(N+1) nodes (Tet: P prolongs N -> N+1; Hex: P = identity),
// interpolated to the quadrature nodes.
// shared: X[3][Nm] | dX[9][Nm] | Jm[Nm] | X1[3][Nm1] | dX1[9][Nm1] | F[3][Nm1] | Lm[9][Nm1]
__global__ void __launch_bounds__(128)
k_geometry_curl3d(MapOps M, long long N_e, double* __restrict__ J_q, double* __restrict__ L_q,
                  double* __restrict__ J_f, double* __restrict__ nJf) {
  SSE_SHARED16(sm);
  const int Nm = M.Nm, Nm1 = M.Nm1, Nq = M.Nq, Nf = M.Nf;
  double* X = sm;
  double* dX = X + 3 * Nm;
  double* Jm = dX + 9 * Nm;
  double* X1 = Jm + Nm;
  double* dX1 = X1 + 3 * Nm1;
  double* F = dX1 + 9 * Nm1;
  double* Lm = F + 3 * Nm1;
  const long long k = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < 3 * Nm; idx += nt) X[idx] = M.xyz[idx / Nm][k * Nm + idx % Nm];
  __syncthreads();
  // degree-N derivatives -> J at the mapping nodes; prolongation to the degree-(N+1) nodes
  for (int idx = tid; idx < 9 * Nm; idx += nt) {
    const int a = idx % Nm, n = (idx / Nm) % 3, m = idx / (3 * Nm);
    const double* Dn = M.D[n] + a * Nm;
    double acc = 0.0;
    for (int b = 0; b < Nm; ++b) acc = fma(__ldg(Dn + b), X[m * Nm + b], acc);
    dX[(m * 3 + n) * Nm + a] = acc;
  }
  for (int idx = tid; idx < 3 * Nm1; idx += nt) {
    const int a = idx % Nm1, m = idx / Nm1;
    double acc;
    if (M.P) {
      acc = 0.0;
      for (int b = 0; b < Nm; ++b) acc = fma(__ldg(M.P + a * Nm + b), X[m * Nm + b], acc);
    } else {
      acc = X[m * Nm + a];
    }
    X1[idx] = acc;
  }
  __syncthreads();
  for (int a = tid; a < Nm; a += nt) {
    const double xr = dX[0 * Nm + a], xs = dX[1 * Nm + a], xt = dX[2 * Nm + a];
    const double yr = dX[3 * Nm + a], ys = dX[4 * Nm + a], yt = dX[5 * Nm + a];
    const double zr = dX[6 * Nm + a], zs = dX[7 * Nm + a], zt = dX[8 * Nm + a];
    Jm[a] = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt);
  }
  for (int idx = tid; idx < 9 * Nm1; idx += nt) {
    const int a = idx % Nm1, n = (idx / Nm1) % 3, m = idx / (3 * Nm1);
    const double* Dn = M.D1[n] + a * Nm1;
    double acc = 0.0;
    for (int b = 0; b < Nm1; ++b) acc = fma(__ldg(Dn + b), X1[m * Nm1 + b], acc);
    dX1[(m * 3 + n) * Nm1 + a] = acc;
  }
  __syncthreads();
  // three curls: column m of Lambda from (u, v) = (y, z), (x, z) [negated], (y, x) [negated]
  for (int c = 0; c < 3; ++c) {
    const int u = (c == 1) ? 0 : 1;          // derivative taken of x_u
    const int v = (c == 2) ? 0 : 2;          // multiplied by x_v
    const double sgn = (c == 0) ? 1.0 : -1.0;
    for (int idx = tid; idx < 3 * Nm1; idx += nt) {
      const int a = idx % Nm1, n = idx / Nm1;
      F[idx] = dX1[(u * 3 + n) * Nm1 + a] * X1[v * Nm1 + a];
    }
    __syncthreads();
    for (int idx = tid; idx < 3 * Nm1; idx += nt) {
      const int a = idx % Nm1, l = idx / Nm1;
      // (r, s, t) components: Dt Fs - Ds Ft, Dr Ft - Dt Fr, Ds Fr - Dr Fs
      const int p = (l + 2) % 3, q = (l + 1) % 3;   // D_p F_q - D_q F_p
      const double* Dp = M.D1[p] + a * Nm1;
      const double* Dq = M.D1[q] + a * Nm1;
      double acc = 0.0;
      for (int b = 0; b < Nm1; ++b)
        acc += __ldg(Dp + b) * F[q * Nm1 + b] - __ldg(Dq + b) * F[p * Nm1 + b];
      Lm[(l * 3 + c) * Nm1 + a] = sgn * acc;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < Nq + Nf; idx += nt) {
    const bool vol = idx < Nq;
    const int i = vol ? idx : idx - Nq;
    const double* Vr = (vol ? M.Vq1 : M.Vf1) + (long long)i * Nm1;
    double L[3][3];
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int m = 0; m < 3; ++m) L[l][m] = 0.0;
    for (int a = 0; a < Nm1; ++a) {
      const double w = __ldg(Vr + a);
#pragma unroll
      for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int m = 0; m < 3; ++m) L[l][m] = fma(w, Lm[(l * 3 + m) * Nm1 + a], L[l][m]);
    }
    if (vol) {
      double J = 0.0;
      const double* Vj = M.Vq + (long long)i * Nm;
      for (int a = 0; a < Nm; ++a) J = fma(__ldg(Vj + a), Jm[a], J);
      J_q[k * Nq + i] = J;
#pragma unroll
      for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int m = 0; m < 3; ++m) L_q[((k * 3 + m) * 3 + l) * Nq + i] = L[l][m];
    } else {
      store_facet<3>(L, M.nrstJ + i * 3, k, i, Nf, J_f, nJf);
    }
  }
}

}  // namespace sse
