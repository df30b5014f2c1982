/*
 * CPU ORACLE, C/OpenMP edition (test + baseline infrastructure only; never linked into the
 * product).  Restates the reference's threaded residual for the flux-differencing form,
 *   semi_discrete_residual!  (src/Solvers/Solvers.jl:498-518: two `Threads.@threads for k`
 *   loops with a barrier between them),
 * with the reference's own algorithms for a ModalTensor/NodalTensor scheme:
 *   loop A  nodal_values!/entropy_projection!   src/Solvers/flux_differencing_form.jl:171-292
 *   loop B  time_derivative!                    src/Solvers/flux_differencing_form.jl:294-347
 *           flux_difference! (sparse S, one flux evaluation per direction m, :37-75)
 *           facet_correction! (sparse C, :130-168)
 *           numerical_flux!                     src/ConservationLaws/ConservationLaws.jl:75-128
 *           mass_matrix_solve! (weight-adjusted / diagonal) src/Solvers/mass_matrix.jl:169-196
 *   V, V^T  WarpedTensorProductMap3D/2D mul!    src/MatrixFreeOperators/warped_product_{2d,3d}.jl
 *           (dense fallback otherwise)
 *   physics src/ConservationLaws/euler_navierstokes.jl:100-195, ConservationLaws.jl:132-156
 * Parity of this file is pinned through tests/test_oracle_invariants.py (C vs NumPy oracle), which
 * in turn reproduces the reference's golden L2 errors (see oracle/sse_oracle.py header).
 *
 * Arrays use the reference's (Julia, column-major) layout, 0-based indices.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int d, N_p, N_q, N_f, N_c, num_faces, n1;   /* n1 > 0: warped tensor-product V tables */
  int64_t N_e;
  int law;            /* 0 advection, 1 burgers, 2 euler */
  double a[3], gamma;
  int inviscid;       /* 0 LF, 1 central, 2 EC */
  double half_lambda;
  int two_point;      /* 0 conservative, 1 EC */
  int proj;           /* 0 none, 1 nodal general R, 2 modal */
  int mass;           /* 0 diagonal, 1 weight adjusted (M^-1 = I) */
  int has_C;
  const double *V;                    /* dense N_q x N_p row-major */
  const double *wA, *wB, *wC; const int32_t *sig;
  const int32_t *R_rp, *R_ci; const double *R_v;      /* CSR R */
  const int32_t *S_cp[3], *S_ri[3]; const double *S_v[3];   /* CSC S_m (as SparseMatrixCSC) */
  const int32_t *C_cp, *C_ri; const double *C_v;      /* CSC C (N_q x N_f) */
  const double *W, *B, *n_ref;
  const double *J_q, *L_q, *J_f, *nJf;
  const int64_t *mapP;
} oracle_problem;

static inline double logmean(double x, double y) {
  double f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y);
  if (f2 < 1.0e-4) return (x + y) * 105 / (210 + f2 * (70 + f2 * (42 + f2 * 30)));
  return (y - x) / log(y / x);
}
static inline double inv_logmean(double x, double y) {
  double f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y);
  if (f2 < 1.0e-4) return (210 + f2 * (70 + f2 * (42 + f2 * 30))) / ((x + y) * 105);
  return log(y / x) / (y - x);
}

/* F[e + N_c*n] */
static void physical_flux(const oracle_problem *P, const double *u, double *F) {
  const int d = P->d;
  if (P->law == 0) { for (int n = 0; n < d; ++n) F[n] = P->a[n] * u[0]; return; }
  if (P->law == 1) { for (int n = 0; n < d; ++n) F[n] = 0.5 * P->a[n] * u[0] * u[0]; return; }
  const int Nc = d + 2;
  double V[3], vv = 0.0;
  for (int m = 0; m < d; ++m) { V[m] = u[m + 1] / u[0]; vv += u[m + 1] * V[m]; }
  double p = (P->gamma - 1.0) * (u[Nc - 1] - 0.5 * vv);
  for (int n = 0; n < d; ++n) {
    F[0 + Nc * n] = u[n + 1];
    for (int m = 0; m < d; ++m) F[m + 1 + Nc * n] = u[m + 1] * V[n] + (m == n ? p : 0.0);
    F[Nc - 1 + Nc * n] = (u[Nc - 1] + p) * V[n];
  }
}

static void two_point_flux(const oracle_problem *P, int kind, const double *uL, const double *uR,
                           double *F) {
  const int d = P->d;
  if (P->law == 0) { double f = 0.5 * (uL[0] + uR[0]); for (int n = 0; n < d; ++n) F[n] = P->a[n] * f; return; }
  if (P->law == 1) {
    double f = kind == 1 ? (uL[0] * uL[0] + uL[0] * uR[0] + uR[0] * uR[0]) / 6 : (uL[0] * uL[0] + uR[0] * uR[0]) * 0.25;
    for (int n = 0; n < d; ++n) F[n] = P->a[n] * f;
    return;
  }
  const int Nc = d + 2;
  if (kind == 0) {
    double FL[15], FR[15];
    physical_flux(P, uL, FL); physical_flux(P, uR, FR);
    for (int q = 0; q < Nc * d; ++q) F[q] = 0.5 * (FL[q] + FR[q]);
    return;
  }
  const double gm1 = P->gamma - 1.0, inv_gm1 = 1.0 / gm1;
  double VL[3], VR[3], vl2 = 0, vr2 = 0, vlvr = 0, Va[3];
  for (int m = 0; m < d; ++m) {
    VL[m] = uL[m + 1] / uL[0]; VR[m] = uR[m + 1] / uR[0];
    vl2 += VL[m] * VL[m]; vr2 += VR[m] * VR[m]; vlvr += VL[m] * VR[m];
    Va[m] = 0.5 * (VL[m] + VR[m]);
  }
  double pL = gm1 * (uL[Nc - 1] - 0.5 * uL[0] * vl2), pR = gm1 * (uR[Nc - 1] - 0.5 * uR[0] * vr2);
  double rho = logmean(uL[0], uR[0]), pa = 0.5 * (pL + pR);
  double C = 0.5 * vlvr + inv_gm1 * inv_logmean(uL[0] / pL, uR[0] / pR);
  for (int n = 0; n < d; ++n) {
    double fr = rho * Va[n];
    F[0 + Nc * n] = fr;
    for (int m = 0; m < d; ++m) F[m + 1 + Nc * n] = rho * Va[m] * Va[n] + (m == n ? pa : 0.0);
    F[Nc - 1 + Nc * n] = fr * C + 0.5 * (pL * VR[n] + pR * VL[n]);
  }
}

static double wave_speed(const oracle_problem *P, const double *ui, const double *uo, const double *n) {
  const int d = P->d;
  if (P->law == 0) { double an = 0; for (int m = 0; m < d; ++m) an += P->a[m] * n[m]; return fabs(an); }
  if (P->law == 1) { double an = 0; for (int m = 0; m < d; ++m) an += P->a[m] * n[m]; return fmax(fabs(an * ui[0]), fabs(an * uo[0])); }
  const int Nc = d + 2;
  const double g = P->gamma, gm1 = g - 1.0;
  double vni = 0, vno = 0, ki = 0, ko = 0;
  for (int m = 0; m < d; ++m) {
    vni += ui[m + 1] / ui[0] * n[m]; vno += uo[m + 1] / uo[0] * n[m];
    ki += ui[m + 1] * ui[m + 1]; ko += uo[m + 1] * uo[m + 1];
  }
  double pi_ = gm1 * (ui[Nc - 1] - (0.5 / ui[0]) * ki), po = gm1 * (uo[Nc - 1] - (0.5 / uo[0]) * ko);
  return fmax(fabs(vni), fabs(vno)) + fmax(sqrt(g * pi_ / ui[0]), sqrt(g * po / uo[0]));
}

static void cons_to_entropy(const oracle_problem *P, const double *u, double *w) {
  if (P->law != 2) { w[0] = u[0]; return; }
  const int d = P->d; const double g = P->gamma, gm1 = g - 1.0;
  double k = 0; for (int m = 0; m < d; ++m) k += u[m + 1] * u[m + 1];
  k *= 0.5 / u[0];
  double p = gm1 * (u[d + 1] - k), ip = 1.0 / p;
  w[0] = (1.0 / gm1) * (g - log(p / pow(u[0], g))) - k * ip;
  for (int m = 0; m < d; ++m) w[m + 1] = u[m + 1] * ip;
  w[d + 1] = -u[0] * ip;
}
static void entropy_to_cons(const oracle_problem *P, const double *win, double *u) {
  if (P->law != 2) { u[0] = win[0]; return; }
  const int d = P->d; const double g = P->gamma, gm1 = g - 1.0, igm1 = 1.0 / gm1;
  double w[5]; for (int e = 0; e < d + 2; ++e) w[e] = win[e] * gm1;
  double k = 0; for (int m = 0; m < d; ++m) k += w[m + 1] * w[m + 1];
  k /= 2 * w[d + 1];
  double s = g - w[0] + k;
  double re = pow(gm1 / pow(-w[d + 1], g), igm1) * exp(-s * igm1);
  u[0] = -w[d + 1] * re;
  for (int m = 0; m < d; ++m) u[m + 1] = w[m + 1] * re;
  u[d + 1] = re * (1 - k);
}

/* y (N_q) = V x (N_p): WarpedTensorProductMap mul! or dense */
static void apply_V(const oracle_problem *P, const double *x, double *y, double *Z, double *Wt) {
  const int n = P->n1, d = P->d;
  if (n == 0) {
    for (int i = 0; i < P->N_q; ++i) { double t = 0; for (int p = 0; p < P->N_p; ++p) t += P->V[i * P->N_p + p] * x[p]; y[i] = t; }
    return;
  }
  if (d == 2) {
    for (int a2 = 0; a2 < n; ++a2) for (int b1 = 0; b1 < n; ++b1) {
      double t = 0; for (int b2 = 0; b2 < n - b1; ++b2) t += P->wB[(a2 * n + b1) * n + b2] * x[P->sig[b1 * n + b2]];
      Z[b1 * n + a2] = t; }
    for (int a1 = 0; a1 < n; ++a1) for (int a2 = 0; a2 < n; ++a2) {
      double t = 0; for (int b1 = 0; b1 < n; ++b1) t += P->wA[a1 * n + b1] * Z[b1 * n + a2];
      y[a1 * n + a2] = t; }
    return;
  }
  for (int b1 = 0; b1 < n; ++b1) for (int b2 = 0; b2 < n - b1; ++b2) for (int a3 = 0; a3 < n; ++a3) {
    double t = 0; for (int b3 = 0; b3 < n - b1 - b2; ++b3) t += P->wC[((a3 * n + b1) * n + b2) * n + b3] * x[P->sig[(b1 * n + b2) * n + b3]];
    Z[(b1 * n + b2) * n + a3] = t; }
  for (int b1 = 0; b1 < n; ++b1) for (int a2 = 0; a2 < n; ++a2) for (int a3 = 0; a3 < n; ++a3) {
    double t = 0; for (int b2 = 0; b2 < n - b1; ++b2) t += P->wB[(a2 * n + b1) * n + b2] * Z[(b1 * n + b2) * n + a3];
    Wt[(b1 * n + a2) * n + a3] = t; }
  for (int a1 = 0; a1 < n; ++a1) for (int a2 = 0; a2 < n; ++a2) for (int a3 = 0; a3 < n; ++a3) {
    double t = 0; for (int b1 = 0; b1 < n; ++b1) t += P->wA[a1 * n + b1] * Wt[(b1 * n + a2) * n + a3];
    y[(a1 * n + a2) * n + a3] = t; }
}
/* y (N_p) = V^T x (N_q) */
static void apply_Vt(const oracle_problem *P, const double *x, double *y, double *Z, double *Wt) {
  const int n = P->n1, d = P->d;
  if (n == 0) {
    for (int p = 0; p < P->N_p; ++p) { double t = 0; for (int i = 0; i < P->N_q; ++i) t += P->V[i * P->N_p + p] * x[i]; y[p] = t; }
    return;
  }
  if (d == 2) {
    for (int b1 = 0; b1 < n; ++b1) for (int a2 = 0; a2 < n; ++a2) {
      double t = 0; for (int a1 = 0; a1 < n; ++a1) t += P->wA[a1 * n + b1] * x[a1 * n + a2];
      Z[b1 * n + a2] = t; }
    for (int b1 = 0; b1 < n; ++b1) for (int b2 = 0; b2 < n - b1; ++b2) {
      double t = 0; for (int a2 = 0; a2 < n; ++a2) t += P->wB[(a2 * n + b1) * n + b2] * Z[b1 * n + a2];
      y[P->sig[b1 * n + b2]] = t; }
    return;
  }
  for (int b1 = 0; b1 < n; ++b1) for (int a2 = 0; a2 < n; ++a2) for (int a3 = 0; a3 < n; ++a3) {
    double t = 0; for (int a1 = 0; a1 < n; ++a1) t += P->wA[a1 * n + b1] * x[(a1 * n + a2) * n + a3];
    Wt[(b1 * n + a2) * n + a3] = t; }
  for (int b1 = 0; b1 < n; ++b1) for (int b2 = 0; b2 < n - b1; ++b2) for (int a3 = 0; a3 < n; ++a3) {
    double t = 0; for (int a2 = 0; a2 < n; ++a2) t += P->wB[(a2 * n + b1) * n + b2] * Wt[(b1 * n + a2) * n + a3];
    Z[(b1 * n + b2) * n + a3] = t; }
  for (int b1 = 0; b1 < n; ++b1) for (int b2 = 0; b2 < n - b1; ++b2) for (int b3 = 0; b3 < n - b1 - b2; ++b3) {
    double t = 0; for (int a3 = 0; a3 < n; ++a3) t += P->wC[((a3 * n + b1) * n + b2) * n + b3] * Z[(b1 * n + b2) * n + a3];
    y[P->sig[(b1 * n + b2) * n + b3]] = t; }
}

static void mass_solve(const oracle_problem *P, int64_t k, double *rhs, double *tmp, double *Z, double *Wt) {
  const int Nq = P->N_q, Np = P->N_p, Nc = P->N_c;
  if (P->mass == 0) { for (int c = 0; c < Nc; ++c) for (int i = 0; i < Np; ++i) rhs[i + Np * c] /= P->W[i] * P->J_q[i + Nq * k]; return; }
  for (int c = 0; c < Nc; ++c) {
    apply_V(P, rhs + Np * c, tmp, Z, Wt);
    for (int i = 0; i < Nq; ++i) tmp[i] *= P->W[i] / P->J_q[i + Nq * k];
    apply_Vt(P, tmp, rhs + Np * c, Z, Wt);
  }
}

int oracle_threads(void) { return omp_get_max_threads(); }
/* bench.py's reference arm: use every host core even when the launcher exported OMP_NUM_THREADS=1
 * (torch.distributed.run does); the reference's own Threads.@threads runs on `julia -t auto`. */
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* u, dudt: (N_p, N_c, N_e); u_q: (N_q, N_c, N_e); u_f: (N_f, N_e, N_c) ("switched order",
 * Solvers.jl:205) -- both scratch arrays are caller-allocated like PreAllocatedArraysFirstOrder. */
int oracle_residual_fluxdiff_range(const oracle_problem *P, const double *u, double *dudt, double *u_q,
                                   double *u_f, int64_t k0, int64_t k1);
int oracle_residual_fluxdiff(const oracle_problem *P, const double *u, double *dudt, double *u_q, double *u_f) {
  return oracle_residual_fluxdiff_range(P, u, dudt, u_q, u_f, 0, P->N_e);
}

/* The two element loops restricted to [k0, k1) (bounded sample of a large mesh for bench.py's
 * reference arm).  Loop B reads the exterior traces u_f of neighbours that may lie outside the
 * range: the caller runs the full-range call once beforehand so that they hold valid values. */
int oracle_residual_fluxdiff_range(const oracle_problem *P, const double *u, double *dudt, double *u_q,
                                   double *u_f, int64_t k0, int64_t k1) {
  const int d = P->d, Np = P->N_p, Nq = P->N_q, Nf = P->N_f, Nc = P->N_c;
  const int64_t Ne = P->N_e;
  const int npf = Nf / P->num_faces;
  const int n3 = P->n1 ? P->n1 * P->n1 * P->n1 : 1;
#pragma omp parallel
  {
    double *w_q = malloc(sizeof(double) * Nq * Nc), *w_f = malloc(sizeof(double) * Nf * Nc);
    double *w = malloc(sizeof(double) * Np * Nc), *tmp = malloc(sizeof(double) * Nq);
    double *Z = malloc(sizeof(double) * n3), *Wt = malloc(sizeof(double) * n3);
    double *r_q = malloc(sizeof(double) * Nq * Nc), *f_f = malloc(sizeof(double) * Nf * Nc);
    double *halfnJq = malloc(sizeof(double) * d * P->num_faces * Nq);
    /* ---- loop A: nodal_values! */
#pragma omp for schedule(static)
    for (int64_t k = k0; k < k1; ++k) {
      double *uq = u_q + (int64_t)Nq * Nc * k;
      for (int c = 0; c < Nc; ++c) apply_V(P, u + (int64_t)Np * (c + Nc * k), uq + Nq * c, Z, Wt);
      if (P->proj == 0) {
        for (int c = 0; c < Nc; ++c) for (int j = 0; j < Nf; ++j) {
          double t = 0; for (int e = P->R_rp[j]; e < P->R_rp[j + 1]; ++e) t += P->R_v[e] * uq[P->R_ci[e] + Nq * c];
          u_f[j + Nf * (k + Ne * c)] = t; }
        continue;
      }
      for (int i = 0; i < Nq; ++i) {
        double ui[5], wi[5];
        for (int c = 0; c < Nc; ++c) ui[c] = uq[i + Nq * c];
        cons_to_entropy(P, ui, wi);
        double sc = P->proj == 2 ? P->W[i] * P->J_q[i + Nq * k] : 1.0;
        for (int c = 0; c < Nc; ++c) w_q[i + Nq * c] = wi[c] * sc;
      }
      if (P->proj == 2) {
        for (int c = 0; c < Nc; ++c) apply_Vt(P, w_q + Nq * c, w + Np * c, Z, Wt);
        mass_solve(P, k, w, tmp, Z, Wt);
        for (int c = 0; c < Nc; ++c) apply_V(P, w + Np * c, w_q + Nq * c, Z, Wt);
      }
      for (int c = 0; c < Nc; ++c) for (int j = 0; j < Nf; ++j) {
        double t = 0; for (int e = P->R_rp[j]; e < P->R_rp[j + 1]; ++e) t += P->R_v[e] * w_q[P->R_ci[e] + Nq * c];
        w_f[j + Nf * c] = t; }
      if (P->proj == 2)
        for (int i = 0; i < Nq; ++i) {
          double wi[5], ui[5];
          for (int c = 0; c < Nc; ++c) wi[c] = w_q[i + Nq * c];
          entropy_to_cons(P, wi, ui);
          for (int c = 0; c < Nc; ++c) uq[i + Nq * c] = ui[c];
        }
      for (int j = 0; j < Nf; ++j) {
        double wi[5], ui[5];
        for (int c = 0; c < Nc; ++c) wi[c] = w_f[j + Nf * c];
        entropy_to_cons(P, wi, ui);
        for (int c = 0; c < Nc; ++c) u_f[j + Nf * (k + Ne * c)] = ui[c];
      }
    }
    /* implicit barrier: all traces written */
    /* ---- loop B: time_derivative! */
#pragma omp for schedule(static)
    for (int64_t k = k0; k < k1; ++k) {
      double *uq = u_q + (int64_t)Nq * Nc * k;
      const double *Lq = P->L_q + (int64_t)Nq * d * d * k;
      /* numerical flux, scaled by B J_f */
      for (int j = 0; j < Nf; ++j) {
        double ui[5], uo[5], n[3], F[15];
        int64_t g = P->mapP[j + (int64_t)Nf * k];
        int64_t jo = g % Nf, ko = g / Nf;
        double Jf = P->J_f[j + (int64_t)Nf * k];
        for (int c = 0; c < Nc; ++c) { ui[c] = u_f[j + Nf * (k + Ne * c)]; uo[c] = u_f[jo + Nf * (ko + Ne * c)]; }
        for (int m = 0; m < d; ++m) n[m] = P->nJf[m + d * (j + (int64_t)Nf * k)] / Jf;
        two_point_flux(P, P->two_point, ui, uo, F);
        double a = P->inviscid == 0 ? P->half_lambda * wave_speed(P, ui, uo, n) : 0.0;
        for (int c = 0; c < Nc; ++c) {
          double t = 0; for (int m = 0; m < d; ++m) t += F[c + Nc * m] * n[m];
          if (P->inviscid == 0) t += a * (ui[c] - uo[c]);
          f_f[j + Nf * c] = P->B[j] * Jf * t;
        }
      }
      /* flux_difference! (sparse: per direction m, re-evaluating the two-point flux) */
      memset(r_q, 0, sizeof(double) * Nq * Nc);
      for (int m = 0; m < d; ++m)
        for (int j = 0; j < Nq; ++j)
          for (int e = P->S_cp[m][j]; e < P->S_cp[m][j + 1]; ++e) {
            int i = P->S_ri[m][e];
            if (i >= j) continue;
            double ui[5], uj[5], F[15];
            for (int c = 0; c < Nc; ++c) { ui[c] = uq[i + Nq * c]; uj[c] = uq[j + Nq * c]; }
            two_point_flux(P, P->two_point, ui, uj, F);
            double Sm = P->S_v[m][e];
            for (int c = 0; c < Nc; ++c) {
              double Fm = 0;
              for (int n = 0; n < d; ++n) Fm += (Lq[i + Nq * (m + d * n)] + Lq[j + Nq * (m + d * n)]) * F[c + Nc * n];
              double df = Sm * Fm;
              r_q[i + Nq * c] -= df; r_q[j + Nq * c] += df;
            }
          }
      /* facet_correction! */
      if (P->has_C) {
        for (int i = 0; i < Nq; ++i) for (int f = 0; f < P->num_faces; ++f) for (int n = 0; n < d; ++n) {
          double t = 0; for (int m = 0; m < d; ++m) t += Lq[i + Nq * (m + d * n)] * P->n_ref[f * d + m];
          halfnJq[n + d * (f + P->num_faces * i)] = 0.5 * t; }
        for (int j = 0; j < Nf; ++j)
          for (int e = P->C_cp[j]; e < P->C_cp[j + 1]; ++e) {
            int i = P->C_ri[e];
            double ui[5], uj[5], F[15];
            for (int c = 0; c < Nc; ++c) { ui[c] = uq[i + Nq * c]; uj[c] = u_f[j + Nf * (k + Ne * c)]; }
            two_point_flux(P, P->two_point, ui, uj, F);
            int f = j / npf;
            for (int c = 0; c < Nc; ++c) {
              double t = 0;
              for (int m = 0; m < d; ++m) t += (0.5 * P->nJf[m + d * (j + (int64_t)Nf * k)] + halfnJq[m + d * (f + P->num_faces * i)]) * F[c + Nc * m];
              double df = P->C_v[e] * t;
              r_q[i + Nq * c] -= df; f_f[j + Nf * c] -= df;
            }
          }
      }
      /* r_q -= R^T f_f */
      for (int c = 0; c < Nc; ++c) for (int j = 0; j < Nf; ++j) {
        double ff = f_f[j + Nf * c];
        for (int e = P->R_rp[j]; e < P->R_rp[j + 1]; ++e) r_q[P->R_ci[e] + Nq * c] -= P->R_v[e] * ff; }
      double *du = dudt + (int64_t)Np * Nc * k;
      for (int c = 0; c < Nc; ++c) apply_Vt(P, r_q + Nq * c, du + Np * c, Z, Wt);
      mass_solve(P, k, du, tmp, Z, Wt);
    }
    free(w_q); free(w_f); free(w); free(tmp); free(Z); free(Wt); free(r_q); free(f_f); free(halfnJq);
  }
  return 0;
}
