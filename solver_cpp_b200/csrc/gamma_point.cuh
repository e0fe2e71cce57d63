// Per-k-point arithmetic of the fused z pass, shared by the shared-memory-only kernel (gamma_kernels.cu) and the
// register-resident one (gamma_fast.cu): xi from the per-axis tables, divergence contribution, Gamma contraction.
// Follows MechBase::calculate_divergence_rms / gamma_convolution (/root/reference/src/grid/mech_base.cpp:302-477); see
// gamma_kernels.cu for the operator definition.
#pragma once
#include "fft_host.h"

__device__ __forceinline__ void load_xi(const DmkGammaParams& gp, int kx, int ky, int kz, double2 (&xi)[3]) {
  double2 ax = __ldg(&gp.xi_x[kx]), ay = __ldg(&gp.xi_y[ky]), az = __ldg(&gp.xi_z[kz]);
  double2 ox = __ldg(&gp.xo_x[kx]), oy = __ldg(&gp.xo_y[ky]), oz = __ldg(&gp.xo_z[kz]);
  xi[0] = cmul(cmul(ax, oy), oz);
  xi[1] = cmul(cmul(ay, oz), ox);
  xi[2] = cmul(cmul(az, ox), oy);
  if (__ldg(&gp.nyq_x[kx])) xi[0] = make_double2(0.0, 0.0);
  if (__ldg(&gp.nyq_y[ky])) xi[1] = make_double2(0.0, 0.0);
  if (__ldg(&gp.nyq_z[kz])) xi[2] = make_double2(0.0, 0.0);
}

// One k-point, xi purely imaginary (continuous / central-difference stencils): xi = i q, so D(l,m) = conj(-xi_l) xi_m =
// -q_l q_m is real, symmetric and of rank one.  That structure is used twice:
//   A(l,m) = sum_{n,o} C_ref(l,n,m,o) D(n,o) = sum_v Csym(l,m,v) D_v over the 6 distinct (n,o) pairs, with
//            Csym(l,m,v) = C_ref(l,n,m,o) + C_ref(l,o,m,n) (n != o) precomputed on the host (dmk_update_gamma): 54 FMA, not 81;
//   A^-1 . tau . D = -(A^-1 (tau q)) q^T : two 3x3 matrix-vector products and an outer product per real/imaginary part
//            (54 flop) instead of two 3x3 matrix-matrix products (108).
// t = 9 complex in record order (component c = i + 3 j).  Csym: 54 doubles, index (l*3+m)*6 + v, v = (00,11,22,12,02,01).
// Returns the divergence contribution.
__device__ __forceinline__ double gamma_point_imag(double2* __restrict__ t, const double* __restrict__ Csym,
                                                   const double q[3], const double resc[3], bool is_dc,
                                                   const double* __restrict__ dc_value) {
  // divergence: prod_l = sum_m tau(l,m) * conj(-xi_m) * resc_m = i * sum_m tau(l,m) q_m resc_m
  double div = 0.0;
  const double v0 = q[0] * resc[0], v1 = q[1] * resc[1], v2 = q[2] * resc[2];
#pragma unroll
  for (int l = 0; l < 3; l++) {
    double pr = t[l].x * v0 + t[l + 3].x * v1 + t[l + 6].x * v2;
    double pi = t[l].y * v0 + t[l + 3].y * v1 + t[l + 6].y * v2;
    div += pr * pr + pi * pi;
  }
  if (is_dc) {
#pragma unroll
    for (int c = 0; c < 9; c++) t[c] = make_double2(dc_value[c], 0.0);
    return div;
  }
  const double Dv[6] = {-q[0] * q[0], -q[1] * q[1], -q[2] * q[2], -q[1] * q[2], -q[0] * q[2], -q[0] * q[1]};
  double A[3][3];
#pragma unroll
  for (int l = 0; l < 3; l++)
#pragma unroll
    for (int m = 0; m < 3; m++) {
      double acc = 0.0;
#pragma unroll
      for (int v = 0; v < 6; v++) acc += Csym[(l * 3 + m) * 6 + v] * Dv[v];
      A[l][m] = acc;
    }
  double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
  if (!(fabs(det) > 1.0e-16)) {
#pragma unroll
    for (int c = 0; c < 9; c++) t[c] = make_double2(0.0, 0.0);
    return div;
  }
  double id = 1.0 / det;
  double Ai[3][3];
  Ai[0][0] = c00 * id; Ai[1][0] = c01 * id; Ai[2][0] = c02 * id;
  Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
  Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
  Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
  Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
  Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
  Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
  // out(l,m) = -u_l q_m,  u = Ai w,  w_n = sum_o tau(n,o) q_o   (real and imaginary parts separately: Ai, q real)
  double2 u[3];
  {
    double2 w[3];
#pragma unroll
    for (int n = 0; n < 3; n++) {
      w[n].x = t[n].x * q[0] + t[n + 3].x * q[1] + t[n + 6].x * q[2];
      w[n].y = t[n].y * q[0] + t[n + 3].y * q[1] + t[n + 6].y * q[2];
    }
#pragma unroll
    for (int l = 0; l < 3; l++) {
      u[l].x = -(Ai[l][0] * w[0].x + Ai[l][1] * w[1].x + Ai[l][2] * w[2].x);
      u[l].y = -(Ai[l][0] * w[0].y + Ai[l][1] * w[1].y + Ai[l][2] * w[2].y);
    }
  }
#pragma unroll
  for (int l = 0; l < 3; l++)
#pragma unroll
    for (int m = 0; m < 3; m++) t[l + 3 * m] = make_double2(u[l].x * q[m], u[l].y * q[m]);
  return div;
}

// general complex xi (forward-backward difference stencil)
static __device__ __noinline__ double gamma_point_cplx(double2* __restrict__ t, const double* __restrict__ C,
                                                const double2 xi[3], const double resc[3], bool is_dc,
                                                const double* __restrict__ dc_value) {
  double div = 0.0;
  double2 v[3];
  for (int m = 0; m < 3; m++) v[m] = cscale(cconj(make_double2(-xi[m].x, -xi[m].y)), resc[m]);
  for (int l = 0; l < 3; l++) {
    double2 p = cadd(cadd(cmul(t[l], v[0]), cmul(t[l + 3], v[1])), cmul(t[l + 6], v[2]));
    div += p.x * p.x + p.y * p.y;
  }
  if (is_dc) { for (int c = 0; c < 9; c++) t[c] = make_double2(dc_value[c], 0.0); return div; }
  double2 D[3][3], A[3][3], Ai[3][3], M[3][3];
  for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++) D[l][m] = cmul(cconj(make_double2(-xi[l].x, -xi[l].y)), xi[m]);
  for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++) {
    double2 acc = make_double2(0.0, 0.0);
    for (int n = 0; n < 3; n++) for (int o = 0; o < 3; o++) acc = cadd(acc, cscale(D[n][o], C[((l * 3 + n) * 3 + m) * 3 + o]));
    A[l][m] = acc;
  }
  double detre = A[0][0].x * (A[1][1].x * A[2][2].x - A[1][2].x * A[2][1].x)
               - A[0][1].x * (A[1][0].x * A[2][2].x - A[1][2].x * A[2][0].x)
               + A[0][2].x * (A[1][0].x * A[2][1].x - A[1][1].x * A[2][0].x);
  if (!(fabs(detre) > 1.0e-16)) { for (int c = 0; c < 9; c++) t[c] = make_double2(0.0, 0.0); return div; }
  double2 c00 = csub(cmul(A[1][1], A[2][2]), cmul(A[1][2], A[2][1]));
  double2 c01 = csub(cmul(A[1][2], A[2][0]), cmul(A[1][0], A[2][2]));
  double2 c02 = csub(cmul(A[1][0], A[2][1]), cmul(A[1][1], A[2][0]));
  double2 det = cadd(cadd(cmul(A[0][0], c00), cmul(A[0][1], c01)), cmul(A[0][2], c02));
  double dn = det.x * det.x + det.y * det.y;
  double2 id = make_double2(det.x / dn, -det.y / dn);
  Ai[0][0] = cmul(c00, id); Ai[1][0] = cmul(c01, id); Ai[2][0] = cmul(c02, id);
  Ai[0][1] = cmul(csub(cmul(A[0][2], A[2][1]), cmul(A[0][1], A[2][2])), id);
  Ai[1][1] = cmul(csub(cmul(A[0][0], A[2][2]), cmul(A[0][2], A[2][0])), id);
  Ai[2][1] = cmul(csub(cmul(A[0][1], A[2][0]), cmul(A[0][0], A[2][1])), id);
  Ai[0][2] = cmul(csub(cmul(A[0][1], A[1][2]), cmul(A[0][2], A[1][1])), id);
  Ai[1][2] = cmul(csub(cmul(A[0][2], A[1][0]), cmul(A[0][0], A[1][2])), id);
  Ai[2][2] = cmul(csub(cmul(A[0][0], A[1][1]), cmul(A[0][1], A[1][0])), id);
  for (int n = 0; n < 3; n++) for (int m = 0; m < 3; m++) {
    double2 acc = make_double2(0.0, 0.0);
    for (int o = 0; o < 3; o++) acc = cadd(acc, cmul(t[n + 3 * o], D[o][m]));
    M[n][m] = acc;
  }
  for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++)
    t[l + 3 * m] = cadd(cadd(cmul(Ai[l][0], M[0][m]), cmul(Ai[l][1], M[1][m])), cmul(Ai[l][2], M[2][m]));
  return div;
}

