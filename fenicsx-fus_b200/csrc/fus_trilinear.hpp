// fus_trilinear.hpp -- geometric factors of a trilinear hexahedron evaluated on the fly.
//
// The reference precomputes G = |det J| w_q K K^T per quadrature point
// (compute_scaled_geometrical_factor, precompute.hpp:101-213) and streams 48 B/point through the
// operator.  Every mesh the reference reads has a degree-1 coordinate element, so J is a trilinear
// function of 8 vertices: 192 B per CELL carry the same information.  These helpers are the whole
// arithmetic of that path; they are plain functions so that the CUDA kernel
// (stiffness_line_kernel<N,FUSE2,2>) and the host entry point fus_trilinear_geometry (CPU tests) run the
// same code.
//
// Cell map  x(xi) = sum_v X_v l_a(xi0) l_b(xi1) l_c(xi2),  v = a + 2b + 4c,  l_0(s) = 1-s, l_1(s) = s
// (DOLFINx tensor vertex order) written in monomials:
//   x = c000 + c100 xi0 + c010 xi1 + c001 xi2 + c110 xi0 xi1 + c101 xi0 xi2 + c011 xi1 xi2
//       + c111 xi0 xi1 xi2
// Stored per cell: FUS_TRI_STRIDE doubles = {c100, c010, c001, c110, c101, c011, c111} (3 each) +
// padding to a multiple of 16 bytes.  Columns of J (J_k = dx/dxi_k):
//   J0 = c100 + c110 xi1 + c101 xi2 + c111 xi1 xi2                  (independent of xi0)
//   J1 = (c010 + c011 xi2) + xi0 (c110 + c111 xi2) = A0 + xi0 dA
//   J2 = (c001 + c011 xi1) + xi0 (c101 + c111 xi1) = B0 + xi0 dB
// Rows of K = J^-1 are r_a / det with r_0 = J1 x J2, r_1 = J2 x J0, r_2 = J0 x J1, det = J0 . r_0,
// hence  |det| w (K K^T f)_a = (w / |det|) r_a . (f_0 r_0 + f_1 r_1 + f_2 r_2).
#pragma once

#if defined(__CUDACC__)
#define FUS_HD __host__ __device__ __forceinline__
#else
#define FUS_HD inline
#endif

#include "fus_b200.h" /* FUS_TRI_STRIDE: doubles per cell, 21 used, rows 16-byte aligned */

namespace fus {

// X[v][i]: coordinate i of vertex v (tensor vertex order) -> the 21 monomial coefficients.
// Differences are taken relative to vertex 0 first so that the result does not depend on where
// the cell sits in space beyond the rounding of the subtractions themselves.
FUS_HD void tri_cell_coeffs(const double X[8][3], double* c) {
  for (int i = 0; i < 3; ++i) {
    const double y1 = X[1][i] - X[0][i], y2 = X[2][i] - X[0][i], y3 = X[3][i] - X[0][i],
                 y4 = X[4][i] - X[0][i], y5 = X[5][i] - X[0][i], y6 = X[6][i] - X[0][i],
                 y7 = X[7][i] - X[0][i];
    c[0 + i] = y1;                                // c100
    c[3 + i] = y2;                                // c010
    c[6 + i] = y4;                                // c001
    c[9 + i] = y3 - y2 - y1;                      // c110
    c[12 + i] = y5 - y4 - y1;                     // c101
    c[15 + i] = y6 - y4 - y2;                     // c011
    c[18 + i] = ((y7 - y6) - (y5 - y4)) - (y3 - y2) + y1; // c111
  }
  c[21] = c[22] = c[23] = 0.0;
}

// Per-line pieces of J for the line (xi1, xi2): J0 and the two affine functions of xi0.
struct TriLine {
  double j0[3], a0[3], da[3], b0[3], db[3];
};

FUS_HD void tri_line_setup(const double* c, double xi1, double xi2, TriLine& L) {
  const double x12 = xi1 * xi2;
  for (int i = 0; i < 3; ++i) {
    L.j0[i] = c[0 + i] + c[9 + i] * xi1 + c[12 + i] * xi2 + c[18 + i] * x12;
    L.a0[i] = c[3 + i] + c[15 + i] * xi2;
    L.da[i] = c[9 + i] + c[18 + i] * xi2;
    L.b0[i] = c[6 + i] + c[15 + i] * xi1;
    L.db[i] = c[12 + i] + c[18 + i] * xi1;
  }
}

FUS_HD void tri_cross(const double* u, const double* v, double* r) {
  r[0] = u[1] * v[2] - u[2] * v[1];
  r[1] = u[2] * v[0] - u[0] * v[2];
  r[2] = u[0] * v[1] - u[1] * v[0];
}

// |det J| at xi0 on the line L: what compute_scaled_jacobian_determinant (precompute.hpp:33-94)
// stores per point, before the quadrature weight.
FUS_HD double tri_abs_det(const TriLine& L, double xi0) {
  double j1[3], j2[3], r0[3];
  for (int i = 0; i < 3; ++i) {
    j1[i] = L.a0[i] + xi0 * L.da[i];
    j2[i] = L.b0[i] + xi0 * L.db[i];
  }
  tri_cross(j1, j2, r0);
  const double det = L.j0[0] * r0[0] + L.j0[1] * r0[1] + L.j0[2] * r0[2];
  return det < 0.0 ? -det : det;
}

// (t0,t1,t2) = scale_w * |det J| * K K^T (f0,f1,f2) at xi0 on the line L, where scale_w carries the
// quadrature weight (and the cell coefficient): stiffness::transform (spectral_op.hpp:113-130)
// with G rebuilt instead of loaded.  Returns |det J|.
FUS_HD double tri_transform(const TriLine& L, double xi0, double scale_w, double f0, double f1,
                            double f2, double& t0, double& t1, double& t2) {
  double j1[3], j2[3], r0[3], r1[3], r2[3];
  for (int i = 0; i < 3; ++i) {
    j1[i] = L.a0[i] + xi0 * L.da[i];
    j2[i] = L.b0[i] + xi0 * L.db[i];
  }
  tri_cross(j1, j2, r0);
  tri_cross(j2, L.j0, r1);
  tri_cross(L.j0, j1, r2);
  const double det = L.j0[0] * r0[0] + L.j0[1] * r0[1] + L.j0[2] * r0[2];
  const double adet = det < 0.0 ? -det : det;
  const double s = scale_w / adet;
  const double w0 = f0 * r0[0] + f1 * r1[0] + f2 * r2[0];
  const double w1 = f0 * r0[1] + f1 * r1[1] + f2 * r2[1];
  const double w2 = f0 * r0[2] + f1 * r1[2] + f2 * r2[2];
  t0 = s * (r0[0] * w0 + r0[1] * w1 + r0[2] * w2);
  t1 = s * (r1[0] * w0 + r1[1] * w1 + r1[2] * w2);
  t2 = s * (r2[0] * w0 + r2[1] * w1 + r2[2] * w2);
  return adet;
}

} // namespace fus
