/*
 * fus_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the fenicsx-fus sum-factorised operator + RK4 path
 * (reference: cpp/fenicsx-sf/common/*.hpp).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * shipped CUDA path never calls it.
 *
 * Parity status: the tensor kernels (contract / transpose) are pinned against the
 * reference's own known-answer test (cpp/mwe/sum_factorisation/main.cpp:10-62) and
 * against the unmodified reference header compiled into oracle/_ref.  Everything that
 * the reference obtains from Basix / DOLFINx / FFCx at run time (GLL rule, 1-D
 * derivative table, Jacobians, facet integrals) has NO golden vector anywhere in the
 * reference tree: for those numbers this oracle is "parity unpinned" and is instead
 * self-validated by closed-form known answers and a dense O(N^6) evaluation in tests/, and
 * anchored outside this repository by the reference's own analytic tests
 * (python/tests/test_{linear,lossy,westervelt}spectral_1d.py: plane wave, attenuated plane
 * wave, Fubini solution -- all within the reference's thresholds), by h^(2P) convergence, and
 * by the reference's 2-D example run on its own shipped mesh and tags.
 *
 * Every function cites the reference lines it follows (paths relative to
 * cpp/fenicsx-sf/common/ unless stated).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* 1-D GLL rule on [0,1] in Basix order [0, 1, interior ascending]            */
/* (Basix make_quadrature(gll, ...) as called at spectral_op.hpp:57-59,160-162)*/
/* ------------------------------------------------------------------------- */

/* Legendre P_n(x) and P_{n-1}(x) by the three-term recurrence. */
static void legendre_pair(int n, double x, double* pn, double* pnm1) {
  double p0 = 1.0, p1 = x;
  if (n == 0) {
    *pn = 1.0;
    *pnm1 = 0.0;
    return;
  }
  for (int k = 2; k <= n; ++k) {
    double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / (double)k;
    p0 = p1;
    p1 = p2;
  }
  *pn = p1;
  *pnm1 = p0;
}

/* m points.  Returns 0 on success. */
int fo_gll(int m, double* pts, double* wts) {
  if (m < 2 || m > 64)
    return 1;
  int n = m - 1; /* polynomial degree of the rule's Legendre polynomial */
  double* x = (double*)malloc(sizeof(double) * m);
  double* w = (double*)malloc(sizeof(double) * m);
  for (int j = 0; j < m; ++j) {
    /* Chebyshev-Gauss-Lobatto start, then the classical fixed-point/Newton map
       x <- x - (x P_n - P_{n-1}) / ((n+1) P_n) whose fixed points are the LGL nodes */
    double xj = -cos(M_PI * j / n);
    for (int it = 0; it < 200; ++it) {
      double pn, pnm1;
      legendre_pair(n, xj, &pn, &pnm1);
      double dx = (xj * pn - pnm1) / ((n + 1.0) * pn);
      xj -= dx;
      if (fabs(dx) < 1e-17)
        break;
    }
    if (j == 0)
      xj = -1.0;
    if (j == n)
      xj = 1.0;
    double pn, pnm1;
    legendre_pair(n, xj, &pn, &pnm1);
    x[j] = xj;
    w[j] = 2.0 / (n * (n + 1.0) * pn * pn);
  }
  /* symmetrise (kills the last-ulp asymmetry of the iteration) */
  for (int j = 0; j < m / 2; ++j) {
    double a = 0.5 * (x[m - 1 - j] - x[j]);
    x[j] = -a;
    x[m - 1 - j] = a;
    double ww = 0.5 * (w[j] + w[m - 1 - j]);
    w[j] = ww;
    w[m - 1 - j] = ww;
  }
  if (m % 2 == 1)
    x[m / 2] = 0.0;
  /* map to [0,1] and rotate to [first, last, interior...] */
  pts[0] = 0.0;
  wts[0] = 0.5 * w[0];
  pts[1] = 1.0;
  wts[1] = 0.5 * w[m - 1];
  for (int j = 1; j < m - 1; ++j) {
    pts[j + 1] = 0.5 + 0.5 * x[j];
    wts[j + 1] = 0.5 * w[j];
  }
  free(x);
  free(w);
  return 0;
}

/* dphi[q*N+i] = phi_i'(pts[q]) for the Lagrange basis on pts
   (tabulate_1d, precompute.hpp:217-234, derivative block spectral_op.hpp:168-170) */
void fo_dphi(int N, const double* pts, double* dphi) {
  double lam[64];
  for (int i = 0; i < N; ++i) {
    double p = 1.0;
    for (int j = 0; j < N; ++j)
      if (j != i)
        p *= (pts[i] - pts[j]);
    lam[i] = 1.0 / p;
  }
  for (int q = 0; q < N; ++q) {
    double diag = 0.0;
    for (int i = 0; i < N; ++i) {
      if (i == q)
        continue;
      double d = (lam[i] / lam[q]) / (pts[q] - pts[i]);
      dphi[q * N + i] = d;
      diag -= d;
    }
    dphi[q * N + q] = diag;
  }
}

/* ------------------------------------------------------------------------- */
/* Tensor kernels (sum_factorisation.hpp:43-49 and :70-86), runtime sizes,    */
/* same loop order and the same accumulate-into-C semantics.                  */
/* ------------------------------------------------------------------------- */
void fo_contract(int Nk, int Na, int Nb, int Nc, int transpose, const double* A, const double* B,
                 double* C) {
  int Nd = Nb * Nc;
  if (transpose) {
    for (int k = 0; k < Nk; k++)
      for (int a = 0; a < Na; a++)
        for (int d = 0; d < Nd; d++)
          C[a * Nd + d] += A[a * Nk + k] * B[k * Nd + d];
  } else {
    for (int k = 0; k < Nk; k++)
      for (int a = 0; a < Na; a++)
        for (int d = 0; d < Nd; d++)
          C[a * Nd + d] += A[k * Na + a] * B[k * Nd + d];
  }
}

void fo_transpose(int Na, int Nb, int Nc, int offa, int offb, int offc, const double* A,
                  double* B) {
  for (int a = 0; a < Na; a++)
    for (int b = 0; b < Nb; b++)
      for (int c = 0; c < Nc; c++)
        B[offa * a + offb * b + offc * c] = A[a * Nb * Nc + b * Nc + c];
}

/* ------------------------------------------------------------------------- */
/* Structured hexahedral box: geometry, cell->vertex map, tensor dofmap       */
/* (stands in for dolfinx::mesh::create_box + create_functionspace +          */
/*  reorder_dofmap, permute.hpp:15-42)                                        */
/* ------------------------------------------------------------------------- */

/* vertices: id = (vx*(ny+1)+vy)*(nz+1)+vz ; cell c=(cx*ny+cy)*nz+cz ;
   local vertex v = a + 2b + 4c  <->  (cx+a, cy+b, cz+c)  (x fastest, DOLFINx hex order) */
void fo_box_mesh(int nx, int ny, int nz, const double* lo, const double* hi, double* xg,
                 int32_t* xdofmap) {
  for (int vx = 0; vx <= nx; ++vx)
    for (int vy = 0; vy <= ny; ++vy)
      for (int vz = 0; vz <= nz; ++vz) {
        size_t id = ((size_t)vx * (ny + 1) + vy) * (nz + 1) + vz;
        xg[3 * id + 0] = lo[0] + (hi[0] - lo[0]) * vx / nx;
        xg[3 * id + 1] = lo[1] + (hi[1] - lo[1]) * vy / ny;
        xg[3 * id + 2] = lo[2] + (hi[2] - lo[2]) * vz / nz;
      }
  for (int cx = 0; cx < nx; ++cx)
    for (int cy = 0; cy < ny; ++cy)
      for (int cz = 0; cz < nz; ++cz) {
        size_t c = ((size_t)cx * ny + cy) * nz + cz;
        for (int v = 0; v < 8; ++v) {
          int a = v & 1, b = (v >> 1) & 1, cc = (v >> 2) & 1;
          xdofmap[8 * c + v] = (int32_t)(((size_t)(cx + a) * (ny + 1) + (cy + b)) * (nz + 1) + (cz + cc));
        }
      }
}

/* position (0..P) along an edge of 1-D node i in Basix order [0,1,interior] */
static inline int node_pos(int i, int P) { return i == 0 ? 0 : (i == 1 ? P : i - 1); }

/* Global dof number of grid node (gx,gy,gz).
   mode 0: lexicographic, x slowest.
   mode 1: cell-blocked: node belongs to the cell (min(gx/P,nx-1),...) ... the lower-corner
           owner; dofs of one owner cell are contiguous, owners in cell order. */
static int64_t box_dof(int P, int nx, int ny, int nz, int mode, int gx, int gy, int gz,
                       const int64_t* blk_off) {
  int Mx = nx * P + 1, My = ny * P + 1, Mz = nz * P + 1;
  (void)Mx;
  if (mode == 0)
    return ((int64_t)gx * My + gy) * Mz + gz;
  /* owner cell and offset inside it; the last layer of nodes joins the last cell */
  int ox = gx / P, oy = gy / P, oz = gz / P;
  if (ox == nx) ox = nx - 1;
  if (oy == ny) oy = ny - 1;
  if (oz == nz) oz = nz - 1;
  int lx = gx - ox * P, ly = gy - oy * P, lz = gz - oz * P;
  int sy = (oy == ny - 1) ? P + 1 : P, sz = (oz == nz - 1) ? P + 1 : P;
  int64_t oc = ((int64_t)ox * ny + oy) * nz + oz;
  return blk_off[oc] + ((int64_t)lx * sy + ly) * sz + lz;
}

/* tensor_dofmap[c*Nd + i0*N*N + i1*N + i2] (a1 in SURVEY section 8) */
int fo_box_dofmap(int P, int nx, int ny, int nz, int mode, int32_t* dofmap) {
  int N = P + 1;
  int64_t nc = (int64_t)nx * ny * nz;
  int64_t* blk_off = NULL;
  if (mode == 1) {
    blk_off = (int64_t*)malloc(sizeof(int64_t) * (nc + 1));
    int64_t acc = 0;
    for (int cx = 0; cx < nx; ++cx)
      for (int cy = 0; cy < ny; ++cy)
        for (int cz = 0; cz < nz; ++cz) {
          int sx = (cx == nx - 1) ? P + 1 : P, sy = (cy == ny - 1) ? P + 1 : P,
              sz = (cz == nz - 1) ? P + 1 : P;
          blk_off[((int64_t)cx * ny + cy) * nz + cz] = acc;
          acc += (int64_t)sx * sy * sz;
        }
    blk_off[nc] = acc;
  }
  for (int cx = 0; cx < nx; ++cx)
    for (int cy = 0; cy < ny; ++cy)
      for (int cz = 0; cz < nz; ++cz) {
        int64_t c = ((int64_t)cx * ny + cy) * nz + cz;
        for (int i0 = 0; i0 < N; ++i0)
          for (int i1 = 0; i1 < N; ++i1)
            for (int i2 = 0; i2 < N; ++i2) {
              int gx = cx * P + node_pos(i0, P), gy = cy * P + node_pos(i1, P),
                  gz = cz * P + node_pos(i2, P);
              dofmap[c * N * N * N + (i0 * N + i1) * N + i2]
                  = (int32_t)box_dof(P, nx, ny, nz, mode, gx, gy, gz, blk_off);
            }
      }
  free(blk_off);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Geometry (precompute.hpp:33-94 detJ, :101-213 G)                           */
/* ------------------------------------------------------------------------- */

/* Jacobian of the trilinear map at reference point xi; J[i][j] = d x_i / d xi_j
   (CoordinateElement::compute_jacobian as called at precompute.hpp:83,183) */
static void q1_jacobian(const double X[8][3], const double xi[3], double J[3][3]) {
  memset(J, 0, sizeof(double) * 9);
  for (int v = 0; v < 8; ++v) {
    int a = v & 1, b = (v >> 1) & 1, c = (v >> 2) & 1;
    double l0 = a ? xi[0] : 1.0 - xi[0], l1 = b ? xi[1] : 1.0 - xi[1],
           l2 = c ? xi[2] : 1.0 - xi[2];
    double d0 = a ? 1.0 : -1.0, d1 = b ? 1.0 : -1.0, d2 = c ? 1.0 : -1.0;
    double g[3] = {d0 * l1 * l2, l0 * d1 * l2, l0 * l1 * d2};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        J[i][j] += X[v][i] * g[j];
  }
}

static double det3(const double J[3][3]) {
  return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
         - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
         + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

static void inv3(const double J[3][3], double K[3][3]) {
  double d = det3(J);
  K[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / d;
  K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / d;
  K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / d;
  K[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / d;
  K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / d;
  K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / d;
  K[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / d;
  K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / d;
  K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / d;
}

/* G[c][q][6] = |detJ| w_q {G00,G01,G02,G11,G12,G22}, G = K K^T  (precompute.hpp:160-209)
   detJ[c][q] = |detJ| w_q                                          (precompute.hpp:67-91)
   q = q0*N*N + q1*N + q2 <-> (pts[q0], pts[q1], pts[q2]); either output may be NULL. */
void fo_geometry(int64_t nc, const double* xg, const int32_t* xdofmap, int N, const double* pts,
                 const double* wts, double* G, double* detJ) {
  int Nd = N * N * N;
  for (int64_t c = 0; c < nc; ++c) {
    double X[8][3];
    for (int v = 0; v < 8; ++v)
      for (int j = 0; j < 3; ++j)
        X[v][j] = xg[3 * (size_t)xdofmap[8 * c + v] + j];
    for (int q0 = 0; q0 < N; ++q0)
      for (int q1 = 0; q1 < N; ++q1)
        for (int q2 = 0; q2 < N; ++q2) {
          int q = (q0 * N + q1) * N + q2;
          double xi[3] = {pts[q0], pts[q1], pts[q2]};
          double w = wts[q0] * wts[q1] * wts[q2];
          double J[3][3], K[3][3];
          q1_jacobian(X, xi, J);
          double dj = fabs(det3(J)) * w;
          if (detJ)
            detJ[c * Nd + q] = dj;
          if (G) {
            inv3(J, K);
            double g[3][3];
            for (int a = 0; a < 3; ++a)
              for (int b = 0; b < 3; ++b) {
                double s = 0.0;
                for (int i = 0; i < 3; ++i)
                  s += K[a][i] * K[b][i];
                g[a][b] = s;
              }
            double* o = G + ((size_t)c * Nd + q) * 6;
            o[0] = dj * g[0][0];
            o[1] = dj * g[0][1];
            o[2] = dj * g[0][2];
            o[3] = dj * g[1][1];
            o[4] = dj * g[1][2];
            o[5] = dj * g[2][2];
          }
        }
  }
}

/* ------------------------------------------------------------------------- */
/* Operators (spectral_op.hpp:69-86 mass, :173-243 stiffness); y += A x       */
/* ------------------------------------------------------------------------- */
void fo_mass_apply(int P, int64_t nc, const int32_t* dofmap, const double* detJ,
                   const double* coeffs, const double* x, double* y) {
  int N = P + 1, Nd = N * N * N;
  double* x_ = (double*)malloc(sizeof(double) * Nd);
  for (int64_t c = 0; c < nc; ++c) {
    for (int i = 0; i < Nd; ++i)
      x_[i] = x[dofmap[c * Nd + i]];
    const double* sdetJ = detJ + c * Nd;
    for (int iq = 0; iq < Nd; ++iq) /* mass::transform, spectral_op.hpp:19-26 */
      x_[iq] = coeffs[c] * x_[iq] * sdetJ[iq];
    for (int i = 0; i < Nd; ++i)
      y[dofmap[c * Nd + i]] += x_[i];
  }
  free(x_);
}

void fo_stiffness_apply(int P, int64_t nc, const int32_t* dofmap, const double* G,
                        const double* dphi, const double* coeffs, const double* x, double* y) {
  int N = P + 1, Nd = N * N * N;
  size_t bytes = sizeof(double) * Nd;
  double* buf = (double*)malloc(bytes * 11);
  double *x_ = buf, *fw0 = buf + Nd, *fw1 = buf + 2 * Nd, *fw2 = buf + 3 * Nd, *y0 = buf + 4 * Nd,
         *y1 = buf + 5 * Nd, *y2 = buf + 6 * Nd, *T1 = buf + 7 * Nd, *T2 = buf + 8 * Nd,
         *T3 = buf + 9 * Nd, *T4 = buf + 10 * Nd;
  for (int64_t c = 0; c < nc; ++c) {
    for (int i = 0; i < Nd; ++i)
      x_[i] = x[dofmap[c * Nd + i]];
    memset(T1, 0, 4 * bytes);
    /* forward contractions, spectral_op.hpp:193-210 */
    memset(fw0, 0, bytes);
    fo_contract(N, N, N, N, 1, dphi, x_, fw0);
    memset(fw1, 0, bytes);
    fo_transpose(N, N, N, N, N * N, 1, x_, T1);
    fo_contract(N, N, N, N, 1, dphi, T1, T2);
    fo_transpose(N, N, N, N, N * N, 1, T2, fw1);
    memset(fw2, 0, bytes);
    fo_transpose(N, N, N, 1, N, N * N, x_, T3);
    fo_contract(N, N, N, N, 1, dphi, T3, T4);
    fo_transpose(N, N, N, 1, N, N * N, T4, fw2);
    /* stiffness::transform, spectral_op.hpp:113-130 */
    const double* Gc = G + (size_t)c * Nd * 6;
    double coeff = coeffs[c];
    for (int iq = 0; iq < Nd; ++iq) {
      const double* _G = Gc + iq * 6;
      double w0 = fw0[iq], w1 = fw1[iq], w2 = fw2[iq];
      fw0[iq] = coeff * (_G[0] * w0 + _G[1] * w1 + _G[2] * w2);
      fw1[iq] = coeff * (_G[1] * w0 + _G[3] * w1 + _G[4] * w2);
      fw2[iq] = coeff * (_G[2] * w0 + _G[4] * w1 + _G[5] * w2);
    }
    memset(T1, 0, 4 * bytes);
    /* transposed contractions, spectral_op.hpp:221-238 */
    memset(y0, 0, bytes);
    fo_contract(N, N, N, N, 0, dphi, fw0, y0);
    memset(y1, 0, bytes);
    fo_transpose(N, N, N, N, N * N, 1, fw1, T1);
    fo_contract(N, N, N, N, 0, dphi, T1, T2);
    fo_transpose(N, N, N, N, N * N, 1, T2, y1);
    memset(y2, 0, bytes);
    fo_transpose(N, N, N, 1, N, N * N, fw2, T3);
    fo_contract(N, N, N, N, 0, dphi, T3, T4);
    fo_transpose(N, N, N, 1, N, N * N, T4, y2);
    for (int i = 0; i < Nd; ++i)
      y[dofmap[c * Nd + i]] += y0[i] + y1[i] + y2[i];
  }
  free(buf);
}

/* ------------------------------------------------------------------------- */
/* Exterior-facet data.  The FFCx `ds` kernels with GLL quadrature are        */
/* collocated, so one facet contributes coef_c * w2d_i * |J_f|(i) * value_i    */
/* to each of its N*N nodes (forms: fenicsx-sf-naive/benchmarks/PH1/SC2-BM1/   */
/* forms.py:35-38, fenicsx-sf/benchmarks/PH1/BM7-SC1/forms.py:37-42).          */
/* Local facet ids follow the DOLFINx hexahedron: 0:z=0 1:y=0 2:x=0 3:x=1      */
/* 4:y=1 5:z=1 with reference axes (xi0,xi1,xi2) = (x,y,z).                    */
/* ------------------------------------------------------------------------- */
static void facet_axes(int lf, int* dir, int* side) {
  static const int d[6] = {2, 1, 0, 0, 1, 2};
  static const int s[6] = {0, 0, 0, 1, 1, 1};
  *dir = d[lf];
  *side = s[lf];
}

/* For facet (cell, lf): local tensor indices of its N*N nodes (fnodes) and the
   scaled surface weight w2d*|J_f| at each (fscale). */
void fo_facet_data(int N, const double* xg, const int32_t* xdofmap, const double* pts,
                   const double* wts, int64_t cell, int lf, int32_t* fnodes, double* fscale) {
  int dir, side;
  facet_axes(lf, &dir, &side);
  int ta = (dir == 0) ? 1 : 0, tb = (dir == 2) ? 1 : 2; /* tangential axes, ascending */
  double X[8][3];
  for (int v = 0; v < 8; ++v)
    for (int j = 0; j < 3; ++j)
      X[v][j] = xg[3 * (size_t)xdofmap[8 * cell + v] + j];
  for (int a = 0; a < N; ++a)
    for (int b = 0; b < N; ++b) {
      int idx[3];
      idx[dir] = side; /* Basix order: node 0 at xi=0, node 1 at xi=1 */
      idx[ta] = a;
      idx[tb] = b;
      double xi[3] = {pts[idx[0]], pts[idx[1]], pts[idx[2]]};
      double J[3][3];
      q1_jacobian(X, xi, J);
      double t1[3] = {J[0][ta], J[1][ta], J[2][ta]}, t2[3] = {J[0][tb], J[1][tb], J[2][tb]};
      double n[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2],
                     t1[0] * t2[1] - t1[1] * t2[0]};
      fnodes[a * N + b] = (idx[0] * N + idx[1]) * N + idx[2];
      fscale[a * N + b] = wts[a] * wts[b] * sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    }
}

/* Exterior facets of a box: returns count; facets[3*k] = {cell, local facet, tag}.
   tag 1 on x=lo face, tag 2 on x=hi face, 0 elsewhere.
   Pass facets=NULL to count only. */
int64_t fo_box_facets(int nx, int ny, int nz, int32_t* facets) {
  int64_t k = 0;
  for (int cx = 0; cx < nx; ++cx)
    for (int cy = 0; cy < ny; ++cy)
      for (int cz = 0; cz < nz; ++cz) {
        int32_t c = (int32_t)(((int64_t)cx * ny + cy) * nz + cz);
        int on[6] = {cz == 0, cy == 0, cx == 0, cx == nx - 1, cy == ny - 1, cz == nz - 1};
        for (int lf = 0; lf < 6; ++lf)
          if (on[lf]) {
            if (facets) {
              facets[3 * k] = c;
              facets[3 * k + 1] = lf;
              facets[3 * k + 2] = (lf == 2) ? 1 : ((lf == 3) ? 2 : 0);
            }
            ++k;
          }
      }
  return k;
}

/* ------------------------------------------------------------------------- */
/* Models and RK4: literal restatement of the reference flow                  */
/*   Linear.hpp:127-134 (m), :171-222 (f0,f1), :228-314 (rk4)                 */
/*   Lossy.hpp:133-141, :196-251 ; Westervelt.hpp:137-144, :216-281           */
/* kind: 0 linear, 1 lossy, 2 westervelt.                                     */
/* Single address space: `nowned` entries are owned, the rest of ndofs are    */
/* ghosts (none on one rank) -- axpy touches owned entries only.              */
/* ------------------------------------------------------------------------- */
typedef struct {
  int kind, P, N, Nd;
  int dim, nfn; /* 3 (hexahedra) or 2 (quadrilaterals); nodes per facet = N^(dim-1) */
  int64_t nc, ndofs, nowned;
  const int32_t* dofmap;
  const double *G, *detJ, *dphi;
  const double *c0, *rho0, *delta0, *beta0;
  int64_t nfacets;
  const int32_t* facets;  /* {cell, lf, tag} */
  const int32_t* fnodes;  /* [nfacets][N*N] local tensor node index */
  const double* fscale;   /* [nfacets][N*N] */
  double freq, p0, s0, w0, period, window_length;
  double src_factor;
  /* operator hooks: default to the plain-C restatement below; oracle/_ref swaps in the
     cell loop built on the reference's own sum_factorisation.hpp */
  void (*stiff)(int, int64_t, const int32_t*, const double*, const double*, const double*,
                const double*, double*);
  void (*mass)(int, int64_t, const int32_t*, const double*, const double*, const double*, double*);
  /* derived */
  double *lin_c, *att_c, *nl1_c, *nl2_c;
  double *m, *m0, *b, *g, *dg, *u_n, *v_n, *w_n;
} fo_model;

/* Vector helpers.  The pragmas only take effect when built with -fopenmp (the timed CPU
   baseline); they do not change any per-entry arithmetic. */
static void vcopy(double* dst, const double* src, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    dst[i] = src[i];
}
static void vzero(double* dst, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    dst[i] = 0.0;
}

static double* vec(int64_t n) { return (double*)calloc((size_t)n, sizeof(double)); }

/* facet part of the bilinear form a (mass-like, applied to u==1) and of L.
   Which facets carry the absorbing term and its mass-like counterpart differs between the two
   variants of the reference: the 3-D forms of cpp/fenicsx-sf integrate them with `ds` WITHOUT an id,
   i.e. over every exterior facet (benchmarks/PH1/BM7-SC1/forms.py:37-42), the 2-D forms of
   cpp/fenicsx-sf-naive over ds(2) only (examples/lossy_planewave2d_1/forms.py:37-42,
   westervelt_planewave2d_1/forms.py:37-42). */
static int absorbing_facet(const fo_model* M, int tag) { return M->dim == 3 || tag == 2; }

static void assemble_facets_a(const fo_model* M, double* out) {
  int NN = M->nfn;
  for (int64_t f = 0; f < M->nfacets; ++f) {
    int32_t c = M->facets[3 * f];
    if (!absorbing_facet(M, M->facets[3 * f + 2]))
      continue;
    double coef = M->delta0[c] / M->rho0[c] / M->c0[c] / M->c0[c] / M->c0[c];
    for (int k = 0; k < NN; ++k) {
      int32_t d = M->dofmap[(int64_t)c * M->Nd + M->fnodes[f * NN + k]];
      out[d] += coef * 1.0 * M->fscale[f * NN + k];
    }
  }
}

static void assemble_facets_L(const fo_model* M, double* b) {
  int NN = M->nfn;
  for (int64_t f = 0; f < M->nfacets; ++f) {
    int32_t c = M->facets[3 * f];
    int tag = M->facets[3 * f + 2];
    double rho = M->rho0[c], cc = M->c0[c];
    for (int k = 0; k < NN; ++k) {
      int32_t d = M->dofmap[(int64_t)c * M->Nd + M->fnodes[f * NN + k]];
      double s = M->fscale[f * NN + k];
      if (tag == 1)
        b[d] += 1.0 / rho * M->g[d] * s;
      if (M->kind == 0) {
        if (tag == 2)
          b[d] -= 1.0 / rho / cc * M->v_n[d] * s;
      } else {
        if (absorbing_facet(M, tag)) /* 3-D: ds without id, every exterior facet; 2-D: ds(2) */
          b[d] -= 1.0 / rho / cc * M->v_n[d] * s;
        if (tag == 1)
          b[d] += M->delta0[c] / rho / cc / cc * M->dg[d] * s;
      }
    }
  }
}

void fo_mass_apply_2d(int P, int64_t nc, const int32_t* dofmap, const double* detJ,
                      const double* coeffs, const double* x, double* y);
void fo_stiffness_apply_2d(int P, int64_t nc, const int32_t* dofmap, const double* G,
                           const double* dphi, const double* coeffs, const double* x, double* y);

static fo_model* model_create(int dim, int kind, int P, int64_t nc, int64_t ndofs, int64_t nowned,
                              const int32_t* dofmap, const double* G, const double* detJ,
                              const double* dphi, const double* c0, const double* rho0,
                              const double* delta0, const double* beta0, int64_t nfacets,
                              const int32_t* facets, const int32_t* fnodes, const double* fscale,
                              double freq, double p0, double s0) {
  fo_model* M = (fo_model*)calloc(1, sizeof(fo_model));
  M->kind = kind;
  M->P = P;
  M->N = P + 1;
  M->dim = dim;
  M->Nd = (dim == 3) ? M->N * M->N * M->N : M->N * M->N;
  M->nfn = (dim == 3) ? M->N * M->N : M->N;
  M->nc = nc;
  M->ndofs = ndofs;
  M->nowned = nowned;
  M->dofmap = dofmap;
  M->G = G;
  M->detJ = detJ;
  M->dphi = dphi;
  M->c0 = c0;
  M->rho0 = rho0;
  M->delta0 = delta0;
  M->beta0 = beta0;
  M->nfacets = nfacets;
  M->facets = facets;
  M->fnodes = fnodes;
  M->fscale = fscale;
  M->freq = freq;
  M->w0 = 2 * M_PI * freq;
  M->p0 = p0;
  M->s0 = s0;
  M->period = 1.0 / freq;
  M->window_length = 4.0;
  M->stiff = (dim == 3) ? fo_stiffness_apply : fo_stiffness_apply_2d;
  M->mass = (dim == 3) ? fo_mass_apply : fo_mass_apply_2d;
  M->src_factor = (kind == 0) ? 1.0 : 2.0; /* Linear.hpp:192 vs Lossy.hpp:216, Westervelt.hpp:237 */
  M->lin_c = vec(nc);
  M->att_c = vec(nc);
  M->nl1_c = vec(nc);
  M->nl2_c = vec(nc);
  double* mcoef = vec(nc);
  for (int64_t i = 0; i < nc; ++i) {
    M->lin_c[i] = -1.0 / rho0[i];
    if (kind >= 1)
      M->att_c[i] = -delta0[i] / rho0[i] / c0[i] / c0[i];
    if (kind == 2) {
      M->nl1_c[i] = -2.0 * beta0[i] / rho0[i] / rho0[i] / c0[i] / c0[i] / c0[i] / c0[i];
      M->nl2_c[i] = 2.0 * beta0[i] / rho0[i] / rho0[i] / c0[i] / c0[i] / c0[i] / c0[i];
    }
    mcoef[i] = 1.0 / rho0[i] / c0[i] / c0[i];
  }
  M->m = vec(ndofs);
  M->m0 = vec(ndofs);
  M->b = vec(ndofs);
  M->g = vec(ndofs);
  M->dg = vec(ndofs);
  M->u_n = vec(ndofs);
  M->v_n = vec(ndofs);
  M->w_n = vec(ndofs);
  /* lumped mass: assemble a with u == 1 (Linear.hpp:127-134) */
  double* ones = vec(ndofs);
  for (int64_t i = 0; i < ndofs; ++i)
    ones[i] = 1.0;
  double* tgt = (kind == 2) ? M->m0 : M->m;
  M->mass(P, nc, dofmap, detJ, mcoef, ones, tgt);
  if (kind >= 1)
    assemble_facets_a(M, tgt);
  free(ones);
  free(mcoef);
  return M;
}

fo_model* fo_model_create(int kind, int P, int64_t nc, int64_t ndofs, int64_t nowned,
                          const int32_t* dofmap, const double* G, const double* detJ,
                          const double* dphi, const double* c0, const double* rho0,
                          const double* delta0, const double* beta0, int64_t nfacets,
                          const int32_t* facets, const int32_t* fnodes, const double* fscale,
                          double freq, double p0, double s0) {
  return model_create(3, kind, P, nc, ndofs, nowned, dofmap, G, detJ, dphi, c0, rho0, delta0,
                      beta0, nfacets, facets, fnodes, fscale, freq, p0, s0);
}

/* {Linear,Lossy,Westervelt}Spectral2D (cpp/fenicsx-sf-naive/common/Linear.hpp:52-350 and the
   2-D classes of Lossy.hpp / Westervelt.hpp there): the same flow on quadrilaterals; G has 3
   entries per point, a facet is an edge with N nodes. */
fo_model* fo_model_create_2d(int kind, int P, int64_t nc, int64_t ndofs, int64_t nowned,
                             const int32_t* dofmap, const double* G, const double* detJ,
                             const double* dphi, const double* c0, const double* rho0,
                             const double* delta0, const double* beta0, int64_t nfacets,
                             const int32_t* facets, const int32_t* fnodes, const double* fscale,
                             double freq, double p0, double s0) {
  return model_create(2, kind, P, nc, ndofs, nowned, dofmap, G, detJ, dphi, c0, rho0, delta0,
                      beta0, nfacets, facets, fnodes, fscale, freq, p0, s0);
}

void fo_model_destroy(fo_model* M) {
  free(M->lin_c); free(M->att_c); free(M->nl1_c); free(M->nl2_c);
  free(M->m); free(M->m0); free(M->b); free(M->g); free(M->dg);
  free(M->u_n); free(M->v_n); free(M->w_n);
  free(M);
}

const double* fo_model_mass(const fo_model* M) { return (M->kind == 2) ? M->m0 : M->m; }

typedef void (*fo_stiff_fn)(int, int64_t, const int32_t*, const double*, const double*,
                            const double*, const double*, double*);
typedef void (*fo_mass_fn)(int, int64_t, const int32_t*, const double*, const double*,
                           const double*, double*);
void fo_model_set_ops(fo_model* M, fo_stiff_fn s, fo_mass_fn m) {
  if (s) M->stiff = s;
  if (m) M->mass = m;
}

/* f1: Linear.hpp:181-222, Lossy.hpp:196-251, Westervelt.hpp:216-281 */
static void f1(fo_model* M, double t, const double* u, const double* v, double* result) {
  int64_t n = M->ndofs;
  double window, dwindow;
  if (t < M->period * M->window_length) {
    window = 0.5 * (1.0 - cos(M->freq * M_PI * t / M->window_length));
    dwindow = 0.5 * M_PI * M->freq / M->window_length * sin(M->freq * M_PI * t / M->window_length);
  } else {
    window = 1.0;
    dwindow = 0.0;
  }
  double gval, dgval = 0.0;
  if (M->kind == 0) {
    gval = window * M->p0 * M->w0 / M->s0 * cos(M->w0 * t);
  } else {
    gval = window * 2.0 * M->p0 * M->w0 / M->s0 * cos(M->w0 * t);
    dgval = dwindow * 2.0 * M->p0 * M->w0 / M->s0 * cos(M->w0 * t)
            - window * 2.0 * M->p0 * M->w0 * M->w0 / M->s0 * sin(M->w0 * t);
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    M->g[i] = gval;
    M->dg[i] = dgval;
  }
  /* scatter_fwd is the identity on one rank */
  vcopy(M->u_n, u, n);
  vcopy(M->v_n, v, n);
  if (M->kind == 2) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
      M->w_n[i] = M->v_n[i] * M->v_n[i];
    vzero(M->m, n);
    M->mass(M->P, M->nc, M->dofmap, M->detJ, M->nl1_c, M->u_n, M->m);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
      M->m[i] = M->m0[i] + M->m[i];
  }
  vzero(M->b, n);
  M->stiff(M->P, M->nc, M->dofmap, M->G, M->dphi, M->lin_c, M->u_n, M->b);
  if (M->kind >= 1)
    M->stiff(M->P, M->nc, M->dofmap, M->G, M->dphi, M->att_c, M->v_n, M->b);
  if (M->kind == 2)
    M->mass(M->P, M->nc, M->dofmap, M->detJ, M->nl2_c, M->w_n, M->b);
  assemble_facets_L(M, M->b);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    result[i] = M->b[i] / M->m[i];
}

/* one stage-level evaluation, exposed for tests: kv = f1(t,u,v) */
void fo_model_f1(fo_model* M, double t, const double* u, const double* v, double* result) {
  f1(M, t, u, v, result);
}

static void axpy(int64_t nowned, double* r, double alpha, const double* x, const double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nowned; ++i)
    r[i] = x[i] * alpha + y[i];
}

/* rk4: Linear.hpp:228-314.  u,v hold u_n,v_n on entry and the solution on exit.
   Returns the number of steps taken. */
int fo_model_rk4(fo_model* M, double startTime, double finalTime, double timeStep, double* u,
                 double* v) {
  int64_t n = M->ndofs;
  double t = startTime, tf = finalTime, dt = timeStep;
  int step = 0;
  double *u_ = vec(n), *v_ = vec(n), *un = vec(n), *vn = vec(n), *u0 = vec(n), *v0 = vec(n),
         *ku = vec(n), *kv = vec(n);
  vcopy(u_, u, n);
  vcopy(v_, v, n);
  vcopy(ku, u_, n);
  vcopy(kv, v_, n);
  const double a_runge[4] = {0.0, 0.5, 0.5, 1.0};
  const double b_runge[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
  const double c_runge[4] = {0.0, 0.5, 0.5, 1.0};
  while (t < tf) {
    dt = fmin(dt, tf - t);
    vcopy(u0, u_, n);
    vcopy(v0, v_, n);
    for (int i = 0; i < 4; i++) {
      vcopy(un, u0, n);
      vcopy(vn, v0, n);
      axpy(M->nowned, un, dt * a_runge[i], ku, un);
      axpy(M->nowned, vn, dt * a_runge[i], kv, vn);
      double tn = t + c_runge[i] * dt;
      vcopy(ku, vn, n); /* f0 */
      f1(M, tn, un, vn, kv);
      axpy(M->nowned, u_, dt * b_runge[i], ku, u_);
      axpy(M->nowned, v_, dt * b_runge[i], kv, v_);
    }
    t += dt;
    step += 1;
  }
  vcopy(u, u_, n);
  vcopy(v, v_, n);
  free(u_); free(v_); free(un); free(vn); free(u0); free(v0); free(ku); free(kv);
  return step;
}

/* ========================================================================= */
/* 2-D quadrilateral variant.  Reference:            */
/* cpp/fenicsx-sf-naive/common/{sum_factorisation,spectral_op,precompute}.hpp */
/* ========================================================================= */

/* transpose<T,Na,Nb,offa,offb> (fenicsx-sf-naive sum_factorisation.hpp:11-19) */
void fo_transpose_2d(int Na, int Nb, int offa, int offb, const double* A, double* B) {
  for (int a = 0; a < Na; ++a)
    for (int b = 0; b < Nb; ++b)
      B[a * offa + b * offb] = A[a * Nb + b];
}

/* contract<T,Na,Nb,Nk>: C[a,b] += A[a,k] * B[b,k] (fenicsx-sf-naive sum_factorisation.hpp:29-39) */
void fo_contract_2d(int Na, int Nb, int Nk, const double* A, const double* B, double* C) {
  for (int a = 0; a < Na; ++a)
    for (int b = 0; b < Nb; ++b)
      for (int k = 0; k < Nk; ++k)
        C[a * Nb + b] += A[a * Nk + k] * B[b * Nk + k];
}

/* Rectangle [lo,hi] with nx x ny quadrilaterals.  vertices: id = vx*(ny+1)+vy, coordinates padded
   to 3 as DOLFINx stores them; cell c = cx*ny+cy; local vertex v = a + 2b <-> (cx+a, cy+b)
   (DOLFINx quadrilateral order, x fastest). */
void fo_rect_mesh(int nx, int ny, const double* lo, const double* hi, double* xg,
                  int32_t* xdofmap) {
  for (int vx = 0; vx <= nx; ++vx)
    for (int vy = 0; vy <= ny; ++vy) {
      size_t id = (size_t)vx * (ny + 1) + vy;
      xg[3 * id + 0] = lo[0] + (hi[0] - lo[0]) * vx / nx;
      xg[3 * id + 1] = lo[1] + (hi[1] - lo[1]) * vy / ny;
      xg[3 * id + 2] = 0.0;
    }
  for (int cx = 0; cx < nx; ++cx)
    for (int cy = 0; cy < ny; ++cy)
      for (int v = 0; v < 4; ++v)
        xdofmap[4 * ((size_t)cx * ny + cy) + v]
            = (int32_t)((size_t)(cx + (v & 1)) * (ny + 1) + (cy + (v >> 1)));
}

/* tensor_dofmap[c*N*N + i0*N + i1], lexicographic node numbering (x slowest) */
void fo_rect_dofmap(int P, int nx, int ny, int32_t* dofmap) {
  int N = P + 1;
  int64_t My = (int64_t)ny * P + 1;
  (void)nx;
  for (int cx = 0; cx < nx; ++cx)
    for (int cy = 0; cy < ny; ++cy)
      for (int i0 = 0; i0 < N; ++i0)
        for (int i1 = 0; i1 < N; ++i1)
          dofmap[((int64_t)cx * ny + cy) * N * N + i0 * N + i1]
              = (int32_t)(((int64_t)cx * P + node_pos(i0, P)) * My + cy * P + node_pos(i1, P));
}

/* J[i][j] = d x_i / d xi_j of the bilinear map */
static void q1_jacobian_2d(const double X[4][2], const double xi[2], double J[2][2]) {
  memset(J, 0, sizeof(double) * 4);
  for (int v = 0; v < 4; ++v) {
    int a = v & 1, b = v >> 1;
    double l0 = a ? xi[0] : 1.0 - xi[0], l1 = b ? xi[1] : 1.0 - xi[1];
    double g[2] = {(a ? 1.0 : -1.0) * l1, l0 * (b ? 1.0 : -1.0)};
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        J[i][j] += X[v][i] * g[j];
  }
}

/* G[c][q][3] = |detJ| w_q {G00, G01, G11}, detJ[c][q] = |detJ| w_q, q = q0*N + q1
   (fenicsx-sf-naive precompute.hpp:101-213 with gdim == 2, :199-203) */
void fo_geometry_2d(int64_t nc, const double* xg, const int32_t* xdofmap, int N, const double* pts,
                    const double* wts, double* G, double* detJ) {
  int Nd = N * N;
  for (int64_t c = 0; c < nc; ++c) {
    double X[4][2];
    for (int v = 0; v < 4; ++v)
      for (int j = 0; j < 2; ++j)
        X[v][j] = xg[3 * (size_t)xdofmap[4 * c + v] + j];
    for (int q0 = 0; q0 < N; ++q0)
      for (int q1 = 0; q1 < N; ++q1) {
        int q = q0 * N + q1;
        double xi[2] = {pts[q0], pts[q1]}, J[2][2];
        q1_jacobian_2d(X, xi, J);
        double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        double dj = fabs(det) * wts[q0] * wts[q1];
        if (detJ)
          detJ[c * Nd + q] = dj;
        if (G) {
          double K[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
          double* o = G + ((size_t)c * Nd + q) * 3;
          o[0] = dj * (K[0][0] * K[0][0] + K[0][1] * K[0][1]);
          o[1] = dj * (K[0][0] * K[1][0] + K[0][1] * K[1][1]);
          o[2] = dj * (K[1][0] * K[1][0] + K[1][1] * K[1][1]);
        }
      }
  }
}

/* MassSpectral2D::operator() (fenicsx-sf-naive spectral_op.hpp:69-86) */
void fo_mass_apply_2d(int P, int64_t nc, const int32_t* dofmap, const double* detJ,
                      const double* coeffs, const double* x, double* y) {
  int N = P + 1, Nd = N * N;
  double* x_ = (double*)malloc(sizeof(double) * Nd);
  for (int64_t c = 0; c < nc; ++c) {
    for (int i = 0; i < Nd; ++i)
      x_[i] = x[dofmap[c * Nd + i]];
    for (int iq = 0; iq < Nd; ++iq)
      x_[iq] = coeffs[c] * x_[iq] * detJ[c * Nd + iq];
    for (int i = 0; i < Nd; ++i)
      y[dofmap[c * Nd + i]] += x_[i];
  }
  free(x_);
}

/* StiffnessSpectral2D::operator() (fenicsx-sf-naive spectral_op.hpp:275-318), call for call;
   the 2-D stiffness::transform is :196-208 */
void fo_stiffness_apply_2d(int P, int64_t nc, const int32_t* dofmap, const double* G,
                           const double* dphi, const double* coeffs, const double* x, double* y) {
  int N = P + 1, Nd = N * N;
  size_t bytes = sizeof(double) * Nd;
  double* buf = (double*)malloc(bytes * 8);
  double *x_ = buf, *fw0 = buf + Nd, *fw1 = buf + 2 * Nd, *y0 = buf + 3 * Nd, *y1 = buf + 4 * Nd,
         *T1 = buf + 5 * Nd, *T2 = buf + 6 * Nd, *dphiT = buf + 7 * Nd;
  fo_transpose_2d(N, N, 1, N, dphi, dphiT); /* :271-272 */
  for (int64_t c = 0; c < nc; ++c) {
    for (int i = 0; i < Nd; ++i)
      x_[i] = x[dofmap[c * Nd + i]];
    memset(T1, 0, bytes);
    memset(T2, 0, bytes);
    memset(fw0, 0, bytes);
    fo_contract_2d(N, N, N, x_, dphi, fw0); /* [i1,i2] x [q2,i2] -> [i1,q2] */
    memset(fw1, 0, bytes);
    fo_transpose_2d(N, N, 1, N, x_, T1);
    fo_contract_2d(N, N, N, T1, dphi, T2);
    fo_transpose_2d(N, N, 1, N, T2, fw1);
    const double* Gc = G + (size_t)c * Nd * 3;
    double coeff = coeffs[c];
    for (int iq = 0; iq < Nd; ++iq) {
      const double* _G = Gc + iq * 3;
      double w0 = fw0[iq], w1 = fw1[iq];
      fw0[iq] = coeff * (_G[2] * w0 + _G[1] * w1);
      fw1[iq] = coeff * (_G[1] * w0 + _G[0] * w1);
    }
    memset(T1, 0, bytes);
    memset(T2, 0, bytes);
    memset(y0, 0, bytes);
    fo_contract_2d(N, N, N, fw0, dphiT, y0);
    memset(y1, 0, bytes);
    fo_transpose_2d(N, N, 1, N, fw1, T1);
    fo_contract_2d(N, N, N, T1, dphiT, T2);
    fo_transpose_2d(N, N, 1, N, T2, y1);
    for (int i = 0; i < Nd; ++i)
      y[dofmap[c * Nd + i]] += y0[i] + y1[i];
  }
  free(buf);
}

/* Edge (cell, lf) of a quadrilateral: local tensor indices of its N nodes and w_a * |dx/ds|.
   DOLFINx quadrilateral facets: 0: xi1=0, 1: xi0=0, 2: xi0=1, 3: xi1=1. */
void fo_facet_data_2d(int N, const double* xg, const int32_t* xdofmap, const double* pts,
                      const double* wts, int64_t cell, int lf, int32_t* fnodes, double* fscale) {
  static const int fdir[4] = {1, 0, 0, 1}, fside[4] = {0, 0, 1, 1};
  int dir = fdir[lf], side = fside[lf], ta = 1 - dir;
  double X[4][2];
  for (int v = 0; v < 4; ++v)
    for (int j = 0; j < 2; ++j)
      X[v][j] = xg[3 * (size_t)xdofmap[4 * cell + v] + j];
  for (int a = 0; a < N; ++a) {
    int idx[2];
    idx[dir] = side;
    idx[ta] = a;
    double xi[2] = {pts[idx[0]], pts[idx[1]]}, J[2][2];
    q1_jacobian_2d(X, xi, J);
    fnodes[a] = idx[0] * N + idx[1];
    fscale[a] = wts[a] * sqrt(J[0][ta] * J[0][ta] + J[1][ta] * J[1][ta]);
  }
}

/* Exterior edges of the rectangle {cell, local facet, tag}: tag 1 on x=lo, 2 on x=hi, 0 elsewhere */
int64_t fo_rect_facets(int nx, int ny, int32_t* facets) {
  int64_t k = 0;
  for (int cx = 0; cx < nx; ++cx)
    for (int cy = 0; cy < ny; ++cy) {
      int on[4] = {cy == 0, cx == 0, cx == nx - 1, cy == ny - 1};
      for (int lf = 0; lf < 4; ++lf)
        if (on[lf]) {
          if (facets) {
            facets[3 * k] = (int32_t)((int64_t)cx * ny + cy);
            facets[3 * k + 1] = lf;
            facets[3 * k + 2] = (lf == 1) ? 1 : ((lf == 2) ? 2 : 0);
          }
          ++k;
        }
    }
  return k;
}
