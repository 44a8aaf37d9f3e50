// fus_host.cpp -- host-side setup of the hot path (runs once per mesh; not the hot loop).
//
// Stands in for the Basix / DOLFINx / FFCx calls made by the reference constructors:
//   GLL rule            basix::quadrature::make_quadrature   (spectral_op.hpp:57-59,160-162)
//   1-D derivative tab  tabulate_1d                          (precompute.hpp:217-234)
//   box mesh            dolfinx::mesh::create_box            (experiments/.../main.cpp:61-65)
//   tensor dofmap       create_functionspace + reorder_dofmap (permute.hpp:15-42)
//   boundary vectors    FFCx `ds` kernels of forms.py + fem::assemble_vector (Linear.hpp:133,205)
// (paths relative to cpp/fenicsx-sf/common/ of the reference)
#include "fus_internal.hpp"
#include "fus_trilinear.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace fus {

// ---- Gauss-Lobatto-Legendre rule ------------------------------------------------------------

namespace {
// P_n and P_n' at x via the Bonnet recurrence and its derivative.
void legendre(int n, double x, double& p, double& dp) {
  double pm1 = 1.0, pc = x, dm1 = 0.0, dc = 1.0;
  if (n == 0) {
    p = 1.0;
    dp = 0.0;
    return;
  }
  for (int k = 1; k < n; ++k) {
    const double pn = ((2 * k + 1) * x * pc - k * pm1) / (k + 1);
    const double dn = dm1 + (2 * k + 1) * pc;
    pm1 = pc;
    pc = pn;
    dm1 = dc;
    dc = dn;
  }
  p = pc;
  dp = dc;
}
} // namespace

// m = P+1 points on [-1,1], ascending.  Interior nodes are the roots of P_n'(x), n = m-1,
// found by Newton on f = P_n' with f' = P_n'' from the Legendre ODE.
static void gll_reference_interval(int m, std::vector<double>& x, std::vector<double>& w) {
  const int n = m - 1;
  x.assign(m, 0.0);
  w.assign(m, 0.0);
  x[0] = -1.0;
  x[n] = 1.0;
  for (int j = 1; j < n; ++j) {
    double xj = -std::cos(M_PI * (j + 0.25) / n - 3.0 / (8.0 * n * M_PI * (j + 0.25)));
    for (int it = 0; it < 100; ++it) {
      double p, dp;
      legendre(n, xj, p, dp);
      const double ddp = (2.0 * xj * dp - n * (n + 1.0) * p) / (1.0 - xj * xj);
      const double dx = dp / ddp;
      xj -= dx;
      if (std::fabs(dx) < 1e-16)
        break;
    }
    x[j] = xj;
  }
  // enforce symmetry about 0
  for (int j = 0; j < m / 2; ++j) {
    const double a = 0.5 * (x[n - j] - x[j]);
    x[j] = -a;
    x[n - j] = a;
  }
  if (m % 2 == 1)
    x[m / 2] = 0.0;
  for (int j = 0; j < m; ++j) {
    double p, dp;
    legendre(n, x[j], p, dp);
    w[j] = 2.0 / (n * (n + 1.0) * p * p);
  }
}

int gll(int P, double* pts, double* wts) {
  if (P < 1 || P > 15)
    return FUS_ERR_ARG;
  const int m = P + 1;
  std::vector<double> x, w;
  gll_reference_interval(m, x, w);
  // [0,1], Basix order: both end points first, then the interior ascending
  pts[0] = 0.0;
  pts[1] = 1.0;
  wts[0] = 0.5 * w[0];
  wts[1] = 0.5 * w[m - 1];
  for (int j = 1; j < m - 1; ++j) {
    pts[j + 1] = 0.5 * (x[j] + 1.0);
    wts[j + 1] = 0.5 * w[j];
  }
  return FUS_OK;
}

int tabulate_dphi(int P, double* dphi) {
  if (P < 1 || P > 15)
    return FUS_ERR_ARG;
  const int N = P + 1;
  std::vector<double> x(N), w(N);
  gll(P, x.data(), w.data());
  // phi_i'(x_q) = sum_{j != i} 1/(x_i - x_j) prod_{l != i,j} (x_q - x_l)/(x_i - x_l)
  for (int q = 0; q < N; ++q)
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      for (int j = 0; j < N; ++j) {
        if (j == i)
          continue;
        double t = 1.0 / (x[i] - x[j]);
        for (int l = 0; l < N; ++l)
          if (l != i && l != j)
            t *= (x[q] - x[l]) / (x[i] - x[l]);
        s += t;
      }
      dphi[q * N + i] = s;
    }
  return FUS_OK;
}

// ---- structured box ---------------------------------------------------------------------------

int box_mesh(const int n[3], const double lo[3], const double hi[3], double* xg, int32_t* xdofmap) {
  if (n[0] < 1 || n[1] < 1 || n[2] < 1)
    return FUS_ERR_ARG;
  const int64_t vy = n[1] + 1, vz = n[2] + 1;
  for (int64_t i = 0; i <= n[0]; ++i)
    for (int64_t j = 0; j <= n[1]; ++j)
      for (int64_t k = 0; k <= n[2]; ++k) {
        double* p = xg + 3 * ((i * vy + j) * vz + k);
        p[0] = lo[0] + (hi[0] - lo[0]) * (double)i / n[0];
        p[1] = lo[1] + (hi[1] - lo[1]) * (double)j / n[1];
        p[2] = lo[2] + (hi[2] - lo[2]) * (double)k / n[2];
      }
  int64_t c = 0;
  for (int64_t i = 0; i < n[0]; ++i)
    for (int64_t j = 0; j < n[1]; ++j)
      for (int64_t k = 0; k < n[2]; ++k, ++c)
        for (int v = 0; v < 8; ++v) // v = a + 2b + 4c, x fastest
          xdofmap[8 * c + v]
              = (int32_t)(((i + (v & 1)) * vy + (j + ((v >> 1) & 1))) * vz + (k + (v >> 2)));
  return FUS_OK;
}

int64_t box_num_dofs(int P, const int n[3]) {
  return ((int64_t)n[0] * P + 1) * ((int64_t)n[1] * P + 1) * ((int64_t)n[2] * P + 1);
}

int box_dofmap(int P, const int n[3], int numbering, int32_t* dm) {
  if (P < 1 || n[0] < 1 || n[1] < 1 || n[2] < 1 || numbering < 0 || numbering > 1)
    return FUS_ERR_ARG;
  if (box_num_dofs(P, n) > INT32_MAX)
    return FUS_ERR_UNSUPPORTED;
  const int N = P + 1;
  const int64_t My = (int64_t)n[1] * P + 1, Mz = (int64_t)n[2] * P + 1;
  // grid offset of 1-D node i (Basix order) inside its cell
  std::vector<int> pos(N);
  pos[0] = 0;
  pos[1] = P;
  for (int i = 2; i < N; ++i)
    pos[i] = i - 1;

  // cell-blocked numbering: each grid node is numbered inside the block of the cell whose
  // lower corner region contains it; blocks are laid out in cell order.
  std::vector<int64_t> block_start;
  auto bsize = [&](int d, int64_t cd) { return (cd == n[d] - 1) ? P + 1 : P; };
  if (numbering == 1) {
    block_start.resize((size_t)n[0] * n[1] * n[2]);
    int64_t acc = 0, c = 0;
    for (int64_t i = 0; i < n[0]; ++i)
      for (int64_t j = 0; j < n[1]; ++j)
        for (int64_t k = 0; k < n[2]; ++k, ++c) {
          block_start[c] = acc;
          acc += (int64_t)bsize(0, i) * bsize(1, j) * bsize(2, k);
        }
  }
  auto number = [&](int64_t gx, int64_t gy, int64_t gz) -> int64_t {
    if (numbering == 0)
      return (gx * My + gy) * Mz + gz;
    int64_t ci = std::min<int64_t>(gx / P, n[0] - 1), cj = std::min<int64_t>(gy / P, n[1] - 1),
            ck = std::min<int64_t>(gz / P, n[2] - 1);
    const int64_t oc = (ci * n[1] + cj) * n[2] + ck;
    return block_start[oc]
           + ((gx - ci * P) * bsize(1, cj) + (gy - cj * P)) * bsize(2, ck) + (gz - ck * P);
  };
  int64_t c = 0;
  for (int64_t i = 0; i < n[0]; ++i)
    for (int64_t j = 0; j < n[1]; ++j)
      for (int64_t k = 0; k < n[2]; ++k, ++c) {
        int32_t* row = dm + c * N * N * N;
        for (int a = 0; a < N; ++a)
          for (int b = 0; b < N; ++b)
            for (int d = 0; d < N; ++d)
              row[(a * N + b) * N + d]
                  = (int32_t)number(i * P + pos[a], j * P + pos[b], k * P + pos[d]);
      }
  return FUS_OK;
}

int64_t box_facets(const int n[3], int32_t* facets) {
  // DOLFINx hexahedron facets: 0:z=0 1:y=0 2:x=0 3:x=1 4:y=1 5:z=1
  int64_t count = 0, c = 0;
  for (int64_t i = 0; i < n[0]; ++i)
    for (int64_t j = 0; j < n[1]; ++j)
      for (int64_t k = 0; k < n[2]; ++k, ++c) {
        const bool on[6] = {k == 0, j == 0, i == 0, i == n[0] - 1, j == n[1] - 1, k == n[2] - 1};
        for (int f = 0; f < 6; ++f) {
          if (!on[f])
            continue;
          if (facets) {
            facets[3 * count + 0] = (int32_t)c;
            facets[3 * count + 1] = f;
            facets[3 * count + 2] = (f == 2) ? 1 : (f == 3 ? 2 : 0);
          }
          ++count;
        }
      }
  return count;
}

// ---- boundary vectors ---------------------------------------------------------------------------

namespace {
// Tangent vectors of the trilinear map of a cell at reference point xi: column `axis` of J.
void tangent(const double X[8][3], const double xi[3], int axis, double t[3]) {
  t[0] = t[1] = t[2] = 0.0;
  for (int v = 0; v < 8; ++v) {
    const int bit[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
    double g = 1.0;
    for (int d = 0; d < 3; ++d) {
      if (d == axis)
        g *= bit[d] ? 1.0 : -1.0;
      else
        g *= bit[d] ? xi[d] : 1.0 - xi[d];
    }
    for (int r = 0; r < 3; ++r)
      t[r] += X[v][r] * g;
  }
}
} // namespace

int boundary_vectors(int kind, int P, int64_t ncells, int64_t ndofs, const double* xg,
                     const int32_t* xdofmap, const int32_t* dm, int64_t nfacets,
                     const int32_t* facets, const double* c0, const double* rho0,
                     const double* delta0, double* src, double* dsrc, double* absb,
                     double* bmass) {
  if (kind < 0 || kind > 2 || !xg || !xdofmap || !dm || !c0 || !rho0)
    return FUS_ERR_ARG;
  if (kind != FUS_LINEAR && !delta0)
    return FUS_ERR_ARG;
  const int N = P + 1, Nd = N * N * N;
  std::vector<double> pts(N), wts(N);
  gll(P, pts.data(), wts.data());
  for (double* v : {src, dsrc, absb, bmass})
    if (v)
      std::fill(v, v + ndofs, 0.0);
  static const int fdir[6] = {2, 1, 0, 0, 1, 2}, fside[6] = {0, 0, 0, 1, 1, 1};
  for (int64_t f = 0; f < nfacets; ++f) {
    const int64_t c = facets[3 * f];
    const int lf = facets[3 * f + 1], tag = facets[3 * f + 2];
    if (c < 0 || c >= ncells || lf < 0 || lf > 5)
      return FUS_ERR_ARG;
    const int dir = fdir[lf], ta = (dir == 0) ? 1 : 0, tb = (dir == 2) ? 1 : 2;
    double X[8][3];
    for (int v = 0; v < 8; ++v)
      for (int r = 0; r < 3; ++r)
        X[v][r] = xg[3 * (int64_t)xdofmap[8 * c + v] + r];
    const double rho = rho0[c], cc = c0[c], del = delta0 ? delta0[c] : 0.0;
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b) {
        int id[3];
        id[dir] = fside[lf]; // node 0 sits at xi=0, node 1 at xi=1
        id[ta] = a;
        id[tb] = b;
        const double xi[3] = {pts[id[0]], pts[id[1]], pts[id[2]]};
        double t1[3], t2[3];
        tangent(X, xi, ta, t1);
        tangent(X, xi, tb, t2);
        const double nx = t1[1] * t2[2] - t1[2] * t2[1], ny = t1[2] * t2[0] - t1[0] * t2[2],
                     nz = t1[0] * t2[1] - t1[1] * t2[0];
        const double s = wts[a] * wts[b] * std::sqrt(nx * nx + ny * ny + nz * nz);
        const int32_t d = dm[c * Nd + (id[0] * N + id[1]) * N + id[2]];
        if (tag == 1 && src)
          src[d] += s / rho;
        if (kind == FUS_LINEAR) {
          if (tag == 2 && absb)
            absb[d] += s / rho / cc;
        } else {
          if (absb)
            absb[d] += s / rho / cc;
          if (tag == 1 && dsrc)
            dsrc[d] += s * del / rho / cc / cc;
          if (bmass)
            bmass[d] += s * del / rho / cc / cc / cc;
        }
      }
  }
  return FUS_OK;
}

// ---- trilinear cell map (option geometry_mode = 2) ----------------------------------------------

int trilinear_coeffs(int64_t ncells, const double* xg, const int32_t* xdofmap, double* coeffs) {
  if (ncells < 0 || !xg || !xdofmap || !coeffs)
    return FUS_ERR_ARG;
  for (int64_t c = 0; c < ncells; ++c) {
    double X[8][3];
    for (int v = 0; v < 8; ++v)
      for (int r = 0; r < 3; ++r)
        X[v][r] = xg[3 * (int64_t)xdofmap[8 * c + v] + r];
    tri_cell_coeffs(X, coeffs + c * FUS_TRI_STRIDE);
  }
  return FUS_OK;
}

// G and detJ of the reference (precompute.hpp:33-213) for a batch of cells, rebuilt point by point
// with exactly the helpers the trilinear stiffness kernel runs: column b of the scaled K K^T is
// the transform of the unit vector e_b.
int trilinear_geometry(int P, int64_t ncells, const double* coeffs, double* G, double* detJ) {
  if (P < 1 || P > 15 || ncells < 0 || !coeffs)
    return FUS_ERR_ARG;
  const int N = P + 1;
  std::vector<double> pts(N), wts(N);
  gll(P, pts.data(), wts.data());
  for (int64_t c = 0; c < ncells; ++c)
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b) {
        TriLine L;
        tri_line_setup(coeffs + c * FUS_TRI_STRIDE, pts[a], pts[b], L);
        for (int i0 = 0; i0 < N; ++i0) {
          const int64_t q = (c * N + i0) * N * N + a * N + b;
          const double w = wts[i0] * (wts[a] * wts[b]);
          double col[3][3], adet = 0.0;
          for (int e = 0; e < 3; ++e)
            adet = tri_transform(L, pts[i0], w, e == 0, e == 1, e == 2, col[e][0], col[e][1],
                                 col[e][2]);
          if (detJ)
            detJ[q] = adet * w;
          if (G) {
            double* g = G + 6 * q;
            g[0] = col[0][0], g[1] = col[1][0], g[2] = col[2][0];
            g[3] = col[1][1], g[4] = col[2][1], g[5] = col[2][2];
          }
        }
      }
  return FUS_OK;
}

// ---- 2-D quadrilateral variant (cpp/fenicsx-sf-naive/common) ------------

int rect_mesh(const int n[2], const double lo[2], const double hi[2], double* xg,
              int32_t* xdofmap) {
  if (n[0] < 1 || n[1] < 1)
    return FUS_ERR_ARG;
  const int64_t vy = n[1] + 1;
  for (int64_t i = 0; i <= n[0]; ++i)
    for (int64_t j = 0; j <= n[1]; ++j) {
      double* p = xg + 3 * (i * vy + j); // padded to 3 coordinates, as DOLFINx stores geometry
      p[0] = lo[0] + (hi[0] - lo[0]) * (double)i / n[0];
      p[1] = lo[1] + (hi[1] - lo[1]) * (double)j / n[1];
      p[2] = 0.0;
    }
  for (int64_t i = 0; i < n[0]; ++i)
    for (int64_t j = 0; j < n[1]; ++j)
      for (int v = 0; v < 4; ++v) // v = a + 2b, x fastest (DOLFINx quadrilateral vertex order)
        xdofmap[4 * (i * n[1] + j) + v] = (int32_t)((i + (v & 1)) * vy + (j + (v >> 1)));
  return FUS_OK;
}

int64_t rect_num_dofs(int P, const int n[2]) {
  return ((int64_t)n[0] * P + 1) * ((int64_t)n[1] * P + 1);
}

int rect_dofmap(int P, const int n[2], int32_t* dm) {
  if (P < 1 || n[0] < 1 || n[1] < 1)
    return FUS_ERR_ARG;
  if (rect_num_dofs(P, n) > INT32_MAX)
    return FUS_ERR_UNSUPPORTED;
  const int N = P + 1;
  const int64_t My = (int64_t)n[1] * P + 1;
  std::vector<int> pos(N);
  pos[0] = 0;
  pos[1] = P;
  for (int i = 2; i < N; ++i)
    pos[i] = i - 1;
  for (int64_t i = 0; i < n[0]; ++i)
    for (int64_t j = 0; j < n[1]; ++j) {
      int32_t* row = dm + (i * n[1] + j) * N * N;
      for (int a = 0; a < N; ++a)
        for (int b = 0; b < N; ++b)
          row[a * N + b] = (int32_t)((i * P + pos[a]) * My + (j * P + pos[b]));
    }
  return FUS_OK;
}

int64_t rect_facets(const int n[2], int32_t* facets) {
  // DOLFINx quadrilateral facets: 0: y=0, 1: x=0, 2: x=1, 3: y=1
  int64_t count = 0;
  for (int64_t i = 0; i < n[0]; ++i)
    for (int64_t j = 0; j < n[1]; ++j) {
      const bool on[4] = {j == 0, i == 0, i == n[0] - 1, j == n[1] - 1};
      for (int f = 0; f < 4; ++f) {
        if (!on[f])
          continue;
        if (facets) {
          facets[3 * count + 0] = (int32_t)(i * n[1] + j);
          facets[3 * count + 1] = f;
          facets[3 * count + 2] = (f == 1) ? 1 : (f == 2 ? 2 : 0);
        }
        ++count;
      }
    }
  return count;
}

// Edge-lumped boundary vectors: the `ds` forms of the 2-D examples
// (cpp/fenicsx-sf-naive/examples/linear_planewave2d_1/forms.py:34-40 and the lossy / Westervelt
// ones) with GLL quadrature are collocated, exactly as in 3-D.
int boundary_vectors_2d(int kind, int P, int64_t ncells, int64_t ndofs, const double* xg,
                        const int32_t* xdofmap, const int32_t* dm, int64_t nfacets,
                        const int32_t* facets, const double* c0, const double* rho0,
                        const double* delta0, double* src, double* dsrc, double* absb,
                        double* bmass) {
  if (kind < 0 || kind > 2 || !xg || !xdofmap || !dm || !c0 || !rho0)
    return FUS_ERR_ARG;
  if (kind != FUS_LINEAR && !delta0)
    return FUS_ERR_ARG;
  const int N = P + 1, Nd = N * N;
  std::vector<double> pts(N), wts(N);
  gll(P, pts.data(), wts.data());
  for (double* v : {src, dsrc, absb, bmass})
    if (v)
      std::fill(v, v + ndofs, 0.0);
  static const int fdir[4] = {1, 0, 0, 1}, fside[4] = {0, 0, 1, 1};
  for (int64_t f = 0; f < nfacets; ++f) {
    const int64_t c = facets[3 * f];
    const int lf = facets[3 * f + 1], tag = facets[3 * f + 2];
    if (c < 0 || c >= ncells || lf < 0 || lf > 3)
      return FUS_ERR_ARG;
    const int dir = fdir[lf], ta = 1 - dir;
    double X[4][2];
    for (int v = 0; v < 4; ++v)
      for (int r = 0; r < 2; ++r)
        X[v][r] = xg[3 * (int64_t)xdofmap[4 * c + v] + r];
    const double rho = rho0[c], cc = c0[c], del = delta0 ? delta0[c] : 0.0;
    for (int a = 0; a < N; ++a) {
      int id[2];
      id[dir] = fside[lf];
      id[ta] = a;
      const double xi[2] = {pts[id[0]], pts[id[1]]};
      // tangent of the bilinear map along reference axis ta
      double t[2] = {0.0, 0.0};
      for (int v = 0; v < 4; ++v) {
        const int bit[2] = {v & 1, v >> 1};
        const double g = (bit[ta] ? 1.0 : -1.0) * (bit[dir] ? xi[dir] : 1.0 - xi[dir]);
        t[0] += X[v][0] * g;
        t[1] += X[v][1] * g;
      }
      const double s = wts[a] * std::sqrt(t[0] * t[0] + t[1] * t[1]);
      const int32_t d = dm[c * Nd + id[0] * N + id[1]];
      if (tag == 1 && src)
        src[d] += s / rho;
      // the 2-D forms of cpp/fenicsx-sf-naive integrate the absorbing term and its mass-like
      // counterpart over ds(2) for every model (examples/lossy_planewave2d_1/forms.py:37-42,
      // westervelt_planewave2d_1/forms.py:37-42) -- unlike the 3-D forms of cpp/fenicsx-sf, which
      // use `ds` without an id for the lossy and Westervelt models (boundary_vectors above)
      if (tag == 2 && absb)
        absb[d] += s / rho / cc;
      if (kind != FUS_LINEAR) {
        if (tag == 1 && dsrc)
          dsrc[d] += s * del / rho / cc / cc;
        if (tag == 2 && bmass)
          bmass[d] += s * del / rho / cc / cc / cc;
      }
    }
  }
  return FUS_OK;
}

} // namespace fus
