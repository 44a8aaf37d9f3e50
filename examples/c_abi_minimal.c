/* c_abi_minimal.c -- the C ABI from plain C (no C++ types cross the boundary): one stiffness
 * application and a few RK4 steps of the linear model on a small box.
 *
 *   gcc -std=c11 -O2 -Iinclude examples/c_abi_minimal.c -Lfenicsx-fus_b200/lib -lfus_b200 -lm \
 *       -Wl,-rpath,$PWD/fenicsx-fus_b200/lib -o examples/c_abi_minimal
 *
 * Mirrors what a driver written for the reference does around StiffnessSpectral3D and
 * LinearSpectral3D (cpp/fenicsx-sf/common/spectral_op.hpp:132-243, Linear.hpp:55-318). */
#include <fus_b200.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(call)                                                                                \
  do {                                                                                             \
    int rc__ = (call);                                                                             \
    if (rc__ != FUS_OK) {                                                                          \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, fus_last_error());                      \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

int main(void) {
  const int P = 3, n[3] = {4, 3, 2}, N = P + 1, Nd = N * N * N;
  const double lo[3] = {0, 0, 0}, hi[3] = {0.008, 0.006, 0.004};
  const int64_t ncells = (int64_t)n[0] * n[1] * n[2];
  const int64_t nverts = (int64_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
  const int64_t ndofs = fus_box_num_dofs(P, n);

  double* xg = malloc(sizeof(double) * 3 * nverts);
  int32_t* xdofmap = malloc(sizeof(int32_t) * 8 * ncells);
  int32_t* dofmap = malloc(sizeof(int32_t) * Nd * ncells);
  CHECK(fus_box_mesh(n, lo, hi, xg, xdofmap));
  CHECK(fus_box_dofmap(P, n, 1, dofmap));
  const int64_t nfacets = fus_box_facets(n, NULL);
  int32_t* facets = malloc(sizeof(int32_t) * 3 * nfacets);
  fus_box_facets(n, facets);

  fus_ctx* ctx = NULL;
  CHECK(fus_ctx_create_from_mesh(P, ncells, ndofs, ndofs, dofmap, nverts, xg, xdofmap, 0, &ctx));

  /* y += K(-1/rho) x */
  double *x = malloc(sizeof(double) * ndofs), *y = calloc(ndofs, sizeof(double));
  double* coeffs = malloc(sizeof(double) * ncells);
  for (int64_t i = 0; i < ndofs; ++i)
    x[i] = sin(0.01 * (double)i);
  for (int64_t c = 0; c < ncells; ++c)
    coeffs[c] = -1.0 / 1000.0;
  CHECK(fus_stiffness_apply_host(ctx, x, coeffs, y));
  double k2 = 0;
  for (int64_t i = 0; i < ndofs; ++i)
    k2 += y[i] * y[i];
  printf("Kx_l2: %.17g\n", sqrt(k2));

  /* LinearSpectral3D: water, 0.5 MHz planar source on x = 0, absorbing x = hi */
  double *c0 = malloc(sizeof(double) * ncells), *rho0 = malloc(sizeof(double) * ncells);
  for (int64_t c = 0; c < ncells; ++c) {
    c0[c] = 1500.0;
    rho0[c] = 1000.0;
  }
  double *src = malloc(sizeof(double) * ndofs), *absb = malloc(sizeof(double) * ndofs);
  CHECK(fus_boundary_vectors(FUS_LINEAR, P, ncells, ndofs, xg, xdofmap, dofmap, nfacets, facets, c0,
                             rho0, NULL, src, NULL, absb, NULL));
  fus_model* model = NULL;
  CHECK(fus_model_create(ctx, FUS_LINEAR, c0, rho0, NULL, NULL, src, NULL, absb, NULL, 0.5e6,
                         60000.0, 1500.0, &model));
  const double dt = 4.0e-8;
  int nsteps = 0;
  CHECK(fus_model_set_state(model, NULL, NULL));
  CHECK(fus_model_rk4(model, 0.0, 9.5 * dt, dt, &nsteps));
  CHECK(fus_model_get_state(model, x, NULL));
  double u2 = 0;
  for (int64_t i = 0; i < ndofs; ++i)
    u2 += x[i] * x[i];
  printf("Number of steps: %d\nu_l2: %.17g\n", nsteps, sqrt(u2));

  if (fus_ctx_destroy(ctx) != FUS_ERR_STATE) { /* refused while the model is alive */
    fprintf(stderr, "fus_ctx_destroy should refuse while a model exists\n");
    return 1;
  }
  CHECK(fus_model_destroy(model));
  CHECK(fus_ctx_destroy(ctx));
  free(xg); free(xdofmap); free(dofmap); free(facets); free(x); free(y); free(coeffs);
  free(c0); free(rho0); free(src); free(absb);
  return 0;
}
