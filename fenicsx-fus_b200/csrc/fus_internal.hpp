// fus_internal.hpp -- declarations shared by the host setup code and the CUDA translation unit.
#pragma once
#include "fus_b200.h"

#include <cstdint>

namespace fus {

// host setup (fus_host.cpp)
int gll(int P, double* pts, double* wts);
int tabulate_dphi(int P, double* dphi);
int box_mesh(const int n[3], const double lo[3], const double hi[3], double* xg, int32_t* xdofmap);
int box_dofmap(int P, const int n[3], int numbering, int32_t* dm);
int64_t box_num_dofs(int P, const int n[3]);
int64_t box_facets(const int n[3], int32_t* facets);
int boundary_vectors(int kind, int P, int64_t ncells, int64_t ndofs, const double* xg,
                     const int32_t* xdofmap, const int32_t* dm, int64_t nfacets,
                     const int32_t* facets, const double* c0, const double* rho0,
                     const double* delta0, double* src, double* dsrc, double* absb, double* bmass);
int trilinear_coeffs(int64_t ncells, const double* xg, const int32_t* xdofmap, double* coeffs);
int trilinear_geometry(int P, int64_t ncells, const double* coeffs, double* G, double* detJ);
// 2-D quadrilateral variant
int rect_mesh(const int n[2], const double lo[2], const double hi[2], double* xg, int32_t* xdofmap);
int rect_dofmap(int P, const int n[2], int32_t* dm);
int64_t rect_num_dofs(int P, const int n[2]);
int64_t rect_facets(const int n[2], int32_t* facets);
int boundary_vectors_2d(int kind, int P, int64_t ncells, int64_t ndofs, const double* xg,
                        const int32_t* xdofmap, const int32_t* dm, int64_t nfacets,
                        const int32_t* facets, const double* c0, const double* rho0,
                        const double* delta0, double* src, double* dsrc, double* absb,
                        double* bmass);

void set_error(const char* fmt, ...);

} // namespace fus
