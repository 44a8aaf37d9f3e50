/*
 * fus_b200.h -- C ABI of the B200-native sum-factorised acoustic operator + RK4 path.
 *
 * This is the drop-in boundary for the hot path of adeebkor/fenicsx-fus
 * (cpp/fenicsx-sf/common/).  Plain pointers and sizes only; no C++/torch types.
 * Each entry point names the reference interface it stands in for (paths relative to
 * cpp/fenicsx-sf/common/ of the reference).  All functions return FUS_OK (0) or a
 * negative error code and never throw; fus_last_error() gives the message.
 *
 * There is NO CPU fallback: every compute entry point fails with FUS_ERR_CUDA when no
 * sm_100 device is usable.
 *
 * Conventions:
 *   N = P+1 nodes/points per direction, Nd = N^3, tensor index i = i0*N*N + i1*N + i2 with i0 <->
 *   reference direction 0; 1-D node order [0, 1, interior ascending] (Basix);
 *   dphi[q*N+i] = phi_i'(xi_q);  G[c][q][6] = {G00,G01,G02,G11,G12,G22} * |detJ| * w_q;
 *   local vectors hold `nowned` owned entries first, then ghosts.
 */
#ifndef FUS_B200_H
#define FUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FUS_OK 0
#define FUS_ERR_ARG (-1)
#define FUS_ERR_CUDA (-2)
#define FUS_ERR_UNSUPPORTED (-3)
#define FUS_ERR_STATE (-4)
#define FUS_ERR_COMM (-5)

#define FUS_LINEAR 0     /* LinearSpectral3D      (Linear.hpp:52)      */
#define FUS_LOSSY 1      /* LossySpectral3D       (Lossy.hpp:54)       */
#define FUS_WESTERVELT 2 /* WesterveltSpectral3D  (Westervelt.hpp:56)  */

typedef struct fus_ctx fus_ctx;     /* what the operator constructors build (spectral_op.hpp:135-171) */
typedef struct fus_model fus_model; /* what the solver constructors build   (Linear.hpp:55-158)       */

const char* fus_last_error(void);
int fus_version(void);
/* number of usable sm_100 devices (0 when none; never an error) */
int fus_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Host-side setup (CPU, runs once): the Basix / DOLFINx calls of the reference constructors.
 * ---------------------------------------------------------------------------------------- */

/* GLL rule with P+1 points on [0,1], Basix order.
   Replaces basix::quadrature::make_quadrature(gll, interval, Qdegree[P]) (spectral_op.hpp:35-59). */
int fus_gll(int P, double* pts, double* wts);

/* Derivative block of tabulate_1d (precompute.hpp:217-234, spectral_op.hpp:168-170): dphi[N*N]. */
int fus_tabulate_dphi(int P, double* dphi);

/* Structured hexahedral box [lo,hi] with n[3] cells: vertex coordinates xg[nverts][3] and the
   cell->vertex map xdofmap[ncells][8] in DOLFINx tensor vertex order (x fastest).
   Replaces dolfinx::mesh::create_box (experiments/measure_fraction_of_peak_performance/main.cpp:61-65). */
int fus_box_mesh(const int n[3], const double lo[3], const double hi[3], double* xg,
                 int32_t* xdofmap);

/* Tensor-product dofmap of degree P on that box: tensor_dofmap[ncells][Nd].
   Replaces create_functionspace + reorder_dofmap (permute.hpp:15-42).
   numbering 0: lexicographic (x slowest); 1: cell-blocked (dofs of one cell contiguous). */
int fus_box_dofmap(int P, const int n[3], int numbering, int32_t* tensor_dofmap);
int64_t fus_box_num_dofs(int P, const int n[3]);

/* Exterior facets of the box as {cell, local facet, tag} triplets; tag 1 on x=lo, 2 on x=hi,
   0 elsewhere.  facets may be NULL to query the count.  Returns the count (>=0) or an error. */
int64_t fus_box_facets(const int n[3], int32_t* facets);

/* Facet-lumped boundary vectors for a model kind.  With GLL quadrature the FFCx `ds` kernels of
   forms.py are collocated, so each boundary term is a diagonal vector times a nodal value
   (benchmarks/PH1/BM7-SC1/forms.py:37-42; fenicsx-sf-naive/benchmarks/PH1/SC2-BM1/forms.py:35-38).
   Inputs: mesh geometry, tensor dofmap, facets {cell, lf, tag}, per-cell c0/rho0/delta0 (delta0 may
   be NULL for FUS_LINEAR).  Outputs are dense over the local dofs (ndofs each, zero off-boundary):
     src   multiplies g(t):   S^(1/rho)(tag 1)
     dsrc  multiplies dg(t):  S^(delta/rho c^2)(tag 1)             [lossy, westervelt]
     absb  multiplies -v:     S^(1/rho c)(tag 2)  for linear,  S^(1/rho c)(all exterior) otherwise
     bmass added to the lumped mass: S^(delta/rho c^3)(all exterior) [lossy, westervelt] */
int fus_boundary_vectors(int kind, int P, int64_t ncells, int64_t ndofs, const double* xg,
                         const int32_t* xdofmap, const int32_t* tensor_dofmap, int64_t nfacets,
                         const int32_t* facets, const double* c0, const double* rho0,
                         const double* delta0, double* src, double* dsrc, double* absb,
                         double* bmass);

/* The trilinear cell map in monomial form: coeffs[ncells][FUS_TRI_STRIDE] doubles, 21 used
   ({c100,c010,c001,c110,c101,c011,c111} x 3 coordinates).  This is what the operator reads per
   CELL with option "geometry_mode" = 2 instead of the 6 doubles per POINT that
   compute_scaled_geometrical_factor stores (precompute.hpp:101-213). */
#define FUS_TRI_STRIDE 24
int fus_trilinear_coeffs(int64_t ncells, const double* xg, const int32_t* xdofmap,
                         double* coeffs);
/* G[ncells][Nd][6] and detJ[ncells][Nd] in the reference layouts (either may be NULL) rebuilt
   from those coefficients by the same arithmetic the trilinear stiffness kernel runs
   (host evaluation; used to check that path against precompute.hpp without a GPU). */
int fus_trilinear_geometry(int P, int64_t ncells, const double* coeffs, double* G, double* detJ);

/* ------------------------------------------------------------------------------------------
 * Operator context (device).  Owns device copies of the cell data; host arrays are only
 * borrowed during the call.
 * ---------------------------------------------------------------------------------------- */

/* From precomputed host arrays in the reference's layouts -- the members the reference operator
   classes hold (spectral_op.hpp:256-264): tensor_dofmap[ncells*Nd], G[ncells*Nd*6] (may be NULL if
   only the mass operator is needed), detJ[ncells*Nd] (may be NULL if only stiffness), dphi[N*N]. */
int fus_ctx_create(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                   const int32_t* tensor_dofmap, const double* G, const double* detJ,
                   const double* dphi, int device, fus_ctx** out);

/* From the mesh: G and detJ are evaluated on the device from the trilinear geometry
   (compute_scaled_geometrical_factor / compute_scaled_jacobian_determinant, precompute.hpp:33-213)
   and the 1-D tables are generated internally (fus_gll / fus_tabulate_dphi). */
int fus_ctx_create_from_mesh(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                             const int32_t* tensor_dofmap, int64_t nverts, const double* xg,
                             const int32_t* xdofmap, int device, fus_ctx** out);

/* Same, but neither G nor detJ is materialised: the context keeps the trilinear cell map
   (192 B per cell) and the dofmap only, "geometry_mode" is fixed at 2 and the mass operator
   rebuilds |det J| w the same way.  Device memory per dof drops from ~116 B to ~10 B of cell
   data at P = 4, so problems ~5x larger fit one GPU.  fus_ctx_get_geometry rebuilds on the host.
   (Setting the environment variable FUS_GEOMETRY_MODE=lean makes fus_ctx_create_from_mesh behave
   like this, FUS_GEOMETRY_MODE=2 makes it start in geometry_mode 2: for unmodified drivers.) */
int fus_ctx_create_from_mesh_lean(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                                  const int32_t* tensor_dofmap, int64_t nverts, const double* xg,
                                  const int32_t* xdofmap, int device, fus_ctx** out);

/* FUS_ERR_STATE (and nothing released) while fus_model objects created on it are alive. */
int fus_ctx_destroy(fus_ctx* ctx);

/* Launch all work of this context on the given cudaStream_t (default: a private stream). */
int fus_ctx_set_stream(fus_ctx* ctx, void* cuda_stream);
/* Tuning/diagnostic knobs (all optional; defaults in brackets):
     "stiffness_variant"  [-1] -1 auto: the fastest kernel per degree from the hardware sweep of
                               DESIGN.md 3.1 (column kernel for P <= 3, line kernel for P = 4,
                               its software pipelines 5 / 5 / 3 for P = 5 / 6 / 7),
                               0 column, 1 per-point (cross-check), 2 line kernel,
                               3..6 line kernel with software pipelines (same results),
                               7 cell-per-thread kernel (P = 2 only, other degrees keep their
                               own; a measured alternative, slower than the column kernel)
     "geometry_mode"      [0]  0 streamed G, 1 affine compression, 2 trilinear on the fly (below)
     "use_graph"          [1]  replay RK4 steps from a captured CUDA graph when possible
     "profile_kernels"    [0]  CUDA event pair around every launch (disables graph replay)
     "l2_persist"         [0]  L2 persistence window on the rhs accumulator (measured slower)
     "halo_overlap"       [0]  NCCL transport: run the exchanges on a side stream
     "halo_reserve_sms"   [4]  SMs left free for NCCL kernels in that mode
     "reverse_operator"   [1]  inside fus_model_rk4 the stiffness kernel walks the cells from the
                               last to the first, so that the epilogue after it (which streams the
                               vectors from the front) finds the part of the right-hand side written
                               last still in L2, and the operator after that the tail of the stage
                               input (same results up to the order of the atomic sums)
     "stage_hints"        [0]  epilogue: streaming vectors marked L2 evict-first (measured slower)
     "col_blocks_per_sm"  [0]  cap on resident blocks of the stiffness kernels (0 = occupancy) */
int fus_ctx_set_option(fus_ctx* ctx, const char* name, int value);
/* "geometry_mode" 1 asks for affine compression: if every cell is a parallelepiped the operator
   keeps 6 numbers per CELL and rebuilds G = w_q * Ghat instead of streaming 48 B per point
   (off by default; falls back to streaming when any cell is not affine).
   "geometry_mode" 2 rebuilds |det J| w K K^T at every point from the trilinear cell map (192 B per
   cell, fus_trilinear_coeffs): valid for any mesh with a degree-1 coordinate element, needs a
   context made by fus_ctx_create_from_mesh (FUS_ERR_STATE otherwise).  Value 3 selects the same
   kernel compiled under a 128-register cap: an occupancy experiment, same results.
   fus_ctx_get_option reads back "geometry_compressed" (the mode in use: 0, 1 or 2),
   "stiffness_variant", "halo_mode". */
int fus_ctx_get_option(fus_ctx* ctx, const char* name, int* value);
int fus_ctx_sync(fus_ctx* ctx);

/* Read the device cell data back in the reference layouts (tests; G or detJ may be NULL). */
int fus_ctx_get_geometry(fus_ctx* ctx, double* G, double* detJ);

/* y += K(coeffs) x  -- StiffnessSpectral3D::operator() (spectral_op.hpp:173-243).
   Accumulates into y (caller zero-fills, Linear.hpp:203); loops local cells only; x must hold
   fresh ghosts; coeffs indexed by local cell.  *_dev take device pointers and are stream-ordered;
   *_host take host buffers of ndofs / ncells entries and include the copies. */
int fus_stiffness_apply_dev(fus_ctx* ctx, const double* x, const double* coeffs, double* y);
int fus_stiffness_apply_host(fus_ctx* ctx, const double* x, const double* coeffs, double* y);

/* y += M(coeffs) x  -- MassSpectral3D::operator() (spectral_op.hpp:69-86). */
int fus_mass_apply_dev(fus_ctx* ctx, const double* x, const double* coeffs, double* y);
int fus_mass_apply_host(fus_ctx* ctx, const double* x, const double* coeffs, double* y);

/* The same two operators in FP32 (StiffnessSpectral3D<float,P> / MassSpectral3D<float,P>, the
   scalar type of the reference's tests/test_operators3d/main.cpp:13 and of its float timing runs):
   float vectors and coefficients, float copies of G / detJ made on the first call (24 B of G per
   point instead of 48).  Hexahedral contexts that store G / detJ only; geometry mode 0. */
int fus_stiffness_apply_f32_dev(fus_ctx* ctx, const float* x, const float* coeffs, float* y);
int fus_stiffness_apply_f32_host(fus_ctx* ctx, const float* x, const float* coeffs, float* y);
int fus_mass_apply_f32_dev(fus_ctx* ctx, const float* x, const float* coeffs, float* y);
int fus_mass_apply_f32_host(fus_ctx* ctx, const float* x, const float* coeffs, float* y);

/* Device buffers for callers without a CUDA runtime of their own (C, ctypes). */
int fus_dev_alloc(fus_ctx* ctx, size_t bytes, void** ptr);
int fus_dev_free(fus_ctx* ctx, void* ptr);
int fus_dev_upload(fus_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int fus_dev_download(fus_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int fus_dev_memset(fus_ctx* ctx, void* dst_dev, int value, size_t bytes);

/* ------------------------------------------------------------------------------------------
 * Models: {Linear,Lossy,Westervelt}Spectral3D (Linear.hpp:52-347, Lossy.hpp:54-373,
 * Westervelt.hpp:56-406).  Per-cell material arrays are host arrays of ncells entries; the
 * boundary vectors are the dense outputs of fus_boundary_vectors (NULL = all zero).
 * The lumped mass m = M(1/rho c^2) 1 + bmass is assembled on the device (Linear.hpp:127-134).
 * ---------------------------------------------------------------------------------------- */
int fus_model_create(fus_ctx* ctx, int kind, const double* c0, const double* rho0,
                     const double* delta0, const double* beta0, const double* src,
                     const double* dsrc, const double* absb, const double* bmass, double freq,
                     double p0, double s0, fus_model** out);
int fus_model_destroy(fus_model* m);

/* init() (Linear.hpp:161-164) when u,v are NULL; otherwise sets u_n, v_n from host arrays. */
int fus_model_set_state(fus_model* m, const double* u, const double* v);
/* u_sol() (Linear.hpp:316): copies u_n (and v_n) back; either may be NULL. */
int fus_model_get_state(fus_model* m, double* u, double* v);
/* Device pointers to u_n / v_n (ndofs doubles each). */
int fus_model_state_dev(fus_model* m, double** u, double** v);
/* The assembled lumped mass (m, or m0 for Westervelt), ndofs entries. */
int fus_model_get_mass(fus_model* m, double* mass);

/* kv = f1(t, u, v) for host u, v (Linear.hpp:181-222); a single stage evaluation, for tests. */
int fus_model_f1(fus_model* m, double t, const double* u, const double* v, double* result);

/* rk4(startTime, finalTime, timeStep) (Linear.hpp:228-314): runs entirely on the device from the
   current state; same host-side time arithmetic as the reference loop.  nsteps (may be NULL)
   receives the number of steps taken.  Asynchronous on the context stream until the state is read. */
int fus_model_rk4(fus_model* m, double t0, double tf, double dt, int* nsteps);

/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t fus_launch_count(void);

/* Per-kernel device timing with CUDA events recorded on the context stream around each launch
   (option "profile_kernels" = 1 to start, 0 to stop).  fus_ctx_profile synchronises, then returns
   the number of launches and the summed device time of one kernel family since profiling was
   enabled: "stiffness" (operator), "stage" (fused RK4 epilogue), "boundary". */
int fus_ctx_profile(fus_ctx* ctx, const char* kernel, int64_t* launches, double* total_ms);

/* ------------------------------------------------------------------------------------------
 * 2-D quadrilateral variant: MassSpectral2D / StiffnessSpectral2D and the 2-D solver classes of
 * cpp/fenicsx-sf-naive/common/{spectral_op.hpp:28-107,226-359, Linear.hpp:52-350, Lossy.hpp,
 * Westervelt.hpp}.  Nd = N^2, tensor index i = i0*N + i1, G[c][q][3] = {G00,G01,G11} |detJ| w_q
 * (precompute.hpp:199-203 there), 4 vertices per cell in DOLFINx order v = a + 2b, vertex
 * coordinates padded to 3 doubles, local facets (edges) 0: y=0, 1: x=0, 2: x=1, 3: y=1.
 * A context made by fus_ctx_create_2d / fus_ctx_create_from_mesh_2d works with every operator,
 * model and halo entry point above (fus_stiffness_apply_*, fus_mass_apply_*, fus_model_*, ...);
 * the boundary vectors come from fus_boundary_vectors_2d.  "geometry_mode" stays 0.
 * ---------------------------------------------------------------------------------------- */
int fus_rect_mesh(const int n[2], const double lo[2], const double hi[2], double* xg,
                  int32_t* xdofmap);
int fus_rect_dofmap(int P, const int n[2], int32_t* tensor_dofmap); /* lexicographic numbering */
int64_t fus_rect_num_dofs(int P, const int n[2]);
/* exterior edges {cell, local facet, tag}: tag 1 on x=lo, 2 on x=hi, 0 elsewhere */
int64_t fus_rect_facets(const int n[2], int32_t* facets);
/* As fus_boundary_vectors, with the facet sets of the 2-D forms of cpp/fenicsx-sf-naive
   (examples/lossy_planewave2d_1/forms.py:37-42, westervelt_planewave2d_1/forms.py:37-42): the
   absorbing term and its mass-like counterpart over ds(2) for EVERY model kind
     absb  S^(1/rho c)(tag 2)     bmass  S^(delta/rho c^3)(tag 2)   [lossy, westervelt]
   -- the 3-D forms of cpp/fenicsx-sf use `ds` without an id for the lossy and Westervelt models. */
int fus_boundary_vectors_2d(int kind, int P, int64_t ncells, int64_t ndofs, const double* xg,
                            const int32_t* xdofmap, const int32_t* tensor_dofmap, int64_t nfacets,
                            const int32_t* facets, const double* c0, const double* rho0,
                            const double* delta0, double* src, double* dsrc, double* absb,
                            double* bmass);
int fus_ctx_create_2d(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                      const int32_t* tensor_dofmap, const double* G, const double* detJ,
                      const double* dphi, int device, fus_ctx** out);
int fus_ctx_create_from_mesh_2d(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                                const int32_t* tensor_dofmap, int64_t nverts, const double* xg,
                                const int32_t* xdofmap, int device, fus_ctx** out);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU halo exchange: replaces la::Vector::scatter_fwd / scatter_rev(std::plus)
 * (call sites Linear.hpp:196,199,206; Westervelt.hpp:243,246,257,265) with NCCL send/recv.
 * One context per GPU/process.  Neighbour k exchanges send_idx[send_off[k]:send_off[k+1]] (local
 * indices of OWNED dofs that are ghosts on rank neigh[k]) and recv_idx[recv_off[k]:...] (local
 * indices of GHOST dofs owned by neigh[k]), in matching order on both sides.
 * `nccl_unique_id` is the 128-byte ncclUniqueId made by fus_comm_unique_id on rank 0 and
 * distributed by the caller (e.g. torch.distributed broadcast).
 * `ninterface_cells`: cells [0, ninterface_cells) touch shared dofs and are applied first so that
 * the reverse exchange overlaps with the remaining interior cells.
 * ---------------------------------------------------------------------------------------- */
int fus_comm_unique_id(void* id128);
int fus_halo_setup(fus_ctx* ctx, int rank, int nranks, const void* nccl_unique_id, int nneigh,
                   const int* neigh, const int64_t* send_off, const int32_t* send_idx,
                   const int64_t* recv_off, const int32_t* recv_idx, int64_t ninterface_cells);
/* Optional FUSED PEER transport for the exchanges inside fus_model_rk4 (after fus_halo_setup).
 * There is no exchange kernel: the RK4 epilogue adds the neighbours' partial sums of the right-hand
 * side for the dofs it shares (scatter_rev) and stores the next stage input straight into the
 * neighbours' mailboxes over NVLink peer memory (scatter_fwd); the stiffness kernel gathers ghost
 * values from the mailbox and, in a launch of its own over the cells that touch shared dofs, ships
 * its ghost partial sums to the owners' mailboxes after its last cell, while the launch over the
 * other cells runs.  Flags with device-side sequence numbers order everything, so a captured CUDA
 * graph replays whole steps; every wait on a neighbour is bounded (FUS_HALO_TIMEOUT_S, default 30 s)
 * and a time-out aborts the run: the remaining kernels return at once and fus_ctx_sync /
 * fus_model_get_state report FUS_ERR_COMM.  fus_model_rk4 begins with a handshake between
 * neighbours, so ranks may enter it at different times (as with the reference's MPI scatters).
 * Requirements on the local numbering (fus_box_partition_* and the Python partitioners provide it;
 * FUS_ERR_UNSUPPORTED otherwise, and the context stays on NCCL): the owned dofs that appear in the
 * send lists are exactly [0, nshared), and recv_idx = nowned + 0, 1, 2, ...
 * Set-up: export this rank's mailbox -- a 64-byte cudaIpcMemHandle_t (other processes) and/or its
 * device pointer (other contexts of this process), plus layout6 = byte offsets {fwd_v, rev, forward
 * flags, reverse flags, ready flags, total size} -- distribute it together with the offset tables,
 * then connect with one handle (or pointer + device) per neighbour and byte_off[nneigh][6] from
 * fus_halo_peer_offsets. */
int fus_halo_peer_export(fus_ctx* ctx, void* ipc_handle64, int64_t* layout6, void** base);
/* Byte layout of a rank's mailbox for given list sizes (host arithmetic, no device). */
int fus_halo_mailbox_layout(int64_t nsend, int64_t nrecv, int nneigh, int64_t* layout6);
/* Where this rank's data go inside neighbour q's mailbox: out6 = byte offsets of {forward-u run,
 * forward-v run, reverse run, forward flag, reverse flag, ready flag}, from q's layout6, q's offset
 * tables and j = this rank's position in q's neighbour list. */
int fus_halo_peer_offsets(const int64_t* q_layout6, const int64_t* q_send_off,
                          const int64_t* q_recv_off, int j, int64_t* out6);
int fus_halo_peer_connect(fus_ctx* ctx, const void* handles, const int64_t* byte_off);
/* The same between contexts of ONE process (one host thread per GPU): bases[k] / devices[k] are
 * neighbour k's mailbox pointer (from its fus_halo_peer_export) and CUDA device. */
int fus_halo_peer_connect_local(fus_ctx* ctx, void* const* bases, const int* devices,
                                const int64_t* byte_off);

/* Partition of the structured box over a pgrid[0] x pgrid[1] x pgrid[2] process grid for rank
 * `rank` (host, once per rank): what DOLFINx's mesh partitioner and common::IndexMap provide to the
 * reference.  Cells are split in contiguous blocks without ghost cells; an interface node is owned
 * by the block with the lowest grid coordinates sharing it; local numbering is owned entries first,
 * (those that a neighbour ghosts before the others), then ghosts grouped by owner rank; cells
 * touching a shared dof come first.
 * fus_box_partition_info: sizes = {ncells, ndofs, nowned, nfacets, nneigh, nsend, nrecv,
 * ninterface_cells, ndofs_global}.  fus_box_partition_arrays copies out (any pointer may be NULL):
 * dofmap[ncells][Nd], xdofmap[ncells][8] (vertex numbering of fus_box_mesh on the local block),
 * cell_global[ncells], global_key[ndofs] (global lexicographic node id of every local dof),
 * facets[nfacets][3], neigh[nneigh], send_off[nneigh+1], send_idx[nsend], recv_off[nneigh+1],
 * recv_idx[nrecv] -- the arguments of fus_halo_setup. */
typedef struct fus_partition fus_partition;
int fus_box_partition_create(int P, const int n_global[3], const int pgrid[3], int rank,
                             int numbering, fus_partition** out);
int fus_box_partition_info(const fus_partition* p, int64_t sizes[9], int32_t n_local[3],
                           int32_t cell_lo[3]);
int fus_box_partition_arrays(const fus_partition* p, int32_t* dofmap, int32_t* xdofmap,
                             int64_t* cell_global, int64_t* global_key, int32_t* facets,
                             int32_t* neigh, int64_t* send_off, int32_t* send_idx,
                             int64_t* recv_off, int32_t* recv_idx);
int fus_box_partition_destroy(fus_partition* p);

/* Stand-alone collectives on device vectors (tests): owner -> ghost, ghost -> owner (+=). */
int fus_scatter_fwd_dev(fus_ctx* ctx, double* x);
int fus_scatter_rev_dev(fus_ctx* ctx, double* x);

#ifdef __cplusplus
}
#endif
#endif /* FUS_B200_H */
