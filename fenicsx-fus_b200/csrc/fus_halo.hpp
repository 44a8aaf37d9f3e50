// fus_halo.hpp -- NCCL send/recv halo exchange replacing DOLFINx's la::Vector::scatter_fwd /
// scatter_rev(std::plus) (call sites Linear.hpp:196,199,206; Westervelt.hpp:243-265).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "fus_halo_kernels.cuh"

namespace fus {

struct Halo;

int halo_unique_id(void* id128);
int halo_create(Halo** out, int device, int rank, int nranks, const void* uid, int nneigh,
                const int* neigh, const int64_t* send_off, const int32_t* send_idx,
                const int64_t* recv_off, const int32_t* recv_idx, int64_t nowned, int64_t ndofs,
                int64_t ninterface_cells);
void halo_destroy(Halo* h);
void halo_set_overlap(Halo* h, int on);
long long halo_interface_cells(const Halo* h);

// owner -> ghost (insert) for one or two vectors in a single message per neighbour
int halo_forward(Halo* h, double* a, double* b, cudaStream_t st);
// ghost -> owner (add) for one or two vectors; complete in stream order on `st`
int halo_reverse(Halo* h, double* a, double* b, cudaStream_t st);
// split forward update: begin after the producer of a/b on `st`; the exchange runs on the halo's
// own high-priority stream while `st` continues with cells that touch no shared dof; end joins.
int halo_forward_begin(Halo* h, double* a, double* b, cudaStream_t st);
int halo_forward_end(Halo* h, double* a, double* b, cudaStream_t st);
int halo_overlap(const Halo* h);
// modes: 0 NCCL in stream order, 1 NCCL on a side stream (overlapped), 2 fused peer transport (the
// stage kernels of fus_model_rk4 exchange over NVLink peer memory themselves; everything else --
// scatter_fwd / scatter_rev entry points, set-up reductions, f1 -- stays on NCCL)
int halo_mode(const Halo* h);
// byte layout of a rank's mailbox (host arithmetic only): {fwd_v, rev, fwd flags, rev flags, ready
// flags, total}
void halo_mailbox_layout(int64_t nsend, int64_t nrecv, int nneigh, int64_t* layout6);
// fused peer transport: export this rank's mailbox (IPC handle and/or device pointer), then connect
// to the neighbours' mailboxes (other processes: IPC handles; same process: device pointers)
int halo_peer_export(Halo* h, void* ipc_handle64, int64_t* layout6, void** base);
int halo_peer_connect(Halo* h, const void* handles, const int64_t* byte_off);
int halo_peer_connect_local(Halo* h, void* const* bases, const int* devices, const int64_t* byte_off);
// 0 when no wait on a neighbour has timed out since creation
int halo_peer_error(Halo* h);
const FusedHalo* halo_fused(const Halo* h);   // device pointer for the stage kernels, or nullptr
HaloLaunch halo_fused_launch(const Halo* h);   // by-value kernel parameter of the interface launch
long long halo_fused_interface_cells(const Halo* h);
int halo_fused_operator_skipped(Halo* h, cudaStream_t st);
int halo_fused_entry(Halo* h, const double* u, const double* v, cudaStream_t st);
int halo_fused_exit(Halo* h, double* u, double* v, cudaStream_t st);
// split form: begin after the interface cells have been applied on `st`; the exchange runs on the
// halo's own stream while `st` continues with interior cells; end joins it back into `st`.
int halo_reverse_begin(Halo* h, double* a, cudaStream_t st);
int halo_reverse_end(Halo* h, double* a, cudaStream_t st);

} // namespace fus
