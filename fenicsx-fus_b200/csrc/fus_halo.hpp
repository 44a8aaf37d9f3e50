// fus_halo.hpp -- NCCL send/recv halo exchange replacing DOLFINx's la::Vector::scatter_fwd /
// scatter_rev(std::plus) (call sites Linear.hpp:196,199,206; Westervelt.hpp:243-265).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

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
// modes: 0 NCCL in stream order, 1 NCCL on a side stream (overlapped), 2 peer-direct one-sided puts
int halo_mode(const Halo* h);
// byte layout of a rank's mailbox (host arithmetic only)
void halo_mailbox_layout(int64_t nsend, int64_t nrecv, int nneigh, int64_t* layout4);
// peer-direct transport: export this rank's mailbox, then connect to the neighbours' mailboxes
int halo_peer_export(Halo* h, void* ipc_handle64, int64_t* layout3);
int halo_peer_connect(Halo* h, const void* handles, const int64_t* byte_off);
// 0 when no peer wait has timed out since creation
int halo_peer_error(Halo* h);
// split form: begin after the interface cells have been applied on `st`; the exchange runs on the
// halo's own stream while `st` continues with interior cells; end joins it back into `st`.
int halo_reverse_begin(Halo* h, double* a, cudaStream_t st);
int halo_reverse_end(Halo* h, double* a, cudaStream_t st);

} // namespace fus
