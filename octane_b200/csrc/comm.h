// comm.h -- row-band halo exchange and scalar all-reduce over NCCL (NVLink 5 /
// NVSwitch).  The reference has no multi-GPU layer at all (SURVEY.md 2,
// "Parallelism strategies": none), so there is no reference file to cite; the
// exchange steps are the ones SURVEY.md 8(e) derives from the stencil radii.
// NCCL is loaded with dlopen so a single-GPU build has no NCCL dependency and a
// process that already holds torch's bundled libnccl shares that copy.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace octane {

struct Comm {
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    // peer-memory path (CUDA IPC over NVLink; common.cuh: P2PWindow).  NCCL stays for the bootstrap
    // and for the per-solve u,v halo rows; the per-iteration dots and r rows go through these.
    bool p2p = false;
    void* window = nullptr;              // this rank's window (device memory, zeroed)
    void* peer_window[16] = { nullptr }; // every rank's window as mapped here ([rank] = window)
    void** d_peers = nullptr;            // device copy of peer_window
    unsigned* d_epoch = nullptr;         // device [3]
    void* nb_arena[2] = { nullptr, nullptr };   // rank-1 / rank+1 workspace as mapped here
    char* d_stage = nullptr;             // staging for the small host-side exchanges (allocated once: a re-plan under
                                         // memory pressure must not fail to allocate it on one rank only)
};

int comm_unique_id(char id[128]);
int comm_init(Comm* c, const char id[128], int rank, int world);
void comm_destroy(Comm* c);
// in-place sum of n doubles on the stream
int comm_allreduce_f64(Comm* c, double* d_buf, int n, cudaStream_t st);
// One exchange step for several planes at once: for each plane p, send `count`
// floats starting at send_up[p] to rank-1 and send_dn[p] to rank+1, receive the
// neighbours' counterparts into recv_up[p] (from rank-1) and recv_dn[p] (from rank+1).
int comm_halo_exchange(Comm* c, int nplanes, float* const* send_up, float* const* recv_up,
                       float* const* send_dn, float* const* recv_dn, size_t count, cudaStream_t st);
// After comm_init: allocate the window, exchange IPC handles (ncclAllGather), map the peers.  Returns 0
// and sets c->p2p = true when every rank succeeded; 0 with p2p = false when any rank could not map
// its peers (the NCCL path is used); -1 on a hard error.
int comm_p2p_init(Comm* c, size_t window_bytes, cudaStream_t st);
// Collective, call when the workspace (re)appears: unmap the neighbours' old workspaces
// (comm_p2p_unmap_arenas, before anybody frees), then map the new ones.
int comm_p2p_unmap_arenas(Comm* c, cudaStream_t st);
int comm_p2p_map_arenas(Comm* c, void* my_arena, int local_ok, cudaStream_t st);
const char* comm_last_error();

}  // namespace octane
