// comm.cc -- see comm.h
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace octane {

namespace {
struct Api {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Api api;
thread_local char errbuf[512] = "";      // per thread, like ctx.cu's g_err: two contexts on two host threads do not share it

bool load()
{
    if (api.h) return true;
    // RTLD_NOLOAD first: reuse a copy the host framework (torch) already mapped
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) break; }
    if (!h) {
        const char* env = getenv("OCTANE_NCCL_LIB");
        if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    for (const char* n : names) { if (h) break; h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
    if (!h) { snprintf(errbuf, sizeof errbuf, "cannot dlopen libnccl: %s", dlerror()); return false; }
#define SYM(field, name) \
    *(void**)(&api.field) = dlsym(h, name); \
    if (!api.field) { snprintf(errbuf, sizeof errbuf, "libnccl lacks %s", name); return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    api.h = h;
    return true;
}

int check(ncclResult_t r, const char* what)
{
    if (r == ncclSuccess) return 0;
    snprintf(errbuf, sizeof errbuf, "%s: %s", what, api.GetErrorString ? api.GetErrorString(r) : "nccl error");
    return -1;
}
}  // namespace

const char* comm_last_error() { return errbuf; }

int comm_unique_id(char id[128])
{
    if (!load()) return -1;
    ncclUniqueId u;
    static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
    if (check(api.GetUniqueId(&u), "ncclGetUniqueId")) return -1;
    memcpy(id, &u, 128);
    return 0;
}

int comm_init(Comm* c, const char id[128], int rank, int world)
{
    if (!load()) return -1;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t comm;
    if (check(api.CommInitRank(&comm, world, u, rank), "ncclCommInitRank")) return -1;
    c->nccl_comm = comm;
    c->rank = rank;
    c->world = world;
    return 0;
}

namespace {

// every rank contributes `bytes` bytes; out (host) receives world * bytes
int allgather_bytes(Comm* c, const void* mine, size_t bytes, void* out, cudaStream_t st)
{
    const size_t stage_bytes = 17 * sizeof(cudaIpcMemHandle_t);     // the largest item exchanged, 16 ranks + own
    if (bytes * (c->world + 1) > stage_bytes) { snprintf(errbuf, sizeof errbuf, "allgather item too large"); return -1; }
    if (!c->d_stage && cudaMalloc(&c->d_stage, stage_bytes) != cudaSuccess) { snprintf(errbuf, sizeof errbuf, "cudaMalloc (allgather)"); return -1; }
    char* d = c->d_stage;
    int rc = 0;
    if (cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (!rc) rc = check(api.AllGather(d, d + bytes, bytes, ncclChar, (ncclComm_t)c->nccl_comm, st), "ncclAllGather");
    if (!rc && cudaMemcpyAsync(out, d + bytes, bytes * c->world, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = -1;
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = -1;
    if (rc && !errbuf[0]) snprintf(errbuf, sizeof errbuf, "allgather of IPC handles failed");
    return rc;
}

// min over ranks of a flag: do all ranks agree that a step worked?
int all_ok(Comm* c, int ok, cudaStream_t st)
{
    std::vector<int> flags(c->world, 0);
    if (allgather_bytes(c, &ok, sizeof ok, flags.data(), st)) return -1;
    for (int f : flags) if (!f) return 0;
    return 1;
}

}  // namespace

int comm_p2p_init(Comm* c, size_t window_bytes, cudaStream_t st)
{
    c->p2p = false;
    if (c->world <= 1 || c->world > 16) return 0;
    const char* mode = getenv("OCTANE_COMM");
    if (mode && !strcmp(mode, "nccl")) return 0;           // developer switch: per-iteration exchanges over NCCL
    int ok = 1;
    if (cudaMalloc(&c->window, window_bytes) != cudaSuccess || cudaMemset(c->window, 0, window_bytes) != cudaSuccess) ok = 0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && cudaIpcGetMemHandle(&mine, c->window) != cudaSuccess) ok = 0;
    std::vector<cudaIpcMemHandle_t> all(c->world);
    if (allgather_bytes(c, &mine, sizeof mine, all.data(), st)) return -1;
    for (int r = 0; r < c->world && ok; r++) {
        if (r == c->rank) { c->peer_window[r] = c->window; continue; }
        if (cudaIpcOpenMemHandle(&c->peer_window[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            c->peer_window[r] = nullptr;
            ok = 0;
        }
    }
    if (ok) {
        if (cudaMalloc(&c->d_peers, sizeof(void*) * 16) != cudaSuccess ||
            cudaMemcpy(c->d_peers, c->peer_window, sizeof(void*) * 16, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(&c->d_epoch, 4 * sizeof(unsigned)) != cudaSuccess ||
            cudaMemset(c->d_epoch, 0, 4 * sizeof(unsigned)) != cudaSuccess)
            ok = 0;
    }
    const int agreed = all_ok(c, ok, st);                  // also a barrier: every window is zeroed and mapped
    if (agreed < 0) return -1;
    c->p2p = agreed == 1;
    return 0;
}

int comm_p2p_unmap_arenas(Comm* c, cudaStream_t st)
{
    if (!c->p2p) return 0;
    for (int k = 0; k < 2; k++) {
        if (c->nb_arena[k]) cudaIpcCloseMemHandle(c->nb_arena[k]);
        c->nb_arena[k] = nullptr;
    }
    return all_ok(c, 1, st) < 0 ? -1 : 0;                  // barrier: nobody frees a workspace a peer still maps
}

int comm_p2p_map_arenas(Comm* c, void* my_arena, int local_ok, cudaStream_t st)
{
    if (!c->p2p) return 0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    // a rank whose own preparation failed (local_ok == 0) still takes part in both exchanges
    // and votes "no": its peers get an error instead of waiting for it forever
    int ok = local_ok && my_arena && cudaIpcGetMemHandle(&mine, my_arena) == cudaSuccess;
    std::vector<cudaIpcMemHandle_t> all(c->world);
    if (allgather_bytes(c, &mine, sizeof mine, all.data(), st)) return -1;
    const int nb[2] = { c->rank - 1, c->rank + 1 };
    for (int k = 0; k < 2 && ok; k++) {
        if (nb[k] < 0 || nb[k] >= c->world) continue;
        if (cudaIpcOpenMemHandle(&c->nb_arena[k], all[nb[k]], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            c->nb_arena[k] = nullptr;
            ok = 0;
        }
    }
    const int agreed = all_ok(c, ok, st);
    if (agreed < 0) return -1;
    if (agreed == 0) { snprintf(errbuf, sizeof errbuf, "a rank could not map its neighbour's workspace (CUDA IPC)"); return -1; }
    return 0;
}

void comm_destroy(Comm* c)
{
    for (int k = 0; k < 2; k++) { if (c->nb_arena[k]) cudaIpcCloseMemHandle(c->nb_arena[k]); c->nb_arena[k] = nullptr; }
    for (int r = 0; r < 16; r++) {
        if (c->peer_window[r] && r != c->rank) cudaIpcCloseMemHandle(c->peer_window[r]);
        c->peer_window[r] = nullptr;
    }
    if (c->window) cudaFree(c->window);
    if (c->d_peers) cudaFree(c->d_peers);
    if (c->d_epoch) cudaFree(c->d_epoch);
    if (c->d_stage) cudaFree(c->d_stage);
    c->d_stage = nullptr;
    c->window = nullptr; c->d_peers = nullptr; c->d_epoch = nullptr; c->p2p = false;
    if (c->nccl_comm && api.CommDestroy) api.CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
    c->world = 1;
    c->rank = 0;
}

int comm_allreduce_f64(Comm* c, double* d_buf, int n, cudaStream_t st)
{
    if (c->world <= 1) return 0;
    return check(api.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)c->nccl_comm, st),
                 "ncclAllReduce");
}

int comm_halo_exchange(Comm* c, int nplanes, float* const* send_up, float* const* recv_up,
                       float* const* send_dn, float* const* recv_dn, size_t count, cudaStream_t st)
{
    if (c->world <= 1 || count == 0) return 0;
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    const int up = c->rank - 1, dn = c->rank + 1;
    if (check(api.GroupStart(), "ncclGroupStart")) return -1;
    int rc = 0;
    for (int p = 0; p < nplanes && !rc; p++) {
        if (up >= 0) {
            rc = check(api.Send(send_up[p], count, ncclFloat32, up, comm, st), "ncclSend");
            if (!rc) rc = check(api.Recv(recv_up[p], count, ncclFloat32, up, comm, st), "ncclRecv");
        }
        if (!rc && dn < c->world) {
            rc = check(api.Send(send_dn[p], count, ncclFloat32, dn, comm, st), "ncclSend");
            if (!rc) rc = check(api.Recv(recv_dn[p], count, ncclFloat32, dn, comm, st), "ncclRecv");
        }
    }
    // the group is always closed, also after a failed call inside it (the first error is the one reported)
    char first[sizeof errbuf];
    memcpy(first, errbuf, sizeof errbuf);
    const int rc_end = check(api.GroupEnd(), "ncclGroupEnd");
    if (rc) { memcpy(errbuf, first, sizeof errbuf); return -1; }
    return rc_end;
}

}  // namespace octane
