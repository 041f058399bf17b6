// comm.cc -- see comm.h
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Api api;
char errbuf[512] = "";

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
    SYM(AllReduce, "ncclAllReduce") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
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

void comm_destroy(Comm* c)
{
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
    for (int p = 0; p < nplanes; p++) {
        if (up >= 0) {
            if (check(api.Send(send_up[p], count, ncclFloat32, up, comm, st), "ncclSend")) return -1;
            if (check(api.Recv(recv_up[p], count, ncclFloat32, up, comm, st), "ncclRecv")) return -1;
        }
        if (dn < c->world) {
            if (check(api.Send(send_dn[p], count, ncclFloat32, dn, comm, st), "ncclSend")) return -1;
            if (check(api.Recv(recv_dn[p], count, ncclFloat32, dn, comm, st), "ncclRecv")) return -1;
        }
    }
    return check(api.GroupEnd(), "ncclGroupEnd");
}

}  // namespace octane
