// common.cuh -- shared device helpers for the octane_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace octane {

// One pyramid level as stored on this rank: global size nx x ny, local rows
// [j0, j0+rows) with row pitch `pitch` floats (multiple of 32 -> every row
// starts on a 128-byte line and float4 accesses are aligned).  Channel planes
// are `plane` floats apart.  Single-GPU runs have j0 = 0, rows = ny.
struct Geom {
    int nx, ny;
    int pitch;
    int j0, rows;
    long long plane;
    __host__ __device__ inline size_t at(int i, int j) const { return (size_t)(j - j0) * pitch + i; }
    __host__ __device__ inline int jlo() const { return j0; }
    __host__ __device__ inline int jhi() const { return j0 + rows; }
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Function attributes (dynamic shared-memory limits) are per device: `mask` remembers the devices a launch
// site has already configured (one bit per device; setting an attribute twice is harmless, so a race
// between host threads is benign).  Returns true the first time the current device is seen.
inline bool first_launch_on_device(unsigned long long* mask)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (*mask & bit) return false;
    *mask |= bit;
    return true;
}

// Device-side scalars of one PCG solve (reference: rkTzk, pkTApk, zktrk,
// z0tr0, residc -- managed floats, src/oct_variational_optical_flow.cu:1319-1328).
struct PcgScalars {
    float rz_old;   // z_{k-1}.r_{k-1}  (z0tr0)
    float rz;       // z_k.r_k          (rkTzk == zktrk)
    float pAp;      // p.Ap             (pkTApk)
    float rr;       // r.r              (residc)
    float alpha;    // rz/pAp of the last executed iteration (x += alpha p is applied one pass later)
    float tol;
    int done;       // stop rule satisfied: !(rr > tol)
    int its;        // iterations executed in this solve
    int halo_err;   // banded runs: warp left the local rows
    int comm_err;   // banded runs: a peer's contribution did not arrive in time
    // merged-reduction solver (pcg_fused.cu): scalar recurrences in double, narrowed once for the vector updates
    double d_gamma; // r.z of the current residual
    double d_alpha; // step of the NEXT launch
    float f_alpha, f_beta;   // what the next launch applies: x += alpha p, p = z + beta p
    float f_alpha_prev;      // alpha of the launch before (its x term is applied every second launch)
};

// ---- peer-memory exchange between the row bands (one process per GPU) -----------------------
// Every rank owns one small window in device memory that all its peers have mapped (CUDA IPC over
// NVLink).  A grid-wide dot product becomes a cross-GPU one inside the same kernel: the last block
// stores this rank's partial straight into every peer's window, releases a sequence number, waits
// for the peers' numbers in its own window and sums the partials in rank order -- the same bits on
// every rank, so the bands take the same stop decision without a host or an NCCL call in between.
// Phases alternate (build / pass 1 / pass 2) and values are double-buffered on the epoch's parity,
// so a rank that runs ahead can never overwrite a value a slower peer still has to read.
constexpr int P2P_MAXW = 16;
enum { P2P_BUILD = 0, P2P_PASS1 = 1, P2P_PASS2 = 2, P2P_NPHASE = 3 };
struct P2PWindow {
    double val[P2P_NPHASE][2][P2P_MAXW][8];    // up to 8 sums per phase (the merged-reduction solver exchanges 6)
    unsigned seq[P2P_NPHASE][2][P2P_MAXW];
};
struct P2P {
    int world, rank;              // world <= 1: disabled
    P2PWindow* const* peers;      // device array [world]: every rank's window as mapped here (peers[rank] = own)
    unsigned* epoch;              // device [P2P_NPHASE], advanced by the publishing block
};

// Scope of the fences / release-acquire pairs that order stores into a peer GPU's memory.  System scope is what
// the PTX memory model requires between two GPUs.  OCTANE_PEER_SCOPE_GPU is a timing experiment only (DESIGN.md
// section 5: what a system-scope fence costs while a device-to-host copy is in flight); never ship it.
#ifdef OCTANE_PEER_SCOPE_GPU
#define OCTANE_FENCE_PEER() __threadfence()
#define OCTANE_PEER_SCOPE "gpu"
#else
#define OCTANE_FENCE_PEER() __threadfence_system()
#define OCTANE_PEER_SCOPE "sys"
#endif
// A thread's stores into a peer GPU's memory (the halo rows of r) are not fenced one by one: every block ends in a
// barrier (block_sum) followed by ONE system-scope fence of its thread 0 before the ticket (grid_sum_finish), which
// is cumulative over the stores the barrier ordered before it, and the publishing block releases at system scope.
// (Per-store fences cost 13 ms of a 1.1 s banded pair while a device-to-host copy was in flight, call 19.)
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release." OCTANE_PEER_SCOPE ".global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire." OCTANE_PEER_SCOPE ".global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v)
{
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// clamp-to-edge on a global index (oct_bc_cu, :26-41, on integer-valued floats)
__device__ __forceinline__ int clampi(int x, int n) { return x < 0 ? 0 : (x >= n ? n - 1 : x); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV doubles (fixed order -> deterministic). blockDim.x*blockDim.y
// must be a multiple of 32 and <= 1024.  Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* NV*32 */)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) smem[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double t = (lane < nwarps) ? smem[k * 32 + lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

// Grid-wide deterministic reduction, second stage: every block has written its
// NV partials to partials[k*nblocks + block]; the last block to arrive (ticket)
// sums them in a fixed order and returns true in ALL its threads with the totals
// in out[] (valid in thread 0).
template <int NV>
__device__ __forceinline__ bool grid_sum_finish(const double (&mine)[NV], double* partials, unsigned* ticket,
                                                double (&out)[NV], double* smem, bool sys_fence = false)
{
    __shared__ bool is_last;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const unsigned nblocks = gridDim.x * gridDim.y;
    const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) partials[(size_t)k * nblocks + bid] = mine[k];
        if (sys_fence) OCTANE_FENCE_PEER();      // the block stored into a peer GPU's memory
        else __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == nblocks - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double s = 0.0;
        for (unsigned b = tid; b < nblocks; b += nthreads) s += __ldcg(&partials[(size_t)k * nblocks + b]);
        out[k] = s;
    }
    __syncthreads();
    block_sum<NV>(out, smem);
    if (tid == 0) *ticket = 0u;
    return true;
}


// Cross-rank sum of NV doubles held by thread 0 of the calling block (the last block of a grid, after
// grid_sum_finish).  All threads of the block must call it; the totals come back in thread 0.  A peer
// that does not answer within 10 s sets *comm_err and the call returns what it has (the host turns
// the flag into OCTANE_ECOMM; nothing spins for ever).
template <int NV>
__device__ __forceinline__ void p2p_allreduce(const P2P& c, int phase, double (&tot)[NV], int* comm_err)
{
    __shared__ double s_val[NV];
    __shared__ unsigned s_epoch;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) {
        const unsigned e = c.epoch[phase] + 1u;
        c.epoch[phase] = e;
        s_epoch = e;
#pragma unroll
        for (int k = 0; k < NV; k++) s_val[k] = tot[k];
    }
    __syncthreads();
    const unsigned e = s_epoch;
    const int par = (int)(e & 1u);
    P2PWindow* mine = c.peers[c.rank];
    if (tid < c.world) {
        P2PWindow* w = c.peers[tid];
#pragma unroll
        for (int k = 0; k < NV; k++) st_relaxed_sys(&w->val[phase][par][c.rank][k], s_val[k]);
        st_release_sys(&w->seq[phase][par][c.rank], e);
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&mine->seq[phase][par][tid]) != e) {
            if (global_ns() - t0 > 10000000000ull) { atomicExch(comm_err, 1); break; }
        }
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = 0.0;
            for (int r = 0; r < c.world; r++) s += ld_relaxed_sys(&mine->val[phase][par][r][k]);
            tot[k] = s;
        }
    }
}

}  // namespace octane
