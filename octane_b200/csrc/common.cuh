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
};

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
                                                double (&out)[NV], double* smem)
{
    __shared__ bool is_last;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const unsigned nblocks = gridDim.x * gridDim.y;
    const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) partials[(size_t)k * nblocks + bid] = mine[k];
        __threadfence();
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

}  // namespace octane
