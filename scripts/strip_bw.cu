// strip_bw.cu -- developer microbenchmark (not part of the product): bandwidth of the PCG
// pass-1 access pattern (persistent CTAs walking 1024-pixel strips of many planes through a
// cp.async.bulk ring) as a function of scene size and plane layout.
//   layout 0: row-major planes (the product's layout): elem (i,j) of plane k at k*P + j*pitch + i
//   layout 1: strip-major planes: k*P + (strip*ny + j)*SW + (i - strip*SW)
//   layout 2: planes interleaved by row: (j*NARR + k)*pitch + i
//   layout 3: row-major planes, every second plane of a twice larger arena (address span x2)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o strip_bw strip_bw.cu
// run:   ./strip_bw nx ny layout [nin nout rs]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int SW = 1024, NSTAGE = 4, MAXIN = 11, CONSUMERS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Args {
    float* base;
    size_t P;          // plane stride (floats), layouts 0 and 1
    int nx, ny, pitch, layout, nin, nout, rs, nstrips, nsegs, narr;
};

__device__ __forceinline__ size_t addr(const Args& a, int k, int strip, int j)
{
    if (a.layout == 0) return (size_t)k * a.P + (size_t)j * a.pitch + (size_t)strip * SW;
    if (a.layout == 3) return (size_t)(2 * k) * a.P + (size_t)j * a.pitch + (size_t)strip * SW;
    if (a.layout == 1) return (size_t)k * a.P + ((size_t)strip * a.ny + j) * SW;
    return ((size_t)j * a.narr + k) * a.pitch + (size_t)strip * SW;
}

__global__ void __launch_bounds__(CONSUMERS + 32, 1) k_strip(Args a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE];
    float* stages = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ntasks = a.nstrips * a.nsegs;
    if (tid >= CONSUMERS) {
        if (tid == CONSUMERS) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
                const int seg = t / a.nstrips, strip = t - seg * a.nstrips;
                const int w = min(SW, a.pitch - strip * SW);
                const int j_a = seg * a.rs, j_b = min(a.ny, j_a + a.rs);
                for (int j = j_a; j < j_b; j++, it++) {
                    const int stg = it % NSTAGE;
                    mbar_wait(&empty_bar[stg], ((it / NSTAGE) & 1u) ^ 1u);
                    float* st = stages + (size_t)stg * MAXIN * SW;
                    mbar_expect_tx(&full_bar[stg], (uint32_t)a.nin * w * 4u);
                    for (int k = 0; k < a.nin; k++) bulk_g2s(st + k * SW, a.base + addr(a, k, strip, j), w * 4u, &full_bar[stg]);
                }
            }
        }
    } else {
        const int lane = tid & 31;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
            const int seg = t / a.nstrips, strip = t - seg * a.nstrips;
            const int w = min(SW, a.pitch - strip * SW);
            const int j_a = seg * a.rs, j_b = min(a.ny, j_a + a.rs);
            for (int j = j_a; j < j_b; j++, it++) {
                const int stg = it % NSTAGE;
                mbar_wait(&full_bar[stg], (it / NSTAGE) & 1u);
                const float* st = stages + (size_t)stg * MAXIN * SW;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tid * 4 < w) {
                    for (int k = 0; k < a.nin; k++) {
                        const float4 v = *reinterpret_cast<const float4*>(st + k * SW + tid * 4);
                        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                    }
                    for (int k = 0; k < a.nout; k++)
                        *reinterpret_cast<float4*>(a.base + addr(a, a.nin + k, strip, j) + tid * 4) = s;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stg]);
            }
        }
    }
}

int main(int argc, char** argv)
{
    if (argc < 4) { printf("usage: strip_bw nx ny layout [nin nout rs]\n"); return 1; }
    Args a;
    a.nx = atoi(argv[1]); a.ny = atoi(argv[2]); a.layout = atoi(argv[3]);
    a.nin = argc > 4 ? atoi(argv[4]) : 11; a.nout = argc > 5 ? atoi(argv[5]) : 6;
    a.narr = 17;                                   // the product's live planes during a solve
    if (a.nin > MAXIN || a.nin + a.nout > a.narr) return 1;
    a.nstrips = (a.nx + SW - 1) / SW;
    a.pitch = a.layout == 1 ? a.nstrips * SW : (a.nx + 31) / 32 * 32;
    a.P = (size_t)a.pitch * a.ny;
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int rs = argc > 6 ? atoi(argv[6]) : 0;
    if (rs <= 0) {                                  // same rule as launch_pcg_pass1_tma
        double best = 1e30;
        for (int r = 24; r <= 256; r++) {
            long long tasks = (long long)((a.ny + r - 1) / r) * a.nstrips, rounds = (tasks + sm - 1) / sm;
            double cost = (double)rounds * (r + 5);
            if (cost < best) { best = cost; rs = r; }
        }
    }
    a.rs = rs; a.nsegs = (a.ny + rs - 1) / rs;
    const size_t bytes = a.P * a.narr * sizeof(float) * (a.layout == 3 ? 2 : 1);
    if (cudaMalloc(&a.base, bytes) != cudaSuccess) { printf("alloc of %.1f GB failed\n", bytes / 1e9); return 1; }
    cudaMemset(a.base, 0, bytes);
    const size_t smem = (size_t)NSTAGE * MAXIN * SW * sizeof(float);
    cudaFuncSetAttribute(k_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = a.nstrips * a.nsegs < sm ? a.nstrips * a.nsegs : sm;
    for (int i = 0; i < 2; i++) k_strip<<<grid, CONSUMERS + 32, smem>>>(a);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; i++) k_strip<<<grid, CONSUMERS + 32, smem>>>(a);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double moved = (double)a.nx * a.ny * 4.0 * (a.nin + a.nout);
    printf("nx=%d ny=%d layout=%d nin=%d nout=%d rs=%d arena=%.1fGB  %.3f ms  %.0f GB/s  (%s)\n", a.nx, a.ny, a.layout,
           a.nin, a.nout, a.rs, bytes / 1e9, ms, moved / ms / 1e6, cudaGetErrorString(e));
    return 0;
}
