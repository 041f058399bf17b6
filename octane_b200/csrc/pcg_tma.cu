// pcg_tma.cu -- PCG pass 1 for large levels: persistent, TMA-fed, warp-specialised.
//
// Same arithmetic as k_pcg_pass1 (pcg.cu) -- p = z + beta p with z = M^-1 r, then
// q = A p row by row with three rows of p rolling in registers, partial p.q --
// but the operands arrive through a shared-memory ring filled by bulk-tensor
// copies (cp.async.bulk -> SASS UBLKCP) that one producer thread issues several
// rows ahead, completion signalled on mbarriers.  The v1 kernel kept one row of
// loads in flight per warp and sat at 25 % occupancy waiting on the long
// scoreboard (profiles/r01_ncu_pass1_v1_conus.txt: 57 % of DRAM peak); here the
// bytes in flight are set by the ring depth (2 rows x 45 KB per SM), not by
// registers.
//
// One CTA per SM: 8 consumer warps (256 threads x 4 pixels = a 1024-pixel strip)
// + 1 producer warp.  A task is a strip x row segment; tasks are dealt
// round-robin to the persistent CTAs.  Stage layout for one row of a strip that
// starts at column i0 (SW = strip width, multiple of 32):
//   RU RV PU PV A1 A4 W : SW+8 floats each, columns i0-4 .. i0+SW+3 (halo for the i-1/i+1 taps)
//   A2 N XU XV          : SW floats each
// W, A2, XU, XV are fetched for the task's own rows only, N also for the row above them.
// Replaces jMatXVec/multiply_row + the p and x updates of
// src/oct_variational_optical_flow.cu:112-139,1138-1146,1161,1172 (reference tree).
#include "kernels.cuh"

namespace octane {

namespace {

constexpr int SWMAX = 1024;                 // pixels per strip (256 consumer threads x 4)
constexpr int HALO = 4;                     // floats of left halo (keeps 16-byte alignment)
constexpr int NSTAGE = 4;
constexpr int HA = SWMAX + 2 * HALO;        // floats per halo array
constexpr int NHALO = 7, NCENTRE = 4;
constexpr int STAGE_FLOATS = NHALO * HA + NCENTRE * SWMAX;
enum { XM_NONE = 0, XM_INIT = 1, XM_ACC = 2 };   // as in pcg.cu
constexpr int CONSUMERS = 256;

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
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float& el(float4& v, int k) { return reinterpret_cast<float*>(&v)[k]; }
__device__ __forceinline__ const float& el(const float4& v, int k) { return reinterpret_cast<const float*>(&v)[k]; }

struct TArgs {
    PcgBuffers b;
    Geom g;
    int ja, jb;
    int cur;
    int store_halo;
    int sw;            // strip width in pixels (multiple of 32, <= SWMAX)
    int rs;            // rows per task
    int nstrips, nsegs;
};

struct PRowT {
    float4 pu, pv, a1, a4;
    float eu_l, ev_l, eu_r, ev_r;     // p of the pixels just left / right of the WARP's 128 pixels
};

__device__ __forceinline__ float mul_lo(int i, int n) { return i == 0 ? 0.f : (i == n - 1 ? 2.f : 1.f); }
__device__ __forceinline__ float mul_hi(int i, int n) { return i == n - 1 ? 0.f : (i == 0 ? 2.f : 1.f); }

// p_new of one staged row for this thread's 4 pixels (+ the warp-edge pixels in lanes 0 / 31);
// when `xdst` is set (the row is one of the task's own), also the previous iteration's pending
// x += alpha_prev p_old (:1172), written straight to global memory.
template <int XM>
__device__ __forceinline__ PRowT p_from_stage(const float* st, int i0s, int tcol, int nx, int lane, float beta,
                                              float alpha_prev, float* xu_dst, float* xv_dst)
{
    constexpr bool FIRST = (XM == XM_NONE);
    // tcol = 4*tid: this thread's first pixel inside the strip; halo arrays are shifted by HALO
    const float* RU = st;
    const float* RV = st + HA;
    const float* PU = st + 2 * HA;
    const float* PV = st + 3 * HA;
    const float* A1 = st + 4 * HA;
    const float* A4 = st + 5 * HA;
    PRowT o;
    o.pu = o.pv = make_float4(0.f, 0.f, 0.f, 0.f);
    o.eu_l = o.ev_l = o.eu_r = o.ev_r = 0.f;
    const int c = tcol + HALO;
    const float4 ru = lds4(RU + c), rv = lds4(RV + c);
    o.a1 = lds4(A1 + c);
    o.a4 = lds4(A4 + c);
    float4 po_u = make_float4(0.f, 0.f, 0.f, 0.f), po_v = po_u;
    if (!FIRST) { po_u = lds4(PU + c); po_v = lds4(PV + c); }
    if (!FIRST && xu_dst) {
        float4 x_u = make_float4(0.f, 0.f, 0.f, 0.f), x_v = x_u;
        if (XM == XM_ACC) {
            const float* XU = st + NHALO * HA + 2 * SWMAX;
            x_u = lds4(XU + tcol);
            x_v = lds4(XU + SWMAX + tcol);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0s + tcol + k < nx) {
                el(x_u, k) = fmaf(alpha_prev, el(po_u, k), el(x_u, k));
                el(x_v, k) = fmaf(alpha_prev, el(po_v, k), el(x_v, k));
            } else {
                el(x_u, k) = 0.f; el(x_v, k) = 0.f;
            }
        }
        stg4(xu_dst, x_u);
        stg4(xv_dst, x_v);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (i0s + tcol + k < nx) {
            const float mu = 1.0f / el(o.a1, k), mv = 1.0f / el(o.a4, k);      // jDiagInv, :142-149
            const float zu = mu * el(ru, k), zv = mv * el(rv, k);              // z = Minv r, :1138
            el(o.pu, k) = FIRST ? zu : fmaf(beta, el(po_u, k), zu);            // p = Bk p + z, :1146
            el(o.pv, k) = FIRST ? zv : fmaf(beta, el(po_v, k), zv);
        }
    }
    if (lane == 0 || lane == 31) {
        const int ce = (lane == 0) ? c - 1 : c + 4;          // column just outside the warp's 128 pixels
        const int ie = i0s + ce - HALO;
        if (ie >= 0 && ie < nx) {
            const float mu = 1.0f / A1[ce], mv = 1.0f / A4[ce];
            const float zu = mu * RU[ce], zv = mv * RV[ce];
            const float eu = FIRST ? zu : fmaf(beta, PU[ce], zu);
            const float ev = FIRST ? zv : fmaf(beta, PV[ce], zv);
            if (lane == 0) { o.eu_l = eu; o.ev_l = ev; } else { o.eu_r = eu; o.ev_r = ev; }
        }
    }
    return o;
}

template <int XM>
__global__ void __launch_bounds__(CONSUMERS + 32, 1) k_pcg_pass1_tma(TArgs a)
{
    constexpr bool FIRST = (XM == XM_NONE);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE];
    const PcgScalars* s = a.b.scal;
    if (s->done) return;
    float* stages = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x;
    const Geom& g = a.g;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NSTAGE; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ntasks = a.nstrips * a.nsegs;
    double dot[1] = { 0.0 };

    if (tid >= CONSUMERS) {
        // ---------------- producer: one thread walks the same (task, row) sequence ----------------
        if (tid == CONSUMERS) {
            const float* src_h[NHALO] = { a.b.ru, a.b.rv, a.b.pu[a.cur], a.b.pv[a.cur], a.b.coef[C_A1], a.b.coef[C_A4],
                                          a.b.coef[C_W] };
            uint32_t it = 0;
            for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
                const int seg = t / a.nstrips, strip = t - seg * a.nstrips;
                const int i0s = strip * a.sw;
                const int j_a = a.ja + seg * a.rs, j_b = min(a.jb, j_a + a.rs);
                // halo arrays: columns [h0, h1) of the row, landing at stage offset h0 - (i0s - HALO)
                const int h0 = max(i0s - HALO, 0), h1 = min(i0s + a.sw + HALO, g.pitch);
                const int w = min(a.sw, g.pitch - i0s);
                const uint32_t hb = (uint32_t)(h1 - h0) * 4u, cb = (uint32_t)w * 4u;
                for (int jr = max(j_a - 1, 0); jr <= min(j_b, g.ny - 1); jr++, it++) {
                    const int stg = it % NSTAGE;
                    const uint32_t par = (it / NSTAGE) & 1u;
                    mbar_wait(&empty_bar[stg], par ^ 1u);
                    float* st = stages + (size_t)stg * STAGE_FLOATS;
                    const bool centre = jr >= j_a && jr < j_b;
                    const bool need_n = jr < j_b;                    // own rows and the row above them
                    const uint32_t nh = (FIRST ? 4u : 6u) + (centre ? 1u : 0u);
                    const uint32_t ncb = (centre ? 1u : 0u) + (need_n ? 1u : 0u) + ((centre && XM == XM_ACC) ? 2u : 0u);
                    mbar_expect_tx(&full_bar[stg], nh * hb + ncb * cb);
                    const size_t row = g.at(0, jr);
#pragma unroll
                    for (int q = 0; q < NHALO; q++) {
                        if (FIRST && (q == 2 || q == 3)) continue;
                        if (q == 6 && !centre) continue;
                        bulk_g2s(st + q * HA + (h0 - (i0s - HALO)), src_h[q] + row + h0, hb, &full_bar[stg]);
                    }
                    float* cst = st + NHALO * HA;
                    if (centre) bulk_g2s(cst, a.b.coef[C_A2] + row + i0s, cb, &full_bar[stg]);
                    if (need_n) bulk_g2s(cst + SWMAX, a.b.coef[C_N] + row + i0s, cb, &full_bar[stg]);
                    if (centre && XM == XM_ACC) {
                        bulk_g2s(cst + 2 * SWMAX, a.b.xu + row + i0s, cb, &full_bar[stg]);
                        bulk_g2s(cst + 3 * SWMAX, a.b.xv + row + i0s, cb, &full_bar[stg]);
                    }
                }
            }
        }
    } else {
        // ---------------- consumers -------------------------------------------------------------
        const float beta = FIRST ? 0.f : s->rz / s->rz_old;           // Bk, :1144
        const float alpha_prev = FIRST ? 0.f : s->alpha;
        const int lane = tid & 31;
        const int tcol = tid * 4;
        float* pu_new = a.b.pu[a.cur ^ 1];
        float* pv_new = a.b.pv[a.cur ^ 1];
        uint32_t it = 0;
        for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
            const int seg = t / a.nstrips, strip = t - seg * a.nstrips;
            const int i0s = strip * a.sw;
            const int i0 = i0s + tcol;
            const int j_a = a.ja + seg * a.rs, j_b = min(a.jb, j_a + a.rs);
            const bool active = tcol < a.sw && i0 < g.nx;
            PRowT up, ce, dn;
            up.pu = up.pv = up.a1 = up.a4 = make_float4(0.f, 0.f, 0.f, 0.f);
            up.eu_l = up.ev_l = up.eu_r = up.ev_r = 0.f;
            ce = up;
            float4 n_up = make_float4(0.f, 0.f, 0.f, 0.f);       // N of the row above the centre row
            float4 n_ce = n_up;                                  // N of the row that is about to become the centre
            int prev_stage = -1;                       // stage holding the centre row's coefficients
            for (int jr = j_a - 1; jr <= j_b; jr++) {
                int stg = -1;
                float4 n_dn = make_float4(0.f, 0.f, 0.f, 0.f);
                if (jr >= 0 && jr < g.ny) {
                    stg = it % NSTAGE;
                    mbar_wait(&full_bar[stg], (it / NSTAGE) & 1u);
                    it++;
                    const float* st = stages + (size_t)stg * STAGE_FLOATS;
                    const bool own = active && jr >= j_a && jr < j_b;
                    const size_t xoff = own ? g.at(i0, jr) : 0;
                    dn = p_from_stage<XM>(st, i0s, tcol, g.nx, lane, beta, alpha_prev,
                                          own ? a.b.xu + xoff : nullptr, own ? a.b.xv + xoff : nullptr);
                    if (!active) { dn.pu = dn.pv = make_float4(0.f, 0.f, 0.f, 0.f); }
                    if (active && jr < j_b) n_dn = lds4(st + NHALO * HA + SWMAX + tcol);
                } else {
                    dn.pu = dn.pv = dn.a1 = dn.a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    dn.eu_l = dn.ev_l = dn.eu_r = dn.ev_r = 0.f;
                }
                const int jc = jr - 1;                 // centre row: up = jc-1, ce = jc, dn = jc+1
                if (jc >= j_a && jc < j_b) {
                    float lu = __shfl_up_sync(0xffffffffu, ce.pu.w, 1), lv = __shfl_up_sync(0xffffffffu, ce.pv.w, 1);
                    float ru_ = __shfl_down_sync(0xffffffffu, ce.pu.x, 1), rv_ = __shfl_down_sync(0xffffffffu, ce.pv.x, 1);
                    if (lane == 0) { lu = ce.eu_l; lv = ce.ev_l; }
                    if (lane == 31) { ru_ = ce.eu_r; rv_ = ce.ev_r; }
                    if (active) {
                        const float* pst = stages + (size_t)prev_stage * STAGE_FLOATS;
                        const float4 a2 = lds4(pst + NHALO * HA + tcol);
                        const float* Wrow = pst + 6 * HA + HALO + tcol;
                        const float4 w = lds4(Wrow);
                        const float wl = (i0 > 0) ? Wrow[-1] : 0.f;     // column i0-1 (never staged at the image edge)
                        const float m6 = mul_lo(jc, g.ny), m8 = mul_hi(jc, g.ny);
                        float4 qu, qv;
                        float part = 0.f;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float pl_u = (k == 0) ? lu : el(ce.pu, k - 1), pl_v = (k == 0) ? lv : el(ce.pv, k - 1);
                            const float pr_u = (k == 3) ? ru_ : el(ce.pu, k + 1), pr_v = (k == 3) ? rv_ : el(ce.pv, k + 1);
                            const float a5 = mul_lo(i0 + k, g.nx) * ((k == 0) ? wl : el(w, k - 1));
                            const float a7 = mul_hi(i0 + k, g.nx) * el(w, k);
                            const float a6 = m6 * el(n_up, k), a8 = m8 * el(n_ce, k);
                            float su = 0.f;                       // multiply_row order: [j-1] [i-1] a1 a2 [i+1] [j+1]
                            su = fmaf(a6, el(up.pu, k), su);
                            su = fmaf(a5, pl_u, su);
                            su = fmaf(el(ce.a1, k), el(ce.pu, k), su);
                            su = fmaf(el(a2, k), el(ce.pv, k), su);
                            su = fmaf(a7, pr_u, su);
                            su = fmaf(a8, el(dn.pu, k), su);
                            float sv = 0.f;
                            sv = fmaf(a6, el(up.pv, k), sv);
                            sv = fmaf(a5, pl_v, sv);
                            sv = fmaf(el(a2, k), el(ce.pu, k), sv);
                            sv = fmaf(el(ce.a4, k), el(ce.pv, k), sv);
                            sv = fmaf(a7, pr_v, sv);
                            sv = fmaf(a8, el(dn.pv, k), sv);
                            const bool in = i0 + k < g.nx;
                            el(qu, k) = in ? su : 0.f;
                            el(qv, k) = in ? sv : 0.f;
                            if (in) part += el(ce.pu, k) * su + el(ce.pv, k) * sv;
                        }
                        const size_t off = g.at(i0, jc);
                        stg4(pu_new + off, ce.pu);
                        stg4(pv_new + off, ce.pv);
                        stg4(a.b.qu + off, qu);
                        stg4(a.b.qv + off, qv);
                        dot[0] += (double)part;
                    }
                } else if (a.store_halo && active && jc >= 0 && jc < g.ny &&
                           ((jc == a.ja - 1 && j_a == a.ja) || (jc == a.jb && j_b == a.jb))) {
                    // banded runs keep p on the halo rows for the next iteration's p_old
                    stg4(pu_new + g.at(i0, jc), ce.pu);
                    stg4(pv_new + g.at(i0, jc), ce.pv);
                }
                // the centre row's stage is no longer needed: hand it back to the producer
                if (prev_stage >= 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[prev_stage]);
                }
                prev_stage = stg;
                up = ce;
                ce = dn;
                n_up = n_ce;
                n_ce = n_dn;
            }
            // last staged row of the task (row j_b, or none when j_b == ny)
            if (a.store_halo && active && j_b == a.jb && j_b < g.ny) {
                stg4(pu_new + g.at(i0, j_b), ce.pu);
                stg4(pv_new + g.at(i0, j_b), ce.pv);
            }
            if (prev_stage >= 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[prev_stage]);
            }
        }
    }
    // ---------------- p.q: fixed-order block + grid reduction (all 288 threads) ----------------
    block_sum<1>(dot, red);
    double tot[1];
    if (grid_sum_finish<1>(dot, a.b.partials, a.b.ticket, tot, red)) {
        if (threadIdx.x == 0) {
            if (a.b.defer) a.b.pending[0] = tot[0];
            else a.b.scal->pAp = (float)tot[0];
        }
    }
}

}  // namespace

bool pcg_pass1_tma_usable(const Geom& g, int nrows)
{
    return g.nx >= 512 && nrows >= 64 && (g.pitch % 32) == 0;
}

void launch_pcg_pass1_tma(const PcgBuffers& b, const Geom& g, int ja, int jb, int ki, int store_halo,
                          int sm_count, cudaStream_t st)
{
    TArgs a;
    a.b = b; a.g = g; a.ja = ja; a.jb = jb; a.cur = ki & 1; a.store_halo = store_halo;
    a.nstrips = (g.nx + SWMAX - 1) / SWMAX;
    a.sw = round_up((g.nx + a.nstrips - 1) / a.nstrips, 32);
    if (a.sw > SWMAX) a.sw = SWMAX;
    a.nstrips = (g.nx + a.sw - 1) / a.sw;
    // rows per task: minimise rounds x (rows + 2 halo rows) over the persistent grid
    const int nrows = jb - ja;
    int best_rs = 64;
    double best_cost = 1e30;
    for (int rs = 24; rs <= 256; rs++) {
        const int nsegs = (nrows + rs - 1) / rs;
        const long long tasks = (long long)nsegs * a.nstrips;
        const long long rounds = (tasks + sm_count - 1) / sm_count;
        const double cost = (double)rounds * (rs + 2 + 3);      // +3: pipeline fill per task
        if (cost < best_cost) { best_cost = cost; best_rs = rs; }
    }
    a.rs = best_rs;
    a.nsegs = (nrows + a.rs - 1) / a.rs;
    const int ntasks = a.nstrips * a.nsegs;
    int grid = ntasks < sm_count ? ntasks : sm_count;
    if (grid > b.max_partial_blocks) grid = b.max_partial_blocks;
    const size_t smem = (size_t)NSTAGE * STAGE_FLOATS * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_pcg_pass1_tma<XM_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pcg_pass1_tma<XM_INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pcg_pass1_tma<XM_ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    if (ki == 0)      k_pcg_pass1_tma<XM_NONE><<<grid, CONSUMERS + 32, smem, st>>>(a);
    else if (ki == 1) k_pcg_pass1_tma<XM_INIT><<<grid, CONSUMERS + 32, smem, st>>>(a);
    else              k_pcg_pass1_tma<XM_ACC><<<grid, CONSUMERS + 32, smem, st>>>(a);
}

}  // namespace octane
