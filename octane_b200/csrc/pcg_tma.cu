// pcg_tma.cu -- PCG pass 1 for large levels: persistent, TMA-fed, warp-specialised.
//
// Same arithmetic as k_pcg_pass1 (pcg.cu) -- p = z + beta p with z = M^-1 r, then
// q = A p row by row with three rows of p rolling in registers, partial p.q --
// but the operands arrive through a shared-memory ring filled by bulk-tensor
// copies (cp.async.bulk -> SASS UBLKCP) that one producer thread issues several
// rows ahead, completion signalled on mbarriers.  The v1 kernel kept one row of
// loads in flight per warp and sat at 25 % occupancy waiting on the long
// scoreboard (profiles/r01_ncu_pass1_v1_conus.txt: 57 % of DRAM peak); here the
// bytes in flight are set by the ring depth (4 rows x 45 KB per SM), not by
// registers: a consumer warp copies what it needs of a staged row into registers
// and hands the stage straight back to the producer.
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
#include <stdlib.h>

#include "kernels.cuh"

namespace octane {

namespace {

constexpr int SWMAX = 1024;                 // widest strip the stage layout holds
constexpr int P1PX = 1;                     // pixels per consumer thread (see Consumers below)
constexpr int HALO = 4;                     // floats of left halo (keeps 16-byte alignment)
#ifndef OCTANE_P1_NSTAGE
#define OCTANE_P1_NSTAGE 4        // measured: 4 x 45.3 KB beats 5 (226 of the SM's 227 KB) by 1-3 %
#endif
constexpr int NSTAGE = OCTANE_P1_NSTAGE;
constexpr int HA = SWMAX + 2 * HALO;        // floats per halo array
constexpr int NHALO = 7, NCENTRE = 4;
constexpr int STAGE_FLOATS = NHALO * HA + NCENTRE * SWMAX;
enum { XM_NONE = 0, XM_INIT = 1, XM_ACC = 2 };   // as in pcg.cu

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

// PX consecutive pixels per thread (2 or 4), moved as one float2 / float4
template <int PX> struct Px { float v[PX]; };
template <int PX> __device__ __forceinline__ Px<PX> ldv(const float* p);
template <> __device__ __forceinline__ Px<4> ldv<4>(const float* p)
{
    const float4 t = *reinterpret_cast<const float4*>(p);
    Px<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <> __device__ __forceinline__ Px<2> ldv<2>(const float* p)
{
    const float2 t = *reinterpret_cast<const float2*>(p);
    Px<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <> __device__ __forceinline__ Px<1> ldv<1>(const float* p)
{
    Px<1> r; r.v[0] = *p; return r;
}
__device__ __forceinline__ void stv(float* p, const Px<1>& a) { *p = a.v[0]; }
__device__ __forceinline__ void stv(float* p, const Px<4>& a) { *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }
__device__ __forceinline__ void stv(float* p, const Px<2>& a) { *reinterpret_cast<float2*>(p) = make_float2(a.v[0], a.v[1]); }
template <int PX> __device__ __forceinline__ Px<PX> zerov()
{
    Px<PX> r;
#pragma unroll
    for (int k = 0; k < PX; k++) r.v[k] = 0.f;
    return r;
}

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

template <int PX> struct PRowT {
    Px<PX> pu, pv, a1, a4;
    float eu_l, ev_l, eu_r, ev_r;     // p of the pixels just left / right of the WARP's 32*PX pixels
};

__device__ __forceinline__ float mul_lo(int i, int n) { return i == 0 ? 0.f : (i == n - 1 ? 2.f : 1.f); }
__device__ __forceinline__ float mul_hi(int i, int n) { return i == n - 1 ? 0.f : (i == 0 ? 2.f : 1.f); }

// p_new of one staged row for this thread's PX pixels (+ the warp-edge pixels in lanes 0 / 31);
// when `xdst` is set (the row is one of the task's own), also the previous iteration's pending
// x += alpha_prev p_old (:1172), written straight to global memory.
template <int XM, int PX>
__device__ __forceinline__ PRowT<PX> p_from_stage(const float* st, int i0s, int tcol, int nx, int lane, float beta,
                                                  float alpha_prev, float* xu_dst, float* xv_dst)
{
    constexpr bool FIRST = (XM == XM_NONE);
    // tcol = PX*tid: this thread's first pixel inside the strip; halo arrays are shifted by HALO
    const float* RU = st;
    const float* RV = st + HA;
    const float* PU = st + 2 * HA;
    const float* PV = st + 3 * HA;
    const float* A1 = st + 4 * HA;
    const float* A4 = st + 5 * HA;
    PRowT<PX> o;
    o.pu = o.pv = zerov<PX>();
    o.eu_l = o.ev_l = o.eu_r = o.ev_r = 0.f;
    const int c = tcol + HALO;
    const Px<PX> ru = ldv<PX>(RU + c), rv = ldv<PX>(RV + c);
    o.a1 = ldv<PX>(A1 + c);
    o.a4 = ldv<PX>(A4 + c);
    Px<PX> po_u = zerov<PX>(), po_v = po_u;
    if (!FIRST) { po_u = ldv<PX>(PU + c); po_v = ldv<PX>(PV + c); }
    if (!FIRST && xu_dst) {
        Px<PX> x_u = zerov<PX>(), x_v = x_u;
        if (XM == XM_ACC) {
            const float* XU = st + NHALO * HA + 2 * SWMAX;
            x_u = ldv<PX>(XU + tcol);
            x_v = ldv<PX>(XU + SWMAX + tcol);
        }
#pragma unroll
        for (int k = 0; k < PX; k++) {
            if (i0s + tcol + k < nx) {
                x_u.v[k] = fmaf(alpha_prev, po_u.v[k], x_u.v[k]);
                x_v.v[k] = fmaf(alpha_prev, po_v.v[k], x_v.v[k]);
            } else {
                x_u.v[k] = 0.f; x_v.v[k] = 0.f;
            }
        }
        stv(xu_dst, x_u);
        stv(xv_dst, x_v);
    }
#pragma unroll
    for (int k = 0; k < PX; k++) {
        if (i0s + tcol + k < nx) {
            const float mu = 1.0f / o.a1.v[k], mv = 1.0f / o.a4.v[k];          // jDiagInv, :142-149
            const float zu = mu * ru.v[k], zv = mv * rv.v[k];                  // z = Minv r, :1138
            o.pu.v[k] = FIRST ? zu : fmaf(beta, po_u.v[k], zu);                // p = Bk p + z, :1146
            o.pv.v[k] = FIRST ? zv : fmaf(beta, po_v.v[k], zv);
        }
    }
    if (lane == 0 || lane == 31) {
        const int ce = (lane == 0) ? c - 1 : c + PX;         // column just outside the warp's 32*PX pixels
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

// The stencil row is latency-bound per warp (dependent FMA chains, shared-memory loads, shuffles),
// so what keeps the copy engine busy is the number of consumer warps.  Measured on the full disk
// (profiles/r02_pass1_levers_ab.txt): 8 warps x 4 pixels 5.86 ms per launch, 16 x 2 5.56 ms, 31 x 1 5.06 ms
// (6.23 TB/s) -- so one pixel per thread ships: 992 consumer threads + the producer warp fill the
// 1024-thread block, strips are at most 992 pixels wide.
template <int PX> struct Consumers { static constexpr int N = SWMAX / PX; };
template <> struct Consumers<1> { static constexpr int N = SWMAX - 32; };
// CWN: in the first GNC stage the build stores W = N = -1 at every pixel (al1 = 1: the couplings are
// -(1 + 0 * psi), build.cu), so the solves of that stage do not read the two planes at all: 8 of pass 1's
// 68 B/px in a third of the solves (measured: full disk 3018 -> 2975 ms per pair, bit-identical results).  The boundary multipliers mul_lo / mul_hi
// zero the couplings that leave the scene, exactly as they do for the stored planes.
template <int XM, int PX, bool CWN>
__global__ void __launch_bounds__(Consumers<PX>::N + 32, 1) k_pcg_pass1_tma(TArgs a)
{
    constexpr bool FIRST = (XM == XM_NONE);
    constexpr int CONSUMERS = Consumers<PX>::N;
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
                    const uint32_t nh = (FIRST ? 4u : 6u) + ((centre && !CWN) ? 1u : 0u);
                    const uint32_t ncb = (centre ? 1u : 0u) + ((need_n && !CWN) ? 1u : 0u) + ((centre && XM == XM_ACC) ? 2u : 0u);
                    mbar_expect_tx(&full_bar[stg], nh * hb + ncb * cb);
                    const size_t row = g.at(0, jr);
#pragma unroll
                    for (int q = 0; q < NHALO; q++) {
                        if (FIRST && (q == 2 || q == 3)) continue;
                        if (q == 6 && (!centre || CWN)) continue;
                        bulk_g2s(st + q * HA + (h0 - (i0s - HALO)), src_h[q] + row + h0, hb, &full_bar[stg]);
                    }
                    float* cst = st + NHALO * HA;
                    if (centre) bulk_g2s(cst, a.b.coef[C_A2] + row + i0s, cb, &full_bar[stg]);
                    if (need_n && !CWN) bulk_g2s(cst + SWMAX, a.b.coef[C_N] + row + i0s, cb, &full_bar[stg]);
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
        const int tcol = tid * PX;
        float* pu_new = a.b.pu[a.cur ^ 1];
        float* pv_new = a.b.pv[a.cur ^ 1];
        uint32_t it = 0;
        for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
            const int seg = t / a.nstrips, strip = t - seg * a.nstrips;
            const int i0s = strip * a.sw;
            const int i0 = i0s + tcol;
            const int j_a = a.ja + seg * a.rs, j_b = min(a.jb, j_a + a.rs);
            const bool active = tcol < a.sw && i0 < g.nx;
            PRowT<PX> up, ce, dn;
            const Px<PX> zero = zerov<PX>();
            up.pu = up.pv = up.a1 = up.a4 = zero;
            up.eu_l = up.ev_l = up.eu_r = up.ev_r = 0.f;
            ce = up;
            Px<PX> n_up = zero;                        // N of the row above the centre row
            Px<PX> n_ce = zero;                        // N, a2, W (+ W of column i0-1) of the centre row
            Px<PX> a2_ce = zero, w_ce = zero;
            float wl_ce = 0.f;
            for (int jr = j_a - 1; jr <= j_b; jr++) {
                Px<PX> n_dn = zero, a2_dn = zero, w_dn = zero;
                float wl_dn = 0.f;
                if (jr >= 0 && jr < g.ny) {
                    const int stg = it % NSTAGE;
                    mbar_wait(&full_bar[stg], (it / NSTAGE) & 1u);
                    it++;
                    const float* st = stages + (size_t)stg * STAGE_FLOATS;
                    const bool centre = jr >= j_a && jr < j_b;
                    const bool own = active && centre;
                    const size_t xoff = own ? g.at(i0, jr) : 0;
                    dn = p_from_stage<XM, PX>(st, i0s, tcol, g.nx, lane, beta, alpha_prev,
                                              own ? a.b.xu + xoff : nullptr, own ? a.b.xv + xoff : nullptr);
                    if (!active) { dn.pu = dn.pv = zero; }
                    if (active && jr < j_b) {
                        if (CWN) {
#pragma unroll
                            for (int k = 0; k < PX; k++) n_dn.v[k] = -1.f;
                        } else {
                            n_dn = ldv<PX>(st + NHALO * HA + SWMAX + tcol);
                        }
                    }
                    if (own) {
                        // everything the row needs as a centre row goes to registers now, so the stage
                        // returns to the producer at once (bytes in flight = the whole ring)
                        a2_dn = ldv<PX>(st + NHALO * HA + tcol);
                        if (CWN) {
#pragma unroll
                            for (int k = 0; k < PX; k++) w_dn.v[k] = -1.f;
                            wl_dn = (i0 > 0) ? -1.f : 0.f;
                        } else {
                            const float* Wrow = st + 6 * HA + HALO + tcol;
                            w_dn = ldv<PX>(Wrow);
                            wl_dn = (i0 > 0) ? Wrow[-1] : 0.f;     // column i0-1 (never staged at the image edge)
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[stg]);
                } else {
                    dn.pu = dn.pv = dn.a1 = dn.a4 = zero;
                    dn.eu_l = dn.ev_l = dn.eu_r = dn.ev_r = 0.f;
                }
                const int jc = jr - 1;                 // centre row: up = jc-1, ce = jc, dn = jc+1
                if (jc >= j_a && jc < j_b) {
                    float lu = __shfl_up_sync(0xffffffffu, ce.pu.v[PX - 1], 1), lv = __shfl_up_sync(0xffffffffu, ce.pv.v[PX - 1], 1);
                    float ru_ = __shfl_down_sync(0xffffffffu, ce.pu.v[0], 1), rv_ = __shfl_down_sync(0xffffffffu, ce.pv.v[0], 1);
                    if (lane == 0) { lu = ce.eu_l; lv = ce.ev_l; }
                    if (lane == 31) { ru_ = ce.eu_r; rv_ = ce.ev_r; }
                    if (active) {
                        const float m6 = mul_lo(jc, g.ny), m8 = mul_hi(jc, g.ny);
                        Px<PX> qu, qv;
                        float part = 0.f;
#pragma unroll
                        for (int k = 0; k < PX; k++) {
                            const float pl_u = (k == 0) ? lu : ce.pu.v[k > 0 ? k - 1 : 0], pl_v = (k == 0) ? lv : ce.pv.v[k > 0 ? k - 1 : 0];
                            const float pr_u = (k == PX - 1) ? ru_ : ce.pu.v[k < PX - 1 ? k + 1 : 0];
                            const float pr_v = (k == PX - 1) ? rv_ : ce.pv.v[k < PX - 1 ? k + 1 : 0];
                            const float a5 = mul_lo(i0 + k, g.nx) * ((k == 0) ? wl_ce : w_ce.v[k > 0 ? k - 1 : 0]);
                            const float a7 = mul_hi(i0 + k, g.nx) * w_ce.v[k];
                            const float a6 = m6 * n_up.v[k], a8 = m8 * n_ce.v[k];
                            float su = 0.f;                       // multiply_row order: [j-1] [i-1] a1 a2 [i+1] [j+1]
                            su = fmaf(a6, up.pu.v[k], su);
                            su = fmaf(a5, pl_u, su);
                            su = fmaf(ce.a1.v[k], ce.pu.v[k], su);
                            su = fmaf(a2_ce.v[k], ce.pv.v[k], su);
                            su = fmaf(a7, pr_u, su);
                            su = fmaf(a8, dn.pu.v[k], su);
                            float sv = 0.f;
                            sv = fmaf(a6, up.pv.v[k], sv);
                            sv = fmaf(a5, pl_v, sv);
                            sv = fmaf(a2_ce.v[k], ce.pu.v[k], sv);
                            sv = fmaf(ce.a4.v[k], ce.pv.v[k], sv);
                            sv = fmaf(a7, pr_v, sv);
                            sv = fmaf(a8, dn.pv.v[k], sv);
                            const bool in = i0 + k < g.nx;
                            qu.v[k] = in ? su : 0.f;
                            qv.v[k] = in ? sv : 0.f;
                            if (in) part += ce.pu.v[k] * su + ce.pv.v[k] * sv;
                        }
                        const size_t off = g.at(i0, jc);
                        stv(pu_new + off, ce.pu);
                        stv(pv_new + off, ce.pv);
                        stv(a.b.qu + off, qu);
                        stv(a.b.qv + off, qv);
                        dot[0] += (double)part;
                    }
                } else if (a.store_halo && active && jc >= 0 && jc < g.ny &&
                           ((jc == a.ja - 1 && j_a == a.ja) || (jc == a.jb && j_b == a.jb))) {
                    // banded runs keep p on the halo rows for the next iteration's p_old
                    stv(pu_new + g.at(i0, jc), ce.pu);
                    stv(pv_new + g.at(i0, jc), ce.pv);
                }
                up = ce;
                ce = dn;
                n_up = n_ce;
                n_ce = n_dn;
                a2_ce = a2_dn;
                w_ce = w_dn;
                wl_ce = wl_dn;
            }
            // last staged row of the task (row j_b, or none when j_b == ny)
            if (a.store_halo && active && j_b == a.jb && j_b < g.ny) {
                stv(pu_new + g.at(i0, j_b), ce.pu);
                stv(pv_new + g.at(i0, j_b), ce.pv);
            }
        }
    }
    // ---------------- p.q: fixed-order block + grid reduction (all threads) -------------------
    block_sum<1>(dot, red);
    double tot[1];
    if (grid_sum_finish<1>(dot, a.b.partials, a.b.ticket, tot, red)) {
        if (a.b.p2p.world > 1) p2p_allreduce<1>(a.b.p2p, P2P_PASS1, tot, &a.b.scal->comm_err);
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

namespace {
template <int PX, bool CWN>
void launch_variant(const TArgs& a, int xm, int grid, size_t smem, cudaStream_t st)
{
    static unsigned long long configured = 0;
    if (first_launch_on_device(&configured)) {
        cudaFuncSetAttribute(k_pcg_pass1_tma<XM_NONE, PX, CWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pcg_pass1_tma<XM_INIT, PX, CWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pcg_pass1_tma<XM_ACC, PX, CWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int threads = Consumers<PX>::N + 32;
    if (xm == XM_NONE)      k_pcg_pass1_tma<XM_NONE, PX, CWN><<<grid, threads, smem, st>>>(a);
    else if (xm == XM_INIT) k_pcg_pass1_tma<XM_INIT, PX, CWN><<<grid, threads, smem, st>>>(a);
    else                    k_pcg_pass1_tma<XM_ACC, PX, CWN><<<grid, threads, smem, st>>>(a);
}
}  // namespace

void launch_pcg_pass1_tma(const PcgBuffers& b, const Geom& g, int ja, int jb, int ki, int store_halo,
                          int sm_count, cudaStream_t st, int const_wn)
{
    TArgs a;
    a.b = b; a.g = g; a.ja = ja; a.jb = jb; a.cur = ki & 1; a.store_halo = store_halo;
    constexpr int swmax = Consumers<P1PX>::N;
    a.nstrips = (g.nx + swmax - 1) / swmax;
    a.sw = round_up((g.nx + a.nstrips - 1) / a.nstrips, 32);
    if (a.sw > swmax) a.sw = swmax;
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
    const int xm = ki == 0 ? XM_NONE : (ki == 1 ? XM_INIT : XM_ACC);
    if (const_wn) launch_variant<P1PX, true>(a, xm, grid, smem, st);
    else          launch_variant<P1PX, false>(a, xm, grid, smem, st);
}

}  // namespace octane
