// pcg.cu -- matrix-free Jacobi-preconditioned conjugate gradient on the 2x2-block
// 5-point system.  Replaces jMatXVec / jVecXVec / jVecPVec / jDiagInv and the
// PCG driver of src/oct_variational_optical_flow.cu:112-205,1105-1195
// (reference tree).  The reference spends ~12 grid-wide barriers and ~520 B/px
// per iteration on explicit CSR; here one iteration is two streaming kernels:
//
//   pass 1  p = z + beta p   (z = M^-1 r, recomputed, never stored)
//           q = A p          (stencil, rolling three rows of p in registers)
//           partial p.q                                   60 B/px
//   pass 2  x += alpha p ; r -= alpha q ; partial r.r and z.r
//           stop rule + scalar roll in the last block      64 B/px
//
// Same recurrence, same fp32 scalar arithmetic (alpha = rz/pAp, beta =
// rz_new/rz_old, stop on !(r.r > tol) or the launch cap), same FMA order inside
// each matrix row as multiply_row (:112-121) over the entry order the build
// writes.  Dot products are reduced in a fixed order in double, so a solve is
// bit-reproducible run to run (the reference's float atomics are not).
// Every launch re-reads the device-side `done` flag and returns at once when
// the stop rule has fired, so the host can enqueue the cap's worth of launches
// (or replay a CUDA graph of them) without synchronising.
#include "kernels.cuh"

namespace octane {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_stream(const float* p)
{
    return __ldcs(reinterpret_cast<const float4*>(p));   // read-once coefficient planes: evict-first
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float& el(float4& v, int k) { return reinterpret_cast<float*>(&v)[k]; }
__device__ __forceinline__ const float& el(const float4& v, int k) { return reinterpret_cast<const float*>(&v)[k]; }

struct P1Args {
    PcgBuffers b;
    Geom g;
    int ja, jb;        // rows whose q = A p this rank computes
    int cur;           // p[cur] = p_old, p[cur^1] = p_new
    int store_halo;    // banded: also store p_new of rows ja-1 and jb
    int rs;            // rows per warp task
    int nstrips, nsegs;
};

struct PRow {
    float4 pu, pv;     // p_new of the lane's 4 pixels
    float4 a1, a4;     // diagonal entries of the same pixels
    float eu, ev;      // p_new of the pixel just outside the warp's strip (lanes 0 and 31)
};

// p_new = (1/M) r + beta p_old for one row segment of the warp's strip.
// (1/M) as jDiagInv (:142-149): 1./M rounded to float; z = Minv*r (:1117,1138);
// p = Bk*p + z (:1146, one FMA).
template <bool FIRST>
__device__ __forceinline__ PRow compute_p(const P1Args& a, int j, int i0, int lane, float beta)
{
    PRow o;
    o.pu = o.pv = o.a1 = o.a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    o.eu = o.ev = 0.f;
    const Geom& g = a.g;
    if (j < 0 || j >= g.ny) return o;
    const float* pu_old = a.b.pu[a.cur];
    const float* pv_old = a.b.pv[a.cur];
    if (i0 < g.nx) {
        const size_t off = g.at(i0, j);
        const float4 ru = ld4(a.b.ru + off), rv = ld4(a.b.rv + off);
        o.a1 = ld4(a.b.coef[0] + off);
        o.a4 = ld4(a.b.coef[2] + off);
        float4 po_u = make_float4(0.f, 0.f, 0.f, 0.f), po_v = po_u;
        if (!FIRST) { po_u = ld4(pu_old + off); po_v = ld4(pv_old + off); }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < g.nx) {
                const float mu = 1.0f / el(o.a1, k), mv = 1.0f / el(o.a4, k);
                const float zu = mu * el(ru, k), zv = mv * el(rv, k);
                el(o.pu, k) = FIRST ? zu : fmaf(beta, el(po_u, k), zu);
                el(o.pv, k) = FIRST ? zv : fmaf(beta, el(po_v, k), zv);
            }
        }
    }
    if (lane == 0 || lane == 31) {
        const int ie = (lane == 0) ? i0 - 1 : i0 + 4;
        if (ie >= 0 && ie < g.nx) {
            const size_t off = g.at(ie, j);
            const float mu = 1.0f / a.b.coef[0][off], mv = 1.0f / a.b.coef[2][off];
            const float zu = mu * a.b.ru[off], zv = mv * a.b.rv[off];
            o.eu = FIRST ? zu : fmaf(beta, pu_old[off], zu);
            o.ev = FIRST ? zv : fmaf(beta, pv_old[off], zv);
        }
    }
    return o;
}

template <bool FIRST>
__global__ void __launch_bounds__(256, 2) k_pcg_pass1(P1Args a)
{
    __shared__ double red[32];
    const PcgScalars* s = a.b.scal;
    if (s->done) return;
    const float beta = FIRST ? 0.f : s->rz / s->rz_old;      // Bk, :1144
    const Geom& g = a.g;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ntasks = a.nstrips * a.nsegs;
    float* pu_new = a.b.pu[a.cur ^ 1];
    float* pv_new = a.b.pv[a.cur ^ 1];
    double dot[1] = { 0.0 };

    for (int t = blockIdx.x * 8 + wib; t < ntasks; t += gridDim.x * 8) {
        const int seg = t / a.nstrips, strip = t - seg * a.nstrips;
        const int i0 = strip * 128 + lane * 4;
        const int j_a = a.ja + seg * a.rs, j_b = min(a.jb, j_a + a.rs);
        const bool active = i0 < g.nx;
        PRow up = compute_p<FIRST>(a, j_a - 1, i0, lane, beta);
        PRow ce = compute_p<FIRST>(a, j_a, i0, lane, beta);
        if (a.store_halo && j_a == a.ja && j_a - 1 >= 0 && active) {
            st4(pu_new + g.at(i0, j_a - 1), up.pu);
            st4(pv_new + g.at(i0, j_a - 1), up.pv);
        }
        for (int j = j_a; j < j_b; j++) {
            PRow dn = compute_p<FIRST>(a, j + 1, i0, lane, beta);
            // horizontal neighbours of the centre row: lanes exchange their edge pixels
            float lu = __shfl_up_sync(0xffffffffu, ce.pu.w, 1), lv = __shfl_up_sync(0xffffffffu, ce.pv.w, 1);
            float ru_ = __shfl_down_sync(0xffffffffu, ce.pu.x, 1), rv_ = __shfl_down_sync(0xffffffffu, ce.pv.x, 1);
            if (lane == 0) { lu = ce.eu; lv = ce.ev; }
            if (lane == 31) { ru_ = ce.eu; rv_ = ce.ev; }
            if (active) {
                const size_t off = g.at(i0, j);
                const float4 a2 = ld4_stream(a.b.coef[1] + off), a5 = ld4_stream(a.b.coef[3] + off),
                             a6 = ld4_stream(a.b.coef[4] + off), a7 = ld4_stream(a.b.coef[5] + off),
                             a8 = ld4_stream(a.b.coef[6] + off);
                float4 qu, qv;
                float part = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float pl_u = (k == 0) ? lu : el(ce.pu, k - 1), pl_v = (k == 0) ? lv : el(ce.pv, k - 1);
                    const float pr_u = (k == 3) ? ru_ : el(ce.pu, k + 1), pr_v = (k == 3) ? rv_ : el(ce.pv, k + 1);
                    // row of u: [j-1] [i-1] a1 a2 [i+1] [j+1]   (multiply_row order)
                    float su = 0.f;
                    su = fmaf(el(a6, k), el(up.pu, k), su);
                    su = fmaf(el(a5, k), pl_u, su);
                    su = fmaf(el(ce.a1, k), el(ce.pu, k), su);
                    su = fmaf(el(a2, k), el(ce.pv, k), su);
                    su = fmaf(el(a7, k), pr_u, su);
                    su = fmaf(el(a8, k), el(dn.pu, k), su);
                    // row of v: [j-1] [i-1] a2 a4 [i+1] [j+1]
                    float sv = 0.f;
                    sv = fmaf(el(a6, k), el(up.pv, k), sv);
                    sv = fmaf(el(a5, k), pl_v, sv);
                    sv = fmaf(el(a2, k), el(ce.pu, k), sv);
                    sv = fmaf(el(ce.a4, k), el(ce.pv, k), sv);
                    sv = fmaf(el(a7, k), pr_v, sv);
                    sv = fmaf(el(a8, k), el(dn.pv, k), sv);
                    const bool in = i0 + k < g.nx;
                    el(qu, k) = in ? su : 0.f;
                    el(qv, k) = in ? sv : 0.f;
                    if (in) part += el(ce.pu, k) * su + el(ce.pv, k) * sv;
                }
                st4(pu_new + off, ce.pu);
                st4(pv_new + off, ce.pv);
                st4(a.b.qu + off, qu);
                st4(a.b.qv + off, qv);
                dot[0] += (double)part;
            }
            up = ce;
            ce = dn;
        }
        if (a.store_halo && j_b == a.jb && j_b < g.ny && active) {     // ce now holds row j_b
            st4(pu_new + g.at(i0, j_b), ce.pu);
            st4(pv_new + g.at(i0, j_b), ce.pv);
        }
    }
    block_sum<1>(dot, red);
    double tot[1];
    if (grid_sum_finish<1>(dot, a.b.partials, a.b.ticket, tot, red)) {
        if (threadIdx.x == 0) {
            if (a.b.defer) a.b.pending[0] = tot[0];
            else a.b.scal->pAp = (float)tot[0];                        // pkTApk, :1165
        }
    }
}

struct P2Args {
    PcgBuffers b;
    Geom g;
    int ja, jb;
    int cur;           // pass 1 of this iteration wrote p[cur^1]
};

template <bool FIRST>
__global__ void __launch_bounds__(256) k_pcg_pass2(P2Args a)
{
    __shared__ double red[2 * 32];
    PcgScalars* s = a.b.scal;
    if (s->done) return;
    const float alphak = s->rz / s->pAp;                      // :1169
    const float nalpha = -1. * alphak;                        // :1174
    const Geom& g = a.g;
    const float* pu = a.b.pu[a.cur ^ 1];
    const float* pv = a.b.pv[a.cur ^ 1];
    const int upr = g.pitch >> 2;                             // float4 units per row
    const long long nunits = (long long)(a.jb - a.ja) * upr;
    double acc[2] = { 0.0, 0.0 };
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < nunits; t += (long long)gridDim.x * 256) {
        const int jr = (int)(t / upr), i0 = (int)(t - (long long)jr * upr) * 4;
        if (i0 >= g.nx) continue;
        const size_t off = g.at(i0, a.ja + jr);
        const float4 p_u = ld4(pu + off), p_v = ld4(pv + off);
        const float4 q_u = ld4_stream(a.b.qu + off), q_v = ld4_stream(a.b.qv + off);
        float4 r_u = ld4(a.b.ru + off), r_v = ld4(a.b.rv + off);
        const float4 a1 = ld4(a.b.coef[0] + off), a4 = ld4(a.b.coef[2] + off);
        float4 x_u = make_float4(0.f, 0.f, 0.f, 0.f), x_v = x_u;
        if (!FIRST) { x_u = ld4(a.b.xu + off); x_v = ld4(a.b.xv + off); }
        float prr = 0.f, prz = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < g.nx) {
                el(x_u, k) = fmaf(alphak, el(p_u, k), el(x_u, k));       // :1172
                el(x_v, k) = fmaf(alphak, el(p_v, k), el(x_v, k));
                const float ru = fmaf(nalpha, el(q_u, k), el(r_u, k));   // :1174
                const float rv = fmaf(nalpha, el(q_v, k), el(r_v, k));
                el(r_u, k) = ru;
                el(r_v, k) = rv;
                const float zu = (1.0f / el(a1, k)) * ru, zv = (1.0f / el(a4, k)) * rv;
                prr += ru * ru + rv * rv;                                // residc, :1178
                prz += zu * ru + zv * rv;                                // zktrk of the next iteration, :1142
            } else {
                el(x_u, k) = 0.f; el(x_v, k) = 0.f; el(r_u, k) = 0.f; el(r_v, k) = 0.f;
            }
        }
        st4(a.b.xu + off, x_u);
        st4(a.b.xv + off, x_v);
        st4(a.b.ru + off, r_u);
        st4(a.b.rv + off, r_v);
        acc[0] += (double)prr;
        acc[1] += (double)prz;
    }
    block_sum<2>(acc, red);
    double tot[2];
    if (grid_sum_finish<2>(acc, a.b.partials, a.b.ticket, tot, red)) {
        if (threadIdx.x == 0 && a.b.defer) {
            a.b.pending[0] = tot[0];
            a.b.pending[1] = tot[1];
        } else if (threadIdx.x == 0) {
            const float rr = (float)tot[0];
            s->rz_old = s->rz;                 // z0tr0 of the next iteration, :1135
            s->rz = (float)tot[1];
            s->rr = rr;
            s->its = s->its + 1;
            s->done = !(rr > s->tol);          // while((*residc) > tol ...), :1131
        }
    }
}

// Banded runs: the dot totals were summed over ranks by an all-reduce on
// `pending`; apply them exactly as the single-GPU kernels' last block does.
__global__ void k_finalize(PcgBuffers b, int kind, float tol)
{
    PcgScalars* s = b.scal;
    if (kind == FINALIZE_BUILD) {
        s->rr = (float)b.pending[0];
        s->rz = (float)b.pending[1];
        s->rz_old = 0.f;
        s->pAp = 0.f;
        s->tol = tol;
        s->its = 0;
        s->done = !((float)b.pending[0] > tol);
        return;
    }
    if (s->done) return;
    if (kind == FINALIZE_PASS1) {
        s->pAp = (float)b.pending[0];
    } else {
        const float rr = (float)b.pending[0];
        s->rz_old = s->rz;
        s->rz = (float)b.pending[1];
        s->rr = rr;
        s->its = s->its + 1;
        s->done = !(rr > s->tol);
    }
}

void launch_finalize(const PcgBuffers& b, int kind, float tol, cudaStream_t st)
{
    k_finalize<<<1, 1, 0, st>>>(b, kind, tol);
}

// u += x, v += x after a solve (:1185-1195); x is undefined when no iteration ran.
__global__ void __launch_bounds__(256)
k_update_uv(float* __restrict__ u, float* __restrict__ v, const float* __restrict__ xu,
            const float* __restrict__ xv, Geom g, int ja, int jb, const PcgScalars* s, int* its_out)
{
    if (blockIdx.x == 0 && threadIdx.x == 0 && its_out) *its_out = s->its;
    if (s->its == 0) return;
    const int upr = g.pitch >> 2;
    const long long nunits = (long long)(jb - ja) * upr;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < nunits; t += (long long)gridDim.x * 256) {
        const int jr = (int)(t / upr), i0 = (int)(t - (long long)jr * upr) * 4;
        if (i0 >= g.nx) continue;
        const size_t off = g.at(i0, ja + jr);
        float4 a = ld4(u + off), b = ld4(v + off);
        const float4 c = ld4(xu + off), d = ld4(xv + off);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (i0 + k < g.nx) { el(a, k) = el(a, k) + el(c, k); el(b, k) = el(b, k) + el(d, k); }
        st4(u + off, a);
        st4(v + off, b);
    }
}

// dst(pitched rows [ja,jb)) = scale * src(dense, row 0 == ja)
__global__ void __launch_bounds__(256)
k_scale_copy(const float* __restrict__ src, float* __restrict__ dst, Geom g, int ja, int jb, float scale)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int j = ja + blockIdx.y;
    if (i < g.nx && j < jb) dst[g.at(i, j)] = src[(size_t)(j - ja) * g.nx + i] * scale;
}

static int pass1_rows_per_task(int nstrips, int nrows, int sm_count)
{
    // enough warp tasks to fill the machine a few times over, at most 64 rows each
    long long want = (long long)sm_count * 16 * 4;
    long long rs = ((long long)nstrips * nrows + want - 1) / want;
    if (rs < 4) rs = 4;
    if (rs > 64) rs = 64;
    return (int)rs;
}

void launch_pcg_pass1(const PcgBuffers& b, const Geom& g, int ja, int jb, int first, int cur, int store_halo,
                      int sm_count, cudaStream_t st)
{
    P1Args a;
    a.b = b; a.g = g; a.ja = ja; a.jb = jb; a.cur = cur; a.store_halo = store_halo;
    a.nstrips = (g.nx + 127) / 128;
    a.rs = pass1_rows_per_task(a.nstrips, jb - ja, sm_count);
    a.nsegs = (jb - ja + a.rs - 1) / a.rs;
    const int ntasks = a.nstrips * a.nsegs;
    int grid = (ntasks + 7) / 8;
    const int cap = sm_count * 16;
    if (grid > cap) grid = cap;
    if (grid > b.max_partial_blocks) grid = b.max_partial_blocks;
    if (first) k_pcg_pass1<true><<<grid, 256, 0, st>>>(a);
    else       k_pcg_pass1<false><<<grid, 256, 0, st>>>(a);
}

void launch_pcg_pass2(const PcgBuffers& b, const Geom& g, int ja, int jb, int first, int cur, int sm_count,
                      cudaStream_t st)
{
    P2Args a;
    a.b = b; a.g = g; a.ja = ja; a.jb = jb; a.cur = cur;
    const long long nunits = (long long)(jb - ja) * (g.pitch >> 2);
    long long grid = (nunits + 255) / 256;
    const int cap = sm_count * 16;
    if (grid > cap) grid = cap;
    if (grid > b.max_partial_blocks) grid = b.max_partial_blocks;
    if (first) k_pcg_pass2<true><<<(int)grid, 256, 0, st>>>(a);
    else       k_pcg_pass2<false><<<(int)grid, 256, 0, st>>>(a);
}

void launch_update_uv(float* u, float* v, const float* xu, const float* xv, const Geom& g, int ja, int jb,
                      const PcgScalars* s, int* its_out, int sm_count, cudaStream_t st)
{
    const long long nunits = (long long)(jb - ja) * (g.pitch >> 2);
    long long grid = (nunits + 255) / 256;
    if (grid > sm_count * 16) grid = sm_count * 16;
    k_update_uv<<<(int)grid, 256, 0, st>>>(u, v, xu, xv, g, ja, jb, s, its_out);
}

void launch_scale_copy(const float* src, float* dst, const Geom& g, int ja, int jb, float scale, cudaStream_t st)
{
    if (jb <= ja) return;
    dim3 grid((g.nx + 255) / 256, jb - ja);
    k_scale_copy<<<grid, 256, 0, st>>>(src, dst, g, ja, jb, scale);
}

}  // namespace octane
