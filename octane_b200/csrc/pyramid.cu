// pyramid.cu -- Gaussian pyramid (blur + decimate), 4th-order gradients and
// bicubic flow prolongation.  Replaces the device functions fill_GK, convh,
// convv, zoom_out, oct_compgrad_cu, zoom_in, oct_bicubic_cu of
// src/oct_variational_optical_flow.cu:208-466 (reference tree).
#include "kernels.cuh"

namespace octane {

// ---- Gaussian taps, :208-228 (one thread, same expression order) -----------
__global__ void k_fill_gk(float* GK, float factor, int R)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float sigma, r, s;
    sigma = 0.6 * sqrt(1.0 / (factor * factor) - 1.0);
    s = 2.0 * sigma * sigma;
    float sum = 0.0;
    for (int x = -R; x <= R; x++) {
        r = x;
        GK[x + R] = (exp(-(r * r) / s)) / (3.14159265358979323846 * s);
        sum += GK[x + R];
    }
    for (int i = 0; i < 2 * R + 1; ++i) GK[i] /= sum;
}

// ---- fused convh + convv + zoom_out, :312-408 --------------------------------
// The reference blurs the whole full-resolution plane (taps kk in [-R,R): the +R
// tap is dropped, :322,344) and then samples it with a bicubic at integer
// coordinates, which is an exact decimation (oct_cell_cu(v,0) == v[1], :236).
// Here each output pixel evaluates the separable blur only where it is sampled,
// with the same tap order (horizontal sums first, then the vertical sum), so
// the result is bit-identical while the full-res blurred planes never exist.
// A block computes a (TX x TY) output tile: the horizontal sums for the
// TY-tile's (TY-1)*step + 2R source rows are staged in shared memory.
template <int TX, int TY>
__global__ void __launch_bounds__(TX* TY)
k_blur_decimate(const float* __restrict__ src, Geom gs, float* __restrict__ dst, Geom gd,
                int ja, int jb, float factor, const float* __restrict__ GK, int R, float scale, int nc)
{
    extern __shared__ float sm[];
    float* gk = sm;                  // 2R+1 taps
    float* hs = sm + 2 * R + 1;      // [nrow_h][TX] horizontal sums
    const int tid = threadIdx.y * TX + threadIdx.x;
    for (int t = tid; t < 2 * R + 1; t += TX * TY) gk[t] = GK[t];
    const int ii = blockIdx.x * TX + threadIdx.x;
    const int jj0 = ja + blockIdx.y * TY;
    // Reference quirk kept for parity: zoom_out reads the blurred plane without a channel
    // offset (:406), so every coarse-level channel is channel 0's blurred, decimated image.
    dst += (size_t)blockIdx.z * gd.plane;
    const int i2 = (int)(ii / factor);                          // :369
    const int jj_last = min(jj0 + TY, jb) - 1;
    const int jsrc0 = (int)(jj0 / factor) - R;                  // first source row of the tile
    const int jsrc1 = (int)(jj_last / factor) + R;              // one past the last
    const int nrow_h = jsrc1 - jsrc0;
    __syncthreads();
    if (ii < gd.nx) {
        for (int rr = threadIdx.y; rr < nrow_h; rr += TY) {
            int js = clampi(jsrc0 + rr, gs.ny);                 // global clamp-to-edge
            js = min(max(js, gs.jlo()), gs.jhi() - 1);          // stay inside the local band
            const float* row = src + gs.at(0, js);
            float wsum = 0;
            for (int kk = -R; kk < R; ++kk)
                wsum = fmaf(gk[kk + R], __ldg(row + clampi(i2 + kk, gs.nx)), wsum);
            hs[rr * TX + threadIdx.x] = wsum;
        }
    }
    __syncthreads();
    const int jj = jj0 + threadIdx.y;
    if (ii < gd.nx && jj < jb) {
        const int j2 = (int)(jj / factor);
        // rows j2-R .. j2+R-1; a clamped source row repeats the clamped row's sum,
        // exactly as convv reading the clamped row of the convh output (:346-347)
        float wsum = 0;
        for (int kk = -R; kk < R; ++kk) {
            int rr = (j2 + kk) - jsrc0;
            wsum = fmaf(gk[kk + R], hs[rr * TX + threadIdx.x], wsum);
        }
        dst[gd.at(ii, jj)] = wsum * scale;
    }
}

// ---- 4th-order central differences, :411-449 ---------------------------------
// numerator in double (the literal 8. promotes it), divided by 12.0, stored float.
__global__ void __launch_bounds__(256)
k_gradient(const float* __restrict__ f, float* __restrict__ gx, float* __restrict__ gy, Geom g,
           int ja, int jb, int nc)
{
    const int i = blockIdx.x * 32 + threadIdx.x;
    const int j = ja + blockIdx.y * 8 + threadIdx.y;
    if (i >= g.nx || j >= jb) return;
    const size_t coff = (size_t)blockIdx.z * g.plane;
    f += coff; gx += coff; gy += coff;
    const int lo = g.jlo(), hi = g.jhi() - 1;
    const int jp1 = min(max(clampi(j + 1, g.ny), lo), hi), jp2 = min(max(clampi(j + 2, g.ny), lo), hi);
    const int jm1 = min(max(clampi(j - 1, g.ny), lo), hi), jm2 = min(max(clampi(j - 2, g.ny), lo), hi);
    const int ip1 = clampi(i + 1, g.nx), ip2 = clampi(i + 2, g.nx);
    const int im1 = clampi(i - 1, g.nx), im2 = clampi(i - 2, g.nx);
    const float* r = f + g.at(0, j);
    gx[g.at(i, j)] = (-r[ip2] + 8. * r[ip1] - 8. * r[im1] + r[im2]) / 12.0;
    gy[g.at(i, j)] = (-f[g.at(i, jp2)] + 8. * f[g.at(i, jp1)] - 8. * f[g.at(i, jm1)] + f[g.at(i, jm2)]) / 12.0;
}

// ---- bicubic, :231-309 ---------------------------------------------------------
__device__ __forceinline__ float oct_cell(const float v[4], float x)
{
    return v[1] + 0.5 * x * (v[2] - v[0] +
           x * (2.0 * v[0] - 5.0 * v[1] + 4.0 * v[2] - v[3] +
           x * (3.0 * (v[1] - v[2]) + v[3] - v[0])));
}

__device__ __forceinline__ float bicubic(const float* __restrict__ in, const Geom& g, float uu, float vv)
{
    // tap indices: (int)-truncation of the float coordinate, then clamp (:266-273)
    const int x = clampi((int)uu, g.nx), y = clampi((int)vv, g.ny);
    const int mx = clampi((int)(uu - 1), g.nx), my = clampi((int)(vv - 1), g.ny);
    const int dx = clampi((int)(uu + 1), g.nx), dy = clampi((int)(vv + 1), g.ny);
    const int ddx = clampi((int)(uu + 2), g.nx), ddy = clampi((int)(vv + 2), g.ny);
    const int xs[4] = { mx, x, dx, ddx };
    const int ys[4] = { my, y, dy, ddy };
    float v[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        float p[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int jr = min(max(ys[b], g.jlo()), g.jhi() - 1);
            p[b] = __ldg(in + g.at(xs[a], jr));
        }
        v[a] = oct_cell(p, vv - y);
    }
    return oct_cell(v, uu - x);
}

// ---- flow prolongation, :453-466 -------------------------------------------------
__global__ void __launch_bounds__(256)
k_zoom_in(const float* __restrict__ flow, Geom gc, float* __restrict__ out, Geom gf, int ja, int jb, float sf)
{
    const int ii = blockIdx.x * 32 + threadIdx.x;
    const int jj = ja + blockIdx.y * 8 + threadIdx.y;
    if (ii >= gf.nx || jj >= jb) return;
    const float factorx = ((float)gf.nx / gc.nx);
    const float factory = ((float)gf.ny / gc.ny);
    float i2 = (float)((ii / factorx) - (0.5 - 0.5 / factorx));
    float j2 = (float)((jj / factory) - (0.5 - 0.5 / factory));
    out[gf.at(ii, jj)] = bicubic(flow, gc, i2, j2) / sf;
}

// ---- dense <-> pitched copies, u += x ---------------------------------------------
__global__ void __launch_bounds__(256)
k_zero_rows(float* __restrict__ a, Geom g, int ja, int jb)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int j = ja + blockIdx.y;
    if (i < g.pitch && j < jb) a[g.at(i, j)] = 0.f;
}

// ---- host launchers ------------------------------------------------------------------
void launch_fill_gk(float* GK, float factor, int R, cudaStream_t st)
{
    k_fill_gk<<<1, 32, 0, st>>>(GK, factor, R);
}

size_t blur_decimate_smem_bytes(float factor, int R)
{
    constexpr int TX = 32, TY = 8;
    const int step = (int)(1.0f / factor) + 1;
    const int nrow_h = (TY - 1) * step + 2 * R + 2;
    return sizeof(float) * ((size_t)nrow_h * TX + 2 * R + 1);
}

void launch_blur_decimate(const float* src, const Geom& gs, float* dst, const Geom& gd, int ja, int jb,
                          float factor, const float* GK, int R, float scale, int nc, cudaStream_t st)
{
    if (jb <= ja) return;
    constexpr int TX = 32, TY = 8;
    const size_t smem = blur_decimate_smem_bytes(factor, R);
    dim3 grid((gd.nx + TX - 1) / TX, (jb - ja + TY - 1) / TY, nc), block(TX, TY);
    static unsigned long long configured = 0;
    if (first_launch_on_device(&configured))
        cudaFuncSetAttribute(k_blur_decimate<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BLUR_DECIMATE_SMEM_LIMIT);
    k_blur_decimate<TX, TY><<<grid, block, smem, st>>>(src, gs, dst, gd, ja, jb, factor, GK, R, scale, nc);
}

void launch_gradient(const float* f, float* gx, float* gy, const Geom& g, int ja, int jb, int nc, cudaStream_t st)
{
    if (jb <= ja) return;
    dim3 grid((g.nx + 31) / 32, (jb - ja + 7) / 8, nc), block(32, 8);
    k_gradient<<<grid, block, 0, st>>>(f, gx, gy, g, ja, jb, nc);
}

void launch_zoom_in(const float* flow, const Geom& gc, float* out, const Geom& gf, int ja, int jb, float sf,
                    cudaStream_t st)
{
    if (jb <= ja) return;
    dim3 grid((gf.nx + 31) / 32, (jb - ja + 7) / 8), block(32, 8);
    k_zoom_in<<<grid, block, 0, st>>>(flow, gc, out, gf, ja, jb, sf);
}

}  // namespace octane
