// ingest.cu -- the stages either side of the flow path that the reference also runs on the GPU:
//   k_navcal : Rad counts -> radiance -> (optional calibration) -> limb taper -> 0..255
//              normalisation, plus fixed-grid latitude / longitude of every pixel.
//              Replaces octnavcalcuda, src/oct_navcal_cuda.cu:12-98 (reference tree).
//   k_uv2pix : first-guess winds (m/s) -> pixel displacements by the haversine destination
//              formula and the forward fixed-grid projection.
//              Replaces octuv2xy + the host loops of oct_uv2pix, src/oct_pix2uv_cuda.cu:223-263,372-476.
// Both are one pass over the scene (2 B in, 12 B out per pixel for navcal; 16 B in, 8 B out for
// uv2pix); navcal with navigation is bound by its ~10 fp64 transcendentals per pixel, not by HBM.
// The reference stages every operand through managed memory filled by single-threaded host
// loops (including per-pixel index arrays icarr/jcarr/lxyzarr); here the pixel's (i, j) come
// from the thread index and the coordinate vectors x[], y[] are read directly.
// Expressions keep the reference's float/double promotion points.
#include "kernels.cuh"

namespace octane {

__global__ void __launch_bounds__(256)
k_navcal(const short* __restrict__ rad, const short* __restrict__ x, const short* __restrict__ y, int nx, int ny,
         CalParams c, float* __restrict__ data3, float* __restrict__ lat, float* __restrict__ lon)
{
    const double PI = 3.14159265359;
    const double DTOR = PI / 180.;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t lxyz = (size_t)j * nx + i;
    const float xScale = c.xScale, xOffset = c.xOffset, yScale = c.yScale, yOffset = c.yOffset;
    const float req = c.req, rpol = c.rpol, H = c.H, lam0 = c.lam0;
    // :31-34 (float arithmetic, then widened)
    double xVal = x[i] * xScale + xOffset;
    double yVal = y[j] * yScale + yOffset;
    double subpoint_dist = xVal * xVal + yVal * yVal;
    float dVal = rad[lxyz] * c.radScale + c.radOffset;
    if (lat && lon) {
        if (c.donav == 1) {          // :36-49
            double a, b, cc, rs, sx, sy, sz;
            a = pow((sin(xVal)), 2) + pow(cos(xVal), 2) * (pow((cos(yVal)), 2) + (pow(req, 2)) / (pow(rpol, 2)) * pow((sin(yVal)), 2));
            b = -2. * H * cos(xVal) * cos(yVal);
            cc = pow(H, 2) - pow(req, 2);
            rs = (-b - sqrt((pow(b, 2) - 4. * a * cc))) / (2. * a);
            sx = rs * cos(xVal) * cos(yVal);
            sy = -rs * sin(xVal);
            sz = rs * cos(xVal) * sin(yVal);
            float la = atan(double((pow(req, 2)) / (pow(rpol, 2))) * (sz / sqrt((pow((H - sx), 2) + pow(sy, 2)))));
            float lo = lam0 - atan(sy / (H - sx));
            la = la / DTOR;
            lo = lo / DTOR;
            lat[lxyz] = la;
            lon[lxyz] = lo;
        } else {
            lat[lxyz] = 0.;
            lon[lxyz] = 0.;
        }
    }
    double dataF;
    if (c.cal == 1) dataF = (c.fk2 / (log((c.fk1 / dVal) + 1.)) - c.bc1) / c.bc2;     // :59-63
    else if (c.cal == 2) dataF = c.kap1 * dVal;                                         // :64-68
    else dataF = dVal;                                                                  // RAW / BRIT / default
    // limb taper, :80-91
    float sdsconst;
    if (subpoint_dist < 0.021) {
        sdsconst = 1.;
    } else {
        if (subpoint_dist >= 0.0212) sdsconst = 0.;
        else sdsconst = c.subpoint_slope * subpoint_dist + c.subpoint_int;
    }
    // :93
    data3[lxyz] = sdsconst * (((dataF - c.minin) / (c.maxin - c.minin)) * (c.maxout - c.minout) + c.minout);
}

void launch_navcal(const short* rad, const short* x, const short* y, int nx, int ny, const CalParams& c,
                   float* data, float* lat, float* lon, cudaStream_t st)
{
    dim3 grid((nx + 255) / 256, ny);
    k_navcal<<<grid, 256, 0, st>>>(rad, x, y, nx, ny, c, data, lat, lon);
}

// One thread per pixel; u, v in: first-guess wind (m/s); out: displacement in pixels over `secs`.
__global__ void __launch_bounds__(256)
k_uv2pix(float* __restrict__ u, float* __restrict__ v, const float* __restrict__ lat, const float* __restrict__ lon,
         const short* __restrict__ xs, const short* __restrict__ ys, int nx, int ny, Uv2PixParams q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t lxyz = (size_t)j * nx + i;
    const double R = 6371000.0;
    const double pi = 3.14159265;
    double rad = pi / 180.;
    double H = q.pph + q.req;
    // :236-244 (the host loop widened u, v, lat, lon to double, :408-413)
    double u1 = u[lxyz];
    double v1 = v[lxyz];
    double latvalv = lat[lxyz];
    double lonvalv = lon[lxyz];
    double dist = sqrt(pow(u1, 2.0) + pow(v1, 2.0)) * (q.secs);
    double brng = (180. + (90. - (atan2(-v1, -u1) / rad))) * rad;
    double latorig = latvalv * rad;
    latvalv = asin(sin(latorig) * cos(dist / R) + cos(latorig) * sin(dist / R) * cos(brng));
    lonvalv = lonvalv * rad + (atan2((sin(brng) * sin(dist / R) * cos(latorig)), (cos(dist / R) - sin(latorig) * sin(latvalv))));
    // :247-252
    double thtc = atan(((q.rpol2) / (q.req2)) * tan(latvalv));
    double rc = q.rpol / sqrt(1. - (q.eval) * pow(cos(thtc), 2.));
    double sx = H - rc * cos(thtc) * cos(lonvalv - q.lam0);
    double sy = -rc * cos(thtc) * sin(lonvalv - q.lam0);
    double sz = rc * sin(thtc);
    double x1v, y1v;
    if ((H * (H - sx)) >= (sy * sy + ((q.req2) / (q.rpol2) * sz * sz))) {        // :253-261
        x1v = (asin(-sy / (sqrt(sx * sx + sy * sy + sz * sz))) - q.xoffset) / q.xscale;
        y1v = (atan(sz / sx) - q.yoffset) / q.yscale;
    } else {
        x1v = -999.;
        y1v = -999.;
    }
    // :451-462
    if (x1v > -998.) {
        u[lxyz] = x1v - xs[i];
        v[lxyz] = y1v - ys[j];
    } else {
        u[lxyz] = 0.;
        v[lxyz] = 0.;
    }
}

void launch_uv2pix(float* u, float* v, const float* lat, const float* lon, const short* xs, const short* ys, int nx,
                   int ny, const Uv2PixParams& q, cudaStream_t st)
{
    dim3 grid((nx + 255) / 256, ny);
    k_uv2pix<<<grid, 256, 0, st>>>(u, v, lat, lon, xs, ys, nx, ny, q);
}


// ---- lat/lon of the two projected grids the reference also ingests: orthographic polar
// (octpolarnavcalcuda, src/oct_polar_navcal_cuda.cu:12-66) and spherical Mercator (octmercnavcalcuda,
// src/oct_merc_navcal_cuda.cu:12-48).  The image passes through unchanged (these files hold floats that
// are already normalised); lat1 / lon0 arrive in radians, narrowed to float, as the reference's host
// wrappers pass them (:141-143 / :122-124) -- including the pole test `lat1 > 89.99999`, which the reference
// applies to the radian value and which therefore never fires.
template <int GRID>   // 1 polar, 2 Mercator
__global__ void __launch_bounds__(256)
k_navcal_grid(const float* __restrict__ data2, const short* __restrict__ x, const short* __restrict__ y, int nx, int ny,
              float xScale, float xOffset, float yScale, float yOffset, float R, float lon0, float lat1, int donav,
              float* __restrict__ data3, float* __restrict__ lat, float* __restrict__ lon)
{
    const double PI = 3.14159265359;
    const double DTOR = PI / 180.;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t lxyz = (size_t)j * nx + i;
    double xVal = x[i] * xScale + xOffset;
    double yVal = y[j] * yScale + yOffset;
    float dVal = data2[lxyz];
    if (lat && lon) {
        float la, lo;
        if (donav == 1) {
            if (GRID == 1) {
                double rho, c;
                rho = sqrt(xVal * xVal + yVal * yVal);
                c = asin(rho / R);
                if (lat1 > 89.99999) {
                    lo = lon0 + atan2(xVal, -yVal);
                } else {
                    lo = lon0 + atan2(xVal * sin(c), (rho * cos(lat1) * cos(c) - yVal * sin(lat1) * sin(c)));
                }
                if (rho > 0.0000001) {
                    la = asin(cos(c) * sin(lat1) + (yVal * sin(c) * cos(lat1) / rho));
                } else {
                    la = lat1;
                }
            } else {
                lo = xVal / R + lon0;
                la = PI / 2. - 2. * atan(exp(-yVal / R));
            }
            la = la / DTOR;
            lo = lo / DTOR;
        } else {
            la = 0.;
            lo = 0.;
        }
        lat[lxyz] = la;
        lon[lxyz] = lo;
    }
    data3[lxyz] = dVal;
}

void launch_navcal_grid(int grid_kind, const float* data2, const short* x, const short* y, int nx, int ny, float xScale,
                        float xOffset, float yScale, float yOffset, float R, float lon0_rad, float lat1_rad, int donav,
                        float* data3, float* lat, float* lon, cudaStream_t st)
{
    dim3 grid((nx + 255) / 256, ny);
    if (grid_kind == 1)
        k_navcal_grid<1><<<grid, 256, 0, st>>>(data2, x, y, nx, ny, xScale, xOffset, yScale, yOffset, R, lon0_rad, lat1_rad, donav, data3, lat, lon);
    else
        k_navcal_grid<2><<<grid, 256, 0, st>>>(data2, x, y, nx, ny, xScale, xOffset, yScale, yOffset, R, lon0_rad, lat1_rad, donav, data3, lat, lon);
}

// ---- regridding of an ancillary field onto the image grid: oct_zoom_in_float, src/oct_zoom.cc:180-222,
// with oct_bicubic_float / oct_cell, src/oct_bicubic.cc:12-29,100-150 (bicubic when interp == 1, nearest
// neighbour otherwise).  The reader uses it for cloud-top heights and extra channels that come on a
// coarser grid than channel 1 (src/oct_fileread.cc:370,796).  One thread per output pixel; the 16 taps
// of a 4x upsampling hit the same cache lines for neighbouring threads, so the pass is bound by the 4 B
// written per pixel.
__device__ __forceinline__ double zcell(double v0, double v1, double v2, double v3, double x)
{
    return v1 + 0.5 * x * (v2 - v0 + x * (2.0 * v0 - 5.0 * v1 + 4.0 * v2 - v3 + x * (3.0 * (v1 - v2) + v3 - v0)));
}
__device__ __forceinline__ int zbc(int x, int n) { return x < 0 ? 0 : (x >= n ? n - 1 : x); }

__global__ void __launch_bounds__(256)
k_zoom_in_float(const float* __restrict__ in, int nx, int ny, float* __restrict__ out, int nxx, int nyy, int interp)
{
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int jj1 = blockIdx.y;
    if (i1 >= nxx || jj1 >= nyy) return;
    const float factorx = ((float)nxx / nx);
    const float factory = ((float)nyy / ny);
    const float val1 = (0.5 - 0.5 / factory);
    const float val2 = (0.5 - 0.5 / factorx);
    const float j2 = (float)((jj1 / factory) - val1);
    const float i2 = (float)((i1 / factorx) - val2);
    float g;
    if (interp == 1) {
        const double uu = i2, vv = j2;
        const int x = zbc((int)uu, nx), y = zbc((int)vv, ny);
        const int mx = zbc((int)(uu - 1), nx), my = zbc((int)(vv - 1), ny);
        const int dx = zbc((int)(uu + 1), nx), dy = zbc((int)(vv + 1), ny);
        const int ddx = zbc((int)(uu + 2), nx), ddy = zbc((int)(vv + 2), ny);
        const int cols[4] = { mx, x, dx, ddx };
        const size_t rows[4] = { (size_t)nx * my, (size_t)nx * y, (size_t)nx * dy, (size_t)nx * ddy };
        double v[4];
#pragma unroll
        for (int c = 0; c < 4; c++)          // pol[c][r] = input[cols[c] + rows[r]]; first along y, then along x
            v[c] = zcell(in[cols[c] + rows[0]], in[cols[c] + rows[1]], in[cols[c] + rows[2]], in[cols[c] + rows[3]], vv - y);
        g = zcell(v[0], v[1], v[2], v[3], uu - x);
    } else {
        const int j3 = int(j2 + 0.5), i3 = int(i2 + 0.5);
        g = in[i3 + (size_t)nx * j3];
    }
    out[i1 + (size_t)nxx * jj1] = g;
}

void launch_zoom_in_float(const float* in, int nx, int ny, float* out, int nxx, int nyy, int interp, cudaStream_t st)
{
    dim3 grid((nxx + 255) / 256, nyy);
    k_zoom_in_float<<<grid, 256, 0, st>>>(in, nx, ny, out, nxx, nyy, interp);
}

// ---- down-scaling of a FINER ancillary field onto the image grid: oct_zoom_out_float,
// src/oct_zoom.cc:51-88 = oct_gaussian (src/oct_gaussian.cc:48-104) on a double copy of the field, then
// oct_bicubic (src/oct_bicubic.cc:36-97) at (ii / factor, jj / factor).  The reference does this on the
// CPU in double; the three passes below keep its operation order and use the explicitly rounded
// __dmul_rn / __dadd_rn so that nvcc cannot contract a product into the following sum: the outputs are
// bit-identical to the reference's CPU object (x86-64, no FMA).  The taps come from the host (the
// reference's exp() is glibc's).  Horizontal pass: 4 B read + 8 B written per input pixel; vertical pass
// 8 + 8; sampling 16 taps per OUTPUT pixel (1 / factor^2 fewer than inputs).
__device__ __forceinline__ double zcell_rn(double v0, double v1, double v2, double v3, double x)
{
    // v1 + 0.5*x*(v2 - v0 + x*(2.0*v0 - 5.0*v1 + 4.0*v2 - v3 + x*(3.0*(v1 - v2) + v3 - v0))), src/oct_bicubic.cc:10-18
    const double t3 = __dsub_rn(__dadd_rn(__dmul_rn(3.0, __dsub_rn(v1, v2)), v3), v0);
    const double t2 = __dadd_rn(__dsub_rn(__dadd_rn(__dsub_rn(__dmul_rn(2.0, v0), __dmul_rn(5.0, v1)), __dmul_rn(4.0, v2)), v3),
                                __dmul_rn(x, t3));
    const double t1 = __dadd_rn(__dsub_rn(v2, v0), __dmul_rn(x, t2));
    return __dadd_rn(v1, __dmul_rn(__dmul_rn(0.5, x), t1));
}

template <typename TIn, bool VERT>
__global__ void __launch_bounds__(256)
k_zoomout_blur(const TIn* __restrict__ in, double* __restrict__ out, int nx, int ny, ZoomOutTaps t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nx || j >= ny) return;
    double wsum = 0;
    for (int k = -t.R; k < t.R; ++k) {              // the +R tap is dropped, src/oct_gaussian.cc:70,91
        const size_t at = VERT ? (size_t)i + (size_t)nx * zbc(j + k, ny) : (size_t)zbc(i + k, nx) + (size_t)nx * j;
        wsum = __dadd_rn(wsum, __dmul_rn(t.gk[k + t.R], (double)in[at]));
    }
    out[i + (size_t)nx * j] = wsum;
}

__global__ void __launch_bounds__(256)
k_zoomout_sample(const double* __restrict__ Is, int nx, int ny, float* __restrict__ out, int nxx, int nyy, double factor)
{
    const int ii = blockIdx.x * blockDim.x + threadIdx.x;
    const int jj = blockIdx.y;
    if (ii >= nxx || jj >= nyy) return;
    const double uu = (double)ii / factor, vv = (double)jj / factor;
    const int x = zbc((int)uu, nx), y = zbc((int)vv, ny);
    const int cols[4] = { zbc((int)(uu - 1), nx), x, zbc((int)(uu + 1), nx), zbc((int)(uu + 2), nx) };
    const size_t rows[4] = { (size_t)nx * zbc((int)(vv - 1), ny), (size_t)nx * y, (size_t)nx * zbc((int)(vv + 1), ny),
                             (size_t)nx * zbc((int)(vv + 2), ny) };
    const double fy = __dsub_rn(vv, (double)y), fx = __dsub_rn(uu, (double)x);
    double v[4];
#pragma unroll
    for (int c = 0; c < 4; c++)
        v[c] = zcell_rn(Is[cols[c] + rows[0]], Is[cols[c] + rows[1]], Is[cols[c] + rows[2]], Is[cols[c] + rows[3]], fy);
    out[ii + (size_t)nxx * jj] = (float)zcell_rn(v[0], v[1], v[2], v[3], fx);
}

__global__ void __launch_bounds__(256)
k_zoomout_copy(const float* __restrict__ in, float* __restrict__ out, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[k];
}

int launch_zoom_out_float(const float* in, int nx, int ny, float* out, int nxx, int nyy, double factor,
                          const ZoomOutTaps* taps, double* tmp_a, double* tmp_b, cudaStream_t st)
{
    if (!taps) {       // factor >= 0.999999: in[ii + nxx*jj] -> out[ii + nxx*jj], src/oct_zoom.cc:79-81
        const size_t n = (size_t)nxx * nyy;
        k_zoomout_copy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
        return 1;
    }
    dim3 grid((nx + 255) / 256, ny);
    k_zoomout_blur<float, false><<<grid, 256, 0, st>>>(in, tmp_a, nx, ny, *taps);
    k_zoomout_blur<double, true><<<grid, 256, 0, st>>>(tmp_a, tmp_b, nx, ny, *taps);
    dim3 grid2((nxx + 255) / 256, nyy);
    k_zoomout_sample<<<grid2, 256, 0, st>>>(tmp_b, nx, ny, out, nxx, nyy, factor);
    return 3;
}

}  // namespace octane
