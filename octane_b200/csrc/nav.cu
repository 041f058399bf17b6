// nav.cu -- pixel displacement -> navigated u/v (m/s), packed to shorts.
// Replaces octnavcalcuda / oct_navpixel_uv_cuda / oct_haversine_cuda and the
// host loops of oct_pix2uv_cuda, src/oct_pix2uv_cuda.cu:13-221,265-370
// (reference tree).  The arithmetic is transcribed expression by expression:
// navigation parity is a rounding-reproduction problem (lat/lon are narrowed
// to float at the haversine call, :13,151,160), so the fp64 expression order
// is kept and nothing is "improved".  What changes is the data movement: the
// reference fills int index arrays and double copies of u,v on the host
// (:308-321, 24 B/px over PCIe/managed memory); here the kernel derives (i,j)
// from the thread index, reads the float flow directly and writes the four
// short planes (8 B in, 8 B out per pixel).
#include "kernels.cuh"

namespace octane {

// The reference packs U_raw/V_raw (and U/V with -pd, and CTP) on the HOST:
// (short)(float) on x86-64 is cvttss2si to int32 followed by truncation to 16 bits,
// which wraps for |x| >= 32768 (e.g. the -9999 fill value times 100) where the GPU's
// cvt.s16.f32 would saturate.  Reproduce the host behaviour.
__device__ __forceinline__ short host_short(float x) { return (short)(int)x; }

__device__ double oct_haversine(float lat1, float lon1, float lat2, float lon2, double rad, double rad2)
{
    const double earthrad = 6371000.00;
    double a, c, r, dlat, dlon;
    dlon = lon2 - lon1;
    dlat = lat2 - lat1;
    a = (pow(sin(dlat * rad2), 2) + cos(lat1 * rad) * cos(lat2 * rad) * pow((sin(dlon * rad2)), 2));
    c = 2. * atan2(sqrt(a), sqrt(1 - a));
    r = earthrad * c;
    return r;
}

__device__ void oct_navpixel_uv(const NavParams& geo, double* xv, int xi, int yi, double dt, double* r,
                                double DTOR, double DTOR2, bool dp, bool dm)
{
    const double PI = 3.14159265359;
    double xVal, yVal, dist;
    double latv[2], lonv[2], sds[2];
    sds[0] = 0.;
    sds[1] = 0.;
    if (dp) {                                   // polar orthographic, :34-67
        for (int iv = 0; iv < 2; ++iv) {
            if (iv == 0) {
                xVal = (xi)*geo.xScale + geo.xOffset;
                yVal = (yi)*geo.yScale + geo.yOffset;
            } else {
                xVal = (xv[0] * dt + xi) * geo.xScale + geo.xOffset;
                yVal = (xv[1] * dt + yi) * geo.yScale + geo.yOffset;
            }
            double rho = sqrt(xVal * xVal + yVal * yVal);
            double c = asin(rho / geo.R);
            if (geo.lat1 > 89.9999) {
                lonv[iv] = geo.lon0 * DTOR + atan2(xVal, -yVal);
            } else {
                lonv[iv] = geo.lon0 * DTOR + atan2(xVal * sin(c), (rho * cos(geo.lat1 * DTOR) * cos(c) - yVal * sin(geo.lat1 * DTOR) * sin(c)));
            }
            if (rho > 0.0000001) {
                latv[iv] = asin(cos(c) * sin(geo.lat1 * DTOR) + (yVal * sin(c) * cos(geo.lat1 * DTOR) / rho));
            } else {
                latv[iv] = geo.lat1 * DTOR;
            }
            latv[iv] = latv[iv] / DTOR;
            lonv[iv] = lonv[iv] / DTOR;
        }
    } else {
        if (dm) {                               // Mercator, :70-87
            for (int iv = 0; iv < 2; ++iv) {
                if (iv == 0) {
                    xVal = (xi)*geo.xScale + geo.xOffset;
                    yVal = (yi)*geo.yScale + geo.yOffset;
                } else {
                    xVal = (xv[0] * dt + xi) * geo.xScale + geo.xOffset;
                    yVal = (xv[1] * dt + yi) * geo.yScale + geo.yOffset;
                }
                latv[iv] = PI / 2. - 2. * atan(exp(-yVal / geo.R));
                lonv[iv] = xVal / geo.R + geo.lon1;
                latv[iv] = latv[iv] / DTOR;
                lonv[iv] = lonv[iv] / DTOR;
            }
        } else {                                // GOES fixed grid, :89-139
            double a, b, c, d, e, rs, sx, sy, sz;
            double H;
            H = geo.pph + geo.req;
            for (int iv = 0; iv < 2; ++iv) {
                if (iv == 0) {
                    xVal = (xi)*geo.xScale + geo.xOffset;
                    yVal = (yi)*geo.yScale + geo.yOffset;
                } else {
                    xVal = (xv[0] * dt + xi) * geo.xScale + geo.xOffset;
                    yVal = (xv[1] * dt + yi) * geo.yScale + geo.yOffset;
                }
                sds[iv] = xVal * xVal + yVal * yVal;
                a = pow((sin(xVal)), 2) + pow(cos(xVal), 2) * (pow((cos(yVal)), 2) + (pow(geo.req, 2)) / (pow(geo.rpol, 2)) * pow((sin(yVal)), 2));
                b = -2. * H * cos(xVal) * cos(yVal);
                c = pow(H, 2) - pow(geo.req, 2);
                d = (pow(b, 2) - 4. * a * c);
                if (d >= 0) {
                    rs = (-b - sqrt(d)) / (2. * a);
                    sx = rs * cos(xVal) * cos(yVal);
                    sy = -rs * sin(xVal);
                    sz = rs * cos(xVal) * sin(yVal);
                    e = (pow((H - sx), 2) + pow(sy, 2));
                    if (sz == 0 || e <= 0 || H - sx == 0) {
                        latv[iv] = -999.;
                        lonv[iv] = -999.;
                    } else {
                        latv[iv] = atan((pow(geo.req, 2)) / (pow(geo.rpol, 2)) * (sz / sqrt(e)));
                        lonv[iv] = geo.lam0 - atan(sy / (H - sx));
                        latv[iv] = latv[iv] / DTOR;
                        lonv[iv] = lonv[iv] / DTOR;
                    }
                } else {
                    latv[iv] = -999.;
                    lonv[iv] = -999.;
                }
            }
        }
    }
    // :144-168
    if ((latv[0] < -998) || (latv[1] < -998) || (sds[0] > 0.021)) {
        r[0] = 0.;
        r[1] = 0.;
    } else {
        dist = oct_haversine(latv[0], lonv[0], latv[0], lonv[1], DTOR, DTOR2);
        if (lonv[1] >= lonv[0]) r[0] = dist / dt;
        else r[0] = -dist / dt;
        dist = oct_haversine(latv[0], lonv[0], latv[1], lonv[0], DTOR, DTOR2);
        if (latv[1] >= latv[0]) r[1] = dist / dt;
        else r[1] = -dist / dt;
    }
}

// one thread per pixel; rows [row0,row0+nrows) of an nx-wide scene, arrays hold just those rows
__global__ void __launch_bounds__(256)
k_pix2uv(NavParams nav, const float* __restrict__ u, const float* __restrict__ v, int nx, int row0, int nrows,
         short* __restrict__ ur, short* __restrict__ vr, short* __restrict__ ur2, short* __restrict__ vr2)
{
    const double pi = 3.14159265;
    const double DTOR = pi / 180.;
    const double DTOR2 = DTOR / 2.;
    const size_t n = (size_t)nx * nrows;
    for (size_t lxyz = (size_t)blockIdx.x * 256 + threadIdx.x; lxyz < n; lxyz += (size_t)gridDim.x * 256) {
        const int jj = (int)(lxyz / nx) + row0, ii = (int)(lxyz % nx);
        const float uf = u[lxyz], vf = v[lxyz];
        ur2[lxyz] = host_short(100 * uf);                  // :335-336
        vr2[lxyz] = host_short(100 * vf);
        if (nav.pixuv) {                                   // :348-356
            ur[lxyz] = host_short(100 * uf);
            vr[lxyz] = host_short(100 * vf);
            continue;
        }
        double dans[2], xans[2];
        const double u1 = uf, v1 = vf;                     // float -> double on the host in the reference, :317-318
        if (u1 > -9998.) {
            dans[0] = u1 / (nav.t2 - nav.t1);
            dans[1] = v1 / (nav.t2 - nav.t1);
            oct_navpixel_uv(nav, dans, ii + nav.minX, jj + nav.minY, nav.t2 - nav.t1, xans, DTOR, DTOR2,
                            nav.dp != 0, nav.dm != 0);
            ur[lxyz] = (short)(100 * (xans[0]));
            vr[lxyz] = (short)(100 * (xans[1]));
        } else {
            ur[lxyz] = (short)(-32768);
            vr[lxyz] = (short)(-32768);
        }
    }
}

// CTP pack of oct_optical_flow.cc:71-88
__global__ void __launch_bounds__(256) k_ctp_pack(const float* __restrict__ cth, short* __restrict__ ctp, size_t n, int ir)
{
    for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256)
        ctp[k] = ir ? host_short((cth[k] - 300) * 100) : host_short(cth[k]);
}

void launch_pix2uv(const NavParams& np, const float* u, const float* v, int nx, int row0, int nrows,
                   short* U, short* V, short* Uraw, short* Vraw, cudaStream_t st)
{
    const size_t n = (size_t)nx * nrows;
    if (!n) return;
    size_t grid = (n + 255) / 256;
    if (grid > 148 * 64) grid = 148 * 64;
    k_pix2uv<<<(unsigned)grid, 256, 0, st>>>(np, u, v, nx, row0, nrows, U, V, Uraw, Vraw);
}

void launch_ctp_pack(const float* cth, short* ctp, size_t n, int ir, cudaStream_t st)
{
    if (!n) return;
    size_t grid = (n + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    k_ctp_pack<<<(unsigned)grid, 256, 0, st>>>(cth, ctp, n, ir);
}

}  // namespace octane
