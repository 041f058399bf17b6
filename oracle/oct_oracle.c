/* oracle/oct_oracle.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * CPU restatement of OCTANE's dense variational optical-flow path, one C
 * function per stage, each citing the reference lines it follows (paths are
 * relative to /root/reference).  It exists so that the CUDA kernels in
 * octane_b200/csrc can be checked on the same seeded inputs; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.
 *
 * Parity pin: the reference ships no tests, fixtures or golden vectors
 * (SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE
 * REFERENCE ITSELF: oracle/_ref/libref_cuda.so (the unmodified reference
 * sources recompiled for sm_100) run on a B200 over the seeded inputs of
 * tests/golden/make_golden.py; the resulting fixtures live in tests/golden/
 * and tests/test_oracle_golden.py checks this file against them.
 *
 * What is restated exactly: pyramid sizes and decimation indices, the
 * dropped +R blur tap, integer truncations in the bicubic tap selection,
 * mirror-without-repeat flow neighbourhoods, clamp + derivative zeroing of the
 * warp, every float/double promotion point of the coefficient build, the
 * boundary-merged stencil entries that the reference encodes in its CSR rows,
 * the PCG recurrence order, stop rule (||r||^2 <= 1e-8 or the iteration cap)
 * and the float narrowing inside the navigation haversine.
 * What is deliberately different: dot products are accumulated in double in a
 * fixed order (the reference uses float atomics and is not run-to-run
 * reproducible), and the matrix is applied stencil-wise instead of through CSR
 * arrays (same entries, same summation order per row).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifndef M_PI            /* -std=c99 hides it; the value is glibc's */
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    double alpha, lambda, lambdac, scaleF;
    int kiters, liters, cgiters, dozim;
} oracle_params;

typedef struct {
    double pph, req, rpol, lam0;
    float xScale, xOffset, yScale, yOffset, g2xOffset, g2yOffset;
    float lat1, lon1, lon0, R;
    int minX, minY;
} oracle_nav;

/* src/oct_variational_optical_flow.cu:26-41 (oct_bc_cu) */
static inline float bc_clamp(float x, int nx, int *bc)
{
    *bc = 0;
    if (x < 0) { x = 0; *bc = 1; }
    if (x >= nx) { x = (float)(nx - 1); *bc = 1; }
    return x;
}

static inline float jsq(float x) { return x * x; }

/* :50-54 zoom_size; the kernel passes its float factor through a double parameter */
void oracle_zoom_size(int nx, int ny, float factor, int *nxx, int *nyy)
{
    double f = (double)factor;
    *nxx = (int)((double)nx * f + 0.5);
    *nyy = (int)((double)ny * f + 0.5);
}

/* :488 factor = pow(scaleFactor, kiters-k-1), scaleFactor float (:1241) */
float oracle_level_factor(double scaleF, int kiters, int k)
{
    float sf = (float)scaleF;
    return (float)pow((double)sf, (double)(kiters - k - 1));
}

/* :521-526 in-kernel blur radius */
int oracle_filter_radius(float factor)
{
    float sigma = (float)(1.0 / sqrt(2. * (double)factor));
    int filtsize = (int)(2 * sigma);
    if (filtsize < 5) filtsize = 5;
    return filtsize;
}

/* :208-228 fill_GK -- 2R+1 taps, normalised over all of them */
void oracle_fill_gk(float *GK, float factor, int R)
{
    float sigma = (float)(0.6 * sqrt(1.0 / (double)(factor * factor) - 1.0));
    float s = (float)(2.0 * (double)sigma * (double)sigma);
    float sum = 0.0f;
    for (int x = -R; x <= R; x++) {
        float r = (float)x;
        GK[x + R] = (float)((double)expf(-(r * r) / s) / (3.14159265358979323846 * (double)s));
        sum += GK[x + R];
    }
    for (int i = 0; i < 2 * R + 1; ++i) GK[i] /= sum;
}

/* :312-351 convh then convv: taps kk in [-R, R) -- the +R tap is dropped */
static void blur_full(const float *img, float *tmp, float *out, const float *GK,
                      int nx, int ny, int R)
{
    int bc;
#pragma omp parallel for schedule(static) private(bc)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            float wsum = 0;
            for (int kk = -R; kk < R; ++kk) {
                int iiv = (int)bc_clamp((float)i + kk, nx, &bc);
                wsum = fmaf(GK[kk + R], img[(size_t)j * nx + iiv], wsum);
            }
            tmp[(size_t)j * nx + i] = wsum;
        }
#pragma omp parallel for schedule(static) private(bc)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            float wsum = 0;
            for (int kk = -R; kk < R; ++kk) {
                int jjv = (int)bc_clamp((float)j + kk, ny, &bc);
                wsum = fmaf(GK[kk + R], tmp[(size_t)jjv * nx + i], wsum);
            }
            out[(size_t)j * nx + i] = wsum;
        }
}

/* :231-239 oct_cell_cu (float taps, double literals) */
static inline float cell(const float v[4], float x)
{
    return (float)((double)v[1] + 0.5 * (double)x * ((double)(v[2] - v[0]) +
           (double)x * (2.0 * (double)v[0] - 5.0 * (double)v[1] + 4.0 * (double)v[2] - (double)v[3] +
           (double)x * (3.0 * (double)(v[1] - v[2]) + (double)v[3] - (double)v[0]))));
}

/* :258-309 oct_bicubic_cu: tap indices are (int)-truncated then clamped */
static float bicubic(const float *in, float uu, float vv, int nx, int ny)
{
    int bc;
    int x = (int)bc_clamp((float)((int)uu), nx, &bc);
    int y = (int)bc_clamp((float)((int)vv), ny, &bc);
    int mx = (int)bc_clamp((float)((int)(uu - 1)), nx, &bc);
    int my = (int)bc_clamp((float)((int)(vv - 1)), ny, &bc);
    int dx = (int)bc_clamp((float)((int)(uu + 1)), nx, &bc);
    int dy = (int)bc_clamp((float)((int)(vv + 1)), ny, &bc);
    int ddx = (int)bc_clamp((float)((int)(uu + 2)), nx, &bc);
    int ddy = (int)bc_clamp((float)((int)(vv + 2)), ny, &bc);
    const int xs[4] = { mx, x, dx, ddx };
    const size_t ys[4] = { (size_t)nx * my, (size_t)nx * y, (size_t)nx * dy, (size_t)nx * ddy };
    float v[4];
    for (int a = 0; a < 4; a++) {           /* column a, interpolated along y first (:248-254) */
        float p[4];
        for (int b = 0; b < 4; b++) p[b] = in[xs[a] + ys[b]];
        v[a] = cell(p, vv - y);
    }
    return cell(v, uu - x);
}

/* :354-408 zoom_out: bicubic at integer coordinates == decimation */
void oracle_blur_decimate(const float *img, int nx, int ny, int nc, float factor, float *out)
{
    int R = oracle_filter_radius(factor);
    float *GK = (float *)malloc(sizeof(float) * (2 * R + 1));
    oracle_fill_gk(GK, factor, R);
    size_t n = (size_t)nx * ny;
    float *tmp = (float *)malloc(n * sizeof(float));
    float *Is = (float *)malloc(n * sizeof(float));
    int nxx, nyy;
    oracle_zoom_size(nx, ny, factor, &nxx, &nyy);
    for (int c = 0; c < nc; c++) {
        /* Reference quirk: zoom_out samples `Is` without a channel offset (:406), so at
         * every coarse level ALL channels are the blurred, decimated CHANNEL 0. */
        if (c == 0) blur_full(img, tmp, Is, GK, nx, ny, R);
        float *o = out + (size_t)nxx * nyy * c;
#pragma omp parallel for schedule(static)
        for (int jj = 0; jj < nyy; jj++)
            for (int ii = 0; ii < nxx; ii++) {
                int i2 = (int)(ii / factor);
                int j2 = (int)(jj / factor);
                o[(size_t)jj * nxx + ii] = bicubic(Is, (float)i2, (float)j2, nx, ny);
            }
    }
    free(GK); free(tmp); free(Is);
}

/* :411-449 oct_compgrad_cu: 4th-order differences, clamp-to-edge, numerator in double */
void oracle_gradient(const float *f, float *gx, float *gy, int xi, int yi, int nc)
{
    size_t n = (size_t)xi * yi;
    for (int c = 0; c < nc; c++) {
        const float *g = f + n * c;
        float *ox = gx + n * c, *oy = gy + n * c;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < yi; j++) {
            int bc;
            int jp1 = (int)bc_clamp((float)j + 1, yi, &bc), jp2 = (int)bc_clamp((float)j + 2, yi, &bc);
            int jm1 = (int)bc_clamp((float)j - 1, yi, &bc), jm2 = (int)bc_clamp((float)j - 2, yi, &bc);
            for (int i = 0; i < xi; i++) {
                int ip1 = (int)bc_clamp((float)i + 1, xi, &bc), ip2 = (int)bc_clamp((float)i + 2, xi, &bc);
                int im1 = (int)bc_clamp((float)i - 1, xi, &bc), im2 = (int)bc_clamp((float)i - 2, xi, &bc);
                size_t row = (size_t)j * xi;
                ox[row + i] = (float)((-(double)g[row + ip2] + 8. * g[row + ip1] - 8. * g[row + im1] + g[row + im2]) / 12.0);
                oy[row + i] = (float)((-(double)g[(size_t)jp2 * xi + i] + 8. * g[(size_t)jp1 * xi + i]
                                       - 8. * g[(size_t)jm1 * xi + i] + g[(size_t)jm2 * xi + i]) / 12.0);
            }
        }
    }
}

/* :453-466 zoom_in (flow prolongation, result divided by sf) */
void oracle_zoom_in(const float *flow, float *out, int nx, int ny, int nxx, int nyy, float sf)
{
    const float factorx = ((float)nxx / nx);
    const float factory = ((float)nyy / ny);
#pragma omp parallel for schedule(static)
    for (int jj = 0; jj < nyy; jj++)
        for (int ii = 0; ii < nxx; ii++) {
            float i2 = (float)((double)(ii / factorx) - (0.5 - 0.5 / (double)factorx));
            float j2 = (float)((double)(jj / factory) - (0.5 - 0.5 / (double)factory));
            out[(size_t)jj * nxx + ii] = bicubic(flow, i2, j2, nx, ny) / sf;
        }
}

/* :72-108 robust-function derivatives */
static inline float psi_smooth(float x) { return (float)(1. / (double)sqrtf((float)((double)x + 1E-6))); }
static inline float psi_data(float x) { return (float)(1. / sqrt((double)x + 1E-6)); }

/* Coefficient build, :611-1097.  Outputs are the seven stencil coefficients
 * per pixel (diagonal block a1,a2,a4; off-diagonals with the reference's
 * boundary merging applied, :929-1077) and the right-hand side (bu,bv).
 * coef planes: [a1,a2,a4,a5,a6,a7,a8] each xi*yi. */
void oracle_build(const float *u, const float *v, const float *uh, const float *vh,
                  const float *g1, const float *g1x, const float *g1y,
                  const float *g2, const float *g2x, const float *g2y,
                  const float *g2xx, const float *g2xy, const float *g2yy,
                  int xi, int yi, int nchan, double alpha, double lambdadalpha, float lambdac,
                  int gnc, int dozim, float *coef, float *bu, float *bv)
{
    const size_t xityi = (size_t)xi * yi;
    const double al1 = 1. - 0.5 * gnc;
    float *A1 = coef, *A2 = coef + xityi, *A4 = coef + 2 * xityi, *A5 = coef + 3 * xityi,
          *A6 = coef + 4 * xityi, *A7 = coef + 5 * xityi, *A8 = coef + 6 * xityi;
#pragma omp parallel for schedule(static)
    for (int jj = 0; jj < yi; jj++)
        for (int ii = 0; ii < xi; ii++) {
            size_t l = (size_t)jj * xi + ii;
            /* mirror-without-repeat neighbours, :629-652 */
            int im = (ii == 0) ? ii + 1 : ii - 1;
            int ip = (ii == xi - 1) ? ii - 1 : ii + 1;
            int jm = (jj == 0) ? jj + 1 : jj - 1;
            int jp = (jj == yi - 1) ? jj - 1 : jj + 1;
#define AT(f, i, j) f[(size_t)(j) * xi + (i)]
            float up1p0 = AT(u, ip, jj), up0p0 = AT(u, ii, jj), up1p1 = AT(u, ip, jp), up1m1 = AT(u, ip, jm);
            float up0p1 = AT(u, ii, jp), up0m1 = AT(u, ii, jm), um1p1 = AT(u, im, jp), um1p0 = AT(u, im, jj);
            float um1m1 = AT(u, im, jm);
            float vp1p0 = AT(v, ip, jj), vp0p0 = AT(v, ii, jj), vp1p1 = AT(v, ip, jp), vp1m1 = AT(v, ip, jm);
            float vp0p1 = AT(v, ii, jp), vp0m1 = AT(v, ii, jm), vm1p1 = AT(v, im, jp), vm1p0 = AT(v, im, jj);
            float vm1m1 = AT(v, im, jm);
#undef AT
            /* :680-683 */
            float Uip1 = jsq(up1p0 - up0p0) + jsq(0.25 * ((up1p1 - up1m1) + (up0p1 - up0m1))) + jsq(vp1p0 - vp0p0) + jsq(0.25 * ((vp1p1 - vp1m1) + (vp0p1 - vp0m1)));
            float Uim1 = jsq(up0p0 - um1p0) + jsq(0.25 * ((um1p1 - um1m1) + (up0p1 - up0m1))) + jsq(vp0p0 - vm1p0) + jsq(0.25 * ((vm1p1 - vm1m1) + (vp0p1 - vp0m1)));
            float Ujp1 = jsq(up0p1 - up0p0) + jsq(0.25 * ((up1p1 - um1p1) + (up1p0 - um1p0))) + jsq(vp0p1 - vp0p0) + jsq(0.25 * ((vp1p1 - vm1p1) + (vp1p0 - vm1p0)));
            float Ujm1 = jsq(up0p0 - up0m1) + jsq(0.25 * ((up1m1 - um1m1) + (up1p0 - um1p0))) + jsq(vp0p0 - vp0m1) + jsq(0.25 * ((vp1m1 - vm1m1) + (vp1p0 - vm1p0)));
            /* :714-724 (dodiscrete is hard-wired false, :1302) */
            float psis1 = psi_smooth(Uim1), psis2 = psi_smooth(Ujm1), psis3 = psi_smooth(Uip1), psis4 = psi_smooth(Ujp1);
            float psistot = psis1 + psis2 + psis3 + psis4;
            float psistotq = 4.;
            float psisnmiu = psis1 * (um1p0) + psis2 * (up0m1) + psis3 * (up1p0) + psis4 * (up0p1);
            float psisnmiv = psis1 * (vm1p0) + psis2 * (vp0m1) + psis3 * (vp1p0) + psis4 * (vp0p1);
            float psisnmiuq = um1p0 + up0m1 + up1p0 + up0p1;
            float psisnmivq = vm1p0 + vp0m1 + vp1p0 + vp0p1;

            float vr1 = 0, vr2 = 0, vr4 = 0, vr5 = 0, vr6 = 0, intcomp = 0;
            float vr12 = 0, vr22 = 0, vr42 = 0, vr52 = 0, vr62 = 0, intcomp2 = 0;
            int bc, bc2 = 0, bc3 = 0;
            /* warp position with clamp, :732-745 */
            float iv = bc_clamp((float)(ii + up0p0), xi, &bc);
            if (bc) bc2 = 1;
            float jv = bc_clamp((float)(jj + vp0p0), yi, &bc);
            if (bc) bc3 = 1;
            int iv1 = (int)iv, jv1 = (int)jv;
            if (iv1 == xi - 1) iv1 = xi - 2;
            if (jv1 == yi - 1) jv1 = yi - 2;
            for (int c = 0; c < nchan; c++) {
                size_t off = xityi * c;
                size_t c1 = (size_t)iv1 + (size_t)xi * jv1 + off, c2 = c1 + 1, c3 = c1 + xi, c4 = c3 + 1;
                /* bilinear weights, :57-71 */
                float x1 = (float)iv1, x2 = (float)(iv1 + 1), y1 = (float)jv1, y2 = (float)(jv1 + 1);
                float p1 = (x2 - iv) / (x2 - x1), p2 = (iv - x1) / (x2 - x1);
                float p3 = ((y2 - jv) / (y2 - y1)), p4 = ((jv - y1) / (y2 - y1));
#define BIL(f) (p3 * ((p1) * f[c1] + (p2) * f[c2]) + p4 * ((p1) * f[c3] + (p2) * f[c4]))
                float g2w = BIL(g2), Ix = BIL(g2x), Iy = BIL(g2y), Ixx = BIL(g2xx), Ixy = BIL(g2xy), Iyy = BIL(g2yy);
#undef BIL
                if (bc2) { Ix = 0.; Ixx = 0.; Ixy = 0.; }
                if (bc3) { Iy = 0.; Ixy = 0.; Iyy = 0.; }
                /* :782-828 */
                float It = g2w - g1[l + off];
                float Ixt = Ix - g1x[l + off];
                float Iyt = Iy - g1y[l + off];
                float IxIx = Ix * Ix, IyIy = Iy * Iy, IxxIxx = Ixx * Ixx, IxyIxy = Ixy * Ixy, IyyIyy = Iyy * Iyy;
                float na, nb, nc;
                if (dozim) {
                    na = 1. / (IxIx + IyIy + 1.);
                    nb = 1. / (IxxIxx + IxyIxy + 1.);
                    nc = 1. / (IxyIxy + IyyIyy + 1.);
                } else { na = 1.; nb = 1.; nc = 1.; }
                intcomp += na * It * It;
                intcomp2 += (nb * Ixt * Ixt + nc * Iyt * Iyt);
                vr1 += (na * IxIx);
                vr12 += (nb * IxxIxx + nc * IxyIxy);
                vr2 += na * Ix * Iy;
                vr22 += (nb * Ixx * Ixy + nc * Iyy * Ixy);
                vr4 += (na * IyIy);
                vr42 += ((nb * IxyIxy + nc * IyyIyy));
                float natIt = -na * It, nbtIxt = nb * Ixt, nctIyt = nc * Iyt;
                vr5 += natIt * Ix;
                vr52 += -(nbtIxt * Ixx + nctIyt * Ixy);
                vr6 += natIt * Iy;
                vr62 += -(nbtIxt * Ixy + nctIyt * Iyy);
            }
            /* :831-864 */
            float psid = psi_data(intcomp) / alpha;
            float psid2 = lambdadalpha * psi_data(intcomp2);
            float a1 = (float)((al1) * ((vr1) / alpha + lambdadalpha * (vr12) + lambdac + psistotq) + (1 - al1) * (psid * (vr1) + psid2 * vr12 + lambdac + psistot));
            float a2 = (float)((al1) * ((vr2) / alpha + lambdadalpha * vr22) + (1 - al1) * (psid * (vr2) + psid2 * vr22));
            float a4 = (float)((al1) * ((vr4) / alpha + lambdadalpha * vr42 + lambdac + psistotq) + (1 - al1) * (psid * (vr4) + psid2 * vr42 + lambdac + psistot));
            float a5 = (float)(-1 * (al1 + (1 - al1) * (psis1)));
            float a6 = (float)(-1 * (al1 + (1 - al1) * (psis2)));
            float a7 = (float)(-1 * (al1 + (1 - al1) * (psis3)));
            float a8 = (float)(-1 * (al1 + (1 - al1) * (psis4)));
            /* boundary-merged CSR entries, :929-1002: a missing neighbour's weight is
             * added to the opposite neighbour's entry; the missing entry is absent (0). */
            A1[l] = a1; A2[l] = a2; A4[l] = a4;
            A6[l] = (jj > 0) ? ((jj < yi - 1) ? a6 : a6 + a8) : 0.f;
            A5[l] = (ii > 0) ? ((ii < xi - 1) ? a5 : a5 + a7) : 0.f;
            A7[l] = (ii < xi - 1) ? ((ii > 0) ? a7 : a7 + a5) : 0.f;
            A8[l] = (jj < yi - 1) ? ((jj > 0) ? a8 : a8 + a6) : 0.f;
            /* right-hand side, :1087-1092 */
            float uvt = uh ? uh[l] : 0.f, vvt = vh ? vh[l] : 0.f;
            float val2 = lambdac * (u[l] - uvt);
            bu[l] = (float)(al1 * ((vr5) / alpha + lambdadalpha * vr52 - val2 + psisnmiuq - psistotq * u[l]) +
                            (1. - al1) * (psid * (vr5) + psid2 * vr52 - val2 + psisnmiu - psistot * u[l]));
            val2 = lambdac * (v[l] - vvt);
            bv[l] = (float)(al1 * ((vr6) / alpha + lambdadalpha * vr62 - val2 + psisnmivq - psistotq * v[l]) +
                            (1 - al1) * (psid * (vr6) + psid2 * vr62 - val2 + psisnmiv - psistot * v[l]));
        }
}

/* y = A x with the row summation order of multiply_row (:112-121) over the
 * entry order the build writes: [j-1] [i-1] diag-block [i+1] [j+1]. */
void oracle_apply(const float *coef, const float *xu, const float *xv, int xi, int yi,
                  float *yu, float *yv)
{
    const size_t n = (size_t)xi * yi;
    const float *A1 = coef, *A2 = coef + n, *A4 = coef + 2 * n, *A5 = coef + 3 * n,
                *A6 = coef + 4 * n, *A7 = coef + 5 * n, *A8 = coef + 6 * n;
#pragma omp parallel for schedule(static)
    for (int jj = 0; jj < yi; jj++)
        for (int ii = 0; ii < xi; ii++) {
            size_t l = (size_t)jj * xi + ii;
            float su = 0, sv = 0;
            if (jj > 0) { su = fmaf(A6[l], xu[l - xi], su); }
            if (ii > 0) { su = fmaf(A5[l], xu[l - 1], su); }
            su = fmaf(A1[l], xu[l], su);
            su = fmaf(A2[l], xv[l], su);
            if (ii < xi - 1) { su = fmaf(A7[l], xu[l + 1], su); }
            if (jj < yi - 1) { su = fmaf(A8[l], xu[l + xi], su); }
            if (jj > 0) { sv = fmaf(A6[l], xv[l - xi], sv); }
            if (ii > 0) { sv = fmaf(A5[l], xv[l - 1], sv); }
            sv = fmaf(A2[l], xu[l], sv);
            sv = fmaf(A4[l], xv[l], sv);
            if (ii < xi - 1) { sv = fmaf(A7[l], xv[l + 1], sv); }
            if (jj < yi - 1) { sv = fmaf(A8[l], xv[l + xi], sv); }
            yu[l] = su; yv[l] = sv;
        }
}

/* fixed-order double dot over both components (replaces jVecXVec :151-186) */
static float dot2(const float *au, const float *av, const float *bu, const float *bv,
                  int xi, int yi, double *rowsum)
{
#pragma omp parallel for schedule(static)
    for (int j = 0; j < yi; j++) {
        double s = 0;
        const size_t o = (size_t)j * xi;
        for (int i = 0; i < xi; i++)
            s += (double)(au[o + i] * bu[o + i]) + (double)(av[o + i] * bv[o + i]);
        rowsum[j] = s;
    }
    double t = 0;
    for (int j = 0; j < yi; j++) t += rowsum[j];
    return (float)t;
}

static double dot2d(const float *au, const float *av, const float *bu, const float *bv,
                    int xi, int yi, double *rowsum)
{
#pragma omp parallel for schedule(static)
    for (int j = 0; j < yi; j++) {
        double s = 0;
        const size_t o = (size_t)j * xi;
        for (int i = 0; i < xi; i++)
            s += (double)(au[o + i] * bu[o + i]) + (double)(av[o + i] * bv[o + i]);
        rowsum[j] = s;
    }
    double t = 0;
    for (int j = 0; j < yi; j++) t += rowsum[j];
    return t;
}

/* Jacobi-PCG, :1105-1182.  x starts at 0, so r0 = b (:1105-1113).
 * work: 8 planes of xi*yi floats.  Returns iterations executed. */
int oracle_pcg(const float *coef, float *bu, float *bv, float *xu, float *xv,
               int xi, int yi, int iters, float tol, float *work)
{
    const size_t n = (size_t)xi * yi;
    const float *A1 = coef, *A4 = coef + 2 * n;
    float *mu = work, *mv = work + n, *zu = work + 2 * n, *zv = work + 3 * n;
    float *pu = work + 4 * n, *pv = work + 5 * n, *ru = work + 6 * n, *rv = work + 7 * n;
    float *qu = (float *)malloc(2 * n * sizeof(float)), *qv = qu + n;
    double *rowsum = (double *)malloc(sizeof(double) * yi);
    for (size_t i = 0; i < n; i++) {
        xu[i] = 0; xv[i] = 0;
        mu[i] = (float)(1. / A1[i]);            /* jDiagInv :142-149 */
        mv[i] = (float)(1. / A4[i]);
        zu[i] = mu[i] * bu[i]; zv[i] = mv[i] * bv[i];   /* :1117 */
        pu[i] = zu[i]; pv[i] = zv[i];                     /* :1119-1122 */
    }
    float residc = dot2(bu, bv, bu, bv, xi, yi, rowsum);  /* :1126 */
    int ki = 0;
    float z0tr0 = 0, zktrk, Bk, rkTzk, pkTApk, alphak;
    while ((residc > tol) && (ki < iters)) {
        if (ki > 0) {
            z0tr0 = dot2(zu, zv, bu, bv, xi, yi, rowsum);             /* :1135 */
            for (size_t i = 0; i < n; i++) { zu[i] = mu[i] * ru[i]; zv[i] = mv[i] * rv[i]; }  /* :1138 */
            zktrk = dot2(zu, zv, ru, rv, xi, yi, rowsum);             /* :1142 */
            Bk = zktrk / z0tr0;                                       /* :1144 */
            for (size_t i = 0; i < n; i++) {                          /* :1146 */
                pu[i] = fmaf(Bk, pu[i], zu[i]);
                pv[i] = fmaf(Bk, pv[i], zv[i]);
            }
            memcpy(bu, ru, n * sizeof(float)); memcpy(bv, rv, n * sizeof(float));  /* :1150-1152 */
        }
        rkTzk = dot2(bu, bv, zu, zv, xi, yi, rowsum);                 /* :1157 */
        oracle_apply(coef, pu, pv, xi, yi, qu, qv);                   /* :1161 (and again :1170) */
        pkTApk = dot2(pu, pv, qu, qv, xi, yi, rowsum);                /* :1165 */
        alphak = rkTzk / pkTApk;                                      /* :1169 */
        float nalpha = (float)(-1. * (double)alphak);
        for (size_t i = 0; i < n; i++) {
            xu[i] = fmaf(alphak, pu[i], xu[i]);                       /* :1172 */
            xv[i] = fmaf(alphak, pv[i], xv[i]);
            ru[i] = fmaf(nalpha, qu[i], bu[i]);                       /* :1174 */
            rv[i] = fmaf(nalpha, qv[i], bv[i]);
        }
        residc = dot2(ru, rv, ru, rv, xi, yi, rowsum);                /* :1178 */
        ki++;
    }
    free(qu); free(rowsum);
    return ki;
}

/* ---- model of the product's single-reduction PCG (octane_b200/csrc/pcg_fused.cu) ------------------
 * NOT a restatement of the reference: the same vector recurrences as oracle_pcg (p = z + beta p, q = A p,
 * x += alpha p, r -= alpha q) with the scalars of the NEXT iteration formed in one dot-product phase from
 * r.z, r.r, z.w, z.q, p.w, p.q (w = A z of the new residual), scalars in double, vectors in float.  It exists so that the distance between
 * the two recurrences can be measured on the CPU (tests/test_oracle_golden.py) and so that the CUDA kernel has a
 * model with the same operation order.  Selected with oracle_set_solver(1); the default (0) is the reference's
 * recurrence, which every parity gate is judged against. */
static int g_solver = 0;
void oracle_set_solver(int s) { g_solver = s; }

int oracle_pcg_merged(const float *coef, float *bu, float *bv, float *xu, float *xv,
                      int xi, int yi, int iters, float tol, float *work)
{
    const size_t n = (size_t)xi * yi;
    const float *A1 = coef, *A4 = coef + 2 * n;
    float *mu = work, *mv = work + n, *zu = work + 2 * n, *zv = work + 3 * n;
    float *pu = work + 4 * n, *pv = work + 5 * n, *wu = work + 6 * n, *wv = work + 7 * n;
    float *qu = (float *)calloc(2 * n, sizeof(float)), *qv = qu + n;
    double *rowsum = (double *)malloc(sizeof(double) * yi);
    float *ru = bu, *rv = bv;                               /* r0 = b (x0 = 0), updated in place */
    for (size_t i = 0; i < n; i++) {
        xu[i] = 0; xv[i] = 0; pu[i] = 0; pv[i] = 0;
        mu[i] = (float)(1. / A1[i]);
        mv[i] = (float)(1. / A4[i]);
        zu[i] = mu[i] * ru[i]; zv[i] = mv[i] * rv[i];
    }
    float rr = dot2(ru, rv, ru, rv, xi, yi, rowsum);
    int ki = 0;
    if (!(rr > tol) || iters <= 0) { free(qu); free(rowsum); return 0; }
    oracle_apply(coef, zu, zv, xi, yi, wu, wv);
    double gamma = dot2d(ru, rv, zu, zv, xi, yi, rowsum);
    double pAp = dot2d(wu, wv, zu, zv, xi, yi, rowsum);     /* p0 = z0: p0.Ap0 = z0.w0 */
    double alpha = gamma / pAp, beta = 0.0;
    while ((rr > tol) && (ki < iters)) {
        const float af = (float)alpha, bf = (float)beta, naf = -af;
        for (size_t i = 0; i < n; i++) {
            pu[i] = fmaf(bf, pu[i], zu[i]);  pv[i] = fmaf(bf, pv[i], zv[i]);      /* p = z + beta p, :1146 */
            xu[i] = fmaf(af, pu[i], xu[i]);  xv[i] = fmaf(af, pv[i], xv[i]);      /* :1172 */
        }
        oracle_apply(coef, pu, pv, xi, yi, qu, qv);                                /* q = A p, :1161 */
        for (size_t i = 0; i < n; i++) {
            ru[i] = fmaf(naf, qu[i], ru[i]); rv[i] = fmaf(naf, qv[i], rv[i]);     /* :1174 */
            zu[i] = mu[i] * ru[i];           zv[i] = mv[i] * rv[i];
        }
        oracle_apply(coef, zu, zv, xi, yi, wu, wv);
        /* one reduction phase: everything the next iteration's scalars need.  The boundary-merged matrix is NOT
         * symmetric (a7(0,j) = 2 W, a5(1,j) = W), so p.Ap of the next direction p' = z + beta' p is expanded without
         * assuming symmetry: p'.Ap' = z.w + beta' (z.q + p.w) + beta'^2 p.q */
        const double gnew = dot2d(ru, rv, zu, zv, xi, yi, rowsum);
        const double zw = dot2d(zu, zv, wu, wv, xi, yi, rowsum), zq = dot2d(zu, zv, qu, qv, xi, yi, rowsum);
        const double pw = dot2d(pu, pv, wu, wv, xi, yi, rowsum), pq = dot2d(pu, pv, qu, qv, xi, yi, rowsum);
        rr = dot2(ru, rv, ru, rv, xi, yi, rowsum);
        beta = gnew / gamma;
        pAp = zw + beta * (zq + pw) + beta * beta * pq;
        alpha = gnew / pAp;
        gamma = gnew;
        ki++;
    }
    free(qu); free(rowsum);
    return ki;
}

/* The whole solve, :487-1210 with the host wrapper's parameter prep, :1229-1241,:1353.
 * u,v: in = first guess (full res), out = flow.  cg_its (may be NULL) receives the
 * iteration count of each of the kiters*3*liters solves.  Returns 0. */
int oracle_variational_flow(const float *img1, const float *img2, int nx, int ny, int nc,
                            const oracle_params *p, float *u, float *v, int *cg_its)
{
    const size_t N = (size_t)nx * ny;
    const double alpha = p->alpha;
    const double lambdadalpha = p->lambda / alpha;
    const float lambdaco = (float)(p->lambdac / alpha);
    const float scaleFactor = (float)p->scaleF;
    const float tol = 0.0001 * 0.0001;
    const int kiters = p->kiters, liters = p->liters, iters = p->cgiters;

    float *geo1 = malloc(N * nc * 4), *geo2 = malloc(N * nc * 4);
    float *g1x = malloc(N * nc * 4), *g1y = malloc(N * nc * 4);
    float *g2x = malloc(N * nc * 4), *g2y = malloc(N * nc * 4);
    float *g2xx = malloc(N * nc * 4), *g2xy = malloc(N * nc * 4), *g2yy = malloc(N * nc * 4);
    float *uval = malloc(N * 4), *vval = malloc(N * 4), *uvalt = malloc(N * 4), *vvalt = malloc(N * 4);
    float *uhval = malloc(N * 4), *vhval = malloc(N * 4);
    float *coef = malloc(7 * N * 4), *bu = malloc(N * 4), *bv = malloc(N * 4);
    float *xu = malloc(N * 4), *xv = malloc(N * 4), *work = malloc(8 * N * 4);
    memcpy(uval, u, N * 4); memcpy(vval, v, N * 4);         /* :1330-1335 */
    memcpy(uhval, u, N * 4); memcpy(vhval, v, N * 4);
    int xi = 0, yi = 0, xio = 0, yio = 0, solve = 0;

    for (int k = 0; k < kiters; k++) {
        float factor = oracle_level_factor(p->scaleF, kiters, k);
        oracle_zoom_size(nx, ny, factor, &xi, &yi);
        const size_t n = (size_t)xi * yi;
        float lambdac = (float)((double)lambdaco * pow(0.5, k));     /* :494 */
        if (k > 0) {                                                 /* :498-503 */
            oracle_zoom_in(uvalt, uval, xio, yio, xi, yi, scaleFactor);
            oracle_zoom_in(vvalt, vval, xio, yio, xi, yi, scaleFactor);
        }
        if (k == kiters - 1) {                                       /* :504-517 */
            memcpy(geo1, img1, N * nc * 4); memcpy(geo2, img2, N * nc * 4);
            memcpy(uvalt, uhval, N * 4); memcpy(vvalt, vhval, N * 4);
        } else {                                                     /* :518-564 */
            oracle_blur_decimate(img1, nx, ny, nc, factor, geo1);
            oracle_blur_decimate(img2, nx, ny, nc, factor, geo2);
            oracle_blur_decimate(uhval, nx, ny, 1, factor, uvalt);
            oracle_blur_decimate(vhval, nx, ny, 1, factor, vvalt);
            for (size_t i = 0; i < n; i++) { uvalt[i] *= factor; vvalt[i] *= factor; }
        }
        if (k == 0) { memcpy(uval, uvalt, n * 4); memcpy(vval, vvalt, n * 4); }   /* :576-585 */
        oracle_gradient(geo1, g1x, g1y, xi, yi, nc);                 /* :587-595 */
        oracle_gradient(geo2, g2x, g2y, xi, yi, nc);
        oracle_gradient(g2x, g2xx, g2xy, xi, yi, nc);
        oracle_gradient(g2y, g2xy, g2yy, xi, yi, nc);                /* overwrites g2xy: d/dx(d/dy g2) */
        for (int gnc = 0; gnc < 3; gnc++)
            for (int l = 0; l < liters; l++) {
                oracle_build(uval, vval, uvalt, vvalt, geo1, g1x, g1y, geo2, g2x, g2y, g2xx, g2xy, g2yy,
                             xi, yi, nc, alpha, lambdadalpha, lambdac, gnc, p->dozim != 0, coef, bu, bv);
                int its = g_solver ? oracle_pcg_merged(coef, bu, bv, xu, xv, xi, yi, iters, tol, work)
                                   : oracle_pcg(coef, bu, bv, xu, xv, xi, yi, iters, tol, work);
                if (cg_its) cg_its[solve] = its;
                solve++;
                for (size_t i = 0; i < n; i++) { uval[i] = uval[i] + xu[i]; vval[i] = vval[i] + xv[i]; }  /* :1185-1195 */
            }
        memcpy(uvalt, uval, n * 4); memcpy(vvalt, vval, n * 4);      /* :1201-1205 */
        xio = xi; yio = yi;
    }
    memcpy(u, uval, N * 4); memcpy(v, vval, N * 4);                  /* :1434-1438 */
    free(geo1); free(geo2); free(g1x); free(g1y); free(g2x); free(g2y); free(g2xx); free(g2xy); free(g2yy);
    free(uval); free(vval); free(uvalt); free(vvalt); free(uhval); free(vhval);
    free(coef); free(bu); free(bv); free(xu); free(xv); free(work);
    return 0;
}

/* ---- navigation: src/oct_pix2uv_cuda.cu ---------------------------------- */

/* :13-25 -- lat/lon arrive narrowed to float */
static double haversine(float lat1, float lon1, float lat2, float lon2, double rad, double rad2)
{
    const double earthrad = 6371000.00;
    double dlon = lon2 - lon1;
    double dlat = lat2 - lat1;
    double a = (pow(sin(dlat * rad2), 2) + cos(lat1 * rad) * cos(lat2 * rad) * pow((sin(dlon * rad2)), 2));
    double c = 2. * atan2(sqrt(a), sqrt(1 - a));
    return earthrad * c;
}

/* :27-172 oct_navpixel_uv_cuda; xv = displacement per second */
static void navpixel_uv(const oracle_nav *geo, const double *xv, int xi, int yi, double dt, double *r,
                        double DTOR, double DTOR2, int dp, int dm)
{
    const double PI = 3.14159265359;
    double xVal, yVal, dist;
    double latv[2], lonv[2], sds[2] = { 0., 0. };
    for (int iv = 0; iv < 2; ++iv) {
        if (iv == 0) {            /* int*float+float: evaluated in float (contracted to an FMA on the GPU) */
            xVal = fmaf((float)xi, geo->xScale, geo->xOffset);
            yVal = fmaf((float)yi, geo->yScale, geo->yOffset);
        } else {
            xVal = (xv[0] * dt + xi) * geo->xScale + geo->xOffset;
            yVal = (xv[1] * dt + yi) * geo->yScale + geo->yOffset;
        }
        if (dp) {                 /* :34-67 polar orthographic */
            double rho = sqrt(xVal * xVal + yVal * yVal);
            double c = asin(rho / geo->R);
            if (geo->lat1 > 89.9999)
                lonv[iv] = geo->lon0 * DTOR + atan2(xVal, -yVal);
            else
                lonv[iv] = geo->lon0 * DTOR + atan2(xVal * sin(c), (rho * cos(geo->lat1 * DTOR) * cos(c) - yVal * sin(geo->lat1 * DTOR) * sin(c)));
            if (rho > 0.0000001)
                latv[iv] = asin(cos(c) * sin(geo->lat1 * DTOR) + (yVal * sin(c) * cos(geo->lat1 * DTOR) / rho));
            else
                latv[iv] = geo->lat1 * DTOR;
            latv[iv] = latv[iv] / DTOR;
            lonv[iv] = lonv[iv] / DTOR;
        } else if (dm) {          /* :70-87 Mercator */
            latv[iv] = PI / 2. - 2. * atan(exp(-yVal / geo->R));
            lonv[iv] = xVal / geo->R + geo->lon1;
            latv[iv] = latv[iv] / DTOR;
            lonv[iv] = lonv[iv] / DTOR;
        } else {                  /* :89-139 GOES fixed grid */
            double H = geo->pph + geo->req;
            sds[iv] = xVal * xVal + yVal * yVal;
            double a = pow((sin(xVal)), 2) + pow(cos(xVal), 2) * (pow((cos(yVal)), 2) + (pow(geo->req, 2)) / (pow(geo->rpol, 2)) * pow((sin(yVal)), 2));
            double b = -2. * H * cos(xVal) * cos(yVal);
            double c = pow(H, 2) - pow(geo->req, 2);
            double d = (pow(b, 2) - 4. * a * c);
            latv[iv] = -999.; lonv[iv] = -999.;
            if (d >= 0) {
                double rs = (-b - sqrt(d)) / (2. * a);
                double sx = rs * cos(xVal) * cos(yVal);
                double sy = -rs * sin(xVal);
                double sz = rs * cos(xVal) * sin(yVal);
                double e = (pow((H - sx), 2) + pow(sy, 2));
                if (!(sz == 0 || e <= 0 || H - sx == 0)) {
                    latv[iv] = atan((pow(geo->req, 2)) / (pow(geo->rpol, 2)) * (sz / sqrt(e)));
                    lonv[iv] = geo->lam0 - atan(sy / (H - sx));
                    latv[iv] = latv[iv] / DTOR;
                    lonv[iv] = lonv[iv] / DTOR;
                }
            }
        }
    }
    if ((latv[0] < -998) || (latv[1] < -998) || (sds[0] > 0.021)) {   /* :144-148 */
        r[0] = 0.; r[1] = 0.;
    } else {
        dist = haversine((float)latv[0], (float)lonv[0], (float)latv[0], (float)lonv[1], DTOR, DTOR2);
        r[0] = (lonv[1] >= lonv[0]) ? dist / dt : -dist / dt;
        dist = haversine((float)latv[0], (float)lonv[0], (float)latv[1], (float)lonv[0], DTOR, DTOR2);
        r[1] = (latv[1] >= latv[0]) ? dist / dt : -dist / dt;
    }
}

/* host wrapper :265-370 + kernel :174-221.  flags: bit0 pixuv, bit1 polar, bit2 mercator */
int oracle_pix2uv(const oracle_nav *nav, double t1, double t2, const float *u, const float *v,
                  int nx, int ny, int flags, short *ur, short *vr, short *ur2, short *vr2, float *dT)
{
    const size_t n = (size_t)nx * ny;
    const double pi = 3.14159265, rad = pi / 180., rad2 = rad / 2.;
    const int pixuv = flags & 1, dp = (flags >> 1) & 1, dm = (flags >> 2) & 1;
    float dxo = nav->xOffset - nav->g2xOffset, dyo = nav->yOffset - nav->g2yOffset;
    *dT = (float)(t2 - t1);
    if (!(((dxo * dxo) < (0.00001 * 0.00001)) && ((dyo * dyo) < (0.00001 * 0.00001)))) {   /* :295,358-368 */
        for (size_t k = 0; k < n; k++) { ur[k] = 0; vr[k] = 0; ur2[k] = 0; vr2[k] = 0; }
        return 1;
    }
    if (pixuv) {                                                      /* :348-356 */
        for (size_t k = 0; k < n; k++) {
            ur[k] = (short)(100 * u[k]); vr[k] = (short)(100 * v[k]);
            ur2[k] = ur[k]; vr2[k] = vr[k];   /* reference leaves ur2/vr2 unwritten here; see DESIGN.md */
        }
        return 0;
    }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            size_t k = (size_t)j * nx + i;
            double u1 = u[k], v1 = v[k];
            if (u1 > -9998.) {
                double dans[2] = { u1 / (t2 - t1), v1 / (t2 - t1) }, xans[2];
                navpixel_uv(nav, dans, i + nav->minX, j + nav->minY, t2 - t1, xans, rad, rad2, dp, dm);
                ur[k] = (short)(100 * (xans[0]));
                vr[k] = (short)(100 * (xans[1]));
            } else {
                ur[k] = (short)(-32768); vr[k] = (short)(-32768);
            }
            ur2[k] = (short)(100 * u[k]);                            /* :335-336 */
            vr2[k] = (short)(100 * v[k]);
        }
    return 0;
}

/* navigation in m/s without the short packing (for tolerance tests) */
int oracle_pix2uv_ms(const oracle_nav *nav, double t1, double t2, const float *u, const float *v,
                     int nx, int ny, int flags, double *ums, double *vms)
{
    const double pi = 3.14159265, rad = pi / 180., rad2 = rad / 2.;
    const int dp = (flags >> 1) & 1, dm = (flags >> 2) & 1;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            size_t k = (size_t)j * nx + i;
            double dans[2] = { (double)u[k] / (t2 - t1), (double)v[k] / (t2 - t1) }, xans[2];
            navpixel_uv(nav, dans, i + nav->minX, j + nav->minY, t2 - t1, xans, rad, rad2, dp, dm);
            ums[k] = xans[0]; vms[k] = xans[1];
        }
    return 0;
}

/* ---- checks of the exact shortcuts taken by the CUDA build kernel (octane_b200/csrc/build.cu) ----
 * The reference's promotion rules (:80,:102,:837-842) give float(1./double(s)) and double(x)/alpha
 * with float s, x; the kernel computes the same values with a single-precision reciprocal and with
 * multiply + one Markstein correction.  These enumerate / sample the operand space and count
 * mismatches against the plain expressions. */
long oracle_check_recip_float(unsigned stride, unsigned start)
{
    long bad = 0;
    if (stride == 0) stride = 1;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long k = start; k < 0x100000000LL; k += stride) {
        uint32_t b = (uint32_t)k;
        float s;
        memcpy(&s, &b, 4);
        if (!(s == s) || s == 0.f || isinf(s)) continue;
        volatile float via_double = (float)(1. / (double)s);   /* the reference's expression */
        volatile float direct = 1.0f / s;                      /* IEEE single division == __frcp_rn */
        if (via_double != direct) bad++;
    }
    return bad;
}

long oracle_check_div_const(double a, unsigned long long seed, long n)
{
    const double y = 1.0 / a;
    long bad = 0;
    unsigned long long s = seed ? seed : 88172645463325252ULL;
    for (long k = 0; k < n; k++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        uint32_t b = (uint32_t)(s >> 16);
        float f;
        memcpy(&f, &b, 4);
        if (!(f == f) || isinf(f)) continue;
        const double x = (double)f;
        const double q = x * y;
        const double r = fma(-q, a, x);
        const double q1 = fma(r, y, q);
        if (q1 != x / a) bad++;
    }
    return bad;
}

/* pyramid.cu: div12 -- the 4th-order difference's numerator (a double built from four floats, :411-449) divided by 12.0
 * through RN(1/12), one exact residual and one correction step must be the double the division gives.  mode 0: four
 * floats of arbitrary exponents (wide numerators); mode 1: image-like values of similar size. */
long oracle_check_div12(unsigned long long seed, long n, int mode)
{
    const double y = 1.0 / 12.0;
    long bad = 0;
    unsigned long long s = seed ? seed : 88172645463325252ULL;
    for (long k = 0; k < n; k++) {
        float v[4];
        for (int t = 0; t < 4; t++) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            uint32_t b = (uint32_t)(s >> 16);
            if (mode == 1) b = (b & 0x807fffffu) | ((uint32_t)(120 + (b >> 23) % 12) << 23);   /* 2^-7 .. 2^4 */
            memcpy(&v[t], &b, 4);
            if (!(v[t] == v[t]) || isinf(v[t])) v[t] = 1.0f;
        }
        const double x = (-v[0] + 8. * v[1] - 8. * v[2] + v[3]);
        if (isinf(x)) continue;
        const double q = x * y;
        const double r = fma(-q, 12.0, x);
        const double q1 = fma(r, y, q);
        if (q1 != x / 12.0 && !(fabs(x / 12.0) < 2.3e-308)) bad++;      /* subnormal quotients: floats cannot produce them */
    }
    return bad;
}

/* ---- ingest: octnavcalcuda, src/oct_navcal_cuda.cu:12-98 (host wrapper :100-207) ---------------
 * Calibration constants as oct_navcal_cuda receives them (all float).  cal: 0 RAW, 1 TEMP, 2 REF,
 * 3 BRIT.  lat/lon may be NULL. */
typedef struct {
    float xScale, xOffset, yScale, yOffset, radScale, radOffset;
    float rpol, req, H, lam0;
    float fk1, fk2, bc1, bc2, kap1;
    float maxin, minin, maxout, minout;
    int cal, donav;
} oracle_cal;

int oracle_navcal(const short *rad, const short *x, const short *y, int nx, int ny, const oracle_cal *c,
                  float *data3, float *lat, float *lon)
{
    const double PI = 3.14159265359, DTOR = PI / 180.;
    const float subpoint_slope = 1. / (0.021 - 0.0212);            /* :178-179 */
    const float subpoint_int = 1. - 0.021 * subpoint_slope;
    const float req = c->req, rpol = c->rpol, H = c->H;
    const float req2 = req * req, rpol2 = rpol * rpol;             /* pow(float,int) is float on the device */
    const float ratio = req2 / rpol2;
    const float cterm = H * H - req2;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            const size_t k = (size_t)j * nx + i;
            /* :31-34: short*float+float in float (an FMA on the GPU), then widened */
            double xVal = fmaf((float)x[i], c->xScale, c->xOffset);
            double yVal = fmaf((float)y[j], c->yScale, c->yOffset);
            double subpoint_dist = xVal * xVal + yVal * yVal;
            float dVal = fmaf((float)rad[k], c->radScale, c->radOffset);
            if (lat && lon) {
                if (c->donav == 1) {                               /* :36-49 */
                    double sxv = sin(xVal), cxv = cos(xVal), syv = sin(yVal), cyv = cos(yVal);
                    double a = sxv * sxv + cxv * cxv * (cyv * cyv + (double)ratio * (syv * syv));
                    double b = -2. * H * cxv * cyv;
                    double cc = cterm;
                    double rs = (-b - sqrt((b * b - 4. * a * cc))) / (2. * a);
                    double sx = rs * cxv * cyv;
                    double sy = -rs * sxv;
                    double sz = rs * cxv * syv;
                    float la = atan((double)ratio * (sz / sqrt(((H - sx) * (H - sx) + sy * sy))));
                    float lo = c->lam0 - atan(sy / (H - sx));
                    la = la / DTOR;
                    lo = lo / DTOR;
                    lat[k] = la; lon[k] = lo;
                } else {
                    lat[k] = 0.f; lon[k] = 0.f;
                }
            }
            double dataF;
            if (c->cal == 1) dataF = (c->fk2 / (log((c->fk1 / dVal) + 1.)) - c->bc1) / c->bc2;
            else if (c->cal == 2) dataF = c->kap1 * dVal;
            else dataF = dVal;
            float sdsconst;                                        /* :80-91 */
            if (subpoint_dist < 0.021) sdsconst = 1.f;
            else if (subpoint_dist >= 0.0212) sdsconst = 0.f;
            else sdsconst = subpoint_slope * subpoint_dist + subpoint_int;
            data3[k] = sdsconst * (((dataF - c->minin) / (c->maxin - c->minin)) * (c->maxout - c->minout) + c->minout);
        }
    return 0;
}

/* ---- first guess: octuv2xy + oct_uv2pix, src/oct_pix2uv_cuda.cu:223-263,372-476 ----------------
 * u, v in: wind (m/s); out: pixel displacement.  Returns 1 when the sector-moved guard zeroed it. */
int oracle_uv2pix(const oracle_nav *nav, double t1, double t2,
                  const float *lat, const float *lon, const short *xs, const short *ys, int nx, int ny,
                  float *u, float *v)
{
    const size_t n = (size_t)nx * ny;
    if (!((nav->xOffset == nav->g2xOffset) && (nav->yOffset == nav->g2yOffset))) {     /* :421, :464-474 */
        for (size_t k = 0; k < n; k++) { u[k] = 0.f; v[k] = 0.f; }
        return 1;
    }
    const double R = 6371000.0, pi = 3.14159265, rad = pi / 180.;
    const double req = nav->req, rpol = nav->rpol, req2 = req * req, rpol2 = rpol * rpol;
    double eval = sqrt((req2 - rpol2) / (req2));
    eval = eval * eval;
    const double H = nav->pph + req, secs = t2 - t1, lam0 = nav->lam0;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            const size_t k = (size_t)j * nx + i;
            double u1 = u[k], v1 = v[k], latvalv = lat[k], lonvalv = lon[k];
            double dist = sqrt(u1 * u1 + v1 * v1) * (secs);
            double brng = (180. + (90. - (atan2(-v1, -u1) / rad))) * rad;
            double latorig = latvalv * rad;
            latvalv = asin(sin(latorig) * cos(dist / R) + cos(latorig) * sin(dist / R) * cos(brng));
            lonvalv = lonvalv * rad + (atan2((sin(brng) * sin(dist / R) * cos(latorig)), (cos(dist / R) - sin(latorig) * sin(latvalv))));
            double thtc = atan(((rpol2) / (req2)) * tan(latvalv));
            double rc = rpol / sqrt(1. - (eval) * (cos(thtc) * cos(thtc)));
            double sx = H - rc * cos(thtc) * cos(lonvalv - lam0);
            double sy = -rc * cos(thtc) * sin(lonvalv - lam0);
            double sz = rc * sin(thtc);
            double x1v, y1v;
            if ((H * (H - sx)) >= (sy * sy + ((req2) / (rpol2) * sz * sz))) {
                x1v = (asin(-sy / (sqrt(sx * sx + sy * sy + sz * sz))) - nav->xOffset) / nav->xScale;
                y1v = (atan(sz / sx) - nav->yOffset) / nav->yScale;
            } else {
                x1v = -999.; y1v = -999.;
            }
            if (x1v > -998.) { u[k] = x1v - xs[i]; v[k] = y1v - ys[j]; }
            else { u[k] = 0.f; v[k] = 0.f; }
        }
    return 0;
}

/* ---- regridding: oct_zoom_in_float, src/oct_zoom.cc:180-222 + oct_bicubic_float / oct_cell,
 * src/oct_bicubic.cc:12-29,100-150 */
static double zcell(const double v[4], double x)
{
    return v[1] + 0.5 * x * (v[2] - v[0] + x * (2.0 * v[0] - 5.0 * v[1] + 4.0 * v[2] - v[3] + x * (3.0 * (v[1] - v[2]) + v[3] - v[0])));
}
static int zbc(int x, int n) { return x < 0 ? 0 : (x >= n ? n - 1 : x); }

int oracle_zoom_in_float(const float *in, int nx, int ny, float *out, int nxx, int nyy, int interp)
{
    const float factorx = ((float)nxx / nx), factory = ((float)nyy / ny);
    const float val1 = (0.5 - 0.5 / factory), val2 = (0.5 - 0.5 / factorx);
#pragma omp parallel for schedule(static)
    for (int jj1 = 0; jj1 < nyy; jj1++) {
        const float j2 = (float)((jj1 / factory) - val1);
        for (int i1 = 0; i1 < nxx; i1++) {
            const float i2 = (float)((i1 / factorx) - val2);
            float g;
            if (interp == 1) {
                const double uu = i2, vv = j2;
                const int x = zbc((int)uu, nx), y = zbc((int)vv, ny);
                const int cols[4] = { zbc((int)(uu - 1), nx), x, zbc((int)(uu + 1), nx), zbc((int)(uu + 2), nx) };
                const int rows[4] = { zbc((int)(vv - 1), ny), y, zbc((int)(vv + 1), ny), zbc((int)(vv + 2), ny) };
                double v[4];
                for (int c = 0; c < 4; c++) {
                    double p[4];
                    for (int r = 0; r < 4; r++) p[r] = in[cols[c] + (size_t)nx * rows[r]];
                    v[c] = zcell(p, vv - y);
                }
                g = zcell(v, uu - x);
            } else {
                const int j3 = (int)(j2 + 0.5), i3 = (int)(i2 + 0.5);
                g = in[i3 + (size_t)nx * j3];
            }
            out[i1 + (size_t)nxx * jj1] = g;
        }
    }
    return 0;
}

/* ---- ingest of the projected grids: octpolarnavcalcuda (src/oct_polar_navcal_cuda.cu:12-66, grid 1) and
 * octmercnavcalcuda (src/oct_merc_navcal_cuda.cu:12-48, grid 2).  lon0 / lat1 in radians as floats (what the
 * host wrappers pass, :141-143 / :122-124); sin / cos of the float lat1 resolve to the float overloads on the
 * device. */
int oracle_navcal_grid(int grid, const float *data2, const short *x, const short *y, int nx, int ny, float xScale,
                       float xOffset, float yScale, float yOffset, float R, float lon0, float lat1, int donav,
                       float *data3, float *lat, float *lon)
{
    const double PI = 3.14159265359, DTOR = PI / 180.;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            const size_t k = (size_t)j * nx + i;
            double xVal = fmaf((float)x[i], xScale, xOffset);
            double yVal = fmaf((float)y[j], yScale, yOffset);
            float la = 0.f, lo = 0.f;
            if (donav == 1) {
                if (grid == 1) {
                    double rho = sqrt(xVal * xVal + yVal * yVal);
                    double c = asin(rho / R);
                    if (lat1 > 89.99999) lo = lon0 + atan2(xVal, -yVal);
                    else lo = lon0 + atan2(xVal * sin(c), (rho * cosf(lat1) * cos(c) - yVal * sinf(lat1) * sin(c)));
                    if (rho > 0.0000001) la = asin(cos(c) * sinf(lat1) + (yVal * sin(c) * cosf(lat1) / rho));
                    else la = lat1;
                } else {
                    lo = xVal / R + lon0;
                    la = PI / 2. - 2. * atan(exp(-yVal / R));
                }
                la = la / DTOR;
                lo = lo / DTOR;
            }
            if (lat && lon) { lat[k] = la; lon[k] = lo; }
            data3[k] = data2[k];
        }
    return 0;
}

/* ---- down-scaling of a finer ancillary field onto the image grid: oct_zoom_out_float,
 * src/oct_zoom.cc:51-88, with oct_gaussian / oct_getGaussian_1D (src/oct_gaussian.cc:34-104) and
 * oct_bicubic (src/oct_bicubic.cc:36-97).  All arithmetic in double on a double copy of the field;
 * the blur drops the +R tap like the solver's (kk < filtsize, :70,91) and its radius is
 * (int)(2 sigma), at least 5 (:54-56).  out is the dense nxx x nyy plane (the reference writes it at
 * imageout + cnum, :84 -- an element offset, not a plane offset; the caller places the plane). */
void oracle_zoom_out_size(int nx, int ny, double factor, int *nxx, int *nyy)
{
    *nxx = (int)((double)nx * factor + 0.5);       /* src/oct_zoom.cc:12-16 */
    *nyy = (int)((double)ny * factor + 0.5);
}

int oracle_gaussian_taps(double sigma, double *GK, int max_taps)
{
    int filtsize = 2 * sigma;                      /* (int) 2*sigma: the product is truncated on assignment */
    if (filtsize < 5) filtsize = 5;
    const int wk = 2 * filtsize + 1;
    if (wk > max_taps) return -1;
    const double s = 2.0 * sigma * sigma;
    double sum = 0.0;
    for (int x = -filtsize; x <= filtsize; x++) {
        const double r = x;
        GK[x + filtsize] = (exp(-(r * r) / s)) / (M_PI * s);
        sum += GK[x + filtsize];
    }
    for (int i = 0; i < wk; ++i) GK[i] /= sum;
    return filtsize;
}

static double zbicubic_d(const double *in, double uu, double vv, int nx, int ny)
{
    const int x = zbc((int)uu, nx), y = zbc((int)vv, ny);
    const int cols[4] = { zbc((int)(uu - 1), nx), x, zbc((int)(uu + 1), nx), zbc((int)(uu + 2), nx) };
    const int rows[4] = { zbc((int)(vv - 1), ny), y, zbc((int)(vv + 1), ny), zbc((int)(vv + 2), ny) };
    double v[4];
    /* pol[c][r] = input[cols[c] + nx rows[r]]; oct_bicubic_cell(pol, uu - x, vv - y) runs oct_cell over r
     * with its second argument (vv - y), then over c with (uu - x) */
    for (int c = 0; c < 4; c++) {
        double p[4];
        for (int r = 0; r < 4; r++) p[r] = in[cols[c] + (size_t)nx * rows[r]];
        v[c] = zcell(p, vv - y);
    }
    return zcell(v, uu - x);
}

int oracle_zoom_out_float(const float *in, int nx, int ny, float *out, double factor)
{
    int nxx, nyy;
    oracle_zoom_out_size(nx, ny, factor, &nxx, &nyy);
    if (!(factor < 0.999999)) {                    /* :74-81: plain copy, read with the OUTPUT stride */
        for (int jj = 0; jj < nyy; jj++)
            for (int ii = 0; ii < nxx; ii++) out[ii + (size_t)nxx * jj] = in[ii + (size_t)nxx * jj];
        return 0;
    }
    const size_t n = (size_t)nx * ny;
    double *Is = (double *)malloc(n * sizeof(double)), *tmp = (double *)malloc(n * sizeof(double));
    double GK[257];
    const double sigma = 0.6 * sqrt(1.0 / (factor * factor) - 1.0);
    const int R = oracle_gaussian_taps(sigma, GK, 257);
    if (!Is || !tmp || R < 0) { free(Is); free(tmp); return -1; }
    for (size_t i = 0; i < n; i++) Is[i] = in[i];
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            double wsum = 0;
            for (int k = -R; k < R; ++k) wsum = wsum + GK[k + R] * Is[zbc(i + k, nx) + (size_t)nx * j];
            tmp[i + (size_t)nx * j] = wsum;
        }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            double wsum = 0;
            for (int l = -R; l < R; ++l) wsum = wsum + GK[l + R] * tmp[i + (size_t)nx * zbc(j + l, ny)];
            Is[i + (size_t)nx * j] = wsum;
        }
#pragma omp parallel for schedule(static)
    for (int jj = 0; jj < nyy; jj++)
        for (int ii = 0; ii < nxx; ii++) {
            const double i2 = (double)ii / factor, j2 = (double)jj / factor;
            out[ii + (size_t)nxx * jj] = zbicubic_d(Is, i2, j2, nx, ny);
        }
    free(Is); free(tmp);
    return 0;
}

/* ---- the optional post-smoother of the pixel displacements (-srsal): octsrsalcuda + oct_srsal_cu,
 * src/oct_srsal_cuda.cu:16-71,73-147.  37 x 37 bilateral filter: spatial Gaussian (sigma 9 px, radius 18,
 * normalised 1-D taps) times a range weight exp(-(dCTH)^2 / (2 * 20^2)) on the cloud-top heights; double
 * accumulators, kc (x offset) outer and lc (y offset) inner; the reflecting index rule of oct_bc_cuda
 * (:16-28: -x below, 2 nx - x - 1 above).  The device code contracts `au += u * a1` into an fma (nvcc
 * default), restated here with fma().  u, v are filtered in place (from a copy). */
int oracle_srsal(float *u, float *v, const float *cth, int nx, int ny)
{
    const double sigpix = 20., sigpix2 = -1. / (sigpix * sigpix * 2.);
    const double filtsigma = 9;
    const int filtsize = 2 * filtsigma;
    double GK[37];
    if (nx <= filtsize || ny <= filtsize) return -1;      /* the reflected index would leave the array */
    {
        const double s = 2.0 * filtsigma * filtsigma;
        double sum = 0.0;
        for (int x = -filtsize; x <= filtsize; x++) { const double r = x; GK[x + filtsize] = (exp(-(r * r) / s)) / (M_PI * s); sum += GK[x + filtsize]; }
        for (int i = 0; i < 2 * filtsize + 1; ++i) GK[i] /= sum;
    }
    const size_t n = (size_t)nx * ny;
    float *u0 = (float *)malloc(n * sizeof(float)), *v0 = (float *)malloc(n * sizeof(float));
    if (!u0 || !v0) { free(u0); free(v0); return -1; }
    memcpy(u0, u, n * sizeof(float)); memcpy(v0, v, n * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int jc = 0; jc < ny; jc++)
        for (int ic = 0; ic < nx; ic++) {
            const float pixc = cth[ic + (size_t)nx * jc];
            double au = 0, av = 0, a2 = 0;
            for (int kc = 0; kc < 2 * filtsize + 1; kc++)
                for (int lc = 0; lc < 2 * filtsize + 1; lc++) {
                    int ivc = ic + kc - filtsize, jvc = jc + lc - filtsize;
                    if (ivc < 0) ivc = 0 - ivc;
                    if (ivc >= nx) ivc = nx - (ivc - nx + 1);
                    if (jvc < 0) jvc = 0 - jvc;
                    if (jvc >= ny) jvc = ny - (jvc - ny + 1);
                    const size_t l2 = ivc + (size_t)jvc * nx;
                    const float pixl = cth[l2];
                    const double pixm = pixl - pixc;            /* float subtraction, then widened (:57) */
                    const double a1 = GK[kc] * GK[lc] * exp((pixm) * (pixm)*sigpix2);
                    a2 += a1;
                    au = fma((double)u0[l2], a1, au);
                    av = fma((double)v0[l2], a1, av);
                }
            u[ic + (size_t)nx * jc] = (au / a2);
            v[ic + (size_t)nx * jc] = (av / a2);
        }
    free(u0); free(v0);
    return 0;
}
