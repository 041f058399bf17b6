// build.cu -- linearised data term + robust smoothness weights -> the 2x2-block
// 5-point system of one inner iteration, matrix-free.
// Replaces the body of the inner loop of octConjugateGradient,
// src/oct_variational_optical_flow.cu:611-1097 (reference tree): same
// neighbourhoods, same warp/clamp rules, same float/double promotion points.
// Instead of CSR arrays (12 nnz x 12 B per pixel) it stores 5 coefficients per
// pixel (the 2x2 diagonal block and the couplings to i+1 and j+1; the other two
// couplings are their neighbours' by symmetry), the right-hand side, and seeds
// the PCG scalars (r0 = b because x0 = 0).
#include "kernels.cuh"

namespace octane {

#define BUILD_ROWS 64

__device__ __forceinline__ float jsq(float x) { return x * x; }

// :72-108
__device__ __forceinline__ float psi_smooth(float x)
{
    float answer;
    answer = 1. / (sqrtf((x + 1E-6)));
    return answer;
}
__device__ __forceinline__ float psi_data(float x)
{
    float answer;
    answer = 1. / (sqrt(x + 1E-6));
    return answer;
}
// :57-71
__device__ __forceinline__ float binterp(float p1, float p2, float p3, float p4, float f11, float f21, float f12,
                                         float f22)
{
    return p3 * ((p1)*f11 + (p2)*f21) + p4 * ((p1)*f12 + (p2)*f22);
}

__global__ void __launch_bounds__(256)
k_build(LevelFields f, PcgBuffers b, Geom g, int ja, int jb, int da, int db, BuildParams bp, int halo_check)
{
    __shared__ double red[2 * 32];
    const int ii = blockIdx.x * 32 + threadIdx.x;
    const int xi = g.nx, yi = g.ny;
    double acc[2] = { 0.0, 0.0 };

    // a block covers 32 columns x BUILD_ROWS rows, 8 rows at a time
    for (int jt = 0; jt < BUILD_ROWS; jt += 8) {
    const int jj = ja + blockIdx.y * BUILD_ROWS + jt + threadIdx.y;
    if (ii < xi && jj < jb) {
        const double alpha = bp.alpha, lambdadalpha = bp.lambdadalpha, al1 = bp.al1;
        const float lambdac = bp.lambdac;
        // mirror-without-repeat neighbours (:629-652)
        const int im = (ii == 0) ? ii + 1 : ii - 1;
        const int ip = (ii == xi - 1) ? ii - 1 : ii + 1;
        const int jm = (jj == 0) ? jj + 1 : jj - 1;
        const int jp = (jj == yi - 1) ? jj - 1 : jj + 1;
        const float* u = f.u;
        const float* v = f.v;
        const size_t rm = g.at(0, jm), r0 = g.at(0, jj), rp = g.at(0, jp);
        float up1p0 = u[r0 + ip], up0p0 = u[r0 + ii], up1p1 = u[rp + ip], up1m1 = u[rm + ip];
        float up0p1 = u[rp + ii], up0m1 = u[rm + ii], um1p1 = u[rp + im], um1p0 = u[r0 + im], um1m1 = u[rm + im];
        float vp1p0 = v[r0 + ip], vp0p0 = v[r0 + ii], vp1p1 = v[rp + ip], vp1m1 = v[rm + ip];
        float vp0p1 = v[rp + ii], vp0m1 = v[rm + ii], vm1p1 = v[rp + im], vm1p0 = v[r0 + im], vm1m1 = v[rm + im];

        // :680-683
        float Uip1 = jsq(up1p0 - up0p0) + jsq(0.25 * ((up1p1 - up1m1) + (up0p1 - up0m1))) + jsq(vp1p0 - vp0p0) + jsq(0.25 * ((vp1p1 - vp1m1) + (vp0p1 - vp0m1)));
        float Uim1 = jsq(up0p0 - um1p0) + jsq(0.25 * ((um1p1 - um1m1) + (up0p1 - up0m1))) + jsq(vp0p0 - vm1p0) + jsq(0.25 * ((vm1p1 - vm1m1) + (vp0p1 - vp0m1)));
        float Ujp1 = jsq(up0p1 - up0p0) + jsq(0.25 * ((up1p1 - um1p1) + (up1p0 - um1p0))) + jsq(vp0p1 - vp0p0) + jsq(0.25 * ((vp1p1 - vm1p1) + (vp1p0 - vm1p0)));
        float Ujm1 = jsq(up0p0 - up0m1) + jsq(0.25 * ((up1m1 - um1m1) + (up1p0 - um1p0))) + jsq(vp0p0 - vp0m1) + jsq(0.25 * ((vp1m1 - vm1m1) + (vp1p0 - vm1p0)));
        // :714-724
        float psis1 = psi_smooth(Uim1);
        float psis2 = psi_smooth(Ujm1);
        float psis3 = psi_smooth(Uip1);
        float psis4 = psi_smooth(Ujp1);
        float psistot = psis1 + psis2 + psis3 + psis4;
        float psistotq = 4.;
        float psisnmiu = psis1 * (um1p0) + psis2 * (up0m1) + psis3 * (up1p0) + psis4 * (up0p1);
        float psisnmiv = psis1 * (vm1p0) + psis2 * (vp0m1) + psis3 * (vp1p0) + psis4 * (vp0p1);
        float psisnmiuq = um1p0 + up0m1 + up1p0 + up0p1;
        float psisnmivq = vm1p0 + vp0m1 + vp1p0 + vp0p1;

        float vr1 = 0, vr2 = 0, vr4 = 0, vr5 = 0, vr6 = 0, intcomp = 0;
        float vr12 = 0, vr22 = 0, vr42 = 0, vr52 = 0, vr62 = 0, intcomp2 = 0;
        // warped position with clamp, :732-745
        bool bc2 = false, bc3 = false;
        float iv = (float)(ii + up0p0);
        if (iv < 0) { iv = 0; bc2 = true; }
        if (iv >= xi) { iv = xi - 1; bc2 = true; }
        float jv = (float)(jj + vp0p0);
        if (jv < 0) { jv = 0; bc3 = true; }
        if (jv >= yi) { jv = yi - 1; bc3 = true; }
        int iv1 = int(iv);
        int jv1 = int(jv);
        if (iv1 == xi - 1) iv1 = xi - 2;
        if (jv1 == yi - 1) jv1 = yi - 2;
        if (halo_check) {     // banded run: rows jv1, jv1+1 must carry valid second derivatives
            const int vlo = (g.jlo() == 0) ? 0 : g.jlo() + 4, vhi = (g.jhi() == yi) ? yi : g.jhi() - 4;
            if (jv1 < vlo || jv1 + 1 >= vhi) {
                atomicOr(&b.scal->halo_err, 1);
                jv1 = min(max(jv1, g.jlo()), g.jhi() - 2);
            }
        }
        const size_t c1 = g.at(iv1, jv1), c2 = c1 + 1, c3 = c1 + g.pitch, c4 = c3 + 1;
        const size_t l = r0 + ii;
        // bilinear weights, :57-65 (x2-x1 == y2-y1 == 1)
        const float x1 = iv1, x2 = iv1 + 1, y1 = jv1, y2 = jv1 + 1;
        const float p1 = (x2 - iv) / (x2 - x1);
        const float p2 = (iv - x1) / (x2 - x1);
        const float p3 = ((y2 - jv) / (y2 - y1));
        const float p4 = ((jv - y1) / (y2 - y1));
        for (int c = 0; c < bp.nchan; c++) {
            const size_t off = (size_t)c * g.plane;
            const float* q;
            q = f.g2 + off;   float g2 = binterp(p1, p2, p3, p4, __ldg(q + c1), __ldg(q + c2), __ldg(q + c3), __ldg(q + c4));
            q = f.g2x + off;  float Ix = binterp(p1, p2, p3, p4, __ldg(q + c1), __ldg(q + c2), __ldg(q + c3), __ldg(q + c4));
            q = f.g2y + off;  float Iy = binterp(p1, p2, p3, p4, __ldg(q + c1), __ldg(q + c2), __ldg(q + c3), __ldg(q + c4));
            q = f.g2xx + off; float Ixx = binterp(p1, p2, p3, p4, __ldg(q + c1), __ldg(q + c2), __ldg(q + c3), __ldg(q + c4));
            q = f.g2xy + off; float Ixy = binterp(p1, p2, p3, p4, __ldg(q + c1), __ldg(q + c2), __ldg(q + c3), __ldg(q + c4));
            q = f.g2yy + off; float Iyy = binterp(p1, p2, p3, p4, __ldg(q + c1), __ldg(q + c2), __ldg(q + c3), __ldg(q + c4));
            // derivative zeroing where the warp was clamped, :768-779
            if (bc2) { Ix = 0.; Ixx = 0.; Ixy = 0.; }
            if (bc3) { Iy = 0.; Ixy = 0.; Iyy = 0.; }
            // :782-828
            float It = g2 - __ldg(f.g1 + off + l);
            float Ixt = Ix - __ldg(f.g1x + off + l);
            float Iyt = Iy - __ldg(f.g1y + off + l);
            float IxIx = Ix * Ix;
            float IyIy = Iy * Iy;
            float IxxIxx = Ixx * Ixx;
            float IxyIxy = Ixy * Ixy;
            float IyyIyy = Iyy * Iyy;
            float na, nb, nc;
            if (bp.dozim) {
                na = 1. / (IxIx + IyIy + 1.);
                nb = 1. / (IxxIxx + IxyIxy + 1.);
                nc = 1. / (IxyIxy + IyyIyy + 1.);
            } else {
                na = 1.; nb = 1.; nc = 1.;
            }
            intcomp += na * It * It;
            intcomp2 += (nb * Ixt * Ixt + nc * Iyt * Iyt);
            vr1 += (na * IxIx);
            vr12 += (nb * IxxIxx + nc * IxyIxy);
            vr2 += na * Ix * Iy;
            vr22 += (nb * Ixx * Ixy + nc * Iyy * Ixy);
            vr4 += (na * IyIy);
            vr42 += ((nb * IxyIxy + nc * IyyIyy));
            float natIt = -na * It;
            float nbtIxt = nb * Ixt;
            float nctIyt = nc * Iyt;
            vr5 += natIt * Ix;
            vr52 += -(nbtIxt * Ixx + nctIyt * Ixy);
            vr6 += natIt * Iy;
            vr62 += -(nbtIxt * Ixy + nctIyt * Iyy);
        }
        // :831-864
        float psid = psi_data(intcomp) / alpha;
        float psid2 = lambdadalpha * psi_data(intcomp2);
        float a1 = (float)((al1) * ((vr1) / alpha + lambdadalpha * (vr12) + lambdac + psistotq) + (1 - al1) * (psid * (vr1) + psid2 * vr12 + lambdac + psistot));
        float a2 = (float)((al1) * ((vr2) / alpha + lambdadalpha * vr22) + (1 - al1) * (psid * (vr2) + psid2 * vr22));
        float a4 = (float)((al1) * ((vr4) / alpha + lambdadalpha * vr42 + lambdac + psistotq) + (1 - al1) * (psid * (vr4) + psid2 * vr42 + lambdac + psistot));
        float a7 = (float)(-1 * (al1 + (1 - al1) * (psis3)));
        float a8 = (float)(-1 * (al1 + (1 - al1) * (psis4)));
        // Only the couplings to i+1 and j+1 are stored: a5(i,j) == a7(i-1,j) and
        // a6(i,j) == a8(i,j-1) bit for bit (same expression, two commuted additions), and at a
        // mirrored edge a5 == a7 (a6 == a8), so the boundary merging of :929-1077 -- the weight
        // of an absent neighbour is added to the opposite one -- is a doubling that the PCG
        // kernels apply on the fly (kernels.cuh, PcgBuffers).
        b.coef[C_A1][l] = a1;
        b.coef[C_A2][l] = a2;
        b.coef[C_A4][l] = a4;
        b.coef[C_W][l] = a7;
        b.coef[C_N][l] = a8;
        // right-hand side, :1087-1092
        float uvt = f.uh ? __ldg(f.uh + l) : 0.f;
        float vvt = f.vh ? __ldg(f.vh + l) : 0.f;
        float val2 = lambdac * (up0p0 - uvt);
        float bu = (float)(al1 * ((vr5) / alpha + lambdadalpha * vr52 - val2 + psisnmiuq - psistotq * up0p0) +
                           (1. - al1) * (psid * (vr5) + psid2 * vr52 - val2 + psisnmiu - psistot * up0p0));
        val2 = lambdac * (vp0p0 - vvt);
        float bv = (float)(al1 * ((vr6) / alpha + lambdadalpha * vr62 - val2 + psisnmivq - psistotq * vp0p0) +
                           (1 - al1) * (psid * (vr6) + psid2 * vr62 - val2 + psisnmiv - psistot * vp0p0));
        b.ru[l] = bu;
        b.rv[l] = bv;
        if (jj >= da && jj < db) {
            // residc = b.b (:1126); rkTzk = b.z with z = (1/M) b (:1115-1117,1157)
            const float mu = 1. / a1, mv = 1. / a4;
            const float zu = mu * bu, zv = mv * bv;
            acc[0] += (double)(bu * bu) + (double)(bv * bv);
            acc[1] += (double)(bu * zu) + (double)(bv * zv);
        }
    }
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const unsigned nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
        b.partials[bid] = acc[0];
        b.partials[(size_t)nblocks + bid] = acc[1];
    }
}

// second stage of the build's dot products (fixed order) + PCG scalar seeding:
// residc = b.b, rkTzk = b.z; stop rule already true -> no iterations (:1131).
__global__ void __launch_bounds__(1024) k_build_finish(PcgBuffers b, unsigned nblocks, float tol)
{
    __shared__ double red[2 * 32];
    double acc[2] = { 0.0, 0.0 };
    for (unsigned i = threadIdx.x; i < nblocks; i += 1024) {
        acc[0] += b.partials[i];
        acc[1] += b.partials[(size_t)nblocks + i];
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0 && b.defer) {
        b.pending[0] = acc[0];
        b.pending[1] = acc[1];
    } else if (threadIdx.x == 0) {
        PcgScalars* s = b.scal;
        s->rr = (float)acc[0];
        s->rz = (float)acc[1];
        s->rz_old = 0.f;
        s->pAp = 0.f;
        s->alpha = 0.f;
        s->tol = tol;
        s->its = 0;
        s->done = !((float)acc[0] > tol);
    }
}

void launch_build(const LevelFields& f, const PcgBuffers& b, const Geom& g, int ja, int jb, int da, int db,
                  const BuildParams& bp, int halo_check, cudaStream_t st)
{
    dim3 grid((g.nx + 31) / 32, (jb - ja + BUILD_ROWS - 1) / BUILD_ROWS), block(32, 8);
    k_build<<<grid, block, 0, st>>>(f, b, g, ja, jb, da, db, bp, halo_check);
    k_build_finish<<<1, 1024, 0, st>>>(b, grid.x * grid.y, bp.tol);
}

int build_partial_blocks(const Geom& g, int nrows)
{
    return ((g.nx + 31) / 32) * ((nrows + BUILD_ROWS - 1) / BUILD_ROWS);
}

}  // namespace octane
