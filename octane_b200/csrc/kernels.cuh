// kernels.cuh -- host-callable launchers of the octane_b200 kernels.
#pragma once
#include "common.cuh"

namespace octane {

// ---- pyramid.cu
void launch_fill_gk(float* GK, float factor, int R, cudaStream_t st);
void launch_blur_decimate(const float* src, const Geom& gs, float* dst, const Geom& gd, int ja, int jb,
                          float factor, const float* GK, int R, float scale, int nc, cudaStream_t st);
void launch_gradient(const float* f, float* gx, float* gy, const Geom& g, int ja, int jb, int nc, cudaStream_t st);
void launch_zoom_in(const float* flow, const Geom& gc, float* out, const Geom& gf, int ja, int jb, float sf,
                    cudaStream_t st);

// ---- build.cu
struct LevelFields {           // device planes of one level, all sharing Geom g
    const float *g1, *g1x, *g1y;                     // image 1 and its gradient   (nc planes each)
    const float *g2, *g2x, *g2y, *g2xx, *g2xy, *g2yy; // image 2 and derivatives    (nc planes each)
    float *u, *v;                                    // current flow
    const float *uh, *vh;                            // hint (first guess at this level) or nullptr
};
struct PcgBuffers {
    float* coef[7];            // a1,a2,a4,a5,a6,a7,a8 (boundary-merged)
    float *ru, *rv;            // rhs, then residual
    float *xu, *xv;            // solution increment
    float *pu[2], *pv[2];      // search direction, ping-pong
    float *qu, *qv;            // A p
    PcgScalars* scal;          // device
    double* partials;          // device, >= 4 * max blocks
    unsigned* ticket;          // device
    int max_partial_blocks;
    double* pending;           // device [2]: rank-local dot totals awaiting the all-reduce
    int defer;                 // 1 (banded runs): kernels leave totals in `pending`;
                               // launch_finalize applies them after the all-reduce
};
enum { FINALIZE_BUILD = 0, FINALIZE_PASS1 = 1, FINALIZE_PASS2 = 2 };
void launch_finalize(const PcgBuffers& b, int kind, float tol, cudaStream_t st);
struct BuildParams {
    double alpha, lambdadalpha, al1;
    float lambdac;
    int dozim, nchan;
    float tol;
};
// rows [ja,jb) get coefficients + rhs; dots (b.b, b.Minv b) over rows [da,db)
void launch_build(const LevelFields& f, const PcgBuffers& b, const Geom& g, int ja, int jb, int da, int db,
                  const BuildParams& bp, int halo_check, cudaStream_t st);
int build_partial_blocks(const Geom& g, int nrows);

// ---- pcg.cu
// One PCG iteration = pass1 (p update fused with the stencil product, dot p.Ap)
// + pass2 (x, r update, dots r.r and z.r, stop rule).  `cur` selects which of
// the ping-pong p buffers holds p_old.
void launch_pcg_pass1(const PcgBuffers& b, const Geom& g, int ja, int jb, int first, int cur, int store_halo,
                      int sm_count, cudaStream_t st);
// large levels: persistent TMA-fed variant of pass 1 (pcg_tma.cu)
bool pcg_pass1_tma_usable(const Geom& g, int nrows);
void launch_pcg_pass1_tma(const PcgBuffers& b, const Geom& g, int ja, int jb, int first, int cur, int store_halo,
                          int sm_count, cudaStream_t st);
void launch_pcg_pass2(const PcgBuffers& b, const Geom& g, int ja, int jb, int first, int cur, int sm_count,
                      cudaStream_t st);
void launch_update_uv(float* u, float* v, const float* xu, const float* xv, const Geom& g, int ja, int jb,
                      const PcgScalars* s, int* its_out, int sm_count, cudaStream_t st);
// dense (stride nx, rows [ja,jb) starting at src row 0) <-> pitched
void launch_scale_copy(const float* src, float* dst, const Geom& g, int ja, int jb, float scale, cudaStream_t st);

// ---- nav.cu
struct NavParams {
    double pph, req, rpol, lam0;
    float xScale, xOffset, yScale, yOffset, lat1, lon1, lon0, R;
    int minX, minY;
    double t1, t2;
    int pixuv, dp, dm;
};
void launch_pix2uv(const NavParams& np, const float* u, const float* v, int nx, int row0, int nrows,
                   short* U, short* V, short* Uraw, short* Vraw, cudaStream_t st);
void launch_ctp_pack(const float* cth, short* ctp, size_t n, int ir, cudaStream_t st);

}  // namespace octane
