// oct_b200_shim.cc -- the reference-side binding of liboctane_b200.so.
//
// The reference has no FFI: its dispatcher (src/oct_optical_flow.cc:11-17)
// declares the two hot-path functions by hand and links their objects.  This
// file defines those two functions with the reference's EXACT C++ signatures
//
//   void oct_variational_optical_flow(Image, Image, float*, float*, float*, int, int, int, OFFlags)
//        replaces src/oct_variational_optical_flow.cu:1213
//   void oct_pix2uv_cuda(GOESVar&, double, float*, float*, short*, short*, short*, short*, OFFlags)
//        replaces src/oct_pix2uv_cuda.cu:265
//
//   void oct_navcal_cuda(short*, short*, short*, short*, short*, short*, int, int, int, int, int, int,
//                        float*, float*, float*, std::string, int, float x19, int, OFFlags)
//        replaces src/oct_navcal_cuda.cu:100 (ingest: calibration, normalisation, lat/lon)
//   void oct_uv2pix(GOESVar&, float*, float*, double, OFFlags)
//        replaces src/oct_pix2uv_cuda.cu:372 (first-guess winds -> pixel displacements)
//   void oct_srsal_cu(float*, float*, float*, int, int, OFFlags)
//        replaces src/oct_srsal_cuda.cu:73 (-srsal post-smoother)
//   void oct_zoom_in_float(float*, float*, int, int, int, int, int, int)
//   void oct_zoom_out_float(float*, float*, int, int, double, int, int)
//        replace src/oct_zoom.cc:180 and :51 (regridding of heights / extra channels; CPU code in the reference)
//
// on top of the C ABI in include/octane_b200.h, so that a maintainer drops the
// two .cu objects from the reference's link line, adds this file and
// -loctane_b200, and every caller above (oct_optical_flow.cc:67,91, main.cc:439)
// stays unmodified.  It is compiled against the reference's own headers
// (include/image.h, goesread.h, offlags.h) in the reference tree; nothing of
// the reference is copied here.  INTEGRATION.md has the build line.
//
// Error behaviour mirrors the reference wrapper: no GPU -> message + exit(0)
// (src/oct_variational_optical_flow.cu:1255-1259); any other failure, which the
// reference ignores (:1421,1431), prints octane_last_error() and exit(1).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "image.h"      // before goesread.h, which uses Image without including it
#include "goesread.h"
#include "offlags.h"
#include "octane_b200.h"

namespace {

octane_ctx* g_ctx[64] = { nullptr };

octane_ctx* context_for(int device)
{
    // this file was compiled against a header with OCTANE_ABI_VERSION; a library built from another one would
    // read / write past the structs passed below
    if (octane_abi_version() != OCTANE_ABI_VERSION) {
        fprintf(stderr, "octane_b200: liboctane_b200.so has ABI version %d, this shim was compiled for version %d\n",
                octane_abi_version(), OCTANE_ABI_VERSION);
        exit(1);
    }
    const int n = octane_device_count();
    if (n == 0) {
        std::cout << "No gpus available for use, exiting\n";
        exit(0);
    }
    if (device > n - 1 || device < 0) {
        std::cout << "Warning: setdevice set to non-existent GPU, setting to default GPU 1\n";
        device = 0;
    }
    if (device >= 64) device = 0;
    if (!g_ctx[device]) {
        int rc = octane_ctx_create(&g_ctx[device], device);
        if (rc) {
            fprintf(stderr, "octane_b200: %s\n", octane_last_error());
            exit(rc == OCTANE_ENODEV ? 0 : 1);
        }
    }
    return g_ctx[device];
}

void params_from(const OFFlags& a, octane_params* p)
{
    octane_params_default(p);
    p->alpha = a.alpha; p->lambda = a.lambda; p->lambdac = a.lambdac;      // :1229-1241
    p->scaleF = a.scaleF; p->scsig = a.scsig;
    p->kiters = a.kiters; p->liters = a.liters; p->cgiters = a.cgiters;
    p->dozim = a.dozim; p->setdevice = a.setdevice;
    p->pixuv = a.pixuv; p->dopolar = a.dopolar; p->domerc = a.domerc;     // src/oct_pix2uv_cuda.cu:291-297
    // The reference always reads uarr/varr as the first guess (:1330-1335); its only caller
    // (src/oct_optical_flow.cc:38-53) passes zeros unless -firstguess was given, and a zero
    // first guess is what first_guess = 0 means in the C ABI (no extra planes, same result).
    p->first_guess = (a.dofirstguess != 0);
    p->doCTH = a.doCTH; p->ir = a.ir;
}

void die(const char* what)
{
    fprintf(stderr, "octane_b200 %s: %s\n", what, octane_last_error());
    exit(1);
}

}  // namespace

void oct_variational_optical_flow(Image geo1i, Image geo2i, float* /*CTHarr: unused, dodiscrete=false :1302*/,
                                  float* uarr, float* varr, int nx, int ny, int nc, OFFlags args)
{
    octane_params p;
    params_from(args, &p);
    octane_ctx* c = context_for(args.setdevice);
    if (octane_variational_flow(c, geo1i.data, geo2i.data, nx, ny, nc, &p, uarr, varr) < 0)
        die("oct_variational_optical_flow");
}

void oct_pix2uv_cuda(GOESVar& goesData, double t2, float* uarr, float* varr, short* ur, short* vr, short* ur2,
                     short* vr2, OFFlags args)
{
    octane_params p;
    params_from(args, &p);
    octane_ctx* c = context_for(args.setdevice);
    const GOESNAVVar& g = goesData.nav;
    octane_nav nav;
    nav.pph = g.pph; nav.req = g.req; nav.rpol = g.rpol; nav.lam0 = g.lam0;
    nav.xScale = g.xScale; nav.xOffset = g.xOffset; nav.yScale = g.yScale; nav.yOffset = g.yOffset;
    nav.g2xOffset = g.g2xOffset; nav.g2yOffset = g.g2yOffset;
    nav.lat1 = g.lat1; nav.lon1 = g.lon1; nav.lon0 = g.lon0; nav.R = g.R;
    nav.minX = g.minX; nav.minY = g.minY;
    float dT = 0.f;
    int rc = octane_pix2uv(c, &nav, goesData.t, t2, uarr, varr, (int)g.nx, (int)g.ny, &p, ur, vr, ur2, vr2, &dT);
    if (rc < 0) die("oct_pix2uv_cuda");
    if (rc == 1)
        std::cout << "MOVE WARNING: Sector Moved, setting motions to 0 " << g.xOffset << " " << g.g2xOffset << " "
                  << g.yOffset << " " << g.g2yOffset << std::endl;
    goesData.dT = dT;          // src/oct_pix2uv_cuda.cu:347,357,369
}

namespace {
void nav_from(const GOESNAVVar& g, octane_nav* nav)
{
    nav->pph = g.pph; nav->req = g.req; nav->rpol = g.rpol; nav->lam0 = g.lam0;
    nav->xScale = g.xScale; nav->xOffset = g.xOffset; nav->yScale = g.yScale; nav->yOffset = g.yOffset;
    nav->g2xOffset = g.g2xOffset; nav->g2yOffset = g.g2yOffset;
    nav->lat1 = g.lat1; nav->lon1 = g.lon1; nav->lon0 = g.lon0; nav->R = g.R;
    nav->minX = g.minX; nav->minY = g.minY;
}
}  // namespace

// Ingest stage called by the GOES reader (src/oct_fileread.cc:359-388).  Like the reference
// wrapper it also fills the sector copies data2s / xs / ys (src/oct_navcal_cuda.cu:145-176).
void oct_navcal_cuda(short* data2, short* data2s, short* x, short* y, short* xs, short* ys, int nx, int ny, int minx,
                     int maxx, int miny, int maxy, float* data3, float* lat, float* lon, std::string cal, int /*datf*/,
                     float xScale, float xOffset, float yScale, float yOffset, float radScale, float radOffset,
                     float rpol, float req, float H, float lam0, float fk1, float fk2, float bc1, float bc2, float kap1,
                     float maxin, float minin, float maxout, float minout, int donav, OFFlags args)
{
    octane_ctx* c = context_for(args.setdevice);
    const int sx = maxx - minx, sy = maxy - miny;
    for (int l = miny; l < maxy && l < ny; l++) {
        for (int i = minx; i < maxx && i < nx; i++) data2s[(long)(i - minx) + (long)sx * (l - miny)] = data2[(long)i + (long)nx * l];
        ys[l - miny] = y[l];
    }
    for (int i = minx; i < maxx && i < nx; i++) xs[i - minx] = x[i];
    octane_nav nav = {};
    nav.req = req; nav.rpol = rpol; nav.pph = 0.; nav.lam0 = lam0;
    nav.xScale = xScale; nav.xOffset = xOffset; nav.yScale = yScale; nav.yOffset = yOffset;
    octane_cal k = {};
    k.radScale = radScale; k.radOffset = radOffset; k.fk1 = fk1; k.fk2 = fk2; k.bc1 = bc1; k.bc2 = bc2; k.kap1 = kap1;
    k.maxin = maxin; k.minin = minin; k.maxout = maxout; k.minout = minout; k.H = H;
    k.cal = (cal == "TEMP") ? 1 : (cal == "REF") ? 2 : (cal == "BRIT") ? 3 : 0;      // :115-118
    k.donav = donav;
    if (octane_navcal(c, data2s, xs, ys, sx, sy, &nav, &k, data3, lat, lon) < 0) die("oct_navcal_cuda");
}

void oct_uv2pix(GOESVar& goesData, float* u, float* v, double t2, OFFlags args)
{
    octane_params p;
    params_from(args, &p);
    octane_ctx* c = context_for(args.setdevice);
    octane_nav nav;
    nav_from(goesData.nav, &nav);
    if (octane_uv2pix(c, &nav, goesData.t, t2, goesData.latVal, goesData.lonVal, goesData.x, goesData.y,
                      (int)goesData.nav.nx, (int)goesData.nav.ny, &p, u, v) < 0)
        die("oct_uv2pix");
}

// Ingest of the projected grids (src/oct_polar_navcal_cuda.cu:69, src/oct_merc_navcal_cuda.cu:51): full sector
// only, as the readers call them (src/oct_fileread.cc:560-567, 728-731).  data2s is zero-filled, xs / ys copied,
// as the reference wrappers do.
static void grid_navcal(int grid, float* data2, short* data2s, short* x, short* y, short* xs, short* ys, int nx, int ny,
                        float* data3, float* lat, float* lon, float xScale, float xOffset, float yScale, float yOffset,
                        float lon0, float lat1, float R, int donav, long plane_offset, OFFlags args)
{
    octane_ctx* c = context_for(args.setdevice);
    for (long k = 0; k < (long)nx * ny; k++) data2s[k] = 0;
    for (int i = 0; i < nx; i++) xs[i] = x[i];
    for (int l = 0; l < ny; l++) ys[l] = y[l];
    octane_nav nav = {};
    nav.xScale = xScale; nav.xOffset = xOffset; nav.yScale = yScale; nav.yOffset = yOffset;
    nav.R = R; nav.lon0 = lon0; nav.lon1 = lon0; nav.lat1 = lat1;
    if (octane_navcal_grid(c, grid, data2, x, y, nx, ny, &nav, donav, data3 + plane_offset, lat, lon) < 0) die("oct_navcal (grid)");
}

void oct_polar_navcal_cuda(float* data2, short* data2s, short* x, short* y, short* xs, short* ys, int nx, int ny, int /*minx*/,
                           int /*maxx*/, int /*miny*/, int /*maxy*/, float* data3, float* lat, float* lon, float xScale,
                           float xOffset, float yScale, float yOffset, float lon0, float lat1, float R, int donav, int chan,
                           OFFlags args)
{
    grid_navcal(1, data2, data2s, x, y, xs, ys, nx, ny, data3, lat, lon, xScale, xOffset, yScale, yOffset, lon0, lat1, R, donav,
                (long)(chan - 1) * nx * ny, args);
}

void oct_merc_navcal_cuda(float* data2, short* data2s, short* x, short* y, short* xs, short* ys, int nx, int ny, int /*minx*/,
                          int /*maxx*/, int /*miny*/, int /*maxy*/, float* data3, float* lat, float* lon, float xScale,
                          float xOffset, float yScale, float yOffset, float lon0, float R, int donav, OFFlags args)
{
    grid_navcal(2, data2, data2s, x, y, xs, ys, nx, ny, data3, lat, lon, xScale, xOffset, yScale, yOffset, lon0, 0.f, R, donav, 0,
                args);
}

// -srsal post-smoother, called by the dispatcher after the navigation (src/oct_optical_flow.cc:100-105);
// replaces src/oct_srsal_cuda.cu:73.  In place on upix / vpix.
void oct_srsal_cu(float* upix, float* vpix, float* CTHsub21, int nx, int ny, OFFlags args)
{
    octane_ctx* c = context_for(args.setdevice);
    if (octane_srsal(c, upix, vpix, CTHsub21, nx, ny) < 0) die("oct_srsal_cu");
}

// Regridding of cloud-top heights / extra channels onto the image grid, CPU code in the reference
// (src/oct_zoom.cc:180 and :51, called by the readers: src/oct_fileread.cc:370,379,796,805).  A maintainer
// who wants them on the GPU removes (or weakens with objcopy, as the drop-in test build does) the two
// definitions in oct_zoom.cc.  Neither function gets OFFlags: device 0.
void oct_zoom_in_float(float* flow, float* flowout, int nx, int ny, int nxx, int nyy, int cnum, int interp)
{
    octane_ctx* c = context_for(0);
    const long cnumt = (long)cnum * ((long)nxx * nyy);                        // plane offset, :190
    if (octane_zoom_in_float(c, flow, nx, ny, flowout + cnumt, nxx, nyy, interp) < 0) die("oct_zoom_in_float");
}

void oct_zoom_out_float(float* image, float* imageout, int nx, int ny, double factor, int verb, int cnum)
{
    if (verb == 1) exit(0);                                                   // :61
    octane_ctx* c = context_for(0);
    // the reference stores pixel k at imageout[k + cnum] -- an ELEMENT offset (:84), kept as it is
    if (octane_zoom_out_float(c, image, nx, ny, imageout + cnum, factor) < 0) die("oct_zoom_out_float");
}
