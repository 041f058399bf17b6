// octane_main.cc -- the `octane` command line on top of liboctane_b200.so.
//
// Host-side counterpart of the reference's driver, readers and writer for the GOES path:
//   src/main.cc:26-484              flags (same names, same defaults, same quirks), sequencing
//   src/oct_fileread.cc:41-417      oct_goesread  (variables / attributes read, call of the ingest kernel)
//   src/oct_fileread.cc:754-859     oct_clavrxread, oct_fgread
//   src/oct_filewrite.cc:17-349     oct_goeswrite (output layout: variable order, types, attributes)
// The compute stages are the C-ABI entry points of include/octane_b200.h: octane_navcal (ingest),
// octane_uv2pix (first guess), octane_optical_flow (solver + CTP pack + navigation).
//
// Differences from the reference, all forced by this image having no netcdf-cxx4 / HDF5:
//   * files are classic NetCDF (CDF-1 / CDF-2), read and written by csrc/cdf.cc; the schema
//     (dimension / variable / attribute names, types, order) is the reference's;
//   * -sosm, -interp and extra channels / first guess on the -Polar / -Merc grids are not built (SURVEY.md
//     section 2: off the variational GOES path); the program says so and stops instead of silently doing
//     less.  Coarser fields are brought up with octane_zoom_in_float (oct_zoom_in_float), finer ones down
//     with octane_zoom_out_float (oct_zoom_out_float); -srsal runs inside the dispatcher (octane_srsal).
//   * an extra channel that is FINER than channel 1 lands in its own channel plane; the reference stores
//     it at an element offset instead (oct_zoom.cc:84) and so overwrites channel 1.
// Quirks kept: -cgiters is documented but never parsed (:144); -corn leaves docorn = 0 (:270-273);
// -scsig stores the square (:229); -set_device is 1-based (:313); -normmax/-normmin only change
// the attribute written, the normalisation always uses the band table (oct_fileread.cc:344-388).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <exception>
#include <new>
#include <string>
#include <vector>

#include "../../include/octane_b200.h"
#include "../csrc/cdf.h"

namespace {

struct Flags {                 // OFFlags, include/offlags.h:4-72 (fields this program uses)
    int farn = 0, pixuv = 0, dosrsal = 0, dopolar = 0, domerc = 0, doahi = 0, dofirstguess = 0, ir = 0;
    int dososm = 0, dointerp = 0, docorn = 0, rad = 2, srad = 2, interpcth = 1, doc2 = 0, doc3 = 0;
    int doinv = 0, doctt = 0, doCTH = 0, dozim = 1, oftype = 1, putinterp = 0, setdevice = 0;
    double lambda = 1., alpha = 5., filtsigma = 3., scaleF = 0.5, alpha2 = 20., lambdac = 0., scsig = 400.;
    int kiters = 4, liters = 3, cgiters = 30, miters = 5;
    float deltat = 60.f;
    float NormMax = 0.f, NormMin = 0.f, NormMax2 = 0.f, NormMin2 = 0.f, NormMax3 = 0.f, NormMin3 = 0.f;
    bool outnav = true, outraw = true, outrad = true, outctp = true;
    bool setNormMax = true, setNormMin = true, setNormMax2 = true, setNormMin2 = true, setNormMax3 = true,
         setNormMin3 = true;
    std::string ftype = "GOES";
    int dump_settings = 0;     // -dump_settings: print the parsed flags as key=value lines and exit (tests)
    int dry_run = 0;           // -dry_run: read the inputs and write outfile.nc with zero motion, no GPU work (tests of the I/O)
};

struct Scene {                 // the parts of GOESVar / GOESNAVVar the GOES path fills
    int nx = 0, ny = 0, band = 0;
    std::vector<short> rad, x, y;          // dataSVal, x, y
    std::vector<float> data, lat, lon;     // data.data (channel 1), latVal, lonVal
    double t = 0.;
    std::string tUnits;
    float xScale = 0, xOffset = 0, yScale = 0, yOffset = 0, radScale = 0, radOffset = 0;
    float req = 0, rpol = 0, pph = 0, lam0 = 0, lpo = 0, lat0 = 0, inverse = 0, gipVal = 0;
    float fk1 = 0, fk2 = 0, bc1 = 0, bc2 = 0, kap1 = 0;
    // -Polar / -Merc files: float image, grid constants of the "grid_mapping" variable
    std::vector<float> radf;
    float lat1 = 0, lon0 = 0, lon1 = 0, R = 0;
};

void usage()
{
    // src/main.cc:112-162
    printf("Optical Flow Toolkit for Atmospheric aNd Earth sciences (OCTANE) -- B200 build\n"
           "input flags:\n\n"
           "-i1 <filename>, -i2 <filename> are the GOES-R file netcdf full paths, i1 is the first image, i2 is the second\n\n"
           "-i1cth <filename>, -i2cth <filename> are optional paths to cloud top height netcdfs \n\n"
           "-nncth instead of default bilinear interpolation, remap the CTH grids with nearest neighbor \n\n"
           "-o <directory> writes the file to the designated directory, include slash at the end (default is ./) \n\n"
           "-pd forces OCTANE to return unnavigated pixel displacements \n\n"
           "-srsal -Polar -Merc -ahi -sosm -rad -srad -interp -deltat -interploc -ic21 -ic22 -ic31 -ic32: accepted, see the header of octane_main.cc\n\n"
           "-ir use this flag to output ir temperatures instead of cloud-top height (changes the scaling of the short variable ctp)\n\n"
           "-normmin(2|3) <value>, -normmax(2|3) <value> image brightness range recorded with the output\n\n"
           "-alpha <value> is a flag to set the smoothness constraint constant for Brox/Zimmer-based approaches \n\n"
           "-lambda <value> is a flag to set the gradient constraint constant for Brox/Zimmer-based approaches \n\n"
           "-lambdac <value> this is to set the weight of a hinting term, only used when -firstguess is active \n\n"
           "-kiters <int value> number of outer iterations/pyramid levels in Brox/Zimmer-based approaches, default is 4 \n\n"
           "-liters <int value> number of inner iterations in Brox/Zimmer-based approaches, default is 3 \n\n"
           "-brox set to perform pure Brox approach (default is modified zimmer) \n\n"
           "-firstguess <filename> set to input a first guess file (Only for GOES files, motions must be navigated) \n\n"
           "-no_outnav -no_outraw -no_outrad -no_outctp turn off groups of output variables\n\n"
           "-set_device <int> sets which gpu to run on, 1-based (default is 1), must be less than # of gpus \n\n");
}

// ---- readers ---------------------------------------------------------------------------------
bool att_f(const cdf::Var* v, const char* name, float* out, std::string* err)
{
    const cdf::Att* a = v ? v->att(name) : nullptr;
    if (!a || a->nelems() < 1) { *err = std::string("missing attribute ") + name; return false; }
    *out = (float)a->as_double();
    return true;
}

// oct_goesread, src/oct_fileread.cc:41-417 (channel 1)
bool read_goes(const std::string& path, Scene& s, std::string* err)
{
    cdf::Reader f;
    if (f.open(path)) { *err = f.error(); return false; }
    uint64_t nx = 0, ny = 0;
    if (f.dim_len("x", &nx) || f.dim_len("y", &ny)) { *err = path + ": no x / y dimension"; return false; }
    s.nx = (int)nx; s.ny = (int)ny;
    const cdf::Var *rad = f.var("Rad"), *xv = f.var("x"), *yv = f.var("y"), *tv = f.var("t"), *bv = f.var("band_id");
    const cdf::Var* gip = f.var("goes_imager_projection");
    if (!rad || !xv || !yv || !tv || !bv || !gip) { *err = path + ": Rad, x, y, t, band_id or goes_imager_projection missing"; return false; }
    if (rad->nelems != nx * ny || xv->nelems != nx || yv->nelems != ny) { *err = path + ": variable shapes do not match the dimensions"; return false; }
    if (!att_f(rad, "scale_factor", &s.radScale, err) || !att_f(rad, "add_offset", &s.radOffset, err) ||
        !att_f(yv, "scale_factor", &s.yScale, err) || !att_f(yv, "add_offset", &s.yOffset, err) ||
        !att_f(xv, "scale_factor", &s.xScale, err) || !att_f(xv, "add_offset", &s.xOffset, err) ||
        !att_f(gip, "longitude_of_projection_origin", &s.lpo, err) || !att_f(gip, "semi_major_axis", &s.req, err) ||
        !att_f(gip, "semi_minor_axis", &s.rpol, err) || !att_f(gip, "inverse_flattening", &s.inverse, err) ||
        !att_f(gip, "latitude_of_projection_origin", &s.lat0, err) || !att_f(gip, "perspective_point_height", &s.pph, err)) {
        *err = path + ": " + *err;
        return false;
    }
    const cdf::Att* tu = tv->att("units");
    if (!tu) { *err = path + ": t:units missing"; return false; }
    s.tUnits = tu->as_text();
    const double PI = 3.14159265359, DTOR = PI / 180.;
    s.lam0 = s.lpo * DTOR;                                   // :171-172 (float * double -> float)
    int gv = 0;
    f.get_int(gip, &gv);
    s.gipVal = (float)gv;
    const char* names[5] = { "planck_fk1", "planck_fk2", "planck_bc1", "planck_bc2", "kappa0" };
    float* dst[5] = { &s.fk1, &s.fk2, &s.bc1, &s.bc2, &s.kap1 };
    for (int k = 0; k < 5; k++) {
        const cdf::Var* v = f.var(names[k]);
        if (!v) { *err = path + ": " + names[k] + " missing"; return false; }
        if (f.get_float(v, dst[k])) { *err = f.error(); return false; }
    }
    s.rad.resize(nx * ny); s.x.resize(nx); s.y.resize(ny);
    if (f.get_short(rad, s.rad.data()) || f.get_short(xv, s.x.data()) || f.get_short(yv, s.y.data()) ||
        f.get_double(tv, &s.t) || f.get_int(bv, &s.band)) { *err = f.error(); return false; }
    return true;
}

// oct_polarread / oct_mercread, src/oct_fileread.cc:418-752: float Rad, short x / y with scale and offset,
// t, and the grid constants as attributes of the variable "grid_mapping" (polar: lat1, lon0, R; Mercator: lon1, R)
bool read_grid(const std::string& path, bool polar, Scene& s, std::string* err)
{
    cdf::Reader f;
    if (f.open(path)) { *err = f.error(); return false; }
    uint64_t nx = 0, ny = 0;
    if (f.dim_len("x", &nx) || f.dim_len("y", &ny)) { *err = path + ": no x / y dimension"; return false; }
    s.nx = (int)nx; s.ny = (int)ny;
    const cdf::Var *rad = f.var("Rad"), *xv = f.var("x"), *yv = f.var("y"), *tv = f.var("t"), *gm = f.var("grid_mapping");
    if (!rad || !xv || !yv || !tv || !gm) { *err = path + ": Rad, x, y, t or grid_mapping missing"; return false; }
    if (rad->nelems != nx * ny || xv->nelems != nx || yv->nelems != ny) { *err = path + ": variable shapes do not match the dimensions"; return false; }
    bool ok = att_f(yv, "scale_factor", &s.yScale, err) && att_f(yv, "add_offset", &s.yOffset, err) &&
              att_f(xv, "scale_factor", &s.xScale, err) && att_f(xv, "add_offset", &s.xOffset, err) && att_f(gm, "R", &s.R, err);
    if (ok && polar) ok = att_f(gm, "lat1", &s.lat1, err) && att_f(gm, "lon0", &s.lon0, err);
    if (ok && !polar) ok = att_f(gm, "lon1", &s.lon1, err);
    if (!ok) { *err = path + ": " + *err; return false; }
    const cdf::Att* tu = tv->att("units");
    if (!tu) { *err = path + ": t:units missing"; return false; }
    s.tUnits = tu->as_text();
    float gv = 0.f;
    f.get_float(gm, &gv);
    s.gipVal = gv;
    s.radf.resize(nx * ny); s.x.resize(nx); s.y.resize(ny);
    if (f.get_float(rad, s.radf.data()) || f.get_short(xv, s.x.data()) || f.get_short(yv, s.y.data()) || f.get_double(tv, &s.t)) {
        *err = f.error();
        return false;
    }
    return true;
}

// a float field with dimensions (ny, nx) as the CLAVR-x / first-guess readers expect (oct_fileread.cc:754-859)
bool read_plane(const std::string& path, const char* var, int* fx_out, int* fy_out, std::vector<float>& out, std::string* err)
{
    cdf::Reader f;
    if (f.open(path)) { *err = f.error(); return false; }
    uint64_t fx = 0, fy = 0;
    if (f.dim_len("nx", &fx) || f.dim_len("ny", &fy)) { *err = path + ": no nx / ny dimension"; return false; }
    const cdf::Var* v = f.var(var);
    if (!v || v->nelems != fx * fy) { *err = path + ": " + var + " missing or misshapen"; return false; }
    out.resize((size_t)fx * fy);
    if (f.get_float(v, out.data())) { *err = f.error(); return false; }
    *fx_out = (int)fx; *fy_out = (int)fy;
    return true;
}

// Bring a field of size fx*fy onto the nx*ny image grid the way the readers do (oct_fileread.cc:361-380,
// 794-806): same size -> unchanged (oct_zoom_out_float with factor 1 copies, oct_zoom.cc:79-86); image wider
// than the field -> oct_zoom_in_float (bicubic, or nearest neighbour with -nncth); otherwise
// oct_zoom_out_float with factor nx / fx, after the reference's check that x and y scale alike.
// ctx == nullptr (dry run): only the same-size case.
bool regrid(octane_ctx* ctx, const std::string& what, std::vector<float>& field, int fx, int fy, int nx, int ny, int interp,
            std::string* err)
{
    if (fx == nx && fy == ny) return true;
    const std::string sizes = what + " is " + std::to_string(fx) + "x" + std::to_string(fy) + ", image is " +
                              std::to_string(nx) + "x" + std::to_string(ny);
    if (!ctx) { *err = sizes + "; regridding needs the GPU (not available in a dry run)"; return false; }
    std::vector<float> out((size_t)nx * ny);
    if (nx > fx) {
        if (ny < fy) { *err = sizes + "; x needs up-scaling and y down-scaling"; return false; }
        if (octane_zoom_in_float(ctx, field.data(), fx, fy, out.data(), nx, ny, interp) < 0) { *err = octane_last_error(); return false; }
    } else {
        const double factor = (double)nx / ((double)fx), factor2 = (double)ny / ((double)fy);
        if (pow(factor - factor2, 2) > 0.000001) {      // oct_fileread.cc:375-378, 800-804 (the reference exits)
            *err = sizes + "; x and y dimensions not compatible for scaling (factor not the same)";
            return false;
        }
        int ox = 0, oy = 0;
        if (octane_zoom_out_size(fx, fy, factor, &ox, &oy) < 0 || ox != nx || oy != ny) {
            *err = sizes + "; scaled field would be " + std::to_string(ox) + "x" + std::to_string(oy);
            return false;
        }
        if (octane_zoom_out_float(ctx, field.data(), fx, fy, out.data(), factor) < 0) { *err = octane_last_error(); return false; }
    }
    field.swap(out);
    return true;
}

// ---- writer: oct_goeswrite, src/oct_filewrite.cc:17-349 ----------------------------------------
bool write_goes(const std::string& path, const Scene& s, const Flags& a, const octane_nav& nav, float dT,
                const short* U, const short* V, const short* Ur, const short* Vr, const float* upix, const float* vpix,
                const short* ctp, std::string* err, const Scene* const* extra = nullptr)
{
    using cdf::Att;
    const Scene* e2 = extra ? extra[0] : nullptr;      // channel 2 / 3 of image 1, when on the same grid
    const Scene* e3 = extra ? extra[1] : nullptr;
    cdf::Writer w;
    if (w.create(path)) { *err = w.error(); return false; }
    const int xd = w.add_dim("x", s.nx), yd = w.add_dim("y", s.ny);
    const int xv = w.add_var("x", cdf::SHORT, { xd }), yv = w.add_var("y", cdf::SHORT, { yd });
    w.put_att(xv, Att::f32("scale_factor", s.xScale)); w.put_att(xv, Att::f32("add_offset", s.xOffset));
    w.put_att(yv, Att::f32("scale_factor", s.yScale)); w.put_att(yv, Att::f32("add_offset", s.yOffset));
    const int tv = w.add_var("t", cdf::DOUBLE, {});
    w.put_att(tv, Att::text("standard_name", "time"));
    w.put_att(tv, Att::text("units", s.tUnits));
    w.put_att(tv, Att::text("axis", "T"));
    w.put_att(tv, Att::text("bounds", "time_bounds"));
    w.put_att(tv, Att::text("long_name", "J2000 epoch mid-point between the start and end image scan in seconds"));
    const std::vector<int> yx = { yd, xd };
    int uV = -1, vV = -1, urV = -1, vrV = -1, upV = -1, vpV = -1, ctpV = -1, radV = -1;
    if (a.outnav) { uV = w.add_var("U", cdf::SHORT, yx); vV = w.add_var("V", cdf::SHORT, yx); }
    if (a.outraw) { urV = w.add_var("U_raw", cdf::SHORT, yx); vrV = w.add_var("V_raw", cdf::SHORT, yx); }
    if (a.pixuv == 1) { upV = w.add_var("Upix", cdf::FLOAT, yx); vpV = w.add_var("Vpix", cdf::FLOAT, yx); }
    if (a.outctp && a.doCTH == 1) ctpV = w.add_var("CTP", cdf::SHORT, yx);
    int rad2V = -1, rad3V = -1;
    if (a.outrad) {
        radV = w.add_var("Rad", cdf::SHORT, yx);
        if (e2) rad2V = w.add_var("Rad2", cdf::SHORT, yx);
        if (e3) rad3V = w.add_var("Rad3", cdf::SHORT, yx);
    }
    const int gipV = w.add_var("goes_imager_projection", cdf::INT, {});
    const int ofV = w.add_var("optical_flow_settings", cdf::INT, {});
    int pk[3][5];
    for (int c = 0; c < 3; c++) for (int k = 0; k < 5; k++) pk[c][k] = -1;
    if (a.outrad) {
        const char* names[5] = { "planck_fk1", "planck_fk2", "planck_bc1", "planck_bc2", "kappa0" };
        const char* suffix[3] = { "", "_2", "_3" };
        const bool have[3] = { true, e2 != nullptr, e3 != nullptr };
        for (int c = 0; c < 3; c++)
            if (have[c]) for (int k = 0; k < 5; k++) pk[c][k] = w.add_var(std::string(names[k]) + suffix[c], cdf::FLOAT, {});
    }
    const char* gm = "goes_imager_projection";
    if (a.outnav) {
        w.put_att(uV, Att::text("long_name", "U")); w.put_att(uV, Att::text("grid_mapping", gm));
        w.put_att(uV, Att::f32("scale_factor", 0.01f));
        w.put_att(uV, Att::text("units", a.pixuv == 0 ? "meters per second" : "x-pixels"));
        w.put_att(vV, Att::text("long_name", "V")); w.put_att(vV, Att::text("grid_mapping", gm));
        w.put_att(vV, Att::f32("scale_factor", 0.01f));
        w.put_att(vV, Att::text("units", a.pixuv == 1 ? "y-pixels" : "meters per second"));
    }
    if (a.outraw) {
        w.put_att(urV, Att::text("long_name", "U Raw")); w.put_att(urV, Att::text("grid_mapping", gm));
        w.put_att(urV, Att::f32("scale_factor", 0.01f)); w.put_att(urV, Att::text("units", "x-pixels"));
        w.put_att(vrV, Att::text("long_name", "V Raw")); w.put_att(vrV, Att::text("grid_mapping", gm));
        w.put_att(vrV, Att::f32("scale_factor", 0.01f)); w.put_att(vrV, Att::text("units", "y-pixels"));
    }
    if (ctpV >= 0) {
        w.put_att(ctpV, Att::text("long_name", "CTP")); w.put_att(ctpV, Att::text("grid_mapping", gm));
        w.put_att(ctpV, Att::f32("interpcth", (float)a.interpcth));
    }
    if (a.outrad) {
        w.put_att(radV, Att::text("long_name", "Rad")); w.put_att(radV, Att::text("grid_mapping", gm));
        w.put_att(radV, Att::f32("scale_factor", s.radScale)); w.put_att(radV, Att::f32("add_offset", s.radOffset));
        const Scene* es[2] = { e2, e3 };
        const int ev[2] = { rad2V, rad3V };
        for (int c = 0; c < 2; c++) {
            if (!es[c]) continue;
            w.put_att(ev[c], Att::text("long_name", "Rad2"));          // sic: both extra channels, oct_filewrite.cc:195,202
            w.put_att(ev[c], Att::text("grid_mapping", gm));
            w.put_att(ev[c], Att::f32("scale_factor", es[c]->radScale)); w.put_att(ev[c], Att::f32("add_offset", es[c]->radOffset));
        }
    }
    w.put_att(gipV, Att::text("long_name", "GOES-R ABI fixed grid projection"));
    w.put_att(gipV, Att::text("grid_mapping_name", "geostationary"));
    w.put_att(gipV, Att::f64("perspective_point_height", nav.pph));
    w.put_att(gipV, Att::f64("semi_major_axis", nav.req));
    w.put_att(gipV, Att::f64("semi_minor_axis", nav.rpol));
    w.put_att(gipV, Att::f64("inverse_flattening", (double)s.inverse));
    w.put_att(gipV, Att::f64("latitude_of_projection_origin", (double)s.lat0));
    w.put_att(gipV, Att::f64("longitude_of_projection_origin", (double)s.lpo));
    w.put_att(gipV, Att::text("sweep_angle_axis", "x"));
    w.put_att(ofV, Att::text("long_name", "Optical Flow Settings"));
    w.put_att(ofV, Att::text("key", "1 = Modified Zimmer et al. (2011), 2 = Farneback, 3 = Brox (2004), 4 = Least Squares"));
    w.put_att(ofV, Att::f32("Image2_xOffset", nav.g2xOffset));
    w.put_att(ofV, Att::f32("Image2_yOffset", nav.g2yOffset));
    w.put_att(ofV, Att::f64("lambda", a.lambda));
    w.put_att(ofV, Att::f64("lambdac", a.lambdac));
    w.put_att(ofV, Att::f64("alpha", a.alpha));
    w.put_att(ofV, Att::f64("filtsigma", a.filtsigma));
    w.put_att(ofV, Att::f64("ScaleF", a.scaleF));
    w.put_att(ofV, Att::i32("K_Iterations", a.kiters));
    w.put_att(ofV, Att::i32("L_Iterations", a.liters));
    w.put_att(ofV, Att::i32("M_Iterations", a.miters));
    w.put_att(ofV, Att::i32("CG_Iterations", a.cgiters));
    w.put_att(ofV, Att::f32("NormMax", a.NormMax));
    w.put_att(ofV, Att::f32("NormMin", a.NormMin));
    w.put_att(ofV, Att::i32("dofirstguess", a.dofirstguess));
    w.put_att(ofV, Att::f32("dt_seconds", dT));
    if (w.enddef()) { *err = w.error(); return false; }
    const uint64_t n = (uint64_t)s.nx * s.ny;
    int rc = 0;
    rc |= w.put_var(xv, s.x.data(), s.nx);
    rc |= w.put_var(yv, s.y.data(), s.ny);
    rc |= w.put_var(tv, &s.t, 1);
    if (a.outnav) { rc |= w.put_var(uV, U, n); rc |= w.put_var(vV, V, n); }
    if (a.outraw) { rc |= w.put_var(urV, Ur, n); rc |= w.put_var(vrV, Vr, n); }
    if (a.pixuv == 1) { rc |= w.put_var(upV, upix, n); rc |= w.put_var(vpV, vpix, n); }
    if (ctpV >= 0) rc |= w.put_var(ctpV, ctp, n);
    if (a.outrad) {
        rc |= w.put_var(radV, s.rad.data(), n);
        if (e2) rc |= w.put_var(rad2V, e2->rad.data(), n);
        if (e3) rc |= w.put_var(rad3V, e3->rad.data(), n);
    }
    const int gv = (int)s.gipVal, ofv = a.oftype;
    rc |= w.put_var(gipV, &gv, 1);
    rc |= w.put_var(ofV, &ofv, 1);
    if (a.outrad) {
        const Scene* sc[3] = { &s, e2, e3 };
        for (int c = 0; c < 3; c++) {
            if (!sc[c]) continue;
            const float pv[5] = { sc[c]->fk1, sc[c]->fk2, sc[c]->bc1, sc[c]->bc2, sc[c]->kap1 };
            for (int k = 0; k < 5; k++) rc |= w.put_var(pk[c][k], &pv[k], 1);
        }
    }
    if (rc || w.close()) { *err = w.error(); return false; }
    return true;
}

// oct_polarwrite (src/oct_filewrite.cc:353-558) and oct_mercwrite (:560-705).  Both declare U and V as doubles:
// the polar file holds the pixel displacements uPix / vPix there, the Mercator file the navigated shorts
// (scale_factor 0.01); the solver settings are written for the Zimmer solver only (oftype == 1).
bool write_grid(const std::string& path, bool polar, const Scene& s, const Flags& a, float dT, const short* U, const short* V,
                const float* upix, const float* vpix, std::string* err)
{
    using cdf::Att;
    cdf::Writer w;
    if (w.create(path)) { *err = w.error(); return false; }
    const int xd = w.add_dim("x", s.nx), yd = w.add_dim("y", s.ny);
    const int xv = w.add_var("x", cdf::SHORT, { xd }), yv = w.add_var("y", cdf::SHORT, { yd });
    w.put_att(xv, Att::f32("scale_factor", s.xScale)); w.put_att(xv, Att::f32("add_offset", s.xOffset));
    w.put_att(yv, Att::f32("scale_factor", s.yScale)); w.put_att(yv, Att::f32("add_offset", s.yOffset));
    const int tv = w.add_var("t", cdf::DOUBLE, {});
    w.put_att(tv, Att::text("standard_name", "time"));
    w.put_att(tv, Att::text("units", s.tUnits));
    w.put_att(tv, Att::text("axis", "T"));
    w.put_att(tv, Att::text("bounds", "time_bounds"));
    w.put_att(tv, Att::text("long_name", "J2000 epoch mid-point between the start and end image scan in seconds"));
    const std::vector<int> yx = { yd, xd };
    const int uV = w.add_var("U", cdf::DOUBLE, yx), vV = w.add_var("V", cdf::DOUBLE, yx);
    int upV = -1, vpV = -1, radV = -1;
    if (a.pixuv == 1) { upV = w.add_var("Upix", cdf::FLOAT, yx); vpV = w.add_var("Vpix", cdf::FLOAT, yx); }
    if (a.outrad) radV = w.add_var("Rad", cdf::FLOAT, yx);
    const int gipV = w.add_var(polar ? "polar_imager_projection" : "merc_imager_projection", cdf::INT, {});
    const int ofV = w.add_var("optical_flow_settings", cdf::INT, {});
    const char* gm = polar ? "polar_orthonormal" : "Mercator Sphere";
    w.put_att(uV, Att::text("long_name", "U")); w.put_att(uV, Att::text("grid_mapping", gm));
    if (!polar) w.put_att(uV, Att::f32("scale_factor", 0.01f));
    w.put_att(uV, Att::text("units", a.pixuv == 0 ? "meters per second" : "x-pixels"));
    w.put_att(vV, Att::text("long_name", "V")); w.put_att(vV, Att::text("grid_mapping", gm));
    if (!polar) w.put_att(vV, Att::f32("scale_factor", 0.01f));
    w.put_att(vV, Att::text("units", a.pixuv == 1 ? "y-pixels" : "meters per second"));
    if (a.outrad) { w.put_att(radV, Att::text("long_name", "Rad")); w.put_att(radV, Att::text("grid_mapping", gm)); }
    if (polar) {
        w.put_att(gipV, Att::text("long_name", "Polar_Orthonormal_Grid"));
        w.put_att(gipV, Att::text("grid_mapping_name", "polar"));
        w.put_att(gipV, Att::f64("lat1", (double)s.lat1));
        w.put_att(gipV, Att::f64("lon0", (double)s.lon0));
    } else {
        w.put_att(gipV, Att::text("long_name", "Mercator_Grid"));
        w.put_att(gipV, Att::text("grid_mapping_name", "Mercator"));
        w.put_att(gipV, Att::f64("lon1", (double)s.lon1));
    }
    w.put_att(gipV, Att::f64("R", (double)s.R));
    w.put_att(ofV, Att::text("long_name", "Optical Flow Settings"));
    w.put_att(ofV, Att::text("key", "1 = Modified Sun (2014), 2 = Farneback, 3 = Brox (2004)"));
    if (a.oftype == 1) {
        w.put_att(ofV, Att::f64("lambda", a.lambda)); w.put_att(ofV, Att::f64("lambdac", a.lambdac));
        w.put_att(ofV, Att::f64("alpha", a.alpha)); w.put_att(ofV, Att::f64("filtsigma", a.filtsigma));
        w.put_att(ofV, Att::f64("ScaleF", a.scaleF));
        w.put_att(ofV, Att::i32("K_Iterations", a.kiters)); w.put_att(ofV, Att::i32("L_Iterations", a.liters));
        w.put_att(ofV, Att::i32("M_Iterations", a.miters)); w.put_att(ofV, Att::i32("CG_Iterations", a.cgiters));
        w.put_att(ofV, Att::f32("NormMax", a.NormMax)); w.put_att(ofV, Att::f32("NormMin", a.NormMin));
        w.put_att(ofV, Att::i32("dofirstguess", a.dofirstguess));
    }
    w.put_att(ofV, Att::f32("dt_seconds", dT));
    if (w.enddef()) { *err = w.error(); return false; }
    const uint64_t n = (uint64_t)s.nx * s.ny;
    std::vector<double> du(n), dv(n);
    for (uint64_t k = 0; k < n; k++) { du[k] = polar ? (double)upix[k] : (double)U[k]; dv[k] = polar ? (double)vpix[k] : (double)V[k]; }
    int rc = 0;
    rc |= w.put_var(xv, s.x.data(), s.nx);
    rc |= w.put_var(yv, s.y.data(), s.ny);
    rc |= w.put_var(tv, &s.t, 1);
    rc |= w.put_var(uV, du.data(), n);
    rc |= w.put_var(vV, dv.data(), n);
    if (a.pixuv == 1) { rc |= w.put_var(upV, upix, n); rc |= w.put_var(vpV, vpix, n); }
    if (a.outrad) rc |= w.put_var(radV, s.data.data(), n);
    const int gv = (int)s.gipVal, ofv = a.oftype;
    rc |= w.put_var(gipV, &gv, 1);
    rc |= w.put_var(ofV, &ofv, 1);
    if (rc || w.close()) { *err = w.error(); return false; }
    return true;
}

int fail(const std::string& msg)
{
    fprintf(stderr, "octane: %s\n", msg.c_str());
    return 1;
}

}  // namespace

static int run(int argc, char* argv[])
{
    Flags args;
    std::string f1, f2, f1c, f2c, f1fg, fc21, fc22 = "none", fc31, fc32 = "none", interploc = "./interpolation", outdir = "./";
    printf("Beginning variational dense optical flow...\n");
    if (argc < 4 && !(argc >= 2 && !strcmp(argv[argc - 1], "-dump_settings"))) {
        usage();
        return 0;
    }
    // src/main.cc:166-350: string compare on every argument, value = the following argument
    for (int i = 0; i < argc; ++i) {
        const std::string s = argv[i];
        const char* nxt = (i + 1 < argc) ? argv[i + 1] : "";
        if (s == "-i1") f1 = nxt;
        if (s == "-i2") f2 = nxt;
        if (s == "-i1cth") { f1c = nxt; args.doCTH = 1; }
        if (s == "-i2cth") f2c = nxt;
        if (s == "-farn") {
            args.farn = 1;
            printf("Farneback disabled for this version of OCTANE, run without -farn, exiting...");
            return 0;
        }
        if (s == "-pd") args.pixuv = 1;
        if (s == "-srsal") args.dosrsal = 1;
        if (s == "-Polar") { args.dopolar = 1; args.ftype = "POLAR"; }
        if (s == "-Merc") { args.domerc = 1; args.ftype = "MERC"; }
        if (s == "-ahi") args.doahi = 1;
        if (s == "-ir") args.ir = 1;
        if (s == "-sosm") args.dososm = 1;
        if (s == "-interp") args.dointerp = 1;
        if (s == "-ic21") { args.doc2 = 1; fc21 = nxt; }
        if (s == "-ic22") fc22 = nxt;
        if (s == "-ic31") { args.doc3 = 1; fc31 = nxt; }
        if (s == "-ic32") fc32 = nxt;
        if (s == "-alpha") args.alpha = atof(nxt);
        if (s == "-lambda") args.lambda = atof(nxt);
        if (s == "-scsig") args.scsig = atof(nxt) * atof(nxt);
        if (s == "-alpha2") args.alpha2 = atof(nxt);
        if (s == "-lambdac") args.lambdac = atof(nxt);
        if (s == "-nncth") args.interpcth = 0;
        if (s == "-inv") args.doinv = 1;
        if (s == "-ctt") args.doctt = 1;
        if (s == "-kiters") args.kiters = atoi(nxt);
        if (s == "-liters") args.liters = atoi(nxt);
        if (s == "-brox") args.dozim = 0;
        if (s == "-corn") args.docorn = 0;
        if (s == "-firstguess") { args.dofirstguess = 1; f1fg = nxt; }
        if (s == "-rad") args.rad = atoi(nxt);
        if (s == "-srad") args.srad = atoi(nxt);
        if (s == "-deltat") args.deltat = (float)atof(nxt);
        if (s == "-interploc") interploc = nxt;
        if (s == "-no_outnav") args.outnav = false;
        if (s == "-no_outraw") args.outraw = false;
        if (s == "-no_outrad") args.outrad = false;
        if (s == "-no_outctp") args.outctp = false;
        if (s == "-set_device") args.setdevice = atoi(nxt) - 1;
        if (s == "-normmax") { args.NormMax = (float)atof(nxt); args.setNormMax = false; }
        if (s == "-normmin") { args.NormMin = (float)atof(nxt); args.setNormMin = false; }
        if (s == "-normmax2") { args.NormMax2 = (float)atof(nxt); args.setNormMax2 = false; }
        if (s == "-normmin2") { args.NormMin2 = (float)atof(nxt); args.setNormMin2 = false; }
        if (s == "-normmax3") { args.NormMax3 = (float)atof(nxt); args.setNormMax3 = false; }
        if (s == "-normmin3") { args.NormMin3 = (float)atof(nxt); args.setNormMin3 = false; }
        if (s == "-o") outdir = nxt;
        if (s == "-dump_settings") args.dump_settings = 1;
        if (s == "-dry_run") args.dry_run = 1;
    }
    // :362-392
    args.oftype = (args.dozim == 0) ? 3 : 1;
    if (args.dososm == 1) args.oftype = 4;
    if (args.dopolar == 1 || args.domerc == 1 || args.doahi == 1) args.doCTH = 0;
    if (args.dump_settings) {
        printf("i1=%s\ni2=%s\ni1cth=%s\nfirstguess=%s\no=%s\n", f1.c_str(), f2.c_str(), f1c.c_str(), f1fg.c_str(), outdir.c_str());
        printf("alpha=%.17g\nlambda=%.17g\nlambdac=%.17g\nscaleF=%.17g\nscsig=%.17g\nfiltsigma=%.17g\n", args.alpha, args.lambda,
               args.lambdac, args.scaleF, args.scsig, args.filtsigma);
        printf("kiters=%d\nliters=%d\ncgiters=%d\nmiters=%d\ndozim=%d\noftype=%d\npixuv=%d\ndoCTH=%d\nir=%d\ndofirstguess=%d\n",
               args.kiters, args.liters, args.cgiters, args.miters, args.dozim, args.oftype, args.pixuv, args.doCTH, args.ir,
               args.dofirstguess);
        printf("setdevice=%d\noutnav=%d\noutraw=%d\noutrad=%d\noutctp=%d\ninterpcth=%d\ndopolar=%d\ndomerc=%d\ndososm=%d\ndocorn=%d\n",
               args.setdevice, (int)args.outnav, (int)args.outraw, (int)args.outrad, (int)args.outctp, args.interpcth,
               args.dopolar, args.domerc, args.dososm, args.docorn);
        return 0;
    }
    if ((args.dopolar || args.domerc) && (args.doc2 || args.doc3 || args.dofirstguess))
        return fail("extra channels / first guess with -Polar / -Merc are not part of this build");
    if (args.dososm) return fail("-sosm (CPU patch-match solver) is not part of this build");
    if ((args.doc2 && fc22 == "none") || (args.doc3 && fc32 == "none")) {       // src/main.cc:352-361
        printf("Missing files for second / third channel...stopping \n");
        return 0;
    }
    if (args.dosrsal && (args.dopolar || args.domerc)) return fail("-srsal with -Polar / -Merc is not part of this build");
    // oct_optical_flow.cc:100-105 hands goesData.CTHVal to the smoother whether or not it was read
    if (args.dosrsal && args.doCTH != 1) return fail("-srsal needs cloud-top heights (-i1cth)");
    if (args.dointerp) printf("Warning: -interp is not part of this build; only outfile.nc is written\n");

    printf("Here are the file names being used: \nFile 1 : %s\nFile 2 : %s\n", f1.c_str(), f2.c_str());
    std::string err;
    Scene g1, g2;
    if (args.dopolar || args.domerc) {
        // ---- projected grids: oct_polarread / oct_mercread -> ingest -> flow -> pix2uv (polar / Mercator branch)
        const bool polar = args.dopolar == 1;
        if (!read_grid(f1, polar, g1, &err) || !read_grid(f2, polar, g2, &err)) return fail(err);
        if (g1.nx != g2.nx || g1.ny != g2.ny) return fail("the two images differ in size");
        const int gnx = g1.nx, gny = g1.ny;
        const size_t gn = (size_t)gnx * gny;
        const std::string gout = outdir + (polar ? "outfile_polar.nc" : "outfile_merc.nc");      // main.cc:442-444
        std::vector<short> U(gn, 0), V(gn, 0), Ur(gn, 0), Vr(gn, 0);
        std::vector<float> up(gn, 0.f), vp(gn, 0.f);
        float dT = (float)(g2.t - g1.t);
        if (args.dry_run) {
            g1.data = g1.radf;
            if (!write_grid(gout, polar, g1, args, dT, U.data(), V.data(), up.data(), vp.data(), &err)) return fail(err);
            printf("%s written (dry run: no motion computed)\n", gout.c_str());
            return 0;
        }
        octane_ctx* gctx = nullptr;
        int grc = octane_ctx_create(&gctx, args.setdevice);
        if (grc == OCTANE_ENODEV) { printf("No gpus available for use, exiting\n"); return 0; }
        if (grc) return fail(octane_last_error());
        octane_nav gnav;
        memset(&gnav, 0, sizeof gnav);
        gnav.xScale = g1.xScale; gnav.xOffset = g1.xOffset; gnav.yScale = g1.yScale; gnav.yOffset = g1.yOffset;
        gnav.g2xOffset = g1.xOffset; gnav.g2yOffset = g1.yOffset;      // main.cc:400-405: the sector guard is GOES-only
        gnav.lat1 = g1.lat1; gnav.lon0 = g1.lon0; gnav.lon1 = g1.lon1; gnav.R = g1.R;
        Scene* pair[2] = { &g1, &g2 };
        for (int k = 0; k < 2; k++) {
            Scene& sc = *pair[k];
            sc.data.resize(gn); sc.lat.resize(gn); sc.lon.resize(gn);
            if (octane_navcal_grid(gctx, polar ? 1 : 2, sc.radf.data(), sc.x.data(), sc.y.data(), gnx, gny, &gnav, k == 0,
                                   sc.data.data(), sc.lat.data(), sc.lon.data()) < 0)
                return fail(octane_last_error());
        }
        octane_params gp;
        octane_params_default(&gp);
        gp.alpha = args.alpha; gp.lambda = args.lambda; gp.lambdac = args.lambdac; gp.scaleF = args.scaleF; gp.scsig = args.scsig;
        gp.kiters = args.kiters; gp.liters = args.liters; gp.cgiters = args.cgiters; gp.dozim = args.dozim;
        gp.setdevice = args.setdevice; gp.pixuv = args.pixuv; gp.dopolar = args.dopolar; gp.domerc = args.domerc;
        grc = octane_optical_flow(gctx, g1.data.data(), g2.data.data(), nullptr, gnx, gny, 1, &gnav, g1.t, g2.t, &gp, up.data(),
                                  vp.data(), U.data(), V.data(), Ur.data(), Vr.data(), nullptr, &dT);
        if (grc < 0) return fail(octane_last_error());
        if (!write_grid(gout, polar, g1, args, dT, U.data(), V.data(), up.data(), vp.data(), &err)) return fail(err);
        printf("%s written\n", gout.c_str());
        octane_ctx_destroy(gctx);
        printf("OCTANE completed, exiting\n");
        return 0;
    }
    if (!read_goes(f1, g1, &err) || !read_goes(f2, g2, &err)) return fail(err);
    if (g1.nx != g2.nx || g1.ny != g2.ny) return fail("the two images differ in size");
    const int nx = g1.nx, ny = g1.ny;
    const size_t n = (size_t)nx * ny;

    if (args.dry_run) {
        octane_nav nav0;
        memset(&nav0, 0, sizeof nav0);
        nav0.pph = g1.pph; nav0.req = g1.req; nav0.rpol = g1.rpol; nav0.lam0 = g1.lam0;
        nav0.xScale = g1.xScale; nav0.xOffset = g1.xOffset; nav0.yScale = g1.yScale; nav0.yOffset = g1.yOffset;
        nav0.g2xOffset = g2.xOffset; nav0.g2yOffset = g2.yOffset;
        if (octane_band_minmax(g1.band, &args.NormMax, &args.NormMin)) return fail("band_id outside 1..16");
        std::vector<short> z(n, 0);
        std::vector<float> zf(n, 0.f), cth0;
        int cx = 0, cy = 0;
        if (args.doCTH == 1 && (!read_plane(f1c, "Cloud_Top_Height_Effective", &cx, &cy, cth0, &err) ||
                                !regrid(nullptr, "cloud-top height field", cth0, cx, cy, nx, ny, args.interpcth, &err)))
            return fail(err);
        for (size_t k = 0; k < cth0.size(); k++) z[k] = args.ir == 1 ? (short)((cth0[k] - 300) * 100) : (short)cth0[k];
        std::vector<short> zero(n, 0);
        const std::string outname0 = outdir + "outfile.nc";
        if (!write_goes(outname0, g1, args, nav0, (float)(g2.t - g1.t), zero.data(), zero.data(), zero.data(), zero.data(),
                        zf.data(), zf.data(), args.doCTH == 1 ? z.data() : nullptr, &err))
            return fail(err);
        printf("%s written (dry run: no motion computed)\n", outname0.c_str());
        return 0;
    }
    octane_ctx* ctx = nullptr;
    int rc = octane_ctx_create(&ctx, args.setdevice);
    if (rc == OCTANE_ENODEV) { printf("No gpus available for use, exiting\n"); return 0; }      // .cu:1255-1259
    if (rc) return fail(octane_last_error());

    // navigation constants: the reader keeps req, rpol, pph, lam0 in floats (oct_fileread.cc:51)
    octane_nav nav;
    memset(&nav, 0, sizeof nav);
    nav.pph = g1.pph; nav.req = g1.req; nav.rpol = g1.rpol; nav.lam0 = g1.lam0;
    nav.xScale = g1.xScale; nav.xOffset = g1.xOffset; nav.yScale = g1.yScale; nav.yOffset = g1.yOffset;
    nav.g2xOffset = g2.xOffset; nav.g2yOffset = g2.yOffset;                                      // main.cc:401-405
    nav.minX = 0; nav.minY = 0;

    // ingest (oct_fileread.cc:341-388): band table range, 0..255, cal "RAW"; navigation for image 1 only
    auto ingest = [&](Scene& s, int donav, int channel) -> bool {
        octane_cal cal;
        memset(&cal, 0, sizeof cal);
        cal.radScale = s.radScale; cal.radOffset = s.radOffset;
        cal.fk1 = s.fk1; cal.fk2 = s.fk2; cal.bc1 = s.bc1; cal.bc2 = s.bc2; cal.kap1 = s.kap1;
        if (octane_band_minmax(s.band, &cal.maxin, &cal.minin)) { err = "band_id outside 1..16"; return false; }
        cal.maxout = 255.f; cal.minout = 0.f; cal.H = s.pph + s.req; cal.cal = 0; cal.donav = donav;
        if (donav && channel == 1) {
            if (args.setNormMax) args.NormMax = cal.maxin;
            if (args.setNormMin) args.NormMin = cal.minin;
        }
        octane_nav ns = nav;
        ns.req = s.req; ns.rpol = s.rpol; ns.pph = s.pph; ns.lam0 = s.lam0;
        ns.xScale = s.xScale; ns.xOffset = s.xOffset; ns.yScale = s.yScale; ns.yOffset = s.yOffset;
        const size_t m = (size_t)s.nx * s.ny;
        s.data.resize(m); s.lat.resize(m); s.lon.resize(m);
        if (octane_navcal(ctx, s.rad.data(), s.x.data(), s.y.data(), s.nx, s.ny, &ns, &cal, s.data.data(), s.lat.data(),
                          s.lon.data()) < 0) { err = octane_last_error(); return false; }
        return true;
    };
    if (!ingest(g1, 1, 1) || !ingest(g2, 0, 1)) return fail(err);

    // extra channels (-ic21/-ic22, -ic31/-ic32; main.cc:414-436, oct_fileread.cc:359-380): channel planes
    // behind channel 1, each brought to channel 1's grid
    const int nc = 1 + args.doc2 + args.doc3;
    std::vector<float> img1(g1.data), img2(g2.data);
    Scene extra[2][2];                     // [channel 2 / 3][image 1 / 2]
    bool extra_same_grid[2] = { false, false };
    {
        const std::string* files[2][2] = { { &fc21, &fc22 }, { &fc31, &fc32 } };
        const int on[2] = { args.doc2, args.doc3 };
        for (int ch = 0; ch < 2; ch++) {
            if (!on[ch]) continue;
            for (int im = 0; im < 2; im++) {
                Scene& e = extra[ch][im];
                if (!read_goes(*files[ch][im], e, &err) || !ingest(e, im == 0, ch + 2)) return fail(err);
                std::vector<float> plane(e.data);
                if (!regrid(ctx, "channel " + std::to_string(ch + 2), plane, e.nx, e.ny, nx, ny, 1, &err)) return fail(err);
                std::vector<float>& dst = im == 0 ? img1 : img2;
                dst.insert(dst.end(), plane.begin(), plane.end());
            }
            extra_same_grid[ch] = extra[ch][0].nx == nx && extra[ch][0].ny == ny;
        }
    }

    octane_params p;
    octane_params_default(&p);
    p.alpha = args.alpha; p.lambda = args.lambda; p.lambdac = args.lambdac; p.scaleF = args.scaleF; p.scsig = args.scsig;
    p.kiters = args.kiters; p.liters = args.liters; p.cgiters = args.cgiters; p.dozim = args.dozim;
    p.setdevice = args.setdevice; p.pixuv = args.pixuv; p.doCTH = args.doCTH; p.ir = args.ir;
    p.first_guess = args.dofirstguess; p.dosrsal = args.dosrsal;

    std::vector<float> cth, upix(n, 0.f), vpix(n, 0.f);
    int cx = 0, cy = 0;
    if (args.doCTH == 1 && (!read_plane(f1c, "Cloud_Top_Height_Effective", &cx, &cy, cth, &err) ||
                            !regrid(ctx, "cloud-top height field", cth, cx, cy, nx, ny, args.interpcth, &err)))
        return fail(err);
    if (args.dofirstguess == 1) {
        // oct_fgread (oct_fileread.cc:817-859) + oct_uv2pix (oct_optical_flow.cc:51-53)
        int ux = 0, uy = 0, vx = 0, vy = 0;
        if (!read_plane(f1fg, "UFG", &ux, &uy, upix, &err) || !read_plane(f1fg, "VFG", &vx, &vy, vpix, &err)) return fail(err);
        if (ux != nx || uy != ny || vx != nx || vy != ny) return fail("first-guess file must have the image's dimensions (offlags.h:13)");
        if (octane_uv2pix(ctx, &nav, g1.t, g2.t, g1.lat.data(), g1.lon.data(), g1.x.data(), g1.y.data(), nx, ny, &p,
                          upix.data(), vpix.data()) < 0)
            return fail(octane_last_error());
    }

    std::vector<short> U(n), V(n), Ur(n), Vr(n), ctp(args.doCTH == 1 ? n : 0);
    float dT = 0.f;
    rc = octane_optical_flow(ctx, img1.data(), img2.data(), args.doCTH == 1 ? cth.data() : nullptr, nx, ny, nc, &nav,
                             g1.t, g2.t, &p, upix.data(), vpix.data(), U.data(), V.data(), Ur.data(), Vr.data(),
                             args.doCTH == 1 ? ctp.data() : nullptr, &dT);
    if (rc < 0) return fail(octane_last_error());
    if (rc == 1)
        printf("MOVE WARNING: Sector Moved, setting motions to 0 %g %g %g %g\n", nav.xOffset, nav.g2xOffset, nav.yOffset, nav.g2yOffset);

    const std::string outname = outdir + "outfile.nc";
    const Scene* ex[2] = { (args.doc2 && extra_same_grid[0]) ? &extra[0][0] : nullptr,
                           (args.doc3 && extra_same_grid[1]) ? &extra[1][0] : nullptr };
    if ((args.doc2 && !ex[0]) || (args.doc3 && !ex[1]))
        printf("Warning: Rad2/Rad3 are written only for channels on channel 1's grid\n");
    if (!write_goes(outname, g1, args, nav, dT, U.data(), V.data(), Ur.data(), Vr.data(), upix.data(), vpix.data(),
                    args.doCTH == 1 ? ctp.data() : nullptr, &err, ex))
        return fail(err);
    printf("%s written\n", outname.c_str());
    octane_ctx_destroy(ctx);
    printf("OCTANE completed, exiting\n");
    return 0;
}

int main(int argc, char* argv[])
{
    try {
        return run(argc, argv);
    } catch (const std::bad_alloc&) {          // e.g. a scene larger than host memory
        fprintf(stderr, "octane: out of host memory\n");
    } catch (const std::exception& e) {
        fprintf(stderr, "octane: %s\n", e.what());
    }
    return 1;
}
