/* tests/stubs/fake_netcdf.c -- a table-driven stand-in for the netCDF C library (the calls of tests/stubs/netcdf.h)
 * serving one small in-memory dataset laid out like a GOES-R ABI L1b radiance file.  Test infrastructure only. */
#include "netcdf.h"

#include <string.h>

typedef struct { const char* name; nc_type type; size_t len; const void* data; } FAtt;
typedef struct { const char* name; nc_type type; int ndims; int dimids[2]; int natts; const FAtt* atts; const void* data; } FVar;

static const float rad_scale = 0.8121064f, rad_off = -25.93664f, xy_scale[2] = { 5.6e-05f, -5.6e-05f }, xy_off[2] = { -0.101332f, 0.128212f };
static const short rad_fill = 4095;
static const double gip_h = 35786023.0, gip_a = 6378137.0, gip_b = 6356752.31414, gip_lon = -75.0;
static const long long big = 1234567890123LL;
static const short rad12[48] = {                     /* "Rad" as operational files hold it: short counts, 12 bits */
    0, 1, 2, 3, 4, 5, 6, 7, 100, 200, 300, 400, 500, 600, 700, 800, 1000, 1001, 1002, 1003, 1004, 1005, 1006, 1007,
    4095, 4094, 4093, 4092, 4091, 4090, 4089, 4088, 9, 8, 7, 6, 5, 4, 3, 2, 2040, 2041, 2042, 2043, 2044, 2045, 2046, 2047 };
static const unsigned short rad[48] = {            /* "RadU": stored as ushort, values above 32767 must survive as numbers */
    0, 1, 2, 3, 4, 5, 6, 7, 100, 200, 300, 400, 500, 600, 700, 800, 1000, 1001, 1002, 1003, 1004, 1005, 1006, 1007,
    4095, 4094, 4093, 4092, 4091, 4090, 4089, 4088, 9, 8, 7, 6, 5, 4, 3, 2, 40000, 40001, 40002, 40003, 40004, 40005, 40006, 40007 };
static const short xs[8] = { 0, 1, 2, 3, 4, 5, 6, 7 }, ys[6] = { 10, 11, 12, 13, 14, 15 };
static const double tval = 7.123456789e8;
static const signed char band = 13;
static const unsigned char dqf[48] = { 0, 1, 2, 3, 255 };
static const float kappa = 0.0123f, fk1 = 202263.0f, fk2 = 3698.19f, bc1 = 0.43361f, bc2 = 0.99939f;
static const double gip_inv = 298.2572221, gip_lat0 = 0.0;
static const int zero = 0;

static const FAtt rad_atts[] = { { "scale_factor", NC_FLOAT, 1, &rad_scale }, { "add_offset", NC_FLOAT, 1, &rad_off },
                                 { "_FillValue", NC_SHORT, 1, &rad_fill }, { "_Unsigned", NC_CHAR, 4, "true" },
                                 { "long_name", NC_STRING, 1, 0 }, { "valid_count", NC_INT64, 1, &big } };
static const FAtt x_atts[] = { { "scale_factor", NC_FLOAT, 1, &xy_scale[0] }, { "add_offset", NC_FLOAT, 1, &xy_off[0] } };
static const FAtt y_atts[] = { { "scale_factor", NC_FLOAT, 1, &xy_scale[1] }, { "add_offset", NC_FLOAT, 1, &xy_off[1] } };
static const FAtt t_atts[] = { { "units", NC_CHAR, 33, "seconds since 2000-01-01 12:00:00" } };
static const FAtt gip_atts[] = { { "perspective_point_height", NC_DOUBLE, 1, &gip_h }, { "semi_major_axis", NC_DOUBLE, 1, &gip_a },
                                 { "semi_minor_axis", NC_DOUBLE, 1, &gip_b }, { "longitude_of_projection_origin", NC_DOUBLE, 1, &gip_lon },
                                 { "inverse_flattening", NC_DOUBLE, 1, &gip_inv }, { "latitude_of_projection_origin", NC_DOUBLE, 1, &gip_lat0 } };
static const FAtt g_atts[] = { { "platform_ID", NC_CHAR, 3, "G16" }, { "title", NC_STRING, 1, 0 } };
static const struct { const char* name; size_t len; } dims[3] = { { "y", 6 }, { "x", 8 }, { "band", 1 } };
static const FVar vars[] = {
    { "Rad", NC_SHORT, 2, { 0, 1 }, 6, rad_atts, rad12 },
    { "RadU", NC_USHORT, 2, { 0, 1 }, 6, rad_atts, rad },
    { "DQF", NC_UBYTE, 2, { 0, 1 }, 0, 0, dqf },
    { "x", NC_SHORT, 1, { 1, 0 }, 2, x_atts, xs },
    { "y", NC_SHORT, 1, { 0, 0 }, 2, y_atts, ys },
    { "t", NC_DOUBLE, 0, { 0, 0 }, 1, t_atts, &tval },
    { "band_id", NC_BYTE, 1, { 2, 0 }, 0, 0, &band },
    { "goes_imager_projection", NC_INT, 0, { 0, 0 }, 6, gip_atts, &zero },
    { "kappa0", NC_FLOAT, 0, { 0, 0 }, 0, 0, &kappa },
    { "planck_fk1", NC_FLOAT, 0, { 0, 0 }, 0, 0, &fk1 },
    { "planck_fk2", NC_FLOAT, 0, { 0, 0 }, 0, 0, &fk2 },
    { "planck_bc1", NC_FLOAT, 0, { 0, 0 }, 0, 0, &bc1 },
    { "planck_bc2", NC_FLOAT, 0, { 0, 0 }, 0, 0, &bc2 },
    { "algorithm_container", NC_STRING, 0, { 0, 0 }, 0, 0, 0 },
};
#define NVARS ((int)(sizeof vars / sizeof vars[0]))
static int opened = 0;

/* the second file of a pair ("...file2...") is the same scene 600 s later */
int nc_open(const char* path, int mode, int* ncidp) { (void)mode; opened++; *ncidp = 65536 + (strstr(path, "file2") ? 1 : 0); return NC_NOERR; }
int nc_close(int ncid) { (void)ncid; opened--; return NC_NOERR; }
int fake_netcdf_open_count(void) { return opened; }
int nc_inq(int ncid, int* nd, int* nv, int* na, int* ul) { (void)ncid; *nd = 3; *nv = NVARS; *na = 2; *ul = -1; return NC_NOERR; }
int nc_inq_dim(int ncid, int d, char* name, size_t* len) { (void)ncid; if (d < 0 || d > 2) return -46; strcpy(name, dims[d].name); *len = dims[d].len; return NC_NOERR; }
int nc_inq_var(int ncid, int v, char* name, nc_type* t, int* nd, int* ids, int* na)
{
    (void)ncid;
    if (v < 0 || v >= NVARS) return -49;
    strcpy(name, vars[v].name); *t = vars[v].type; *nd = vars[v].ndims; *na = vars[v].natts;
    for (int i = 0; i < vars[v].ndims; i++) ids[i] = vars[v].dimids[i];
    return NC_NOERR;
}
static const FAtt* find(int v, const char* name, int k)
{
    const FAtt* a = v == NC_GLOBAL ? g_atts : vars[v].atts;
    const int n = v == NC_GLOBAL ? 2 : vars[v].natts;
    if (!name) return k < n ? &a[k] : 0;
    for (int i = 0; i < n; i++) if (!strcmp(a[i].name, name)) return &a[i];
    return 0;
}
int nc_inq_attname(int ncid, int v, int k, char* name) { (void)ncid; const FAtt* a = find(v, 0, k); if (!a) return -43; strcpy(name, a->name); return NC_NOERR; }
int nc_inq_att(int ncid, int v, const char* name, nc_type* t, size_t* len) { (void)ncid; const FAtt* a = find(v, name, 0); if (!a) return -43; *t = a->type; *len = a->len; return NC_NOERR; }
static double elem(nc_type t, const void* p, size_t i)
{
    switch (t) {
    case NC_BYTE: return ((const signed char*)p)[i];
    case NC_UBYTE: return ((const unsigned char*)p)[i];
    case NC_SHORT: return ((const short*)p)[i];
    case NC_USHORT: return ((const unsigned short*)p)[i];
    case NC_INT: return ((const int*)p)[i];
    case NC_FLOAT: return ((const float*)p)[i];
    case NC_DOUBLE: return ((const double*)p)[i];
    case NC_INT64: return (double)((const long long*)p)[i];
    default: return 0;
    }
}
int nc_get_att_text(int ncid, int v, const char* name, char* ip) { (void)ncid; const FAtt* a = find(v, name, 0); if (!a || a->type != NC_CHAR) return -56; memcpy(ip, a->data, a->len); return NC_NOERR; }
#define GET_ATT(fn, T)                                                                                      \
    int fn(int ncid, int v, const char* name, T* ip)                                                        \
    {                                                                                                       \
        (void)ncid;                                                                                         \
        const FAtt* a = find(v, name, 0);                                                                   \
        if (!a || a->type == NC_CHAR || a->type == NC_STRING) return -56;                                   \
        for (size_t i = 0; i < a->len; i++) ip[i] = (T)elem(a->type, a->data, i);                           \
        return NC_NOERR;                                                                                    \
    }
GET_ATT(nc_get_att_schar, signed char)
GET_ATT(nc_get_att_short, short)
GET_ATT(nc_get_att_int, int)
GET_ATT(nc_get_att_float, float)
GET_ATT(nc_get_att_double, double)
#define GET_VAR(fn, T, LO, HI)                                                                              \
    int fn(int ncid, int v, T* ip)                                                                          \
    {                                                                                                       \
        if (v < 0 || v >= NVARS || !vars[v].data) return -49;                                               \
        size_t n = 1;                                                                                       \
        int rc = NC_NOERR;                                                                                  \
        for (int i = 0; i < vars[v].ndims; i++) n *= dims[vars[v].dimids[i]].len;                           \
        for (size_t i = 0; i < n; i++) {                                                                    \
            const double e = elem(vars[v].type, vars[v].data, i) + ((ncid & 1) && !strcmp(vars[v].name, "t") ? 600.0 : 0.0); \
            if (e < (LO) || e > (HI)) rc = NC_ERANGE;        /* the library reports, and still converts */  \
            ip[i] = (T)(long long)e == e ? (T)(long long)e : (T)e;                                          \
        }                                                                                                   \
        return rc;                                                                                          \
    }
GET_VAR(nc_get_var_short, short, -32768.0, 32767.0)
GET_VAR(nc_get_var_int, int, -2147483648.0, 2147483647.0)
GET_VAR(nc_get_var_float, float, -3.4e38, 3.4e38)
GET_VAR(nc_get_var_double, double, -1.7e308, 1.7e308)
const char* nc_strerror(int e) { return e == NC_NOERR ? "No error" : "fake netCDF error"; }
