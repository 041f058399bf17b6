// cdf.cc -- classic NetCDF (CDF-1 / CDF-2) reader and writer, see cdf.h.
// Replaces the reference's use of netcdf-cxx4 in src/oct_fileread.cc / src/oct_filewrite.cc.
#include "cdf.h"

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

namespace cdf {

namespace {

constexpr uint32_t TAG_DIMENSION = 10, TAG_VARIABLE = 11, TAG_ATTRIBUTE = 12;
constexpr size_t CHUNK = 1 << 20;     // elements converted per fwrite / fread

uint64_t pad4(uint64_t n) { return (n + 3) & ~(uint64_t)3; }

// host (little- or big-endian) -> big-endian, element size es
void to_be(unsigned char* dst, const unsigned char* src, size_t n, size_t es)
{
    const uint16_t probe = 1;
    const bool little = *(const unsigned char*)&probe == 1;
    if (!little || es == 1) { memcpy(dst, src, n * es); return; }
    for (size_t i = 0; i < n; i++)
        for (size_t b = 0; b < es; b++) dst[i * es + b] = src[i * es + (es - 1 - b)];
}

struct Buf {                          // header assembly
    std::vector<unsigned char> b;
    void u32(uint32_t v) { for (int s = 24; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
    void u64(uint64_t v) { for (int s = 56; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
    void name(const std::string& s)
    {
        u32((uint32_t)s.size());
        b.insert(b.end(), s.begin(), s.end());
        while (b.size() & 3) b.push_back(0);
    }
    void att(const Att& a)
    {
        name(a.name);
        u32((uint32_t)a.type);
        u32((uint32_t)a.nelems());
        const size_t es = type_size(a.type), at = b.size();
        b.resize(at + a.raw.size());
        to_be(b.data() + at, a.raw.data(), a.nelems(), es);
        while (b.size() & 3) b.push_back(0);
    }
    void att_list(const std::vector<Att>& v)
    {
        if (v.empty()) { u32(0); u32(0); return; }
        u32(TAG_ATTRIBUTE);
        u32((uint32_t)v.size());
        for (auto& a : v) att(a);
    }
};

}  // namespace

size_t type_size(int t)
{
    switch (t) {
        case BYTE: case CHAR: return 1;
        case SHORT: return 2;
        case INT: case FLOAT: return 4;
        case DOUBLE: return 8;
    }
    return 0;
}

Att Att::text(const std::string& name, const std::string& v)
{
    Att a; a.name = name; a.type = CHAR; a.raw.assign(v.begin(), v.end()); return a;
}
Att Att::f32(const std::string& name, float v)
{
    Att a; a.name = name; a.type = FLOAT; a.raw.resize(4); memcpy(a.raw.data(), &v, 4); return a;
}
Att Att::f64(const std::string& name, double v)
{
    Att a; a.name = name; a.type = DOUBLE; a.raw.resize(8); memcpy(a.raw.data(), &v, 8); return a;
}
Att Att::i32(const std::string& name, int v)
{
    Att a; a.name = name; a.type = INT; a.raw.resize(4); memcpy(a.raw.data(), &v, 4); return a;
}
double Att::as_double(size_t k) const
{
    if (k >= nelems()) return 0.0;
    const unsigned char* p = raw.data() + k * type_size(type);
    switch (type) {
        case BYTE: return (double)*(const signed char*)p;
        case CHAR: return (double)*p;
        case SHORT: { short v; memcpy(&v, p, 2); return v; }
        case INT: { int v; memcpy(&v, p, 4); return v; }
        case FLOAT: { float v; memcpy(&v, p, 4); return v; }
        case DOUBLE: { double v; memcpy(&v, p, 8); return v; }
    }
    return 0.0;
}
std::string Att::as_text() const { return std::string(raw.begin(), raw.end()); }

const Att* Var::att(const std::string& n) const
{
    for (auto& a : atts) if (a.name == n) return &a;
    return nullptr;
}

// ---------------------------------------------------------------------------------- writer
Writer::~Writer() { close(); }

int Writer::create(const std::string& path)
{
    close();
    dims_.clear(); vars_.clear(); defined_ = false;
    fp_ = fopen(path.c_str(), "wb");
    if (!fp_) { err_ = "cannot create " + path; return -1; }
    return 0;
}

int Writer::add_dim(const std::string& name, uint64_t len)
{
    Dim d; d.name = name; d.len = len;
    dims_.push_back(d);
    return (int)dims_.size() - 1;
}

int Writer::add_var(const std::string& name, int type, const std::vector<int>& dimids)
{
    Var v; v.name = name; v.type = type; v.dimids = dimids; v.nelems = 1;
    for (int d : dimids) v.nelems *= dims_[d].len;
    vars_.push_back(v);
    return (int)vars_.size() - 1;
}

void Writer::put_att(int varid, const Att& a)
{
    for (auto& e : vars_[varid].atts)
        if (e.name == a.name) { e = a; return; }
    vars_[varid].atts.push_back(a);
}

int Writer::enddef()
{
    if (!fp_ || defined_) { err_ = "enddef: not open or already defined"; return -1; }
    // two passes: the header's size does not depend on the offsets it contains
    uint64_t hdr = 0;
    Buf out;
    for (int pass = 0; pass < 2; pass++) {
        Buf h;
        h.b.push_back('C'); h.b.push_back('D'); h.b.push_back('F'); h.b.push_back(2);   // 64-bit offsets
        h.u32(0);                                                                       // no records
        if (dims_.empty()) { h.u32(0); h.u32(0); }
        else {
            h.u32(TAG_DIMENSION); h.u32((uint32_t)dims_.size());
            for (auto& d : dims_) {
                if (d.len > 0xffffffffull) { err_ = "dimension too long for the classic format"; return -1; }
                h.name(d.name); h.u32((uint32_t)d.len);
            }
        }
        h.u32(0); h.u32(0);                                                             // no global attributes
        if (vars_.empty()) { h.u32(0); h.u32(0); }
        else {
            h.u32(TAG_VARIABLE); h.u32((uint32_t)vars_.size());
            uint64_t off = pad4(hdr);
            for (auto& v : vars_) {
                h.name(v.name);
                h.u32((uint32_t)v.dimids.size());
                for (int d : v.dimids) h.u32((uint32_t)d);
                h.att_list(v.atts);
                h.u32((uint32_t)v.type);
                const uint64_t vsize = pad4(v.nelems * type_size(v.type));
                h.u32(vsize > 0xfffffffcull ? 0xffffffffu : (uint32_t)vsize);
                v.begin = off;
                h.u64(off);
                off += vsize;
            }
        }
        hdr = h.b.size();
        out = h;
    }
    FILE* f = (FILE*)fp_;
    if (fwrite(out.b.data(), 1, out.b.size(), f) != out.b.size()) { err_ = "short write (header)"; return -1; }
    defined_ = true;
    return 0;
}

int Writer::put_var(int varid, const void* data, uint64_t n)
{
    if (!fp_ || !defined_) { err_ = "put_var before enddef"; return -1; }
    const Var& v = vars_[varid];
    if (n != v.nelems) { err_ = "put_var: element count mismatch for " + v.name; return -1; }
    FILE* f = (FILE*)fp_;
    if (fseeko(f, (off_t)v.begin, SEEK_SET) != 0) { err_ = "seek failed"; return -1; }
    const size_t es = type_size(v.type);
    std::vector<unsigned char> tmp(std::min<uint64_t>(n, CHUNK) * es);
    const unsigned char* src = (const unsigned char*)data;
    for (uint64_t done = 0; done < n;) {
        const size_t m = (size_t)std::min<uint64_t>(CHUNK, n - done);
        to_be(tmp.data(), src + done * es, m, es);
        if (fwrite(tmp.data(), es, m, f) != m) { err_ = "short write (" + v.name + ")"; return -1; }
        done += m;
    }
    const uint64_t bytes = n * es, padded = pad4(bytes);
    const unsigned char zero[4] = { 0, 0, 0, 0 };
    if (padded > bytes && fwrite(zero, 1, (size_t)(padded - bytes), f) != padded - bytes) { err_ = "short write"; return -1; }
    return 0;
}

int Writer::close()
{
    if (!fp_) return 0;
    FILE* f = (FILE*)fp_;
    int rc = 0;
    if (defined_ && !vars_.empty()) {
        // make the file as long as its last variable even when that variable was never put
        const Var& v = vars_.back();
        const uint64_t end = v.begin + pad4(v.nelems * type_size(v.type));
        if (fseeko(f, 0, SEEK_END) == 0 && (uint64_t)ftello(f) < end) {
            fseeko(f, (off_t)end - 1, SEEK_SET);
            fputc(0, f);
        }
    }
    if (fclose(f) != 0) { err_ = "close failed"; rc = -1; }
    fp_ = nullptr;
    return rc;
}

// ---------------------------------------------------------------------------------- reader
namespace {

struct Cursor {
    FILE* f;
    bool ok = true;
    uint32_t u32()
    {
        unsigned char b[4];
        if (fread(b, 1, 4, f) != 4) { ok = false; return 0; }
        return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    }
    uint64_t u64() { uint64_t hi = u32(); return (hi << 32) | u32(); }
    std::string name()
    {
        const uint32_t n = u32();
        if (!ok || n > (1u << 20)) { ok = false; return ""; }
        std::string s(n, '\0');
        if (n && fread(&s[0], 1, n, f) != n) ok = false;
        const uint32_t pad = (4 - (n & 3)) & 3;
        if (pad) fseeko(f, pad, SEEK_CUR);
        return s;
    }
    bool att_list(std::vector<Att>& out)
    {
        const uint32_t tag = u32(), n = u32();
        if (!ok) return false;
        if (tag == 0 && n == 0) return true;
        if (tag != TAG_ATTRIBUTE) return false;
        for (uint32_t i = 0; i < n && ok; i++) {
            Att a;
            a.name = name();
            a.type = (int)u32();
            const uint32_t ne = u32();
            const size_t es = type_size(a.type);
            if (!ok || es == 0 || (uint64_t)ne * es > (1u << 26)) return false;
            std::vector<unsigned char> be((size_t)ne * es);
            if (!be.empty() && fread(be.data(), 1, be.size(), f) != be.size()) return false;
            a.raw.resize(be.size());
            to_be(a.raw.data(), be.data(), ne, es);      // the swap is its own inverse
            const size_t pad = (4 - (be.size() & 3)) & 3;
            if (pad) fseeko(f, (off_t)pad, SEEK_CUR);
            out.push_back(a);
        }
        return ok;
    }
};

}  // namespace

Reader::~Reader() { close(); }

#ifdef OCTANE_HAVE_NETCDF
}  // namespace cdf
#include <netcdf.h>
namespace cdf {
bool has_netcdf4() { return true; }

// NetCDF-4 input through the netCDF C library (what the reference's netcdf-cxx4 calls end in,
// src/oct_fileread.cc:71-263): the header is mirrored into the same Dim / Var / Att records the classic parser
// fills, so the callers do not know which container they read.  Unsigned and 64-bit attribute types are kept as the
// next wider classic type; string-typed attributes and user-defined types are skipped (the OCTANE readers use neither).
static int classic_type(int t)
{
    switch (t) {
    case NC_BYTE: case NC_UBYTE: return BYTE;
    case NC_CHAR: return CHAR;
    case NC_SHORT: return SHORT;
    case NC_USHORT: case NC_INT: return INT;
    case NC_FLOAT: return FLOAT;
    case NC_UINT: case NC_INT64: case NC_UINT64: case NC_DOUBLE: return DOUBLE;
    default: return 0;
    }
}

static bool read_atts(int ncid, int varid, int natts, std::vector<Att>& out)
{
    for (int k = 0; k < natts; k++) {
        char name[NC_MAX_NAME + 1];
        nc_type t;
        size_t len = 0;
        if (nc_inq_attname(ncid, varid, k, name) != NC_NOERR || nc_inq_att(ncid, varid, name, &t, &len) != NC_NOERR) return false;
        Att a;
        a.name = name;
        a.type = classic_type(t);
        if (a.type == 0) continue;
        a.raw.resize(len * type_size(a.type));
        int rc = NC_NOERR;
        if (len > 0) switch (a.type) {
        case CHAR:   rc = nc_get_att_text(ncid, varid, name, (char*)a.raw.data()); break;
        case BYTE:   rc = nc_get_att_schar(ncid, varid, name, (signed char*)a.raw.data()); break;
        case SHORT:  rc = nc_get_att_short(ncid, varid, name, (short*)a.raw.data()); break;
        case INT:    rc = nc_get_att_int(ncid, varid, name, (int*)a.raw.data()); break;
        case FLOAT:  rc = nc_get_att_float(ncid, varid, name, (float*)a.raw.data()); break;
        case DOUBLE: rc = nc_get_att_double(ncid, varid, name, (double*)a.raw.data()); break;
        }
        if (rc != NC_NOERR && rc != NC_ERANGE) return false;
        out.push_back(a);
    }
    return true;
}

int Reader::open_nc4(const std::string& path)
{
    int ncid = -1, rc = nc_open(path.c_str(), NC_NOWRITE, &ncid);
    if (rc != NC_NOERR) { err_ = path + ": " + nc_strerror(rc); return -1; }
    ncid_ = ncid;
    int ndims = 0, nvars = 0, ngatts = 0, unlim = -1;
    if (nc_inq(ncid, &ndims, &nvars, &ngatts, &unlim) != NC_NOERR) { err_ = path + ": nc_inq failed"; return -1; }
    for (int d = 0; d < ndims; d++) {
        char name[NC_MAX_NAME + 1];
        size_t len = 0;
        if (nc_inq_dim(ncid, d, name, &len) != NC_NOERR) { err_ = path + ": nc_inq_dim failed"; return -1; }
        Dim dm; dm.name = name; dm.len = len;
        dims_.push_back(dm);
    }
    if (!read_atts(ncid, NC_GLOBAL, ngatts, gatts_)) { err_ = path + ": cannot read the global attributes"; return -1; }
    for (int k = 0; k < nvars; k++) {
        char name[NC_MAX_NAME + 1];
        nc_type t;
        int nd = 0, natts = 0, dimids[NC_MAX_VAR_DIMS];
        if (nc_inq_var(ncid, k, name, &t, &nd, dimids, &natts) != NC_NOERR) { err_ = path + ": nc_inq_var failed"; return -1; }
        Var v;
        v.name = name;
        v.type = classic_type(t);
        if (v.type == 0 || v.type == CHAR) continue;      // string / user-defined / text variables: not read by OCTANE
        v.ncvarid = k;
        v.nelems = 1;
        for (int i = 0; i < nd; i++) {
            if (dimids[i] < 0 || dimids[i] >= (int)dims_.size()) { err_ = path + ": bad dimension id"; return -1; }
            v.dimids.push_back(dimids[i]);
            const uint64_t len = dims_[dimids[i]].len;
            if (len != 0 && v.nelems > UINT64_MAX / len) { err_ = path + ": variable " + v.name + " is too large"; return -1; }
            v.nelems *= len;
        }
        if (!read_atts(ncid, k, natts, v.atts)) { err_ = path + ": cannot read the attributes of " + v.name; return -1; }
        vars_.push_back(v);
    }
    return 0;
}

static int nc_get(int ncid, int varid, short* out) { return nc_get_var_short(ncid, varid, out); }
static int nc_get(int ncid, int varid, int* out) { return nc_get_var_int(ncid, varid, out); }
static int nc_get(int ncid, int varid, float* out) { return nc_get_var_float(ncid, varid, out); }
static int nc_get(int ncid, int varid, double* out) { return nc_get_var_double(ncid, varid, out); }
#define OCTANE_NC_GET(ncid, v, out)                                                              \
    do {                                                                                         \
        const int rc_ = nc_get(ncid, (v)->ncvarid, out);   /* NC_ERANGE: a value outside T, as a C cast would give */ \
        if (rc_ != NC_NOERR && rc_ != NC_ERANGE) { err_ = "read of " + (v)->name + ": " + nc_strerror(rc_); return -1; } \
        return 0;                                                                                \
    } while (0)
static void nc_close_id(int ncid) { nc_close(ncid); }
#else
bool has_netcdf4() { return false; }
int Reader::open_nc4(const std::string& path) { err_ = path + ": built without the netCDF library"; return -1; }
static void nc_close_id(int) {}
#endif

void Reader::close()
{
    if (fp_) fclose((FILE*)fp_);
    fp_ = nullptr;
    if (ncid_ >= 0) nc_close_id(ncid_);
    ncid_ = -1;
}

int Reader::open(const std::string& path)
{
    close();
    dims_.clear(); vars_.clear(); gatts_.clear();
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err_ = "cannot open " + path; return -1; }
    fp_ = f;
    unsigned char magic[4];
    if (fread(magic, 1, 4, f) != 4) { err_ = path + ": empty file"; return -1; }
    if (magic[0] == 0x89 && magic[1] == 'H' && magic[2] == 'D' && magic[3] == 'F') {
        if (has_netcdf4()) { close(); return open_nc4(path); }
        err_ = path + ": NetCDF-4/HDF5 container; this build reads the classic format only (convert with `nccopy -k cdf2`, "
                      "or rebuild with the netCDF C library present)";
        return -1;
    }
    if (magic[0] != 'C' || magic[1] != 'D' || magic[2] != 'F' || (magic[3] != 1 && magic[3] != 2)) {
        err_ = path + ": not a classic NetCDF (CDF-1/CDF-2) file";
        return -1;
    }
    version_ = magic[3];
    Cursor c{ f };
    c.u32();                                            // numrecs (record variables unsupported)
    uint32_t tag = c.u32(), n = c.u32();
    if (tag == TAG_DIMENSION) {
        for (uint32_t i = 0; i < n && c.ok; i++) { Dim d; d.name = c.name(); d.len = c.u32(); dims_.push_back(d); }
    } else if (!(tag == 0 && n == 0)) { err_ = path + ": malformed dimension list"; return -1; }
    if (!c.att_list(gatts_)) { err_ = path + ": malformed global attributes"; return -1; }
    tag = c.u32(); n = c.u32();
    if (tag == TAG_VARIABLE) {
        for (uint32_t i = 0; i < n && c.ok; i++) {
            Var v;
            v.name = c.name();
            const uint32_t nd = c.u32();
            if (nd > 64) { c.ok = false; break; }
            v.nelems = 1;
            for (uint32_t k = 0; k < nd; k++) {
                const uint32_t id = c.u32();
                if (id >= dims_.size()) { c.ok = false; break; }
                v.dimids.push_back((int)id);
                if (dims_[id].len != 0 && v.nelems > UINT64_MAX / dims_[id].len) { c.ok = false; break; }
                v.nelems *= dims_[id].len;
            }
            if (!c.ok || !c.att_list(v.atts)) { c.ok = false; break; }
            v.type = (int)c.u32();
            c.u32();                                    // vsize (recomputed from the dimensions)
            v.begin = (version_ == 2) ? c.u64() : c.u32();
            if (type_size(v.type) == 0) { c.ok = false; break; }
            vars_.push_back(v);
        }
    } else if (!(tag == 0 && n == 0)) { err_ = path + ": malformed variable list"; return -1; }
    if (!c.ok) { err_ = path + ": truncated or malformed header"; return -1; }
    // every variable must lie inside the file: callers size their buffers from the header, and a corrupt
    // dimension would otherwise turn into an absurd allocation long before the short read is noticed
    if (fseeko(f, 0, SEEK_END) != 0) { err_ = path + ": seek failed"; return -1; }
    const uint64_t fsize = (uint64_t)ftello(f);
    for (auto& v : vars_) {
        const uint64_t es = type_size(v.type);
        if (v.begin > fsize || v.nelems > (fsize - v.begin) / es) {
            err_ = path + ": variable " + v.name + " extends beyond the end of the file (truncated or corrupt)";
            return -1;
        }
    }
    return 0;
}

int Reader::dim_len(const std::string& name, uint64_t* len) const
{
    for (auto& d : dims_) if (d.name == name) { *len = d.len; return 0; }
    return -1;
}

const Var* Reader::var(const std::string& name) const
{
    for (auto& v : vars_) if (v.name == name) return &v;
    return nullptr;
}

template <class T> int Reader::get_as(const Var* v, T* out)
{
#ifdef OCTANE_HAVE_NETCDF
    if (ncid_ >= 0 && v && v->ncvarid >= 0) OCTANE_NC_GET(ncid_, v, out);
#endif
    if (!fp_ || !v) { err_ = "get: no such variable"; return -1; }
    FILE* f = (FILE*)fp_;
    if (fseeko(f, (off_t)v->begin, SEEK_SET) != 0) { err_ = "seek failed"; return -1; }
    const size_t es = type_size(v->type);
    std::vector<unsigned char> be(std::min<uint64_t>(v->nelems, CHUNK) * es), he(be.size());
    for (uint64_t done = 0; done < v->nelems;) {
        const size_t m = (size_t)std::min<uint64_t>(CHUNK, v->nelems - done);
        if (fread(be.data(), es, m, f) != m) { err_ = "short read (" + v->name + ")"; return -1; }
        to_be(he.data(), be.data(), m, es);
        const unsigned char* p = he.data();
        switch (v->type) {
            case BYTE: for (size_t i = 0; i < m; i++) out[done + i] = (T)((const signed char*)p)[i]; break;
            case CHAR: for (size_t i = 0; i < m; i++) out[done + i] = (T)p[i]; break;
            case SHORT: for (size_t i = 0; i < m; i++) { short x; memcpy(&x, p + 2 * i, 2); out[done + i] = (T)x; } break;
            case INT: for (size_t i = 0; i < m; i++) { int x; memcpy(&x, p + 4 * i, 4); out[done + i] = (T)x; } break;
            case FLOAT: for (size_t i = 0; i < m; i++) { float x; memcpy(&x, p + 4 * i, 4); out[done + i] = (T)x; } break;
            case DOUBLE: for (size_t i = 0; i < m; i++) { double x; memcpy(&x, p + 8 * i, 8); out[done + i] = (T)x; } break;
        }
        done += m;
    }
    return 0;
}

int Reader::get_short(const Var* v, short* out) { return get_as<short>(v, out); }
int Reader::get_float(const Var* v, float* out) { return get_as<float>(v, out); }
int Reader::get_double(const Var* v, double* out) { return get_as<double>(v, out); }
int Reader::get_int(const Var* v, int* out) { return get_as<int>(v, out); }

}  // namespace cdf
