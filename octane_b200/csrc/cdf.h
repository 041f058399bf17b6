// cdf.h -- self-contained classic NetCDF (CDF-1 / CDF-2 "64-bit offset") reader and writer.
//
// The reference does its file I/O through netcdf-cxx4 (src/oct_fileread.cc, src/oct_filewrite.cc);
// neither that library nor HDF5 exists in this image, so the `octane` host program carries its
// own implementation of the classic on-disk format (big-endian header: dimensions, global
// attributes, variables with attributes, then the fixed-size variable data).  Only what the
// OCTANE files need: fixed dimensions (no record variables), types byte/char/short/int/float/
// double, scalar and n-d variables.
//
// Operational GOES-R L1b / CLAVR-x files are NetCDF-4 (HDF5 container).  When the build finds the netCDF C library
// (csrc/Makefile: `nc-config`, or NETCDF_CFLAGS / NETCDF_LIBS) it defines OCTANE_HAVE_NETCDF and the Reader opens such
// files through it behind the same interface (cdf.cc: open_nc4); without the library the Reader reports what to do
// (`nccopy -k cdf2`).  The Writer always produces the classic container, which every netCDF tool reads.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace cdf {

enum Type { BYTE = 1, CHAR = 2, SHORT = 3, INT = 4, FLOAT = 5, DOUBLE = 6 };
size_t type_size(int t);

struct Att {
    std::string name;
    int type = CHAR;
    std::vector<unsigned char> raw;     // host-endian values, nelems * type_size
    size_t nelems() const { return raw.size() / type_size(type); }
    static Att text(const std::string& name, const std::string& v);
    static Att f32(const std::string& name, float v);
    static Att f64(const std::string& name, double v);
    static Att i32(const std::string& name, int v);
    double as_double(size_t k = 0) const;   // numeric attributes of any type
    std::string as_text() const;
};

struct Dim { std::string name; uint64_t len = 0; };

struct Var {
    std::string name;
    int type = FLOAT;
    std::vector<int> dimids;
    std::vector<Att> atts;
    uint64_t begin = 0;                 // file offset of the data (filled by the reader / by File::enddef)
    uint64_t nelems = 1;
    int ncvarid = -1;                   // variable id in the netCDF library (OCTANE_HAVE_NETCDF builds, NetCDF-4 input)
    const Att* att(const std::string& n) const;
};

// ---- writer: define everything, enddef(), then put each variable once ----------------------
class Writer {
public:
    Writer() = default;
    ~Writer();
    int create(const std::string& path);                 // 0 ok
    int add_dim(const std::string& name, uint64_t len);  // -> dimid
    int add_var(const std::string& name, int type, const std::vector<int>& dimids);   // -> varid
    void put_att(int varid, const Att& a);               // same name replaces, as NcVar::putAtt does
    int enddef();                                        // writes the header; 0 ok
    // host-endian values in, big-endian on disk; n must equal the variable's element count
    int put_var(int varid, const void* data, uint64_t n);
    int close();
    const std::string& error() const { return err_; }

private:
    std::vector<Dim> dims_;
    std::vector<Var> vars_;
    void* fp_ = nullptr;
    bool defined_ = false;
    std::string err_;
};

// ---- reader ------------------------------------------------------------------------------
class Reader {
public:
    Reader() = default;
    ~Reader();
    int open(const std::string& path);                   // 0 ok
    int dim_len(const std::string& name, uint64_t* len) const;
    const Var* var(const std::string& name) const;
    // reads the whole variable converted to T (host-endian); out must hold var->nelems values
    int get_short(const Var* v, short* out);
    int get_float(const Var* v, float* out);
    int get_double(const Var* v, double* out);
    int get_int(const Var* v, int* out);
    const std::vector<Dim>& dims() const { return dims_; }
    const std::vector<Var>& vars() const { return vars_; }
    const std::string& error() const { return err_; }
    void close();

private:
    template <class T> int get_as(const Var* v, T* out);
    std::vector<Dim> dims_;
    std::vector<Var> vars_;
    std::vector<Att> gatts_;
    void* fp_ = nullptr;
    int version_ = 1;
    int ncid_ = -1;                     // >= 0: the file is open through the netCDF library (see open_nc4)
    int open_nc4(const std::string& path);
    std::string err_;
};

// true when this build links the netCDF C library and can therefore read NetCDF-4 / HDF5 files as well
bool has_netcdf4();

}  // namespace cdf
