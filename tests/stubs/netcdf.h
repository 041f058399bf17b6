/* tests/stubs/netcdf.h -- the part of the netCDF C library's public interface that octane_b200/csrc/cdf.cc uses,
 * declared from the library's documentation so that the OCTANE_HAVE_NETCDF code path can be compiled and exercised
 * in an image without the library (tests/test_cli.py::test_netcdf4_backend_against_a_stub_library links it against
 * tests/stubs/fake_netcdf.c, a table-driven stand-in).  Test infrastructure only. */
#ifndef OCTANE_TEST_STUB_NETCDF_H
#define OCTANE_TEST_STUB_NETCDF_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int nc_type;
#define NC_NAT 0
#define NC_BYTE 1
#define NC_CHAR 2
#define NC_SHORT 3
#define NC_INT 4
#define NC_FLOAT 5
#define NC_DOUBLE 6
#define NC_UBYTE 7
#define NC_USHORT 8
#define NC_UINT 9
#define NC_INT64 10
#define NC_UINT64 11
#define NC_STRING 12
#define NC_NOWRITE 0
#define NC_GLOBAL (-1)
#define NC_MAX_NAME 256
#define NC_MAX_VAR_DIMS 1024
#define NC_NOERR 0
#define NC_ERANGE (-60)
int nc_open(const char* path, int mode, int* ncidp);
int nc_close(int ncid);
int nc_inq(int ncid, int* ndimsp, int* nvarsp, int* nattsp, int* unlimdimidp);
int nc_inq_dim(int ncid, int dimid, char* name, size_t* lenp);
int nc_inq_var(int ncid, int varid, char* name, nc_type* xtypep, int* ndimsp, int* dimidsp, int* nattsp);
int nc_inq_attname(int ncid, int varid, int attnum, char* name);
int nc_inq_att(int ncid, int varid, const char* name, nc_type* xtypep, size_t* lenp);
int nc_get_att_text(int ncid, int varid, const char* name, char* ip);
int nc_get_att_schar(int ncid, int varid, const char* name, signed char* ip);
int nc_get_att_short(int ncid, int varid, const char* name, short* ip);
int nc_get_att_int(int ncid, int varid, const char* name, int* ip);
int nc_get_att_float(int ncid, int varid, const char* name, float* ip);
int nc_get_att_double(int ncid, int varid, const char* name, double* ip);
int nc_get_var_short(int ncid, int varid, short* ip);
int nc_get_var_int(int ncid, int varid, int* ip);
int nc_get_var_float(int ncid, int varid, float* ip);
int nc_get_var_double(int ncid, int varid, double* ip);
const char* nc_strerror(int ncerr);
#ifdef __cplusplus
}
#endif
#endif
