// tests/stubs/nc4_reader_check.cc -- drives cdf::Reader's netCDF-library backend against tests/stubs/fake_netcdf.c:
// a file with the HDF5 signature must be routed to the library and come back through the same Dim / Var / Att records
// the classic parser fills.  Prints OK or the first difference.  Test infrastructure only.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../octane_b200/csrc/cdf.h"

extern "C" int fake_netcdf_open_count(void);

#define CHECK(c) do { if (!(c)) { printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "wb");
    const unsigned char sig[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n' };
    fwrite(sig, 1, 8, f);
    fclose(f);
    CHECK(cdf::has_netcdf4());
    {
        cdf::Reader r;
        CHECK(r.open(argv[1]) == 0);
        CHECK(fake_netcdf_open_count() == 1);
        uint64_t n = 0;
        CHECK(r.dim_len("y", &n) == 0 && n == 6);
        CHECK(r.dim_len("x", &n) == 0 && n == 8);
        CHECK(r.var("algorithm_container") == nullptr);               // string variable: skipped
        const cdf::Var* rad12 = r.var("Rad");
        CHECK(rad12 && rad12->type == cdf::SHORT && rad12->nelems == 48);
        const cdf::Var* rad = r.var("RadU");
        CHECK(rad && rad->nelems == 48 && rad->dimids.size() == 2 && rad->type == cdf::INT);   // ushort -> the next wider type
        CHECK(rad->att("scale_factor") && fabs(rad->att("scale_factor")->as_double() - 0.8121064) < 1e-6);
        CHECK(rad->att("_FillValue") && rad->att("_FillValue")->as_double() == 4095.0);
        CHECK(rad->att("_Unsigned") && rad->att("_Unsigned")->as_text() == "true");
        CHECK(rad->att("long_name") == nullptr);                        // string attribute: skipped
        CHECK(rad->att("valid_count") && rad->att("valid_count")->as_double() == 1234567890123.0);
        std::vector<float> vf(48);
        CHECK(r.get_float(rad, vf.data()) == 0 && vf[0] == 0.f && vf[9] == 200.f && vf[47] == 40007.f);
        std::vector<int> vi(48);
        CHECK(r.get_int(rad, vi.data()) == 0 && vi[40] == 40000);
        std::vector<short> vs(48);
        CHECK(r.get_short(rad, vs.data()) == 0 && vs[24] == 4095);      // out-of-range elements elsewhere: not an error
        const cdf::Var* x = r.var("x");
        CHECK(x && x->nelems == 8 && x->type == cdf::SHORT && fabs(x->att("add_offset")->as_double() + 0.101332) < 1e-7);
        short xs[8];
        CHECK(r.get_short(x, xs) == 0 && xs[7] == 7);
        const cdf::Var* t = r.var("t");
        double tv = 0;
        CHECK(t && t->nelems == 1 && r.get_double(t, &tv) == 0 && tv == 7.123456789e8);
        CHECK(t->att("units") && t->att("units")->as_text() == "seconds since 2000-01-01 12:00:00");
        const cdf::Var* b = r.var("band_id");
        int bv = 0;
        CHECK(b && r.get_int(b, &bv) == 0 && bv == 13);
        const cdf::Var* gip = r.var("goes_imager_projection");
        CHECK(gip && gip->att("perspective_point_height")->as_double() == 35786023.0 &&
              gip->att("longitude_of_projection_origin")->as_double() == -75.0);
        const cdf::Var* dq = r.var("DQF");
        std::vector<int> dv(48);
        CHECK(dq && dq->type == cdf::BYTE && r.get_int(dq, dv.data()) == 0 && dv[4] == 255);
        r.close();
        CHECK(fake_netcdf_open_count() == 0);
        CHECK(r.open(argv[1]) == 0);                                    // reopen; the destructor closes it
    }
    CHECK(fake_netcdf_open_count() == 0);
    printf("OK\n");
    return 0;
}
