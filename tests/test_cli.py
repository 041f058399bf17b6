"""The `octane` host program (octane_b200/bin/octane): flag surface of reference src/main.cc, the
self-contained classic-NetCDF reader/writer (octane_b200/csrc/cdf.cc) and, on a GPU, the whole
path file -> ingest -> flow -> navigation -> outfile.nc against the library called directly."""
import os
import subprocess

import numpy as np
import pytest

import cases
import goes_files as G
import octane_b200 as ob
from octane_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "octane_b200", "bin", "octane")


def run(*args, check=True):
    r = subprocess.run([EXE, *args], capture_output=True, text=True, timeout=600)
    if check:
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


def settings(*args):
    out = run(*args, "-dump_settings").stdout
    return dict(ln.split("=", 1) for ln in out.splitlines() if "=" in ln)


def test_usage_and_defaults():
    r = run()                                   # fewer than 3 arguments: usage text, exit 0 (src/main.cc:112-164)
    assert "-i1 <filename>" in r.stdout and "-alpha" in r.stdout
    s = settings()
    # src/main.cc:53-108
    assert (float(s["alpha"]), float(s["lambda"]), float(s["lambdac"]), float(s["scaleF"])) == (5.0, 1.0, 0.0, 0.5)
    assert (int(s["kiters"]), int(s["liters"]), int(s["cgiters"]), int(s["miters"])) == (4, 3, 30, 5)
    assert (int(s["dozim"]), int(s["oftype"]), int(s["pixuv"]), int(s["setdevice"])) == (1, 1, 0, 0)
    assert all(int(s[k]) == 1 for k in ("outnav", "outraw", "outrad", "outctp", "interpcth"))


def test_flags_and_reference_quirks():
    s = settings("-i1", "a.nc", "-i2", "b.nc", "-alpha", "3.5", "-lambda", "0.25", "-lambdac", "0.1", "-kiters", "5", "-liters",
                 "2", "-pd", "-brox", "-ir", "-i1cth", "c.nc", "-firstguess", "fg.nc", "-o", "/tmp/x/", "-set_device", "3",
                 "-no_outraw", "-no_outctp", "-nncth", "-scsig", "7", "-cgiters", "99", "-corn")
    assert (s["i1"], s["i2"], s["i1cth"], s["firstguess"], s["o"]) == ("a.nc", "b.nc", "c.nc", "fg.nc", "/tmp/x/")
    assert (float(s["alpha"]), float(s["lambda"]), float(s["lambdac"])) == (3.5, 0.25, 0.1)
    assert (int(s["kiters"]), int(s["liters"]), int(s["pixuv"]), int(s["ir"]), int(s["doCTH"])) == (5, 2, 1, 1, 1)
    assert int(s["dozim"]) == 0 and int(s["oftype"]) == 3          # -brox: :265-268, :365-371
    assert int(s["setdevice"]) == 2                                 # 1-based on the command line, :311-314
    assert (int(s["outraw"]), int(s["outctp"]), int(s["outnav"]), int(s["interpcth"])) == (0, 0, 1, 0)
    assert float(s["scsig"]) == 49.0                                # stores the square, :229
    assert int(s["cgiters"]) == 30                                  # documented (:144) but never parsed
    assert int(s["docorn"]) == 0                                    # -corn is a no-op, :270-273
    assert int(settings("-Polar", "-i1cth", "c.nc")["doCTH"]) == 0  # :380-391
    r = run("-i1", "a", "-i2", "b", "-farn")
    assert "Farneback disabled" in r.stdout


def _pair_files(tmp, nx=96, ny=80, band=2, cth=False, sector="meso_0.5km", seed=21, prefix=""):
    c = dict(sector=sector, nx=nx, ny=ny, x0=10, y0=20)
    xs, ys, xo, yo, dt = S.SECTORS[sector]
    i1, i2, _, _ = S.make_pair(nx, ny, seed)
    maxin, minin = ob.band_minmax(band)
    radScale, radOffset = (maxin - minin) / 4000.0, minin
    xc = (np.arange(nx) + c["x0"]).astype(np.int16); yc = (np.arange(ny) + c["y0"]).astype(np.int16)
    f1, f2 = os.path.join(tmp, prefix + "g1.nc"), os.path.join(tmp, prefix + "g2.nc")
    r1 = G.counts_from_image(i1, maxin, minin, radScale, radOffset); r2 = G.counts_from_image(i2, maxin, minin, radScale, radOffset)
    G.write_goes_l1b(f1, r1, xc, yc, 1000.0, band, xs, xo, ys, yo, radScale, radOffset)
    G.write_goes_l1b(f2, r2, xc, yc, 1000.0 + dt, band, xs, xo, ys, yo, radScale, radOffset)
    return dict(f1=f1, f2=f2, r1=r1, r2=r2, xc=xc, yc=yc, dt=dt, xs=xs, ys=ys, xo=xo, yo=yo, radScale=radScale,
                radOffset=radOffset, band=band, nx=nx, ny=ny)


def test_netcdf_layout_matches_the_reference_writer(tmp_path):
    """dry run (no GPU): read two GOES-shaped files, write outfile.nc; scipy must read it back and find
    the reference's schema (src/oct_filewrite.cc:17-349): names, order, types, attributes."""
    from scipy.io import netcdf_file
    tmp = str(tmp_path)
    d = _pair_files(tmp)
    cth = (5000 + 4000 * np.sin(np.arange(d["nx"])[None, :] / 9.0) * np.ones((d["ny"], 1))).astype(np.float32)
    G.write_plane_file(os.path.join(tmp, "cth.nc"), Cloud_Top_Height_Effective=cth)
    run("-i1", d["f1"], "-i2", d["f2"], "-i1cth", os.path.join(tmp, "cth.nc"), "-o", tmp + "/", "-pd", "-dry_run")
    f = netcdf_file(os.path.join(tmp, "outfile.nc"), "r", mmap=False)
    assert list(f.dimensions.items()) == [("x", d["nx"]), ("y", d["ny"])]
    want = ["x", "y", "t", "U", "V", "U_raw", "V_raw", "Upix", "Vpix", "CTP", "Rad", "goes_imager_projection",
            "optical_flow_settings", "planck_fk1", "planck_fk2", "planck_bc1", "planck_bc2", "kappa0"]
    assert list(f.variables) == want
    v = f.variables
    assert v["U"].dimensions == ("y", "x") and v["U"].data.dtype == np.dtype(">i2") and v["Upix"].data.dtype == np.dtype(">f4")
    assert v["t"].data.dtype == np.dtype(">f8") and float(v["t"].getValue()) == 1000.0
    assert v["U"].long_name == b"U" and v["U"].grid_mapping == b"goes_imager_projection" and v["U"].units == b"x-pixels"
    assert v["V"].units == b"y-pixels" and abs(float(v["U"].scale_factor) - 0.01) < 1e-9
    assert v["U_raw"].long_name == b"U Raw" and v["V_raw"].units == b"y-pixels"
    assert np.array_equal(v["x"][:], d["xc"]) and np.array_equal(v["y"][:], d["yc"]) and np.array_equal(v["Rad"][:], d["r1"])
    assert float(v["x"].scale_factor) == np.float32(d["xs"]) and float(v["y"].add_offset) == np.float32(d["yo"])
    assert np.array_equal(v["CTP"][:], cth.astype(np.int16)) and float(v["CTP"].interpcth) == 1.0
    g = v["goes_imager_projection"]
    assert g.grid_mapping_name == b"geostationary" and g.sweep_angle_axis == b"x"
    assert float(g.perspective_point_height) == float(np.float32(35786023.0)) and float(g.longitude_of_projection_origin) == -75.0
    o = v["optical_flow_settings"]
    assert int(o.getValue()) == 1 and int(o.K_Iterations) == 4 and int(o.L_Iterations) == 3 and int(o.CG_Iterations) == 30
    assert float(o.alpha) == 5.0 and float(o.getattr("lambda") if hasattr(o, "getattr") else o.__dict__["_attributes"]["lambda"]) == 1.0
    mx, mn = ob.band_minmax(2)
    assert abs(float(o.NormMax) - mx) < 1e-4 and abs(float(o.NormMin) - mn) < 1e-5 and float(o.dt_seconds) == d["dt"]
    assert float(o.Image2_xOffset) == np.float32(d["xo"])
    assert abs(float(v["planck_fk1"].getValue()) - 202263.0) < 1e-2
    f.close()
    # without -pd / -i1cth the optional variables are absent; -no_out* drop their groups
    run("-i1", d["f1"], "-i2", d["f2"], "-o", tmp + "/", "-no_outraw", "-no_outrad", "-dry_run")
    f = netcdf_file(os.path.join(tmp, "outfile.nc"), "r", mmap=False)
    assert list(f.variables) == ["x", "y", "t", "U", "V", "goes_imager_projection", "optical_flow_settings"]
    assert f.variables["U"].units == b"meters per second"
    f.close()


def test_reader_rejects_what_it_cannot_read(tmp_path):
    tmp = str(tmp_path)
    bad = os.path.join(tmp, "hdf5.nc")
    open(bad, "wb").write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    r = run("-i1", bad, "-i2", bad, "-dry_run", check=False)
    assert r.returncode == 1 and "classic format only" in r.stderr
    r = run("-i1", os.path.join(tmp, "missing.nc"), "-i2", bad, "-dry_run", check=False)
    assert r.returncode == 1 and "cannot open" in r.stderr
    d = _pair_files(tmp)
    G.write_plane_file(os.path.join(tmp, "cth_small.nc"), Cloud_Top_Height_Effective=np.zeros((10, 12), np.float32))
    r = run("-i1", d["f1"], "-i2", d["f2"], "-i1cth", os.path.join(tmp, "cth_small.nc"), "-dry_run", check=False)
    assert r.returncode == 1 and "regridding" in r.stderr
    r = run("-i1", d["f1"], "-i2", d["f2"], "-srsal", "-dry_run", check=False)       # the smoother's range weight needs heights
    assert r.returncode == 1 and "-srsal needs cloud-top heights" in r.stderr


def test_reader_survives_corrupt_headers(tmp_path):
    """truncated files and headers with absurd counts end in an error message and exit code 1, never in an
    abort: the reader checks every variable's extent against the file size before anybody allocates"""
    import random
    tmp = str(tmp_path)
    d = _pair_files(tmp, nx=48, ny=40)
    good = open(d["f1"], "rb").read()
    rng = random.Random(5)
    bad = os.path.join(tmp, "bad.nc")
    seen = set()
    for it in range(60):
        b = bytearray(good)
        mode = it % 4
        if mode == 0:
            b = b[:rng.randrange(4, len(b))]
        else:
            k = rng.randrange(0, 1400 // 4) * 4
            b[k:k + 4] = {1: bytes([0x7f, 0xff, 0xff, rng.randrange(256)]), 2: b"\xff\xff\xff\xff",
                          3: bytes(rng.randrange(256) for _ in range(4))}[mode]
        open(bad, "wb").write(bytes(b))
        r = run("-i1", bad, "-i2", d["f2"], "-o", tmp + "/", "-dry_run", check=False)
        assert r.returncode in (0, 1), (it, mode, r.returncode, r.stderr[-300:])
        seen.add(r.returncode)
    assert 1 in seen
    open(bad, "wb").write(good[:len(good) - 100])                       # the last variable is cut short
    r = run("-i1", bad, "-i2", d["f2"], "-o", tmp + "/", "-dry_run", check=False)
    assert r.returncode == 1 and "extends beyond the end of the file" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "pd_cth", "firstguess", "cth_coarse_nn", "two_channels", "cth_fine_srsal",
                                  "channel2_fine"])
def test_cli_end_to_end_matches_the_library(tmp_path, ctx, mode):
    """file -> octane -> outfile.nc equals ingest + flow + navigation called through the C ABI"""
    from scipy.io import netcdf_file
    tmp = str(tmp_path)
    d = _pair_files(tmp, nx=160, ny=128)
    nx, ny = d["nx"], d["ny"]
    args = ["-i1", d["f1"], "-i2", d["f2"], "-o", tmp + "/"]
    p = ob.default_params()
    cth = None
    u0 = np.zeros((ny, nx), np.float32); v0 = np.zeros((ny, nx), np.float32)
    nav = ob.goes_nav(d["xs"], d["ys"], d["xo"], d["yo"], pph=float(np.float32(35786023.0)), req=float(np.float32(6378137.0)),
                      rpol=float(np.float32(6356752.31414)))
    nav.lam0 = float(np.float32(np.float32(-75.0) * (3.14159265359 / 180.)))
    cal = ob.goes_cal(d["radScale"], d["radOffset"], band=d["band"], fk1=202263.0, fk2=3698.19, bc1=0.43361, bc2=0.99939,
                      kap1=0.0019486)
    img1, lat, lon = ctx.oct_navcal_cuda(d["r1"], d["xc"], d["yc"], nav, cal)
    cal.donav = 0
    img2, _, _ = ctx.oct_navcal_cuda(d["r2"], d["xc"], d["yc"], nav, cal)
    if mode == "pd_cth":
        cth = (6000 + 5000 * np.cos(np.arange(nx)[None, :] / 11.0) * np.ones((ny, 1))).astype(np.float32)
        G.write_plane_file(os.path.join(tmp, "cth.nc"), Cloud_Top_Height_Effective=cth)
        args += ["-pd", "-i1cth", os.path.join(tmp, "cth.nc"), "-alpha", "8", "-kiters", "3"]
        p = ob.default_params(pixuv=1, doCTH=1, alpha=8.0, kiters=3)
    nc = 1
    if mode == "cth_coarse_nn":       # CTH on a 4x coarser grid, nearest-neighbour remap (-nncth)
        small = (6000 + 5000 * np.cos(np.arange(nx // 4)[None, :] / 5.0) * np.sin(np.arange(ny // 4)[:, None] / 7.0)).astype(np.float32)
        G.write_plane_file(os.path.join(tmp, "cth4.nc"), Cloud_Top_Height_Effective=small)
        args += ["-i1cth", os.path.join(tmp, "cth4.nc"), "-nncth", "-ir"]
        p = ob.default_params(doCTH=1, ir=1)
        cth = ctx.oct_zoom_in_float(small, nx, ny, 0)
    if mode == "cth_fine_srsal":      # CTH on a 2x finer grid (oct_zoom_out_float), -srsal smoothing of Upix / Vpix
        yy, xx = np.mgrid[0:2 * ny, 0:2 * nx].astype(np.float32)
        fine = (6000 + 50 * np.cos(xx / 23.0) * np.sin(yy / 17.0) + 3000 * (xx > nx)).astype(np.float32)
        G.write_plane_file(os.path.join(tmp, "cth_fine.nc"), Cloud_Top_Height_Effective=fine)
        args += ["-pd", "-i1cth", os.path.join(tmp, "cth_fine.nc"), "-srsal"]
        p = ob.default_params(pixuv=1, doCTH=1, dosrsal=1)
        cth = ctx.oct_zoom_out_float(fine, 0.5)
        assert cth.shape == (ny, nx)
    if mode == "channel2_fine":       # -ic21 / -ic22: a second channel on a 2x finer grid, blurred and decimated
        d2 = _pair_files(os.path.join(tmp), nx=320, ny=256, band=2, seed=34, prefix="c2f_")
        args += ["-ic21", d2["f1"], "-ic22", d2["f2"]]
        cal2 = ob.goes_cal(d2["radScale"], d2["radOffset"], band=2, fk1=202263.0, fk2=3698.19, bc1=0.43361, bc2=0.99939,
                           kap1=0.0019486)
        a2, _, _ = ctx.oct_navcal_cuda(d2["r1"], d2["xc"], d2["yc"], nav, cal2)
        cal2.donav = 0
        b2, _, _ = ctx.oct_navcal_cuda(d2["r2"], d2["xc"], d2["yc"], nav, cal2)
        img1 = np.ascontiguousarray(np.stack([img1, ctx.oct_zoom_out_float(a2, 0.5)]))
        img2 = np.ascontiguousarray(np.stack([img2, ctx.oct_zoom_out_float(b2, 0.5)]))
        nc = 2
    if mode == "two_channels":        # -ic21 / -ic22: a second channel on the same grid
        d2 = _pair_files(os.path.join(tmp), nx=160, ny=128, band=13, seed=33, prefix="c2_")
        args += ["-ic21", d2["f1"], "-ic22", d2["f2"]]
        cal2 = ob.goes_cal(d2["radScale"], d2["radOffset"], band=13, fk1=202263.0, fk2=3698.19, bc1=0.43361, bc2=0.99939,
                           kap1=0.0019486)
        a2, _, _ = ctx.oct_navcal_cuda(d2["r1"], d2["xc"], d2["yc"], nav, cal2)
        cal2.donav = 0
        b2, _, _ = ctx.oct_navcal_cuda(d2["r2"], d2["xc"], d2["yc"], nav, cal2)
        img1 = np.ascontiguousarray(np.stack([img1, a2])); img2 = np.ascontiguousarray(np.stack([img2, b2]))
        nc = 2
    if mode == "firstguess":
        ufg, vfg = cases.uv2pix_winds(nx, ny)
        G.write_plane_file(os.path.join(tmp, "fg.nc"), UFG=ufg, VFG=vfg)
        args += ["-firstguess", os.path.join(tmp, "fg.nc"), "-lambdac", "0.5"]
        p = ob.default_params(first_guess=1, lambdac=0.5)
        u0, v0 = ufg.copy(), vfg.copy()
        ctx.oct_uv2pix(nav, 1000.0, 1000.0 + d["dt"], lat, lon, d["xc"], d["yc"], u0, v0, p)
    run(*args)
    want = ctx.oct_optical_flow(img1, img2, nav, 1000.0, 1000.0 + d["dt"], p, cth=cth, upix=u0, vpix=v0, nc=nc)
    f = netcdf_file(os.path.join(tmp, "outfile.nc"), "r", mmap=False)
    v = f.variables
    for name, key in (("U", "uVal"), ("V", "vVal"), ("U_raw", "uVal2"), ("V_raw", "vVal2")):
        assert np.array_equal(v[name][:], want[key]), name
    if mode == "pd_cth":
        assert np.array_equal(v["Upix"][:], want["uPix"]) and np.array_equal(v["Vpix"][:], want["vPix"])
        assert np.array_equal(v["CTP"][:], want["CTP"]) and int(v["optical_flow_settings"].K_Iterations) == 3
    if mode == "cth_coarse_nn":
        assert np.array_equal(v["CTP"][:], want["CTP"]) and float(v["CTP"].interpcth) == 0.0
    if mode == "cth_fine_srsal":
        assert np.array_equal(v["Upix"][:], want["uPix"]) and np.array_equal(v["Vpix"][:], want["vPix"])
        assert np.array_equal(v["CTP"][:], want["CTP"])
        raw = v["U_raw"][:].astype(np.float32) / 100.0          # the shorts come from the UNSMOOTHED flow
        assert np.abs(want["uPix"] - raw).max() > 0.02
    if mode == "channel2_fine":
        assert "Rad2" not in v                                   # written only for channels on channel 1's grid
    if mode == "two_channels":
        assert np.array_equal(v["Rad2"][:], d2["r1"]) and "planck_fk1_2" in v and "Rad3" not in v
    assert int(v["optical_flow_settings"].dofirstguess) == int(mode == "firstguess")
    assert np.abs(v["U_raw"][:].astype(int)).max() > 20        # a real flow field came out (>0.2 px)
    f.close()


def _grid_files(tmp, kind, nx=128, ny=96):
    c = cases.GRIDNAV["gridnav_polar" if kind == "polar" else "gridnav_merc"]
    i1, i2, _, _ = S.make_pair(nx, ny, 41)
    xc = np.arange(nx, dtype=np.int16); yc = np.arange(ny, dtype=np.int16)
    f1, f2 = os.path.join(tmp, kind + "1.nc"), os.path.join(tmp, kind + "2.nc")
    kw = dict(lat1=c["lat1"], lon0=c["lon0"]) if kind == "polar" else dict(lon1=c["lon0"])
    G.write_grid_file(f1, i1, xc, yc, 5000.0, c["xScale"], c["xOffset"], c["yScale"], c["yOffset"], c["R"], **kw)
    G.write_grid_file(f2, i2, xc, yc, 5000.0 + 3600.0, c["xScale"], c["xOffset"], c["yScale"], c["yOffset"], c["R"], **kw)
    return dict(f1=f1, f2=f2, i1=i1, i2=i2, xc=xc, yc=yc, c=c, nx=nx, ny=ny)


@pytest.mark.parametrize("kind", ["polar", "merc"])
def test_projected_grid_layout(tmp_path, kind):
    """-Polar / -Merc dry run: outfile_polar.nc / outfile_merc.nc with the reference's schema
    (src/oct_filewrite.cc:353-705): U, V declared double, the grid constants on *_imager_projection"""
    from scipy.io import netcdf_file
    tmp = str(tmp_path)
    d = _grid_files(tmp, kind)
    flag = "-Polar" if kind == "polar" else "-Merc"
    run("-i1", d["f1"], "-i2", d["f2"], flag, "-pd", "-o", tmp + "/", "-dry_run")
    f = netcdf_file(os.path.join(tmp, f"outfile_{kind}.nc"), "r", mmap=False)
    proj = "polar_imager_projection" if kind == "polar" else "merc_imager_projection"
    assert list(f.variables) == ["x", "y", "t", "U", "V", "Upix", "Vpix", "Rad", proj, "optical_flow_settings"]
    v = f.variables
    assert v["U"].data.dtype == np.dtype(">f8") and v["Rad"].data.dtype == np.dtype(">f4") and v["U"].units == b"x-pixels"
    assert np.array_equal(v["Rad"][:], d["i1"]) and float(v["t"].getValue()) == 5000.0
    assert float(v["optical_flow_settings"].dt_seconds) == 3600.0 and int(v["optical_flow_settings"].K_Iterations) == 4
    if kind == "polar":
        assert v[proj].grid_mapping_name == b"polar" and float(v[proj].lat1) == d["c"]["lat1"] and float(v[proj].lon0) == d["c"]["lon0"]
        assert v["U"].grid_mapping == b"polar_orthonormal" and not hasattr(v["U"], "scale_factor")
    else:
        assert v[proj].grid_mapping_name == b"Mercator" and float(v[proj].lon1) == d["c"]["lon0"]
        assert abs(float(v["U"].scale_factor) - 0.01) < 1e-9
    assert float(v[proj].R) == float(np.float32(d["c"]["R"]))
    f.close()
    r = run("-i1", d["f1"], "-i2", d["f2"], flag, "-firstguess", "x.nc", "-dry_run", check=False)
    assert r.returncode == 1 and "not part of this build" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["polar", "merc"])
def test_cli_projected_grids_end_to_end(tmp_path, ctx, kind):
    from scipy.io import netcdf_file
    tmp = str(tmp_path)
    d = _grid_files(tmp, kind)
    c = d["c"]
    run("-i1", d["f1"], "-i2", d["f2"], "-Polar" if kind == "polar" else "-Merc", "-o", tmp + "/")
    nav = ob.goes_nav(c["xScale"], c["yScale"], c["xOffset"], c["yOffset"])
    nav.R = float(np.float32(c["R"])); nav.lon0 = c["lon0"] if kind == "polar" else 0.0
    nav.lon1 = 0.0 if kind == "polar" else c["lon0"]; nav.lat1 = c["lat1"] if kind == "polar" else 0.0
    grid = 1 if kind == "polar" else 2
    img1, _, _ = ctx.oct_navcal_grid(grid, d["i1"], d["xc"], d["yc"], nav, 1)
    img2, _, _ = ctx.oct_navcal_grid(grid, d["i2"], d["xc"], d["yc"], nav, 0)
    p = ob.default_params(dopolar=int(kind == "polar"), domerc=int(kind == "merc"))
    want = ctx.oct_optical_flow(img1, img2, nav, 5000.0, 8600.0, p)
    f = netcdf_file(os.path.join(tmp, f"outfile_{kind}.nc"), "r", mmap=False)
    v = f.variables
    if kind == "polar":     # the polar file carries the pixel displacements in its double U / V
        assert np.array_equal(v["U"][:], want["uPix"].astype(np.float64)) and np.array_equal(v["V"][:], want["vPix"].astype(np.float64))
    else:
        assert np.array_equal(v["U"][:], want["uVal"].astype(np.float64)) and np.array_equal(v["V"][:], want["vVal"].astype(np.float64))
    assert np.abs(want["uPix"]).max() > 0.2
    f.close()


def test_netcdf4_backend_against_a_stub_library(tmp_path):
    """Operational GOES-R L1b files are NetCDF-4 / HDF5.  The image has no netCDF library, so the library backend of
    cdf::Reader (csrc/cdf.cc, OCTANE_HAVE_NETCDF; selected by csrc/Makefile when `nc-config` is found) is compiled against
    the declarations in tests/stubs/netcdf.h and run against tests/stubs/fake_netcdf.c, a table-driven stand-in that
    serves one GOES-shaped dataset: (1) the Reader mirrors dimensions, variables and attributes (unsigned and 64-bit
    types widened, string types skipped) and converts on read; (2) the `octane` program built that way reads two files
    that carry the HDF5 signature and writes the reference's schema (dry run, no GPU)."""
    from scipy.io import netcdf_file
    tmp = str(tmp_path)
    stubs = os.path.join(ROOT, "tests", "stubs")
    csrc = os.path.join(ROOT, "octane_b200", "csrc")
    cxx = ["g++", "-O2", "-std=c++17", "-Wall", "-D_FILE_OFFSET_BITS=64", "-I", stubs]
    subprocess.check_call(cxx + ["-DOCTANE_HAVE_NETCDF", "-c", "-o", f"{tmp}/cdf_nc.o", os.path.join(csrc, "cdf.cc")])
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", stubs, "-c", "-o", f"{tmp}/fake_nc.o", os.path.join(stubs, "fake_netcdf.c")])
    subprocess.check_call(cxx + ["-o", f"{tmp}/check", os.path.join(stubs, "nc4_reader_check.cc"), f"{tmp}/cdf_nc.o", f"{tmp}/fake_nc.o"])
    r = subprocess.run([f"{tmp}/check", f"{tmp}/probe.nc"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr
    lib = os.path.join(ROOT, "octane_b200", "lib")
    subprocess.check_call(cxx + ["-o", f"{tmp}/octane_nc4", os.path.join(ROOT, "octane_b200", "cli", "octane_main.cc"), f"{tmp}/cdf_nc.o",
                                 f"{tmp}/fake_nc.o", "-L", lib, "-loctane_b200", f"-Wl,-rpath,{lib}"])
    for name in ("file1.nc", "file2.nc"):
        open(os.path.join(tmp, name), "wb").write(b"\x89HDF\r\n\x1a\n")
    r = subprocess.run([f"{tmp}/octane_nc4", "-i1", f"{tmp}/file1.nc", "-i2", f"{tmp}/file2.nc", "-o", tmp + "/", "-dry_run"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    f = netcdf_file(os.path.join(tmp, "outfile.nc"), "r", mmap=False)
    assert list(f.dimensions.items()) == [("x", 8), ("y", 6)]
    v = f.variables
    assert np.array_equal(v["x"][:], np.arange(8)) and np.array_equal(v["y"][:], np.arange(10, 16))
    assert v["Rad"][:].shape == (6, 8) and list(v["Rad"][:][3]) == list(range(4095, 4087, -1)) and int(v["Rad"][:][5, 7]) == 2047
    assert float(v["x"].scale_factor) == np.float32(5.6e-05) and float(v["y"].add_offset) == np.float32(0.128212)
    assert float(v["t"].getValue()) == 7.123456789e8 and float(v["optical_flow_settings"].dt_seconds) == 600.0
    assert float(v["goes_imager_projection"].longitude_of_projection_origin) == -75.0
    assert abs(float(v["planck_fk1"].getValue()) - 202263.0) < 1e-2 and abs(float(v["kappa0"].getValue()) - 0.0123) < 1e-7
    f.close()
    # the shipped binary (no library in this image) still says what to do with such a file
    r = run("-i1", f"{tmp}/file1.nc", "-i2", f"{tmp}/file2.nc", "-dry_run", check=False)
    assert r.returncode == 1 and "classic format only" in r.stderr
