"""ctypes view of the CHECKERS (test infrastructure; never imported by the
product package octane_b200):

  liboracle.so           CPU restatement, oracle/oct_oracle.c
  _ref/libref_cpu.so     reference CPU stages compiled in place
  _ref/libref_cuda.so    reference CUDA path recompiled for sm_100 (needs a GPU)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i16 = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


class OracleParams(C.Structure):
    _fields_ = [("alpha", C.c_double), ("lambda_", C.c_double), ("lambdac", C.c_double),
                ("scaleF", C.c_double), ("kiters", C.c_int), ("liters", C.c_int),
                ("cgiters", C.c_int), ("dozim", C.c_int)]


class OracleNav(C.Structure):
    _fields_ = [("pph", C.c_double), ("req", C.c_double), ("rpol", C.c_double), ("lam0", C.c_double),
                ("xScale", C.c_float), ("xOffset", C.c_float), ("yScale", C.c_float), ("yOffset", C.c_float),
                ("g2xOffset", C.c_float), ("g2yOffset", C.c_float),
                ("lat1", C.c_float), ("lon1", C.c_float), ("lon0", C.c_float), ("R", C.c_float),
                ("minX", C.c_int), ("minY", C.c_int)]


class RefParams(C.Structure):
    _fields_ = [("alpha", C.c_double), ("lambda_", C.c_double), ("lambdac", C.c_double),
                ("scaleF", C.c_double), ("scsig", C.c_double),
                ("kiters", C.c_int), ("liters", C.c_int), ("cgiters", C.c_int), ("dozim", C.c_int),
                ("setdevice", C.c_int), ("pixuv", C.c_int), ("dopolar", C.c_int), ("domerc", C.c_int),
                ("dososm", C.c_int), ("rad", C.c_int), ("srad", C.c_int), ("doCTH", C.c_int),
                ("ir", C.c_int), ("dofirstguess", C.c_int)]


RefNav = OracleNav  # same field order in oracle/ref_driver.cc


class OracleCal(C.Structure):
    """the float scalars oct_navcal_cuda receives (src/oct_navcal_cuda.cu:100-107)"""
    _fields_ = [("xScale", C.c_float), ("xOffset", C.c_float), ("yScale", C.c_float), ("yOffset", C.c_float),
                ("radScale", C.c_float), ("radOffset", C.c_float),
                ("rpol", C.c_float), ("req", C.c_float), ("H", C.c_float), ("lam0", C.c_float),
                ("fk1", C.c_float), ("fk2", C.c_float), ("bc1", C.c_float), ("bc2", C.c_float), ("kap1", C.c_float),
                ("maxin", C.c_float), ("minin", C.c_float), ("maxout", C.c_float), ("minout", C.c_float),
                ("cal", C.c_int), ("donav", C.c_int)]


RefCal = OracleCal  # same field order in oracle/ref_driver.cc


def goes_cal(nav, radScale, radOffset, maxin, minin, fk1=0.0, fk2=0.0, bc1=0.0, bc2=1.0, kap1=0.0, cal=0, donav=1):
    """as oct_goesread assembles the call (src/oct_fileread.cc:51,306,341-388): req, rpol, pph, lam0 are floats there"""
    f = np.float32
    return OracleCal(nav.xScale, nav.xOffset, nav.yScale, nav.yOffset, radScale, radOffset,
                     f(nav.rpol), f(nav.req), f(f(nav.pph) + f(nav.req)), f(nav.lam0),
                     fk1, fk2, bc1, bc2, kap1, maxin, minin, 255.0, 0.0, cal, donav)


def params(alpha=5.0, lambda_=1.0, lambdac=0.0, scaleF=0.5, kiters=4, liters=3, cgiters=30, dozim=1):
    return OracleParams(alpha, lambda_, lambdac, scaleF, kiters, liters, cgiters, dozim)


def ref_params(alpha=5.0, lambda_=1.0, lambdac=0.0, scaleF=0.5, kiters=4, liters=3, cgiters=30, dozim=1,
               pixuv=0, dopolar=0, domerc=0, dososm=0, rad=2, srad=2, doCTH=0, ir=0, dofirstguess=0):
    return RefParams(alpha, lambda_, lambdac, scaleF, 400.0, kiters, liters, cgiters, dozim, 0,
                     pixuv, dopolar, domerc, dososm, rad, srad, doCTH, ir, dofirstguess)


def goes_nav(xScale, yScale, xOffset, yOffset, pph=35786023.0, req=6378137.0, rpol=6356752.31414,
             lon0_deg=-75.0, minX=0, minY=0, g2xOffset=None, g2yOffset=None, cls=OracleNav):
    lam0 = lon0_deg * (3.14159265 / 180.0)
    return cls(pph, req, rpol, lam0, xScale, xOffset, yScale, yOffset,
               xOffset if g2xOffset is None else g2xOffset, yOffset if g2yOffset is None else g2yOffset,
               0.0, 0.0, 0.0, 6371000.0, minX, minY)


def build(ref: bool = True) -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"] + (["ref"] if ref else []))


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.oracle_variational_flow.argtypes = [_f32, _f32, C.c_int, C.c_int, C.c_int, C.POINTER(OracleParams),
                                              _f32, _f32, C.c_void_p]
        L.oracle_blur_decimate.argtypes = [_f32, C.c_int, C.c_int, C.c_int, C.c_float, _f32]
        L.oracle_gradient.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int, C.c_int]
        L.oracle_zoom_in.argtypes = [_f32, _f32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        L.oracle_zoom_size.argtypes = [C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_level_factor.argtypes = [C.c_double, C.c_int, C.c_int]
        L.oracle_level_factor.restype = C.c_float
        L.oracle_filter_radius.argtypes = [C.c_float]
        L.oracle_fill_gk.argtypes = [_f32, C.c_float, C.c_int]
        L.oracle_build.argtypes = [_f32, _f32, C.c_void_p, C.c_void_p] + [_f32] * 9 + \
            [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_float, C.c_int, C.c_int, _f32, _f32, _f32]
        L.oracle_apply.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int, _f32, _f32]
        L.oracle_pcg.argtypes = [_f32, _f32, _f32, _f32, _f32, C.c_int, C.c_int, C.c_int, C.c_float, _f32]
        L.oracle_pcg_merged.argtypes = L.oracle_pcg.argtypes       # model of the product's merged-reduction kernel
        L.oracle_set_solver.argtypes = [C.c_int]
        L.oracle_pix2uv.argtypes = [C.POINTER(OracleNav), C.c_double, C.c_double, _f32, _f32, C.c_int, C.c_int,
                                    C.c_int, _i16, _i16, _i16, _i16, C.POINTER(C.c_float)]
        L.oracle_pix2uv_ms.argtypes = [C.POINTER(OracleNav), C.c_double, C.c_double, _f32, _f32, C.c_int, C.c_int,
                                       C.c_int, _f64, _f64]
        _lib = L
    return _lib


def level_dims(nx, ny, kiters=4, scaleF=0.5):
    L = lib()
    out = []
    for k in range(kiters):
        f = L.oracle_level_factor(scaleF, kiters, k)
        a, b = C.c_int(), C.c_int()
        L.oracle_zoom_size(nx, ny, f, C.byref(a), C.byref(b))
        out.append((a.value, b.value))
    return out


def variational_flow(img1, img2, p=None, u0=None, v0=None, nc=1):
    """img*: (nc*)ny x nx float32. Returns (u, v, cg_its)."""
    p = p or params()
    img1 = np.ascontiguousarray(img1, np.float32)
    img2 = np.ascontiguousarray(img2, np.float32)
    ny, nx = img1.shape[-2:]
    u = np.zeros((ny, nx), np.float32) if u0 is None else np.array(u0, np.float32, copy=True)
    v = np.zeros((ny, nx), np.float32) if v0 is None else np.array(v0, np.float32, copy=True)
    its = np.zeros(p.kiters * 3 * p.liters, np.int32)
    lib().oracle_variational_flow(img1, img2, nx, ny, nc, C.byref(p), u, v, its.ctypes.data)
    return u, v, its


def pix2uv(nav, t1, t2, u, v, flags=0):
    u = np.ascontiguousarray(u, np.float32); v = np.ascontiguousarray(v, np.float32)
    ny, nx = u.shape
    o = [np.zeros((ny, nx), np.int16) for _ in range(4)]
    dT = C.c_float()
    rc = lib().oracle_pix2uv(C.byref(nav), t1, t2, u, v, nx, ny, flags, *o, C.byref(dT))
    return (*o, dT.value, rc)


def pix2uv_ms(nav, t1, t2, u, v, flags=0):
    u = np.ascontiguousarray(u, np.float32); v = np.ascontiguousarray(v, np.float32)
    ny, nx = u.shape
    a = np.zeros((ny, nx)); b = np.zeros((ny, nx))
    lib().oracle_pix2uv_ms(C.byref(nav), t1, t2, u, v, nx, ny, flags, a, b)
    return a, b


def navcal(rad, x, y, cal):
    rad = np.ascontiguousarray(rad, np.int16); x = np.ascontiguousarray(x, np.int16); y = np.ascontiguousarray(y, np.int16)
    ny, nx = rad.shape
    L = lib()
    L.oracle_navcal.argtypes = [_i16, _i16, _i16, C.c_int, C.c_int, C.POINTER(OracleCal), _f32, _f32, _f32]
    data = np.zeros((ny, nx), np.float32); lat = np.zeros((ny, nx), np.float32); lon = np.zeros((ny, nx), np.float32)
    L.oracle_navcal(rad, x, y, nx, ny, C.byref(cal), data, lat, lon)
    return data, lat, lon


def uv2pix(nav, t1, t2, lat, lon, x, y, u, v):
    """returns (u_pix, v_pix, rc); inputs are not modified"""
    L = lib()
    L.oracle_uv2pix.argtypes = [C.POINTER(OracleNav), C.c_double, C.c_double, _f32, _f32, _i16, _i16, C.c_int, C.c_int,
                                _f32, _f32]
    u = np.array(u, np.float32, copy=True); v = np.array(v, np.float32, copy=True)
    ny, nx = u.shape
    rc = L.oracle_uv2pix(C.byref(nav), t1, t2, np.ascontiguousarray(lat, np.float32), np.ascontiguousarray(lon, np.float32),
                         np.ascontiguousarray(x, np.int16), np.ascontiguousarray(y, np.int16), nx, ny, u, v)
    return u, v, rc


def navcal_grid(grid, data, x, y, xScale, xOffset, yScale, yOffset, R, lon0_deg, lat1_deg=0.0, donav=1):
    """grid 1 polar / 2 Mercator; angles in degrees as the readers hold them (converted like the reference's wrappers)"""
    data = np.ascontiguousarray(data, np.float32); x = np.ascontiguousarray(x, np.int16); y = np.ascontiguousarray(y, np.int16)
    ny, nx = data.shape
    L = lib()
    f = C.c_float
    L.oracle_navcal_grid.argtypes = [C.c_int, _f32, _i16, _i16, C.c_int, C.c_int, f, f, f, f, f, f, f, C.c_int, _f32, _f32, _f32]
    DTOR = 3.14159265359 / 180.
    out = np.zeros((ny, nx), np.float32); lat = np.zeros((ny, nx), np.float32); lon = np.zeros((ny, nx), np.float32)
    L.oracle_navcal_grid(grid, data, x, y, nx, ny, xScale, xOffset, yScale, yOffset, R, np.float32(np.float32(lon0_deg) * DTOR),
                         np.float32(np.float32(lat1_deg) * DTOR), donav, out, lat, lon)
    return out, lat, lon


def ref_navcal_grid(grid, data, x, y, xScale, xOffset, yScale, yOffset, R, lon0_deg, lat1_deg=0.0, donav=1, L=None):
    L = L or ref_cuda()
    data = np.ascontiguousarray(data, np.float32); x = np.ascontiguousarray(x, np.int16); y = np.ascontiguousarray(y, np.int16)
    ny, nx = data.shape
    f = C.c_float
    L.ref_navcal_grid.argtypes = [C.c_int, _f32, _i16, _i16, C.c_int, C.c_int, f, f, f, f, f, f, f, C.c_int, C.POINTER(RefParams),
                                  _f32, _f32, _f32]
    out = np.zeros((ny, nx), np.float32); lat = np.zeros((ny, nx), np.float32); lon = np.zeros((ny, nx), np.float32)
    rp = ref_params()
    bad = L.ref_navcal_grid(grid, data, x, y, nx, ny, xScale, xOffset, yScale, yOffset, R, lon0_deg, lat1_deg, donav, C.byref(rp),
                            out, lat, lon)
    assert bad == 0
    return out, lat, lon


def zoom_in_float(field, nxx, nyy, interp=1):
    field = np.ascontiguousarray(field, np.float32)
    ny, nx = field.shape
    L = lib()
    L.oracle_zoom_in_float.argtypes = [_f32, C.c_int, C.c_int, _f32, C.c_int, C.c_int, C.c_int]
    out = np.zeros((nyy, nxx), np.float32)
    L.oracle_zoom_in_float(field, nx, ny, out, nxx, nyy, interp)
    return out


def ref_zoom_in_float(field, nxx, nyy, interp=1, L=None):
    """the reference's own CPU function (oracle/_ref/libref_cpu.so; no GPU needed); L = ref_shim(): the shim's
    definition under the same signature"""
    field = np.ascontiguousarray(field, np.float32)
    ny, nx = field.shape
    L = L or ref_cpu()
    L.ref_zoom_in_float.argtypes = [_f32, _f32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    out = np.zeros((nyy, nxx), np.float32)
    L.ref_zoom_in_float(field, out, nx, ny, nxx, nyy, interp)
    return out


def zoom_out_size(nx, ny, factor):
    L = lib()
    L.oracle_zoom_out_size.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    a, b = C.c_int(), C.c_int()
    L.oracle_zoom_out_size(nx, ny, factor, C.byref(a), C.byref(b))
    return a.value, b.value


def zoom_out_float(field, factor):
    """oct_zoom_out_float restated (src/oct_zoom.cc:51-88): blur in double, bicubic sample at ii / factor"""
    field = np.ascontiguousarray(field, np.float32)
    ny, nx = field.shape
    nxx, nyy = zoom_out_size(nx, ny, factor)
    L = lib()
    L.oracle_zoom_out_float.argtypes = [_f32, C.c_int, C.c_int, _f32, C.c_double]
    out = np.zeros((nyy, nxx), np.float32)
    assert L.oracle_zoom_out_float(field, nx, ny, out, factor) == 0
    return out


def ref_zoom_out_float(field, factor, cnum=0, L=None):
    """the reference's own CPU function (oracle/_ref/libref_cpu.so); returns the raw output buffer of
    nxx*nyy + cnum floats (the plane starts at element cnum) reshaped when cnum == 0.  L = ref_shim(): the
    shim's definition under the same signature"""
    field = np.ascontiguousarray(field, np.float32)
    ny, nx = field.shape
    nxx, nyy = zoom_out_size(nx, ny, factor)
    L = L or ref_cpu()
    L.ref_zoom_out_float.argtypes = [_f32, _f32, C.c_int, C.c_int, C.c_double, C.c_int]
    out = np.zeros(nxx * nyy + cnum, np.float32)
    L.ref_zoom_out_float(field, out, nx, ny, factor, cnum)
    return out.reshape(nyy, nxx) if cnum == 0 else out


def ref_zoom_out(field, factor):
    """the reference's CPU pyramid stage in double (oct_zoom_out, src/oct_zoom.cc:17): blur + bicubic sampling"""
    field = np.ascontiguousarray(field, np.float64)
    ny, nx = field.shape
    nxx, nyy = zoom_out_size(nx, ny, factor)
    out = np.zeros((nyy, nxx), np.float64)
    ref_cpu().ref_zoom_out(field, out, nx, ny, float(factor))
    return out


def ref_zoom_in(field, nxx, nyy):
    """the reference's CPU prolongation in double (oct_zoom_in, src/oct_zoom.cc:154)"""
    field = np.ascontiguousarray(field, np.float64)
    ny, nx = field.shape
    out = np.zeros((nyy, nxx), np.float64)
    ref_cpu().ref_zoom_in(field, out, nx, ny, nxx, nyy)
    return out


def srsal(u, v, cth):
    """-srsal bilateral post-smoother restated (src/oct_srsal_cuda.cu:35-71)"""
    u = np.array(u, np.float32, order="C"); v = np.array(v, np.float32, order="C")
    cth = np.ascontiguousarray(cth, np.float32)
    ny, nx = u.shape
    L = lib()
    L.oracle_srsal.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int]
    assert L.oracle_srsal(u, v, cth, nx, ny) == 0
    return u, v


def ref_srsal(u, v, cth, L=None):
    """the reference's oct_srsal_cu (needs a GPU)"""
    L = L or ref_cuda()
    u = np.array(u, np.float32, order="C"); v = np.array(v, np.float32, order="C")
    cth = np.ascontiguousarray(cth, np.float32)
    ny, nx = u.shape
    L.ref_srsal.argtypes = [_f32, _f32, _f32, C.c_int, C.c_int, C.POINTER(RefParams)]
    rp = ref_params()
    L.ref_srsal(u, v, cth, nx, ny, C.byref(rp))
    return u, v


# ---- the reference itself -------------------------------------------------
_ref_cpu = None
_ref_cuda = None


def ref_cpu():
    global _ref_cpu
    if _ref_cpu is None:
        L = C.CDLL(os.path.join(HERE, "_ref", "libref_cpu.so"))
        L.ref_patch_match.argtypes = [_f32, _f32, _f32, _f32, C.c_int, C.c_int, C.POINTER(RefParams)]
        L.ref_zoom_out.argtypes = [_f64, _f64, C.c_int, C.c_int, C.c_double]
        L.ref_zoom_in.argtypes = [_f64, _f64, C.c_int, C.c_int, C.c_int, C.c_int]
        _ref_cpu = L
    return _ref_cpu


def ref_cuda():
    global _ref_cuda
    if _ref_cuda is None:
        L = C.CDLL(os.path.join(HERE, "_ref", "libref_cuda.so"))
        L.ref_variational.argtypes = [_f32, _f32, C.c_int, C.c_int, C.c_int, _f32, _f32, C.POINTER(RefParams)]
        L.ref_pix2uv.argtypes = [C.POINTER(RefNav), C.c_double, C.c_double, _f32, _f32, C.c_int, C.c_int,
                                 C.POINTER(RefParams), _i16, _i16, _i16, _i16, C.POINTER(C.c_float)]
        L.ref_optical_flow.argtypes = [_f32, _f32, C.c_void_p, C.c_int, C.c_int, C.POINTER(RefNav),
                                       C.c_double, C.c_double, C.POINTER(RefParams),
                                       _f32, _f32, _i16, _i16, _i16, _i16, C.c_void_p, C.POINTER(C.c_float)]
        L.ref_patch_match.argtypes = [_f32, _f32, _f32, _f32, C.c_int, C.c_int, C.POINTER(RefParams)]
        ingest_argtypes(L)
        _ref_cuda = L
    return _ref_cuda


_ref_shim = None


def ref_shim():
    """The reference's UNMODIFIED dispatcher object (oct_optical_flow.cc) linked against
    octane_b200/shim/oct_b200_shim.cc + liboctane_b200.so: the drop-in proof."""
    global _ref_shim
    if _ref_shim is None:
        L = C.CDLL(os.path.join(HERE, "_ref", "libref_shim.so"))
        L.ref_optical_flow.argtypes = ref_cuda_argtypes()
        ingest_argtypes(L)
        _ref_shim = L
    return _ref_shim


def ref_cuda_argtypes():
    return [_f32, _f32, C.c_void_p, C.c_int, C.c_int, C.POINTER(RefNav), C.c_double, C.c_double,
            C.POINTER(RefParams), _f32, _f32, _i16, _i16, _i16, _i16, C.c_void_p, C.POINTER(C.c_float)]


def ingest_argtypes(L):
    L.ref_navcal.argtypes = [_i16, _i16, _i16, C.c_int, C.c_int, C.POINTER(RefCal), C.POINTER(RefParams), _f32, _f32, _f32]
    L.ref_uv2pix.argtypes = [C.POINTER(RefNav), C.c_double, C.c_double, _f32, _f32, _i16, _i16, C.c_int, C.c_int,
                             C.POINTER(RefParams), _f32, _f32]


def ref_navcal(rad, x, y, cal, L=None):
    """oct_navcal_cuda() of library L (ref_cuda(): the reference; ref_shim(): our shim under the reference's signature)"""
    L = L or ref_cuda()
    rad = np.ascontiguousarray(rad, np.int16); x = np.ascontiguousarray(x, np.int16); y = np.ascontiguousarray(y, np.int16)
    ny, nx = rad.shape
    data = np.zeros((ny, nx), np.float32); lat = np.zeros((ny, nx), np.float32); lon = np.zeros((ny, nx), np.float32)
    rp = ref_params()
    bad = L.ref_navcal(rad, x, y, nx, ny, C.byref(cal), C.byref(rp), data, lat, lon)
    assert bad == 0, "sector copies data2s/xs/ys differ from the inputs"
    return data, lat, lon


def ref_uv2pix(nav, t1, t2, lat, lon, x, y, u, v, L=None):
    L = L or ref_cuda()
    u = np.array(u, np.float32, copy=True); v = np.array(v, np.float32, copy=True)
    ny, nx = u.shape
    rp = ref_params(dofirstguess=1)
    L.ref_uv2pix(C.byref(nav), t1, t2, np.ascontiguousarray(lat, np.float32), np.ascontiguousarray(lon, np.float32),
                 np.ascontiguousarray(x, np.int16), np.ascontiguousarray(y, np.int16), nx, ny, C.byref(rp), u, v)
    return u, v


def ref_dispatch(L, img1, img2, nav, t1, t2, rp=None, cth=None):
    """oct_optical_flow() of library L (ref_cuda() or ref_shim()) on host arrays."""
    rp = rp or ref_params()
    img1 = np.ascontiguousarray(img1, np.float32); img2 = np.ascontiguousarray(img2, np.float32)
    ny, nx = img1.shape
    up = np.zeros((ny, nx), np.float32); vp = np.zeros((ny, nx), np.float32)
    o = [np.zeros((ny, nx), np.int16) for _ in range(4)]
    ctp = np.zeros((ny, nx), np.int16)
    dT = C.c_float()
    cthp = None if cth is None else np.ascontiguousarray(cth, np.float32).ctypes.data
    L.ref_optical_flow(img1, img2, cthp, nx, ny, C.byref(nav), t1, t2, C.byref(rp), up, vp, *o,
                       ctp.ctypes.data, C.byref(dT))
    return dict(uPix=up, vPix=vp, U=o[0], V=o[1], U_raw=o[2], V_raw=o[3], CTP=ctp, dT=dT.value)


def ref_variational(img1, img2, rp=None, u0=None, v0=None, nc=1):
    rp = rp or ref_params()
    img1 = np.ascontiguousarray(img1, np.float32); img2 = np.ascontiguousarray(img2, np.float32)
    ny, nx = img1.shape[-2:]
    u = np.zeros((ny, nx), np.float32) if u0 is None else np.array(u0, np.float32, copy=True)
    v = np.zeros((ny, nx), np.float32) if v0 is None else np.array(v0, np.float32, copy=True)
    ref_cuda().ref_variational(img1, img2, nx, ny, nc, u, v, C.byref(rp))
    return u, v


def ref_pix2uv(nav, t1, t2, u, v, rp=None):
    rp = rp or ref_params()
    u = np.ascontiguousarray(u, np.float32); v = np.ascontiguousarray(v, np.float32)
    ny, nx = u.shape
    o = [np.zeros((ny, nx), np.int16) for _ in range(4)]
    dT = C.c_float()
    ref_cuda().ref_pix2uv(C.byref(nav), t1, t2, u, v, nx, ny, C.byref(rp), *o, C.byref(dT))
    return (*o, dT.value)


def ref_patch_match(img1, img2, rp=None):
    rp = rp or ref_params(dososm=1)
    img1 = np.ascontiguousarray(img1, np.float32); img2 = np.ascontiguousarray(img2, np.float32)
    ny, nx = img1.shape
    u = np.zeros((ny, nx), np.float32); v = np.zeros((ny, nx), np.float32)
    ref_cpu().ref_patch_match(img1, img2, u, v, nx, ny, C.byref(rp))
    return u, v
