"""Synthetic GOES-R-shaped classic NetCDF files for the `octane` CLI tests (scipy writes CDF-1/CDF-2;
the image has no netCDF4/HDF5).  Variable / attribute names are the ones oct_goesread reads
(reference src/oct_fileread.cc:99-263)."""
import numpy as np
from scipy.io import netcdf_file


def write_goes_l1b(path, rad, x, y, t, band, xScale, xOffset, yScale, yOffset, radScale, radOffset,
                   lon0=-75.0, req=6378137.0, rpol=6356752.31414, pph=35786023.0):
    ny, nx = rad.shape
    f = netcdf_file(path, "w", version=2)
    f.createDimension("y", ny); f.createDimension("x", nx); f.createDimension("num_bands", 1)
    v = f.createVariable("Rad", "h", ("y", "x")); v[:] = rad
    v.scale_factor = np.float32(radScale); v.add_offset = np.float32(radOffset)
    v = f.createVariable("x", "h", ("x",)); v[:] = x
    v.scale_factor = np.float32(xScale); v.add_offset = np.float32(xOffset)
    v = f.createVariable("y", "h", ("y",)); v[:] = y
    v.scale_factor = np.float32(yScale); v.add_offset = np.float32(yOffset)
    v = f.createVariable("t", "d", ()); v.data[...] = t; v.units = "seconds since 2000-01-01 12:00:00"
    v = f.createVariable("band_id", "b", ("num_bands",)); v[:] = np.array([band], np.int8)
    v = f.createVariable("goes_imager_projection", "i", ()); v.data[...] = -2147483647
    v.longitude_of_projection_origin = np.float64(lon0); v.semi_major_axis = np.float64(req)
    v.semi_minor_axis = np.float64(rpol); v.inverse_flattening = np.float64(298.2572221)
    v.latitude_of_projection_origin = np.float64(0.0); v.perspective_point_height = np.float64(pph)
    for name, val in (("planck_fk1", 202263.0), ("planck_fk2", 3698.19), ("planck_bc1", 0.43361),
                      ("planck_bc2", 0.99939), ("kappa0", 0.0019486)):
        v = f.createVariable(name, "f", ()); v.data[...] = np.float32(val)
    f.close()


def write_plane_file(path, **planes):
    """CLAVR-x / first-guess shaped file: dims ny, nx; float variables"""
    first = next(iter(planes.values()))
    ny, nx = first.shape
    f = netcdf_file(path, "w", version=2)
    f.createDimension("ny", ny); f.createDimension("nx", nx)
    for name, a in planes.items():
        v = f.createVariable(name, "f", ("ny", "nx")); v[:] = a.astype(np.float32)
    f.close()


def counts_from_image(img, maxin, minin, radScale, radOffset):
    """inverse of the ingest normalisation: brightness 0..255 -> radiance -> short counts"""
    radv = img.astype(np.float64) / 255.0 * (maxin - minin) + minin
    return np.clip(np.rint((radv - radOffset) / radScale), -32768, 32767).astype(np.int16)


def write_grid_file(path, rad, x, y, t, xScale, xOffset, yScale, yOffset, R, lat1=None, lon0=None, lon1=None):
    """-Polar / -Merc shaped file (reference oct_polarread / oct_mercread, src/oct_fileread.cc:418-752):
    float Rad, grid constants on the variable grid_mapping"""
    ny, nx = rad.shape
    f = netcdf_file(path, "w", version=2)
    f.createDimension("y", ny); f.createDimension("x", nx)
    v = f.createVariable("Rad", "f", ("y", "x")); v[:] = rad.astype(np.float32)
    v = f.createVariable("x", "h", ("x",)); v[:] = x
    v.scale_factor = np.float32(xScale); v.add_offset = np.float32(xOffset)
    v = f.createVariable("y", "h", ("y",)); v[:] = y
    v.scale_factor = np.float32(yScale); v.add_offset = np.float32(yOffset)
    v = f.createVariable("t", "d", ()); v.data[...] = t; v.units = "seconds since 2000-01-01 12:00:00"
    v = f.createVariable("grid_mapping", "i", ()); v.data[...] = 7
    v.R = np.float32(R)
    if lat1 is not None:
        v.lat1 = np.float32(lat1); v.lon0 = np.float32(lon0)
    if lon1 is not None:
        v.lon1 = np.float32(lon1)
    f.close()
