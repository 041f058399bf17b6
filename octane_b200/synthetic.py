"""Seeded synthetic GOES-R-shaped inputs.

Cloud texture = band-limited power-law random field (spectrum amplitude
~ k^-1.7 * exp(-(k/0.25)^2), uniform random phases), scaled to the 0..255
brightness range the reference's ingest emits (src/oct_navcal_cuda.cu:93).
Frame 2 is frame 1 displaced by a known flow (backward cubic warp), so
I1(x) = I2(x + w(x)) to interpolation accuracy.

numpy/scipy path for test sizes; a torch path (periodic 4096^2 tile, device
side) for the CONUS / full-disk bench sizes.  Data generation is not part of
the hot path and is never timed.
"""
from __future__ import annotations

import numpy as np


def texture(nx: int, ny: int, seed: int) -> np.ndarray:
    """ny x nx float32 field in [0, 255]."""
    rng = np.random.default_rng(seed)
    ky = np.fft.fftfreq(ny)[:, None]
    kx = np.fft.rfftfreq(nx)[None, :]
    k = np.sqrt(kx * kx + ky * ky)
    amp = np.zeros_like(k)
    nz = k > 0
    amp[nz] = k[nz] ** -1.7 * np.exp(-((k[nz] / 0.25) ** 2))
    phase = rng.uniform(0.0, 2.0 * np.pi, size=k.shape)
    f = np.fft.irfft2(amp * np.exp(1j * phase), s=(ny, nx))
    f -= f.min()
    f *= 255.0 / f.max()
    return f.astype(np.float32)


def flow_field(nx: int, ny: int, kind: str = "vortex", drift=(0.8, -0.4), peak: float = 2.0):
    """Known displacement (u, v) in pixels, float64 ny x nx."""
    y, x = np.mgrid[0:ny, 0:nx].astype(np.float64)
    if kind == "shift":
        return np.full((ny, nx), drift[0]), np.full((ny, nx), drift[1])
    cx, cy, rad = 0.5 * (nx - 1), 0.5 * (ny - 1), nx / 4.0
    dx, dy = (x - cx) / rad, (y - cy) / rad
    r2 = dx * dx + dy * dy
    # Gaussian vortex whose speed peaks at `peak` px (at r = rad/sqrt(2))
    s = peak * np.sqrt(2.0 * np.e) * np.exp(-r2)
    return drift[0] - s * dy, drift[1] + s * dx


def warp_pair(img1: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Frame 2 = frame 1 sampled at (x - u, y - v), cubic spline, edge-clamped."""
    from scipy.ndimage import map_coordinates

    ny, nx = img1.shape
    y, x = np.mgrid[0:ny, 0:nx].astype(np.float64)
    out = map_coordinates(img1.astype(np.float64), [y - v, x - u], order=3, mode="nearest")
    return np.clip(out, 0.0, 255.0).astype(np.float32)


def make_pair(nx: int, ny: int, seed: int, kind: str = "vortex", drift=(0.8, -0.4), peak: float = 2.0):
    """(img1, img2, u_true, v_true), all ny x nx."""
    img1 = texture(nx, ny, seed)
    u, v = flow_field(nx, ny, kind, drift, peak)
    return img1, warp_pair(img1, u, v), u.astype(np.float32), v.astype(np.float32)


# GOES-R ABI fixed-grid constants (public values; any self-consistent set works)
GOES_NAV = dict(pph=35786023.0, req=6378137.0, rpol=6356752.31414, lon0_deg=-75.0)
SECTORS = {
    # name: (xScale, yScale, xOffset, yOffset, dt seconds)
    "fulldisk_0.5km": (1.4e-5, -1.4e-5, -0.151865, 0.151865, 600.0),
    "conus_0.5km": (1.4e-5, -1.4e-5, -0.101353, 0.128233, 300.0),
    "meso_0.5km": (1.4e-5, -1.4e-5, -0.030000, 0.100000, 60.0),
    "meso_2km": (5.6e-5, -5.6e-5, -0.030000, 0.100000, 60.0),
}


def make_pair_torch(nx: int, ny: int, seed: int, device, limb_taper: bool = False,
                    drift=(0.8, -0.4), peak: float = 2.0, tile: int = 4096, rows=None, out=None):
    """Large-scene generator on `device` (torch): periodic `tile`^2 texture
    replicated over the scene with a slow brightness modulation, frame 2 =
    frame 1 sampled at (x-u, y-v) (bilinear on the periodic tile, exact
    wrap-around).  `rows=(r0, r1)` generates only that row band of the scene
    (identical values to the same rows of the full scene), `out=(a, b)` writes
    into existing (pinned host or device) tensors.  Returns two float32 tensors
    of shape (r1-r0) x nx."""
    import torch

    t = min(tile, 1 << int(np.ceil(np.log2(max(nx, ny)))))
    base = torch.from_numpy(texture(t, t, seed)).to(device)

    def sample(xs, ys):
        x0 = torch.floor(xs); y0 = torch.floor(ys)
        fx = xs - x0; fy = ys - y0
        x0 = x0.long() % t; y0 = y0.long() % t
        x1 = (x0 + 1) % t; y1 = (y0 + 1) % t
        return ((1 - fy) * ((1 - fx) * base[y0, x0] + fx * base[y0, x1])
                + fy * ((1 - fx) * base[y1, x0] + fx * base[y1, x1]))

    r0, r1 = rows if rows is not None else (0, ny)
    if out is not None:
        img1, img2 = out
    else:
        img1 = torch.empty((r1 - r0, nx), dtype=torch.float32, device=device)
        img2 = torch.empty((r1 - r0, nx), dtype=torch.float32, device=device)
    cx, cy, rad = 0.5 * (nx - 1), 0.5 * (ny - 1), nx / 4.0
    xs_full = torch.arange(nx, device=device, dtype=torch.float32)[None, :]
    chunk = max(1, (1 << 24) // nx)
    for j0 in range(r0, r1, chunk):
        j1 = min(r1, j0 + chunk)
        ys = torch.arange(j0, j1, device=device, dtype=torch.float32)[:, None]
        xs = xs_full.expand(j1 - j0, nx)
        ysb = ys.expand(j1 - j0, nx)
        dx, dy = (xs - cx) / rad, (ysb - cy) / rad
        s = peak * float(np.sqrt(2.0 * np.e)) * torch.exp(-(dx * dx + dy * dy))
        u = drift[0] - s * dy
        v = drift[1] + s * dx
        mod = 0.75 + 0.25 * torch.sin(xs * (2 * np.pi / (3.7 * t))) * torch.cos(ysb * (2 * np.pi / (2.9 * t)))
        a = sample(xs, ysb) * mod
        xb, yb = xs - u, ysb - v
        modb = 0.75 + 0.25 * torch.sin(xb * (2 * np.pi / (3.7 * t))) * torch.cos(yb * (2 * np.pi / (2.9 * t)))
        b = sample(xb, yb) * modb
        if limb_taper:  # src/oct_navcal_cuda.cu:81-91: 1 inside x^2+y^2<0.021, linear to 0 at 0.0212
            xsc, ysc, xo, yo, _ = SECTORS["fulldisk_0.5km"]
            r2 = (xs * xsc + xo) ** 2 + (ysb * ysc + yo) ** 2
            w = torch.clamp((0.0212 - r2) / (0.0212 - 0.021), 0.0, 1.0)
            a = a * w; b = b * w
        img1[j0 - r0:j1 - r0].copy_(a)
        img2[j0 - r0:j1 - r0].copy_(b)
    return img1, img2
