"""Host-side mirror of picaso/disco.py: geometry tables stay on the host (O(ng) work),
the disk integrations run on the GPU through the C ABI."""
import numpy as np

from . import _lib
from ._lib import PB_HOST, addr
from .synth import _GAUSS

__all__ = ["get_angles_1d", "get_angles_3d", "compute_disco", "compress_disco", "compress_thermal"]


def get_angles_1d(ngauss):
    """disco.get_angles_1d (disco.py:52-89): half-sphere Gauss points, ngauss in 5..8."""
    if ngauss not in _GAUSS:
        raise Exception("Please enter ngauss=5,6,7 or 8.")
    g, w = _GAUSS[ngauss]
    return np.array(g), np.array(w), np.array([0]), np.array([1])


def get_angles_3d(num_gangle, num_tangle):
    """disco.get_angles_3d (disco.py:92-115): Gauss-Legendre x Chebyshev."""
    i = np.linspace(1, num_tangle, num_tangle)
    tangle = np.cos(i * np.pi / (num_tangle + 1))
    tweight = np.pi / (num_tangle + 1) * np.sin(i * np.pi / (num_tangle + 1)) ** 2.0
    gangle, gweight = np.polynomial.legendre.leggauss(num_gangle)
    return gangle, gweight, tangle, tweight


def compute_disco(ng, nt, gangle, tangle, phase_angle):
    """disco.compute_disco (disco.py:8-50): ubar0, ubar1, cos_theta, latitude, longitude."""
    gangle = np.asarray(gangle, dtype=np.float64)
    tangle = np.asarray(tangle, dtype=np.float64)
    cos_theta = np.cos(phase_angle)
    lon = np.arcsin((gangle - (cos_theta - 1.0) / (cos_theta + 1.0)) / (2.0 / (cos_theta + 1)))
    if phase_angle > np.pi:
        lon = -lon
    colat = np.arccos(tangle)
    lat = np.pi / 2 - colat
    f = np.sin(colat)
    ubar0 = np.outer(np.cos(lon - phase_angle), f)
    ubar1 = np.outer(np.cos(lon), f)
    return ubar0, ubar1, cos_theta, lat, lon


def compress_disco(nwno, cos_theta, xint_at_top, gweight, tweight, F0PI, *, ctx=None):
    """CUDA replacement of disco.compress_disco (disco.py:118-149)."""
    ctx = ctx or _lib.default_context()
    gw = np.ascontiguousarray(gweight, dtype=np.float64)
    tw = np.ascontiguousarray(tweight, dtype=np.float64)
    x = np.ascontiguousarray(xint_at_top, dtype=np.float64).reshape(gw.size * tw.size, nwno)
    f0 = np.asarray(F0PI, dtype=np.float64)
    f0 = np.full(nwno, float(f0)) if f0.ndim == 0 else np.ascontiguousarray(f0)
    out = np.zeros(nwno)
    if nwno > 0:
        ctx.check(ctx.lib.pb_compress_disco(ctx.h, nwno, float(cos_theta), addr(x), addr(gw), gw.size,
                                            addr(tw), tw.size, addr(f0), addr(out), PB_HOST))
    return out


def compress_thermal(nwno, flux_at_top, gweight, tweight, *, ctx=None):
    """CUDA replacement of disco.compress_thermal (disco.py:152-180); accepts
    [ng,nt,nwno] or [ng,nt,nlevel,nwno] like the reference."""
    ctx = ctx or _lib.default_context()
    gw = np.ascontiguousarray(gweight, dtype=np.float64)
    tw = np.ascontiguousarray(tweight, dtype=np.float64)
    x = np.ascontiguousarray(flux_at_top, dtype=np.float64)
    tail = x.shape[2:]
    n = int(np.prod(tail))
    x = x.reshape(gw.size * tw.size, n)
    out = np.zeros(n)
    if n > 0:
        ctx.check(ctx.lib.pb_compress_thermal(ctx.h, n, addr(x), addr(gw), gw.size, addr(tw), tw.size,
                                              addr(out), PB_HOST))
    return out.reshape(tail)
