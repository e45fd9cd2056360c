"""Host mirror of justplotit.mean_regrid (picaso/justplotit.py:31-63) and its planning step.

`mean_regrid(x, y, newx=None, R=None)` has the reference's signature and return value; `y` may be one
spectrum [nwno], a stack [nbatch, nwno], or a `DeviceArray` (then the rebinned stack stays in HBM).
The bin membership - np.digitize plus scipy's "samples on the rightmost edge belong to the last bin"
rule (scipy/stats/_binned_statistic.py:_bin_numbers) - is planned once per (x, bin edges) pair and
cached: the retrieval driver rebins every model spectrum from the same model grid onto the same data
grid (picaso/driver.py:229-232).  Only index bookkeeping happens here; the sums run on the device.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import PB_DEVICE, PB_HOST, addr

__all__ = ["RegridPlan", "mean_regrid", "create_grid", "bin_edges", "plan_ranges"]


def create_grid(min_wavelength, max_wavelength, constant_R):
    """picaso/opacity_factory.py:712-739: wavenumber grid at constant resolving power."""
    spacing = (2. * constant_R + 1.) / (2. * constant_R - 1.)
    npts = np.log(max_wavelength / min_wavelength) / np.log(spacing)
    wsize = int(np.ceil(npts)) + 1
    newwl = np.zeros(wsize)
    newwl[0] = min_wavelength
    for j in range(1, wsize):
        newwl[j] = newwl[j - 1] * spacing
    return 1e4 / newwl[::-1]


def bin_edges(x, newx=None, R=None):
    """the `bins` argument mean_regrid hands to binned_statistic (justplotit.py:51-58)"""
    if newx is None and R is not None:
        return np.asarray(create_grid(1e4 / max(x), 1e4 / min(x), R), dtype=np.float64)
    if newx is not None and R is None:
        newx = np.asarray(newx, dtype=np.float64)
        d = np.diff(newx)
        return np.array([newx[0] - d[0] / 2] + list(newx[0:-1] + d / 2.0) + [newx[-1] + d[-1] / 2])
    raise Exception('Please either enter a newx or a R')


def plan_ranges(x, edges):
    """(start, count) of the contiguous index range of x that falls into each bin, with the bin
    membership of scipy/stats/_binned_statistic.py:_bin_numbers (np.digitize; samples that round to the
    rightmost edge belong to the last bin)."""
    x = np.asarray(x, dtype=np.float64)
    edges = np.asarray(edges, dtype=np.float64)
    if edges.ndim != 1 or edges.size < 2 or np.any(np.diff(edges) < 0):
        raise ValueError("bin edges must be a monotonically increasing 1-D array")  # as scipy
    if not (np.isfinite(x).all()):
        raise ValueError("x contains non-finite values")  # as scipy
    nbins = edges.size - 1
    binnum = np.digitize(x, edges)
    dmin = np.diff(edges).min()
    if dmin == 0:
        raise ValueError('The smallest edge difference is numerically 0.')
    decimal = int(-np.log10(dmin)) + 6
    on_edge = np.where((x >= edges[-1]) & (np.around(x, decimal) == np.around(edges[-1], decimal)))[0]
    binnum[on_edge] -= 1
    inside = (binnum >= 1) & (binnum <= nbins)
    idx = np.nonzero(inside)[0]
    b = binnum[idx] - 1
    count = np.bincount(b, minlength=nbins).astype(np.int32)
    start = np.zeros(nbins, dtype=np.int32)
    if idx.size:
        first = np.full(nbins, np.iinfo(np.int64).max)
        np.minimum.at(first, b, idx)
        last = np.full(nbins, -1)
        np.maximum.at(last, b, idx)
        has = count > 0
        if np.any(last[has] - first[has] + 1 != count[has]):
            raise _lib.PicasoB200Error("mean_regrid: x must be monotonic (every bin a contiguous index range)")
        start[has] = first[has]
    return start, count


class RegridPlan:
    """(start, count) index range of every bin, resident in HBM."""

    def __init__(self, ctx, x, edges):
        x = np.asarray(x, dtype=np.float64)
        edges = np.asarray(edges, dtype=np.float64)
        start, count = plan_ranges(x, edges)
        nbins = edges.size - 1
        self.ctx, self.nbins, self.nwno = ctx, nbins, x.size
        self.edges = edges
        self.centers = (edges[0:-1] + edges[1:]) / 2.0
        self.start, self.count = start, count
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.pb_regrid_plan_create(ctx.h, nbins, addr(start), addr(count), ctypes.byref(h)))
        self.h = h

    def apply(self, y, scale=1.0, out=None):
        """y: numpy [nwno] | [nbatch, nwno] -> numpy; DeviceArray [nbatch, nwno] -> DeviceArray [nbatch, nbins]
        (`out`: an existing DeviceArray to write into)"""
        ctx = self.ctx
        if hasattr(y, "ptr") and hasattr(y, "ctx"):
            from .optics import DeviceArray
            shp = y.shape if len(y.shape) == 2 else (1, y.shape[0])
            if shp[1] != self.nwno:
                raise _lib.PicasoB200Error("mean_regrid: spectrum has %d points, plan %d" % (shp[1], self.nwno))
            if out is None:
                out = DeviceArray(ctx, (shp[0], self.nbins))
            ctx.check(ctx.lib.pb_mean_regrid(ctx.h, self.h, shp[0], self.nwno, self.nwno, y.ptr, float(scale),
                                             out.ptr, PB_DEVICE))
            return out
        y = np.asarray(y, dtype=np.float64)
        one = y.ndim == 1
        y2 = np.ascontiguousarray(y.reshape(-1, y.shape[-1]))
        if y2.shape[1] != self.nwno:
            raise _lib.PicasoB200Error("mean_regrid: spectrum has %d points, plan %d" % (y2.shape[1], self.nwno))
        out = np.empty((y2.shape[0], self.nbins))
        if y2.shape[0]:
            ctx.check(ctx.lib.pb_mean_regrid(ctx.h, self.h, y2.shape[0], self.nwno, self.nwno, addr(y2),
                                             float(scale), addr(out), PB_HOST))
        return out[0] if one else out

    def close(self):
        if getattr(self, "h", None) is not None and self.ctx.h is not None:
            self.ctx.lib.pb_regrid_plan_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_PLANS = {}


def _plan(ctx, x, edges):
    x = np.ascontiguousarray(x, dtype=np.float64)
    key = (id(ctx), x.size, edges.size, hash(x.tobytes()), hash(edges.tobytes()))
    p = _PLANS.get(key)
    if p is None or p.h is None:
        if len(_PLANS) > 16:
            _PLANS.clear()
        p = _PLANS[key] = RegridPlan(ctx, x, edges)
    return p


def mean_regrid(x, y, newx=None, R=None, *, ctx=None):
    """CUDA replacement of justplotit.mean_regrid (picaso/justplotit.py:31-63): returns (newx, y)."""
    ctx = ctx or _lib.default_context()
    edges = bin_edges(x, newx, R)
    p = _plan(ctx, x, edges)
    return p.centers, p.apply(y)
