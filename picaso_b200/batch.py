"""Batched thermal-emission forward model for retrieval loops (BASELINE cfg5).

The reference's retrieval driver (picaso/driver.py:176-245, MODEL) builds one atmosphere per sample,
calls spectrum() -> get_thermal_1d + compress_thermal, scales by 1e-8 (R/d)^2 and rebins onto the data
grid with mean_regrid - one sample at a time.  `thermal_batch` runs the same three steps for a stack of
atmospheres in two launches (the nbatch axis of the thermal TOA kernel with the disk integration fused,
then the rebinning kernel); nothing of size [nbatch, nwno] returns to the host when `newx` is given.
`shard` splits the atmospheres over ranks (one process per GPU, no exchange on the data path).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import PB_DEVICE, ThermalArgs, addr
from .regrid import _plan, bin_edges

__all__ = ["thermal_batch", "shard"]


def shard(nbatch, rank, world):
    """contiguous block of atmospheres owned by `rank` (same convention as sharded.wave_slab)"""
    base, rem = divmod(nbatch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _buf(ctx, name, nbytes):
    return ctx.workspace("batch_" + name, nbytes)


def _cached_vec(ctx, name, v):
    """device copy of a small host vector, re-uploaded only when its contents change"""
    cache = ctx.__dict__.setdefault("_batch_vec", {})
    ent = cache.get(name)
    if ent is None or ent[0].shape != v.shape or not np.array_equal(ent[0], v):
        ptr = _buf(ctx, "vec_" + name, v.nbytes)
        ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, ptr, v.ctypes.data, v.nbytes))
        ctx.sync()
        ent = cache[name] = (v.copy(), ptr)
    return ent[1]


def _dev(ctx, a, shape, name):
    """device pointer of a [nbatch, ...] input: DeviceArray as is, numpy staged through the workspace"""
    if hasattr(a, "ptr") and hasattr(a, "ctx"):
        if tuple(a.shape) != tuple(shape):
            raise _lib.PicasoB200Error("thermal_batch: array of shape %s, expected %s" % (a.shape, shape))
        return a.ptr
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise _lib.PicasoB200Error("thermal_batch: array of shape %s, expected %s" % (a.shape, shape))
    ptr = _buf(ctx, name, a.nbytes)
    ctx.check(ctx.lib.pb_memcpy_h2d(ctx.h, ptr, a.ctypes.data, a.nbytes))
    return ptr


def thermal_batch(wno, tlevel, plevel, dtau, w0, cosb, ubar1, gweight, tweight, surf_reflect=0.0, hard_surface=0,
                  dwno=None, calc_type=0, newx=None, R=None, scale=1.0, *, ctx=None, device_output=False):
    """tlevel, plevel [nbatch, nlevel]; dtau, w0, cosb [nbatch, nlayer, nwno] (numpy or DeviceArray).

    Per atmosphere b: compress_thermal(get_thermal_1d(...)[0]) * scale (fluxes.py:1683-1912, disco.py:152-180,
    driver.py:226), then mean_regrid onto `newx` / constant `R` (justplotit.py:31-63) if one is given.
    Returns (x, y[nbatch, len(x)]): x = wno without rebinning, else the bin centres.  With
    device_output=True y is a DeviceArray view of a workspace that the next call overwrites."""
    ctx = ctx or _lib.default_context()
    from .optics import DeviceArray
    tl = np.ascontiguousarray(tlevel, dtype=np.float64)
    pl = np.ascontiguousarray(plevel, dtype=np.float64)
    B, V = tl.shape
    L = V - 1
    wn = np.ascontiguousarray(wno, dtype=np.float64)
    W = wn.size
    u1 = np.ascontiguousarray(ubar1, dtype=np.float64)
    ng, nt = u1.shape
    gw = np.ascontiguousarray(gweight, dtype=np.float64)
    tw = np.ascontiguousarray(tweight, dtype=np.float64)
    a = ThermalArgs()
    a.nlayer, a.nwno, a.numg, a.numt, a.nbatch, a.ld = L, W, ng, nt, B, W
    a.dtau, a.w0, a.cosb = (_dev(ctx, x, (B, L, W), n) for x, n in ((dtau, "dtau"), (w0, "w0"), (cosb, "cosb")))
    a.wno = _cached_vec(ctx, "wno", wn)
    if calc_type == 1:
        a.dwno = _cached_vec(ctx, "dwno", np.ascontiguousarray(np.broadcast_to(np.asarray(dwno, dtype=np.float64), (W,))))
    sr = np.asarray(surf_reflect, dtype=np.float64)
    if sr.ndim == 0 and float(sr) == 0.0:
        a.surf_reflect = None  # NULL = 0 (include/picaso_b200.h)
    else:
        a.surf_reflect = _dev(ctx, np.broadcast_to(sr, (B, W)) if sr.ndim < 2 else sr, (B, W), "surf")
    a.tlevel, a.plevel, a.ubar1, a.gweight, a.tweight = addr(tl), addr(pl), addr(u1.reshape(-1)), addr(gw), addr(tw)
    a.hard_surface, a.calc_type = int(hard_surface), int(calc_type)
    spec = DeviceArray(ctx, (B, W), ptr=_buf(ctx, "spec", B * W * 8))
    a.thermal = spec.ptr
    if B and W:
        ctx.check(ctx.lib.pb_thermal_toon_1d(ctx.h, ctypes.byref(a), PB_DEVICE))
    if newx is None and R is None:
        if scale != 1.0:
            raise _lib.PicasoB200Error("thermal_batch: scale is applied by the rebinning kernel; pass newx or R")
        x, out = wn, spec
    else:
        plan = _plan(ctx, wn, bin_edges(wn, newx, R))
        out = DeviceArray(ctx, (B, plan.nbins), ptr=_buf(ctx, "rebinned", B * plan.nbins * 8))
        x, out = plan.centers, plan.apply(spec, scale=scale, out=out)
    if device_output:
        return x, out
    return x, out.numpy()
